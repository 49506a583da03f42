## Package hooks of the GPU-backed build: free device workspaces, streams, plans and NCCL communicators when the
## package is unloaded (spw_shutdown), then drop the shared object.
.onUnload <- function(libpath) {
  try(.Call(spw_r_shutdown), silent = TRUE)
  library.dynam.unload("spWombling", libpath)
}
