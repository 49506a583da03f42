"""Device-resident entry points (torch tensors as HBM buffers, torch streams, torch.distributed).

PyTorch is plumbing here -- device memory, the current CUDA stream and the NCCL process group -- while every
kernel that runs is libspwombling's own (reached through the ``*_dev`` functions of the C ABI).

  * ``gradient_shard``        one rank's contiguous sample range of the spatial_gradient loop
                              (R/spatial_gradient.R:199-338; the reference's mclapply chunk, :29-32,200)
  * ``spatial_gradient_sharded``  shard by sample over the ranks of a process group, ONE gather of the draws
                              to rank 0 over NCCL / NVLink (the reference's rbind of the children's lists,
                              R/spatial_gradient.R:378-390), batch-median / HPD summary on rank 0 (:391-419)
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, load
from .api import COMP_NAMES, EST_NAMES, _code, ncomp_of


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64(t, device):
    return torch.as_tensor(t, dtype=torch.float64, device=device).contiguous()


def coords_to_device(coords, device):
    """N x 2 host array -> device tensor [2][N] (the column-major layout the C ABI expects)."""
    c = np.ascontiguousarray(np.asarray(coords, dtype=np.float64).T)
    return torch.from_numpy(c).to(device)


def gradient_shard(coords_d, grid_d, cov_type, phi_d, sig2_d, z_d, eps_d, want_moments=False, out=None):
    """Run samples of this rank.  coords_d [2][N], grid_d [2][G], phi_d / sig2_d [S], z_d [S][ldz >= N],
    eps_d [S][G][c]  ->  draws [S][c][G] (and mean [S][G][c], var [S][G][c][c] when ``want_moments``)."""
    lib = load()
    code = _code(cov_type)
    c = ncomp_of(cov_type)
    n, g, S = coords_d.shape[1], grid_d.shape[1], phi_d.shape[0]
    dev = coords_d.device
    draws = out if out is not None else torch.empty((S, c, g), dtype=torch.float64, device=dev)
    mean = torch.empty((S, g, c), dtype=torch.float64, device=dev) if want_moments else None
    var = torch.empty((S, g, c, c), dtype=torch.float64, device=dev) if want_moments else None
    info = (ctypes.c_int * 4)()
    with torch.cuda.device(dev):
        rc = lib.spw_gradient_shard_dev(code, n, _p(coords_d), g, _p(grid_d), S, _p(phi_d), _p(sig2_d), _p(z_d),
                                        int(z_d.stride(0)), _p(eps_d), _p(draws), _p(mean), _p(var),
                                        ctypes.cast(info, ctypes.c_void_p), _stream_ptr())
    check(rc, list(info))
    if want_moments:
        return draws, mean, var.transpose(2, 3)
    return draws


def hpd_summary_dev(draws_d, nbatch, prob=0.95):
    """draws [S][c][G] (device) -> [c][3][G] (estimate, lower, upper)."""
    S, c, g = draws_d.shape
    out = torch.empty((c, 3, g), dtype=torch.float64, device=draws_d.device)
    with torch.cuda.device(draws_d.device):
        check(load().spw_hpd_summary_dev(S, g, c, int(nbatch), float(prob), _p(draws_d), _p(out), _stream_ptr()))
    return out


def draws_to_r_layout(draws_d):
    """[S][c][G] -> [c][G][S]: the c matrices S x G in R's column-major order."""
    S, c, g = draws_d.shape
    out = torch.empty((c, g, S), dtype=torch.float64, device=draws_d.device)
    with torch.cuda.device(draws_d.device):
        check(load().spw_draws_to_r_layout_dev(S, g, c, _p(draws_d), _p(out), _stream_ptr()))
    return out


def shard_bounds(S, world, rank):
    """Contiguous sample range of ``rank``.  Any contiguous split gives bit-identical draws (samples are
    independent); this is the even split the reference's ``split(1:S, ceiling(...))`` aims for."""
    return (S * rank) // world, (S * (rank + 1)) // world


def gather_shards(mine, S, group=None):
    """The path's one exchange step: every rank contributes its [S_r][c][G] block, rank 0 receives the
    sample-major concatenation [S][c][G] (other ranks get None).  Equal shards use one ``gather``
    collective; ragged shards use point-to-point sends into rank 0's slices."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_bounds(S, world, r)[1] - shard_bounds(S, world, r)[0] for r in range(world)]
    assert mine.shape[0] == counts[rank]
    even = len(set(counts)) == 1
    if rank == 0:
        full = torch.empty((S,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        parts = list(torch.split(full, counts, dim=0))
        if even:
            dist.gather(mine, gather_list=parts, dst=0, group=group)
        else:
            parts[0].copy_(mine)
            reqs = [dist.irecv(parts[r], src=r, group=group) for r in range(1, world) if counts[r] > 0]
            for q in reqs:
                q.wait()
        return full
    if even:
        dist.gather(mine, gather_list=None, dst=0, group=group)
    elif counts[rank] > 0:
        dist.send(mine, dst=0, group=group)
    return None


def spatial_gradient_sharded(coords, model, cov_type, grid_points, nbatch=200, return_mcmc=True, eps=None,
                             group=None, device=None, rounded=True, prob=0.95,
                             compute=None, summarise=None, to_r_layout=None, local=None):
    """``spatial_gradient`` over the ranks of a torch.distributed process group (one process per GPU).

    Every rank passes the full host inputs and works on its ``shard_bounds`` range; rank 0 returns the
    reference's list as a dict, the other ranks return None.  Alternatively ``local`` hands over this rank's shard
    directly: a dict of (pinned) host tensors ``phi, sig2, z, eps`` covering exactly ``shard_bounds(S_total, ...)``
    plus ``S_total`` (``model`` / ``eps`` are then ignored).  ``compute`` / ``summarise`` / ``to_r_layout``
    default to the CUDA entry points (they exist as arguments so that the host logic -- sharding, gather,
    assembly -- can be exercised on CPU with the gloo backend in the tests).
    """
    import torch.distributed as dist
    compute = gradient_shard if compute is None else compute
    summarise = hpd_summary_dev if summarise is None else summarise
    to_r_layout = draws_to_r_layout if to_r_layout is None else to_r_layout
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    c = ncomp_of(cov_type)
    coords_d = coords_to_device(coords, device)
    grid_d = coords_to_device(grid_points, device)
    n, g = coords_d.shape[1], grid_d.shape[1]
    def up(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device, non_blocking=True)

    if local is not None:
        S = int(local["S_total"])
        lo, hi = shard_bounds(S, world, rank)
        phi_d, sig2_d, z_d, eps_d = up(local["phi"]), up(local["sig2"]), up(local["z"]), up(local["eps"])
        if phi_d.shape[0] != hi - lo:
            raise ValueError("local shard must hold exactly the samples of shard_bounds(S_total, world, rank)")
    else:
        phi = np.asarray(model["phi"], dtype=np.float64).reshape(-1)
        S = phi.shape[0]
        lo, hi = shard_bounds(S, world, rank)
        sl = slice(lo, hi)
        phi_d = up(phi[sl])
        sig2_d = up(np.asarray(model["sig2"], dtype=np.float64).reshape(-1)[sl])
        z_d = up(np.asarray(model["z"], dtype=np.float64).reshape(S, n)[sl])
        eps_d = up(np.asarray(eps, dtype=np.float64).reshape(S, g, c)[sl])
    mine = compute(coords_d, grid_d, cov_type, phi_d, sig2_d, z_d, eps_d)
    draws_all = gather_shards(mine, S, group) if multi else mine
    if rank != 0:
        return None
    summ = summarise(draws_all, nbatch, prob).cpu().numpy()
    digits = 2 if cov_type == "matern1" else 4
    out = {}
    for k in range(c):
        est = np.ascontiguousarray(summ[k].T)
        out[EST_NAMES[c][k]] = np.round(est, digits) if rounded else est
    if return_mcmc:
        r = to_r_layout(draws_all).cpu().numpy()
        for k in range(c):
            out["grad.%s.mcmc" % COMP_NAMES[c][k]] = r[k].T
    return out
