// Host side of tma.cuh: cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked).
#include <mutex>
#include <utility>
#include <vector>
#include "tma.cuh"

namespace spw {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SPW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled is not available in this driver");
      return SPW_ERR_CUDA;
    }
    fn = (EncodeTiledFn)p;
  }
  *out = fn;
  return SPW_OK;
}

// Encoded maps are cached: the workspaces they describe are long-lived and the same (pointer, shape) recurs for
// every launch of a factorisation, so the driver call is paid once per distinct operand.
struct MapKey {
  const void* p;
  long long ld, stride;
  int cols, rows, batch, box_cols, box_rows;
  bool operator==(const MapKey& o) const {
    return p == o.p && ld == o.ld && stride == o.stride && cols == o.cols && rows == o.rows && batch == o.batch &&
           box_cols == o.box_cols && box_rows == o.box_rows;
  }
};
static std::vector<std::pair<MapKey, CUtensorMap>> g_map_cache;
static std::mutex g_map_mu;

void tensor_map_cache_clear() {
  std::lock_guard<std::mutex> lk(g_map_mu);
  g_map_cache.clear();
}

int make_tensor_map(CUtensorMap* tm, const MatView& M, int cols, int rows, int batch, int box_cols, int box_rows) {
  const MapKey key{M.p, M.ld, M.stride, cols, rows, batch, box_cols, box_rows};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    for (auto& kv : g_map_cache)
      if (kv.first == key) {
        *tm = kv.second;
        return SPW_OK;
      }
  }
  EncodeTiledFn enc;
  SPW_TRY(get_encode_fn(&enc));
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  // the batch stride of a single matrix is irrelevant but must still be a valid (16-byte multiple, >= row extent) value
  const cuuint64_t bstride = (M.stride > 0 ? (cuuint64_t)M.stride : (cuuint64_t)M.ld * rows) * sizeof(double);
  cuuint64_t strides[2] = {(cuuint64_t)M.ld * sizeof(double), bstride};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)M.p, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): cols=%d rows=%d batch=%d ld=%lld", (int)r, cols, rows, batch, M.ld);
    return SPW_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_map_cache.size() >= 64) g_map_cache.clear();
  g_map_cache.emplace_back(key, *tm);
  return SPW_OK;
}

}  // namespace spw
