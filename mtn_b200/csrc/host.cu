#include "host.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include <vector>

namespace mtn {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTN_B200_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode(CUtensorMap* out, const void* base, uint32_t rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box, TmSwizzle swz,
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  EncodeTiledFn fn = encode_fn();
  MTN_REQUIRE(fn != nullptr, MTN_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  MTN_REQUIRE(aligned16(base), MTN_E_ALIGN, "TMA base pointer %p is not 16-byte aligned", base);
  for (uint32_t i = 0; i + 1 < rank; ++i)
    MTN_REQUIRE(strides_bytes[i] % 16 == 0, MTN_E_ALIGN, "TMA stride %llu B is not a multiple of 16",
                (unsigned long long)strides_bytes[i]);
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype, rank, const_cast<void*>(base), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz == TM_SWZ_128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTN_REQUIRE(r == CUDA_SUCCESS, MTN_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MTN_OK;
}

typedef CUresult (*StreamGetGreenCtxFn)(CUstream, CUgreenCtx*);
typedef CUresult (*GreenCtxGetDevResourceFn)(CUgreenCtx, CUdevResource*, CUdevResourceType);

int stream_sm_count(cudaStream_t st) {
  static int dev_sms = 0;
  static StreamGetGreenCtxFn get_ctx = nullptr;
  static GreenCtxGetDevResourceFn get_res = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamGetGreenCtx", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      get_ctx = reinterpret_cast<StreamGetGreenCtxFn>(p);
    if (cudaGetDriverEntryPoint("cuGreenCtxGetDevResource", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      get_res = reinterpret_cast<GreenCtxGetDevResourceFn>(p);
  });
  if (st == nullptr || get_ctx == nullptr || get_res == nullptr) return dev_sms;
  static std::mutex mu;
  static std::vector<std::pair<cudaStream_t, int>> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : cache)
    if (e.first == st) return e.second;
  int n = dev_sms;
  CUgreenCtx g = nullptr;
  if (get_ctx(reinterpret_cast<CUstream>(st), &g) == CUDA_SUCCESS && g != nullptr) {
    CUdevResource res;
    if (get_res(g, &res, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS && res.sm.smCount > 0) n = (int)res.sm.smCount;
  }
  if (cache.size() < 64) cache.emplace_back(st, n);
  return n;
}

int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld,
                     uint32_t box_cols, uint32_t box_rows, TmSwizzle swz) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(out, base, 2, dims, strides, box, swz);
}

int make_tmap_3d_f16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint64_t ld, uint64_t batch_stride, uint32_t box_cols, uint32_t box_rows,
                     TmSwizzle swz) {
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * 2, batch_stride * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  return encode(out, base, 3, dims, strides, box, swz);
}

int make_tmap_3d_f32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint64_t ld, uint64_t batch_stride, uint32_t box_cols, uint32_t box_rows,
                     TmSwizzle swz) {
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld * 4, batch_stride * 4};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  return encode(out, base, 3, dims, strides, box, swz, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

}  // namespace mtn

extern "C" int mtn_abi_version(void) { return MTN_B200_ABI_VERSION; }
extern "C" const char* mtn_last_error(void) { return mtn::last_error(); }
