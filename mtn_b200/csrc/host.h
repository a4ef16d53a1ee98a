// Host-side helpers shared by the C-ABI translation units: error reporting and
// TMA tensor-map construction (driver entry point resolved at run time, so the
// library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mtn_b200.h"

namespace mtn {

int set_error(int code, const char* fmt, ...);
const char* last_error();

#define MTN_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::mtn::set_error(MTN_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define MTN_REQUIRE(cond, code, ...)                               \
  do {                                                             \
    if (!(cond)) return ::mtn::set_error(code, __VA_ARGS__);       \
  } while (0)

// Decoding-step programs (csrc/decode_rows.cu): while one is being recorded the few-row entry points and
// mtn_layernorm_fwd append stages instead of launching; every other launch is an error (it would run out of order).
bool prog_recording();
int prog_push_layernorm(const float* x, const float* a2, const float* b2, float eps, int rows, int d, int rows_per_group,
                        float* y32, void* y16);

// Launch with the programmatic-stream-serialization attribute (PDL) unless MTN_B200_PDL=0.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, unsigned cluster_x, Args&&... args) {
  if (prog_recording()) return cudaErrorNotPermitted;   // this kernel cannot be a stage of a decoding-step program
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  return launch_kernel_cluster(kernel, grid, block, smem, st, 1u, static_cast<Args&&>(args)...);
}

// SMs a kernel launched on `st` can use: the SM count of the stream's green context (CUDA SM partitioning: the engine
// may run its side chains on a small partition, mtn_b200/parallel.py) or of the device.  Persistent kernels size their
// grids with it.  Cached per stream handle.
int stream_sm_count(cudaStream_t st);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

enum TmSwizzle { TM_SWZ_64 = 64, TM_SWZ_128 = 128 };

// f16 tensor map, rank 2: dims {cols, rows}, row pitch ld (elements), box {box_cols, box_rows}.
int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld,
                     uint32_t box_cols, uint32_t box_rows, TmSwizzle swz);
// rank 3: dims {cols, rows, batch}, strides {ld, rows_stride (elements)}; box {box_cols, box_rows, 1}.
int make_tmap_3d_f16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint64_t ld, uint64_t batch_stride, uint32_t box_cols, uint32_t box_rows,
                     TmSwizzle swz);

// f32, rank 3 (the residual stream as the target of TMA reduce-add stores): strides in elements
int make_tmap_3d_f32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint64_t ld, uint64_t batch_stride, uint32_t box_cols, uint32_t box_rows,
                     TmSwizzle swz);

}  // namespace mtn
