// Row-wise HBM-bound kernels of the MTN hot path: the reference's custom LayerNorm,
// f32 -> f16 operand packing and mask bit-packing.  All are one pass over the data
// with 128-bit coalesced accesses; the roofline that bounds them is HBM bandwidth
// (algorithmic bytes: LN reads 4 B/elem and writes 2 (f16) and/or 4 (f32) B/elem).
#include "common.cuh"
#include "host.h"

namespace mtn {

// ----------------------------------------------------------------------------
// LayerNorm (mtn.py:111-114): y = a*(x-mean)/(std_unbiased + eps) + b.
// One warp per row; the row lives in registers (VPL float4 per lane), two-pass
// mean / sum of squared deviations exactly like the reference (no E[x^2]-mean^2
// cancellation).
// ----------------------------------------------------------------------------
template <int VPL>  // float4 per lane: d = 128 * VPL
__global__ void __launch_bounds__(256)
    layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ a2,
                          const float* __restrict__ b2, float eps, int rows, float* __restrict__ y32,
                          __half* __restrict__ y16) {
  constexpr int D = 128 * VPL;
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / (sqrtf(ss * (1.f / (D - 1))) + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    const float4 a = __ldg(reinterpret_cast<const float4*>(a2) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(b2) + c4);
    float4 o;
    o.x = a.x * v[i].x * inv + b.x;
    o.y = a.y * v[i].y * inv + b.y;
    o.z = a.z * v[i].z * inv + b.z;
    o.w = a.w * v[i].w * inv + b.w;
    if (y32) reinterpret_cast<float4*>(y32 + (size_t)row * D)[c4] = o;
    if (y16)
      reinterpret_cast<uint2*>(y16 + (size_t)row * D)[c4] =
          make_uint2(pack_f16x2_sat(o.x, o.y), pack_f16x2_sat(o.z, o.w));
  }
}

// any d (used for odd sizes such as the d=4 known-answer test): one warp per row,
// three passes through L1/L2.
__global__ void layernorm_generic_kernel(const float* __restrict__ x, const float* __restrict__ a2,
                                         const float* __restrict__ b2, float eps, int rows, int d,
                                         float* __restrict__ y32, __half* __restrict__ y16) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += xr[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float t = xr[i] - mean;
    ss += t * t;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / (sqrtf(ss / (d - 1)) + eps);
  for (int i = lane; i < d; i += 32) {
    const float o = a2[i] * (xr[i] - mean) * inv + b2[i];
    if (y32) y32[(size_t)row * d + i] = o;
    if (y16) {
      const uint32_t p = pack_f16x2_sat(o, 0.f);
      y16[(size_t)row * d + i] = __ushort_as_half((unsigned short)(p & 0xffff));
    }
  }
}

// ----------------------------------------------------------------------------
// f32 -> f16 (rn, saturating), 8 elements per thread when rows are 16-B aligned.
// ----------------------------------------------------------------------------
__global__ void cast_f16_vec8_kernel(const float* __restrict__ src, int ld_src, __half* __restrict__ dst,
                                     int ld_dst, int rows, int cols8) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols8) return;
  const int r = (int)(i / cols8), c = (int)(i % cols8) * 8;
  const float4 lo = *reinterpret_cast<const float4*>(src + (size_t)r * ld_src + c);
  const float4 hi = *reinterpret_cast<const float4*>(src + (size_t)r * ld_src + c + 4);
  *reinterpret_cast<uint4*>(dst + (size_t)r * ld_dst + c) =
      make_uint4(pack_f16x2_sat(lo.x, lo.y), pack_f16x2_sat(lo.z, lo.w), pack_f16x2_sat(hi.x, hi.y),
                 pack_f16x2_sat(hi.z, hi.w));
}
__global__ void cast_f16_scalar_kernel(const float* __restrict__ src, int ld_src, __half* __restrict__ dst,
                                       int ld_dst, int rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const uint32_t p = pack_f16x2_sat(src[(size_t)r * ld_src + c], 0.f);
  dst[(size_t)r * ld_dst + c] = __ushort_as_half((unsigned short)(p & 0xffff));
}

// ----------------------------------------------------------------------------
// mask bytes -> bit words.  One warp per 32 keys: ballot.
// ----------------------------------------------------------------------------
__global__ void mask_pack_kernel(const uint8_t* __restrict__ m, int nrows, int Lk, int words,
                                 uint32_t* __restrict__ bits) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (size_t)nrows * words) return;
  const int row = (int)(w / words), word = (int)(w % words);
  const int k = word * 32 + (threadIdx.x & 31);
  const bool keep = (k < Lk) && (m[(size_t)row * Lk + k] != 0);
  const uint32_t b = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) bits[w] = b;
}

}  // namespace mtn

extern "C" int mtn_layernorm_fwd(const float* x, const float* a_2, const float* b_2, float eps, int rows,
                                 int d, float* y_f32, void* y_f16, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(x && a_2 && b_2 && (y_f32 || y_f16), MTN_E_ARG, "layernorm: NULL pointer");
  MTN_REQUIRE(rows > 0 && d > 1, MTN_E_SHAPE, "layernorm: rows=%d d=%d", rows, d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* y16 = reinterpret_cast<__half*>(y_f16);
  const int wpb = 8;  // warps (= rows) per block
  dim3 grid((rows + wpb - 1) / wpb);
  const bool vec = (d % 128 == 0) && d <= 1024 && aligned16(x) && aligned16(a_2) && aligned16(b_2) &&
                   (!y_f32 || aligned16(y_f32)) && (!y_f16 || aligned16(y_f16));
  if (vec && d == 128) MTN_CHECK_CUDA(launch_kernel(layernorm_rows_kernel<1>, grid, dim3(32 * wpb), 0, st, x, a_2, b_2, eps, rows, y_f32, y16));
  else if (vec && d == 256) MTN_CHECK_CUDA(launch_kernel(layernorm_rows_kernel<2>, grid, dim3(32 * wpb), 0, st, x, a_2, b_2, eps, rows, y_f32, y16));
  else if (vec && d == 512) MTN_CHECK_CUDA(launch_kernel(layernorm_rows_kernel<4>, grid, dim3(32 * wpb), 0, st, x, a_2, b_2, eps, rows, y_f32, y16));
  else if (vec && d == 1024) MTN_CHECK_CUDA(launch_kernel(layernorm_rows_kernel<8>, grid, dim3(32 * wpb), 0, st, x, a_2, b_2, eps, rows, y_f32, y16));
  else MTN_CHECK_CUDA(launch_kernel(layernorm_generic_kernel, grid, dim3(32 * wpb), 0, st, x, a_2, b_2, eps, rows, d, y_f32, y16));
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}

extern "C" int mtn_cast_f32_to_f16(const float* src, int ld_src, void* dst, int ld_dst, int rows, int cols,
                                   void* stream) {
  using namespace mtn;
  MTN_REQUIRE(src && dst, MTN_E_ARG, "cast: NULL pointer");
  MTN_REQUIRE(rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, MTN_E_SHAPE,
              "cast: rows=%d cols=%d ld_src=%d ld_dst=%d", rows, cols, ld_src, ld_dst);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* d16 = reinterpret_cast<__half*>(dst);
  if (cols % 8 == 0 && ld_src % 4 == 0 && ld_dst % 8 == 0 && aligned16(src) && aligned16(dst)) {
    const size_t n = (size_t)rows * (cols / 8);
    MTN_CHECK_CUDA(launch_kernel(cast_f16_vec8_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, src, ld_src, d16,
                                 ld_dst, rows, cols / 8));
  } else {
    const size_t n = (size_t)rows * cols;
    MTN_CHECK_CUDA(launch_kernel(cast_f16_scalar_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, src, ld_src,
                                 d16, ld_dst, rows, cols));
  }
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}

extern "C" int mtn_mask_words(int Lk) { return ((Lk + 127) / 128) * 4; }  // padded to whole 128-key tiles

extern "C" int mtn_mask_pack(const uint8_t* mask_u8, int B, int rows_q, int Lk, uint32_t* bits, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(mask_u8 && bits, MTN_E_ARG, "mask_pack: NULL pointer");
  MTN_REQUIRE(B > 0 && rows_q > 0 && Lk > 0, MTN_E_SHAPE, "mask_pack: B=%d rows_q=%d Lk=%d", B, rows_q, Lk);
  const int words = mtn_mask_words(Lk);
  const size_t nw = (size_t)B * rows_q * words;
  MTN_CHECK_CUDA(launch_kernel(mask_pack_kernel, dim3((unsigned)((nw + 7) / 8)), dim3(256), 0,
                               static_cast<cudaStream_t>(stream), mask_u8, B * rows_q, Lk, words, bits));
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}
