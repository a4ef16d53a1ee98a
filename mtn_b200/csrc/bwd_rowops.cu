// Row-wise HBM-bound kernels of the BACKWARD pass of the MTN hot path (training, train.py:33-39):
// the reference's custom LayerNorm (mtn.py:111-114), embedding scatter, bias gradients (column sums fused
// with the f32 -> f16 operand cast of the incoming gradient), the per-row dot products the attention
// backward needs, the generator's log-softmax and the label-smoothed KL criterion.  One pass over the data
// each, 128-bit accesses; parameter gradients are accumulated with f32 L2 reductions.
//
// Gradient scaling.  Tensor-core operands of the backward GEMMs are f16 like the forward ones; so that
// small gradients do not underflow, a backward pass multiplies the incoming gradient by a power of two S
// chosen on the device from its magnitude (mtn_grad_scale_*), keeps every intermediate scaled, and multiplies
// each RESULT (parameter / input gradient) by 1/S in the producing kernel.  `scale` / `alpha` arguments below
// are pointers to those device scalars (NULL = 1).
//
// Input pointers are `const T*` WITHOUT __restrict__, and device scalars are read with __ldcg: nearly every input
// here is produced by the kernel launched just before, whose writes overlap this kernel's lifetime under
// programmatic dependent launch, so the non-coherent load path is off limits (rule in common.cuh).  Only weights
// (embedding table, positional table, a_2) keep __ldg.
#include <math_constants.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// ----------------------------------------------------------------------------
// LayerNorm backward.  y = a (x - mean) / (sigma + eps) + b, sigma = unbiased std.  With c = x - mean,
// s = sigma + eps, g = dy * a:
//   dx_i = (g_i - mean(g)) / s  -  c_i * (sum_j g_j c_j) / (s^2 (D-1) sigma)
//   da  += sum_rows dy * c / s        db += sum_rows dy
// One warp per row, grid-stride over rows with the parameter-gradient partials in registers; one smem
// reduction + one set of atomics per block.  EMBED: the LN input is recomputed as lut[id]*scale + pe[pos]
// (mtn_embed_fwd) and dx is scattered into the embedding-table gradient instead of stored.
// ----------------------------------------------------------------------------
struct LnBwdParams {
  const float* x;
  const long long* ids; const float* lut; const float* pe; int L; int vocab; float emb_scale;
  const float* a2;
  const float* dy;
  const float* dy_scale;
  const float* param_alpha;
  const float* dres;
  float* dx;
  float* dlut;
  float* da; float* db;
  float eps; int rows; int has_ln;
  __half* dx16;   // optional f16 copy of dx: the operand of the next backward GEMM
  float* colsum;  // optional: += param_alpha * sum_rows dx  (bias gradient of the layer whose output this is)
  int multimem;   // da / db / colsum are NVLS multicast addresses (data-parallel gradient reduction in the switch)
  DropCfg drop;   // !EMBED: dropout of the sublayer output that dx16 / colsum flow into (mtn.py:127), applied to
                  // them only.  EMBED: dropout of the embedding (mtn.py:309), applied before the scatter.
};

template <int VPL, bool EMBED>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const LnBwdParams p) {
  constexpr int D = 128 * VPL;
  __shared__ float red[8 * D];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float dys = p.dy_scale ? __ldcg(p.dy_scale) : 1.f;
  const float pal = p.param_alpha ? __ldcg(p.param_alpha) : 1.f;
  const unsigned long long dseed = p.drop.seed ? __ldcg(p.drop.seed) : 0ull;
  // keep decisions of this lane's float4 number i of `row` (columns 4*(lane+32i) .. +3): 4 bits
  auto keep4 = [&](int row, int i) -> uint32_t {
    if (p.drop.seed == nullptr) return 0xfu;
    const int c4 = lane + 32 * i;
    return (drop_keep8(p.drop, dseed, (unsigned long long)row * (D / 8) + (c4 >> 1)) >> (4 * (c4 & 1))) & 0xfu;
  };
  float4 da_acc[VPL], db_acc[VPL], cs_acc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) da_acc[i] = db_acc[i] = cs_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int row = blockIdx.x * 8 + warp; row < p.rows; row += gridDim.x * 8) {
    float4 c[VPL], g[VPL];
    long long id = 0;
    float s1 = 0.f;
    if (EMBED) {
      id = p.ids[row];
      id = id < 0 ? 0 : (id >= p.vocab ? p.vocab - 1 : id);
      const float4* er = reinterpret_cast<const float4*>(p.lut + (size_t)id * D);
      const float4* pr = reinterpret_cast<const float4*>(p.pe + (size_t)(row % p.L) * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 e = __ldg(er + lane + 32 * i), q = __ldg(pr + lane + 32 * i);
        c[i] = make_float4(e.x * p.emb_scale + q.x, e.y * p.emb_scale + q.y, e.z * p.emb_scale + q.z,
                           e.w * p.emb_scale + q.w);
        const uint32_t kb = keep4(row, i);
        const float ik = p.drop.inv_keep;
        c[i] = make_float4((kb & 1u) ? c[i].x * ik : 0.f, (kb & 2u) ? c[i].y * ik : 0.f, (kb & 4u) ? c[i].z * ik : 0.f,
                           (kb & 8u) ? c[i].w * ik : 0.f);
      }
    } else {
      const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)row * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) c[i] = xr[lane + 32 * i];
    }
    const float4* dyr = reinterpret_cast<const float4*>(p.dy + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 t = dyr[lane + 32 * i];
      g[i] = make_float4(t.x * dys, t.y * dys, t.z * dys, t.w * dys);  // g holds dy for now
      s1 += (c[i].x + c[i].y) + (c[i].z + c[i].w);
    }
    float4 dxv[VPL];
    if (p.has_ln) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      const float mean = s1 * (1.f / D);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        c[i].x -= mean; c[i].y -= mean; c[i].z -= mean; c[i].w -= mean;
        ss += (c[i].x * c[i].x + c[i].y * c[i].y) + (c[i].z * c[i].z + c[i].w * c[i].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float sigma = sqrtf(ss * (1.f / (D - 1)));
      const float inv = 1.f / (sigma + p.eps);
      float sg = 0.f, sgc = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.a2) + lane + 32 * i);
        const float4 dyv = g[i];
        // parameter gradients use dy itself
        da_acc[i].x += dyv.x * c[i].x * inv; da_acc[i].y += dyv.y * c[i].y * inv;
        da_acc[i].z += dyv.z * c[i].z * inv; da_acc[i].w += dyv.w * c[i].w * inv;
        db_acc[i].x += dyv.x; db_acc[i].y += dyv.y; db_acc[i].z += dyv.z; db_acc[i].w += dyv.w;
        g[i] = make_float4(dyv.x * a.x, dyv.y * a.y, dyv.z * a.z, dyv.w * a.w);
        sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        sgc += (g[i].x * c[i].x + g[i].y * c[i].y) + (g[i].z * c[i].z + g[i].w * c[i].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sgc += __shfl_xor_sync(0xffffffffu, sgc, o);
      }
      const float mg = sg * (1.f / D);
      const float coef = sigma > 0.f ? sgc * inv * inv / ((float)(D - 1) * sigma) : 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        dxv[i] = make_float4((g[i].x - mg) * inv - c[i].x * coef, (g[i].y - mg) * inv - c[i].y * coef,
                             (g[i].z - mg) * inv - c[i].z * coef, (g[i].w - mg) * inv - c[i].w * coef);
    } else {
#pragma unroll
      for (int i = 0; i < VPL; ++i) dxv[i] = g[i];
    }
    if (EMBED) {
      float* gr = p.dlut + (size_t)id * D;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const uint32_t kb = keep4(row, i);
        const float f = p.emb_scale * pal * p.drop.inv_keep;   // inv_keep == 1 without dropout
        red_add_v4(gr + 4 * (lane + 32 * i), make_float4((kb & 1u) ? dxv[i].x * f : 0.f, (kb & 2u) ? dxv[i].y * f : 0.f,
                                                         (kb & 4u) ? dxv[i].z * f : 0.f, (kb & 8u) ? dxv[i].w * f : 0.f));
      }
    } else {
      float4* dxr = reinterpret_cast<float4*>(p.dx + (size_t)row * D);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float4 o = dxv[i];
        if (p.dres != nullptr) {
          const float4 r = reinterpret_cast<const float4*>(p.dres + (size_t)row * D)[lane + 32 * i];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        dxr[lane + 32 * i] = o;
        if (p.dx16 != nullptr || p.colsum != nullptr) {
          const uint32_t kb = keep4(row, i);
          const float ik = p.drop.inv_keep;
          o = make_float4((kb & 1u) ? o.x * ik : 0.f, (kb & 2u) ? o.y * ik : 0.f, (kb & 4u) ? o.z * ik : 0.f,
                          (kb & 8u) ? o.w * ik : 0.f);
        }
        if (p.dx16 != nullptr)
          reinterpret_cast<uint2*>(p.dx16 + (size_t)row * D)[lane + 32 * i] =
              make_uint2(pack_f16x2_sat(o.x, o.y), pack_f16x2_sat(o.z, o.w));
        cs_acc[i].x += o.x; cs_acc[i].y += o.y; cs_acc[i].z += o.z; cs_acc[i].w += o.w;
      }
    }
  }
  // block reduction of the parameter-gradient partials (8 warps), then one atomic per column
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? p.da : (pass == 1 ? p.db : p.colsum);
    if (dst == nullptr || (pass < 2 && !p.has_ln)) continue;  // uniform over the block
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      reinterpret_cast<float4*>(red + warp * D)[lane + 32 * i] = pass == 0 ? da_acc[i] : (pass == 1 ? db_acc[i] : cs_acc[i]);
    __syncthreads();
    for (int col = threadIdx.x; col < D; col += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w * D + col];
      grad_red_f32(dst + col, t * pal, p.multimem);
    }
  }
}

// any d (odd sizes such as the d=4 known-answer test): one warp per row, atomics per element.
__global__ void layernorm_bwd_generic_kernel(const LnBwdParams p, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int lane = threadIdx.x & 31;
  const float dys = p.dy_scale ? p.dy_scale[0] : 1.f, pal = p.param_alpha ? p.param_alpha[0] : 1.f;
  const float* xr = p.x + (size_t)row * d;
  const float* dyr = p.dy + (size_t)row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += xr[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / d;
  float ss = 0.f, sg = 0.f, sgc = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float c = xr[i] - mean, g = dyr[i] * dys * p.a2[i];
    ss += c * c; sg += g; sgc += g * c;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sgc += __shfl_xor_sync(0xffffffffu, sgc, o);
  }
  const float sigma = sqrtf(ss / (d - 1)), inv = 1.f / (sigma + p.eps);
  const float mg = sg / d, coef = sigma > 0.f ? sgc * inv * inv / ((float)(d - 1) * sigma) : 0.f;
  for (int i = lane; i < d; i += 32) {
    const float c = xr[i] - mean, dyv = dyr[i] * dys;
    float o = (dyv * p.a2[i] - mg) * inv - c * coef;
    if (p.dres) o += p.dres[(size_t)row * d + i];
    p.dx[(size_t)row * d + i] = o;
    if (p.da) {
      grad_red_f32(p.da + i, dyv * c * inv * pal, p.multimem);
      grad_red_f32(p.db + i, dyv * pal, p.multimem);
    }
  }
}

// ----------------------------------------------------------------------------
// Operand cast + bias gradient:  dst16 = f16(src * scale [masked by relu_mask > 0]),
// colsum[c] += alpha * sum_rows (src * scale [masked]).  TIn = float or __half; dst16 / colsum optional.
// Block (32 x 8): 32 eight-column vectors x 8 row lanes, ROWS_PER_BLOCK rows per block.
// ----------------------------------------------------------------------------
constexpr int CC_ROWS_PER_BLOCK = 64;

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(256)
    cast_colsum_kernel(const TIn* src, int ld_src, __half* __restrict__ dst, int ld_dst,
                       const __half* relu_mask, int ld_mask, int rows, int vcols, int rows_per_block,
                       const float* scale, const float* alpha, float* __restrict__ colsum,
                       const DropCfg drop, int multimem) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long dseed = drop.seed ? __ldcg(drop.seed) : 0ull;
  const int vc = blockIdx.x * 32 + threadIdx.x;
  const bool active = vc < vcols;
  const float sc = scale ? __ldcg(scale) : 1.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int r_end = active ? min(rows, (int)(blockIdx.y + 1) * rows_per_block) : 0;
  const int r0 = blockIdx.y * rows_per_block + threadIdx.y;
  constexpr int U = 4;  // rows per thread and batch: all loads of a batch are issued before any use
  float v[U][8], m[U][8];
#pragma unroll 1
  for (int rb = r0; rb < r_end; rb += 8 * U) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int r = rb + 8 * u;
    if (r < r_end) {
      load8(src + (size_t)r * ld_src + 8 * vc, v[u]);
      if (relu_mask != nullptr) load8(relu_mask + (size_t)r * ld_mask + 8 * vc, m[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int r = rb + 8 * u;
    if (r >= r_end) continue;
    const uint32_t kb = drop.seed ? drop_keep8(drop, dseed, (unsigned long long)r * vcols + vc) : 0xffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = v[u][j];
      if (relu_mask != nullptr) x = m[u][j] > 0.f ? x : 0.f;
      x = ((kb >> j) & 1u) ? x * drop.inv_keep : 0.f;
      x *= sc;
      v[u][j] = x;
      acc[j] += x;
    }
    if (dst != nullptr)
      *reinterpret_cast<uint4*>(dst + (size_t)r * ld_dst + 8 * vc) =
          make_uint4(pack_f16x2_sat(v[u][0], v[u][1]), pack_f16x2_sat(v[u][2], v[u][3]), pack_f16x2_sat(v[u][4], v[u][5]),
                     pack_f16x2_sat(v[u][6], v[u][7]));
  }
  }
  if (colsum != nullptr) {
    const float al = alpha ? __ldcg(alpha) : 1.f;
    // the 8 row lanes of a block share columns: combine them through shared memory first
    __shared__ float sh[8][32][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x][j];
        grad_red_f32(colsum + 8 * vc + j, t * al, multimem);
      }
    }
  }
}

// ----------------------------------------------------------------------------
// Gradient scale: S = 2^k with absmax * S in [2^7, 2^8); out = {S, 1/S}.  absmax is collected into a
// self-resetting slot (uint bits of a non-negative float order like the float).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* x, size_t n4, unsigned* __restrict__ slot) {
  __shared__ float sh[8];
  pdl_launch_dependents();
  pdl_wait();
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]);
    if (!(m <= 3.0e38f)) m = 0.f;  // NaN / inf: ignore (scale falls back to 1)
    atomicMax(slot, __float_as_uint(m));
  }
}
__global__ void grad_scale_kernel(unsigned* __restrict__ slot, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const float m = __uint_as_float(*slot);
  *slot = 0u;
  float S = 1.f;
  if (m > 0.f) {
    int e;
    frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(8 - e) in [128, 256)
    int k = 8 - e;
    k = k > 60 ? 60 : (k < -60 ? -60 : k);
    S = ldexpf(1.f, k);
  }
  out[0] = S;
  out[1] = 1.f / S;
}

// ----------------------------------------------------------------------------
// Fused Adam over a flat parameter arena (torch.optim.Adam semantics, train.py:190-191):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// one pass: also refreshes the f16 tensor-core copy of the parameters and zeroes the gradient.  Hyper-parameters
// live in device memory (state[8] = lr, b1, b2, eps, bc1, bc2, step, -) so that a captured step can advance them:
// adam_advance_kernel increments the step, updates the bias corrections and, for the reference's NoamOpt schedule
// (data_utils.py:92-117), the learning rate.
// ----------------------------------------------------------------------------
__global__ void adam_advance_kernel(float* st, float noam_factor, float model_size, float warmup) {
  pdl_launch_dependents();
  pdl_wait();
  const float step = st[6] + 1.f;
  st[6] = step;
  st[4] = 1.f - powf(st[1], step);
  st[5] = 1.f - powf(st[2], step);
  if (noam_factor > 0.f) st[0] = noam_factor * (rsqrtf(model_size) * fminf(rsqrtf(step), step * powf(warmup, -1.5f)));
}

__global__ void __launch_bounds__(256)
    adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     __half* __restrict__ p16, size_t n4, const float* st, int zero_grad) {
  pdl_launch_dependents();
  pdl_wait();
  const float lr = __ldcg(st), b1 = __ldcg(st + 1), b2 = __ldcg(st + 2), eps = __ldcg(st + 3);
  const float step_size = lr / __ldcg(st + 4), inv_sqrt_bc2 = rsqrtf(__ldcg(st + 5));
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 Gv = reinterpret_cast<const float4*>(g)[i];
    float4 Mv = reinterpret_cast<float4*>(m)[i], Vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &P.x; const float* gg = &Gv.x; float* mm = &Mv.x; float* vv = &Vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = b1 * mm[j] + (1.f - b1) * gg[j];
      vv[j] = b2 * vv[j] + (1.f - b2) * gg[j] * gg[j];
      pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = Mv;
    reinterpret_cast<float4*>(v)[i] = Vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p16 != nullptr) reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_f16x2_sat(P.x, P.y), pack_f16x2_sat(P.z, P.w));
  }
}

__global__ void seed_bump_kernel(unsigned long long* seed) {
  pdl_launch_dependents();
  pdl_wait();
  *seed += 1ull;
}

// y = (accumulate ? y : 0) + x * alpha: un-scaling of an input gradient that leaves the backward pass
__global__ void __launch_bounds__(256) scale_f32_kernel(const float* x, const float* alpha,
                                                        float* __restrict__ y, size_t n4, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const float a = alpha ? __ldcg(alpha) : 1.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v = make_float4(v.x * a, v.y * a, v.z * a, v.w * a);
    if (accumulate) {
      const float4 o = reinterpret_cast<const float4*>(y)[i];
      v = make_float4(v.x + o.x, v.y + o.y, v.z + o.z, v.w + o.w);
    }
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

// ----------------------------------------------------------------------------
// Per-(row, head) dot product  delta[b, h, q] = sum_c dO[row, h*dk + c] * O[row, h*dk + c]  (f16 operands,
// f32 result): the softmax-backward row term of attention (dS = P * (dP - delta)).  One warp per row.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    attn_delta_kernel(const __half* dO, int lddo, const __half* O, int ldo, int B, int Lq, int h,
                      int dk, float* __restrict__ delta) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B * Lq) return;
  const int lane = threadIdx.x & 31;
  const int vph = dk / 8;  // 8-element vectors per head (4 or 8)
  const int b = row / Lq, q = row % Lq;
  for (int v = lane; v < h * vph; v += 32) {
    float a[8], o[8];
    load8(dO + (size_t)row * lddo + 8 * v, a);
    load8(O + (size_t)row * ldo + 8 * v, o);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(a[j], o[j], s);
    for (int off = vph >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((lane % vph) == 0) delta[((size_t)b * h + v / vph) * Lq + q] = s;
  }
}

// ----------------------------------------------------------------------------
// Generator tail backward (mtn.py:68-69): y = log_softmax(z)  =>  dz = dy - exp(y) * sum_v dy.
// One 128-thread block per row; columns [V, ld_dz) of dz are zeroed (vocabulary padded to x8).
// ----------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_128(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return (sh[0] + sh[1]) + (sh[2] + sh[3]);
}
__device__ __forceinline__ float block_max_128(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
}

__global__ void __launch_bounds__(128)
    log_softmax_bwd_kernel(const float* y, int ldy, const float* dy, int lddy, int V,
                           float* __restrict__ dz, int lddz) {
  __shared__ float sh[4];
  pdl_launch_dependents();
  pdl_wait();
  const float* yr = y + (size_t)blockIdx.x * ldy;
  const float* dyr = dy + (size_t)blockIdx.x * lddy;
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 128) s += dyr[i];
  const float tot = block_sum_128(s, sh);
  float* dzr = dz + (size_t)blockIdx.x * lddz;
  for (int i = threadIdx.x; i < lddz; i += 128) dzr[i] = i < V ? dyr[i] - __expf(yr[i]) * tot : 0.f;
}

// ----------------------------------------------------------------------------
// Label-smoothed KL criterion backward (label_smoothing.py:20-32 + KLDivLoss(sum)), from logits OR
// log-probabilities z (same formula as the forward kernel in rowops.cu):
//   dz_v = g * (T_r softmax(z)_v - t_v),  t = smoothed target row, T_r = sum_v t_v.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    label_smoothing_bwd_kernel(const float* z, int ldz, int V, const long long* tgt, long long pad,
                               float smoothing, const unsigned long long* pad_index_sum, float gscale,
                               const float* gout, float* __restrict__ dz, int lddz) {
  __shared__ float sh[4];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const float* zr = z + (size_t)r * ldz;
  float* dzr = dz + (size_t)r * lddz;
  const long long y = tgt[r];
  const float g = gscale * (gout ? __ldcg(gout) : 1.f);
  const float s = smoothing / (float)(V - 2), conf = 1.f - smoothing;
  const bool is_pad = (y == pad);
  if (is_pad && *pad_index_sum > 0ull) {  // zeroed padding row
    for (int i = threadIdx.x; i < lddz; i += 128) dzr[i] = 0.f;
    return;
  }
  float mx = -3.4e38f;
  for (int i = threadIdx.x; i < V; i += 128) mx = fmaxf(mx, zr[i]);
  const float m = block_max_128(mx, sh);
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += 128) se += __expf(zr[i] - m);
  const float inv = 1.f / block_sum_128(se, sh);
  const float T = is_pad ? s * (float)(V - 1) : conf + s * (float)(V - 2);
  for (int i = threadIdx.x; i < lddz; i += 128) {
    float o = 0.f;
    if (i < V) {
      const float t = (i == pad) ? 0.f : ((!is_pad && i == y) ? conf : s);
      o = g * (T * __expf(zr[i] - m) * inv - t);
    }
    dzr[i] = o;
  }
}

// sum of the indices of the padding rows (label_smoothing.py:26-30 quirk; same as the forward's kernel)
__global__ void __launch_bounds__(1024) pad_index_sum_bwd_kernel(const long long* tgt, int rows, long long pad,
                                                                 unsigned long long* __restrict__ out) {
  __shared__ unsigned long long sh[32];
  pdl_launch_dependents();
  pdl_wait();
  unsigned long long acc = 0;
  for (int r = threadIdx.x; r < rows; r += 1024) acc += (tgt[r] == pad) ? (unsigned long long)r : 0ull;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 32; ++w) t += sh[w];
    *out = t;
  }
}

}  // namespace mtn

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int mtn_layernorm_bwd(const MtnLayerNormBwdArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->x && a->a_2 && a->dy && a->dx, MTN_E_ARG, "layernorm_bwd: NULL pointer");
  MTN_REQUIRE((a->da_2 == nullptr) == (a->db_2 == nullptr), MTN_E_ARG, "layernorm_bwd: da_2 and db_2 go together");
  MTN_REQUIRE(a->rows > 0 && a->d > 1, MTN_E_SHAPE, "layernorm_bwd: rows=%d d=%d", a->rows, a->d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LnBwdParams p = {};
  p.x = a->x; p.a2 = a->a_2; p.dy = a->dy; p.dy_scale = a->dy_scale; p.param_alpha = a->param_alpha;
  p.dres = a->dres; p.dx = a->dx; p.da = a->da_2; p.db = a->db_2; p.eps = a->eps; p.rows = a->rows; p.has_ln = 1;
  p.dx16 = reinterpret_cast<__half*>(a->dx_f16); p.colsum = a->dx_colsum; p.multimem = a->multimem;
  MTN_REQUIRE(a->drop_thresh < 65536u, MTN_E_ARG, "layernorm_bwd: drop_thresh=%u", a->drop_thresh);
  p.drop = DropCfg{reinterpret_cast<const unsigned long long*>(a->drop_seed), a->drop_site, a->drop_thresh,
                   a->drop_seed ? 1.f / (1.f - a->drop_thresh / 65536.f) : 1.f};
  const int d = a->d;
  const bool vec = (d == 128 || d == 256 || d == 512 || d == 1024) && aligned16(a->x) && aligned16(a->a_2) &&
                   aligned16(a->dy) && aligned16(a->dx) && (!a->dres || aligned16(a->dres));
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MTN_CHECK_CUDA(cudaGetDevice(&dev));
    MTN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int blocks = (a->rows + 7) / 8;
  if (blocks > 4 * sms) blocks = 4 * sms;
  if (vec && d == 128) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<1, false>, dim3(blocks), dim3(256), 0, st, p));
  else if (vec && d == 256) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<2, false>, dim3(blocks), dim3(256), 0, st, p));
  else if (vec && d == 512) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<4, false>, dim3(blocks), dim3(256), 0, st, p));
  else if (vec && d == 1024) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<8, false>, dim3(blocks), dim3(256), 0, st, p));
  else {
    MTN_REQUIRE(a->dx_f16 == nullptr && a->dx_colsum == nullptr && a->drop_seed == nullptr, MTN_E_SHAPE,
                "layernorm_bwd: dx_f16 / dx_colsum need d in {128, 256, 512, 1024} and 16-byte aligned pointers");
    MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_generic_kernel, dim3((a->rows + 7) / 8), dim3(256), 0, st, p, d));
  }
  return MTN_OK;
}

extern "C" int mtn_embed_bwd(const MtnEmbedBwdArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->ids && a->lut && a->pe && a->dy && a->dlut, MTN_E_ARG, "embed_bwd: NULL pointer");
  MTN_REQUIRE((a->a_2 == nullptr) == (a->da_2 == nullptr) && (a->da_2 == nullptr) == (a->db_2 == nullptr), MTN_E_ARG,
              "embed_bwd: a_2, da_2 and db_2 go together");
  MTN_REQUIRE(a->rows > 0 && a->L > 0 && a->vocab > 0, MTN_E_SHAPE, "embed_bwd: rows=%d L=%d vocab=%d", a->rows, a->L,
              a->vocab);
  const int d = a->d;
  MTN_REQUIRE(d == 128 || d == 256 || d == 512 || d == 1024, MTN_E_SHAPE, "embed_bwd: d=%d (supported: 128, 256, 512, 1024)", d);
  MTN_REQUIRE(aligned16(a->lut) && aligned16(a->pe) && aligned16(a->dy) && aligned16(a->dlut), MTN_E_ALIGN,
              "embed_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LnBwdParams p = {};
  p.ids = reinterpret_cast<const long long*>(a->ids); p.lut = a->lut; p.pe = a->pe; p.L = a->L; p.vocab = a->vocab;
  p.emb_scale = a->scale; p.a2 = a->a_2; p.dy = a->dy; p.dy_scale = nullptr; p.param_alpha = a->param_alpha;
  p.dlut = a->dlut; p.da = a->da_2; p.db = a->db_2; p.eps = a->eps; p.rows = a->rows; p.has_ln = a->a_2 != nullptr;
  MTN_REQUIRE(a->drop_thresh < 65536u, MTN_E_ARG, "embed_bwd: drop_thresh=%u", a->drop_thresh);
  p.drop = DropCfg{reinterpret_cast<const unsigned long long*>(a->drop_seed), a->drop_site, a->drop_thresh,
                   a->drop_seed ? 1.f / (1.f - a->drop_thresh / 65536.f) : 1.f};
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MTN_CHECK_CUDA(cudaGetDevice(&dev));
    MTN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int blocks = (a->rows + 7) / 8;
  if (blocks > 4 * sms) blocks = 4 * sms;
  if (d == 128) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<1, true>, dim3(blocks), dim3(256), 0, st, p));
  else if (d == 256) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<2, true>, dim3(blocks), dim3(256), 0, st, p));
  else if (d == 512) MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<4, true>, dim3(blocks), dim3(256), 0, st, p));
  else MTN_CHECK_CUDA(launch_kernel(layernorm_bwd_kernel<8, true>, dim3(blocks), dim3(256), 0, st, p));
  return MTN_OK;
}

extern "C" int mtn_cast_colsum(const void* src, int src_is_f16, int ld_src, void* dst_f16, int ld_dst,
                               const void* relu_mask, int ld_mask, int rows, int cols, const float* scale,
                               const float* alpha, float* colsum, const void* drop_seed, uint32_t drop_site,
                               uint32_t drop_thresh, int multimem, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(drop_thresh < 65536u, MTN_E_ARG, "cast_colsum: drop_thresh=%u", drop_thresh);
  const DropCfg drop{reinterpret_cast<const unsigned long long*>(drop_seed), drop_site, drop_thresh,
                     drop_seed ? 1.f / (1.f - drop_thresh / 65536.f) : 1.f};
  MTN_REQUIRE(src && (dst_f16 || colsum), MTN_E_ARG, "cast_colsum: NULL pointer");
  MTN_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ld_src >= cols && ld_src % 8 == 0, MTN_E_SHAPE,
              "cast_colsum: rows=%d cols=%d ld_src=%d (cols and ld must be multiples of 8)", rows, cols, ld_src);
  MTN_REQUIRE(aligned16(src) && (!dst_f16 || (aligned16(dst_f16) && ld_dst % 8 == 0 && ld_dst >= cols)) &&
                  (!relu_mask || (aligned16(relu_mask) && ld_mask % 8 == 0 && ld_mask >= cols)),
              MTN_E_ALIGN, "cast_colsum: alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vcols = cols / 8;
  // rows per block: at least 64; large inputs get about 8 blocks per SM in total, so that the column-sum atomics per
  // address stay in the tens (they serialise in L2) instead of rows / 64
  const int gx = (vcols + 31) / 32;
  int rpb = CC_ROWS_PER_BLOCK;
  const int want_y = (8 * 148 + gx - 1) / gx;
  if ((rows + rpb - 1) / rpb > want_y) rpb = ((rows + want_y - 1) / want_y + 31) / 32 * 32;
  dim3 grid(gx, (rows + rpb - 1) / rpb), block(32, 8);
  __half* d16 = reinterpret_cast<__half*>(dst_f16);
  const __half* mk = reinterpret_cast<const __half*>(relu_mask);
  if (src_is_f16)
    MTN_CHECK_CUDA(launch_kernel(cast_colsum_kernel<__half>, grid, block, 0, st, reinterpret_cast<const __half*>(src), ld_src,
                                 d16, ld_dst, mk, ld_mask, rows, vcols, rpb, scale, alpha, colsum, drop, multimem));
  else
    MTN_CHECK_CUDA(launch_kernel(cast_colsum_kernel<float>, grid, block, 0, st, reinterpret_cast<const float*>(src), ld_src,
                                 d16, ld_dst, mk, ld_mask, rows, vcols, rpb, scale, alpha, colsum, drop, multimem));
  return MTN_OK;
}

extern "C" int mtn_grad_absmax(const float* x, size_t n, uint32_t* slot, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(x && slot, MTN_E_ARG, "grad_absmax: NULL pointer");
  MTN_REQUIRE(n > 0 && n % 4 == 0 && aligned16(x), MTN_E_ALIGN, "grad_absmax: n=%zu must be a multiple of 4, x 16-byte aligned", n);
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  MTN_CHECK_CUDA(launch_kernel(absmax_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, n4,
                               reinterpret_cast<unsigned*>(slot)));
  return MTN_OK;
}

extern "C" int mtn_adam_advance(float* state, float noam_factor, float model_size, float warmup, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(state != nullptr, MTN_E_ARG, "adam_advance: NULL pointer");
  MTN_CHECK_CUDA(launch_kernel(adam_advance_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), state, noam_factor,
                               model_size, warmup));
  return MTN_OK;
}

extern "C" int mtn_adam_step(float* p, float* g, float* m, float* v, void* p_f16, size_t n, const float* state, int zero_grad,
                             void* stream) {
  using namespace mtn;
  MTN_REQUIRE(p && g && m && v && state, MTN_E_ARG, "adam_step: NULL pointer");
  MTN_REQUIRE(n > 0 && n % 4 == 0 && aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) &&
                  (!p_f16 || (reinterpret_cast<uintptr_t>(p_f16) & 7) == 0),
              MTN_E_ALIGN, "adam_step: n=%zu must be a multiple of 4 and the buffers 16-byte aligned", n);
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 4736) blocks = 4736;
  MTN_CHECK_CUDA(launch_kernel(adam_step_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g, m,
                               v, reinterpret_cast<__half*>(p_f16), n4, state, zero_grad));
  return MTN_OK;
}

extern "C" int mtn_seed_bump(uint64_t* seed, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(seed != nullptr, MTN_E_ARG, "seed_bump: NULL pointer");
  MTN_CHECK_CUDA(launch_kernel(seed_bump_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream),
                               reinterpret_cast<unsigned long long*>(seed)));
  return MTN_OK;
}

extern "C" int mtn_zero(void* p, size_t bytes, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(p != nullptr, MTN_E_ARG, "zero: NULL pointer");
  MTN_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, static_cast<cudaStream_t>(stream)));
  return MTN_OK;
}

extern "C" int mtn_scale_f32(const float* x, const float* alpha, float* y, size_t n, int accumulate, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(x && y, MTN_E_ARG, "scale_f32: NULL pointer");
  MTN_REQUIRE(n > 0 && n % 4 == 0 && aligned16(x) && aligned16(y), MTN_E_ALIGN,
              "scale_f32: n=%zu must be a multiple of 4, pointers 16-byte aligned", n);
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 2368) blocks = 2368;
  MTN_CHECK_CUDA(launch_kernel(scale_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x,
                               alpha, y, n4, accumulate));
  return MTN_OK;
}

extern "C" int mtn_grad_scale(uint32_t* slot, float* scale2, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(slot && scale2, MTN_E_ARG, "grad_scale: NULL pointer");
  MTN_CHECK_CUDA(launch_kernel(grad_scale_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream),
                               reinterpret_cast<unsigned*>(slot), scale2));
  return MTN_OK;
}

extern "C" int mtn_attn_delta(const void* dO, int lddo, const void* O, int ldo, int B, int Lq, int h, int d_k, float* delta,
                              void* stream) {
  using namespace mtn;
  MTN_REQUIRE(dO && O && delta, MTN_E_ARG, "attn_delta: NULL pointer");
  MTN_REQUIRE(B > 0 && Lq > 0 && h > 0 && (d_k == 32 || d_k == 64), MTN_E_SHAPE, "attn_delta: B=%d Lq=%d h=%d d_k=%d", B, Lq,
              h, d_k);
  MTN_REQUIRE(aligned16(dO) && aligned16(O) && lddo % 8 == 0 && ldo % 8 == 0 && lddo >= h * d_k && ldo >= h * d_k, MTN_E_ALIGN,
              "attn_delta: alignment / leading dimensions");
  MTN_CHECK_CUDA(launch_kernel(attn_delta_kernel, dim3((B * Lq + 7) / 8), dim3(256), 0, static_cast<cudaStream_t>(stream),
                               reinterpret_cast<const __half*>(dO), lddo, reinterpret_cast<const __half*>(O), ldo, B, Lq, h,
                               d_k, delta));
  return MTN_OK;
}

extern "C" int mtn_log_softmax_bwd(const float* y, int ldy, const float* dy, int lddy, int rows, int V, float* dz, int lddz,
                                   void* stream) {
  using namespace mtn;
  MTN_REQUIRE(y && dy && dz, MTN_E_ARG, "log_softmax_bwd: NULL pointer");
  MTN_REQUIRE(rows > 0 && V > 0 && ldy >= V && lddy >= V && lddz >= V, MTN_E_SHAPE, "log_softmax_bwd: rows=%d V=%d", rows, V);
  MTN_CHECK_CUDA(launch_kernel(log_softmax_bwd_kernel, dim3(rows), dim3(128), 0, static_cast<cudaStream_t>(stream), y, ldy, dy,
                               lddy, V, dz, lddz));
  return MTN_OK;
}

extern "C" int mtn_label_smoothing_loss_bwd(const float* z, int ld, int rows, int V, const int64_t* target,
                                            int64_t padding_idx, float smoothing, float gscale, const float* gout,
                                            float* dz, int lddz, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(z && target && dz, MTN_E_ARG, "label_smoothing_bwd: NULL pointer");
  MTN_REQUIRE(rows > 0 && V > 2 && ld >= V && lddz >= V && padding_idx >= 0 && padding_idx < V, MTN_E_SHAPE,
              "label_smoothing_bwd: rows=%d V=%d ld=%d lddz=%d padding_idx=%lld", rows, V, ld, lddz, (long long)padding_idx);
  MTN_REQUIRE(workspace && workspace_bytes >= 256 && aligned16(workspace), MTN_E_WORKSPACE,
              "label_smoothing_bwd: workspace (>= 256 bytes, 16-byte aligned) required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* flag = reinterpret_cast<unsigned long long*>(workspace);
  const long long* t64 = reinterpret_cast<const long long*>(target);
  MTN_CHECK_CUDA(launch_kernel(pad_index_sum_bwd_kernel, dim3(1), dim3(1024), 0, st, t64, rows, (long long)padding_idx, flag));
  MTN_CHECK_CUDA(launch_kernel(label_smoothing_bwd_kernel, dim3(rows), dim3(128), 0, st, z, ld, V, t64, (long long)padding_idx,
                               smoothing, (const unsigned long long*)flag, gscale, gout, dz, lddz));
  return MTN_OK;
}
