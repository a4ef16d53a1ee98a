// Small-M kernels for KV-cached, last-token-only decoding (reference call form
// data_utils.py:202-210): at one target position per dialogue a decoding step is ~170 dependent launches over M = B
// (<= 128) rows.  The tcgen05 kernels pay their fixed costs (TMEM allocation, TMA pipeline fill, 128-row tiles that are
// mostly padding) on every one of them -- ~6 us per launch for microseconds of work.  These two kernels do the same
// arithmetic (f16 operands, f32 accumulate, the same softmax formulas) for few rows with nothing to set up:
//
//   rows_linear_kernel   out[M, N] = act(A[M, K] W[N, K]^T + bias) (+ residual):  one CTA per 8 output columns, one warp
//                        per 16 rows, mma.sync.m16n8k16 with both operands read straight from global memory by 16-byte
//                        loads (a k-permutation that A and B share makes the fragments line up without shared memory);
//                        the weight matrix is spread over N/8 CTAs, i.e. streamed by the whole machine.
//   decode_attn_kernel   O = softmax(mask(Q K^T / sqrt(d_k))) V for R <= 8 query rows per batch element and d_k = 64:
//                        one warp per (batch element, head); keys across lanes for the scores, dims across lanes for
//                        P V; online softmax in the log2 domain with the reference's FINITE -1e9 (mtn.py:227).
#include <math_constants.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "host.h"

namespace mtn {

struct RowsLinearParams {
  const __half* A; int lda;
  const __half* W; int ldw;
  const float* bias;
  int M, N, K, act;
  float* x32; int ld32;      // optional f32 output; with `residual` it is updated in place: x32 += result
  int residual;
  __half* out16; int ld16;   // optional f16 output
  uint32_t zero;             // always 0 (see the load-scheduling note in the kernel)
};

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// volatile 16-byte loads: the compiler keeps them in program order ahead of the (volatile) mma sequence, i.e. ALL of a
// warp's loads are in flight at once instead of being sunk next to their uses to save registers
__device__ __forceinline__ uint4 ld_cg_v4(const void* ptr) {    // L2 (coherent): activations written by the predecessor kernel
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* ptr) {    // read-only path: weights
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}

// grid = (N / 8, ceil(M / 16)), block = 32 * KS warps: CTA = 16 rows x 8 output columns, the contraction split over its
// KS warps (each issues ALL of its 16-byte loads before the first mma: one memory round trip per warp), partial sums
// combined through shared memory in a fixed order (deterministic).  K % (32 * KS) == 0.
constexpr int RL_MAX_KS = 8;
// The kernels' bodies are device functions of a VIRTUAL block index (bx, by): the stand-alone kernels pass blockIdx, the
// persistent decoding-step kernel at the end of this file walks the virtual blocks of one launch after the other.
template <int CPW>   // 32-wide k chunks per warp
__device__ __forceinline__ void rows_linear_body(const RowsLinearParams& p, int bx, int by, int KS, float (*part)[32][4]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int n0 = bx * 8;
  const int r0 = by * 16 + g, r1 = r0 + 8;  // the two rows of this thread's fragments
  const int kbase = warp * CPW * 32;
  // rows beyond M read row M-1 (valid memory) and are not stored
  const __half* a_lo = p.A + (size_t)min(r0, p.M - 1) * p.lda + 8 * q + kbase;
  const __half* a_hi = p.A + (size_t)min(r1, p.M - 1) * p.lda + 8 * q + kbase;
  const __half* w_row = p.W + (size_t)(n0 + g) * p.ldw + 8 * q + kbase;
  // Per 32-wide k chunk a thread loads 8 consecutive k of its rows (A: rows g / g + 8; W: output column g) with one
  // 16-byte load each.  Both mma k16 steps then use the SAME mapping logical-k -> actual-k for A and B
  // (logical 2q+j -> 8q+j, logical 2q+8+j -> 8q+2+j; second step: +4), so the products pair up correctly.
  uint4 xa[CPW], xb[CPW], w[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    xa[c] = ld_cg_v4(a_lo + 32 * c);
    xb[c] = ld_cg_v4(a_hi + 32 * c);
    w[c] = ld_nc_v4(w_row + 32 * c);
  }
  // The accumulator's initial value is made to DEPEND on every loaded register (their XOR, masked with a run-time
  // zero the compiler cannot see through): ptxas must then complete all loads before the first mma instead of sinking
  // each load next to its use to save registers (load -> mma -> load ...: one memory round trip per chunk).
  uint32_t fold = 0u;
#pragma unroll
  for (int c = 0; c < CPW; ++c)
    fold ^= (xa[c].x ^ xa[c].y ^ xa[c].z ^ xa[c].w) ^ (xb[c].x ^ xb[c].y ^ xb[c].z ^ xb[c].w) ^ (w[c].x ^ w[c].y ^ w[c].z ^ w[c].w);
  float c4[4] = {__uint_as_float(fold & p.zero), 0.f, 0.f, 0.f};   // p.zero == 0: exactly 0.0f
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    mma_16816(c4, xa[c].x, xb[c].x, xa[c].y, xb[c].y, w[c].x, w[c].y);
    mma_16816(c4, xa[c].z, xb[c].z, xa[c].w, xb[c].w, w[c].z, w[c].w);
  }
  if (KS > 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) part[warp][lane][i] = c4[i];
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float acc = part[0][lane][i];
      for (int ws = 1; ws < KS; ++ws) acc += part[ws][lane][i];   // fixed order
      c4[i] = acc;
    }
  }
  const int col = n0 + 2 * q;
  float b0 = 0.f, b1 = 0.f;
  if (p.bias != nullptr) {
    b0 = __ldg(p.bias + col);
    b1 = __ldg(p.bias + col + 1);
  }
  float v[4] = {c4[0] + b0, c4[1] + b1, c4[2] + b0, c4[3] + b1};
  if (p.act == MTN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int r = hh == 0 ? r0 : r1;
    if (r >= p.M) continue;
    float y0 = v[2 * hh], y1 = v[2 * hh + 1];
    if (p.x32 != nullptr) {
      float2* o = reinterpret_cast<float2*>(p.x32 + (size_t)r * p.ld32 + col);
      if (p.residual) {  // every element has exactly one owner thread: a plain read-modify-write, deterministic
        const float2 old = __ldcg(o);
        y0 += old.x;
        y1 += old.y;
      }
      *o = make_float2(y0, y1);
    }
    if (p.out16 != nullptr) *reinterpret_cast<uint32_t*>(p.out16 + (size_t)r * p.ld16 + col) = pack_f16x2_sat(y0, y1);
  }
}

template <int CPW>
__global__ void __launch_bounds__(32 * RL_MAX_KS) rows_linear_kernel(const RowsLinearParams p) {
  __shared__ float part[RL_MAX_KS][32][4];
  pdl_launch_dependents();
  pdl_wait();
  rows_linear_body<CPW>(p, blockIdx.x, blockIdx.y, blockDim.x >> 5, part);
}

// ----------------------------------------------------------------------------------------------------------------
// The same with the reference's LayerNorm (mtn.py:111-114) in front:  out16[M, N] = act(LN(x)[M, K] W[N, K]^T + bias),
// x f32.  A CTA (16 rows x 8 output columns, KS warps over the contraction) first requests everything it will need --
// the raw x values of its mma fragments, a_2 / b_2, its weights -- then its warps compute mean / 1/(std + eps) of the 16
// rows with exactly the arithmetic (and summation order) of layernorm_rows_kernel, and the fragments are normalised and
// rounded to f16 in registers: bit-identical to mtn_layernorm_fwd + mtn_rows_linear_fwd, one launch, one memory round
// trip.  The statistics are recomputed by every CTA of a row block (N / 8 of them): 2 KB per row out of L2.
struct RowsLnLinearParams {
  const float* x; int ldx;
  const float* a2; const float* b2; float eps;
  const __half* W; int ldw;
  const float* bias;
  int M, N, act;
  __half* out16; int ld16;
  uint32_t zero;
};

template <int VPL>   // K = 128 * VPL
__device__ __forceinline__ void rows_ln_linear_body(const RowsLnLinearParams& p, int bx, int by, float (*part)[32][4], float2* stats) {
  constexpr int K = 128 * VPL;
  constexpr int KS = (K / 32 >= 8) ? 8 : K / 32;   // warps over the contraction
  constexpr int CPW = K / (32 * KS);              // 32-wide chunks per warp
  constexpr int RPW = 16 / KS;                    // rows whose statistics a warp computes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int n0 = bx * 8;
  const int rb = by * 16;
  const int r0 = rb + g, r1 = r0 + 8;
  const int kbase = warp * CPW * 32 + 8 * q;
  const float* x_lo = p.x + (size_t)min(r0, p.M - 1) * p.ldx + kbase;
  const float* x_hi = p.x + (size_t)min(r1, p.M - 1) * p.ldx + kbase;
  // (1) everything the mma phase needs, requested now
  float4 xl[CPW][2], xh[CPW][2], ga[CPW][2], gb[CPW][2];
  uint4 w[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      xl[c][u] = __ldcg(reinterpret_cast<const float4*>(x_lo + 32 * c) + u);
      xh[c][u] = __ldcg(reinterpret_cast<const float4*>(x_hi + 32 * c) + u);
      ga[c][u] = __ldg(reinterpret_cast<const float4*>(p.a2 + kbase + 32 * c) + u);
      gb[c][u] = __ldg(reinterpret_cast<const float4*>(p.b2 + kbase + 32 * c) + u);
    }
    w[c] = ld_nc_v4(p.W + (size_t)(n0 + g) * p.ldw + kbase + 32 * c);
  }
  // (2) row statistics: the arithmetic of layernorm_rows_kernel<VPL> (lane l holds float4 l + 32 i of the row)
#pragma unroll
  for (int rr = 0; rr < RPW; ++rr) {
    const int rl = warp * RPW + rr;
    const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)min(rb + rl, p.M - 1) * p.ldx);
    float4 v[VPL];
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = __ldcg(xr + lane + 32 * i);
      sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
    const float mean = sm * (1.f / K);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) stats[rl] = make_float2(mean, 1.f / (sqrtf(ss * (1.f / (K - 1))) + p.eps));
  }
  __syncthreads();
  // (3) normalise the fragments: a_2 * (x - mean) * inv + b_2, rounded to f16 like the LayerNorm kernel's y_f16
  const float2 sl = stats[g], sh = stats[g + 8];
  float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    uint32_t al[4], ah[4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float4 a = ga[c][u], bq = gb[c][u], l4 = xl[c][u], h4 = xh[c][u];
      al[2 * u] = pack_f16x2_sat(a.x * (l4.x - sl.x) * sl.y + bq.x, a.y * (l4.y - sl.x) * sl.y + bq.y);
      al[2 * u + 1] = pack_f16x2_sat(a.z * (l4.z - sl.x) * sl.y + bq.z, a.w * (l4.w - sl.x) * sl.y + bq.w);
      ah[2 * u] = pack_f16x2_sat(a.x * (h4.x - sh.x) * sh.y + bq.x, a.y * (h4.y - sh.x) * sh.y + bq.y);
      ah[2 * u + 1] = pack_f16x2_sat(a.z * (h4.z - sh.x) * sh.y + bq.z, a.w * (h4.w - sh.x) * sh.y + bq.w);
    }
    mma_16816(c4, al[0], ah[0], al[1], ah[1], w[c].x, w[c].y);
    mma_16816(c4, al[2], ah[2], al[3], ah[3], w[c].z, w[c].w);
  }
  if (KS > 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) part[warp][lane][i] = c4[i];
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float acc = part[0][lane][i];
#pragma unroll
      for (int ws = 1; ws < KS; ++ws) acc += part[ws][lane][i];
      c4[i] = acc;
    }
  }
  const int col = n0 + 2 * q;
  float b0 = 0.f, b1 = 0.f;
  if (p.bias != nullptr) {
    b0 = __ldg(p.bias + col);
    b1 = __ldg(p.bias + col + 1);
  }
  float v[4] = {c4[0] + b0, c4[1] + b1, c4[2] + b0, c4[3] + b1};
  if (p.act == MTN_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (r0 < p.M) *reinterpret_cast<uint32_t*>(p.out16 + (size_t)r0 * p.ld16 + col) = pack_f16x2_sat(v[0], v[1]);
  if (r1 < p.M) *reinterpret_cast<uint32_t*>(p.out16 + (size_t)r1 * p.ld16 + col) = pack_f16x2_sat(v[2], v[3]);
}

template <int VPL>
__global__ void __launch_bounds__(256) rows_ln_linear_kernel(const RowsLnLinearParams p) {
  constexpr int KS = (128 * VPL / 32 >= 8) ? 8 : 128 * VPL / 32;
  __shared__ float part[KS][32][4];
  __shared__ float2 stats[16];
  pdl_launch_dependents();
  pdl_wait();
  rows_ln_linear_body<VPL>(p, blockIdx.x, blockIdx.y, part, stats);
}

// ----------------------------------------------------------------------------------------------------------------
struct DecodeAttnParams {
  const __half *q, *k, *v;
  __half* out;
  int ldq, ldk, ldv, ldo;
  long long sq, sk, sv, so;  // batch strides (elements)
  const uint32_t* mask_bits;
  int mask_rows_q, mask_words;
  int B, h, R, Lk;
  float scale;
};

__device__ __forceinline__ float da_ex2(float x) {   // the tensor-core path's exponential (csrc/attn.cu)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int DA_MAXR = 8;
constexpr int DA_WARPS = 8;

__device__ __forceinline__ uint32_t ld_cg_b32(const void* ptr) {
  uint32_t r;
  asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(r) : "l"(ptr));
  return r;
}

// One CTA (8 warps) per (batch element, head), d_k = 64, R <= 8 query rows.  The keys are dealt to the warps in chunks
// of 32 (warp w takes chunks w, w + 8, ...; up to 256 keys in one pass).  Per chunk a warp requests EVERYTHING first --
// its lane's key row (128 B) and, for P V, its two output dims of all 32 value rows -- so a launch is one memory round
// trip deep (the K / V caches of a decoding step do not fit the L2: this kernel is HBM-latency bound otherwise); then an
// online softmax (lane = key for the scores, lane = two dims for P V) and a fixed-order merge of the warps' partial
// (max, sum, O) triples through shared memory.
template <int R>
struct DecodeAttnShared {
  float sq_[R][64];
  float sm_[DA_WARPS][R], sl_[DA_WARPS][R], so_[DA_WARPS][R][64];
};

template <int R>
__device__ __forceinline__ void decode_attn_body(const DecodeAttnParams& p, int bx, DecodeAttnShared<R>& sh) {
  float (&sq_)[R][64] = sh.sq_;
  float (&sm_)[DA_WARPS][R] = sh.sm_;
  float (&sl_)[DA_WARPS][R] = sh.sl_;
  float (&so_)[DA_WARPS][R][64] = sh.so_;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = bx / p.h, hd = bx % p.h;
  constexpr float LOG2E = 1.4426950408889634f;
  const float c1 = p.scale * LOG2E, t_masked = -1e9f * LOG2E;
  const __half* kb = p.k + (size_t)b * p.sk + hd * 64;
  const __half* vb = p.v + (size_t)b * p.sv + hd * 64;
  float m_run[R], l_run[R], o0[R], o1[R];
#pragma unroll
  for (int r = 0; r < R; ++r) m_run[r] = -CUDART_INF_F, l_run[r] = 0.f, o0[r] = 0.f, o1[r] = 0.f;
  bool have_q = false;
  for (int k0 = warp * 32; k0 < p.Lk || !have_q; k0 += 32 * DA_WARPS) {
    const bool active = k0 < p.Lk;            // (a warp without keys still takes part in the query staging below)
    const int key = k0 + lane;
    const bool inb = key < p.Lk;
    // ---- request the chunk: key row of this lane, value rows (this lane's two dims), mask words
    uint4 kv[8];
    uint32_t vv[32];
    uint32_t mw[R];
    if (active) {
      const uint8_t* kr = reinterpret_cast<const uint8_t*>(kb + (size_t)min(key, p.Lk - 1) * p.ldk);
#pragma unroll
      for (int c = 0; c < 8; ++c) kv[c] = ld_cg_v4(kr + 16 * c);
#pragma unroll
      for (int u = 0; u < 32; ++u) vv[u] = ld_cg_b32(vb + (size_t)min(k0 + u, p.Lk - 1) * p.ldv + 2 * lane);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        mw[r] = 0xffffffffu;
        if (p.mask_bits != nullptr) {
          const int mq = p.mask_rows_q == 1 ? 0 : r;
          mw[r] = __ldcg(p.mask_bits + ((size_t)b * p.mask_rows_q + mq) * p.mask_words + (k0 >> 5));
        }
      }
    }
    if (!have_q) {   // queries -> shared memory as f32 (requested after the chunk's loads, used after the barrier)
      for (int i = threadIdx.x; i < R * 32; i += 32 * DA_WARPS) {
        const int r = i >> 5, l2 = i & 31;
        const float2 f = __half22float2(__ldcg(reinterpret_cast<const __half2*>(p.q + (size_t)b * p.sq + (size_t)r * p.ldq + hd * 64) + l2));
        sq_[r][2 * l2] = f.x;
        sq_[r][2 * l2 + 1] = f.y;
      }
      __syncthreads();
      have_q = true;
    }
    if (!active) break;
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const __half2* hp = reinterpret_cast<const __half2*>(&kv[c]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 kf = __half22float2(hp[j]);
#pragma unroll
        for (int r = 0; r < R; ++r) s[r] = fmaf(sq_[r][8 * c + 2 * j], kf.x, fmaf(sq_[r][8 * c + 2 * j + 1], kf.y, s[r]));
      }
    }
    // ---- online softmax (log2 domain)
    float pr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool keep = (mw[r] >> lane) & 1u;
      const float t = inb ? (keep ? s[r] * c1 : t_masked) : -CUDART_INF_F;
      float mx = t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[r], mx);
      const float e = da_ex2(t - m_new);     // 0 for keys beyond Lk
      float sum = e;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float alpha = da_ex2(m_run[r] - m_new);  // 0 on the first chunk (m_run = -inf)
      l_run[r] = l_run[r] * alpha + sum;
      o0[r] *= alpha;
      o1[r] *= alpha;
      m_run[r] = m_new;
      pr[r] = __half2float(__float2half_rn(e));     // P is rounded to f16 before P V, like the tensor-core path
    }
    // ---- P V: lane owns dims 2*lane, 2*lane+1 (rows beyond Lk were clamped to a valid row; their P is 0)
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&vv[u]));
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float pk = __shfl_sync(0xffffffffu, pr[r], u);
        o0[r] = fmaf(pk, vf.x, o0[r]);
        o1[r] = fmaf(pk, vf.y, o1[r]);
      }
    }
  }
  // ---- merge the warps' partial results (fixed order)
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (lane == 0) sm_[warp][r] = m_run[r], sl_[warp][r] = l_run[r];
    so_[warp][r][2 * lane] = o0[r];
    so_[warp][r][2 * lane + 1] = o1[r];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * 32; i += 32 * DA_WARPS) {
    const int r = i >> 5, l2 = i & 31;
    float m = sm_[0][r];
#pragma unroll
    for (int w = 1; w < DA_WARPS; ++w) m = fmaxf(m, sm_[w][r]);
    float l = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float f = da_ex2(sm_[w][r] - m);   // 0 for a warp that saw no key (its max is -inf)
      l = fmaf(sl_[w][r], f, l);
      a0 = fmaf(so_[w][r][2 * l2], f, a0);
      a1 = fmaf(so_[w][r][2 * l2 + 1], f, a1);
    }
    const float inv = 1.f / l;
    *reinterpret_cast<uint32_t*>(p.out + (size_t)b * p.so + (size_t)r * p.ldo + hd * 64 + 2 * l2) = pack_f16x2_sat(a0 * inv, a1 * inv);
  }
}

template <int R>
__global__ void __launch_bounds__(32 * DA_WARPS) decode_attn_kernel(const DecodeAttnParams p) {
  __shared__ DecodeAttnShared<R> sh;
  pdl_launch_dependents();
  pdl_wait();
  decode_attn_body<R>(p, blockIdx.x, sh);
}

// ----------------------------------------------------------------------------------------------------------------
// One decoding step as ONE persistent kernel.  A KV-cached step is ~130 dependent few-row launches (above); inside a CUDA
// graph each still costs a kernel boundary (~3 us) for ~1 us of work.  While a program is being RECORDED
// (mtn_prog_begin), the few-row entry points below and mtn_layernorm_fwd append their launch -- kernel kind, grid, the
// parameter struct they would have launched with -- to a stage list instead of launching; mtn_prog_launch then runs
// the list in one cooperative kernel: every CTA walks the virtual blocks of stage s (the same device functions as the
// stand-alone kernels: identical arithmetic, bit-identical results), then all CTAs meet at a grid barrier (one atomic
// per CTA + acquire polling) before stage s+1.  The next stage's descriptor is fetched under the current stage's work.
// ----------------------------------------------------------------------------------------------------------------
enum { PK_LINEAR = 1, PK_LN_LINEAR = 2, PK_ATTN = 3, PK_LN = 4 };

struct LnStageParams {
  const float* x; const float* a2; const float* b2;
  float eps; int rows, rpg;
  float* y32; __half* y16;
};

constexpr int PROG_PARAM_BYTES = 128;
struct ProgStage {
  int kind, tparam, gx, gy;
  alignas(8) unsigned char params[PROG_PARAM_BYTES];
};
static_assert(sizeof(RowsLinearParams) <= PROG_PARAM_BYTES && sizeof(RowsLnLinearParams) <= PROG_PARAM_BYTES &&
                  sizeof(DecodeAttnParams) <= PROG_PARAM_BYTES && sizeof(LnStageParams) <= PROG_PARAM_BYTES,
              "stage parameter area");
constexpr int PROG_STAGE_WORDS = sizeof(ProgStage) / 4;
static_assert(sizeof(ProgStage) % 4 == 0 && PROG_STAGE_WORDS <= 64, "stage descriptor");

// LayerNorm of 8 rows per virtual block, d = 512: the arithmetic (and summation order) of layernorm_rows_kernel<4>
__device__ __forceinline__ void ln_rows_body(const LnStageParams& p, int bx) {
  constexpr int VPL = 4, D = 512;
  const int row = bx * 8 + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const float* a2 = p.a2 + (size_t)(row / p.rpg) * D;
  const float* b2 = p.b2 + (size_t)(row / p.rpg) * D;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)row * D);
  float4 v[VPL];
  float sm = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = __ldcg(xr + lane + 32 * i);
    sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
  const float mean = sm * (1.f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / (sqrtf(ss * (1.f / (D - 1))) + p.eps);
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
    if (p.y32) reinterpret_cast<float4*>(p.y32 + (size_t)row * D)[c4] = o;
    if (p.y16) reinterpret_cast<uint2*>(p.y16 + (size_t)row * D)[c4] = make_uint2(pack_f16x2_sat(o.x, o.y), pack_f16x2_sat(o.z, o.w));
  }
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* ptr) {
  uint32_t r;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(ptr));
  return r;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned* ptr) {
  uint32_t r;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(ptr) : "memory");
  return r;
}

// All CTAs of the (co-resident, cooperative) grid have finished the stage: the writes of every CTA are visible to every
// CTA afterwards.  `target` = arrivals expected so far (monotonic counter, zeroed before the launch).  Bounded: a
// protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void prog_grid_sync(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // release: the CTA's writes (ordered before this thread by the barrier above) become visible before the arrival
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target) {
      if (++spins > (1u << 27)) {
        printf("mtn_b200: decode program grid barrier timeout, block %d target %u\n", blockIdx.x, target);
        __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 2) decode_prog_kernel(const ProgStage* prog, int n_stages, unsigned* counter) {
  __shared__ ProgStage sst[2];
  __shared__ float part[RL_MAX_KS][32][4];
  __shared__ float2 stats[16];
  __shared__ DecodeAttnShared<1> ash;
  if (threadIdx.x < PROG_STAGE_WORDS)
    reinterpret_cast<uint32_t*>(&sst[0])[threadIdx.x] = ld_cg_u32(reinterpret_cast<const uint32_t*>(prog) + threadIdx.x);
  __syncthreads();
  for (int s = 0; s < n_stages; ++s) {
    uint32_t nextw = 0u;   // the next stage's descriptor, in flight under this stage's work
    if (s + 1 < n_stages && threadIdx.x < PROG_STAGE_WORDS)
      nextw = ld_cg_u32(reinterpret_cast<const uint32_t*>(prog + s + 1) + threadIdx.x);
    const ProgStage& st = sst[s & 1];
    const int kind = st.kind, tparam = st.tparam, gx = st.gx, nvb = st.gx * st.gy;
    for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
      const int bx = vb % gx, by = vb / gx;
      if (kind == PK_LINEAR) {
        const RowsLinearParams p = *reinterpret_cast<const RowsLinearParams*>(st.params);
        if (tparam == 2) rows_linear_body<2>(p, bx, by, RL_MAX_KS, part);
        else rows_linear_body<8>(p, bx, by, RL_MAX_KS, part);
      } else if (kind == PK_LN_LINEAR) {
        const RowsLnLinearParams p = *reinterpret_cast<const RowsLnLinearParams*>(st.params);
        rows_ln_linear_body<4>(p, bx, by, part, stats);
      } else if (kind == PK_ATTN) {
        const DecodeAttnParams p = *reinterpret_cast<const DecodeAttnParams*>(st.params);
        decode_attn_body<1>(p, bx, ash);
      } else {
        const LnStageParams p = *reinterpret_cast<const LnStageParams*>(st.params);
        ln_rows_body(p, bx);
      }
      __syncthreads();   // the shared scratch is reused by the next virtual block
    }
    if (s + 1 < n_stages) {
      if (threadIdx.x < PROG_STAGE_WORDS) reinterpret_cast<uint32_t*>(&sst[(s + 1) & 1])[threadIdx.x] = nextw;
      prog_grid_sync(counter, (unsigned)(s + 1) * gridDim.x);
    }
  }
}

// ---- recorder (host side; one recording at a time per thread)
static thread_local std::vector<ProgStage>* g_prog = nullptr;

bool prog_recording() { return g_prog != nullptr; }

static int prog_push(int kind, int tparam, dim3 grid, const void* params, size_t bytes) {
  ProgStage st;
  memset(&st, 0, sizeof(st));
  st.kind = kind; st.tparam = tparam; st.gx = (int)grid.x; st.gy = (int)grid.y;
  memcpy(st.params, params, bytes);
  g_prog->push_back(st);
  return MTN_OK;
}

int prog_push_layernorm(const float* x, const float* a2, const float* b2, float eps, int rows, int d, int rows_per_group,
                        float* y32, void* y16) {
  MTN_REQUIRE(d == 512 && aligned16(x) && aligned16(a2) && aligned16(b2) && (!y32 || aligned16(y32)) && (!y16 || aligned16(y16)),
              MTN_E_SHAPE, "decode program: LayerNorm stage needs d = 512 and 16-byte aligned pointers (d = %d)", d);
  LnStageParams p{x, a2, b2, eps, rows, rows_per_group, y32, reinterpret_cast<__half*>(y16)};
  return prog_push(PK_LN, 4, dim3((rows + 7) / 8, 1), &p, sizeof(p));
}

}  // namespace mtn

extern "C" int mtn_prog_begin(void) {
  using namespace mtn;
  MTN_REQUIRE(g_prog == nullptr, MTN_E_ARG, "prog_begin: a program is already being recorded");
  g_prog = new std::vector<ProgStage>();
  return MTN_OK;
}
extern "C" int mtn_prog_recording(void) { return mtn::g_prog != nullptr ? 1 : 0; }
extern "C" int mtn_prog_stage_bytes(void) { return (int)sizeof(mtn::ProgStage); }
// Ends the recording.  host_dst (capacity bytes; pinned memory if it is to be uploaded asynchronously) receives the stage
// list; returns the number of stages through *n_stages.  host_dst == NULL: the recording is discarded.
extern "C" int mtn_prog_end(void* host_dst, size_t capacity, int* n_stages) {
  using namespace mtn;
  MTN_REQUIRE(g_prog != nullptr, MTN_E_ARG, "prog_end: no program is being recorded");
  std::vector<ProgStage>* v = g_prog;
  g_prog = nullptr;
  const size_t n = v->size(), bytes = n * sizeof(ProgStage);
  int rc = MTN_OK;
  if (host_dst != nullptr) {
    if (bytes > capacity) rc = set_error(MTN_E_WORKSPACE, "prog_end: %zu stages need %zu bytes, buffer has %zu", n, bytes, capacity);
    else if (n > 0) memcpy(host_dst, v->data(), bytes);
  }
  if (n_stages != nullptr) *n_stages = (int)n;
  delete v;
  return rc;
}
// Runs a recorded program (device copy of the stage list) as one cooperative kernel on `stream`; `counter`: 4 bytes of
// device memory for the grid barrier (zeroed here, stream-ordered).  Graph-capturable.
extern "C" int mtn_prog_launch(const void* dev_prog, int n_stages, void* counter, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(dev_prog && counter && n_stages > 0, MTN_E_ARG, "prog_launch: NULL pointer / empty program");
  MTN_REQUIRE(!prog_recording(), MTN_E_ARG, "prog_launch: a program is being recorded");
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    MTN_CHECK_CUDA(cudaGetDevice(&dev));
    MTN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MTN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_prog_kernel, 256, 0));
    MTN_REQUIRE(per_sm > 0, MTN_E_CUDA, "prog_launch: the program kernel does not fit an SM");
    grid = sms * (per_sm > 2 ? 2 : per_sm);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTN_CHECK_CUDA(cudaMemsetAsync(counter, 0, 4, st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const ProgStage* pp = static_cast<const ProgStage*>(dev_prog);
  unsigned* cc = static_cast<unsigned*>(counter);
  MTN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decode_prog_kernel, pp, n_stages, cc));
  return MTN_OK;
}

extern "C" int mtn_rows_linear_supported(int M, int N, int K) {
  if (!(M > 0 && M <= 128 && N > 0 && N % 8 == 0 && K > 0 && K % 32 == 0)) return 0;
  const int chunks = K / 32;
  for (int ks = 1; ks <= mtn::RL_MAX_KS; ++ks)
    if (chunks % ks == 0 && chunks / ks <= 8) return 1;
  return 0;
}

// Same contract as mtn_linear_fwd for few rows (M <= 128): A f16 [M, K], W f16 [N, K], optional bias / ReLU, f16 and/or
// f32 outputs; addend must be the f32 output itself (in-place residual, x += ...) or NULL.
extern "C" int mtn_rows_linear_fwd(const MtnLinearArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->A && a->W && (a->out_f32 || a->out_f16), MTN_E_ARG, "rows_linear: NULL pointer");
  MTN_REQUIRE(mtn_rows_linear_supported(a->M, a->N, a->K), MTN_E_SHAPE, "rows_linear: M=%d N=%d K=%d (M <= 128, N %% 8 == 0, K %% 32 == 0)",
              a->M, a->N, a->K);
  MTN_REQUIRE(a->addend == nullptr || (a->addend == a->out_f32 && a->ld_add == a->ld32 && a->add_period == 0), MTN_E_ARG,
              "rows_linear: the addend must be the f32 output itself (in-place residual)");
  MTN_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0 && aligned16(a->A) && aligned16(a->W) && a->batch <= 1 && a->drop_seed == nullptr &&
                  !a->out16_pre_add,
              MTN_E_ALIGN, "rows_linear: alignment / unsupported option");
  MTN_REQUIRE((!a->out_f32 || (a->ld32 % 2 == 0 && (reinterpret_cast<uintptr_t>(a->out_f32) & 7) == 0)) &&
                  (!a->out_f16 || (a->ld16 % 2 == 0 && (reinterpret_cast<uintptr_t>(a->out_f16) & 3) == 0)),
              MTN_E_ALIGN, "rows_linear: output alignment");
  RowsLinearParams p{reinterpret_cast<const __half*>(a->A), a->lda, reinterpret_cast<const __half*>(a->W), a->ldw, a->bias,
                     a->M, a->N, a->K, a->act, a->out_f32, a->ld32, a->addend != nullptr ? 1 : 0,
                     reinterpret_cast<__half*>(a->out_f16), a->ld16, 0u};
  // contraction split: at most 8 warps, at most 8 chunks (of 32) per warp in flight; K = 512 -> 4 warps x 4 chunks,
  // K = 2048 -> 8 warps x 8 chunks; other K: the largest split that divides it
  const int chunks = a->K / 32;
  int ks = 0;
  for (int c = RL_MAX_KS; c >= 1 && ks == 0; --c)
    if (chunks % c == 0 && chunks / c <= 8) ks = c;
  if (ks == 0) ks = RL_MAX_KS + 1;
  MTN_REQUIRE(ks <= RL_MAX_KS && chunks % ks == 0, MTN_E_SHAPE, "rows_linear: K=%d has no supported split (K / 32 must factor into <= 8 warps x <= 8 chunks)", a->K);
  const int cpw = chunks / ks;
  dim3 grid(a->N / 8, (a->M + 15) / 16), block(32 * ks);
  if (prog_recording()) {
    MTN_REQUIRE(ks == RL_MAX_KS && (cpw == 2 || cpw == 8), MTN_E_SHAPE,
                "decode program: linear stage needs K = 512 or 2048 (K = %d)", a->K);
    return prog_push(PK_LINEAR, cpw, grid, &p, sizeof(p));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MTN_RL(C) MTN_CHECK_CUDA(launch_kernel(rows_linear_kernel<C>, grid, block, 0, st, p))
  switch (cpw) {
    case 1: MTN_RL(1); break;
    case 2: MTN_RL(2); break;
    case 3: MTN_RL(3); break;
    case 4: MTN_RL(4); break;
    case 5: MTN_RL(5); break;
    case 6: MTN_RL(6); break;
    case 7: MTN_RL(7); break;
    default: MTN_RL(8); break;
  }
#undef MTN_RL
  return MTN_OK;
}

extern "C" int mtn_rows_ln_linear_supported(int M, int N, int d) {
  return (M > 0 && M <= 128 && N > 0 && N % 8 == 0 && (d == 128 || d == 256 || d == 512 || d == 1024)) ? 1 : 0;
}

// out_f16[M, N] = act(LN(x)[M, d] W[N, d]^T + bias) for M <= 128: the contract of mtn_ln_linear_fwd (and bit-identical to
// mtn_layernorm_fwd + mtn_rows_linear_fwd), one launch.
extern "C" int mtn_rows_ln_linear_fwd(const float* x, int ldx, const float* a_2, const float* b_2, float eps, int M, int d,
                                      const void* W, int ldw, const float* bias, int N, int act, void* out_f16, int ld16,
                                      void* stream) {
  using namespace mtn;
  MTN_REQUIRE(x && a_2 && b_2 && W && out_f16, MTN_E_ARG, "rows_ln_linear: NULL pointer");
  MTN_REQUIRE(mtn_rows_ln_linear_supported(M, N, d), MTN_E_SHAPE, "rows_ln_linear: M=%d N=%d d=%d", M, N, d);
  MTN_REQUIRE(ldx >= d && ldx % 4 == 0 && ldw >= d && ldw % 8 == 0 && ld16 >= N && ld16 % 2 == 0 && aligned16(x) && aligned16(a_2) &&
                  aligned16(b_2) && aligned16(W) && (reinterpret_cast<uintptr_t>(out_f16) & 3) == 0,
              MTN_E_ALIGN, "rows_ln_linear: alignment / leading dimensions");
  MTN_REQUIRE(act == MTN_ACT_NONE || act == MTN_ACT_RELU, MTN_E_ARG, "rows_ln_linear: act=%d", act);
  RowsLnLinearParams p{x, ldx, a_2, b_2, eps, reinterpret_cast<const __half*>(W), ldw, bias, M, N, act,
                       reinterpret_cast<__half*>(out_f16), ld16, 0u};
  dim3 grid(N / 8, (M + 15) / 16);
  if (prog_recording()) {
    MTN_REQUIRE(d == 512, MTN_E_SHAPE, "decode program: LayerNorm + linear stage needs d = 512 (d = %d)", d);
    return prog_push(PK_LN_LINEAR, 4, grid, &p, sizeof(p));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ks = d / 32 >= 8 ? 8 : d / 32;
  if (d == 128) MTN_CHECK_CUDA(launch_kernel(rows_ln_linear_kernel<1>, grid, dim3(32 * ks), 0, st, p));
  else if (d == 256) MTN_CHECK_CUDA(launch_kernel(rows_ln_linear_kernel<2>, grid, dim3(32 * ks), 0, st, p));
  else if (d == 512) MTN_CHECK_CUDA(launch_kernel(rows_ln_linear_kernel<4>, grid, dim3(32 * ks), 0, st, p));
  else MTN_CHECK_CUDA(launch_kernel(rows_ln_linear_kernel<8>, grid, dim3(32 * ks), 0, st, p));
  return MTN_OK;
}

extern "C" int mtn_decode_attn_supported(int Lq, int d_k) { return (Lq >= 1 && Lq <= mtn::DA_MAXR && d_k == 64) ? 1 : 0; }

// Same contract as mtn_attn_core_fwd for Lq <= 8 query rows per batch element and d_k = 64.
extern "C" int mtn_decode_attn_fwd(const MtnAttnCoreArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->q && a->k && a->v && a->out, MTN_E_ARG, "decode_attn: NULL pointer");
  MTN_REQUIRE(a->B > 0 && a->h > 0 && a->Lk > 0 && mtn_decode_attn_supported(a->Lq, a->d_k), MTN_E_SHAPE,
              "decode_attn: B=%d h=%d Lq=%d Lk=%d d_k=%d (Lq <= 8, d_k = 64)", a->B, a->h, a->Lq, a->Lk, a->d_k);
  MTN_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 && aligned16(a->q) && aligned16(a->k) &&
                  aligned16(a->v) && aligned16(a->out) && a->q_batch_stride % 8 == 0 && a->k_batch_stride % 8 == 0 &&
                  a->v_batch_stride % 8 == 0 && a->o_batch_stride % 8 == 0,
              MTN_E_ALIGN, "decode_attn: leading dimensions / strides must be multiples of 8 elements, pointers 16-byte aligned");
  MTN_REQUIRE(a->mask_bits == nullptr || a->mask_rows_q == 1 || a->mask_rows_q == a->Lq, MTN_E_SHAPE,
              "decode_attn: mask_rows_q=%d must be 1 or Lq=%d", a->mask_rows_q, a->Lq);
  MTN_REQUIRE(a->stats == nullptr && a->drop_seed == nullptr, MTN_E_ARG, "decode_attn: inference only");
  auto bs = [](long long given, int L, int ld) { return given > 0 ? given : (long long)L * ld; };
  DecodeAttnParams p{reinterpret_cast<const __half*>(a->q), reinterpret_cast<const __half*>(a->k), reinterpret_cast<const __half*>(a->v),
                     reinterpret_cast<__half*>(a->out), a->ldq, a->ldk, a->ldv, a->ldo,
                     bs(a->q_batch_stride, a->Lq, a->ldq), bs(a->k_batch_stride, a->Lk, a->ldk), bs(a->v_batch_stride, a->Lk, a->ldv),
                     bs(a->o_batch_stride, a->Lq, a->ldo), a->mask_bits, a->mask_rows_q, mtn_mask_words(a->Lk), a->B, a->h, a->Lq, a->Lk,
                     1.0f / sqrtf(64.f)};
  dim3 grid(a->B * a->h), block(32 * DA_WARPS);
  if (prog_recording()) {
    MTN_REQUIRE(a->Lq == 1, MTN_E_SHAPE, "decode program: attention stage needs one query row per batch element (Lq = %d)", a->Lq);
    return prog_push(PK_ATTN, 1, dim3(grid.x, 1), &p, sizeof(p));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MTN_DA(RR) MTN_CHECK_CUDA(launch_kernel(decode_attn_kernel<RR>, grid, block, 0, st, p))
  switch (a->Lq) {
    case 1: MTN_DA(1); break;
    case 2: MTN_DA(2); break;
    case 3: MTN_DA(3); break;
    case 4: MTN_DA(4); break;
    case 5: MTN_DA(5); break;
    case 6: MTN_DA(6); break;
    case 7: MTN_DA(7); break;
    default: MTN_DA(8); break;
  }
#undef MTN_DA
  return MTN_OK;
}
