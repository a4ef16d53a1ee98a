// One feed-forward sublayer in ONE kernel, the hidden activation never leaves the SM:
//
//     x[rows, d] += relu(xn[rows, d] W1^T + b1) W2^T + b2          (xn = LayerNorm(x) as f16, d = 512, d_ff = 2048)
//
// Replaces  SublayerConnection.forward(x, PositionwiseFeedForward)  (mtn.py:125-127 around mtn.py:279-280) behind the
// LayerNorm: the two linear launches of mtn_ffn_fwd and the [rows, d_ff] f16 hidden buffer between them.
//
// Why it looks the way it does.  A [128 rows x 512] f32 output tile IS the whole tensor memory of an SM (512 columns),
// so a CTA cannot hold the output tile and a hidden-chunk accumulator at once.  Here a CTA owns 128 rows and HALF of the
// output columns (256 TMEM columns) and walks the hidden dimension in chunks of 128: chunk c of the hidden activation
// H_c[128, 128] = relu(xn W1_c^T + b1_c) is accumulated in one of two 128-column TMEM buffers, drained by the epilogue
// warps (bias, ReLU, f16) and written BACK over the first 64 columns of the same buffer as packed f16 (like the attention
// probabilities of csrc/attn.cu), and immediately consumed from tensor memory as the A operand (TS-form tcgen05.mma) of
// Y[128, 256] += H_c W2[cols, c]^T -- the hidden activation touches neither HBM nor shared memory.  The two CTAs of a row block (one per column half) each compute the full hidden
// activation: 1.5x the FLOPs of the two-GEMM form, in exchange for no cluster, no exchange and no HBM round trip of the
// hidden activation (2 x rows x d_ff x 2 bytes).  xn stays resident in shared memory (128 KB); W1 / W2 stream through a
// ring of six [128 x 64] f16 tiles (16 KB each: everything that is left); every MMA is 128 x 128 x 16.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..9 epilogue (two per TMEM lane
// quarter).  Tensor-pipe program order: G1(0), G1(1), G2(0), G1(2), G2(1), ... so the hidden chunk c+1 is being
// accumulated while chunk c is drained; the residual add leaves as TMA reduce-add stores (csrc/gemm.cu).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int FF_THREADS = 320;
constexpr int FF_ROWS = 128;    // rows per CTA (UMMA M)
constexpr int FF_D = 512;       // model width (contraction of the first GEMM)
constexpr int FF_NCOL = 256;    // output columns per CTA
constexpr int FF_HC = 128;      // hidden units per chunk
constexpr int FF_STAGES = 6;

struct FfCfg {
  static constexpr int TILE = 128 * 128;                  // one [128 x 64] f16 tile: 16 KB
  static constexpr int OFF_XN = 0;                        // 8 k-panels of xn: 128 KB
  static constexpr int OFF_RING = OFF_XN + (FF_D / 64) * TILE;
  static constexpr int OFF_BAR = OFF_RING + FF_STAGES * TILE;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

enum { FB_FULL = 0 /* +5 */, FB_EMPTY = 6 /* +5 */, FB_XN = 12, FB_D1_FULL = 13 /* +1 */, FB_HS_FULL = 15 /* +1 */, FB_G2_DONE = 17 /* +1 */, FB_Y_FULL = 19, FB_COUNT = 20 };

struct FfParams {
  int rows, d_ff;
  const float* b1;
  const float* b2;
};

__global__ void __launch_bounds__(FF_THREADS, 1)
    ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const FfParams p) {
  using C = FfCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sXN = base + C::OFF_XN, sRING = base + C::OFF_RING;
  const uint32_t bars = base + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * FB_COUNT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * FB_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * FF_NCOL;     // this CTA's output columns
  const int row0 = blockIdx.y * FF_ROWS;
  const int nch = p.d_ff / FF_HC;            // hidden chunks

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < FB_COUNT; ++i) mbar_init(bar(i), (i == FB_HS_FULL || i == FB_HS_FULL + 1) ? 256u : 1u);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tY = tmem_base;                       // [128, 256] f32
  const uint32_t tD1 = tmem_base + FF_NCOL;            // two [128, 128] f32 buffers
  pdl_wait();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(FB_XN), (FF_D / 64) * C::TILE);
      for (int kp = 0; kp < FF_D / 64; ++kp) tma_load_2d(sXN + kp * C::TILE, &tmX, bar(FB_XN), kp * 64, row0);
      uint32_t it = 0;
      auto load = [&](const CUtensorMap* m, int c0, int c1) {
        const uint32_t s = it % FF_STAGES;
        mbar_wait(bar(FB_EMPTY + s), ((it / FF_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(FB_FULL + s), C::TILE);
        tma_load_2d(sRING + s * C::TILE, m, bar(FB_FULL + s), c0, c1);
        ++it;
      };
      auto load_w1 = [&](int c) {       // W1 rows [128 c, +128), k-panels 0..7
        for (int kp = 0; kp < FF_D / 64; ++kp) load(&tmW1, kp * 64, c * FF_HC);
      };
      auto load_w2 = [&](int c) {       // W2 rows [col0 + 128 nh, +128), hidden columns [128 c + 64 kp, +64)
        for (int kp = 0; kp < FF_HC / 64; ++kp)
          for (int nh = 0; nh < FF_NCOL / 128; ++nh) load(&tmW2, c * FF_HC + kp * 64, col0 + nh * 128);
      };
      load_w1(0);
      for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) load_w1(c + 1);
        load_w2(c);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, 0);
    uint32_t it = 0;
    mbar_wait(bar(FB_XN), 0);
    auto g1 = [&](int c) {              // D1[c & 1] = xn W1_c^T
      const uint32_t d = tD1 + (uint32_t)(c & 1) * FF_HC;
      // the buffer still holds the hidden chunk c-2 as the A operand of its second GEMM: that GEMM must have COMPLETED (the
      // tensor pipe does not order an accumulator write behind an earlier MMA's read of a TMEM operand, csrc/attn.cu)
      if (c >= 2) mbar_wait(bar(FB_G2_DONE + (c & 1)), ((c - 2) >> 1) & 1);
      for (int kp = 0; kp < FF_D / 64; ++kp, ++it) {
        const uint32_t s = it % FF_STAGES;
        mbar_wait(bar(FB_FULL + s), (it / FF_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(sXN + kp * C::TILE, 16, 1024, SWZ_128B);
          const uint64_t db = make_smem_desc(sRING + s * C::TILE, 16, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16(d, da + 2 * k, db + 2 * k, idesc, (kp | k) != 0);
          tc_commit(bar(FB_EMPTY + s));
          if (kp == FF_D / 64 - 1) tc_commit(bar(FB_D1_FULL + (c & 1)));
        }
        __syncwarp();
      }
    };
    auto g2 = [&](int c) {              // Y += H_c W2[cols, c]^T, H_c from tensor memory
      mbar_wait(bar(FB_HS_FULL + (c & 1)), (c >> 1) & 1);   // the packed hidden chunk is in D1[c & 1]
      const uint32_t th = tD1 + (uint32_t)(c & 1) * FF_HC;
      for (int kp = 0; kp < FF_HC / 64; ++kp)
        for (int nh = 0; nh < FF_NCOL / 128; ++nh, ++it) {
          const uint32_t s = it % FF_STAGES;
          mbar_wait(bar(FB_FULL + s), (it / FF_STAGES) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t db = make_smem_desc(sRING + s * C::TILE, 16, 1024, SWZ_128B);
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 16 hidden units = 8 packed columns per k-step
              tc_mma_f16_ts(tY + nh * 128, th + (uint32_t)(kp * 4 + k) * 8, db + 2 * k, idesc, (c | kp | k) != 0);
            tc_commit(bar(FB_EMPTY + s));
          }
          __syncwarp();
        }
      if (lane == 0) {
        tc_commit(bar(FB_G2_DONE + (c & 1)));   // D1[c & 1] may be overwritten by the first GEMM of chunk c+2
        if (c == nch - 1) tc_commit(bar(FB_Y_FULL));
      }
      __syncwarp();
    };
    g1(0);
    for (int c = 0; c < nch; ++c) {
      if (c + 1 < nch) g1(c + 1);
      g2(c);
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps (8)
    const int ew = warp - 2;
    const int q = warp & 3;       // TMEM lane quarter
    const int half = ew >> 2;     // hidden 64-block of the chunk (drain) / output 128-column half (final epilogue)
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    // b1 of this warp's 64 hidden units of the next chunk: lane l holds units l and 32 + l, broadcast by shuffles
    float bA = __ldg(p.b1 + half * 64 + lane), bB = __ldg(p.b1 + half * 64 + 32 + lane);
    for (int c = 0; c < nch; ++c) {
      const float cA = bA, cB = bB;
      if (c + 1 < nch) {
        bA = __ldg(p.b1 + (c + 1) * FF_HC + half * 64 + lane);
        bB = __ldg(p.b1 + (c + 1) * FF_HC + half * 64 + 32 + lane);
      }
      mbar_wait(bar(FB_D1_FULL + (c & 1)), (c >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t t = tD1 + (uint32_t)(c & 1) * FF_HC + lane_off + half * 64;
      tc_ld32(t, r0);
      tc_ld32(t + 32, r1);
      tc_wait_ld();
      uint32_t pk[32];   // 64 hidden units of this row as f16 pairs
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float b0 = __shfl_sync(0xffffffffu, cA, 2 * i), b1v = __shfl_sync(0xffffffffu, cA, 2 * i + 1);
        pk[i] = pack_f16x2_sat(fmaxf(__uint_as_float(r0[2 * i]) + b0, 0.f), fmaxf(__uint_as_float(r0[2 * i + 1]) + b1v, 0.f));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float b0 = __shfl_sync(0xffffffffu, cB, 2 * i), b1v = __shfl_sync(0xffffffffu, cB, 2 * i + 1);
        pk[16 + i] = pack_f16x2_sat(fmaxf(__uint_as_float(r1[2 * i]) + b0, 0.f), fmaxf(__uint_as_float(r1[2 * i + 1]) + b1v, 0.f));
      }
      // Both warps of this lane quarter have read their f32 columns before either overwrites them: the packed hidden
      // units of warp `half` go to columns [32 half, 32 half + 32) of the buffer, inside the f32 columns of warp 0.
      named_bar_sync(1 + q, 64);
      tc_st32(tD1 + (uint32_t)(c & 1) * FF_HC + lane_off + half * 32, pk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar(FB_HS_FULL + (c & 1)));
    }
    // ---- final epilogue: x[rows, col0 + 128 half ..] += Y + b2, as TMA reduce-add stores from the (dead) xn region
    float b2v[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) b2v[cc] = __ldg(p.b2 + col0 + half * 128 + cc * 32 + lane);
    mbar_wait(bar(FB_Y_FULL), 0);
    tc_fence_after();
    const uint32_t tiles = sXN + (uint32_t)ew * (4u * 4096u);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const uint32_t tile = tiles + (uint32_t)cc * 4096u;
      const float bmine = cc == 0 ? b2v[0] : (cc == 1 ? b2v[1] : (cc == 2 ? b2v[2] : b2v[3]));
      uint32_t acc[32];
      tc_ld32(tY + lane_off + half * 128 + cc * 32, acc);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v0 = __uint_as_float(acc[4 * j]) + __shfl_sync(0xffffffffu, bmine, 4 * j);
        const float v1 = __uint_as_float(acc[4 * j + 1]) + __shfl_sync(0xffffffffu, bmine, 4 * j + 1);
        const float v2 = __uint_as_float(acc[4 * j + 2]) + __shfl_sync(0xffffffffu, bmine, 4 * j + 2);
        const float v3 = __uint_as_float(acc[4 * j + 3]) + __shfl_sync(0xffffffffu, bmine, 4 * j + 3);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + lane * 128 + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4)),
                     "f"(v0), "f"(v1), "f"(v2), "f"(v3)
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmOut, tile, col0 + half * 128 + cc * 32, row0 + q * 32, 0);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();   // (the kernel boundary completes the writes)
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mtn

extern "C" int mtn_ffn_fused_supported(int rows, int d, int d_ff) {
  return (rows > 0 && d == mtn::FF_D && d_ff > 0 && d_ff % mtn::FF_HC == 0) ? 1 : 0;
}

// x += relu(xn W1^T + b1) W2^T + b2.   xn: [rows, ld_xn] f16 (LayerNorm(x), mtn_layernorm_fwd), x: [rows, ld_x] f32
// updated in place, w_1: [d_ff, d] f16, w_2: [d, d_ff] f16 (the reference's nn.Linear layouts).  d = 512.
extern "C" int mtn_ffn_fused_fwd(const void* xn_f16, int ld_xn, float* x, int ld_x, int rows, int d, int d_ff, const void* w_1,
                                 const float* b_1, const void* w_2, const float* b_2, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(xn_f16 && x && w_1 && b_1 && w_2 && b_2, MTN_E_ARG, "ffn_fused: NULL pointer");
  MTN_REQUIRE(mtn_ffn_fused_supported(rows, d, d_ff), MTN_E_SHAPE, "ffn_fused: rows=%d d=%d d_ff=%d (d = 512, d_ff %% 128 == 0)", rows,
              d, d_ff);
  MTN_REQUIRE(ld_xn >= d && ld_xn % 8 == 0 && ld_x >= d && ld_x % 4 == 0 && aligned16(xn_f16) && aligned16(x) && aligned16(w_1) &&
                  aligned16(w_2) && aligned16(b_1) && aligned16(b_2),
              MTN_E_ALIGN, "ffn_fused: leading dimensions / alignment");
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FfCfg::TOTAL));
    attr_set = true;
  }
  CUtensorMap tx, tw1, tw2, to;
  int rc = make_tmap_2d_f16(&tx, xn_f16, d, rows, ld_xn, 64, FF_ROWS, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&tw1, w_1, d, d_ff, d, 64, 128, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&tw2, w_2, d_ff, d, d_ff, 64, 128, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_3d_f32(&to, x, d, rows, 1, ld_x, (uint64_t)rows * ld_x, 32, 32, TM_SWZ_128);
  if (rc) return rc;
  FfParams p{rows, d_ff, b_1, b_2};
  dim3 grid(d / FF_NCOL, (rows + FF_ROWS - 1) / FF_ROWS);
  MTN_CHECK_CUDA(launch_kernel(ffn_fused_kernel, grid, dim3(FF_THREADS), FfCfg::TOTAL, static_cast<cudaStream_t>(stream), tx, tw1,
                               tw2, to, p));
  return MTN_OK;
}
