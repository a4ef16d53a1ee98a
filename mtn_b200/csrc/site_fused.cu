// One attention SITE in one kernel (hoisted K/V):
//     x += Wo . concat_h softmax(mask((xn Wq_h^T + bq_h) K_h^T / sqrt(d_k))) V_h + bo
// i.e. SublayerConnection.forward (mtn.py:125-127) around MultiHeadedAttention.forward (mtn.py:248-267) around
// attention() (mtn.py:221-231), for a site whose K / V projections already exist (static memories: their K/V for all
// layers come out of one hoisted GEMM; engine.py).  xn = LayerNorm(x) in f16 is the input (mtn_layernorm_fwd).
//
// Work decomposition: one CLUSTER OF TWO CTAs per (batch element, 128-query tile).  CTA `rank` owns half of the heads
// (HH = h / 2) and half of the output columns (NH = HH * 64 = d / 2); everything between the LayerNorm output and the
// residual stream stays on chip:
//   phase 1  Q[128, NH] = xn[128, d] Wq[rank half]^T on the tensor core (TMA ring, tcgen05.mma into TMEM), drained
//            (+ bias, -> f16) into shared memory as HH K-major 128B-swizzled [128 x 64] head tiles -- the A operand
//            layout of Q K^T.  The Q projection never goes to HBM.
//   phase 2  per head: S = Q_h K_h^T, online softmax (f32, log2 domain, the reference's FINITE -1e9 mask value), P V
//            -- the pipeline of csrc/attn.cu (K/V by 3-D TMA out of the hoisted buffer, two S accumulators, P through
//            shared memory).  O_h / l overwrites the head's Q tile (f16): after the last head the CTA holds its half
//            of concat_h O as the A operand of the output projection.  Each finished tile is also pushed into the
//            PEER CTA's shared memory by one bulk async copy (cp.async.bulk shared::cta -> shared::cluster) that
//            completes on the peer's mbarrier.
//   phase 3  Y[128, NH] = O[128, d] Wo[rank half]^T: A = the 2 HH head tiles (own + received), B = Wo rows by TMA;
//            epilogue: + bias, transposed through shared memory, added to the f32 residual stream with coalesced
//            red.global.add.v4.f32 (one f32 add per element; the CTA pair covers every column exactly once).
// Versus the launch sequence LayerNorm -> Q GEMM -> core -> out-proj GEMM this removes two launches and the HBM/L2 round
// trips of Q and O per site.
//
// Shared memory (HH = 4, 96-key tiles): Q/O tiles 64 KB + peer tiles 64 KB + a 96 KB region that is, in turn, the
// phase-1 operand ring (2 x 48 KB), the K/V/P buffers of phase 2 (72 KB), the phase-3 Wo ring (3 x 32 KB) and the
// epilogue's transpose tiles.  TMEM: 512 columns = Q / Y accumulator (NH) + two S buffers + O.  One CTA per SM.
//
// Warp roles (224 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 softmax / drains / epilogue
// (TMEM lane quarter = warp % 4), warp 6 sends finished O tiles to the peer.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int SF_THREADS = 224;
constexpr int SF_QT = 128;

template <int HH, int KT>
struct SfCfg {
  static constexpr int DK = 64;
  static constexpr int NH = HH * DK;        // columns of Q / of the output owned by one CTA
  static constexpr int D = 2 * NH;          // model width
  static constexpr int NKB = D / 64;        // 64-wide k-blocks of the two projections
  static constexpr int TILE = SF_QT * 128;  // one [128 x 64] f16 tile, K-major, 128B swizzle: 16 KB
  static constexpr int OFF_QO = 0;
  static constexpr int OFF_PEER = OFF_QO + HH * TILE;
  static constexpr int OFF_RING = OFF_PEER + HH * TILE;
  static constexpr int A_BYTES = TILE;      // xn k-block [128 x 64]
  static constexpr int B_BYTES = NH * 128;  // weight k-block [NH x 64]
  static constexpr int ST1 = 2, ST3 = 3;
  static constexpr int RING1 = ST1 * (A_BYTES + B_BYTES);
  static constexpr int RING3 = ST3 * B_BYTES;
  static constexpr int KV_BYTES = KT * 128;
  static constexpr int P_BYTES = (KT / 64) * SF_QT * 128 + ((KT % 64) ? SF_QT * 64 : 0);
  static constexpr int OFF_K = 0, OFF_V = 2 * KV_BYTES, OFF_P = (4 * KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int RING2 = OFF_P + P_BYTES;
  static constexpr int XPOSE = 4 * 4096;
  static constexpr int RING_A = RING1 > RING2 ? RING1 : RING2;
  static constexpr int RING_B = RING3 > XPOSE ? RING3 : XPOSE;
  static constexpr int RING = RING_A > RING_B ? RING_A : RING_B;
  static constexpr int OFF_BAR = OFF_RING + RING;
  static constexpr int TOTAL = OFF_BAR + 512 + 1024;  // barriers + 1 KB alignment slack
  static constexpr uint32_t D_COL = 0, S_COL = NH, O_COL = NH + 2 * KT;
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(NH + 2 * KT + DK <= 512, "TMEM budget");
  static_assert(KV_BYTES % 1024 == 0, "K/V buffers must stay swizzle-atom aligned");
  static_assert(NH <= 256 && NH % 16 == 0, "UMMA N");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

enum {
  SB_R1_FULL = 0 /* +1 */, SB_R1_EMPTY = 2 /* +1 */, SB_D1_FULL = 4, SB_Q_READY = 5,
  SB_K_FULL = 6 /* +1 */, SB_K_EMPTY = 8 /* +1 */, SB_V_FULL = 10 /* +1 */, SB_V_EMPTY = 12 /* +1 */,
  SB_S_FULL = 14 /* +1 */, SB_S_FREE = 16 /* +1 */, SB_P_FULL = 18, SB_PV_DONE = 19, SB_ATT_DONE = 20,
  SB_OWN_O = 21 /* +3 */, SB_PEER_O = 25, SB_R3_FULL = 26 /* +2 */, SB_R3_EMPTY = 29 /* +2 */, SB_D3_FULL = 32,
  SB_COUNT = 33
};

struct SfParams {
  int B, h, Lq, Lk, nqt;
  const float* b_q;
  const float* b_o;
  float* x;
  int ld_x;
  const uint32_t* mask_bits;
  int mask_rows_q, mask_words;
  float scale;
};

__device__ __forceinline__ float sf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t sf_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// bulk async copy local shared memory -> a peer CTA's shared memory; completes `bytes` on the PEER's mbarrier
__device__ __forceinline__ void sf_dsmem_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}

template <int HH, int KT>
__global__ void __launch_bounds__(SF_THREADS, 1)
    attn_site_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                           const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                           const __grid_constant__ CUtensorMap tmWo, const SfParams p) {
  using C = SfCfg<HH, KT>;
  constexpr int DK = C::DK;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sQO = base + C::OFF_QO, sPEER = base + C::OFF_PEER, sRING = base + C::OFF_RING;
  const uint32_t sK = sRING + C::OFF_K, sV = sRING + C::OFF_V, sP = sRING + C::OFF_P;
  const uint32_t bars = base + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * SB_COUNT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * SB_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int item = blockIdx.x >> 1;  // (batch element, query tile) of this cluster
  const int qt = item % p.nqt, b = item / p.nqt;
  const int nt = (p.Lk + KT - 1) / KT;
  const uint32_t total = (uint32_t)HH * (uint32_t)nt;  // key tiles of this CTA, all heads

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmWo);
    for (int i = 0; i < SB_COUNT; ++i) {
      const bool all128 = (i == SB_Q_READY || i == SB_S_FREE || i == SB_S_FREE + 1 || i == SB_P_FULL ||
                           (i >= SB_OWN_O && i < SB_OWN_O + 4));
      mbar_init(bar(i), all128 ? 128u : 1u);
    }
    // the peer's HH finished head tiles land here (armed before the cluster barrier below, i.e. before any send)
    mbar_arrive_expect_tx(bar(SB_PEER_O), (uint32_t)(HH * C::TILE));
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers exist before either can signal the other
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tD = tmem_base + C::D_COL, tO = tmem_base + C::O_COL;
  pdl_wait();

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      // ---- phase 1: xn k-blocks + Wq k-blocks
      for (int kb = 0; kb < C::NKB; ++kb) {
        const int s = kb % C::ST1;
        mbar_wait(bar(SB_R1_EMPTY + s), ((kb / C::ST1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(SB_R1_FULL + s), C::A_BYTES + C::B_BYTES);
        const uint32_t st = sRING + s * (C::A_BYTES + C::B_BYTES);
        tma_load_3d(st, &tmX, bar(SB_R1_FULL + s), kb * 64, qt * SF_QT, b);
        tma_load_2d(st + C::A_BYTES, &tmWq, bar(SB_R1_FULL + s), kb * 64, (int)rank * C::NH);
      }
      // ---- phase 2: K / V tiles of my heads (the ring region is free once the projection MMAs have retired)
      mbar_wait(bar(SB_D1_FULL), 0);
      auto load_k = [&](uint32_t g) {
        const uint32_t hh = g / nt, j = g % nt, kb = g & 1, kph = (g >> 1) & 1;
        mbar_wait(bar(SB_K_EMPTY + kb), kph ^ 1);
        mbar_arrive_expect_tx(bar(SB_K_FULL + kb), C::KV_BYTES);
        tma_load_3d(sK + kb * C::KV_BYTES, &tmK, bar(SB_K_FULL + kb), ((int)rank * HH + (int)hh) * DK, j * KT, b);
      };
      auto load_v = [&](uint32_t g) {
        const uint32_t hh = g / nt, j = g % nt, vb = g & 1, vph = (g >> 1) & 1;
        mbar_wait(bar(SB_V_EMPTY + vb), vph ^ 1);
        mbar_arrive_expect_tx(bar(SB_V_FULL + vb), C::KV_BYTES);
        tma_load_3d(sV + vb * C::KV_BYTES, &tmV, bar(SB_V_FULL + vb), ((int)rank * HH + (int)hh) * DK, j * KT, b);
      };
      load_k(0);
      for (uint32_t g = 0; g < total; ++g) {
        if (g + 1 < total) load_k(g + 1);
        load_v(g);
      }
      // ---- phase 3: Wo k-blocks (the region is free once the last P V has retired)
      mbar_wait(bar(SB_ATT_DONE), 0);
      for (int kc = 0; kc < C::NKB; ++kc) {
        const int s = kc % C::ST3;
        mbar_wait(bar(SB_R3_EMPTY + s), ((kc / C::ST3) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(SB_R3_FULL + s), C::B_BYTES);
        tma_load_2d(sRING + s * C::B_BYTES, &tmWo, bar(SB_R3_FULL + s), kc * 64, (int)rank * C::NH);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    constexpr uint32_t idesc_p = make_idesc_f16(SF_QT, C::NH, 0, 0);  // both projections: K-major x K-major
    constexpr uint32_t idesc_s = make_idesc_f16(SF_QT, KT, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(SF_QT, DK, 0, 1);     // V is MN-major
    // ---- phase 1
    for (int kb = 0; kb < C::NKB; ++kb) {
      const int s = kb % C::ST1;
      mbar_wait(bar(SB_R1_FULL + s), (kb / C::ST1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = sRING + s * (C::A_BYTES + C::B_BYTES);
        const uint64_t da = make_smem_desc(st, 16, 1024, SWZ_128B);
        const uint64_t db = make_smem_desc(st + C::A_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tD, da + 2 * k, db + 2 * k, idesc_p, (kb | k) != 0);
        tc_commit(bar(SB_R1_EMPTY + s));
        if (kb == C::NKB - 1) tc_commit(bar(SB_D1_FULL));
      }
      __syncwarp();
    }
    // ---- phase 2
    mbar_wait(bar(SB_Q_READY), 0);  // the Q head tiles are in shared memory
    tc_fence_after();
    auto issue_qk = [&](uint32_t g) {
      const uint32_t hh = g / nt, sb = g & 1, ph2 = (g >> 1) & 1;
      mbar_wait(bar(SB_K_FULL + sb), ph2);
      mbar_wait(bar(SB_S_FREE + sb), ph2 ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dq = make_smem_desc(sQO + hh * C::TILE, 16, 1024, SWZ_128B);
        const uint64_t dk = make_smem_desc(sK + sb * C::KV_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k)
          tc_mma_f16(tmem_base + C::S_COL + sb * KT, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(bar(SB_K_EMPTY + sb));
        tc_commit(bar(SB_S_FULL + sb));
      }
      __syncwarp();
    };
    issue_qk(0);
    for (uint32_t g = 0; g < total; ++g) {
      if (g + 1 < total) issue_qk(g + 1);
      const uint32_t ph = g & 1, j = g % nt;
      mbar_wait(bar(SB_V_FULL + ph), (g >> 1) & 1);
      mbar_wait(bar(SB_P_FULL), ph);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const bool rem = (KT % 64) != 0 && kk >= (KT / 64) * 4;
          const uint64_t dp = rem ? make_smem_desc(sP + (KT / 64) * (SF_QT * 128) + (kk & 3) * 32, 16, 512, SWZ_64B)
                                  : make_smem_desc(sP + (kk >> 2) * (SF_QT * 128) + (kk & 3) * 32, 16, 1024, SWZ_128B);
          const uint64_t dv = make_smem_desc(sV + ph * C::KV_BYTES + kk * 16 * 128, KT * 128, 1024, SWZ_128B);
          tc_mma_f16(tO, dp, dv, idesc_o, (j | (uint32_t)kk) != 0);
        }
        tc_commit(bar(SB_V_EMPTY + ph));
        tc_commit(bar(SB_PV_DONE));
        if (g == total - 1) tc_commit(bar(SB_ATT_DONE));
      }
      __syncwarp();
    }
    // ---- phase 3: Y = [O own | O peer] Wo^T, head tiles in global head order
    for (int hh = 0; hh < HH; ++hh) mbar_wait(bar(SB_OWN_O + hh), 0);
    mbar_wait(bar(SB_PEER_O), 0);
    tc_fence_after();
    for (int kc = 0; kc < C::NKB; ++kc) {
      const int s = kc % C::ST3;
      mbar_wait(bar(SB_R3_FULL + s), (kc / C::ST3) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_tile = ((uint32_t)(kc / HH) == rank ? sQO : sPEER) + (uint32_t)(kc % HH) * C::TILE;
        const uint64_t da = make_smem_desc(a_tile, 16, 1024, SWZ_128B);
        const uint64_t db = make_smem_desc(sRING + s * C::B_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tD, da + 2 * k, db + 2 * k, idesc_p, (kc | k) != 0);
        tc_commit(bar(SB_R3_EMPTY + s));
        if (kc == C::NKB - 1) tc_commit(bar(SB_D3_FULL));
      }
      __syncwarp();
    }
  } else if (warp == 6) {
    // ================================================================ DSMEM sender: my finished head tiles -> the peer
    if (lane == 0) {
      const uint32_t peer = rank ^ 1u;
      const uint32_t peer_bar = sf_mapa(bar(SB_PEER_O), peer);
      for (int hh = 0; hh < HH; ++hh) {
        mbar_wait(bar(SB_OWN_O + hh), 0);  // all 128 rows written and fenced for the async proxy
        sf_dsmem_copy(sf_mapa(sPEER + hh * C::TILE, peer), sQO + hh * C::TILE, (uint32_t)C::TILE, peer_bar);
      }
    }
    __syncwarp();
  } else {
    // ================================================================ softmax warps: Q drain, softmax, O drain, epilogue
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    constexpr float LOG2E = 1.4426950408889634f;
    const float c1 = p.scale * LOG2E;
    const float t_masked = -1e9f * LOG2E;
    constexpr int NCH = KT / 32;

    // ---- phase 1 drain: Q = acc + bias -> f16 -> HH K-major swizzled head tiles
    mbar_wait(bar(SB_D1_FULL), 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < C::NH / 32; ++c) {
      uint32_t r[32];
      tc_ld32(tD + lane_off + c * 32, r);
      const float* bq = p.b_q + rank * C::NH + c * 32;
      float4 bb[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) bb[t] = __ldg(reinterpret_cast<const float4*>(bq) + t);
      tc_wait_ld();
      const uint32_t tile = sQO + (uint32_t)(c >> 1) * C::TILE + row * 128;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t chunk = ((uint32_t)((c & 1) * 4 + t)) ^ sw;
        const float4 b0 = bb[2 * t], b1 = bb[2 * t + 1];
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + chunk * 16),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t]) + b0.x, __uint_as_float(r[8 * t + 1]) + b0.y)),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 2]) + b0.z, __uint_as_float(r[8 * t + 3]) + b0.w)),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 4]) + b1.x, __uint_as_float(r[8 * t + 5]) + b1.y)),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 6]) + b1.z, __uint_as_float(r[8 * t + 7]) + b1.w))
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar(SB_Q_READY));

    // ---- phase 2: softmax over the key tiles of each head (the per-tile code of csrc/attn.cu)
    const bool live = qt * SF_QT + q4 * 32 < p.Lq;  // some query row of this warp exists
    const uint32_t* mrow = nullptr;
    if (p.mask_bits != nullptr) {
      const int mq = (p.mask_rows_q == 1) ? 0 : min(qt * SF_QT + row, p.Lq - 1);
      mrow = p.mask_bits + ((size_t)b * p.mask_rows_q + mq) * p.mask_words;
    }
    uint32_t mw_pref[NCH];
    auto fetch_mask = [&](int jn) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int k0 = jn * KT + c * 32;
        mw_pref[c] = (mrow != nullptr && k0 < p.Lk) ? __ldcg(mrow + (k0 >> 5)) : 0xffffffffu;
      }
    };
    fetch_mask(0);
    uint32_t g = 0;
    for (int hh = 0; hh < HH; ++hh) {
      float m_run = -CUDART_INF_F, l_run = 0.f;
      if (!live) {
        // every query row of this warp is padding: keep the barrier protocol alive, skip the arithmetic (these rows of
        // P / O hold whatever shared / tensor memory holds; rows are independent and never stored)
        for (int j = 0; j < nt; ++j, ++g) {
          const uint32_t ph = g & 1;
          mbar_wait(bar(SB_S_FULL + ph), (g >> 1) & 1);
          if (j > 0) mbar_wait(bar(SB_PV_DONE), ph ^ 1);
          mbar_arrive(bar(SB_S_FREE + ph));
          mbar_arrive(bar(SB_P_FULL));
        }
        mbar_wait(bar(SB_PV_DONE), (g - 1) & 1);
        mbar_arrive(bar(SB_OWN_O + hh));
        continue;
      }
      for (int j = 0; j < nt; ++j, ++g) {
        const uint32_t ph = g & 1;
        const uint32_t tS = tmem_base + C::S_COL + ph * KT;
        uint32_t mwv[NCH];
        bool plain = (j + 1) * KT <= p.Lk;
        {
          uint32_t w = 0xffffffffu;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            mwv[c] = mw_pref[c];
            w &= mwv[c];
          }
          if (mrow != nullptr) plain = plain && __all_sync(0xffffffffu, w == 0xffffffffu);
        }
        fetch_mask(j + 1 < nt ? j + 1 : 0);  // the mask does not depend on the head
        mbar_wait(bar(SB_S_FULL + ph), (g >> 1) & 1);
        tc_fence_after();
        auto store_chunk = [&](int c, const uint32_t(&e)[32]) {
          const bool rem = (KT % 64) != 0 && c == NCH - 1;
          const uint32_t panel = sP + (c >> 1) * (SF_QT * 128) + row * (rem ? 64 : 128);
          const uint32_t x = rem ? ((uint32_t)(row >> 1) & 3u) : sw, c4 = rem ? 0u : (uint32_t)(c & 1) * 4u;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(panel + ((c4 + t) ^ x) * 16),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t]), __uint_as_float(e[8 * t + 1]))),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t + 2]), __uint_as_float(e[8 * t + 3]))),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t + 4]), __uint_as_float(e[8 * t + 5]))),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t + 6]), __uint_as_float(e[8 * t + 7])))
                         : "memory");
          }
        };
        // the P buffer and the O accumulator are in use by P V of the previous tile until it retires (across heads the
        // previous head's epilogue has already waited for its last P V)
        auto wait_p_buffer = [&]() {
          if (j > 0) {
            mbar_wait(bar(SB_PV_DONE), ph ^ 1);
            tc_fence_after();
          }
        };
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        float m_new;
        bool rescale = false;
        auto pick_max = [&](float m_tile) {
          m_new = fmaxf(m_run, m_tile);
          if (j > 0) {
            rescale = __any_sync(0xffffffffu, m_new - m_run > 8.f);
            if (!rescale) m_new = m_run;
          }
        };
        if (plain) {
          uint32_t r[NCH][32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc_ld32(tS + lane_off + c * 32, r[c]);
          tc_wait_ld();
          tc_fence_before();
          mbar_arrive(bar(SB_S_FREE + ph));
          float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r[c][i]));
          }
          pick_max(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c1);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float e = sf_ex2(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
              l4[i & 3] += e;
              r[c][i] = __float_as_uint(e);
            }
          }
          wait_p_buffer();
#pragma unroll
          for (int c = 0; c < NCH; ++c) store_chunk(c, r[c]);
        } else {
          float m_tile = -CUDART_INF_F;
          auto chunk_mask = [&](int c) { return c == 0 ? mwv[0] : (c == 1 ? mwv[1] : mwv[NCH - 1]); };
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int nvalid = p.Lk - (j * KT + c * 32);
            if (nvalid <= 0) break;
            const uint32_t inb = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
            const uint32_t mw = chunk_mask(c);
            if (__all_sync(0xffffffffu, (mw & inb) == 0u)) {
              m_tile = fmaxf(m_tile, t_masked);
              continue;
            }
            uint32_t r[32];
            tc_ld32(tS + lane_off + c * 32, r);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float t = __uint_as_float(r[i]) * c1;
              t = ((mw >> i) & 1u) ? t : t_masked;
              t = ((inb >> i) & 1u) ? t : -CUDART_INF_F;
              m_tile = fmaxf(m_tile, t);
            }
          }
          pick_max(m_tile);
          wait_p_buffer();
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int nvalid = p.Lk - (j * KT + c * 32);
            const uint32_t inb = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
            const uint32_t mw = chunk_mask(c);
            uint32_t e[32];
            if (nvalid > 0 && __all_sync(0xffffffffu, (mw & inb) == 0u)) {
              const float pm = sf_ex2(t_masked - m_new);
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = ((inb >> i) & 1u) ? __float_as_uint(pm) : 0u;
              l4[0] += pm * (float)__popc(inb);
            } else if (nvalid > 0) {
              tc_ld32(tS + lane_off + c * 32, e);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float t = __uint_as_float(e[i]) * c1;
                t = ((mw >> i) & 1u) ? t : t_masked;
                t = ((inb >> i) & 1u) ? t : -CUDART_INF_F;
                const float x = sf_ex2(t - m_new);
                l4[i & 3] += x;
                e[i] = __float_as_uint(x);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = 0u;
            }
            store_chunk(c, e);
          }
          tc_fence_before();
          mbar_arrive(bar(SB_S_FREE + ph));
        }
        const float alpha = sf_ex2(m_run - m_new);
        l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
        m_run = m_new;
        if (rescale) {
#pragma unroll
          for (int c = 0; c < DK / 32; ++c) {
            uint32_t o[32];
            tc_ld32(tO + lane_off + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc_st32(tO + lane_off + c * 32, o);
          }
          tc_wait_st();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar(SB_P_FULL));
      }
      // ---- head epilogue: O / l -> f16 -> the head's (now idle) Q tile: the A operand of the output projection
      mbar_wait(bar(SB_PV_DONE), (g - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const uint32_t tile = sQO + (uint32_t)hh * C::TILE + row * 128;
#pragma unroll
      for (int c = 0; c < DK / 32; ++c) {
        uint32_t r[32];
        tc_ld32(tO + lane_off + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t chunk = (uint32_t)(c * 4 + t) ^ sw;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + chunk * 16),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t]) * inv_l, __uint_as_float(r[8 * t + 1]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 2]) * inv_l, __uint_as_float(r[8 * t + 3]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 4]) * inv_l, __uint_as_float(r[8 * t + 5]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 6]) * inv_l, __uint_as_float(r[8 * t + 7]) * inv_l))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(SB_OWN_O + hh));
    }

    // ---- phase 3 epilogue: Y + bias, transposed through shared memory, added to the residual stream (coalesced red.add)
    mbar_wait(bar(SB_D3_FULL), 0);
    tc_fence_after();
    float4* xp = reinterpret_cast<float4*>(smem + C::OFF_RING + (warp - 2) * 4096);
    const int sub_r = lane >> 2, c8 = lane & 3;
    const int wr_base = lane * 8, wr_sw = lane & 7;
    const int rd0 = sub_r * 8 + ((2 * c8) ^ sub_r), rd1 = sub_r * 8 + ((2 * c8 + 1) ^ sub_r);
    const int rows_valid = min(SF_QT, p.Lq - qt * SF_QT);
    const size_t grow0 = (size_t)b * p.Lq + (size_t)qt * SF_QT;
#pragma unroll 1
    for (int c = 0; c < C::NH / 32; ++c) {
      const int col = (int)rank * C::NH + c * 32 + c8 * 8;
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.b_o + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.b_o + col + 4));
      uint32_t acc[32];
      tc_ld32(tD + lane_off + c * 32, acc);
      tc_wait_ld();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        xp[wr_base + (jj ^ wr_sw)] = make_float4(__uint_as_float(acc[4 * jj]), __uint_as_float(acc[4 * jj + 1]),
                                                 __uint_as_float(acc[4 * jj + 2]), __uint_as_float(acc[4 * jj + 3]));
      __syncwarp();
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = xp[i * 64 + rd0];
        v[2 * i + 1] = xp[i * 64 + rd1];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = q4 * 32 + sub_r + i * 8;
        if (rl < rows_valid) {
          float* o = p.x + (grow0 + rl) * p.ld_x + col;
          grad_red_v4(o, v[2 * i].x + b0.x, v[2 * i].y + b0.y, v[2 * i].z + b0.z, v[2 * i].w + b0.w, 0);
          grad_red_v4(o + 4, v[2 * i + 1].x + b1.x, v[2 * i + 1].y + b1.y, v[2 * i + 1].z + b1.z, v[2 * i + 1].w + b1.w, 0);
        }
      }
    }
    tc_fence_before();
  }
  // nobody leaves while the peer may still copy into / out of its shared memory or signal its barriers
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ================================================================================================================
// Debug instrument (tools/site_phases.py): when a stamp buffer is registered, a few role-leader threads of the v2 kernel
// record clock64() at phase boundaries: stamps[blockIdx.x * 16 + i].  NULL (the default) costs one load per CTA.
__device__ unsigned long long* g_sf_stamps = nullptr;
#define SF_STAMP(i) do { if (stamps != nullptr) stamps[blockIdx.x * 16 + (i)] = (unsigned long long)clock64(); } while (0)

// v2: TWO attention engines per CTA.  v1 runs the HH heads of a CTA one after the other on one softmax warpgroup; the
// attention phase is a per-row latency chain (TMEM load -> max -> exp -> P store -> P V -> O drain), so the tensor
// pipe idles and phase 2 is ~40 % of the kernel.  Here the CTA carries two complete engines -- K/V producer warp, MMA
// warp, four softmax warps, own barriers, own K/V/P buffers (64-key tiles: 48 KB each, together the 96 KB ring region),
// own S/O columns of tensor memory -- that work on different heads at the same time; the two softmax warpgroups also
// split the Q drain and the epilogue's columns.  416 threads (13 warps): 0 producer A, 1 MMA A, 2..5 softmax 0,
// 6 sender, 7 producer B, 8 MMA B, 9..12 softmax 1.  Everything else is v1.
// P goes to the tensor core THROUGH TENSOR MEMORY as in csrc/attn.cu (packed f16 over the first KT/2 columns of the
// tile's S buffer, TS-form tcgen05.mma): the engines keep no P panel in shared memory, and the 32 KB that frees hold
// the first W_o stage of phase 3, requested while the attention phase still runs.
// ================================================================================================================
constexpr int SF2_THREADS = 416;

template <int HH>
struct Sf2Cfg {
  static constexpr int DK = 64, KT = 64;
  static constexpr int NH = HH * DK;
  static constexpr int D = 2 * NH;
  static constexpr int NKB = D / 64;
  static constexpr int TILE = SF_QT * 128;
  static constexpr int OFF_QO = 0;
  static constexpr int OFF_PEER = OFF_QO + HH * TILE;
  static constexpr int OFF_RING = OFF_PEER + HH * TILE;
  static constexpr int A_BYTES = TILE;
  static constexpr int B_BYTES = NH * 128;
  // phase 1: two stages in the ring region plus one in the (still unused) Q/O tile region -- the projection is bound by
  // the latency of its operand loads, not by their bandwidth
  static constexpr int ST1 = 3, ST3 = 3;
  static constexpr int RING1 = (ST1 - 1) * (A_BYTES + B_BYTES);
  static_assert(A_BYTES + B_BYTES <= HH * TILE, "third phase-1 stage lives in the Q/O tile region");
  static constexpr int RING3 = ST3 * B_BYTES;
  static constexpr int KV_BYTES = KT * 128;            // 8 KB
  static constexpr int ENG_BYTES = 4 * KV_BYTES;       // K0 K1 V0 V1 = 32 KB (P lives in tensor memory)
  static constexpr int RING2 = 2 * ENG_BYTES;          // the engines use ring bytes [0, 64 KB)
  // W_o stage ST3-1 of phase 3 sits behind the engines' buffers, so its first k-block can be requested while the
  // attention phase still runs; the other stages start at the ring base
  static constexpr int OFF_WO_LAST = RING2 > (ST3 - 1) * B_BYTES ? RING2 : (ST3 - 1) * B_BYTES;
  static constexpr int XPOSE = 0;           // (the epilogue stages in the O tile regions)
  static constexpr int RING_A = RING1 > RING2 ? RING1 : RING2;
  static constexpr int RING3B = OFF_WO_LAST + B_BYTES;
  static constexpr int RING_B0 = RING3 > XPOSE ? RING3 : XPOSE;
  static constexpr int RING_B = RING_B0 > RING3B ? RING_B0 : RING3B;
  static constexpr int RING = RING_A > RING_B ? RING_A : RING_B;
  static constexpr int OFF_BAR = OFF_RING + RING;
  static constexpr int OFF_BIAS = OFF_BAR + 448;        // b_q of this CTA's NH columns (f32)
  static constexpr int TOTAL = OFF_BIAS + NH * 4 + 1024;
  static constexpr uint32_t ENG_COLS = 2 * KT + DK;    // S0 S1 O = 192 columns per engine
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(2 * ENG_COLS <= 512 && NH <= 256, "TMEM budget");
  static_assert(TOTAL <= 232448, "shared memory budget");
  static_assert(HH % 2 == 0, "heads are split over two engines");
};

enum {
  S2_R1_FULL = 0 /* +2 */, S2_R1_EMPTY = 3 /* +2 */, S2_D1_FULL = 6, S2_Q_READY = 7, S2_OWN_O = 8 /* +3 */, S2_PEER_O = 12,
  S2_R3_FULL = 13 /* +2 */, S2_R3_EMPTY = 16 /* +2 */, S2_D3_FULL = 19,
  S2_ENG = 20,   // per engine (+E_COUNT each): K_FULL 0,1  K_EMPTY 2,3  V_FULL 4,5  V_EMPTY 6,7  S_FULL 8,9  P_FULL 10,11  PV_DONE 12,13  ATT_DONE 14
  S2_COUNT = 20 + 30
};
static_assert(8 * S2_COUNT + 8 <= 448, "barrier area");
enum { E_K_FULL = 0, E_K_EMPTY = 2, E_V_FULL = 4, E_V_EMPTY = 6, E_S_FULL = 8, E_P_FULL = 10, E_PV_DONE = 12, E_ATT_DONE = 14, E_COUNT = 15 };

template <int HH>
__global__ void __launch_bounds__(SF2_THREADS, 1)
    attn_site_fused2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                            const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                            const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmXo,
                            const SfParams p) {
  using C = Sf2Cfg<HH>;
  constexpr int DK = C::DK, KT = C::KT, HE = HH / 2;   // heads per engine
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sQO = base + C::OFF_QO, sPEER = base + C::OFF_PEER, sRING = base + C::OFF_RING;
  const uint32_t bars = base + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * S2_COUNT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * S2_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int item = blockIdx.x >> 1;
  const int qt = item % p.nqt, b = item / p.nqt;
  const int nt = (p.Lk + KT - 1) / KT;
  const uint32_t total = (uint32_t)HE * (uint32_t)nt;  // key tiles of one engine, all its heads

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmWo);
    tma_prefetch_desc(&tmXo);
    for (int i = 0; i < S2_COUNT; ++i) {
      uint32_t cnt = 1u;
      if (i == S2_Q_READY) cnt = 256u;
      if (i >= S2_OWN_O && i < S2_OWN_O + 4) cnt = 128u;
      if (i >= S2_ENG) {
        const int k = (i - S2_ENG) % E_COUNT;
        if (k == E_P_FULL || k == E_P_FULL + 1) cnt = 128u;
      }
      mbar_init(bar(i), cnt);
    }
    mbar_arrive_expect_tx(bar(S2_PEER_O), (uint32_t)(HH * C::TILE));
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tD = tmem_base;   // Q / Y accumulator: columns [0, NH); the engines reuse columns [0, 384) in phase 2
  unsigned long long* const stamps = (lane == 0) ? g_sf_stamps : nullptr;
  float* const s_bq = reinterpret_cast<float*>(smem + C::OFF_BIAS);
  if (warp >= 2 && warp != 6 && warp != 7 && warp != 8) {   // the 8 softmax warps: b_q (a parameter, not produced upstream) -> shared memory
    const int t = ((warp >= 9 ? warp - 5 : warp - 2) << 5) + lane;   // 0 .. 255
    if (t < C::NH) s_bq[t] = __ldg(p.b_q + rank * C::NH + t);
    named_bar_sync(1, 256);
  }
  pdl_wait();
  if (warp == 0) SF_STAMP(0);

  // which engine this warp serves (producer / MMA / softmax warps), or -1
  const int eng = (warp == 0 || warp == 1 || (warp >= 2 && warp <= 5)) ? 0 : ((warp == 7 || warp == 8 || warp >= 9) ? 1 : -1);
  const int EB = S2_ENG + E_COUNT * (eng < 0 ? 0 : eng);
  auto r1_stage = [&](int st) { return st < C::ST1 - 1 ? sRING + (uint32_t)st * (C::A_BYTES + C::B_BYTES) : sQO; };
  // phase 3 walks the output projection's contraction OWN heads first, then the peer's: step i is k-block k3(i)
  auto k3 = [&](int i) { return (int)((rank * HH + (uint32_t)i) % (uint32_t)C::NKB); };
  auto wo_stage = [&](int st) { return sRING + (uint32_t)(st == C::ST3 - 1 ? C::OFF_WO_LAST : st * C::B_BYTES); };
  // P V of this engine's (global) tile t has completed: one barrier per tile parity (see csrc/attn.cu)
  auto wait_pv = [&](uint32_t t) { mbar_wait(bar(EB + E_PV_DONE + (t & 1)), (t >> 1) & 1); };
  const uint32_t sE = sRING + (uint32_t)(eng < 0 ? 0 : eng) * C::ENG_BYTES;
  const uint32_t sK = sE, sV = sE + 2 * C::KV_BYTES;
  const uint32_t tS = tmem_base + (uint32_t)(eng < 0 ? 0 : eng) * C::ENG_COLS;   // S0 S1 at +0 / +KT, O at +2 KT
  const uint32_t tO = tS + 2 * KT;
  const int head0 = (eng < 0 ? 0 : eng) * HE;   // first local head of this engine

  if (warp == 0 || warp == 7) {
    // ================================================================ TMA producers
    if (lane == 0) {
      if (warp == 0) {
        for (int kb = 0; kb < C::NKB; ++kb) {
          const int s = kb % C::ST1;
          mbar_wait(bar(S2_R1_EMPTY + s), ((kb / C::ST1) & 1) ^ 1);
          mbar_arrive_expect_tx(bar(S2_R1_FULL + s), C::A_BYTES + C::B_BYTES);
          const uint32_t st = r1_stage(s);
          tma_load_3d(st, &tmX, bar(S2_R1_FULL + s), kb * 64, qt * SF_QT, b);
          tma_load_2d(st + C::A_BYTES, &tmWq, bar(S2_R1_FULL + s), kb * 64, (int)rank * C::NH);
        }
        SF_STAMP(1);
      }
      mbar_wait(bar(S2_D1_FULL), 0);   // the ring region is free once the projection MMAs have retired
      if (warp == 0) {
        // first W_o k-block of phase 3 -> the ring bytes the engines do not use (stage ST3-1); k-block kc lives in
        // stage (kc + ST3 - 1) % ST3, so the stage-reuse parity of k-block kc is (kc / ST3) & 1 as before
        mbar_arrive_expect_tx(bar(S2_R3_FULL + C::ST3 - 1), C::B_BYTES);
        tma_load_2d(wo_stage(C::ST3 - 1), &tmWo, bar(S2_R3_FULL + C::ST3 - 1), k3(0) * 64, (int)rank * C::NH);
      }
      auto load_k = [&](uint32_t g) {
        const uint32_t hh = g / nt, j = g % nt, kb = g & 1, kph = (g >> 1) & 1;
        mbar_wait(bar(EB + E_K_EMPTY + kb), kph ^ 1);
        mbar_arrive_expect_tx(bar(EB + E_K_FULL + kb), C::KV_BYTES);
        tma_load_3d(sK + kb * C::KV_BYTES, &tmK, bar(EB + E_K_FULL + kb), ((int)rank * HH + head0 + (int)hh) * DK, j * KT, b);
      };
      auto load_v = [&](uint32_t g) {
        const uint32_t hh = g / nt, j = g % nt, vb = g & 1, vph = (g >> 1) & 1;
        mbar_wait(bar(EB + E_V_EMPTY + vb), vph ^ 1);
        mbar_arrive_expect_tx(bar(EB + E_V_FULL + vb), C::KV_BYTES);
        tma_load_3d(sV + vb * C::KV_BYTES, &tmV, bar(EB + E_V_FULL + vb), ((int)rank * HH + head0 + (int)hh) * DK, j * KT, b);
      };
      load_k(0);
      for (uint32_t g = 0; g < total; ++g) {
        if (g + 1 < total) load_k(g + 1);
        load_v(g);
      }
      if (warp == 0) {
        mbar_wait(bar(S2_ENG + E_ATT_DONE), 0);        // both engines have retired their last P V
        mbar_wait(bar(S2_ENG + E_COUNT + E_ATT_DONE), 0);
        for (int kc = 1; kc < C::NKB; ++kc) {          // (k-block 0 was requested before the attention phase, below)
          const int s = (kc + C::ST3 - 1) % C::ST3;
          mbar_wait(bar(S2_R3_EMPTY + s), ((kc / C::ST3) & 1) ^ 1);
          mbar_arrive_expect_tx(bar(S2_R3_FULL + s), C::B_BYTES);
          tma_load_2d(wo_stage(s), &tmWo, bar(S2_R3_FULL + s), k3(kc) * 64, (int)rank * C::NH);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 8) {
    // ================================================================ MMA issuers
    constexpr uint32_t idesc_p = make_idesc_f16(SF_QT, C::NH, 0, 0);
    constexpr uint32_t idesc_s = make_idesc_f16(SF_QT, KT, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(SF_QT, DK, 0, 1);
    if (warp == 1) {
      for (int kb = 0; kb < C::NKB; ++kb) {
        const int s = kb % C::ST1;
        mbar_wait(bar(S2_R1_FULL + s), (kb / C::ST1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = r1_stage(s);
          const uint64_t da = make_smem_desc(st, 16, 1024, SWZ_128B);
          const uint64_t db = make_smem_desc(st + C::A_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16(tD, da + 2 * k, db + 2 * k, idesc_p, (kb | k) != 0);
          tc_commit(bar(S2_R1_EMPTY + s));
          if (kb == C::NKB - 1) tc_commit(bar(S2_D1_FULL));
        }
        __syncwarp();
      }
      SF_STAMP(2);
    }
    mbar_wait(bar(S2_Q_READY), 0);   // every Q tile is in shared memory AND the whole Q accumulator has been read
    tc_fence_after();
    auto issue_qk = [&](uint32_t g) {
      const uint32_t hh = g / nt, sb = g & 1, ph2 = (g >> 1) & 1;
      mbar_wait(bar(EB + E_K_FULL + sb), ph2);
      if (g >= 2) wait_pv(g - 2);   // this S buffer held P of tile g-2: its P V must have COMPLETED (csrc/attn.cu)
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dq = make_smem_desc(sQO + (uint32_t)(head0 + (int)hh) * C::TILE, 16, 1024, SWZ_128B);
        const uint64_t dk = make_smem_desc(sK + sb * C::KV_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) tc_mma_f16(tS + sb * KT, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(bar(EB + E_K_EMPTY + sb));
        tc_commit(bar(EB + E_S_FULL + sb));
      }
      __syncwarp();
    };
    issue_qk(0);
    for (uint32_t g = 0; g < total; ++g) {
      if (g + 1 < total) issue_qk(g + 1);
      const uint32_t ph = g & 1, j = g % nt;
      mbar_wait(bar(EB + E_V_FULL + ph), (g >> 1) & 1);
      mbar_wait(bar(EB + E_P_FULL + ph), (g >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
          const uint64_t dv = make_smem_desc(sV + ph * C::KV_BYTES + kk * 16 * 128, KT * 128, 1024, SWZ_128B);
          tc_mma_f16_ts(tO, tS + ph * KT + kk * 8, dv, idesc_o, (j | (uint32_t)kk) != 0);   // P: 16 keys = 8 columns of S buffer ph
        }
        tc_commit(bar(EB + E_V_EMPTY + ph));
        tc_commit(bar(EB + E_PV_DONE + ph));
        if (g == total - 1) tc_commit(bar(EB + E_ATT_DONE));
      }
      __syncwarp();
    }
    if (warp == 1) {
      // ---- phase 3 (after BOTH engines: their tensor-memory columns become the Y accumulator)
      for (int hh = 0; hh < HH; ++hh) mbar_wait(bar(S2_OWN_O + hh), 0);
      mbar_wait(bar(S2_ENG + E_COUNT + E_ATT_DONE), 0);
      SF_STAMP(11);
      tc_fence_after();
      for (int kc = 0; kc < C::NKB; ++kc) {   // step kc: k-block k3(kc) -- own O tiles first, the peer's arrive meanwhile
        const int s = (kc + C::ST3 - 1) % C::ST3;
        if (kc == HH) {
          mbar_wait(bar(S2_PEER_O), 0);
          SF_STAMP(7);
        }
        mbar_wait(bar(S2_R3_FULL + s), (kc / C::ST3) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_tile = (kc < HH ? sQO : sPEER) + (uint32_t)(kc % HH) * C::TILE;
          const uint64_t da = make_smem_desc(a_tile, 16, 1024, SWZ_128B);
          const uint64_t db = make_smem_desc(wo_stage(s), 16, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16(tD, da + 2 * k, db + 2 * k, idesc_p, (kc | k) != 0);
          tc_commit(bar(S2_R3_EMPTY + s));
          if (kc == C::NKB - 1) tc_commit(bar(S2_D3_FULL));
        }
        __syncwarp();
      }
    }
  } else if (warp == 6) {
    // ================================================================ DSMEM sender
    if (lane == 0) {
      const uint32_t peer = rank ^ 1u;
      const uint32_t peer_bar = sf_mapa(bar(S2_PEER_O), peer);
      for (int hh = 0; hh < HH; ++hh) {
        mbar_wait(bar(S2_OWN_O + hh), 0);
        sf_dsmem_copy(sf_mapa(sPEER + hh * C::TILE, peer), sQO + hh * C::TILE, (uint32_t)C::TILE, peer_bar);
      }
    }
    __syncwarp();
  } else {
    // ================================================================ softmax warpgroups (engine 0: warps 2..5, engine 1: 9..12)
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    constexpr float LOG2E = 1.4426950408889634f;
    const float c1 = p.scale * LOG2E;
    const float t_masked = -1e9f * LOG2E;
    constexpr int NCH = KT / 32;
    constexpr int CH = C::NH / 64;   // 32-column chunks of the Q / Y accumulator per warpgroup

    // ---- phase 1 drain of this group's heads: columns [eng * NH/2, (eng+1) * NH/2)
    mbar_wait(bar(S2_D1_FULL), 0);
    tc_fence_after();
    if (warp == 2) SF_STAMP(3);
#pragma unroll 1
    for (int c = eng * CH; c < (eng + 1) * CH; ++c) {
      uint32_t r[32];
      tc_ld32(tD + lane_off + c * 32, r);
      float4 bb[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) bb[t] = reinterpret_cast<const float4*>(s_bq + c * 32)[t];   // broadcast reads
      tc_wait_ld();
      const uint32_t tile = sQO + (uint32_t)(c >> 1) * C::TILE + row * 128;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t chunk = ((uint32_t)((c & 1) * 4 + t)) ^ sw;
        const float4 b0 = bb[2 * t], b1 = bb[2 * t + 1];
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + chunk * 16),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t]) + b0.x, __uint_as_float(r[8 * t + 1]) + b0.y)),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 2]) + b0.z, __uint_as_float(r[8 * t + 3]) + b0.w)),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 4]) + b1.x, __uint_as_float(r[8 * t + 5]) + b1.y)),
                     "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 6]) + b1.z, __uint_as_float(r[8 * t + 7]) + b1.w))
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar(S2_Q_READY));
    if (warp == 2) SF_STAMP(4);

    // ---- phase 2
    const bool live = qt * SF_QT + q4 * 32 < p.Lq;
    const uint32_t* mrow = nullptr;
    if (p.mask_bits != nullptr) {
      const int mq = (p.mask_rows_q == 1) ? 0 : min(qt * SF_QT + row, p.Lq - 1);
      mrow = p.mask_bits + ((size_t)b * p.mask_rows_q + mq) * p.mask_words;
    }
    uint32_t mw_pref[NCH];
    auto fetch_mask = [&](int jn) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int k0 = jn * KT + c * 32;
        mw_pref[c] = (mrow != nullptr && k0 < p.Lk) ? __ldcg(mrow + (k0 >> 5)) : 0xffffffffu;
      }
    };
    fetch_mask(0);
    uint32_t g = 0;
    for (int hh = 0; hh < HE; ++hh) {
      const int hl = head0 + hh;   // local head index (tile / OWN_O barrier)
      float m_run = -CUDART_INF_F, l_run = 0.f;
      if (!live) {
        for (int j = 0; j < nt; ++j, ++g) {
          const uint32_t ph = g & 1;
          mbar_wait(bar(EB + E_S_FULL + ph), (g >> 1) & 1);
          mbar_arrive(bar(EB + E_P_FULL + ph));
        }
        wait_pv(g - 1);
        mbar_arrive(bar(S2_OWN_O + hl));
        continue;
      }
      for (int j = 0; j < nt; ++j, ++g) {
        const uint32_t ph = g & 1;
        const uint32_t tSb = tS + ph * KT;
        uint32_t mwv[NCH];
        bool plain = (j + 1) * KT <= p.Lk;
        {
          uint32_t w = 0xffffffffu;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            mwv[c] = mw_pref[c];
            w &= mwv[c];
          }
          if (mrow != nullptr) plain = plain && __all_sync(0xffffffffu, w == 0xffffffffu);
        }
        fetch_mask(j + 1 < nt ? j + 1 : 0);
        mbar_wait(bar(EB + E_S_FULL + ph), (g >> 1) & 1);
        tc_fence_after();
        if (warp == 2 && g == 0) SF_STAMP(5);
        if (warp == 2 && g == 1) SF_STAMP(15);
        auto store_chunk = [&](int c, const uint32_t(&e)[32]) {   // keys 32 c .. +31 of this row -> columns [16 c, 16 c + 16) of the S buffer
          uint32_t pk[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2_sat(__uint_as_float(e[2 * t]), __uint_as_float(e[2 * t + 1]));
          tc_st16(tSb + lane_off + c * 16, pk);
        };
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        float m_new;
        bool rescale = false;
        auto pick_max = [&](float m_tile) {
          m_new = fmaxf(m_run, m_tile);
          if (j > 0) {
            rescale = __any_sync(0xffffffffu, m_new - m_run > 8.f);
            if (!rescale) m_new = m_run;
          }
        };
        if (plain) {
          uint32_t r[NCH][32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc_ld32(tSb + lane_off + c * 32, r[c]);
          tc_wait_ld();
          float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r[c][i]));
          }
          pick_max(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c1);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float e = sf_ex2(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
              l4[i & 3] += e;
              r[c][i] = __float_as_uint(e);
            }
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) store_chunk(c, r[c]);
        } else {
          // masked / ragged tile, register-resident like a plain one (csrc/attn.cu): per 32-key chunk a warp-uniform choice
          // between select-free code (every row keeps every key), a constant (no row keeps any key) and per-key selects
          uint32_t r[NCH][32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc_ld32(tSb + lane_off + c * 32, r[c]);
          uint32_t inbv[NCH], kmv[NCH];
          int kind[NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int nvalid = p.Lk - (j * KT + c * 32);
            inbv[c] = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
            kmv[c] = mwv[c] & inbv[c];
            kind[c] = __all_sync(0xffffffffu, kmv[c] == 0xffffffffu) ? 0 : (__all_sync(0xffffffffu, kmv[c] == 0u) ? 1 : 2);
          }
          tc_wait_ld();
          float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
          uint32_t masked_any = 0u;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            masked_any |= inbv[c] & ~mwv[c];
            if (kind[c] == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r[c][i]));
            } else if (kind[c] == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                mx[i & 3] = fmaxf(mx[i & 3], ((kmv[c] >> i) & 1u) ? __uint_as_float(r[c][i]) : -CUDART_INF_F);
            }
          }
          pick_max(fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c1, masked_any ? t_masked : -CUDART_INF_F));
          const float pm = sf_ex2(t_masked - m_new);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (kind[c] == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float e = sf_ex2(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
                l4[i & 3] += e;
                r[c][i] = __float_as_uint(e);
              }
            } else if (kind[c] == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) r[c][i] = ((inbv[c] >> i) & 1u) ? __float_as_uint(pm) : 0u;
              l4[0] += pm * (float)__popc(inbv[c]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = sf_ex2(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
                const float e = ((kmv[c] >> i) & 1u) ? x : (((inbv[c] >> i) & 1u) ? pm : 0.f);
                l4[i & 3] += e;
                r[c][i] = __float_as_uint(e);
              }
            }
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) store_chunk(c, r[c]);
        }
        const float alpha = sf_ex2(m_run - m_new);
        l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
        m_run = m_new;
        if (rescale) {
          wait_pv(g - 1);   // (rescale implies j > 0) the accumulator is idle once the previous tile's P V has retired
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < DK / 32; ++c) {
            uint32_t o[32];
            tc_ld32(tO + lane_off + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc_st32(tO + lane_off + c * 32, o);
          }
          tc_wait_st();
        }
        tc_wait_st();   // P (and a rescaled O) have landed in tensor memory
        tc_fence_before();
        mbar_arrive(bar(EB + E_P_FULL + ph));
        if (warp == 2 && g == 0) SF_STAMP(12);
      }
      wait_pv(g - 1);
      tc_fence_after();
      if (warp == 2 && hh == 0) SF_STAMP(13);
      const float inv_l = 1.f / l_run;
      const uint32_t tile = sQO + (uint32_t)hl * C::TILE + row * 128;
#pragma unroll
      for (int c = 0; c < DK / 32; ++c) {
        uint32_t r[32];
        tc_ld32(tO + lane_off + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t chunk = (uint32_t)(c * 4 + t) ^ sw;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + chunk * 16),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t]) * inv_l, __uint_as_float(r[8 * t + 1]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 2]) * inv_l, __uint_as_float(r[8 * t + 3]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 4]) * inv_l, __uint_as_float(r[8 * t + 5]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 6]) * inv_l, __uint_as_float(r[8 * t + 7]) * inv_l))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar(S2_OWN_O + hl));
      if (warp == 2 && hh == 0) SF_STAMP(14);
    }
    if (warp == 2) SF_STAMP(6);
    if (warp == 9) SF_STAMP(10);

    // ---- phase 3 epilogue: this group's half of the CTA's output columns.  The residual add x += Y + b_o goes out as
    // TMA REDUCE-ADD stores: per warp and 32-column chunk, the [32 rows x 32 cols] f32 tile is staged in shared memory
    // (row per thread, 128B-swizzled like the tensor map) and one bulk operation adds it into the residual stream.
    // (Per-lane red.global.add retires ~1 lane per clock per SM: 128 KB per CTA took ~8.5 k cycles, the longest
    // phase of the kernel; tools/site_phases.py.)  Rows beyond Lq are clipped by the tensor map.
    const int t256 = ((warp >= 9 ? warp - 5 : warp - 2) << 5) + lane;
    const float bo_mine = t256 < C::NH ? __ldg(p.b_o + rank * C::NH + t256) : 0.f;   // requested while phase 3 still runs
    mbar_wait(bar(S2_Q_READY), 0);          // (long complete) every warp has finished reading b_q from shared memory
    if (t256 < C::NH) s_bq[t256] = bo_mine;   // the bias area now holds b_o of this CTA's columns
    named_bar_sync(1, 256);
    mbar_wait(bar(S2_D3_FULL), 0);
    tc_fence_after();
    if (warp == 2) SF_STAMP(8);
    // Staging: one 4 KB tile per (warp, chunk) in the Q/O and peer-O tile regions, which are dead once the phase-3 MMAs
    // have retired (D3_FULL): no tile is reused, every chunk is issued as soon as it is staged.  What bounds this phase is
    // the SM's write path into L2 (128 KB per CTA at ~26 B per clock measured, whatever the box size: 16 KB boxes issued
    // by one thread per group were no faster), so the bulk operations should start as early as possible.
    const int ew = eng * 4 + q4;
    const uint32_t tiles = sQO + (uint32_t)ew * (uint32_t)(CH * 4096);
    static_assert(8 * CH * 4096 <= 2 * HH * C::TILE, "epilogue staging fits the O tile regions");
#pragma unroll 1
    for (int cc = 0; cc < CH; ++cc) {   // (sw == row % 8 == lane % 8: the 16-byte chunk XOR of the 128B swizzle)
      const int c = eng * CH + cc;
      const uint32_t tile = tiles + (uint32_t)cc * 4096u;
      uint32_t acc[32];
      tc_ld32(tD + lane_off + c * 32, acc);
      float4 bb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bb[j] = reinterpret_cast<const float4*>(s_bq + c * 32)[j];
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + lane * 128 + (((uint32_t)j ^ sw) << 4)),
                     "f"(__uint_as_float(acc[4 * j]) + bb[j].x), "f"(__uint_as_float(acc[4 * j + 1]) + bb[j].y),
                     "f"(__uint_as_float(acc[4 * j + 2]) + bb[j].z), "f"(__uint_as_float(acc[4 * j + 3]) + bb[j].w)
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(&tmXo, tile, (int)rank * C::NH + c * 32, qt * SF_QT + q4 * 32, b);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();   // (the kernel boundary completes the writes)
    tc_fence_before();
    if (warp == 2) SF_STAMP(9);
  }
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int HH>
static int launch_site_fused2(const MtnAttnSiteFusedArgs& a, cudaStream_t st) {
  using C = Sf2Cfg<HH>;
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(attn_site_fused2_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    attr_set = true;
  }
  const int d = a.d;
  CUtensorMap tx, twq, tk, tv, two;
  int rc = make_tmap_3d_f16(&tx, a.xn_f16, d, a.Lq, a.B, a.ld_xn, (uint64_t)a.Lq * a.ld_xn, 64, SF_QT, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&twq, a.w_q, d, d, a.ld_wq > 0 ? a.ld_wq : d, 64, C::NH, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&two, a.w_o, d, d, a.ld_wo > 0 ? a.ld_wo : d, 64, C::NH, TM_SWZ_128);
  if (rc) return rc;
  const uint8_t* kp = static_cast<const uint8_t*>(a.kv) + (size_t)a.kv_k_col * 2;
  const uint8_t* vp = static_cast<const uint8_t*>(a.kv) + (size_t)a.kv_v_col * 2;
  rc = make_tmap_3d_f16(&tk, kp, d, a.Lk, a.B, a.ld_kv, (uint64_t)a.Lk * a.ld_kv, 64, C::KT, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tv, vp, d, a.Lk, a.B, a.ld_kv, (uint64_t)a.Lk * a.ld_kv, 64, C::KT, TM_SWZ_128);
  if (rc) return rc;
  const int nqt = (a.Lq + SF_QT - 1) / SF_QT;
  SfParams p{a.B, a.h, a.Lq, a.Lk, nqt, a.b_q, a.b_o, a.x, a.ld_x, a.mask_bits, a.mask_rows_q, mtn_mask_words(a.Lk),
             1.0f / sqrtf(64.f)};
  dim3 grid(2u * (unsigned)(a.B * nqt));
  CUtensorMap txo;   // the f32 residual stream, target of the epilogue's reduce-add stores
  rc = make_tmap_3d_f32(&txo, a.x, d, a.Lq, a.B, a.ld_x, (uint64_t)a.Lq * a.ld_x, 32, 32, TM_SWZ_128);
  if (rc) return rc;
  MTN_CHECK_CUDA(launch_kernel_cluster(attn_site_fused2_kernel<HH>, grid, dim3(SF2_THREADS), C::TOTAL, st, 2u, tx, twq, tk, tv,
                                       two, txo, p));
  return MTN_OK;
}

template <int HH, int KT>
static int launch_site_fused(const MtnAttnSiteFusedArgs& a, cudaStream_t st) {
  using C = SfCfg<HH, KT>;
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(attn_site_fused_kernel<HH, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    attr_set = true;
  }
  const int d = a.d;
  CUtensorMap tx, twq, tk, tv, two;
  int rc = make_tmap_3d_f16(&tx, a.xn_f16, d, a.Lq, a.B, a.ld_xn, (uint64_t)a.Lq * a.ld_xn, 64, SF_QT, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&twq, a.w_q, d, d, a.ld_wq > 0 ? a.ld_wq : d, 64, C::NH, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&two, a.w_o, d, d, a.ld_wo > 0 ? a.ld_wo : d, 64, C::NH, TM_SWZ_128);
  if (rc) return rc;
  const uint8_t* kp = static_cast<const uint8_t*>(a.kv) + (size_t)a.kv_k_col * 2;
  const uint8_t* vp = static_cast<const uint8_t*>(a.kv) + (size_t)a.kv_v_col * 2;
  rc = make_tmap_3d_f16(&tk, kp, d, a.Lk, a.B, a.ld_kv, (uint64_t)a.Lk * a.ld_kv, 64, KT, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tv, vp, d, a.Lk, a.B, a.ld_kv, (uint64_t)a.Lk * a.ld_kv, 64, KT, TM_SWZ_128);
  if (rc) return rc;
  const int nqt = (a.Lq + SF_QT - 1) / SF_QT;
  SfParams p{a.B, a.h, a.Lq, a.Lk, nqt, a.b_q, a.b_o, a.x, a.ld_x, a.mask_bits, a.mask_rows_q, mtn_mask_words(a.Lk),
             1.0f / sqrtf(64.f)};
  dim3 grid(2u * (unsigned)(a.B * nqt));
  MTN_CHECK_CUDA(launch_kernel_cluster(attn_site_fused_kernel<HH, KT>, grid, dim3(SF_THREADS), C::TOTAL, st, 2u, tx, twq, tk, tv,
                                       two, p));
  return MTN_OK;
}

}  // namespace mtn

// Debug (not part of include/mtn_b200.h): register / clear the phase-stamp buffer of the v2 kernel, [grid * 16] u64.
extern "C" int mtn_debug_site_stamps(void* dev_ptr) {
  unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
  MTN_CHECK_CUDA(cudaMemcpyToSymbol(mtn::g_sf_stamps, &p, sizeof(p)));
  return MTN_OK;
}

extern "C" int mtn_attn_site_fused_supported(int d, int h) {
  return (h > 0 && d == h * 64 && (d == 512 || d == 256)) ? 1 : 0;
}

extern "C" int mtn_attn_site_fused_fwd(const MtnAttnSiteFusedArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->xn_f16 && a->x && a->w_q && a->b_q && a->w_o && a->b_o && a->kv, MTN_E_ARG, "attn_site_fused: NULL pointer");
  MTN_REQUIRE(a->B > 0 && a->Lq > 0 && a->Lk > 0, MTN_E_SHAPE, "attn_site_fused: B=%d Lq=%d Lk=%d", a->B, a->Lq, a->Lk);
  MTN_REQUIRE(mtn_attn_site_fused_supported(a->d, a->h), MTN_E_SHAPE,
              "attn_site_fused: d=%d h=%d (supported: d_k = 64 with d in {256, 512}; other shapes take mtn_attn_site_fwd)", a->d, a->h);
  MTN_REQUIRE(a->ld_xn >= a->d && a->ld_xn % 8 == 0 && a->ld_kv % 8 == 0 && a->ld_x >= a->d && a->ld_x % 4 == 0 &&
                  a->kv_k_col % 8 == 0 && a->kv_v_col % 8 == 0 && a->kv_k_col >= 0 && a->kv_v_col >= 0 &&
                  a->ld_kv >= a->kv_k_col + a->d && a->ld_kv >= a->kv_v_col + a->d,
              MTN_E_ALIGN, "attn_site_fused: leading dimensions / K,V columns");
  MTN_REQUIRE(aligned16(a->xn_f16) && aligned16(a->x) && aligned16(a->w_q) && aligned16(a->w_o) && aligned16(a->kv) &&
                  aligned16(a->b_q) && aligned16(a->b_o),
              MTN_E_ALIGN, "attn_site_fused: pointers must be 16-byte aligned");
  MTN_REQUIRE(a->mask_bits == nullptr || a->mask_rows_q == 1 || a->mask_rows_q == a->Lq, MTN_E_SHAPE,
              "attn_site_fused: mask_rows_q=%d must be 1 or Lq=%d", a->mask_rows_q, a->Lq);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static int version = 0;   // MTN_B200_SITE_FUSED_V=1: the single-engine kernel (v1); default: two engines per CTA (v2)
  if (version == 0) {
    const char* e = getenv("MTN_B200_SITE_FUSED_V");
    version = (e != nullptr && e[0] == '1') ? 1 : 2;
  }
  if (version == 2) return a->d == 512 ? launch_site_fused2<4>(*a, st) : launch_site_fused2<2>(*a, st);
  if (a->d == 512) return a->Lk <= 64 ? launch_site_fused<4, 64>(*a, st) : launch_site_fused<4, 96>(*a, st);
  return a->Lk <= 64 ? launch_site_fused<2, 64>(*a, st) : launch_site_fused<2, 96>(*a, st);
}
