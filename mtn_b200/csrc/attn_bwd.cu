// Attention core BACKWARD on tcgen05 (training; autograd of attention(), mtn.py:221-231).
//
// Given Q, K, V (f16 head slices of the packed projection buffers, as in the forward), dO (f16), the
// forward's per-row softmax statistics {m, 1/l} and delta = rowsum(dO * O):
//     P  = softmax(mask(Q K^T / sqrt(d_k)))             (recomputed, never stored)
//     dV = P^T dO        dP = dO V^T        dS = P * (dP - delta) / sqrt(d_k)   [0 where mask == 0]
//     dQ = dS K          dK = dS^T Q
// Masked scores are the constant -1e9 (mtn.py:227), so no gradient flows through them into Q / K, but their
// probabilities (non-zero only in fully masked rows) still weight dV exactly as autograd does.
//
// Work item = (128-key tile, head, batch element); persistent CTAs (one per SM) walk the items and, per item, loop
// over 128-query tiles.  All barrier phases run on per-CTA global counters (item counter for the double-buffered K/V
// tiles, tile counter for everything else), so the producer streams the next item's K / V / Q / dO and the tensor
// core computes its first scores while the current item is still in its gradient MMAs and drains.  Per query tile the
// tensor core computes S and dP (both K-major operands, like the forward), 256 threads (one TMEM lane = one
// query row, two warps per lane quarter splitting the key columns) turn them into P and dS, written as f16
// to shared memory in the 128-byte-swizzled panel layout, and three more MMAs consume those panels:
//   dV += P^T dO   and   dK += dS^T Q   read the panels as MN-major A operands (the transpose is free:
//                                       UMMA's major-ness bit) and dO / Q as MN-major B operands;
//   dQ  = dS K     reads dS as a K-major A operand and K as an MN-major B operand.
// dV / dK accumulate in tensor memory over the whole query loop and are stored once (f16); dQ (two TMEM
// buffers, drained while the next tile is in flight) is accumulated across key tiles into an f32 buffer
// with red.global.add.  TMEM: S 128 + dP 128 + dV + dK + 2 x dQ = 256 + 4 d_k columns (512 at d_k = 64).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..9 softmax backward.
#include <math_constants.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int AB_THREADS = 320;
constexpr int AB_T = 128;  // queries per step == keys per CTA

template <int DK>
struct AttnBwdCfg {
  static constexpr int ROWB = DK * 2;
  static constexpr int TILE = AB_T * ROWB;             // one 128-row operand tile
  static constexpr int PANEL = AB_T * 128;             // [128 rows x 64 f16] swizzled panel
  static constexpr int OFF_K = 0;                      // two buffers of [K | V] (next item's keys / values)
  static constexpr int OFF_V = OFF_K + TILE;
  static constexpr int KV_STRIDE = 2 * TILE;
  static constexpr int OFF_Q = OFF_K + 2 * KV_STRIDE;  // two buffers
  static constexpr int OFF_DO = OFF_Q + 2 * TILE;      // two buffers
  static constexpr int OFF_P = OFF_DO + 2 * TILE;      // two panels (keys 0-63, 64-127)
  static constexpr int OFF_DS = OFF_P + 2 * PANEL;
  static constexpr int OFF_BAR = OFF_DS + 2 * PANEL;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;
  static constexpr uint64_t SWZ = (DK == 64) ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SBO = 8 * ROWB;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 256 + DK, COL_DQ = 256 + 2 * DK;
  static_assert(TILE % 1024 == 0, "operand tiles must stay swizzle-atom aligned");
  static_assert(256 + 4 * DK <= 512, "TMEM budget");
};

struct AttnBwdParams {
  int n_items, nkt;  // work items = B * h * nkt, ordered (b, head, key tile) with the key tile fastest
  const uint32_t* mask_bits;
  int mask_rows_q, mask_words;
  int B, h, Lq, Lk;
  float scale;
  const float2* stats;
  const float* delta;
  float* dq; int lddq;
  __half* dq16; int lddq16;  // single key tile (Lk <= 128): dQ is complete inside the CTA and stored as f16 directly
  __half* dk; int lddk;
  __half* dv; int lddv;
  DropCfg drop;  // the forward's dropout of the probabilities, regenerated
  int Lk32;
};

enum { ABAR_KV_FULL = 0 /* +1 */, ABAR_KV_EMPTY = 2 /* +1 */, ABAR_QDO_FULL = 4 /* +1 */, ABAR_QDO_EMPTY = 6 /* +1 */,
       ABAR_S_FULL = 8, ABAR_P_FULL, ABAR_G_DONE, ABAR_COUNT };

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DK>
__global__ void __launch_bounds__(AB_THREADS, 1)
    attn_core_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                            const AttnBwdParams p) {
  using C = AttnBwdCfg<DK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sK = base + C::OFF_K, sV = base + C::OFF_V, sQ = base + C::OFF_Q, sDO = base + C::OFF_DO;
  const uint32_t sP = base + C::OFF_P, sDS = base + C::OFF_DS;
  const uint32_t bars = base + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * ABAR_COUNT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * ABAR_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (p.Lq + AB_T - 1) / AB_T;
  auto item_of = [&](int it, int& kt, int& hd, int& b) {
    kt = it % p.nkt;
    hd = (it / p.nkt) % p.h;
    b = it / (p.nkt * p.h);
  };

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    for (int i = 0; i < ABAR_COUNT; ++i) mbar_init(bar(i), i == ABAR_P_FULL ? 256u : 1u);  // P_FULL: all softmax threads
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    // -------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t n = 0, g = 0;  // items / query tiles issued so far by this CTA
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x, ++n) {
        int kt, hd, b;
        item_of(it, kt, hd, b);
        const uint32_t kb = n & 1;
        mbar_wait(bar(ABAR_KV_EMPTY + kb), ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(ABAR_KV_FULL + kb), 2 * C::TILE);
        tma_load_3d(sK + kb * C::KV_STRIDE, &tmK, bar(ABAR_KV_FULL + kb), hd * DK, kt * AB_T, b);
        tma_load_3d(sV + kb * C::KV_STRIDE, &tmV, bar(ABAR_KV_FULL + kb), hd * DK, kt * AB_T, b);
        for (int i = 0; i < nq; ++i, ++g) {
          const uint32_t buf = g & 1;
          mbar_wait(bar(ABAR_QDO_EMPTY + buf), ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(bar(ABAR_QDO_FULL + buf), 2 * C::TILE);
          tma_load_3d(sQ + buf * C::TILE, &tmQ, bar(ABAR_QDO_FULL + buf), hd * DK, i * AB_T, b);
          tma_load_3d(sDO + buf * C::TILE, &tmDO, bar(ABAR_QDO_FULL + buf), hd * DK, i * AB_T, b);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = make_idesc_f16(AB_T, AB_T, 0, 0);  // S = Q K^T, dP = dO V^T
    constexpr uint32_t idesc_g = make_idesc_f16(AB_T, DK, 1, 1);    // dV += P^T dO, dK += dS^T Q
    constexpr uint32_t idesc_q = make_idesc_f16(AB_T, DK, 0, 1);    // dQ = dS K
    const uint32_t tS = tmem_base + C::COL_S, tDP = tmem_base + C::COL_DP;
    const uint32_t tDV = tmem_base + C::COL_DV, tDK = tmem_base + C::COL_DK;
    const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_items * (uint32_t)nq;  // query tiles of this CTA, all items
    // scores of global tile g (item g / nq, its query tile g % nq): whole warp waits for the operands, lane 0 issues
    auto issue_scores = [&](uint32_t g) {
      const uint32_t n = g / nq, buf = g & 1, kb = n & 1;
      if (g % nq == 0) mbar_wait(bar(ABAR_KV_FULL + kb), (n >> 1) & 1);
      mbar_wait(bar(ABAR_QDO_FULL + buf), (g >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dq = make_smem_desc(sQ + buf * C::TILE, 16, C::SBO, C::SWZ);
        const uint64_t ddo = make_smem_desc(sDO + buf * C::TILE, 16, C::SBO, C::SWZ);
        const uint64_t dk = make_smem_desc(sK + kb * C::KV_STRIDE, 16, C::SBO, C::SWZ);
        const uint64_t dv = make_smem_desc(sV + kb * C::KV_STRIDE, 16, C::SBO, C::SWZ);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) tc_mma_f16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) tc_mma_f16(tDP, ddo + 2 * k, dv + 2 * k, idesc_s, k != 0);
        tc_commit(bar(ABAR_S_FULL));
      }
      __syncwarp();
    };
    if (total > 0) issue_scores(0);
    for (uint32_t g = 0; g < total; ++g) {
      const uint32_t n = g / nq, i = g % nq, buf = g & 1, kb = n & 1;
      mbar_wait(bar(ABAR_P_FULL), g & 1);  // P_g / dS_g are in shared memory; S / dP and dQ[buf] may be overwritten
      // scores of the next tile first (possibly the first tile of the NEXT item): its softmax overlaps the MMAs below
      if (g + 1 < total) issue_scores(g + 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t tDQ = tmem_base + C::COL_DQ + buf * DK;
        const uint32_t sKb = sK + kb * C::KV_STRIDE;
#pragma unroll
        for (int kk = 0; kk < AB_T / 16; ++kk) {  // contraction over the 128 queries, 16 per step
          // A: panel rows [16 kk, +16) (2048 B), both 64-key panels (LBO = panel stride), MN-major
          const uint64_t dp = make_smem_desc(sP + kk * 2048, C::PANEL, 1024, SWZ_128B);
          const uint64_t dds = make_smem_desc(sDS + kk * 2048, C::PANEL, 1024, SWZ_128B);
          // B: dO / Q rows [16 kk, +16), d_k contiguous (MN-major, one swizzle atom wide)
          const uint64_t bdo = make_smem_desc(sDO + buf * C::TILE + kk * 16 * C::ROWB, C::TILE, C::SBO, C::SWZ);
          const uint64_t bq = make_smem_desc(sQ + buf * C::TILE + kk * 16 * C::ROWB, C::TILE, C::SBO, C::SWZ);
          tc_mma_f16(tDV, dp, bdo, idesc_g, (i | (uint32_t)kk) != 0);
          tc_mma_f16(tDK, dds, bq, idesc_g, (i | (uint32_t)kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < AB_T / 16; ++kk) {  // contraction over the 128 keys
          const uint64_t ads = make_smem_desc(sDS + (kk >> 2) * C::PANEL + (kk & 3) * 32, 16, 1024, SWZ_128B);
          const uint64_t bk = make_smem_desc(sKb + kk * 16 * C::ROWB, C::TILE, C::SBO, C::SWZ);
          tc_mma_f16(tDQ, ads, bk, idesc_q, kk != 0);
        }
        tc_commit(bar(ABAR_QDO_EMPTY + buf));
        if (i == (uint32_t)nq - 1) tc_commit(bar(ABAR_KV_EMPTY + kb));  // last use of this item's keys / values
        tc_commit(bar(ABAR_G_DONE));
      }
      __syncwarp();
    }
  } else {
    // -------------------------------------------------------------- softmax backward + drains
    const int q4 = warp & 3;           // TMEM lane quarter
    const int half = (warp - 2) >> 2;  // 0: keys 0-63 (panel 0), 1: keys 64-127 (panel 1)
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    const float c1 = p.scale * LOG2E;
    const float t_masked = -1e9f * LOG2E;
    const uint32_t sw = (uint32_t)(row & 7);
    const unsigned long long dseed = p.drop.seed != nullptr ? __ldcg(p.drop.seed) : 0ull;
    uint32_t g = 0;  // query tiles processed so far by this CTA (all items): barrier phases and the dQ buffer index

    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
    int kt, hd, b;
    item_of(it, kt, hd, b);
    const size_t bh = (size_t)b * p.h + hd;

    auto drain_dq = [&](uint32_t gt, int i) {  // dQ of query tile i (global tile gt, TMEM buffer gt & 1) -> global memory
      if (half * 32 < DK) {
        const int qi = i * AB_T + row;
        uint32_t r[32];
        tc_ld32(tmem_base + C::COL_DQ + (gt & 1) * DK + half * 32 + lane_off, r);
        tc_wait_ld();
        if (qi < p.Lq && p.dq16 != nullptr) {
          uint4* o = reinterpret_cast<uint4*>(p.dq16 + ((size_t)b * p.Lq + qi) * p.lddq16 + hd * DK + half * 32);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            o[t] = make_uint4(pack_f16x2_sat(__uint_as_float(r[8 * t]), __uint_as_float(r[8 * t + 1])),
                              pack_f16x2_sat(__uint_as_float(r[8 * t + 2]), __uint_as_float(r[8 * t + 3])),
                              pack_f16x2_sat(__uint_as_float(r[8 * t + 4]), __uint_as_float(r[8 * t + 5])),
                              pack_f16x2_sat(__uint_as_float(r[8 * t + 6]), __uint_as_float(r[8 * t + 7])));
        } else if (qi < p.Lq) {
          float* o = p.dq + ((size_t)b * p.Lq + qi) * p.lddq + hd * DK + half * 32;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * t), "f"(__uint_as_float(r[4 * t])),
                         "f"(__uint_as_float(r[4 * t + 1])), "f"(__uint_as_float(r[4 * t + 2])),
                         "f"(__uint_as_float(r[4 * t + 3]))
                         : "memory");
        }
      }
    };

    for (int i = 0; i < nq; ++i, ++g) {
      const int qi = i * AB_T + row;
      const bool valid = qi < p.Lq;
      float m_row = 0.f, inv_l = 0.f, delta = 0.f;
      if (valid) {
        const float2 st = __ldcg(p.stats + bh * p.Lq + qi);
        m_row = st.x;
        inv_l = st.y;
        delta = __ldcg(p.delta + bh * p.Lq + qi);  // written by the predecessor (attn_delta): coherent load, common.cuh
      }
      const uint32_t* mrow = nullptr;
      if (p.mask_bits != nullptr) {
        const int mq = (p.mask_rows_q == 1) ? 0 : min(qi, p.Lq - 1);
        mrow = p.mask_bits + ((size_t)b * p.mask_rows_q + mq) * p.mask_words;
      }
      mbar_wait(bar(ABAR_S_FULL), g & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        const int k0 = kt * AB_T + c * 32;
        const int nvalid = p.Lk - k0;
        const uint32_t inb = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
        const uint32_t mw = (mrow != nullptr && nvalid > 0) ? __ldcg(mrow + (k0 >> 5)) : 0xffffffffu;
        uint32_t pk_p[16], pk_d[16];
        if (nvalid > 0) {  // warp-uniform
          uint32_t s[32], d[32];
          tc_ld32(tmem_base + C::COL_S + lane_off + c * 32, s);
          tc_ld32(tmem_base + C::COL_DP + lane_off + c * 32, d);
          tc_wait_ld();
          uint32_t dropk = 0xffffffffu;  // keep decisions of this row's 32 keys (P dropout of the forward)
          if (p.drop.seed != nullptr) {
            const unsigned long long e0 = (((unsigned long long)bh * p.Lq + min(qi, p.Lq - 1)) * p.Lk32 + k0) >> 3;
            dropk = 0u;
#pragma unroll
            for (int t = 0; t < 4; ++t) dropk |= drop_keep8(p.drop, dseed, e0 + t) << (8 * t);
          }
          const float ik = p.drop.inv_keep;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float pv[2], dv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const bool keep = (mw >> (j + u)) & 1u, in = (inb >> (j + u)) & 1u;
              const float dm = ((dropk >> (j + u)) & 1u) ? ik : 0.f;   // dropout mask / keep probability
              float t = __uint_as_float(s[j + u]) * c1;
              t = keep ? t : t_masked;
              const float pr = (valid && in) ? ex2_approx_b(t - m_row) * inv_l : 0.f;
              pv[u] = pr * dm;                                            // dropped P feeds dV = P_drop^T dO
              dv[u] = (keep && in) ? pr * (__uint_as_float(d[j + u]) * dm - delta) * p.scale : 0.f;
            }
            pk_p[j >> 1] = pack_f16x2_sat(pv[0], pv[1]);
            pk_d[j >> 1] = pack_f16x2_sat(dv[0], dv[1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk_p[j] = pk_d[j] = 0u;
        }
        // the panels are still being read by the gradient MMAs of the previous tile until G_DONE: the first chunk's
        // TMEM loads, exponentials and dropout decisions above overlap them, only the stores wait
        if (cc == 0 && g > 0) {
          mbar_wait(bar(ABAR_G_DONE), (g - 1) & 1);
          tc_fence_after();  // (also orders the dQ drain below after the MMAs that produced it)
        }
        const uint32_t off = (uint32_t)(c >> 1) * C::PANEL + row * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t chunk = (uint32_t)((c & 1) * 4 + t) ^ sw;  // 128B swizzle: 16-B chunk ^= row % 8
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off + chunk * 16), "r"(pk_p[4 * t]),
                       "r"(pk_p[4 * t + 1]), "r"(pk_p[4 * t + 2]), "r"(pk_p[4 * t + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off + chunk * 16), "r"(pk_d[4 * t]),
                       "r"(pk_d[4 * t + 1]), "r"(pk_d[4 * t + 2]), "r"(pk_d[4 * t + 3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();  // panel stores (generic proxy) -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(bar(ABAR_P_FULL));
      if (i > 0) drain_dq(g - 1, i - 1);  // overlaps the MMAs of tile i (they write the other dQ buffer)
    }
    mbar_wait(bar(ABAR_G_DONE), (g - 1) & 1);  // g already counts this item's last tile
    tc_fence_after();
    drain_dq(g - 1, nq - 1);
    // ---- dV, dK: f16, head hd's column slice, keys of this tile
    if (half * 32 < DK) {
      const int key = kt * AB_T + row;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint32_t r[32];
        tc_ld32(tmem_base + (which == 0 ? C::COL_DV : C::COL_DK) + half * 32 + lane_off, r);
        tc_wait_ld();
        if (key < p.Lk) {
          __half* dst = which == 0 ? p.dv + ((size_t)b * p.Lk + key) * p.lddv : p.dk + ((size_t)b * p.Lk + key) * p.lddk;
          uint4* o = reinterpret_cast<uint4*>(dst + hd * DK + half * 32);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            o[t] = make_uint4(pack_f16x2_sat(__uint_as_float(r[8 * t]), __uint_as_float(r[8 * t + 1])),
                              pack_f16x2_sat(__uint_as_float(r[8 * t + 2]), __uint_as_float(r[8 * t + 3])),
                              pack_f16x2_sat(__uint_as_float(r[8 * t + 4]), __uint_as_float(r[8 * t + 5])),
                              pack_f16x2_sat(__uint_as_float(r[8 * t + 6]), __uint_as_float(r[8 * t + 7])));
        }
      }
    }
    }  // work items
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int DK>
static int launch_attn_bwd(const MtnAttnCoreBwdArgs& a, cudaStream_t st) {
  using C = AttnBwdCfg<DK>;
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(attn_core_bwd_tc_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    attr_set = true;
  }
  const TmSwizzle swz = DK == 64 ? TM_SWZ_128 : TM_SWZ_64;
  const uint64_t cols = (uint64_t)a.h * DK;
  CUtensorMap tq, tk, tv, tdo;
  int rc = make_tmap_3d_f16(&tq, a.q, cols, a.Lq, a.B, a.ldq, (uint64_t)a.Lq * a.ldq, DK, AB_T, swz);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tk, a.k, cols, a.Lk, a.B, a.ldk, (uint64_t)a.Lk * a.ldk, DK, AB_T, swz);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tv, a.v, cols, a.Lk, a.B, a.ldv, (uint64_t)a.Lk * a.ldv, DK, AB_T, swz);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tdo, a.dO, cols, a.Lq, a.B, a.lddo, (uint64_t)a.Lq * a.lddo, DK, AB_T, swz);
  if (rc) return rc;
  const int nkt = (a.Lk + AB_T - 1) / AB_T;
  const int n_items = nkt * a.h * a.B;
  AttnBwdParams p{n_items, nkt, a.mask_bits, a.mask_rows_q, mtn_mask_words(a.Lk), a.B, a.h, a.Lq, a.Lk, 1.0f / sqrtf((float)DK),
                  reinterpret_cast<const float2*>(a.stats), a.delta, a.dq, a.lddq,
                  reinterpret_cast<__half*>(a.dq_f16), a.lddq16, reinterpret_cast<__half*>(a.dk), a.lddk, reinterpret_cast<__half*>(a.dv), a.lddv,
                  DropCfg{reinterpret_cast<const unsigned long long*>(a.drop_seed), a.drop_site, a.drop_thresh,
                          a.drop_seed ? 1.f / (1.f - a.drop_thresh / 65536.f) : 1.f},
                  (a.Lk + 31) / 32 * 32};
  static int sms = 0;  // persistent CTAs: one per SM (512 TMEM columns, ~193 KB shared memory each)
  if (sms == 0) {
    int dev = 0;
    MTN_CHECK_CUDA(cudaGetDevice(&dev));
    MTN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  dim3 grid(n_items < sms ? n_items : sms);
  MTN_CHECK_CUDA(launch_kernel(attn_core_bwd_tc_kernel<DK>, grid, dim3(AB_THREADS), C::TOTAL, st, tq, tk, tv, tdo, p));
  return MTN_OK;
}

}  // namespace mtn

extern "C" int mtn_attn_core_bwd(const MtnAttnCoreBwdArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->q && a->k && a->v && a->dO && a->stats && a->delta && (a->dq || a->dq_f16) && a->dk && a->dv, MTN_E_ARG,
              "attn_core_bwd: NULL pointer");
  MTN_REQUIRE((a->dq != nullptr) != (a->dq_f16 != nullptr), MTN_E_ARG, "attn_core_bwd: give exactly one of dq (f32, accumulated) and dq_f16");
  MTN_REQUIRE(a->dq_f16 == nullptr || (a->Lk <= 128 && aligned16(a->dq_f16) && a->lddq16 % 8 == 0 && a->lddq16 >= a->h * a->d_k),
              MTN_E_SHAPE, "attn_core_bwd: dq_f16 needs Lk <= 128 (one key tile), 16-byte alignment, lddq16 >= h*d_k");
  MTN_REQUIRE(a->B > 0 && a->B <= 65535 && a->h > 0 && a->h <= 65535 && a->Lq > 0 && a->Lk > 0, MTN_E_SHAPE,
              "attn_core_bwd: B=%d h=%d Lq=%d Lk=%d", a->B, a->h, a->Lq, a->Lk);
  MTN_REQUIRE(a->d_k == 32 || a->d_k == 64, MTN_E_SHAPE, "attn_core_bwd: d_k=%d (supported: 32, 64)", a->d_k);
  const int w = a->h * a->d_k;
  MTN_REQUIRE(a->ldq >= w && a->ldk >= w && a->ldv >= w && a->lddo >= w && (!a->dq || a->lddq >= w) && a->lddk >= w && a->lddv >= w,
              MTN_E_SHAPE, "attn_core_bwd: leading dimension smaller than h*d_k=%d", w);
  MTN_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->lddo % 8 == 0 && a->lddk % 8 == 0 &&
                  a->lddv % 8 == 0 && (!a->dq || a->lddq % 4 == 0),
              MTN_E_ALIGN, "attn_core_bwd: leading dimensions must keep rows 16-byte aligned");
  MTN_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->dO) && (!a->dq || aligned16(a->dq)) &&
                  aligned16(a->dk) && aligned16(a->dv) && (reinterpret_cast<uintptr_t>(a->stats) & 7) == 0,
              MTN_E_ALIGN, "attn_core_bwd: pointers must be 16-byte aligned");
  MTN_REQUIRE(a->mask_bits == nullptr || a->mask_rows_q == 1 || a->mask_rows_q == a->Lq, MTN_E_SHAPE,
              "attn_core_bwd: mask_rows_q=%d must be 1 or Lq=%d", a->mask_rows_q, a->Lq);
  MTN_REQUIRE(a->drop_thresh < 65536u, MTN_E_ARG, "attn_core_bwd: drop_thresh=%u", a->drop_thresh);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->d_k == 64 ? launch_attn_bwd<64>(*a, st) : launch_attn_bwd<32>(*a, st);
}
