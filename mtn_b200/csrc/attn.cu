// Attention core on tcgen05:  O_h = softmax(mask(Q_h K_h^T / sqrt(d_k))) V_h
//
// Replaces attention() (mtn.py:221-231) and the head split / concat copies around it
// (mtn.py:257, 265-266).  One CTA per (128-query tile, head, batch element):
//   * Q, K, V head slices are fetched straight out of the packed projection buffers
//     with 3-D TMA tensor maps (columns, sequence, batch) -- no transpose copies;
//     out-of-range rows are zero-filled by TMA.
//   * S = Q K^T (128 x 128 keys per tile) and O (128 x d_k) accumulate in tensor
//     memory; tcgen05.mma is issued by one thread.
//   * softmax is f32 with one thread per query row (TMEM lane == row): masked
//     scores become the FINITE -1e9 of mtn.py:227, so a fully masked row yields the
//     uniform average the reference produces; keys beyond Lk get -inf (weight 0).
//   * P is rounded to f16 and handed to the tensor core THROUGH TENSOR MEMORY (PT, the default): the softmax
//     warps overwrite the first KT/2 columns of the tile's S buffer with the packed probabilities (tcgen05.st)
//     and P V is the TS form of tcgen05.mma (A operand in TMEM, V as MN-major shared-memory operand) -- no
//     swizzled shared-memory stores, no generic->async proxy fence, and no wait for the previous tile's P V
//     before P is written (each tile's P lives in its own S buffer).  The S buffer is recycled by issue order:
//     Q K^T of tile g+2 is issued after P V of tile g, and the tensor pipe executes in issue order.
//     (PT = false keeps the round-1 form: P through a swizzled shared-memory panel; MTN_B200_ATTN_PTMEM=0.)
//     The running max/sum rescale of O is done in TMEM (online softmax).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer,
// warps 2..5 softmax / epilogue.  K and V use separate single-slot buffers with their
// own full/empty barriers so the next K tile streams in while softmax / PV run.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int ATT_THREADS = 192;
constexpr int ATT_QT = 128;  // queries per CTA (UMMA M)

// KT = keys per tile (UMMA N of the S MMA): 96, or 64 for short memories (Lk <= 64: caption / query /
// auto-encoder sites).  96 rather than 128 so that TWO S accumulators plus O fit in 256 TMEM columns
// (2 x 96 + 64): the tensor core computes Q K^T of tile j+1 while the softmax warps are still on tile j,
// and two CTAs still share an SM.
template <int DK, int ATT_KT>
struct AttnCfg {
  static constexpr int ROWB = DK * 2;  // bytes per smem row of Q/K/V == swizzle span
  static constexpr int Q_BYTES = ATT_QT * ROWB;
  static constexpr int KV_BYTES = ATT_KT * ROWB;
  // P: [128 x 64-key] f16 panels, K-major, 128B swizzle; a 32-key remainder (KT = 96) is a [128 x 32] panel
  // with 64-byte rows and the 64B swizzle
  static constexpr int P_BYTES = (ATT_KT / 64) * ATT_QT * 128 + ((ATT_KT % 64) ? ATT_QT * 64 : 0);
  static constexpr int OFF_Q = 0;                       // two Q buffers (next work item's queries)
  static constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;     // two K buffers (keys of the next two tiles)
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;    // two V buffers
  static constexpr int OFF_P = (OFF_V + 2 * KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  // At most TWO CTAs may share an SM (2 x 256 TMEM columns): the request is padded above a third of the
  // shared memory so a third CTA can never be co-resident and sit in tcgen05.alloc behind a persistent peer.
  static constexpr int NEEDED = OFF_BAR + 192 + 1024;
  static constexpr int TOTAL = NEEDED > 78 * 1024 ? NEEDED : 78 * 1024;
  static constexpr uint64_t SWZ = (DK == 64) ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SBO = 8 * ROWB;  // 8-row swizzle atom
  static constexpr uint32_t TMEM_COLS = 256;  // S0: [0,KT)  S1: [KT,2KT)  O: [2KT, 2KT+DK)
  static constexpr uint32_t O_COL = 2 * ATT_KT;
  static_assert(2 * ATT_KT + DK <= 256, "TMEM budget");
  static_assert(ATT_KT % 64 == 0 || ATT_KT % 64 == 32, "P panels");
  static_assert((KV_BYTES % 1024) == 0 || DK == 32, "K/V buffers must stay swizzle-atom aligned");
};

struct AttnParams {
  int n_items, nqt;  // work items = B * h * nqt, ordered (b, head, q-tile) with the q-tile fastest
  const uint32_t* mask_bits;
  int mask_rows_q;
  int mask_words;
  int B, h, Lq, Lk;
  float scale;
  __half* out;
  int ldo;
  float2* stats;  // optional [B, h, Lq] {row maximum (log2 domain), 1 / row sum}: saved for the backward pass
  DropCfg drop;   // dropout of the probabilities (mtn.py:229-230): element index ((b*h + head)*Lq + q) * Lk32 + key
  int Lk32;       // Lk rounded up to a multiple of 32
};

enum {  // "+1": two barriers, one per buffer
  BAR_Q_FULL = 0 /* +1 */, BAR_Q_EMPTY = 2 /* +1 */, BAR_K_FULL = 4 /* +1 */, BAR_K_EMPTY = 6 /* +1 */,
  BAR_V_FULL = 8 /* +1 */, BAR_V_EMPTY = 10 /* +1 */, BAR_S_FULL = 12 /* +1 */, BAR_S_FREE = 14 /* +1 */, BAR_P_FULL = 16 /* +1 */,
  BAR_PV_DONE = 18 /* +1 */,
  BAR_COUNT = 20
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DK, int ATT_KT, bool DROP, bool PT>
__global__ void __launch_bounds__(ATT_THREADS, 2)
    attn_core_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                        const AttnParams p) {
  using C = AttnCfg<DK, ATT_KT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sQ = base + C::OFF_Q, sK = base + C::OFF_K, sV = base + C::OFF_V, sP = base + C::OFF_P;
  const uint32_t bars = base + C::OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * i; };
  // P V of (global) tile t has completed.  One barrier per tile parity: a waiter may lag the tensor core by a whole
  // tile (Q K^T runs one tile ahead, so a fast softmax warp can finish tile t while a slow one still holds up P V of
  // tile t-1); with a single barrier its parity test would then be satisfied by the phase of tile t-2.
  auto wait_pv = [&](uint32_t t) { mbar_wait(bar(BAR_PV_DONE + (t & 1)), (t >> 1) & 1); };
  const uint32_t tmem_slot = bars + 8u * BAR_COUNT;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 8 * BAR_COUNT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = (p.Lk + ATT_KT - 1) / ATT_KT;
  // Persistent CTA: work items it = blockIdx.x, += gridDim.x.  All barrier phases run on global
  // counters (item counter `n` for the Q buffers, tile counter `g` for everything else), so the
  // producer streams the next item's Q / K / V while the current item is still in softmax / PV.

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < BAR_COUNT; ++i)
      mbar_init(bar(i), (i == BAR_S_FREE || i == BAR_S_FREE + 1 || i == BAR_P_FULL || i == BAR_P_FULL + 1) ? 128u : 1u);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tO = tmem_base + C::O_COL;  // S buffers: tmem_base + (g & 1) * KT
  pdl_wait();

  if (warp == 0) {
    // -------------------------------------------------------------- TMA producer
    // Loads are issued in the order their buffers come free (Q K^T runs a tile ahead of P V):
    // K(0); then K(g+1), V(g) for every tile g -- the keys of a tile are requested two softmax periods
    // before its Q K^T, the values two periods before its P V, across work-item boundaries.
    if (lane == 0) {
      const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const uint32_t total = (uint32_t)my_items * (uint32_t)nt;
      auto coords = [&](uint32_t n, int& qt, int& hd, int& b) {
        const int it = blockIdx.x + (int)n * gridDim.x;
        qt = it % p.nqt, hd = (it / p.nqt) % p.h, b = it / (p.nqt * p.h);
      };
      auto load_k = [&](uint32_t g) {
        const uint32_t n = g / nt, j = g % nt, kb = g & 1, kph = (g >> 1) & 1;
        int qt, hd, b;
        coords(n, qt, hd, b);
        if (j == 0) {
          const uint32_t qb = n & 1;
          mbar_wait(bar(BAR_Q_EMPTY + qb), ((n >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(bar(BAR_Q_FULL + qb), C::Q_BYTES);
          tma_load_3d(sQ + qb * C::Q_BYTES, &tmQ, bar(BAR_Q_FULL + qb), hd * DK, qt * ATT_QT, b);
        }
        mbar_wait(bar(BAR_K_EMPTY + kb), kph ^ 1);
        mbar_arrive_expect_tx(bar(BAR_K_FULL + kb), C::KV_BYTES);
        tma_load_3d(sK + kb * C::KV_BYTES, &tmK, bar(BAR_K_FULL + kb), hd * DK, j * ATT_KT, b);
      };
      auto load_v = [&](uint32_t g) {
        const uint32_t n = g / nt, j = g % nt, vb = g & 1, vph = (g >> 1) & 1;
        int qt, hd, b;
        coords(n, qt, hd, b);
        mbar_wait(bar(BAR_V_EMPTY + vb), vph ^ 1);
        mbar_arrive_expect_tx(bar(BAR_V_FULL + vb), C::KV_BYTES);
        tma_load_3d(sV + vb * C::KV_BYTES, &tmV, bar(BAR_V_FULL + vb), hd * DK, j * ATT_KT, b);
      };
      if (total > 0) load_k(0);
      for (uint32_t g = 0; g < total; ++g) {
        if (g + 1 < total) load_k(g + 1);
        load_v(g);
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = make_idesc_f16(ATT_QT, ATT_KT, 0, 0);  // S = Q K^T, both K-major
    constexpr uint32_t idesc_o = make_idesc_f16(ATT_QT, DK, 0, 1);      // O += P V,  V is MN-major
    const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_items * (uint32_t)nt;  // tiles of this CTA, all items
    // Q K^T runs ONE tile ahead of P V (it only needs the keys and a free S buffer), so the scores of
    // tile g+1 -- possibly the first tile of the next work item -- are ready when the softmax warps
    // finish tile g.
    auto issue_qk = [&](uint32_t g) {
      const uint32_t n = g / nt, j = g % nt, qb = n & 1, sb = g & 1, ph2 = (g >> 1) & 1;
      if (j == 0) mbar_wait(bar(BAR_Q_FULL + qb), (n >> 1) & 1);
      mbar_wait(bar(BAR_K_FULL + sb), ph2);
      // softmax has finished reading this S buffer (tile g-2).  PT: the buffer also holds P of tile g-2; its P V was
      // issued before this Q K^T (after P_FULL of tile g-2, i.e. after the softmax warps' last access) and the
      // tensor pipe executes in issue order -- no barrier needed.
      if (!PT) mbar_wait(bar(BAR_S_FREE + sb), ph2 ^ 1);
      // ... and P V of tile g-2 has COMPLETED: a later MMA's accumulator write is not ordered behind an earlier MMA's
      // read of a tensor-memory A operand (measured: without this wait P is occasionally overwritten under the P V
      // that reads it).  P V of tile g-2 retired long ago in steady state -- the wait is free.
      if (PT && g >= 2) wait_pv(g - 2);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dq = make_smem_desc(sQ + qb * C::Q_BYTES, 16, C::SBO, C::SWZ);
        const uint64_t dk = make_smem_desc(sK + sb * C::KV_BYTES, 16, C::SBO, C::SWZ);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) tc_mma_f16(tmem_base + sb * ATT_KT, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(bar(BAR_K_EMPTY + sb));
        tc_commit(bar(BAR_S_FULL + sb));
        if (j == (uint32_t)nt - 1) tc_commit(bar(BAR_Q_EMPTY + qb));  // last use of this item's queries
      }
      __syncwarp();
    };
    if (total > 0) issue_qk(0);
    for (uint32_t g = 0; g < total; ++g) {
      if (g + 1 < total) issue_qk(g + 1);
      const uint32_t ph = g & 1, j = g % nt;
      mbar_wait(bar(BAR_V_FULL + ph), (g >> 1) & 1);
      // P_g is in place, O has been rescaled.  One barrier per S buffer: a softmax warp can run at most two tiles ahead
      // of the slowest one (S of tile g+2 exists only after P V of tile g was issued, i.e. after this phase completed),
      // so its arrival for tile g+2 can never be counted in the phase of tile g.
      mbar_wait(bar(BAR_P_FULL + ph), (g >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < ATT_KT / 16; ++kk) {
          // A: P panel kk/4 (64 keys per 128-B swizzled row), +32 B per 16 keys inside the row; the 32-key
          // remainder panel of KT = 96 has 64-B rows (64B swizzle, 512-B atoms)
          const bool rem = (ATT_KT % 64) != 0 && kk >= (ATT_KT / 64) * 4;
          const uint64_t dp = rem ? make_smem_desc(sP + (ATT_KT / 64) * (ATT_QT * 128) + (kk & 3) * 32, 16, 512, SWZ_64B)
                                  : make_smem_desc(sP + (kk >> 2) * (ATT_QT * 128) + (kk & 3) * 32, 16, 1024, SWZ_128B);
          // B: V rows [16 kk, 16 kk + 16): two 8-row swizzle atoms, d_k contiguous (MN-major)
          const uint64_t dv = make_smem_desc(sV + ph * C::KV_BYTES + kk * 16 * C::ROWB, ATT_KT * C::ROWB, C::SBO, C::SWZ);
          if (PT) tc_mma_f16_ts(tO, tmem_base + ph * ATT_KT + kk * 8, dv, idesc_o, (j | (uint32_t)kk) != 0);  // P: 16 keys = 8 columns
          else tc_mma_f16(tO, dp, dv, idesc_o, (j | (uint32_t)kk) != 0);
        }
        tc_commit(bar(BAR_V_EMPTY + ph));
        tc_commit(bar(BAR_PV_DONE + ph));
      }
      __syncwarp();
    }
  } else {
    // -------------------------------------------------------------- softmax + epilogue
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    // Softmax runs in the log2 domain: t = s * (scale * log2 e); masked t = -1e9 * log2 e (the
    // reference's finite -1e9, mtn.py:227); p = 2^(t - m).  Chunks (32 keys) whose keys are all
    // kept and in range -- the common case -- take a select-free fast path.
    constexpr float LOG2E = 1.4426950408889634f;
    const float c1 = p.scale * LOG2E;
    const float t_masked = -1e9f * LOG2E;
    const uint32_t sw = (uint32_t)(row & 7);
    constexpr int NCH = ATT_KT / 32;
    uint32_t g = 0;
    const unsigned long long dseed = (DROP && p.drop.seed != nullptr) ? __ldcg(p.drop.seed) : 0ull;

    // The mask words of a tile are requested one tile ahead (the first tile of the next work item during the
    // last tile of the current one): their latency would otherwise sit at the head of every tile.
    auto mask_row = [&](int it) -> const uint32_t* {
      if (p.mask_bits == nullptr || it >= p.n_items) return nullptr;
      const int qt = it % p.nqt, b = it / (p.nqt * p.h);
      const int mq = (p.mask_rows_q == 1) ? 0 : min(qt * ATT_QT + row, p.Lq - 1);
      return p.mask_bits + ((size_t)b * p.mask_rows_q + mq) * p.mask_words;
    };
    uint32_t mw_pref[NCH];
    auto fetch_mask = [&](const uint32_t* mr, int jn) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int k0 = jn * ATT_KT + c * 32;
        mw_pref[c] = (mr != nullptr && k0 < p.Lk) ? __ldcg(mr + (k0 >> 5)) : 0xffffffffu;  // coherent load (PDL, common.cuh)
      }
    };
    const uint32_t* mrow = mask_row(blockIdx.x);
    fetch_mask(mrow, 0);
    bool staged = false;  // this warp's TMA store of the previous item's output rows may still be reading the P buffer
    // Output staging tile of this warp (32 rows x d_k f16, TMA box layout): the first bytes of the warp's OWN 32 rows of
    // P panel 0 (128 B per row = 4 KB per warp), so that only the owning warp -- which waits for its TMA store to have
    // read the tile before it writes the next item's P -- ever touches it.  (For d_k = 64 the tile fills those 4 KB
    // exactly; for d_k = 32 it is their first half: a dense [row * 64 B] layout would alias other warps' P rows.)
    const uint32_t sO = sP + (uint32_t)q4 * (32u * 128u);

    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const int qt = it % p.nqt, hd = (it / p.nqt) % p.h, b = it / (p.nqt * p.h);
      const int qi = qt * ATT_QT + row;
      const uint32_t* mrow_next = mask_row(it + gridDim.x);
      float m_run = -CUDART_INF_F, l_run = 0.f;
      if (qt * ATT_QT + q4 * 32 >= p.Lq) {
        // every query row of this warp is padding (Lq <= 64 sites use half of the 128-row tile): skip the
        // softmax work, keep the barrier protocol alive.  The P rows of these queries stay whatever is in
        // shared memory; they only feed O rows that are never stored.
        for (int j = 0; j < nt; ++j, ++g) {
          const uint32_t ph = g & 1;
          mbar_wait(bar(BAR_S_FULL + ph), (g >> 1) & 1);
          if (!PT) {
            if (j > 0) wait_pv(g - 1);
            mbar_arrive(bar(BAR_S_FREE + ph));
          }
          mbar_arrive(bar(BAR_P_FULL + ph));
        }
        wait_pv(g - 1);
        mrow = mrow_next;
        fetch_mask(mrow, 0);
        continue;
      }

      for (int j = 0; j < nt; ++j, ++g) {
        const uint32_t ph = g & 1;
        const uint32_t tS = tmem_base + ph * ATT_KT;  // S buffer of this tile
        // A tile is "plain" when every key of it is inside the sequence and kept for every row of the warp (the
        // common case): select-free straight-line code on a register-resident score row.  Everything else
        // (ragged end of the sequence, padding tail, causal diagonal) takes the general chunk loop.
        uint32_t mwv[NCH];
        bool plain = !DROP && (j + 1) * ATT_KT <= p.Lk;  // (training-mode dropout: the chunk loop leaves registers for Philox)
        {
          uint32_t w = 0xffffffffu;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            mwv[c] = mw_pref[c];
            w &= mwv[c];
          }
          if (mrow != nullptr) plain = plain && __all_sync(0xffffffffu, w == 0xffffffffu);
        }
        if (j + 1 < nt) fetch_mask(mrow, j + 1);
        else fetch_mask(mrow_next, 0);
        mbar_wait(bar(BAR_S_FULL + ph), (g >> 1) & 1);
        tc_fence_after();
        auto store_chunk = [&](int c, const uint32_t(&e)[32]) {  // f16 P chunk -> tensor memory / swizzled K-major panel
          if (PT) {  // keys 32 c .. 32 c + 31 of this row -> columns [16 c, 16 c + 16) of the tile's S buffer
            uint32_t pk[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2_sat(__uint_as_float(e[2 * t]), __uint_as_float(e[2 * t + 1]));
            tc_st16(tS + lane_off + c * 16, pk);
            return;
          }
          const bool rem = (ATT_KT % 64) != 0 && c == NCH - 1;     // 32-key remainder panel: 64-B rows
          const uint32_t panel = sP + (c >> 1) * (ATT_QT * 128) + row * (rem ? 64 : 128);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            // 128B swizzle: 16-B chunk ^= row % 8;  64B swizzle: 16-B chunk ^= (row / 2) % 4
            const uint32_t chunk = rem ? ((uint32_t)t ^ ((uint32_t)(row >> 1) & 3u)) : ((uint32_t)((c & 1) * 4 + t) ^ sw);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(panel + chunk * 16),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t]), __uint_as_float(e[8 * t + 1]))),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t + 2]), __uint_as_float(e[8 * t + 3]))),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t + 4]), __uint_as_float(e[8 * t + 5]))),
                         "r"(pack_f16x2_sat(__uint_as_float(e[8 * t + 6]), __uint_as_float(e[8 * t + 7])))
                         : "memory");
          }
        };
        auto store_chunk_dyn = [&](int c, const uint32_t(&e)[32]) {  // same, chunk index known at run time
          if (PT) {
            uint32_t pk[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) pk[t] = pack_f16x2_sat(__uint_as_float(e[2 * t]), __uint_as_float(e[2 * t + 1]));
            tc_st16(tS + lane_off + c * 16, pk);
            return;
          }
          const bool rem = (ATT_KT % 64) != 0 && c == NCH - 1;
          const uint32_t panel = sP + (c >> 1) * (ATT_QT * 128) + row * (rem ? 64 : 128);
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
        auto drop_chunk = [&](int c, uint32_t(&e)[32]) {
          // dropout AFTER the softmax: the row sum keeps every key, only the P V operand is thinned
          const int k0 = j * ATT_KT + c * 32;
          const unsigned long long e0 =
              ((((unsigned long long)b * p.h + hd) * p.Lq + min(qi, p.Lq - 1)) * p.Lk32 + k0) >> 3;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t kb = drop_keep8(p.drop, dseed, e0 + t);
#pragma unroll
            for (int u = 0; u < 8; ++u)
              e[8 * t + u] = ((kb >> u) & 1u) ? __float_as_uint(__uint_as_float(e[8 * t + u]) * p.drop.inv_keep) : 0u;
          }
        };
        // P buffer and O accumulator are in use by PV of the previous tile of this item until it retires; across
        // items the P buffer stages the previous item's output tile until its TMA store has read it
        auto wait_p_buffer = [&]() {
          if (PT) return;  // P of this tile goes to its own S buffer; the staging tile is waited for in the epilogue
          if (j > 0) {
            wait_pv(g - 1);
            tc_fence_after();
          } else if (staged) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
        };
        float l4[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent partial row sums
        float m_new;
        bool rescale = false;
        // lazy rescale: the running maximum only follows the tile maximum when some row of the warp grew by
        // more than 2^8 (p <= 256 stays far inside f16 / f32 range and the common factor cancels in O / l);
        // otherwise the O accumulator in tensor memory is left alone.
        auto pick_max = [&](float m_tile) {
          m_new = fmaxf(m_run, m_tile);
          if (j > 0) {
            rescale = __any_sync(0xffffffffu, m_new - m_run > 8.f);
            if (!rescale) m_new = m_run;
          }
        };
        if (plain) {
          // ---- the whole score row of the tile goes to registers with ONE TMEM round trip (all chunk loads in
          // flight, one wait) and the S buffer is released to Q K^T of tile g+2 right away
          uint32_t r[NCH][32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc_ld32(tS + lane_off + c * 32, r[c]);
          tc_wait_ld();
          if (!PT) {
            tc_fence_before();
            mbar_arrive(bar(BAR_S_FREE + ph));
          }
          float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r[c][i]));  // 4 independent chains
          }
          pick_max(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c1);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float e = ex2_approx(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
              l4[i & 3] += e;
              r[c][i] = __float_as_uint(e);
            }
          }
          wait_p_buffer();
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (DROP && p.drop.seed != nullptr) drop_chunk(c, r[c]);
            store_chunk(c, r[c]);
          }
        } else if (!DROP && ATT_KT == 64) {
          // (64-key tiles only, i.e. the single-tile sites of short memories: measured 8.4 -> 7.5 us there, but with three
          // chunks per tile the unrolled code of this path costs more in instruction-cache misses than it saves --
          // causal self-attention 16.1 -> 22.1 us -- so 96-key tiles keep the compact chunk loop below.)
          // ---- masked / ragged tile (padding tail, end of the sequence, causal diagonal, holes), still register-resident:
          // ONE TMEM round trip, then per 32-key chunk a warp-uniform choice between the select-free code of a plain tile
          // (every row keeps every key), a constant (no row keeps any key: every in-range score is the -1e9 of
          // mtn.py:227, keys beyond Lk weigh 0) and per-key selects.
          uint32_t r[NCH][32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc_ld32(tS + lane_off + c * 32, r[c]);
          uint32_t inbv[NCH], kmv[NCH];
          int kind[NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int nvalid = p.Lk - (j * ATT_KT + c * 32);  // keys of this chunk inside the sequence (warp-uniform)
            inbv[c] = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
            kmv[c] = mwv[c] & inbv[c];
            kind[c] = __all_sync(0xffffffffu, kmv[c] == 0xffffffffu) ? 0 : (__all_sync(0xffffffffu, kmv[c] == 0u) ? 1 : 2);
          }
          tc_wait_ld();
          if (!PT) {
            tc_fence_before();
            mbar_arrive(bar(BAR_S_FREE + ph));
          }
          float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
          uint32_t masked_any = 0u;  // some in-range key of this row is masked
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
          const float pm = ex2_approx(t_masked - m_new);  // weight of a masked key: 0 unless the whole row is masked so far
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (kind[c] == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float e = ex2_approx(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
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
                const float x = ex2_approx(fmaf(__uint_as_float(r[c][i]), c1, -m_new));
                const float e = ((kmv[c] >> i) & 1u) ? x : (((inbv[c] >> i) & 1u) ? pm : 0.f);
                l4[i & 3] += e;
                r[c][i] = __float_as_uint(e);
              }
            }
          }
          wait_p_buffer();
#pragma unroll
          for (int c = 0; c < NCH; ++c) store_chunk(c, r[c]);
        } else {
          // ---- general tile (96-key tiles, and training-mode dropout: Philox needs the registers): two passes over the S
          // buffer, one 32-key chunk at a time, compact code.  Per chunk (warp-uniform):
          // beyond the sequence -> zeros; no kept key for ANY row of the warp (padding tail of a key-padding mask)
          // -> every in-range score is the constant -1e9, nothing to read from TMEM; otherwise per-key selects.
          float m_tile = -CUDART_INF_F;
          auto chunk_mask = [&](int c) { return c == 0 ? mwv[0] : (c == 1 ? mwv[1] : mwv[NCH - 1]); };
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int nvalid = p.Lk - (j * ATT_KT + c * 32);  // keys of this chunk inside the sequence (warp-uniform)
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
            if (__all_sync(0xffffffffu, (mw & inb) == 0xffffffffu)) {  // every row keeps every key of the chunk
              float mx[4] = {__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3])};
#pragma unroll
              for (int i = 4; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(r[i]));
              m_tile = fmaxf(m_tile, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c1);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float t = __uint_as_float(r[i]) * c1;
                t = ((mw >> i) & 1u) ? t : t_masked;
                t = ((inb >> i) & 1u) ? t : -CUDART_INF_F;
                m_tile = fmaxf(m_tile, t);
              }
            }
          }
          pick_max(m_tile);
          wait_p_buffer();
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int nvalid = p.Lk - (j * ATT_KT + c * 32);
            const uint32_t inb = nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
            const uint32_t mw = chunk_mask(c);
            uint32_t e[32];
            if (nvalid > 0 && __all_sync(0xffffffffu, (mw & inb) == 0u)) {
              // p = 2^(-1e9 log2e - m) is one value per row (0 unless the whole row has been masked so far,
              // then 1 -> uniform average)
              const float pm = ex2_approx(t_masked - m_new);
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = ((inb >> i) & 1u) ? __float_as_uint(pm) : 0u;
              l4[0] += pm * (float)__popc(inb);
            } else if (nvalid > 0) {
              tc_ld32(tS + lane_off + c * 32, e);
              tc_wait_ld();
              if (__all_sync(0xffffffffu, (mw & inb) == 0xffffffffu)) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float x = ex2_approx(fmaf(__uint_as_float(e[i]), c1, -m_new));
                  l4[i & 3] += x;
                  e[i] = __float_as_uint(x);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  float t = __uint_as_float(e[i]) * c1;
                  t = ((mw >> i) & 1u) ? t : t_masked;
                  t = ((inb >> i) & 1u) ? t : -CUDART_INF_F;
                  const float x = ex2_approx(t - m_new);
                  l4[i & 3] += x;
                  e[i] = __float_as_uint(x);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = 0u;
            }
            if (DROP && p.drop.seed != nullptr && nvalid > 0) drop_chunk(c, e);
            store_chunk_dyn(c, e);
          }
          if (!PT) {
            tc_fence_before();
            mbar_arrive(bar(BAR_S_FREE + ph));  // this S buffer may be overwritten by Q K^T of tile g+2
          }
        }
        const float alpha = ex2_approx(m_run - m_new);  // 1 when the maximum was kept; 0 for j == 0 (l_run is 0)
        l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
        m_run = m_new;
        if (rescale) {
          // rescale the running O accumulator in tensor memory (warp-uniform branch)
          if (PT) {  // (without PT every tile has already waited for the previous P V before writing P)
            wait_pv(g - 1);
            tc_fence_after();
          }
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
        if (PT) tc_wait_st();               // P (and a rescaled O) have landed in tensor memory
        else fence_proxy_async_smem();      // P (generic-proxy stores) -> visible to the tensor core
        tc_fence_before();
        mbar_arrive(bar(BAR_P_FULL + ph));
      }
      // ---- epilogue: O / l -> f16 -> this warp's 32 rows of the (now idle) P buffer in the TMA tile layout -> one
      // TMA store per warp into head hd's column slice of the output (rows beyond Lq are clipped by the tensor
      // map); no CTA-wide synchronisation
      wait_pv(g - 1);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      if (PT && staged) {  // this warp's previous output tile must have left the staging buffer
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      if (p.stats != nullptr && qi < p.Lq) p.stats[((size_t)b * p.h + hd) * p.Lq + qi] = make_float2(m_run, inv_l);
#pragma unroll
      for (int c = 0; c < DK / 32; ++c) {
        uint32_t r[32];
        tc_ld32(tO + lane_off + c * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          // 128B swizzle (d_k = 64): 16-B chunk ^= row % 8;  64B swizzle (d_k = 32): chunk ^= (row / 2) % 4
          const uint32_t chunk = DK == 64 ? ((uint32_t)(c * 4 + t) ^ sw) : ((uint32_t)t ^ ((uint32_t)(row >> 1) & 3u));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sO + lane * C::ROWB + chunk * 16),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t]) * inv_l, __uint_as_float(r[8 * t + 1]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 2]) * inv_l, __uint_as_float(r[8 * t + 3]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 4]) * inv_l, __uint_as_float(r[8 * t + 5]) * inv_l)),
                       "r"(pack_f16x2_sat(__uint_as_float(r[8 * t + 6]) * inv_l, __uint_as_float(r[8 * t + 7]) * inv_l))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmO, sO, hd * DK, qt * ATT_QT + q4 * 32, b);
        tma_store_commit();
      }
      staged = true;
      mrow = mrow_next;
    }
    if (lane == 0) tma_store_wait_read();  // (the kernel boundary completes the writes)
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int DK, int ATT_KT, bool DROP, bool PT>
static int launch_attn_v(const MtnAttnCoreArgs& a, cudaStream_t st) {
  using C = AttnCfg<DK, ATT_KT>;
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(attn_core_tc_kernel<DK, ATT_KT, DROP, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::TOTAL));
    attr_set = true;
  }
  const TmSwizzle swz = DK == 64 ? TM_SWZ_128 : TM_SWZ_64;
  const uint64_t cols = (uint64_t)a.h * DK;
  CUtensorMap tq, tk, tv;
  auto bstride = [](long long given, int L, int ld) { return given > 0 ? (uint64_t)given : (uint64_t)L * ld; };
  int rc = make_tmap_3d_f16(&tq, a.q, cols, a.Lq, a.B, a.ldq, bstride(a.q_batch_stride, a.Lq, a.ldq), DK, ATT_QT, swz);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tk, a.k, cols, a.Lk, a.B, a.ldk, bstride(a.k_batch_stride, a.Lk, a.ldk), DK, ATT_KT, swz);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tv, a.v, cols, a.Lk, a.B, a.ldv, bstride(a.v_batch_stride, a.Lk, a.ldv), DK, ATT_KT, swz);
  if (rc) return rc;
  CUtensorMap to;
  rc = make_tmap_3d_f16(&to, a.out, cols, a.Lq, a.B, a.ldo, bstride(a.o_batch_stride, a.Lq, a.ldo), DK, 32, swz);
  if (rc) return rc;
  const int nqt = (a.Lq + ATT_QT - 1) / ATT_QT;
  const int n_items = nqt * a.h * a.B;
  AttnParams p{n_items, nqt, a.mask_bits, a.mask_rows_q, mtn_mask_words(a.Lk), a.B, a.h, a.Lq, a.Lk,
               1.0f / sqrtf((float)DK), reinterpret_cast<__half*>(a.out), a.ldo, reinterpret_cast<float2*>(a.stats),
               DropCfg{reinterpret_cast<const unsigned long long*>(a.drop_seed), a.drop_site, a.drop_thresh,
                       a.drop_seed ? 1.f / (1.f - a.drop_thresh / 65536.f) : 1.f},
               (a.Lk + 31) / 32 * 32};
  const int slots = 2 * stream_sm_count(st);  // resident CTAs: 2 per SM of the stream's SM partition
  dim3 grid(n_items < slots ? n_items : slots);
  MTN_CHECK_CUDA(launch_kernel(attn_core_tc_kernel<DK, ATT_KT, DROP, PT>, grid, dim3(ATT_THREADS), C::TOTAL, st, tq, tk, tv, to, p));
  return MTN_OK;
}

template <int DK, int ATT_KT, bool DROP>
static int launch_attn(const MtnAttnCoreArgs& a, cudaStream_t st) {
  static int pt = -1;  // MTN_B200_ATTN_PTMEM=0: P through shared memory (round-1 form)
  if (pt < 0) {
    const char* e = getenv("MTN_B200_ATTN_PTMEM");
    pt = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return pt ? launch_attn_v<DK, ATT_KT, DROP, true>(a, st) : launch_attn_v<DK, ATT_KT, DROP, false>(a, st);
}

static int validate_attn(const MtnAttnCoreArgs* a) {
  MTN_REQUIRE(a && a->q && a->k && a->v && a->out, MTN_E_ARG, "attn_core: NULL pointer");
  MTN_REQUIRE(a->B > 0 && a->h > 0 && a->Lq > 0 && a->Lk > 0, MTN_E_SHAPE, "attn_core: B=%d h=%d Lq=%d Lk=%d",
              a->B, a->h, a->Lq, a->Lk);
  MTN_REQUIRE(a->d_k == 32 || a->d_k == 64, MTN_E_SHAPE, "attn_core: d_k=%d (supported: 32, 64)", a->d_k);
  const int w = a->h * a->d_k;
  MTN_REQUIRE(a->ldq >= w && a->ldk >= w && a->ldv >= w && a->ldo >= w, MTN_E_SHAPE,
              "attn_core: leading dimension smaller than h*d_k=%d", w);
  MTN_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, MTN_E_ALIGN,
              "attn_core: leading dimensions must be multiples of 8 elements");
  MTN_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->out), MTN_E_ALIGN,
              "attn_core: pointers must be 16-byte aligned");
  MTN_REQUIRE(a->mask_bits == nullptr || a->mask_rows_q == 1 || a->mask_rows_q == a->Lq, MTN_E_SHAPE,
              "attn_core: mask_rows_q=%d must be 1 or Lq=%d", a->mask_rows_q, a->Lq);
  MTN_REQUIRE(a->drop_thresh < 65536u, MTN_E_ARG, "attn_core: drop_thresh=%u", a->drop_thresh);
  MTN_REQUIRE(a->q_batch_stride >= 0 && a->k_batch_stride >= 0 && a->v_batch_stride >= 0 && a->o_batch_stride >= 0 &&
                  a->q_batch_stride % 8 == 0 && a->k_batch_stride % 8 == 0 && a->v_batch_stride % 8 == 0 &&
                  a->o_batch_stride % 8 == 0,
              MTN_E_ALIGN, "attn_core: batch strides must be non-negative multiples of 8 elements");
  MTN_REQUIRE(a->stats == nullptr || (a->q_batch_stride == 0 && a->o_batch_stride == 0), MTN_E_ARG,
              "attn_core: saved statistics (training) need the dense batch layout");
  return MTN_OK;
}

// ----------------------------------------------------------------------------
// self-check kernel (tests only): one warp per (batch, head, query) row, same
// arithmetic contract (f16 operands, f32 scores/softmax, P rounded to f16 for PV).
// ----------------------------------------------------------------------------
__global__ void attn_core_check_kernel(const __half* q, int ldq, const __half* k, int ldk, const __half* v,
                                       int ldv, AttnParams p, int dk, size_t sq, size_t sk, size_t sv, size_t so) {
  extern __shared__ float sc[];  // Lk scores
  const int qi = blockIdx.x, hd = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x;
  const __half* qr = q + (size_t)b * sq + (size_t)qi * ldq + hd * dk;
  const uint32_t* mrow = nullptr;
  if (p.mask_bits) mrow = p.mask_bits + ((size_t)b * p.mask_rows_q + (p.mask_rows_q == 1 ? 0 : qi)) * p.mask_words;
  float mx = -CUDART_INF_F;
  for (int j = lane; j < p.Lk; j += 32) {
    const __half* kr = k + (size_t)b * sk + (size_t)j * ldk + hd * dk;
    float s = 0.f;
    for (int c = 0; c < dk; ++c) s = fmaf(__half2float(qr[c]), __half2float(kr[c]), s);
    s *= p.scale;
    if (mrow && !((mrow[j >> 5] >> (j & 31)) & 1u)) s = -1e9f;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < p.Lk; j += 32) {
    const float e = __expf(sc[j] - mx);
    sum += e;
    sc[j] = __half2float(__float2half_rn(e));
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  for (int c = lane; c < dk; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < p.Lk; ++j)
      acc = fmaf(sc[j], __half2float(v[(size_t)b * sv + (size_t)j * ldv + hd * dk + c]), acc);
    const uint32_t pk = pack_f16x2_sat(acc / sum, 0.f);
    p.out[(size_t)b * so + (size_t)qi * p.ldo + hd * dk + c] = __ushort_as_half((unsigned short)(pk & 0xffff));
  }
}

}  // namespace mtn

extern "C" int mtn_attn_core_fwd(const MtnAttnCoreArgs* a, void* stream) {
  int rc = mtn::validate_attn(a);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->drop_seed != nullptr) {      // training: probability dropout compiled in
    if (a->Lk <= 64) return a->d_k == 64 ? mtn::launch_attn<64, 64, true>(*a, st) : mtn::launch_attn<32, 64, true>(*a, st);
    return a->d_k == 64 ? mtn::launch_attn<64, 96, true>(*a, st) : mtn::launch_attn<32, 96, true>(*a, st);
  }
  if (a->Lk <= 64) return a->d_k == 64 ? mtn::launch_attn<64, 64, false>(*a, st) : mtn::launch_attn<32, 64, false>(*a, st);
  return a->d_k == 64 ? mtn::launch_attn<64, 96, false>(*a, st) : mtn::launch_attn<32, 96, false>(*a, st);
}

extern "C" int mtn_check_attn_core_fwd(const MtnAttnCoreArgs* a, void* stream) {
  int rc = mtn::validate_attn(a);
  if (rc) return rc;
  mtn::AttnParams p{0, 0, a->mask_bits, a->mask_rows_q, mtn_mask_words(a->Lk), a->B, a->h, a->Lq, a->Lk,
                    1.0f / sqrtf((float)a->d_k), reinterpret_cast<__half*>(a->out), a->ldo, nullptr,
                    mtn::DropCfg{nullptr, 0, 0, 1.f}, 0};
  dim3 grid(a->Lq, a->h, a->B);
  auto bs = [](long long given, int L, int ld) { return given > 0 ? (size_t)given : (size_t)L * ld; };
  mtn::attn_core_check_kernel<<<grid, 32, a->Lk * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(a->q), a->ldq, reinterpret_cast<const __half*>(a->k), a->ldk,
      reinterpret_cast<const __half*>(a->v), a->ldv, p, a->d_k, bs(a->q_batch_stride, a->Lq, a->ldq),
      bs(a->k_batch_stride, a->Lk, a->ldk), bs(a->v_batch_stride, a->Lk, a->ldv), bs(a->o_batch_stride, a->Lq, a->ldo));
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}
