// Fused linear layer on tcgen05:  C = act(A W^T + bias) + addend
//
// Replaces the nn.Linear call sites of the MTN hot path (mtn.py:256-258 Q/K/V
// projections, :267 output projection, :280 FFN w_1/w_2, :35 video encoder).
// A [M,K] and W [N,K] are both K-major f16, so the GEMM is the "TN" form UMMA
// consumes directly: 128 x BN x 64 tiles (BN = 256 or 128) are staged by TMA (128-byte
// swizzle) into a STAGES-deep shared-memory ring, one elected thread issues 128xBNx16
// tcgen05.mma (f32 accumulate in one of TWO tensor-memory buffers) and four epilogue
// warps drain the other buffer with tcgen05.ld, apply bias / ReLU / residual-or-positional
// addend and store f32 and/or f16 with line-coalesced accesses.  CTAs are persistent.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (warp w may touch TMEM lanes 32*(w%4) .. +31; two warps per quarter).
//
// The same kernel is the backward of those call sites (mtn_linear_dgrad / mtn_linear_wgrad): an operand
// may be "MN-major" (stored [k, mn], mn contiguous), which UMMA consumes through the transpose bits of the
// instruction descriptor, so   dX = dY W   reads W in its forward [out, in] layout and   dW = dY^T X
// reads dY and X in their forward [rows, features] layouts -- no transposed copies exist anywhere.
// Weight gradients (small output, long contraction) run split-K: the contraction is cut over CTAs and the
// partial tiles are accumulated into the f32 gradient with L2 reductions (red.global.add).
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

struct GemmEpi {
  const float* bias;
  int act;
  const float* addend;
  int ld_add;
  int add_period;
  float* out32;
  int ld32;
  __half* out16;
  int ld16;
  const float* alpha;      // device scalar multiplied into the accumulator first (NULL: 1)
  const __half* relu_mask; // [M, ld_mask] forward activation: result zeroed where it is <= 0 (ReLU backward)
  int ld_mask;
  int accumulate;          // out32 += result (red.global.add), required for split-K
  int out16_pre_add;       // out16 receives the value BEFORE the addend (video encoder: relu(.) without PE)
  int multimem;            // accumulate mode: out32 / colsum_a are NVLS multicast addresses (multimem.red)
  float* colsum_a;         // BIAS kernels: colsum_a[m] += alpha * sum_k A(m, k)  (the bias gradient of a wgrad GEMM)
  float mask_scale;        // multiplies the elements that pass relu_mask (1/(1-p) of the dropout after the ReLU); 0 = 1
  DropCfg drop;            // dropout of the result (element index = row * N + col)
  int drop_after_add;      // 0: before the addend (SublayerConnection, mtn.py:127)  1: after it (PositionalEncoding, mtn.py:309)
  // strided batch (element strides between consecutive problems; batch == 1: unused)
  long long s_bias, s_add, s_out32, s_out16;
  int tma_red;             // in-place residual / accumulating f32 output goes out as TMA reduce-add stores (tmC)
};

constexpr int BM = 128;
constexpr int BK = 64;  // 64 f16 = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int STAGE_TILE_BYTES = 32 * 32 * 4;  // per-epilogue-warp 32x32 f32 transpose buffer

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int XPOSE_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = XPOSE_OFF + 8 * STAGE_TILE_BYTES;
  static constexpr int NBARS = 2 * STAGES + 4;
  static constexpr int ONES_OFF = (BAR_OFF + 8 * NBARS + 16 + 255) / 256 * 256;  // 1 KB of f16 1.0 (bias-gradient MMA)
  static constexpr int TOTAL = ONES_OFF + 1024 + 1024;  // + 1 KB alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// Persistent kernel: each CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  (n fastest, so
// the CTAs running at the same time share the same A row panels and the whole of W in L2).  The TMA
// ring and the two TMEM accumulator buffers run across tile boundaries: while the epilogue warps drain
// accumulator i the tensor core is already filling accumulator i+1.
//
// CL > 1: thread-block clusters of CL CTAs along M.  The CTAs of a cluster work on the same n-block and
// CL consecutive m-blocks at the same time; each loads 1/CL of the W tile and TMA-multicasts it to all of
// them, so W crosses L2->SM once per cluster instead of once per CTA (the 128xBN tiles are L2-bandwidth
// bound otherwise).  A slot is refilled only after every CTA of the cluster has consumed it: the MMA
// warp's tcgen05.commit arrives on the "empty" barrier of all CL CTAs (multicast commit).
//
// A_MN / B_MN: the operand is MN-major.  Its 64-row k-block is then staged as (BM or BN)/64 TMA boxes of
// [64 k-rows x 64 mn-columns] (128-byte swizzled rows), i.e. the canonical UMMA MN-major layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units with LBO = 8192 B (next 64 mn), SBO = 1024 B (next 8 k).
// ksplit > 1: tile index also enumerates contraction slices of kb_per_split k-blocks (split-K).
// DROP: compile the dropout of the result into the epilogue (training forward only; the inference kernels carry
// none of its registers or branches).
// BIAS: the weight-gradient form additionally accumulates the row sums of its A operand (dY^T), i.e. the bias
// gradient, on the tensor core: one extra N=16 MMA per k-step against a constant tile of ones (any layout of an
// all-ones tile is the same tile), 16 more TMEM columns per accumulator buffer, one atomic per output row.
template <int BN, int STAGES, int CL, int A_MN, int B_MN, int DROP, int BIAS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const GemmEpi epi0, int M, int N, int K, int tiles_n, int tiles_per_batch, int num_tiles,
                       int tiles_mn, int kb_per_split) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle-128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + L::BAR_OFF;
  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto bar_acc_full = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto bar_acc_empty = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * L::NBARS;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * L::NBARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb_all = (K + BK - 1) / BK;  // K tail: TMA zero-fills columns >= K
  static_assert(CL == 1 || (A_MN == 0 && B_MN == 0), "multicast path is K-major only");
  // k-block range of tile-in-batch index tl (split-K slice tl / tiles_mn)
  auto kb_range = [&](int tl, int& kb0, int& kb1) {
    kb0 = (tl / tiles_mn) * kb_per_split;
    kb1 = min(nkb_all, kb0 + kb_per_split);
  };
  // num_tiles counts cluster-level "super tiles" (CL m-blocks x 1 n-block); this CTA takes m-block
  // (mg * CL + rank) of each.  Out-of-range m-blocks still take part in the multicast and barriers:
  // their A rows are zero-filled by TMA and their stores are predicated off.
  const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0u;
  const int first_tile = blockIdx.x / CL, tile_stride = gridDim.x / CL;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (epi0.tma_red) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full(s), 1);   // producer arrive + tx bytes
      mbar_init(bar_empty(s), CL);  // tcgen05.commit of every CTA in the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);   // tcgen05.commit after the last k-block of a tile
      mbar_init(bar_acc_empty(b), 8);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  __syncwarp();  // cluster barriers are .aligned: every warp must arrive converged
  static_assert(!BIAS || (A_MN == 1 && B_MN == 1 && BN == 128), "BIAS: weight-gradient form, BN = 128");
  // BIAS kernels (one-wave split-K launches: a CTA rarely sees a second tile) keep ONE accumulator buffer, so that
  // accumulator + row sums still fit 256 TMEM columns and two CTAs of consecutive launches can overlap on an SM (PDL)
  constexpr uint32_t NBUF = BIAS ? 1u : 2u;
  constexpr uint32_t TMEM_COLS = BIAS ? 256u : 2u * BN;
  constexpr uint32_t BIAS_COL = BN;  // BIAS: columns [BN, BN + 16)
  if (BIAS && warp == 2) {  // the constant ones tile
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + L::ONES_OFF);
#pragma unroll
    for (int i = 0; i < 8; ++i) ones[lane + 32 * i] = 0x3C003C00u;  // two f16 1.0
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // peers' barriers must exist before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();  // everything above overlapped the previous kernel; global memory is touched below

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;  // k-blocks issued so far (ring position across tiles)
      for (int t = first_tile; t < num_tiles; t += tile_stride) {
        const int bt = t / tiles_per_batch, tl = t % tiles_per_batch;  // problem of the strided batch, tile in it
        const int tmn = tl % tiles_mn;
        const int m0 = ((tmn / tiles_n) * CL + rank) * BM, n0 = (tmn % tiles_n) * BN;
        int kb0, kb1;
        kb_range(tl, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_empty(s), ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full(s), L::STAGE_BYTES);
          const uint32_t sA = base + s * L::STAGE_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_3d(sA + j * 8192, &tmA, bar_full(s), m0 + 64 * j, kb * BK, bt);
          } else {
            tma_load_3d(sA, &tmA, bar_full(s), kb * BK, m0, bt);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_3d(sA + L::A_BYTES + j * 8192, &tmB, bar_full(s), n0 + 64 * j, kb * BK, bt);
          } else if (CL == 1) {
            tma_load_3d(sA + L::A_BYTES, &tmB, bar_full(s), kb * BK, n0, bt);
          } else {  // my 1/CL of the W tile, delivered to every CTA of the cluster
            constexpr int SL = BN / CL;
            tma_load_3d_mc(sA + L::A_BYTES + rank * (SL * BK * 2), &tmB, bar_full(s), kb * BK, n0 + rank * SL, bt, kMask);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, A_MN, B_MN);
    uint32_t it = 0, lt = 0;  // ring position, local tile counter
    for (int t = first_tile; t < num_tiles; t += tile_stride, ++lt) {
      const uint32_t buf = lt % NBUF;
      mbar_wait(bar_acc_empty(buf), ((lt / NBUF) & 1) ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      const uint32_t d_bias = tmem_base + BIAS_COL;
      // BIAS: the n-tiles of a row block share the row-sum work -- tile j takes the k-blocks with kb % tiles_n == j
      const int nidx = ((t % tiles_per_batch) % tiles_mn) % tiles_n;
      bool bias_started = false;
      int kb0, kb1;
      kb_range(t % tiles_per_batch, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(bar_full(s), (it / STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sA = base + s * L::STAGE_BYTES;
          // K-major: +32 B along K inside the swizzled row (= +2 encoded) per 16-wide k-step.
          // MN-major: a k-step is 16 k-rows of 128 B = +2048 B (= +128 encoded).
          const uint64_t da = A_MN ? make_smem_desc(sA, 8192, 1024, SWZ_128B) : make_smem_desc(sA, 16, 1024, SWZ_128B);
          const uint64_t db = B_MN ? make_smem_desc(sA + L::A_BYTES, 8192, 1024, SWZ_128B)
                                   : make_smem_desc(sA + L::A_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma_f16(d_tmem, da + (A_MN ? 128 : 2) * k, db + (B_MN ? 128 : 2) * k, idesc, ((kb - kb0) | k) != 0);
          if (BIAS && kb % tiles_n == nidx) {
            constexpr uint32_t idesc_b = make_idesc_f16(BM, 16, A_MN, 1);
            const uint64_t dones = make_smem_desc(base + L::ONES_OFF, 128, 256, SWZ_NONE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_f16(d_bias, da + (A_MN ? 128 : 2) * k, dones, idesc_b, (bias_started || k != 0) ? 1u : 0u);
            bias_started = true;
          }
          if (CL == 1) tc_commit(bar_empty(s));  // frees the stage when these MMAs retire
          else tc_commit_mc(bar_empty(s), kMask);
          if (kb == kb1 - 1) tc_commit(bar_acc_full(buf));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (8 warps)
    // TMEM gives each thread one accumulator ROW (32 columns per load).  Each warp transposes its
    // 32x32 block through a private XOR-swizzled shared-memory tile so that all global traffic is
    // line-coalesced (8 lanes cover 128 contiguous bytes of a row, a warp covers 4 rows), and all
    // element-wise work (bias, ReLU, residual) happens in that coalesced layout with its global
    // operands requested before the TMEM load.  Two warps share each TMEM lane quarter and take
    // alternate 32-column chunks.
    const int ew = warp - 2;
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int half = ew >> 2;  // 0: even chunks, 1: odd chunks
    float4* xp = reinterpret_cast<float4*>(smem + L::XPOSE_OFF + ew * STAGE_TILE_BYTES);
    // In-place residual (x += A W^T + b, the common SublayerConnection case): the add is done by L2
    // reductions (red.global.add.v4.f32) -- the 16 MB residual read that makes these launches
    // L2-bandwidth bound disappears and each element still receives exactly one f32 add.
    const bool red_add = epi0.accumulate != 0 ||
                         (epi0.addend != nullptr && epi0.addend == epi0.out32 && epi0.add_period == 0 &&
                          epi0.ld_add == epi0.ld32 && epi0.out16 == nullptr && epi0.s_add == epi0.s_out32);
    const bool has_add = epi0.addend != nullptr && epi0.accumulate == 0;
    constexpr bool MASKED = (A_MN == 0 && B_MN == 1);  // only the data-gradient form carries the ReLU-mask epilogue
    const bool f16_only = epi0.out16 != nullptr && epi0.out32 == nullptr && epi0.addend == nullptr &&
                          !(MASKED && epi0.relu_mask != nullptr);
    const float alpha = epi0.alpha != nullptr ? __ldcg(epi0.alpha) : 1.f;  // written by a recent kernel (gradient scale): coherent load
    const unsigned long long drop_seed = (DROP && epi0.drop.seed != nullptr) ? __ldcg(epi0.drop.seed) : 0ull;
    const float mscale = epi0.mask_scale != 0.f ? epi0.mask_scale : 1.f;
    const int sub_r = lane >> 2, c8 = lane & 3;  // coalesced phase: 4 lanes x 8 columns per row, 8 rows per pass
    constexpr int NCHUNK = BN / 32;
    // shared-memory slots of the transpose tile (float4 units); (row & 7) == sub_r for every row this lane reads
    const int wr_base = lane * 8, wr_sw = lane & 7;
    const int rd0 = sub_r * 8 + ((2 * c8) ^ sub_r), rd1 = sub_r * 8 + ((2 * c8 + 1) ^ sub_r);
    uint32_t lt = 0;
    for (int t = first_tile; t < num_tiles; t += tile_stride, ++lt) {
      const int bt = t / tiles_per_batch, tmn = (t % tiles_per_batch) % tiles_mn;
      const int m0 = ((tmn / tiles_n) * CL + rank) * BM, n0 = (tmn % tiles_n) * BN;
      GemmEpi epi = epi0;  // this problem's operands
      if (bt > 0) {
        if (epi.bias) epi.bias += bt * epi.s_bias;
        if (epi.addend) epi.addend += bt * epi.s_add;
        if (epi.out32) epi.out32 += bt * epi.s_out32;
        if (epi.out16) epi.out16 += bt * epi.s_out16;
      }
      const uint32_t buf = lt % NBUF;
      const int row0 = m0 + q * 32 + sub_r;  // this lane's first row; the others are +8, +16, +24
      // per-row element offsets, hoisted out of the chunk loop
      size_t off32[4], off16[4], offad[4], offmk[4];
      bool row_ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + i * 8;
        row_ok[i] = r < M;
        off32[i] = (size_t)r * epi.ld32;
        off16[i] = (size_t)r * epi.ld16;
        offmk[i] = (size_t)r * epi.ld_mask;
        offad[i] = (size_t)(epi.add_period > 0 ? r % epi.add_period : r) * epi.ld_add;
      }
      mbar_wait(bar_acc_full(buf), (lt / NBUF) & 1);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      int bkb0, bkb1;
      kb_range(t % tiles_per_batch, bkb0, bkb1);
      const int nidx = tmn % tiles_n;
      const bool has_bias = bkb0 + ((nidx - bkb0 % tiles_n) + tiles_n) % tiles_n < bkb1;  // this tile summed >= 1 k-block
      if (BIAS && epi.colsum_a != nullptr && has_bias && half == 0) {  // partial row sums of A: the bias gradient
        uint32_t rb[32];
        tc_ld32(tmem_base + BIAS_COL + ((uint32_t)(q * 32) << 16), rb);
        tc_wait_ld();
        const int r = m0 + q * 32 + lane;
        if (r < M) grad_red_f32(epi.colsum_a + r, __uint_as_float(rb[0]) * alpha, epi.multimem);
      }
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        if (epi0.tma_red) {
          // x += A W^T + b (in-place residual) or a split-K partial tile: the [32 rows x 32 cols] f32 block of this warp
          // is staged row-per-thread (the TMEM layout: no transpose) in its 128B-swizzled shared-memory tile and ONE TMA
          // reduce-add operation adds it into the output; rows >= M / columns >= N are clipped by the tensor map.
          // Per-lane red.global.add retires ~16 B per clock per SM, the bulk form ~26 B (the SM's L2 write path):
          // single-tile launches (N = d residual projections) spend most of their time in this epilogue.
          const int cb = n0 + c * 32;
          float4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (epi.bias != nullptr && cb + 4 * j < N) bb[j] = __ldg(reinterpret_cast<const float4*>(epi.bias + cb + 4 * j));
          }
          uint32_t acc[32];
          tc_ld32(t_acc + c * 32, acc);
          tc_wait_ld();
          if (c + 2 >= NCHUNK) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(buf));
          }
          if (lane == 0) tma_store_wait_read();   // this warp's previous block has left the staging tile
          __syncwarp();
          const uint32_t tile = base + L::XPOSE_OFF + ew * STAGE_TILE_BYTES + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0 = fmaf(__uint_as_float(acc[4 * j]), alpha, bb[j].x), v1 = fmaf(__uint_as_float(acc[4 * j + 1]), alpha, bb[j].y);
            float v2 = fmaf(__uint_as_float(acc[4 * j + 2]), alpha, bb[j].z), v3 = fmaf(__uint_as_float(acc[4 * j + 3]), alpha, bb[j].w);
            if (epi.act == MTN_ACT_RELU) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4)),
                         "f"(v0), "f"(v1), "f"(v2), "f"(v3)
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_3d(&tmC, base + L::XPOSE_OFF + ew * STAGE_TILE_BYTES, cb, m0 + q * 32, bt);
            tma_store_commit();
          }
          continue;
        }
        if (f16_only) {
          // f16-only outputs (Q/K/V projections, FFN hidden: most of the FLOPs): bias + activation in the
          // row-per-thread layout, round to f16 BEFORE the transpose -- half the shared-memory traffic of the
          // f32 path, and shared-memory bandwidth is what the epilogue and the mainloop compete for.
          const int cb = n0 + c * 32;
          float4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // warp-uniform addresses: broadcast loads, issued before the TMEM load
            bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (epi.bias != nullptr && cb + 4 * j < N) bb[j] = __ldg(reinterpret_cast<const float4*>(epi.bias + cb + 4 * j));
          }
          uint32_t acc[32];
          tc_ld32(t_acc + c * 32, acc);
          tc_wait_ld();
          if (c + 2 >= NCHUNK) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(buf));
          }
          uint32_t pk[16];
          uint32_t keep8 = 0xffu;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v0 = fmaf(__uint_as_float(acc[4 * j]), alpha, bb[j].x), v1 = fmaf(__uint_as_float(acc[4 * j + 1]), alpha, bb[j].y);
            float v2 = fmaf(__uint_as_float(acc[4 * j + 2]), alpha, bb[j].z), v3 = fmaf(__uint_as_float(acc[4 * j + 3]), alpha, bb[j].w);
            if (epi.act == MTN_ACT_RELU) {
              v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
            }
            if (DROP && epi.drop.seed != nullptr) {  // this thread's row, columns cb + 4j .. +3: half of an 8-element group
              if ((j & 1) == 0) keep8 = drop_keep8(epi.drop, drop_seed, ((unsigned long long)(m0 + q * 32 + lane) * N + cb + 4 * j) >> 3);
              const uint32_t kb = keep8 >> (4 * (j & 1));
              v0 = (kb & 1u) ? v0 * epi.drop.inv_keep : 0.f; v1 = (kb & 2u) ? v1 * epi.drop.inv_keep : 0.f;
              v2 = (kb & 4u) ? v2 * epi.drop.inv_keep : 0.f; v3 = (kb & 8u) ? v3 * epi.drop.inv_keep : 0.f;
            }
            pk[2 * j] = pack_f16x2_sat(v0, v1);
            pk[2 * j + 1] = pack_f16x2_sat(v2, v3);
          }
          // [32 rows x 64 B] tile, 16-B slots XOR-swizzled by (row >> 1) & 3: conflict-free both ways
          uint4* hp = reinterpret_cast<uint4*>(xp);
          const int wsw = (lane >> 1) & 3;
#pragma unroll
          for (int sI = 0; sI < 4; ++sI)
            hp[lane * 4 + (sI ^ wsw)] = make_uint4(pk[4 * sI], pk[4 * sI + 1], pk[4 * sI + 2], pk[4 * sI + 3]);
          __syncwarp();
          uint4 hv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + sub_r;
            hv[i] = hp[rl * 4 + (c8 ^ ((rl >> 1) & 3))];
          }
          __syncwarp();
          const int colh = cb + c8 * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (row_ok[i] && colh < N) *reinterpret_cast<uint4*>(epi.out16 + off16[i] + colh) = hv[i];
          continue;
        }
        const int col = n0 + c * 32 + c8 * 8;  // this lane's 8 columns in the coalesced phase
        const bool col_ok = col < N;           // N % 8 == 0: all 8 or none
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (epi.bias != nullptr && col_ok) {
          b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + col));
          b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + col + 4));
        }
        float4 res[8];
        uint4 mk[4];
        if (MASKED && epi.relu_mask != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mk[i] = make_uint4(0u, 0u, 0u, 0u);
            if (row_ok[i] && col_ok) mk[i] = *reinterpret_cast<const uint4*>(epi.relu_mask + offmk[i] + col);
          }
        }
        if (has_add && !red_add) {  // residual / positional operand: requested before touching TMEM
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            res[2 * i] = res[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok[i] && col_ok) {
              const float4* ap = reinterpret_cast<const float4*>(epi.addend + offad[i] + col);
              res[2 * i] = ap[0];
              res[2 * i + 1] = ap[1];
            }
          }
        }
        uint32_t acc[32];
        tc_ld32(t_acc + c * 32, acc);
        tc_wait_ld();
        if (c + 2 >= NCHUNK) {  // this warp's last read of the accumulator: release its share of the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          xp[wr_base + (j ^ wr_sw)] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                                  __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
        __syncwarp();
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // all shared-memory reads first (independent of the global stores below)
          v[2 * i] = xp[i * 64 + rd0];
          v[2 * i + 1] = xp[i * 64 + rd1];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 x0 = v[2 * i], x1 = v[2 * i + 1];
          x0.x = fmaf(x0.x, alpha, b0.x); x0.y = fmaf(x0.y, alpha, b0.y); x0.z = fmaf(x0.z, alpha, b0.z); x0.w = fmaf(x0.w, alpha, b0.w);
          x1.x = fmaf(x1.x, alpha, b1.x); x1.y = fmaf(x1.y, alpha, b1.y); x1.z = fmaf(x1.z, alpha, b1.z); x1.w = fmaf(x1.w, alpha, b1.w);
          if (epi.act == MTN_ACT_RELU) {
            x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
            x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
          }
          if (MASKED && epi.relu_mask != nullptr) {  // ReLU backward: keep the gradient where the forward activation was > 0
            const __half2* hm = reinterpret_cast<const __half2*>(&mk[i]);
            const float2 m0_ = __half22float2(hm[0]), m1_ = __half22float2(hm[1]), m2_ = __half22float2(hm[2]),
                         m3_ = __half22float2(hm[3]);
            x0.x = m0_.x > 0.f ? x0.x * mscale : 0.f; x0.y = m0_.y > 0.f ? x0.y * mscale : 0.f;
            x0.z = m1_.x > 0.f ? x0.z * mscale : 0.f; x0.w = m1_.y > 0.f ? x0.w * mscale : 0.f;
            x1.x = m2_.x > 0.f ? x1.x * mscale : 0.f; x1.y = m2_.y > 0.f ? x1.y * mscale : 0.f;
            x1.z = m3_.x > 0.f ? x1.z * mscale : 0.f; x1.w = m3_.y > 0.f ? x1.w * mscale : 0.f;
          }
          uint32_t keep8 = 0xffu;
          if (DROP && epi.drop.seed != nullptr)
            keep8 = drop_keep8(epi.drop, drop_seed, ((unsigned long long)(row0 + i * 8) * N + col) >> 3);
          auto apply_drop = [&]() {
            const float ik = epi.drop.inv_keep;
            x0.x = (keep8 & 1u) ? x0.x * ik : 0.f; x0.y = (keep8 & 2u) ? x0.y * ik : 0.f;
            x0.z = (keep8 & 4u) ? x0.z * ik : 0.f; x0.w = (keep8 & 8u) ? x0.w * ik : 0.f;
            x1.x = (keep8 & 16u) ? x1.x * ik : 0.f; x1.y = (keep8 & 32u) ? x1.y * ik : 0.f;
            x1.z = (keep8 & 64u) ? x1.z * ik : 0.f; x1.w = (keep8 & 128u) ? x1.w * ik : 0.f;
          };
          if (DROP && epi.drop.seed != nullptr && !epi.drop_after_add) apply_drop();
          uint4 pre16 = make_uint4(0u, 0u, 0u, 0u);
          if (epi.out16_pre_add)
            pre16 = make_uint4(pack_f16x2_sat(x0.x, x0.y), pack_f16x2_sat(x0.z, x0.w), pack_f16x2_sat(x1.x, x1.y),
                               pack_f16x2_sat(x1.z, x1.w));
          if (has_add && !red_add) {
            x0.x += res[2 * i].x; x0.y += res[2 * i].y; x0.z += res[2 * i].z; x0.w += res[2 * i].w;
            x1.x += res[2 * i + 1].x; x1.y += res[2 * i + 1].y; x1.z += res[2 * i + 1].z; x1.w += res[2 * i + 1].w;
          }
          if (DROP && epi.drop.seed != nullptr && epi.drop_after_add) apply_drop();
          if (row_ok[i] && col_ok) {
            if (red_add) {
              float* o = epi.out32 + off32[i] + col;
              const int mm = epi.accumulate ? epi.multimem : 0;
              grad_red_v4(o, x0.x, x0.y, x0.z, x0.w, mm);
              grad_red_v4(o + 4, x1.x, x1.y, x1.z, x1.w, mm);
            } else if (epi.out32 != nullptr) {
              float4* o = reinterpret_cast<float4*>(epi.out32 + off32[i] + col);
              o[0] = x0;
              o[1] = x1;
            }
            if (epi.out16 != nullptr)
              *reinterpret_cast<uint4*>(epi.out16 + off16[i] + col) =
                  epi.out16_pre_add ? pre16
                                    : make_uint4(pack_f16x2_sat(x0.x, x0.y), pack_f16x2_sat(x0.z, x0.w),
                                                 pack_f16x2_sat(x1.x, x1.y), pack_f16x2_sat(x1.z, x1.w));
          }
        }
      }
    }
    if (epi0.tma_red && lane == 0) tma_store_wait_read();   // (the kernel boundary completes the writes)
    tc_fence_before();
  }
  // no CTA may leave while a peer can still multicast into its shared memory / arrive on its barriers
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int g_num_sms = 0;

static int ensure_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    MTN_CHECK_CUDA(cudaGetDevice(&dev));
    MTN_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    g_num_sms = n;
  }
  return MTN_OK;
}

// Tensor map of one operand.  K-major: dims {K, MN, batch}, box {64, box_mn}.  MN-major: dims {MN, K, batch},
// box {64 mn-columns, 64 k-rows} (the kernel issues box_mn / 64 of them per k-block).
static int make_operand_map(CUtensorMap* tm, const void* p, int mn_major, int MN, int K, int ld, int batch,
                            long long stride, int box_mn) {
  if (mn_major)
    return make_tmap_3d_f16(tm, p, MN, K, batch, ld, batch > 1 ? (uint64_t)stride : (uint64_t)K * ld, 64, BK, TM_SWZ_128);
  return make_tmap_3d_f16(tm, p, K, MN, batch, ld, batch > 1 ? (uint64_t)stride : (uint64_t)MN * ld, BK, box_mn, TM_SWZ_128);
}

template <int BN, int STAGES, int CL, int A_MN, int B_MN, int DROP = 0, int BIAS = 0>
static int launch_gemm(const MtnGemmArgs& a, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  static int max_clusters_dev = 0;  // co-resident clusters on the whole device (1 CTA per SM)
  int& max_clusters = max_clusters_dev;
  if (max_clusters == 0) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<BN, STAGES, CL, A_MN, B_MN, DROP, BIAS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    if (CL == 1) {
      max_clusters = g_num_sms;
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(g_num_sms / CL * CL);
      cfg.blockDim = dim3(GEMM_THREADS);
      cfg.dynamicSmemBytes = L::TOTAL;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = 0;
      MTN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_f16_tc_kernel<BN, STAGES, CL, A_MN, B_MN, DROP, BIAS>, &cfg));
      MTN_REQUIRE(n > 0, MTN_E_CUDA, "gemm: no cluster of %d CTAs fits on this device", CL);
      max_clusters = n;
    }
  }
  // a stream of an SM partition (green context, mtn_b200/parallel.py): the grid follows the partition's SM count
  const int part_sms = stream_sm_count(st);
  int max_clusters_here = max_clusters_dev;
  if (part_sms < g_num_sms) max_clusters_here = (part_sms / CL) < max_clusters_dev ? (part_sms / CL) : max_clusters_dev;
  if (max_clusters_here < 1) max_clusters_here = 1;
#define max_clusters max_clusters_here
  const int batch = a.batch > 1 ? a.batch : 1;
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, a.A, A_MN, a.M, a.K, a.lda, batch, a.stride_A, BM);
  if (rc) return rc;
  rc = make_operand_map(&tmB, a.B, B_MN, a.N, a.K, a.ldb, batch, a.stride_B, BN / CL);
  if (rc) return rc;
  GemmEpi epi{a.bias, a.act, a.addend, a.ld_add, a.add_period, a.out_f32, a.ld32,
              reinterpret_cast<__half*>(a.out_f16), a.ld16, a.alpha, reinterpret_cast<const __half*>(a.relu_mask),
              a.ld_mask, a.accumulate, a.out16_pre_add, a.multimem, a.colsum_a, a.mask_scale,
              DropCfg{reinterpret_cast<const unsigned long long*>(a.drop_seed), a.drop_site, a.drop_thresh,
                      a.drop_thresh ? 1.f / (1.f - a.drop_thresh / 65536.f) : 1.f},
              a.drop_after_add, a.stride_bias, a.stride_add, a.stride_out_f32, a.stride_out_f16, 0};
  // TMA reduce-add epilogue: the result is ADDED into an f32 output and nothing else is written
  CUtensorMap tmC = tmA;
  {
    static int tma_red_on = -1;   // MTN_B200_GEMM_TMA_RED=0: per-lane red.global.add (round-1 form)
    if (tma_red_on < 0) {
      const char* e = getenv("MTN_B200_GEMM_TMA_RED");
      tma_red_on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const bool inplace = a.addend != nullptr && a.addend == a.out_f32 && a.add_period == 0 && a.ld_add == a.ld32 &&
                         a.stride_add == a.stride_out_f32;
    const bool eligible = tma_red_on && a.out_f32 != nullptr && a.out_f16 == nullptr && a.relu_mask == nullptr &&
                          a.drop_seed == nullptr && !a.multimem && (a.accumulate ? a.addend == nullptr : inplace) &&
                          a.ld32 % 4 == 0 && aligned16(a.out_f32) && (batch == 1 || a.stride_out_f32 % 4 == 0);
    if (eligible) {
      rc = make_tmap_3d_f32(&tmC, a.out_f32, a.N, a.M, batch, a.ld32,
                            batch > 1 ? (uint64_t)a.stride_out_f32 : (uint64_t)a.M * a.ld32, 32, 32, TM_SWZ_128);
      if (rc) return rc;
      epi.tma_red = 1;
    }
  }
  const int tiles_n = (a.N + BN - 1) / BN, tiles_m = (a.M + BM - 1) / BM;
  const int tiles_mn = tiles_n * ((tiles_m + CL - 1) / CL);
  // split-K (accumulating outputs only): cut the contraction so that the launch is ONE wave of CTAs -- every slice
  // pays a full-tile red.add epilogue, so fewer, longer slices beat many short ones
  const int nkb = (a.K + BK - 1) / BK;
  int kb_per_split = nkb;
  if (a.accumulate) {
    long want = (long)max_clusters / ((long)tiles_mn * batch);  // slices per output tile
    if (want < 1) want = 1;
    long per = (nkb + want - 1) / want;
    if (per < 2) per = nkb < 2 ? nkb : 2;
    kb_per_split = (int)per;
  }
  const int ksplit = (nkb + kb_per_split - 1) / kb_per_split;
  const int tiles_per_batch = tiles_mn * ksplit;
  const int num_super = tiles_per_batch * batch;
  const int clusters = num_super < max_clusters ? num_super : max_clusters;
  MTN_CHECK_CUDA(launch_kernel_cluster(gemm_f16_tc_kernel<BN, STAGES, CL, A_MN, B_MN, DROP, BIAS>, dim3(clusters * CL),
                                       dim3(GEMM_THREADS), L::TOTAL, st, (unsigned)CL, tmA, tmB, tmC, epi, a.M, a.N, a.K,
                                       tiles_n, tiles_per_batch, num_super, tiles_mn, kb_per_split));
#undef max_clusters
  return MTN_OK;
}

static int validate_gemm(const MtnGemmArgs* a) {
  MTN_REQUIRE(a != nullptr && a->A != nullptr && a->B != nullptr, MTN_E_ARG, "gemm: NULL operand");
  MTN_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, MTN_E_SHAPE, "gemm: M=%d N=%d K=%d", a->M, a->N, a->K);
  MTN_REQUIRE(a->N % 8 == 0, MTN_E_SHAPE, "gemm: N=%d must be a multiple of 8", a->N);
  // 16-byte aligned rows of every operand: the contiguous extent must be a multiple of 8 elements
  MTN_REQUIRE((a->a_mn ? a->M : a->K) % 8 == 0, MTN_E_SHAPE, "gemm: contiguous extent of A (%s=%d) must be a multiple of 8",
              a->a_mn ? "M" : "K", a->a_mn ? a->M : a->K);
  MTN_REQUIRE(a->b_mn || a->K % 8 == 0, MTN_E_SHAPE, "gemm: K=%d must be a multiple of 8", a->K);
  MTN_REQUIRE(a->lda >= (a->a_mn ? a->M : a->K) && a->ldb >= (a->b_mn ? a->N : a->K) && a->lda % 8 == 0 &&
                  a->ldb % 8 == 0,
              MTN_E_ALIGN, "gemm: lda=%d ldb=%d must cover the contiguous extent and be multiples of 8", a->lda, a->ldb);
  MTN_REQUIRE(aligned16(a->A) && aligned16(a->B), MTN_E_ALIGN, "gemm: A/B not 16-byte aligned");
  MTN_REQUIRE(a->out_f32 != nullptr || a->out_f16 != nullptr, MTN_E_ARG, "gemm: no output");
  MTN_REQUIRE(a->act == MTN_ACT_NONE || a->act == MTN_ACT_RELU, MTN_E_ARG, "gemm: act=%d", a->act);
  if (a->out_f32)
    MTN_REQUIRE(aligned16(a->out_f32) && a->ld32 % 4 == 0 && a->ld32 >= a->N, MTN_E_ALIGN,
                "gemm: out_f32 alignment / ld32=%d", a->ld32);
  if (a->out_f16)
    MTN_REQUIRE(aligned16(a->out_f16) && a->ld16 % 8 == 0 && a->ld16 >= a->N, MTN_E_ALIGN,
                "gemm: out_f16 alignment / ld16=%d", a->ld16);
  if (a->addend)
    MTN_REQUIRE(aligned16(a->addend) && a->ld_add % 4 == 0 && a->ld_add >= a->N && a->add_period >= 0,
                MTN_E_ALIGN, "gemm: addend alignment / ld_add=%d", a->ld_add);
  if (a->bias) MTN_REQUIRE(aligned16(a->bias), MTN_E_ALIGN, "gemm: bias not 16-byte aligned");
  if (a->relu_mask)
    MTN_REQUIRE(aligned16(a->relu_mask) && a->ld_mask % 8 == 0 && a->ld_mask >= a->N && !a->a_mn && a->b_mn, MTN_E_ALIGN,
                "gemm: relu_mask (data-gradient form only) alignment / ld_mask=%d", a->ld_mask);
  if (a->drop_seed != nullptr)
    MTN_REQUIRE(a->drop_thresh < 65536u && a->batch <= 1 && !a->accumulate && !a->a_mn && !a->b_mn, MTN_E_ARG,
                "gemm: dropout needs thresh < 65536, the forward (K-major) form, no batch, no accumulate");
  MTN_REQUIRE(!a->multimem || a->accumulate, MTN_E_ARG, "gemm: multimem belongs to the accumulating form");
  if (a->colsum_a != nullptr)
    MTN_REQUIRE(a->a_mn && a->b_mn && a->accumulate && a->M % 8 == 0, MTN_E_ARG,
                "gemm: colsum_a (bias gradient) belongs to the accumulating weight-gradient form");
  if (a->accumulate)
    MTN_REQUIRE(a->out_f32 != nullptr && a->out_f16 == nullptr && a->bias == nullptr && a->addend == nullptr &&
                    a->act == MTN_ACT_NONE && a->relu_mask == nullptr,
                MTN_E_ARG, "gemm: accumulate takes only alpha and out_f32");
  if (a->batch > 1) {
    MTN_REQUIRE(a->stride_A % 8 == 0 && a->stride_B % 8 == 0 && a->stride_A > 0 && a->stride_B > 0 &&
                    a->stride_bias % 4 == 0 && a->stride_add % 4 == 0 && a->stride_out_f32 % 4 == 0 &&
                    a->stride_out_f16 % 8 == 0,
                MTN_E_ALIGN, "gemm: batch strides must keep every problem 16-byte aligned");
    MTN_REQUIRE(a->batch <= 65535, MTN_E_SHAPE, "gemm: batch=%d", a->batch);
    MTN_REQUIRE(a->relu_mask == nullptr, MTN_E_ARG, "gemm: relu_mask is not batched");
  }
  return MTN_OK;
}

static int run_gemm(const MtnGemmArgs* a, void* stream) {
  int rc = validate_gemm(a);
  if (rc) return rc;
  rc = ensure_sms();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 128x256 tiles halve the operand bytes per FLOP; use them when they still fill the machine.
  // Wide GEMMs with many waves additionally run as clusters of 2 CTAs along M that share each W tile
  // through TMA multicast (measured +3 %; on the small single-wave GEMMs clusters only constrain
  // placement next to the other streams' kernels, so they stay unclustered).  MTN_B200_CLUSTER=1 disables.
  static int cl = -1;
  if (cl < 0) {
    const char* e = getenv("MTN_B200_CLUSTER");
    cl = e ? atoi(e) : 2;
  }
  const int tiles_m = (a->M + 127) / 128;
  const long tiles256 = (long)((a->N + 255) / 256) * tiles_m * (a->batch > 1 ? a->batch : 1);
  static int big_pct = -1;  // 128x256 tiles once they fill this percentage of the SMs
  if (big_pct < 0) {
    const char* e = getenv("MTN_B200_BIG_PCT");
    big_pct = e ? atoi(e) : 45;
  }
  // split-K launches fill the machine through their slices; small outputs take 128x128 tiles (half the red.add
  // bytes per slice, twice the slices' length), large ones the wide tile
  const bool big = a->N >= 256 && (a->accumulate ? tiles256 * 2 >= g_num_sms : tiles256 * 100 >= (long)big_pct * g_num_sms);
  const int form = (a->a_mn ? 2 : 0) | (a->b_mn ? 1 : 0);
  switch (form) {
    case 0:
      if (a->drop_seed != nullptr) return big ? launch_gemm<256, 4, 1, 0, 0, 1>(*a, st) : launch_gemm<128, 6, 1, 0, 0, 1>(*a, st);
      if (cl >= 2 && big && !a->accumulate && tiles256 >= 4L * g_num_sms) return launch_gemm<256, 4, 2, 0, 0>(*a, st);
      return big ? launch_gemm<256, 4, 1, 0, 0>(*a, st) : launch_gemm<128, 6, 1, 0, 0>(*a, st);
    case 1:
      return big ? launch_gemm<256, 4, 1, 0, 1>(*a, st) : launch_gemm<128, 6, 1, 0, 1>(*a, st);
    case 3:
      if (a->colsum_a != nullptr) {
        if (!big && a->batch <= 1) return launch_gemm<128, 6, 1, 1, 1, 0, 1>(*a, st);   // bias gradient on the tensor core
        // wide tiles have no TMEM columns to spare: the row sums of A take a separate pass over it
        rc = mtn_cast_colsum(a->A, 1, a->lda, nullptr, 0, nullptr, 0, a->K, a->M, nullptr, a->alpha, a->colsum_a, nullptr, 0, 0,
                             a->multimem, stream);
        if (rc) return rc;
      }
      return big ? launch_gemm<256, 4, 1, 1, 1>(*a, st) : launch_gemm<128, 6, 1, 1, 1>(*a, st);
    default:
      return set_error(MTN_E_ARG, "gemm: the (A MN-major, B K-major) form is not instantiated");
  }
}

static MtnGemmArgs from_linear(const MtnLinearArgs& a) {
  MtnGemmArgs g = {};
  g.A = a.A; g.lda = a.lda; g.B = a.W; g.ldb = a.ldw;
  g.M = a.M; g.N = a.N; g.K = a.K;
  g.bias = a.bias; g.act = a.act;
  g.addend = a.addend; g.ld_add = a.ld_add; g.add_period = a.add_period;
  g.out_f32 = a.out_f32; g.ld32 = a.ld32; g.out_f16 = a.out_f16; g.ld16 = a.ld16;
  g.out16_pre_add = a.out16_pre_add;
  g.drop_seed = a.drop_seed; g.drop_site = a.drop_site; g.drop_thresh = a.drop_thresh; g.drop_after_add = a.drop_after_add;
  g.batch = a.batch; g.stride_A = a.stride_A; g.stride_B = a.stride_W; g.stride_bias = a.stride_bias;
  g.stride_add = a.stride_add; g.stride_out_f32 = a.stride_out_f32; g.stride_out_f16 = a.stride_out_f16;
  return g;
}

// ----------------------------------------------------------------------------
// self-check kernel (tests only): one thread per output element, same arithmetic
// contract (f16 operands, f32 accumulate, same epilogue order).
// ----------------------------------------------------------------------------
__global__ void gemm_f16_check_kernel(const __half* A, int lda, int a_mn, const __half* B, int ldb, int b_mn,
                                      GemmEpi epi, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float av = __half2float(a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]);
    const float bv = __half2float(b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]);
    acc = fmaf(av, bv, acc);
  }
  if (epi.colsum_a != nullptr && n == 0) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += __half2float(a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]);
    epi.colsum_a[m] += s * (epi.alpha ? epi.alpha[0] : 1.f);
  }
  if (epi.alpha) acc *= epi.alpha[0];
  if (epi.bias) acc += epi.bias[n];
  if (epi.act == MTN_ACT_RELU) acc = fmaxf(acc, 0.f);
  if (epi.relu_mask) acc = (__half2float(epi.relu_mask[(size_t)m * epi.ld_mask + n]) > 0.f) ? acc * (epi.mask_scale != 0.f ? epi.mask_scale : 1.f) : 0.f;
  const float pre = acc;
  bool keep = true;
  if (epi.drop.seed) keep = (drop_keep8(epi.drop, epi.drop.seed[0], ((unsigned long long)m * N + n) >> 3) >> (n & 7)) & 1u;
  if (epi.drop.seed && !epi.drop_after_add) acc = keep ? acc * epi.drop.inv_keep : 0.f;
  if (epi.addend && !epi.accumulate)
    acc += epi.addend[(size_t)(epi.add_period > 0 ? m % epi.add_period : m) * epi.ld_add + n];
  if (epi.drop.seed && epi.drop_after_add) acc = keep ? acc * epi.drop.inv_keep : 0.f;
  if (epi.out32) {
    if (epi.accumulate) epi.out32[(size_t)m * epi.ld32 + n] += acc;
    else epi.out32[(size_t)m * epi.ld32 + n] = acc;
  }
  if (epi.out16) {
    const uint32_t p = pack_f16x2_sat(epi.out16_pre_add ? pre : acc, 0.f);
    epi.out16[(size_t)m * epi.ld16 + n] = __ushort_as_half((unsigned short)(p & 0xffff));
  }
}

static int run_check_gemm(const MtnGemmArgs* a, void* stream) {
  int rc = validate_gemm(a);
  if (rc) return rc;
  GemmEpi epi{a->bias, a->act, a->addend, a->ld_add, a->add_period, a->out_f32, a->ld32,
              reinterpret_cast<__half*>(a->out_f16), a->ld16, a->alpha, reinterpret_cast<const __half*>(a->relu_mask),
              a->ld_mask, a->accumulate, a->out16_pre_add, 0, a->colsum_a, a->mask_scale,
              DropCfg{reinterpret_cast<const unsigned long long*>(a->drop_seed), a->drop_site, a->drop_thresh,
                      a->drop_thresh ? 1.f / (1.f - a->drop_thresh / 65536.f) : 1.f},
              a->drop_after_add, 0, 0, 0, 0};
  MTN_REQUIRE(a->batch <= 1, MTN_E_ARG, "check_gemm: the check kernel is not batched");
  dim3 grid((a->N + 127) / 128, a->M);
  gemm_f16_check_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(a->A), a->lda, a->a_mn, reinterpret_cast<const __half*>(a->B), a->ldb, a->b_mn,
      epi, a->M, a->N, a->K);
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}

}  // namespace mtn

extern "C" int mtn_gemm_f16(const MtnGemmArgs* a, void* stream) { return mtn::run_gemm(a, stream); }
extern "C" int mtn_check_gemm_f16(const MtnGemmArgs* a, void* stream) { return mtn::run_check_gemm(a, stream); }

extern "C" int mtn_linear_fwd(const MtnLinearArgs* a, void* stream) {
  MTN_REQUIRE(a != nullptr, MTN_E_ARG, "linear: NULL args");
  const MtnGemmArgs g = mtn::from_linear(*a);
  return mtn::run_gemm(&g, stream);
}

extern "C" int mtn_check_linear_fwd(const MtnLinearArgs* a, void* stream) {
  MTN_REQUIRE(a != nullptr, MTN_E_ARG, "linear: NULL args");
  const MtnGemmArgs g = mtn::from_linear(*a);
  return mtn::run_check_gemm(&g, stream);
}

// dX[M, K] = alpha * (dY[M, N] W[N, K])  (optionally masked by the forward ReLU output): the contraction runs
// over the layer's OUTPUT features, W is read in its forward [out, in] layout as an MN-major operand.
extern "C" int mtn_linear_dgrad(const MtnLinearDgradArgs* a, void* stream) {
  MTN_REQUIRE(a != nullptr, MTN_E_ARG, "linear_dgrad: NULL args");
  MtnGemmArgs g = {};
  g.A = a->dY; g.lda = a->lddy; g.a_mn = 0;
  g.B = a->W; g.ldb = a->ldw; g.b_mn = 1;
  g.M = a->M; g.N = a->K; g.K = a->N;
  g.alpha = a->alpha;
  g.relu_mask = a->relu_mask; g.ld_mask = a->ld_mask; g.mask_scale = a->mask_scale;
  g.addend = a->addend; g.ld_add = a->ld_add;
  g.out_f32 = a->dX_f32; g.ld32 = a->ld32; g.out_f16 = a->dX_f16; g.ld16 = a->ld16;
  g.batch = a->batch; g.stride_A = a->stride_dY; g.stride_B = a->stride_W;
  g.stride_add = a->stride_add; g.stride_out_f32 = a->stride_dX_f32; g.stride_out_f16 = a->stride_dX_f16;
  return mtn::run_gemm(&g, stream);
}

// dW[N, K] += alpha * (dY[M, N]^T X[M, K]): both operands MN-major (forward layouts), contraction over the
// M rows, split-K with red.global.add accumulation into the f32 gradient.
extern "C" int mtn_linear_wgrad(const MtnLinearWgradArgs* a, void* stream) {
  MTN_REQUIRE(a != nullptr, MTN_E_ARG, "linear_wgrad: NULL args");
  MtnGemmArgs g = {};
  g.A = a->dY; g.lda = a->lddy; g.a_mn = 1;
  g.B = a->X; g.ldb = a->ldx; g.b_mn = 1;
  g.M = a->N; g.N = a->K; g.K = a->M;
  g.alpha = a->alpha;
  g.accumulate = 1;
  g.colsum_a = a->dbias;
  g.multimem = a->multimem;
  g.out_f32 = a->dW; g.ld32 = a->lddw;
  g.batch = a->batch; g.stride_A = a->stride_dY; g.stride_B = a->stride_X; g.stride_out_f32 = a->stride_dW;
  return mtn::run_gemm(&g, stream);
}
