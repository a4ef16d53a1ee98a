// One KV-cached decoding step (reference call form data_utils.py:202-210, one new target position per
// dialogue) as ONE kernel in which every DIALOGUE GROUP is owned by one thread-block cluster.
//
// Why: a cached step is a chain of ~135 dependent few-row launches (csrc/decode_rows.cu), each a kernel boundary (~3 us)
// plus one exposed memory round trip; a grid-wide persistent kernel (decode_prog_kernel) pays a grid barrier (three L2
// round trips) per stage instead and is no faster.  But the dependency chain of a decoding step is PER DIALOGUE: nothing
// in DecoderLayer.forward (mtn.py:181-218) mixes dialogues.  So a cluster of 8 CTAs takes G = ceil(B / #clusters)
// dialogues through the whole N-layer step; CTA `rank` owns head `rank` (d_k = 64) and the matching 64 output columns of
// every projection (256 of the feed-forward hidden layer), and nothing in the step synchronises wider than the cluster:
//
//   per attention sublayer (mtn.py:125-127 around :248-267 around :221-231)
//     LayerNorm of the G residual rows (every CTA holds the full rows, f32, in shared memory)
//     q (self: q, k, v) of head `rank`:  mma.sync over the CTA's 64 weight rows (8 per warp)
//     attention of head `rank`: the static memory's K / V (or the self-attention cache rows 0..t-1 plus the new row,
//       which is also appended to the cache), online softmax in the log2 domain with the reference's FINITE -1e9
//     the head's output -> every CTA of the cluster (st.async into distributed shared memory)
//     output projection of the CTA's 64 columns + residual -> every CTA's copy of the rows (st.async)
//   feed-forward sublayer (mtn.py:279-280): LayerNorm, w_1 + ReLU (256 hidden columns per CTA) -> every CTA,
//     w_2 (64 columns, K = 2048) + residual -> every CTA
//   last: Decoder.norm (mtn.py:164) and, for greedy decoding, the generator's projection + arg-max (mtn.py:68-69,
//     data_utils.py:183) over vocabulary slices, the per-CTA winners gathered in CTA 0.
//
// EXCHANGES.  A value another CTA needs is stored straight into that CTA's shared memory with st.async, which completes
// its bytes on an mbarrier of the RECEIVER; the receiver arms the barrier with the byte count of the exchange (8 CTAs x
// its rows x 64 columns) and waits on it -- no fence, no cluster barrier inside the step (a barrier.cluster.arrive.release
// costs ~1300 cycles here, the st.async hand-off ~400).  Single buffers suffice: a CTA can only send sublayer s + 1 data
// after it has received every CTA's sublayer-s data, i.e. after every reader of the old contents is done.
//
// OPERAND STREAM.  Everything a CTA reads from global memory inside the step -- its weight rows and its head's K / V rows
// -- has an address that does not depend on computed data.  Warps 8..15 are PRODUCERS: producer w walks the chunk sequence
// of compute warp w (weight chunk = the 8 rows x 512 k of one projection job; attention-unit chunk = 32 keys of K and of
// V) and copies it with cp.async (16 bytes per lane: full 128-byte lines per row) into a two-slot ring, the slot's FULL
// mbarrier collecting the copies' completions; it runs ahead ACROSS sublayer boundaries and while the compute warp waits
// for the cluster.  Measured alternatives: the compute warps issuing their own cp.async (the address
// arithmetic lands on the dependency chain), bulk copies (one per row: serialised by the hardware's one-lane-at-a-time
// issue), four smaller slots (per-chunk hand-off costs more than the deeper look-ahead returns), operands prefetched
// straight into registers (64-byte requests: the L1 request rate, not bandwidth, becomes the bound).
//
// Arithmetic = the few-row kernels' (f16 operands, f32 accumulate, f16-rounded q / k / v / P / O / hidden, LayerNorm
// with the summation order of layernorm_rows_kernel); only the summation order of the projections differs (one warp
// accumulates the whole contraction).  d = 512, h = 8, d_ff = 2048, one query row per dialogue (greedy decoding).
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int DC_CS = 8;            // CTAs per cluster = heads
constexpr int DC_CWARPS = 8;        // compute warps; warp DC_CWARPS + w streams the operands of compute warp w
constexpr int DC_CTHREADS = 32 * DC_CWARPS, DC_THREADS = 2 * DC_CTHREADS;
constexpr int DC_G = 8;             // dialogues (rows) per cluster, at most: the m16n8k16 fragments carry rows 0..7
constexpr int DC_D = 512, DC_DFF = 2048, DC_DK = 64;
// ring slots: a WEIGHT chunk is 8 weight rows x 512 k (row pitch 1088 B); an attention-UNIT chunk is 32 keys x 64 dims of K
// (row pitch 144 B) followed, at DC_VOFF, by the same of V.  The pitches make the fragment / lane-per-key reads
// conflict-free.  Two slots per warp: with everything else the step keeps on chip, shared memory has room for no more.
constexpr int DC_NSLOT = 2, DC_SLOT = 9216;
constexpr int DC_WPITCH = 1088, DC_KPITCH = 144, DC_VOFF = 32 * DC_KPITCH;
static_assert(8 * DC_WPITCH <= DC_SLOT && 2 * DC_VOFF <= DC_SLOT, "slot size");
constexpr int DC_MAX_SITES = 64;

// shared memory (bytes).  The f16 A operands (LayerNorm output, gathered attention output, gathered hidden activation)
// have a row stride == 64 (mod 128) so the 16-byte fragment loads of a quarter-warp (two rows x four k-segments) hit
// distinct banks.
constexpr int DC_LDA = DC_D * 2 + 64;
constexpr int DC_LDH = DC_DFF * 2 + 64;
constexpr int DC_OFF_XS = 0;                                      // f32 [8][512]  residual rows
constexpr int DC_OFF_XN = DC_OFF_XS + DC_G * DC_D * 4;            // f16 [8][LDA]  LayerNorm(x)
constexpr int DC_OFF_OB = DC_OFF_XN + DC_G * DC_LDA;              // f16 [8][LDA]  attention output, all heads
constexpr int DC_OFF_HID = DC_OFF_OB + DC_G * DC_LDA;             // f16 [8][LDH]  hidden activation, all columns
constexpr int DC_OFF_PO = DC_OFF_HID;                             // f32 [8 warps][8][64] partial attention outputs (alias: hid is
                                                                  // dead between the feed-forward sublayers)
constexpr int DC_OFF_QS = DC_OFF_HID + DC_G * DC_LDH;             // f32 [3][8][64] q, new k, new v of this head (f16-rounded)
constexpr int DC_OFF_PM = DC_OFF_QS + 3 * DC_G * DC_DK * 4;       // f32 [8 warps][8]      partial row maxima
constexpr int DC_OFF_PL = DC_OFF_PM + DC_CWARPS * DC_G * 4;       // f32 [8 warps][8]      partial row sums
constexpr int DC_OFF_BAR = DC_OFF_PL + DC_CWARPS * DC_G * 4;      // mbarriers
constexpr int DC_OFF_RING = (DC_OFF_BAR + 1024 + 1023) / 1024 * 1024;
constexpr int DC_SMEM = DC_OFF_RING + DC_CWARPS * DC_NSLOT * DC_SLOT;
static_assert(DC_SMEM <= 232448, "shared memory budget");
static_assert(DC_CWARPS * DC_G * DC_DK * 4 <= DC_G * DC_LDH, "partial outputs fit the hidden-activation buffer");
// mbarriers: FULL / EMPTY per (compute warp, ring slot); O / X / H: the cluster-wide exchanges (attention outputs,
// residual rows, hidden activation) complete their bytes on the RECEIVER's barrier
constexpr int DC_BAR_FULL = 0, DC_BAR_EMPTY = DC_CWARPS * DC_NSLOT, DC_BAR_O = 2 * DC_CWARPS * DC_NSLOT, DC_BAR_X = DC_BAR_O + 1,
              DC_BAR_H = DC_BAR_O + 2, DC_BAR_G = DC_BAR_O + 3, DC_BAR_COUNT = DC_BAR_O + 4;
constexpr int DC_OFF_GBUF = DC_OFF_BAR + 512;   // u64 [8 CTAs][8 rows]: every CTA's best (logit, column) per row, gathered in CTA 0
static_assert(DC_BAR_COUNT * 8 <= 512, "barrier area");

struct DcTable {
  MtnDecodeSite s[DC_MAX_SITES];
};

struct DcParams {
  int n_sites, B, G, t;
  int R;               // target rows per dialogue (beam search: R hypotheses share their dialogue's memories; greedy: 1)
  const float* x_in;
  float* out;
  const float* norm_a;
  const float* norm_b;
  float norm_eps;
  float* taps;
  long long* stamps;   // optional [n_sites][8] clock64 stamps of CTA 0, thread 0 (debug: where a sublayer spends its time)
  // optional last stage: arg-max of the generator's logits (mtn.py:62-69 + data_utils.py:183) of the output rows
  const __half* gen_w;       // [gen_V8, 512] f16, rows >= gen_V zero
  const float* gen_b;        // [gen_V8]
  int gen_V, gen_V8;
  long long* tokens;         // token of row r -> tokens[r * tokens_stride]
  long long tokens_stride;
  int kv_prefetch;           // L2 prefetch of the next cross-attention sublayer's K / V (MTN_B200_DECODE_KV_PREFETCH=1; default off)
};

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 dc_lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t dc_lds32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ uint32_t dc_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// remote (or local) shared-memory store that completes its bytes on the mbarrier `bar` of the SAME CTA as `addr`: the
// receiver waits on its own barrier -- no fence, no cluster barrier
__device__ __forceinline__ void dc_st_async_u32(uint32_t addr, uint32_t v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(bar) : "memory");
}
__device__ __forceinline__ void dc_st_async_v2f32(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a), "f"(b), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void dc_cbar() { named_bar_sync(1, DC_CTHREADS); }   // the compute warps
__device__ __forceinline__ float dc_ex2(float x) {   // the exponential of the other attention kernels
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// rows 8..15 of the A fragment are zero: only c[0], c[1] (row lane / 4, columns 2 (lane % 4), + 1) carry results
__device__ __forceinline__ void dc_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------------------------
// The operand stream of one compute warp.  Its chunk sequence is a pure function of (site list, t, rows of the cluster,
// CTA rank, warp): per attention sublayer [w_in: 2 chunks per projection] [per attention unit: K chunk, V chunk]
// [w_out: 2], per feed-forward sublayer [w_1: 4 column groups x 2] [w_2: 8].  A weight chunk is 8 weight rows x 256 k
// (4 KB), a K / V chunk 32 keys x 64 dims of the CTA's head.  An attention UNIT is (row of the cluster, 32-key chunk);
// the units of a sublayer are dealt round-robin to the warps (unit u -> warp u % 8).  The PRODUCER warp of the pair walks
// this sequence (dc_produce) as far ahead as the ring allows -- across sublayer boundaries and while the compute warp
// waits for the cluster -- and the compute warp takes the chunks in exactly the same order (dc_acquire / dc_release).
struct DcCtx {
  int n_sites, t, nrows, row0, rank, warp, lane;   // warp: index of the COMPUTE warp of the pair
  int R;                                           // rows per dialogue
  uint32_t ring, bars;                             // shared-memory addresses: this pair's ring, the barrier area
};

__device__ __forceinline__ int dc_site_lk(const DcCtx& c, const MtnDecodeSite& d) { return d.kind == 0 ? c.t : d.Lk; }
__device__ __forceinline__ int dc_units(const DcCtx& c, int Lk) {
  const int total = c.nrows * ((Lk + 31) >> 5);
  return total > c.warp ? (total - c.warp + 7) >> 3 : 0;
}
__device__ __forceinline__ void dc_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int DOFF, int SOFF>
__device__ __forceinline__ void dc_cp16i(uint32_t dst, const void* src) {   // immediate offsets: no address arithmetic per copy
  asm volatile("cp.async.cg.shared.global [%0 + %2], [%1 + %3], 16;" ::"r"(dst), "l"(src), "n"(DOFF), "n"(SOFF) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its cp.async issued so far have landed
__device__ __forceinline__ void dc_cp_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dc_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 8 weight rows x 512 k -> slot: lane l copies 16-byte segments l and l + 32 of every row.  ROWB = bytes between rows.
template <int ROWB>
__device__ __forceinline__ void dc_issue_w(uint32_t dst, const uint8_t* src) {
#define DC_W2(R) dc_cp16i<(R) * DC_WPITCH, (R) * ROWB>(dst, src); dc_cp16i<(R) * DC_WPITCH + 512, (R) * ROWB + 512>(dst, src);
  DC_W2(0) DC_W2(1) DC_W2(2) DC_W2(3) DC_W2(4) DC_W2(5) DC_W2(6) DC_W2(7)
#undef DC_W2
}

// producer warp: the whole chunk sequence of its compute warp, slot by slot (cp.async, 16 bytes per lane and copy; the
// slot's FULL barrier collects one deferred arrival per lane)
__device__ __forceinline__ void dc_produce(const DcCtx& c, const DcTable& tab, const void* gen_w, int gen_groups, int pf_on) {
  uint32_t n = 0;
  const int lane = c.lane;
  const uint32_t full0 = c.bars + 8u * (DC_BAR_FULL + c.warp * DC_NSLOT), empty0 = c.bars + 8u * (DC_BAR_EMPTY + c.warp * DC_NSLOT);
  auto slot_wait = [&]() -> uint32_t {   // index of the next slot, free
    const uint32_t sl = n % DC_NSLOT;
    if (n >= DC_NSLOT) mbar_wait(empty0 + 8u * sl, ((n / DC_NSLOT) - 1u) & 1u);
    ++n;
    return sl;
  };
  auto weights = [&](const void* W, int K, int row, int nchunks) {   // rows [row, row + 8), nchunks k-blocks of 512
    const uint8_t* src = static_cast<const uint8_t*>(W) + (size_t)row * K * 2 + lane * 16;
#pragma unroll 1
    for (int kp = 0; kp < nchunks; ++kp) {
      const uint32_t sl = slot_wait();
      const uint32_t dst = c.ring + sl * DC_SLOT + lane * 16;
      if (K == DC_D) dc_issue_w<DC_D * 2>(dst, src + kp * 1024);
      else dc_issue_w<DC_DFF * 2>(dst, src + kp * 1024);
      dc_cp_arrive(full0 + 8u * sl);
    }
  };
  // (experiment, opt-in) K / V of a cross-attention sublayer -> L2, one sublayer ahead of its use.  Measured without
  // effect: the step is bound by its dependency chain, not by where the K / V come from.
  // Line i of the CTA's (row, key, K | V) lines goes to producer lane (i % 256).
  auto kv_prefetch = [&](const MtnDecodeSite& d) {
    const int lines = c.nrows * d.Lk * 2;
    for (int i = c.warp * 32 + lane; i < lines; i += 256) {
      const int kvsel = i & 1, rk = i >> 1;
      const int g = rk / d.Lk, key = rk - g * d.Lk;
      const uint8_t* a = static_cast<const uint8_t*>(kvsel ? d.v : d.k) +
                         ((size_t)((c.row0 + g) / c.R) * d.kv_batch_stride + (size_t)key * d.ld_kv + c.rank * DC_DK) * 2;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  };
  const bool pf = pf_on != 0;
#pragma unroll 1
  for (int s = 0; s < c.n_sites; ++s) {
    const MtnDecodeSite& d = tab.s[s];
    if (pf) {   // the next cross-attention sublayer after s (s itself if it is the first one)
      if (s == 0 && d.kind == 1) kv_prefetch(d);
      for (int s2 = s + 1; s2 < c.n_sites; ++s2)
        if (tab.s[s2].kind != 2) {
          if (tab.s[s2].kind == 1) kv_prefetch(tab.s[s2]);
          break;
        }
    }
    if (d.kind == 2) {
#pragma unroll 1
      for (int j = 0; j < 4; ++j) weights(d.w_in, DC_D, c.rank * 256 + (j * 8 + c.warp) * 8, 1);
      weights(d.w_out, DC_DFF, c.rank * 64 + c.warp * 8, 4);
      continue;
    }
    const int nin = d.kind == 0 ? 3 : 1;
#pragma unroll 1
    for (int pj = 0; pj < nin; ++pj) weights(d.w_in, DC_D, pj * DC_D + c.rank * 64 + c.warp * 8, 1);
    const int Lk = dc_site_lk(c, d);
    const int nck = (Lk + 31) >> 5;
    const int nun = dc_units(c, Lk);
    const long long vdelta = static_cast<const uint8_t*>(d.v) - static_cast<const uint8_t*>(d.k);
    const int seg = lane & 7;
#pragma unroll 1
    for (int j = 0; j < nun; ++j) {
      // attention unit (row g, keys [32 ck, 32 ck + 32)): 128 bytes of head `rank` per key, K then V; lane l copies
      // segment l % 8 of keys l / 8 + 4 it
      const int u = c.warp + 8 * j;
      const int g = u / nck, ck = u - g * nck;
      const int bidx = d.kind == 1 ? (c.row0 + g) / c.R : c.row0 + g;   // static memories are stored once per dialogue
      const uint8_t* base = static_cast<const uint8_t*>(d.k) + ((size_t)bidx * d.kv_batch_stride + c.rank * DC_DK + seg * 8) * 2;
      const uint32_t sl = slot_wait();
      const uint32_t dst = c.ring + sl * DC_SLOT + (lane >> 3) * DC_KPITCH + seg * 16;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint8_t* src = base + (size_t)min(ck * 32 + 4 * it + (lane >> 3), Lk - 1) * d.ld_kv * 2;
        dc_cp16(dst + it * 4 * DC_KPITCH, src);
        dc_cp16(dst + it * 4 * DC_KPITCH + DC_VOFF, src + vdelta);
      }
      dc_cp_arrive(full0 + 8u * sl);
    }
    weights(d.w_out, DC_D, c.rank * 64 + c.warp * 8, 1);
  }
  if (gen_w != nullptr)   // generator arg-max: 8-row groups of the vocabulary, group index = rank + 8 (warp + 8 i)
    for (int gi = c.rank + 8 * c.warp; gi < gen_groups; gi += 64) weights(gen_w, DC_D, gi * 8, 1);
  dc_cp_wait_all();
}

// compute warp: the next chunk of its stream (shared-memory address) / hand its slot back to the producer
__device__ __forceinline__ uint32_t dc_acquire(const DcCtx& c, uint32_t& taken) {
  const uint32_t sl = taken % DC_NSLOT;
  mbar_wait(c.bars + 8u * (DC_BAR_FULL + c.warp * DC_NSLOT + sl), (taken / DC_NSLOT) & 1u);
  return c.ring + sl * DC_SLOT;
}
__device__ __forceinline__ void dc_release(const DcCtx& c, uint32_t& taken) {
  __syncwarp();
  if (c.lane == 0) mbar_arrive(c.bars + 8u * (DC_BAR_EMPTY + c.warp * DC_NSLOT + (taken % DC_NSLOT)));
  ++taken;
}

// One 8-column group of a projection over K = 256 * KCH: A rows from shared memory (byte offset a_off, row stride lda),
// the group's weight rows from the stream.  Per 32-wide k chunk a thread takes 8 consecutive k of its row (A: row
// lane / 4; W: output column lane / 4) with one 16-byte load each; both mma k16 steps use the same logical -> actual k
// mapping for A and B (csrc/decode_rows.cu), so the products pair up.  Four independent accumulators.  Returns this
// thread's two outputs: row lane / 4, columns 2 (lane % 4), + 1 of the group.
template <int KCH>   // K = 512 * KCH
__device__ __forceinline__ float2 dc_proj(const DcCtx& c, uint32_t& taken, uint32_t smem_base, int a_off, int lda) {
  const int g = c.lane >> 2, q = c.lane & 3;
  // A fragment rows beyond the cluster's rows are zero: their quarter-warps (two rows each) do not touch shared memory
  const bool live = (g & ~1) < c.nrows;
  // EIGHT independent accumulators: a dependent mma.sync costs ~90 cycles here, and a chunk is 32 of them
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 1
  for (int kp = 0; kp < KCH; ++kp) {
    const uint32_t arow = smem_base + a_off + g * lda + (kp * 512 + 8 * q) * 2;
    uint32_t wrow = 0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      uint4 a[4], w[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) a[cc] = live ? dc_lds128(arow + (4 * h + cc) * 64) : make_uint4(0u, 0u, 0u, 0u);
      if (h == 0) wrow = dc_acquire(c, taken) + g * DC_WPITCH + (q << 4);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) w[cc] = dc_lds128(wrow + (4 * h + cc) * 64);
      if (h == 3) dc_release(c, taken);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        dc_mma(acc[2 * cc], a[cc].x, a[cc].y, w[cc].x, w[cc].y);
        dc_mma(acc[2 * cc + 1], a[cc].z, a[cc].w, w[cc].z, w[cc].w);
      }
    }
  }
  return make_float2(((acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0])) + ((acc[4][0] + acc[5][0]) + (acc[6][0] + acc[7][0])),
                     ((acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1])) + ((acc[4][1] + acc[5][1]) + (acc[6][1] + acc[7][1])));
}

// LayerNorm parameters of one row pass: lane l holds float4 (l + 32 i) of a_2 / b_2
struct DcLn {
  float4 a[4], b[4];
  float eps;
};
__device__ __forceinline__ void dc_ln_load(DcLn& ln, const float* a2, const float* b2, float eps, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ln.a[i] = __ldg(reinterpret_cast<const float4*>(a2) + lane + 32 * i);
    ln.b[i] = __ldg(reinterpret_cast<const float4*>(b2) + lane + 32 * i);
  }
  ln.eps = eps;
}
// The reference's LayerNorm (mtn.py:111-114: unbiased std, eps added to std) of one 512-wide row held in shared memory,
// with the arithmetic and summation order of layernorm_rows_kernel<4>.  One warp.
__device__ __forceinline__ void dc_ln_row(const float* xrow, const DcLn& ln, int lane, float4 (&o)[4]) {
  float4 v[4];
  float sm = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = reinterpret_cast<const float4*>(xrow)[lane + 32 * i];
    sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, off);
  const float mean = sm * (1.f / DC_D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float inv = 1.f / (sqrtf(ss * (1.f / (DC_D - 1))) + ln.eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[i].x = ln.a[i].x * v[i].x * inv + ln.b[i].x;
    o[i].y = ln.a[i].y * v[i].y * inv + ln.b[i].y;
    o[i].z = ln.a[i].z * v[i].z * inv + ln.b[i].z;
    o[i].w = ln.a[i].w * v[i].w * inv + ln.b[i].w;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DC_THREADS, 1) decode_cluster_kernel(const __grid_constant__ DcTable tab, const DcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (int)(blockIdx.x / DC_CS) * p.G;
  const int nrows = min(p.G, p.B - row0);
  float* xs = reinterpret_cast<float*>(smem + DC_OFF_XS);
  float* qs = reinterpret_cast<float*>(smem + DC_OFF_QS);
  float* pm = reinterpret_cast<float*>(smem + DC_OFF_PM);
  float* pl = reinterpret_cast<float*>(smem + DC_OFF_PL);
  float* po = reinterpret_cast<float*>(smem + DC_OFF_PO);
  const uint32_t bars = sbase + DC_OFF_BAR;
  constexpr float LOG2E = 1.4426950408889634f;
  const float c1 = 0.125f * LOG2E, t_masked = -1e9f * LOG2E;   // 1 / sqrt(d_k), d_k = 64

  pdl_launch_dependents();
  // ---- barriers; A operand buffers zeroed (rows >= nrows stay zero: their products are never stored)
  if (threadIdx.x == 0) {
    for (int i = 0; i < DC_CWARPS * DC_NSLOT; ++i) {
      mbar_init(bars + 8u * (DC_BAR_FULL + i), 32u);    // one deferred arrival per producer lane (cp.async.mbarrier.arrive.noinc)
      mbar_init(bars + 8u * (DC_BAR_EMPTY + i), 1u);
    }
    mbar_init(bars + 8u * DC_BAR_O, 1u);
    mbar_init(bars + 8u * DC_BAR_X, 1u);
    mbar_init(bars + 8u * DC_BAR_H, 1u);
    mbar_init(bars + 8u * DC_BAR_G, 1u);
    mbar_fence_init();
  }
  {
    uint4* z = reinterpret_cast<uint4*>(smem + DC_OFF_XN);
    for (int i = threadIdx.x; i < (DC_OFF_QS - DC_OFF_XN) / 16; i += DC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  pdl_wait();   // the embedding rows and the caches are written by preceding kernels
  if (warp < DC_CWARPS) {
    // residual rows: warp w loads row w (zeros beyond the cluster's rows)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (warp < nrows) v = __ldcg(reinterpret_cast<const float4*>(p.x_in + (size_t)(row0 + warp) * DC_D) + lane + 32 * i);
      reinterpret_cast<float4*>(xs + warp * DC_D)[lane + 32 * i] = v;
    }
  }
  __syncthreads();
  cluster_sync_all();   // every CTA of the cluster is running, its barriers initialised, before anyone writes into a peer

  DcCtx ctx;
  ctx.n_sites = p.n_sites; ctx.t = p.t; ctx.nrows = nrows; ctx.row0 = row0; ctx.R = p.R;
  ctx.rank = (int)rank; ctx.warp = warp & (DC_CWARPS - 1); ctx.lane = lane;
  ctx.ring = sbase + DC_OFF_RING + ctx.warp * (DC_NSLOT * DC_SLOT);
  ctx.bars = bars;

  if (warp >= DC_CWARPS) {
    dc_produce(ctx, tab, p.gen_w, p.gen_V8 >> 3, p.kv_prefetch);
  } else {
    uint32_t taken = 0;
    const int col = (int)rank * 64 + warp * 8 + 2 * q;   // this thread's output columns of a d-wide projection
    const uint32_t xs_s = sbase + DC_OFF_XS, ob_s = sbase + DC_OFF_OB, hid_s = sbase + DC_OFF_HID;
    // bytes every exchange delivers into one CTA: 8 CTAs x its rows x their 64 (256) columns
    const uint32_t o_bytes = (uint32_t)nrows * DC_D * 2, x_bytes = (uint32_t)nrows * DC_D * 4, h_bytes = (uint32_t)nrows * DC_DFF * 2;
    uint32_t n_o = 0, n_x = 0, n_h = 0;   // completed phases of the exchange barriers
    DcLn ln;
    dc_ln_load(ln, tab.s[0].ln_a, tab.s[0].ln_b, tab.s[0].ln_eps, lane);
    const bool stamping = p.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define DC_STAMP(k) do { if (stamping) p.stamps[s * 8 + (k)] = clock64(); } while (0)

#pragma unroll 1
    for (int s = 0; s < p.n_sites; ++s) {
      const MtnDecodeSite& d = tab.s[s];
      const int kind = d.kind;
      DC_STAMP(0);
      // ---- everything of this sublayer whose address is known now is requested before the LayerNorm: biases, mask words
      const float2 bo = __ldg(reinterpret_cast<const float2*>(d.b_out + col));
      float2 bi[3];
      bi[0] = __ldg(reinterpret_cast<const float2*>(d.b_in + (kind == 2 ? (int)rank * 256 + warp * 8 + 2 * q : col)));
      bi[1] = bi[2] = bi[0];
      if (kind == 0) {
        bi[1] = __ldg(reinterpret_cast<const float2*>(d.b_in + DC_D + col));
        bi[2] = __ldg(reinterpret_cast<const float2*>(d.b_in + 2 * DC_D + col));
      }
      const int Lk = kind == 2 ? 0 : dc_site_lk(ctx, d);
      const int nck = (Lk + 31) >> 5;
      const int nun = kind == 2 ? 0 : dc_units(ctx, Lk);
      // mask word of attention unit j in lane j (key padding mask, one query row per dialogue; NULL: all keys kept)
      uint32_t mwords = 0xffffffffu;
      if (kind == 1 && d.mask_bits != nullptr && lane < nun) {
        const int u = warp + 8 * lane;
        const int gu = u / nck;
        mwords = __ldg(d.mask_bits + (size_t)((row0 + gu) / p.R) * d.mask_words + (u - gu * nck));
      }
      if (lane < DC_G) pm[warp * DC_G + lane] = -CUDART_INF_F;
      // ---- LayerNorm: warp w normalises row w -> f16 A operand
      {
        float4 o[4];
        dc_ln_row(xs + warp * DC_D, ln, lane, o);
        uint8_t* xr = smem + DC_OFF_XN + warp * DC_LDA;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<uint2*>(xr)[lane + 32 * i] = make_uint2(pack_f16x2_sat(o[i].x, o[i].y), pack_f16x2_sat(o[i].z, o[i].w));
      }
      dc_cbar();
      DC_STAMP(1);
      if (kind == 2) {
        // ================================================================ feed-forward sublayer
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bars + 8u * DC_BAR_H, h_bytes);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int hc = (int)rank * 256 + (j * 8 + warp) * 8 + 2 * q;
          const float2 b1 = j == 0 ? bi[0] : __ldg(reinterpret_cast<const float2*>(d.b_in + hc));
          const float2 y = dc_proj<1>(ctx, taken, sbase, DC_OFF_XN, DC_LDA);
          const uint32_t h2 = pack_f16x2_sat(fmaxf(y.x + b1.x, 0.f), fmaxf(y.y + b1.y, 0.f));
          if (g < nrows) {
            const uint32_t a = hid_s + g * DC_LDH + hc * 2;
#pragma unroll
            for (uint32_t r = 0; r < DC_CS; ++r) dc_st_async_u32(dc_mapa(a, r), h2, dc_mapa(bars + 8u * DC_BAR_H, r));
          }
        }
        DC_STAMP(4);
        mbar_wait(bars + 8u * DC_BAR_H, n_h & 1u);
        ++n_h;
      } else {
        // ================================================================ attention sublayer
        const int nin = kind == 0 ? 3 : 1;
        if (threadIdx.x == 0) mbar_arrive_expect_tx(bars + 8u * DC_BAR_O, o_bytes);
#pragma unroll 1
        for (int pj = 0; pj < nin; ++pj) {
          const float2 y = dc_proj<1>(ctx, taken, sbase, DC_OFF_XN, DC_LDA);
          const float2 bb = pj == 0 ? bi[0] : (pj == 1 ? bi[1] : bi[2]);
          const uint32_t h2 = pack_f16x2_sat(y.x + bb.x, y.y + bb.y);
          if (g < nrows) {
            const float2 r = __half22float2(*reinterpret_cast<const __half2*>(&h2));
            *reinterpret_cast<float2*>(qs + (pj * DC_G + g) * DC_DK + warp * 8 + 2 * q) = r;
            if (kind == 0) {   // the new row of the self-attention cache: [Q | K | V] of position t
              __half* base = pj == 0 ? static_cast<__half*>(d.q_cache)
                                     : const_cast<__half*>(static_cast<const __half*>(pj == 1 ? d.k : d.v));
              if (base != nullptr)
                *reinterpret_cast<uint32_t*>(base + (size_t)(row0 + g) * d.kv_batch_stride + (size_t)p.t * d.ld_kv + col) = h2;
            }
          }
        }
        dc_cbar();   // q (new k, v) of every warp's columns are in shared memory
        DC_STAMP(2);
        {
          int gcur = -1;
          float m_run = -CUDART_INF_F, l_run = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll 1
          for (int j = 0; j < nun; ++j) {
            const int u = warp + 8 * j;
            const int gu = u / nck, ck = u - gu * nck;
            if (gu != gcur) {
              if (gcur >= 0) {
                if (lane == 0) pm[warp * DC_G + gcur] = m_run, pl[warp * DC_G + gcur] = l_run;
                *reinterpret_cast<float2*>(po + (warp * DC_G + gcur) * DC_DK + 2 * lane) = make_float2(o0, o1);
              }
              gcur = gu; m_run = -CUDART_INF_F; l_run = 0.f; o0 = 0.f; o1 = 0.f;
            }
            const uint32_t mw = __shfl_sync(0xffffffffu, mwords, j);
            const int key = ck * 32 + lane;
            // ---- scores: lane = key; four independent partial sums
            float4 qv[16];
            const float4* q4 = reinterpret_cast<const float4*>(qs + gu * DC_DK);
#pragma unroll
            for (int i = 0; i < 16; ++i) qv[i] = q4[i];
            const uint32_t ks = dc_acquire(ctx, taken);
            uint4 kr[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) kr[cc] = dc_lds128(ks + lane * DC_KPITCH + (cc << 4));
            float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const float4 qa = qv[2 * cc], qb = qv[2 * cc + 1];
              const __half2* hp = reinterpret_cast<const __half2*>(&kr[cc]);
              const float2 k0 = __half22float2(hp[0]), k1 = __half22float2(hp[1]), k2 = __half22float2(hp[2]), k3 = __half22float2(hp[3]);
              sc[0] = fmaf(qa.x, k0.x, fmaf(qa.y, k0.y, sc[0]));
              sc[1] = fmaf(qa.z, k1.x, fmaf(qa.w, k1.y, sc[1]));
              sc[2] = fmaf(qb.x, k2.x, fmaf(qb.y, k2.y, sc[2]));
              sc[3] = fmaf(qb.z, k3.x, fmaf(qb.w, k3.y, sc[3]));
            }
            const float sdot = (sc[0] + sc[1]) + (sc[2] + sc[3]);
            const bool keep = (mw >> lane) & 1u;
            const float tt = key < Lk ? (keep ? sdot * c1 : t_masked) : -CUDART_INF_F;
            float mx = tt;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run, mx);
            const float e = dc_ex2(tt - m_new);
            float sum = e;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            const float alpha = dc_ex2(m_run - m_new);
            l_run = l_run * alpha + sum;
            m_run = m_new;
            const float pr = __half2float(__float2half_rn(e));   // P rounded to f16 before P V, like the tensor-core path
            // ---- P V: lane = two output dims; two independent partial sums per dim
            uint32_t vv[32];
#pragma unroll
            for (int uu = 0; uu < 32; ++uu) vv[uu] = dc_lds32(ks + DC_VOFF + uu * DC_KPITCH + lane * 4);
            dc_release(ctx, taken);
            float oa0 = o0 * alpha, oa1 = o1 * alpha, ob0 = 0.f, ob1 = 0.f;
#pragma unroll
            for (int uu = 0; uu < 32; uu += 2) {
              const float2 va = __half22float2(*reinterpret_cast<const __half2*>(&vv[uu]));
              const float2 vb = __half22float2(*reinterpret_cast<const __half2*>(&vv[uu + 1]));
              const float pa = __shfl_sync(0xffffffffu, pr, uu), pb = __shfl_sync(0xffffffffu, pr, uu + 1);
              oa0 = fmaf(pa, va.x, oa0);
              oa1 = fmaf(pa, va.y, oa1);
              ob0 = fmaf(pb, vb.x, ob0);
              ob1 = fmaf(pb, vb.y, ob1);
            }
            o0 = oa0 + ob0;
            o1 = oa1 + ob1;
          }
          if (gcur >= 0) {
            if (lane == 0) pm[warp * DC_G + gcur] = m_run, pl[warp * DC_G + gcur] = l_run;
            *reinterpret_cast<float2*>(po + (warp * DC_G + gcur) * DC_DK + 2 * lane) = make_float2(o0, o1);
          }
        }
        dc_cbar();
        DC_STAMP(3);
        // ---- merge the warps' partial (max, sum, O) of row `warp` in a fixed order; self-attention adds the new key
        if (warp < nrows) {
          const int gr = warp;
          float m = -CUDART_INF_F;
#pragma unroll
          for (int w = 0; w < DC_CWARPS; ++w) m = fmaxf(m, pm[w * DC_G + gr]);
          float m_new = -CUDART_INF_F;
          float2 vnew = make_float2(0.f, 0.f);
          if (kind == 0) {
            const float2 qq = *reinterpret_cast<const float2*>(qs + (0 * DC_G + gr) * DC_DK + 2 * lane);
            const float2 kk = *reinterpret_cast<const float2*>(qs + (1 * DC_G + gr) * DC_DK + 2 * lane);
            vnew = *reinterpret_cast<const float2*>(qs + (2 * DC_G + gr) * DC_DK + 2 * lane);
            float sn = fmaf(qq.x, kk.x, qq.y * kk.y);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sn += __shfl_xor_sync(0xffffffffu, sn, off);
            m_new = sn * c1;
            m = fmaxf(m, m_new);
          }
          float l = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int w = 0; w < DC_CWARPS; ++w) {
            const float pmw = pm[w * DC_G + gr];
            if (pmw > -CUDART_INF_F) {
              const float f = dc_ex2(pmw - m);
              const float2 ov = *reinterpret_cast<const float2*>(po + (w * DC_G + gr) * DC_DK + 2 * lane);
              l = fmaf(pl[w * DC_G + gr], f, l);
              a0 = fmaf(ov.x, f, a0);
              a1 = fmaf(ov.y, f, a1);
            }
          }
          if (kind == 0) {
            const float f = dc_ex2(m_new - m);
            l += f;
            a0 = fmaf(vnew.x, f, a0);
            a1 = fmaf(vnew.y, f, a1);
          }
          const float inv = 1.f / l;
          const uint32_t h2 = pack_f16x2_sat(a0 * inv, a1 * inv);
          const uint32_t a = ob_s + gr * DC_LDA + ((int)rank * DC_DK + 2 * lane) * 2;
#pragma unroll
          for (uint32_t r = 0; r < DC_CS; ++r) dc_st_async_u32(dc_mapa(a, r), h2, dc_mapa(bars + 8u * DC_BAR_O, r));
        }
        DC_STAMP(4);
        mbar_wait(bars + 8u * DC_BAR_O, n_o & 1u);
        ++n_o;
      }
      DC_STAMP(5);
      // ---- output projection (attention: A = all heads' outputs; feed-forward: A = hidden activation) + residual
      if (threadIdx.x == 0) mbar_arrive_expect_tx(bars + 8u * DC_BAR_X, x_bytes);
      {
        const float2 y = kind == 2 ? dc_proj<4>(ctx, taken, sbase, DC_OFF_HID, DC_LDH) : dc_proj<1>(ctx, taken, sbase, DC_OFF_OB, DC_LDA);
        if (g < nrows) {
          const float2 xo = *reinterpret_cast<const float2*>(xs + g * DC_D + col);
          const float x0 = xo.x + (y.x + bo.x), x1 = xo.y + (y.y + bo.y);
          const uint32_t a = xs_s + (g * DC_D + col) * 4;
#pragma unroll
          for (uint32_t r = 0; r < DC_CS; ++r) dc_st_async_v2f32(dc_mapa(a, r), x0, x1, dc_mapa(bars + 8u * DC_BAR_X, r));
        }
      }
      DC_STAMP(6);
      if (s + 1 < p.n_sites) dc_ln_load(ln, tab.s[s + 1].ln_a, tab.s[s + 1].ln_b, tab.s[s + 1].ln_eps, lane);
      else dc_ln_load(ln, p.norm_a, p.norm_b, p.norm_eps, lane);
      mbar_wait(bars + 8u * DC_BAR_X, n_x & 1u);
      ++n_x;
      DC_STAMP(7);
      if (p.taps != nullptr && rank == 0 && warp < nrows) {
        float4* tp = reinterpret_cast<float4*>(p.taps + ((size_t)s * p.B + row0 + warp) * DC_D);
#pragma unroll
        for (int i = 0; i < 4; ++i) tp[lane + 32 * i] = reinterpret_cast<const float4*>(xs + warp * DC_D)[lane + 32 * i];
      }
    }
    // ---- final LayerNorm (mtn.py:164) -> out (CTA 0) and, for the generator stage, the f16 A operand (every CTA)
    {
      float4 o[4];
      dc_ln_row(xs + warp * DC_D, ln, lane, o);
      if (rank == 0 && warp < nrows) {
        float4* op = reinterpret_cast<float4*>(p.out + (size_t)(row0 + warp) * DC_D);
#pragma unroll
        for (int i = 0; i < 4; ++i) op[lane + 32 * i] = o[i];
      }
      uint8_t* xr = smem + DC_OFF_XN + warp * DC_LDA;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint2*>(xr)[lane + 32 * i] = make_uint2(pack_f16x2_sat(o[i].x, o[i].y), pack_f16x2_sat(o[i].z, o[i].w));
    }
    if (p.gen_w != nullptr) {
      // ---- generator arg-max (Generator.forward mtn.py:68-69 + the arg-max of data_utils.py:183: log_softmax is monotone,
      // so the arg-max of the logits): vocabulary groups of 8 dealt over the cluster's 64 warps; a key packs (orderable
      // logit, ~column) so that the maximum key is the largest logit, the FIRST column among equal ones (torch.argmax)
      dc_cbar();
      if (rank == 0 && threadIdx.x == 0) mbar_arrive_expect_tx(bars + 8u * DC_BAR_G, (uint32_t)(DC_CS * nrows * 8));
      unsigned long long best = 0ull;
      auto key_of = [](float v, int colx) -> unsigned long long {
        uint32_t u = __float_as_uint(v);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // monotone map f32 -> u32
        return ((unsigned long long)u << 32) | (uint32_t)(0x7fffffff - colx);
      };
      const int ngroups = p.gen_V8 >> 3;
#pragma unroll 1
      for (int gi = (int)rank + 8 * warp; gi < ngroups; gi += 64) {
        const int c0 = gi * 8 + 2 * q;
        const float2 bb = __ldg(reinterpret_cast<const float2*>(p.gen_b + c0));
        const float2 y = dc_proj<1>(ctx, taken, sbase, DC_OFF_XN, DC_LDA);
        if (c0 < p.gen_V) best = max(best, key_of(y.x + bb.x, c0));
        if (c0 + 1 < p.gen_V) best = max(best, key_of(y.y + bb.y, c0 + 1));
      }
      best = max(best, __shfl_xor_sync(0xffffffffu, best, 1));
      best = max(best, __shfl_xor_sync(0xffffffffu, best, 2));
      unsigned long long* wbest = reinterpret_cast<unsigned long long*>(smem + DC_OFF_PM);   // [8 warps][8 rows] (pm / pl are dead)
      if (q == 0) wbest[warp * DC_G + g] = best;
      dc_cbar();
      if (threadIdx.x < nrows) {
        unsigned long long b8 = 0ull;
#pragma unroll
        for (int w = 0; w < DC_CWARPS; ++w) b8 = max(b8, wbest[w * DC_G + threadIdx.x]);
        const uint32_t a = sbase + DC_OFF_GBUF + ((int)rank * DC_G + threadIdx.x) * 8;
        asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];" ::"r"(dc_mapa(a, 0)), "l"(b8),
                     "r"(dc_mapa(bars + 8u * DC_BAR_G, 0))
                     : "memory");
      }
      if (rank == 0) {
        mbar_wait(bars + 8u * DC_BAR_G, 0u);
        if (threadIdx.x < nrows) {
          const unsigned long long* gb = reinterpret_cast<const unsigned long long*>(smem + DC_OFF_GBUF);
          unsigned long long b8 = 0ull;
#pragma unroll
          for (int r = 0; r < DC_CS; ++r) b8 = max(b8, gb[r * DC_G + threadIdx.x]);
          p.tokens[(size_t)(row0 + threadIdx.x) * p.tokens_stride] = (long long)(0x7fffffff - (int)(uint32_t)(b8 & 0xffffffffu));
        }
      }
    }
  }
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while a peer could still address its shared memory
}

static int dc_kv_prefetch_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTN_B200_DECODE_KV_PREFETCH");   // opt-in: measured without effect (380 vs 375 us per step)
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v;
}

static int dc_max_clusters(int* out) {
  static int cached = 0;
  if (cached == 0) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(DC_CS * 16);
    cfg.blockDim = dim3(DC_THREADS);
    cfg.dynamicSmemBytes = DC_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = DC_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    MTN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, decode_cluster_kernel, &cfg));
    MTN_REQUIRE(n > 0, MTN_E_CUDA, "decode_cluster: no cluster of %d CTAs with %d bytes of shared memory fits this device", DC_CS, DC_SMEM);
    cached = n;
  }
  *out = cached;
  return MTN_OK;
}

}  // namespace mtn

// B rows fit when ceil(B / co-resident clusters) <= 8 rows per cluster.  The cluster count is a property of the device (13 on
// the B200s this was measured on: 104 rows); without a usable device (host-only callers, tests) 16 clusters are assumed.
extern "C" int mtn_decode_cluster_supported(int B, int d, int h, int d_ff) {
  if (!(B >= 1 && d == mtn::DC_D && h == mtn::DC_CS && d_ff == mtn::DC_DFF)) return 0;
  int clusters = 0, ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || mtn::dc_max_clusters(&clusters) != MTN_OK) {
    cudaGetLastError();
    clusters = 16;
  }
  if (clusters > 16) clusters = 16;
  return B <= clusters * mtn::DC_G ? 1 : 0;
}

extern "C" int mtn_decode_cluster_max_sites(void) { return mtn::DC_MAX_SITES; }

extern "C" int mtn_decode_cluster_fwd(const MtnDecodeClusterArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->sites && a->x_in && a->out && a->norm_a && a->norm_b, MTN_E_ARG, "decode_cluster: NULL pointer");
  MTN_REQUIRE(mtn_decode_cluster_supported(a->B, a->d, a->h, a->d_ff), MTN_E_SHAPE,
              "decode_cluster: B=%d d=%d h=%d d_ff=%d (d = 512, h = 8, d_ff = 2048, B <= 8 rows x co-resident clusters)", a->B, a->d, a->h,
              a->d_ff);
  MTN_REQUIRE(a->rows_per_dialogue <= 1 || a->B % a->rows_per_dialogue == 0, MTN_E_SHAPE, "decode_cluster: B=%d is not a multiple of rows_per_dialogue=%d",
              a->B, a->rows_per_dialogue);
  MTN_REQUIRE(a->n_sites >= 1 && a->n_sites <= DC_MAX_SITES && a->t >= 0, MTN_E_SHAPE, "decode_cluster: n_sites=%d (<= %d), t=%d", a->n_sites,
              DC_MAX_SITES, a->t);
  MTN_REQUIRE(aligned16(a->x_in) && aligned16(a->out) && aligned16(a->norm_a) && aligned16(a->norm_b) && (!a->taps || aligned16(a->taps)),
              MTN_E_ALIGN, "decode_cluster: x_in / out / norm / taps must be 16-byte aligned");
  if (a->gen_w != nullptr) {
    MTN_REQUIRE(a->gen_b && a->tokens && a->gen_V >= 1 && a->gen_V8 >= a->gen_V && a->gen_V8 % 8 == 0 && a->tokens_stride >= 1, MTN_E_ARG,
                "decode_cluster: generator stage: bias / tokens NULL or V=%d V8=%d stride=%lld", a->gen_V, a->gen_V8, a->tokens_stride);
    MTN_REQUIRE(aligned16(a->gen_w) && (reinterpret_cast<uintptr_t>(a->gen_b) & 7) == 0 && (reinterpret_cast<uintptr_t>(a->tokens) & 7) == 0,
                MTN_E_ALIGN, "decode_cluster: generator stage: alignment");
  }
  DcTable tab;
  memset(&tab, 0, sizeof(tab));
  for (int s = 0; s < a->n_sites; ++s) {
    const MtnDecodeSite& d = a->sites[s];
    MTN_REQUIRE(d.kind >= 0 && d.kind <= 2 && d.ln_a && d.ln_b && d.w_in && d.b_in && d.w_out && d.b_out, MTN_E_ARG,
                "decode_cluster: site %d: kind %d / NULL pointer", s, d.kind);
    MTN_REQUIRE(aligned16(d.ln_a) && aligned16(d.ln_b) && aligned16(d.w_in) && aligned16(d.w_out) &&
                    (reinterpret_cast<uintptr_t>(d.b_in) & 7) == 0 && (reinterpret_cast<uintptr_t>(d.b_out) & 7) == 0,
                MTN_E_ALIGN, "decode_cluster: site %d: parameter alignment", s);
    if (d.kind != 2) {
      MTN_REQUIRE(d.k && d.v && aligned16(d.k) && aligned16(d.v) && d.ld_kv % 8 == 0 && d.kv_batch_stride % 8 == 0 && d.ld_kv >= DC_D,
                  MTN_E_ALIGN, "decode_cluster: site %d: K / V pointers, ld_kv=%d, batch stride %lld", s, d.ld_kv, d.kv_batch_stride);
      if (d.kind == 1) {
        MTN_REQUIRE(d.Lk >= 1 && d.Lk <= 1024, MTN_E_SHAPE, "decode_cluster: site %d: Lk=%d (1..1024)", s, d.Lk);
        MTN_REQUIRE(d.mask_bits == nullptr || d.mask_words >= (d.Lk + 31) / 32, MTN_E_SHAPE, "decode_cluster: site %d: mask_words=%d", s,
                    d.mask_words);
      } else {
        MTN_REQUIRE(a->t <= 1024 && (d.q_cache == nullptr || (reinterpret_cast<uintptr_t>(d.q_cache) & 3) == 0), MTN_E_SHAPE,
                    "decode_cluster: site %d: t=%d", s, a->t);
      }
    }
    tab.s[s] = d;
  }
  int max_clusters = 0;
  const int rc = dc_max_clusters(&max_clusters);
  if (rc != MTN_OK) return rc;
  if (max_clusters > 16) max_clusters = 16;
  const int G = (a->B + max_clusters - 1) / max_clusters;
  MTN_REQUIRE(G <= DC_G, MTN_E_SHAPE, "decode_cluster: B=%d needs %d rows per cluster (%d clusters fit), at most %d", a->B, G, max_clusters, DC_G);
  const int nclusters = (a->B + G - 1) / G;
  DcParams p{a->n_sites, a->B, G, a->t, a->rows_per_dialogue > 1 ? a->rows_per_dialogue : 1, a->x_in, a->out, a->norm_a, a->norm_b, a->norm_eps, a->taps, a->stamps,
             static_cast<const __half*>(a->gen_w), a->gen_b, a->gen_V, a->gen_V8, reinterpret_cast<long long*>(a->tokens), a->tokens_stride, dc_kv_prefetch_enabled()};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTN_CHECK_CUDA(launch_kernel_cluster(decode_cluster_kernel, dim3(nclusters * DC_CS), dim3(DC_THREADS), DC_SMEM, st, DC_CS, tab, p));
  return MTN_OK;
}
