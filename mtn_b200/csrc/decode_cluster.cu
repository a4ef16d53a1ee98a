// One KV-cached decoding step (reference call form data_utils.py:202-210, one new target position per
// dialogue) as ONE kernel in which every DIALOGUE GROUP is owned by one thread-block cluster.
//
// Why: a cached step is a chain of ~135 dependent few-row launches (csrc/decode_rows.cu), each a kernel boundary (~3 us)
// plus one exposed memory round trip; a grid-wide persistent kernel (decode_prog_kernel) pays a grid barrier (three L2
// round trips) per stage instead and is no faster.  But the dependency chain of a decoding step is PER DIALOGUE: nothing
// in DecoderLayer.forward (mtn.py:181-218) mixes dialogues.  So a cluster of 8 CTAs takes G = ceil(B / #clusters)
// dialogues through the whole N-layer step; CTA `rank` owns head `rank` (d_k = 64) and the matching 64 output columns of
// every projection (256 of the feed-forward hidden layer), and the only synchronisation is the hardware cluster barrier:
//
//   per attention sublayer (mtn.py:125-127 around :248-267 around :221-231)
//     LayerNorm of the G residual rows (every CTA holds the full rows, f32, in shared memory)
//     q (self: q, k, v) of head `rank`:  mma.sync over the CTA's 64 weight rows (8 per warp)
//     attention of head `rank`: the static memory's K / V (or the self-attention cache rows 0..t-1 plus the new row,
//       which is also appended to the cache), online softmax in the log2 domain with the reference's FINITE -1e9
//     the head's output -> every CTA of the cluster (st.shared::cluster), cluster barrier
//     output projection of the CTA's 64 columns + residual -> every CTA's copy of the rows, cluster barrier
//   feed-forward sublayer (mtn.py:279-280): LayerNorm, w_1 + ReLU (256 hidden columns per CTA) -> every CTA, barrier,
//     w_2 (64 columns, K = 2048) + residual -> every CTA, barrier.
//
// Everything a CTA reads from global memory inside the step -- its weight rows and its head's K / V rows -- has an
// address that does not depend on computed data, so each WARP streams its own operands through a private ring of
// cp.async buffers (4 x 4 KB), running up to three chunks ahead ACROSS sublayer boundaries: the weights of the next
// sublayer arrive while the current one waits at its barriers.  No kernel boundary, no grid barrier, no exposed memory
// round trip on the chain; HBM sees each K / V byte once, the f16 weights (72 MB) come out of L2 once per cluster.
//
// Arithmetic = the few-row kernels' (f16 operands, f32 accumulate, f16-rounded q / k / v / P / O / hidden, LayerNorm
// with the summation order of layernorm_rows_kernel); only the summation order of the projections differs (one warp
// accumulates the whole contraction).  d = 512, h = 8, d_ff = 2048, one query row per dialogue (greedy decoding).
#include <math_constants.h>
#include <string.h>

#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int DC_CS = 8;            // CTAs per cluster = heads
constexpr int DC_WARPS = 8, DC_THREADS = 32 * DC_WARPS;
constexpr int DC_G = 8;             // dialogues (rows) per cluster, at most: the m16n8k16 fragments carry rows 0..7
constexpr int DC_D = 512, DC_DFF = 2048, DC_DK = 64;
constexpr int DC_NSLOT = 4, DC_SLOT = 4096;
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
constexpr int DC_OFF_QS = DC_OFF_HID + DC_G * DC_LDH;             // f32 [3][8][64] q, new k, new v of this head (f16-rounded)
constexpr int DC_OFF_PM = DC_OFF_QS + 3 * DC_G * DC_DK * 4;       // f32 [8 warps][8]      partial row maxima
constexpr int DC_OFF_PL = DC_OFF_PM + DC_WARPS * DC_G * 4;        // f32 [8 warps][8]      partial row sums
constexpr int DC_OFF_PO = DC_OFF_PL + DC_WARPS * DC_G * 4;        // f32 [8 warps][8][64]  partial outputs
constexpr int DC_OFF_SITES = DC_OFF_PO + DC_WARPS * DC_G * DC_DK * 4;
constexpr int DC_OFF_RING = (DC_OFF_SITES + DC_MAX_SITES * (int)sizeof(MtnDecodeSite) + 1023) / 1024 * 1024;
constexpr int DC_SMEM = DC_OFF_RING + DC_WARPS * DC_NSLOT * DC_SLOT;
static_assert(DC_SMEM <= 232448, "shared memory budget");
static_assert(sizeof(MtnDecodeSite) % 8 == 0, "site descriptors are copied word-wise");

struct DcTable {
  MtnDecodeSite s[DC_MAX_SITES];
};

struct DcParams {
  int n_sites, B, G, t;
  const float* x_in;
  float* out;
  const float* norm_a;
  const float* norm_b;
  float norm_eps;
  float* taps;
  long long* stamps;   // optional [n_sites][8] clock64 stamps of CTA 0, warp 0 (debug: where a sublayer spends its time)
};

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dc_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void dc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dc_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 dc_lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t dc_lds32(uint32_t a) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ uint32_t dc_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dc_st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void dc_st_cluster_v2f32(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void dc_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void dc_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
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
// The operand stream of one warp.  Its chunk sequence is a pure function of (site list, t, rows of the cluster, CTA
// rank, warp): per attention sublayer [w_in: 2 chunks per projection] [per attention unit: K chunk, V chunk] [w_out: 2],
// per feed-forward sublayer [w_1: 4 column groups x 2] [w_2: 8].  A weight chunk is 8 weight rows x 256 k (4 KB), a K / V
// chunk 32 keys x 64 dims of the CTA's head.  An attention UNIT is (row of the cluster, 32-key chunk); the units of a
// sublayer are dealt round-robin to the warps (unit u -> warp u % 8).  The consumer code below takes chunks in exactly
// this order; dc_phase_count is the single source of truth for how many chunks a phase has.
struct DcCtx {
  const MtnDecodeSite* sites;   // shared-memory copy
  int n_sites, t, nrows, row0, rank, warp, lane;
  uint32_t ring;                // shared-memory address of this warp's ring
};
struct DcStream {
  int s, ph, i;                 // producer cursor: site, phase, chunk inside the phase
  uint32_t issued, taken;
};

__device__ __forceinline__ int dc_site_lk(const DcCtx& c, const MtnDecodeSite& d) { return d.kind == 0 ? c.t : d.Lk; }
__device__ __forceinline__ int dc_units(const DcCtx& c, int Lk) {
  const int total = c.nrows * ((Lk + 31) >> 5);
  return total > c.warp ? (total - c.warp + 7) >> 3 : 0;
}
__device__ __forceinline__ int dc_phases(const MtnDecodeSite& d) { return d.kind == 2 ? 2 : (d.kind == 0 ? 5 : 3); }
__device__ __forceinline__ int dc_phase_count(const DcCtx& c, const MtnDecodeSite& d, int ph) {
  if (d.kind == 2) return 8;
  const int nin = d.kind == 0 ? 3 : 1;
  if (ph == nin) return 2 * dc_units(c, dc_site_lk(c, d));
  return 2;
}

__device__ __forceinline__ void dc_issue(const DcCtx& c, DcStream& st) {
  if (st.s < c.n_sites) {
    const MtnDecodeSite& d = c.sites[st.s];
    const uint32_t slot = c.ring + (st.issued % DC_NSLOT) * DC_SLOT;
    const int lane = c.lane;
    const int nin = d.kind == 0 ? 3 : 1;
    if (d.kind == 2 || st.ph != nin) {
      // ---- weight chunk: rows [row, row + 8) x k [256 kp, 256 kp + 256); lane l copies 16-byte segment l of each row.
      // Row i lands at i * 512 with its segment index XORed by 4 for odd rows (conflict-free fragment loads).
      const __half* W;
      int K, row, kp;
      if (d.kind == 2) {
        if (st.ph == 0) { W = static_cast<const __half*>(d.w_in); K = DC_D; row = c.rank * 256 + ((st.i >> 1) * 8 + c.warp) * 8; kp = st.i & 1; }
        else { W = static_cast<const __half*>(d.w_out); K = DC_DFF; row = c.rank * 64 + c.warp * 8; kp = st.i; }
      } else if (st.ph < nin) {
        W = static_cast<const __half*>(d.w_in); K = DC_D; row = st.ph * DC_D + c.rank * 64 + c.warp * 8; kp = st.i;
      } else {
        W = static_cast<const __half*>(d.w_out); K = DC_D; row = c.rank * 64 + c.warp * 8; kp = st.i;
      }
      const __half* src = W + (size_t)row * K + kp * 256 + lane * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) dc_cp16(slot + i * 512 + (((uint32_t)lane ^ ((uint32_t)(i & 1) << 2)) << 4), src + (size_t)i * K);
    } else {
      // ---- K (even i) or V (odd i) chunk of attention unit st.i / 2: 32 keys x 128 bytes of head `rank`.  K rows are
      // stored with their 16-byte segment index XORed by (key % 8): lane = key reads are conflict-free.
      const int Lk = dc_site_lk(c, d);
      const int nck = (Lk + 31) >> 5;
      const int u = c.warp + 8 * (st.i >> 1);
      const int g = u / nck, ck = u - g * nck;
      const bool isv = (st.i & 1) != 0;
      const int seg = lane & 7;
      const __half* base = static_cast<const __half*>(isv ? d.v : d.k) + (size_t)(c.row0 + g) * d.kv_batch_stride + c.rank * DC_DK + seg * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + (lane >> 3);
        const int key = min(ck * 32 + r, Lk - 1);
        const uint32_t sg = isv ? (uint32_t)seg : ((uint32_t)seg ^ (uint32_t)(r & 7));
        dc_cp16(slot + r * 128 + (sg << 4), base + (size_t)key * d.ld_kv);
      }
    }
    ++st.i;
    while (st.s < c.n_sites && st.i >= dc_phase_count(c, c.sites[st.s], st.ph)) {
      st.i = 0;
      if (++st.ph >= dc_phases(c.sites[st.s])) { st.ph = 0; ++st.s; }
    }
  }
  dc_cp_commit();   // (an empty group past the end keeps the group arithmetic of dc_acquire uniform)
  ++st.issued;
}

// The next chunk of this warp's stream: waits for it, refills the slot consumed BEFORE it (so the caller holds exactly
// one chunk at a time) and returns its shared-memory address.
__device__ __forceinline__ uint32_t dc_acquire(const DcCtx& c, DcStream& st) {
  dc_cp_wait<DC_NSLOT - 2>();
  __syncwarp();
  dc_issue(c, st);
  const uint32_t slot = c.ring + (st.taken % DC_NSLOT) * DC_SLOT;
  ++st.taken;
  return slot;
}

// One 8-column group of a projection over K = 256 * KCH: A rows from shared memory (byte offset a_off, row stride lda),
// the group's weight rows from the stream.  Per 32-wide k chunk a thread takes 8 consecutive k of its row (A: row
// lane / 4; W: output column lane / 4) with one 16-byte load each; both mma k16 steps use the same logical -> actual k
// mapping for A and B (csrc/decode_rows.cu), so the products pair up.  Returns this thread's two outputs: row lane / 4,
// columns 2 (lane % 4), + 1 of the group.
template <int KCH>
__device__ __forceinline__ float2 dc_proj(const DcCtx& c, DcStream& st, uint32_t smem_base, int a_off, int lda) {
  const int g = c.lane >> 2, q = c.lane & 3;
  float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t wsw = (uint32_t)(g & 1) << 2;
#pragma unroll 1
  for (int kp = 0; kp < KCH; ++kp) {
    const uint32_t slot = dc_acquire(c, st);
    const uint32_t arow = smem_base + a_off + g * lda + (kp * 256 + 8 * q) * 2;
    const uint32_t wrow = slot + g * 512;
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const uint4 w = dc_lds128(wrow + ((((uint32_t)(4 * cc + q)) ^ wsw) << 4));
      const uint4 a = dc_lds128(arow + cc * 64);
      if (cc & 1) {
        dc_mma(acc1, a.x, a.y, w.x, w.y);
        dc_mma(acc1, a.z, a.w, w.z, w.w);
      } else {
        dc_mma(acc0, a.x, a.y, w.x, w.y);
        dc_mma(acc0, a.z, a.w, w.z, w.w);
      }
    }
  }
  return make_float2(acc0[0] + acc1[0], acc0[1] + acc1[1]);
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
  MtnDecodeSite* sites = reinterpret_cast<MtnDecodeSite*>(smem + DC_OFF_SITES);
  constexpr float LOG2E = 1.4426950408889634f;
  const float c1 = 0.125f * LOG2E, t_masked = -1e9f * LOG2E;   // 1 / sqrt(d_k), d_k = 64

  pdl_launch_dependents();
  // ---- site table -> shared memory; A operand buffers zeroed (rows >= nrows stay zero: their products are never stored)
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&tab);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sites);
    const int words = p.n_sites * (int)(sizeof(MtnDecodeSite) / 4);
    for (int i = threadIdx.x; i < words; i += DC_THREADS) dst[i] = src[i];
    uint4* z = reinterpret_cast<uint4*>(smem + DC_OFF_XN);
    for (int i = threadIdx.x; i < (DC_OFF_QS - DC_OFF_XN) / 16; i += DC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  pdl_wait();   // the embedding rows and the caches are written by preceding kernels
  {
    // residual rows: warp w loads row w (zeros beyond the cluster's rows)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (warp < nrows) v = __ldcg(reinterpret_cast<const float4*>(p.x_in + (size_t)(row0 + warp) * DC_D) + lane + 32 * i);
      reinterpret_cast<float4*>(xs + warp * DC_D)[lane + 32 * i] = v;
    }
  }
  __syncthreads();
  dc_cluster_arrive();   // every CTA of the cluster is running and initialised before anyone writes into a peer
  dc_cluster_wait();

  DcCtx ctx;
  ctx.sites = sites; ctx.n_sites = p.n_sites; ctx.t = p.t; ctx.nrows = nrows; ctx.row0 = row0;
  ctx.rank = (int)rank; ctx.warp = warp; ctx.lane = lane;
  ctx.ring = sbase + DC_OFF_RING + warp * (DC_NSLOT * DC_SLOT);
  DcStream st = {0, 0, 0, 0u, 0u};
#pragma unroll 1
  for (int i = 0; i < DC_NSLOT - 1; ++i) dc_issue(ctx, st);

  const int col = (int)rank * 64 + warp * 8 + 2 * q;   // this thread's output columns of a d-wide projection
  const uint32_t xs_s = sbase + DC_OFF_XS, ob_s = sbase + DC_OFF_OB, hid_s = sbase + DC_OFF_HID;
  DcLn ln;
  dc_ln_load(ln, sites[0].ln_a, sites[0].ln_b, sites[0].ln_eps, lane);

  const bool stamping = p.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define DC_STAMP(k) do { if (stamping) p.stamps[s * 8 + (k)] = clock64(); } while (0)
#pragma unroll 1
  for (int s = 0; s < p.n_sites; ++s) {
    const MtnDecodeSite& d = sites[s];
    DC_STAMP(0);
    // ---- LayerNorm: warp w normalises row w -> f16 A operand
    {
      float4 o[4];
      dc_ln_row(xs + warp * DC_D, ln, lane, o);
      uint8_t* xr = smem + DC_OFF_XN + warp * DC_LDA;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint2*>(xr)[lane + 32 * i] = make_uint2(pack_f16x2_sat(o[i].x, o[i].y), pack_f16x2_sat(o[i].z, o[i].w));
    }
    __syncthreads();
    DC_STAMP(1);
    float2 bo;
    if (d.kind == 2) {
      // ================================================================ feed-forward sublayer
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int hc = (int)rank * 256 + (j * 8 + warp) * 8 + 2 * q;
        const float2 b1 = __ldg(reinterpret_cast<const float2*>(d.b_in + hc));
        float2 y = dc_proj<2>(ctx, st, sbase, DC_OFF_XN, DC_LDA);
        const uint32_t h2 = pack_f16x2_sat(fmaxf(y.x + b1.x, 0.f), fmaxf(y.y + b1.y, 0.f));
        if (g < nrows) {
          const uint32_t a = hid_s + g * DC_LDH + hc * 2;
#pragma unroll
          for (uint32_t r = 0; r < DC_CS; ++r) dc_st_cluster_u32(dc_mapa(a, r), h2);
        }
      }
      DC_STAMP(4);
      dc_cluster_arrive();
      bo = __ldg(reinterpret_cast<const float2*>(d.b_out + col));
      dc_cluster_wait();
    } else {
      // ================================================================ attention sublayer
      const int nin = d.kind == 0 ? 3 : 1;
#pragma unroll 1
      for (int pj = 0; pj < nin; ++pj) {
        const float2 bi = __ldg(reinterpret_cast<const float2*>(d.b_in + pj * DC_D + col));
        float2 y = dc_proj<2>(ctx, st, sbase, DC_OFF_XN, DC_LDA);
        const uint32_t h2 = pack_f16x2_sat(y.x + bi.x, y.y + bi.y);
        if (g < nrows) {
          const float2 r = __half22float2(*reinterpret_cast<const __half2*>(&h2));
          *reinterpret_cast<float2*>(qs + (pj * DC_G + g) * DC_DK + warp * 8 + 2 * q) = r;
          if (d.kind == 0) {   // the new row of the self-attention cache: [Q | K | V] of position t
            __half* base = pj == 0 ? static_cast<__half*>(d.q_cache)
                                   : const_cast<__half*>(static_cast<const __half*>(pj == 1 ? d.k : d.v));
            if (base != nullptr)
              *reinterpret_cast<uint32_t*>(base + (size_t)(row0 + g) * d.kv_batch_stride + (size_t)p.t * d.ld_kv + col) = h2;
          }
        }
      }
      const int Lk = dc_site_lk(ctx, d);
      const int nck = (Lk + 31) >> 5;
      const int nun = dc_units(ctx, Lk);
      // mask word of unit j in lane j (key padding mask, one query row per dialogue; NULL: all keys kept)
      uint32_t mwords = 0xffffffffu;
      if (d.mask_bits != nullptr && lane < nun) {
        const int u = warp + 8 * lane;
        const int gu = u / nck;
        mwords = __ldg(d.mask_bits + (size_t)(row0 + gu) * d.mask_words + (u - gu * nck));
      }
      if (lane < DC_G) pm[warp * DC_G + lane] = -CUDART_INF_F;
      __syncthreads();   // q (new k, v) of every warp's columns are in shared memory
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
          // ---- scores: lane = key
          const uint32_t ks = dc_acquire(ctx, st);
          float sc = 0.f;
          const float4* q4 = reinterpret_cast<const float4*>(qs + gu * DC_DK);
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const uint4 kv = dc_lds128(ks + lane * 128 + (((uint32_t)cc ^ (uint32_t)(lane & 7)) << 4));
            const float4 qa = q4[2 * cc], qb = q4[2 * cc + 1];
            const __half2* hp = reinterpret_cast<const __half2*>(&kv);
            const float2 k0 = __half22float2(hp[0]), k1 = __half22float2(hp[1]), k2 = __half22float2(hp[2]), k3 = __half22float2(hp[3]);
            sc = fmaf(qa.x, k0.x, fmaf(qa.y, k0.y, sc));
            sc = fmaf(qa.z, k1.x, fmaf(qa.w, k1.y, sc));
            sc = fmaf(qb.x, k2.x, fmaf(qb.y, k2.y, sc));
            sc = fmaf(qb.z, k3.x, fmaf(qb.w, k3.y, sc));
          }
          const bool keep = (mw >> lane) & 1u;
          const float tt = key < Lk ? (keep ? sc * c1 : t_masked) : -CUDART_INF_F;
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
          o0 *= alpha;
          o1 *= alpha;
          m_run = m_new;
          const float pr = __half2float(__float2half_rn(e));   // P rounded to f16 before P V, like the tensor-core path
          // ---- P V: lane = two output dims
          const uint32_t vs = dc_acquire(ctx, st);
#pragma unroll
          for (int uu = 0; uu < 32; ++uu) {
            const uint32_t vv = dc_lds32(vs + uu * 128 + lane * 4);
            const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&vv));
            const float pk = __shfl_sync(0xffffffffu, pr, uu);
            o0 = fmaf(pk, vf.x, o0);
            o1 = fmaf(pk, vf.y, o1);
          }
        }
        if (gcur >= 0) {
          if (lane == 0) pm[warp * DC_G + gcur] = m_run, pl[warp * DC_G + gcur] = l_run;
          *reinterpret_cast<float2*>(po + (warp * DC_G + gcur) * DC_DK + 2 * lane) = make_float2(o0, o1);
        }
      }
      __syncthreads();
      DC_STAMP(3);
      // ---- merge the warps' partial (max, sum, O) of row `warp` in a fixed order; self-attention adds the new key
      if (warp < nrows) {
        const int gr = warp;
        float m = -CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < DC_WARPS; ++w) m = fmaxf(m, pm[w * DC_G + gr]);
        float m_new = -CUDART_INF_F;
        float2 vnew = make_float2(0.f, 0.f);
        if (d.kind == 0) {
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
        for (int w = 0; w < DC_WARPS; ++w) {
          const float pmw = pm[w * DC_G + gr];
          if (pmw > -CUDART_INF_F) {
            const float f = dc_ex2(pmw - m);
            const float2 ov = *reinterpret_cast<const float2*>(po + (w * DC_G + gr) * DC_DK + 2 * lane);
            l = fmaf(pl[w * DC_G + gr], f, l);
            a0 = fmaf(ov.x, f, a0);
            a1 = fmaf(ov.y, f, a1);
          }
        }
        if (d.kind == 0) {
          const float f = dc_ex2(m_new - m);
          l += f;
          a0 = fmaf(vnew.x, f, a0);
          a1 = fmaf(vnew.y, f, a1);
        }
        const float inv = 1.f / l;
        const uint32_t h2 = pack_f16x2_sat(a0 * inv, a1 * inv);
        const uint32_t a = ob_s + gr * DC_LDA + ((int)rank * DC_DK + 2 * lane) * 2;
#pragma unroll
        for (uint32_t r = 0; r < DC_CS; ++r) dc_st_cluster_u32(dc_mapa(a, r), h2);
      }
      DC_STAMP(4);
      dc_cluster_arrive();
      bo = __ldg(reinterpret_cast<const float2*>(d.b_out + col));
      dc_cluster_wait();
    }
    DC_STAMP(5);
    // ---- output projection (attention: A = all heads' outputs; feed-forward: A = hidden activation) + residual
    {
      const float2 y = d.kind == 2 ? dc_proj<8>(ctx, st, sbase, DC_OFF_HID, DC_LDH) : dc_proj<2>(ctx, st, sbase, DC_OFF_OB, DC_LDA);
      if (g < nrows) {
        const float2 xo = *reinterpret_cast<const float2*>(xs + g * DC_D + col);
        const float x0 = xo.x + (y.x + bo.x), x1 = xo.y + (y.y + bo.y);
        const uint32_t a = xs_s + (g * DC_D + col) * 4;
#pragma unroll
        for (uint32_t r = 0; r < DC_CS; ++r) dc_st_cluster_v2f32(dc_mapa(a, r), x0, x1);
      }
    }
    DC_STAMP(6);
    dc_cluster_arrive();
    if (s + 1 < p.n_sites) dc_ln_load(ln, sites[s + 1].ln_a, sites[s + 1].ln_b, sites[s + 1].ln_eps, lane);
    else dc_ln_load(ln, p.norm_a, p.norm_b, p.norm_eps, lane);
    dc_cluster_wait();
    DC_STAMP(7);
    if (p.taps != nullptr && rank == 0 && warp < nrows) {
      float4* tp = reinterpret_cast<float4*>(p.taps + ((size_t)s * p.B + row0 + warp) * DC_D);
#pragma unroll
      for (int i = 0; i < 4; ++i) tp[lane + 32 * i] = reinterpret_cast<const float4*>(xs + warp * DC_D)[lane + 32 * i];
    }
  }
  // ---- final LayerNorm (mtn.py:164) -> out
  if (rank == 0 && warp < nrows) {
    float4 o[4];
    dc_ln_row(xs + warp * DC_D, ln, lane, o);
    float4* op = reinterpret_cast<float4*>(p.out + (size_t)(row0 + warp) * DC_D);
#pragma unroll
    for (int i = 0; i < 4; ++i) op[lane + 32 * i] = o[i];
  }
  dc_cp_wait<0>();
  dc_cluster_arrive();   // no CTA leaves while a peer could still address its shared memory
  dc_cluster_wait();
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

extern "C" int mtn_decode_cluster_supported(int B, int d, int h, int d_ff) {
  return (B >= 1 && B <= 16 * mtn::DC_G && d == mtn::DC_D && h == mtn::DC_CS && d_ff == mtn::DC_DFF) ? 1 : 0;
}

extern "C" int mtn_decode_cluster_max_sites(void) { return mtn::DC_MAX_SITES; }

extern "C" int mtn_decode_cluster_fwd(const MtnDecodeClusterArgs* a, void* stream) {
  using namespace mtn;
  MTN_REQUIRE(a && a->sites && a->x_in && a->out && a->norm_a && a->norm_b, MTN_E_ARG, "decode_cluster: NULL pointer");
  MTN_REQUIRE(mtn_decode_cluster_supported(a->B, a->d, a->h, a->d_ff), MTN_E_SHAPE,
              "decode_cluster: B=%d d=%d h=%d d_ff=%d (d = 512, h = 8, d_ff = 2048, B <= 128)", a->B, a->d, a->h, a->d_ff);
  MTN_REQUIRE(a->n_sites >= 1 && a->n_sites <= DC_MAX_SITES && a->t >= 0, MTN_E_SHAPE, "decode_cluster: n_sites=%d (<= %d), t=%d", a->n_sites,
              DC_MAX_SITES, a->t);
  MTN_REQUIRE(aligned16(a->x_in) && aligned16(a->out) && aligned16(a->norm_a) && aligned16(a->norm_b) && (!a->taps || aligned16(a->taps)),
              MTN_E_ALIGN, "decode_cluster: x_in / out / norm / taps must be 16-byte aligned");
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
  DcParams p{a->n_sites, a->B, G, a->t, a->x_in, a->out, a->norm_a, a->norm_b, a->norm_eps, a->taps, a->stamps};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MTN_CHECK_CUDA(launch_kernel_cluster(decode_cluster_kernel, dim3(nclusters * DC_CS), dim3(DC_THREADS), DC_SMEM, st, DC_CS, tab, p));
  return MTN_OK;
}
