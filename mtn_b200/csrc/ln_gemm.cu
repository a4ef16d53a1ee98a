// LayerNorm fused into the projection that consumes it:   Y = act( LN(x) W^T + bias )   (f16 out)
//
// Every sublayer of the MTN hot path starts with SublayerConnection's norm (mtn.py:126) followed by an
// nn.Linear of the normalised rows (mtn.py:256 Q projection / packed Q|K|V, :280 w_1).  As two kernels
// that is a pass over the f32 residual stream that writes an f16 copy, a launch boundary, and a GEMM that
// reads the copy back.  On the latency-bound chains of the decoder (one dependent kernel after another on
// a [B*T, d] stream) the boundary costs more than the arithmetic, so this kernel does both:
//
//   * a CTA owns 128 rows.  Its 8 worker warps normalise them (one warp per row, the row in registers,
//     two-pass unbiased variance -- the arithmetic of layernorm_rows_kernel; at d = 256 / 512 the results
//     are bit-identical to it, at d = 128 the compiler contracts one FMA differently) and write the
//     f16 result straight into shared memory as the UMMA A operand: d/64 K-major panels of
//     [128 rows x 128 B] in the 128-byte swizzle, i.e. exactly what TMA would have produced;
//   * the W tiles ([128 out-features x 64 k] f16, 16 KB) stream through a 5-stage TMA ring; one thread
//     issues 128x128x16 tcgen05.mma into one of two TMEM accumulators; the A panels are reused by every
//     n-tile of the CTA;
//   * the worker warps then turn into the epilogue (bias, ReLU, f16, coalesced stores), overlapping the
//     MMAs of the next n-tile.
//
// Grid = (row blocks, nsplit): the n-tiles of a row block are dealt round-robin to nsplit CTAs (each
// repeats the cheap normalisation) so that small problems still spread over the machine.
// d in {128, 256, 512}: the A panels of a row block (d * 256 B) must fit next to the ring.
//
// STATUS (profiles/r01d_ln_linear.txt): correct on every tested shape, but 3.5-5 us SLOWER per sublayer than
// the two launches it replaces (with programmatic dependent launch the boundary costs ~3 us, while here the
// normalisation of a row block runs on 8 warps of ONE SM -- 16 dependent rows per warp -- before the first MMA
// can issue; the standalone LayerNorm spreads the same rows over 64 warps on each of 148 SMs).  The engine
// therefore keeps the two-launch form (MTN_B200_LN_FUSED=<max rows> opts in); the kernel is the building
// block of the row-block-resident site fusion planned next.
#include "common.cuh"
#include "host.h"

namespace mtn {

constexpr int LG_BM = 128, LG_BN = 128, LG_BK = 64;
constexpr int LG_THREADS = 320;  // TMA warp, MMA warp, 8 worker (LayerNorm, then epilogue) warps
constexpr int LG_STAGES = 5;

template <int VPL>
struct LnGemmSmem {
  static constexpr int D = 128 * VPL;
  static constexpr int NKB = D / LG_BK;
  static constexpr int A_KB_BYTES = LG_BM * LG_BK * 2;  // one k-block panel: 128 rows x 128 B
  static constexpr int B_BYTES = LG_BN * LG_BK * 2;
  static constexpr int RING_OFF = NKB * A_KB_BYTES;
  static constexpr int XPOSE_OFF = RING_OFF + LG_STAGES * B_BYTES;  // 8 x [32 rows x 64 B] f16 transpose tiles
  static constexpr int BAR_OFF = XPOSE_OFF + 8 * 2048;
  static constexpr int NBARS = 2 * LG_STAGES + 5;
  static constexpr int TOTAL = BAR_OFF + 8 * NBARS + 16 + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget");
};

template <int VPL>
__global__ void __launch_bounds__(LG_THREADS, 1)
    ln_gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ x,
                          const float* __restrict__ a2, const float* __restrict__ b2, float eps,
                          const float* __restrict__ bias, int act, __half* __restrict__ out16, int ld16, int M, int N,
                          int tiles_n, long long* __restrict__ ts) {
  using L = LnGemmSmem<VPL>;
  constexpr int D = L::D, NKB = L::NKB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + L::BAR_OFF;
  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (LG_STAGES + s); };
  auto bar_acc_full = [&](int b) { return bar_base + 8u * (2 * LG_STAGES + b); };
  auto bar_acc_empty = [&](int b) { return bar_base + 8u * (2 * LG_STAGES + 2 + b); };
  const uint32_t bar_a_ready = bar_base + 8u * (2 * LG_STAGES + 4);
  const uint32_t tmem_slot = bar_base + 8u * L::NBARS;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * L::NBARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * LG_BM;
  const int first_tile = blockIdx.y, tile_stride = gridDim.y;
  constexpr uint32_t TMEM_COLS = 2u * LG_BN;

  // phase timestamps of CTA (0, 0) for tools/ln_linear_bench.py --phases (ts == NULL in production)
  const bool stamp = ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  if (stamp && warp == 0) ts[0] = clock64();
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < LG_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), 8);
    }
    mbar_init(bar_a_ready, 8);  // one arrive per worker warp
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  if (stamp && warp == 0) ts[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (W tiles)
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = first_tile; t < tiles_n; t += tile_stride) {
        for (int kb = 0; kb < NKB; ++kb, ++it) {
          const int s = it % LG_STAGES;
          mbar_wait(bar_empty(s), ((it / LG_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full(s), L::B_BYTES);
          tma_load_2d(base + L::RING_OFF + s * L::B_BYTES, &tmW, bar_full(s), kb * LG_BK, t * LG_BN);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(LG_BM, LG_BN, 0, 0);
    mbar_wait(bar_a_ready, 0);  // the normalised rows are in shared memory (generic writes fenced by the workers)
    tc_fence_after();
    if (stamp) ts[3] = clock64();
    uint32_t it = 0, lt = 0;
    for (int t = first_tile; t < tiles_n; t += tile_stride, ++lt) {
      const uint32_t buf = lt & 1u;
      mbar_wait(bar_acc_empty(buf), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * LG_BN;
      for (int kb = 0; kb < NKB; ++kb, ++it) {
        const int s = it % LG_STAGES;
        mbar_wait(bar_full(s), (it / LG_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc(base + kb * L::A_KB_BYTES, 16, 1024, SWZ_128B);
          const uint64_t db = make_smem_desc(base + L::RING_OFF + s * L::B_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < LG_BK / 16; ++k) tc_mma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          tc_commit(bar_empty(s));
          if (kb == NKB - 1) tc_commit(bar_acc_full(buf));
          if (stamp && kb == NKB - 1 && lt == 0) ts[4] = clock64();
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ workers: LayerNorm of the row block
    const int ew = warp - 2;
    {
      constexpr int RB = 4;               // rows in flight per warp (loads of the next batch are issued
      constexpr int NBATCH = 16 / RB;     // before the arithmetic of the current one)
      float4 v[2][RB][VPL];
      auto load_batch = [&](int b, float4(&dst)[RB][VPL]) {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int g = m0 + ew + 8 * (b * RB + j);
          const float4* xr = reinterpret_cast<const float4*>(x + (size_t)(g < M ? g : 0) * D);
#pragma unroll
          for (int i = 0; i < VPL; ++i) dst[j][i] = (g < M) ? __ldcg(xr + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_batch(0, v[0]);
#pragma unroll
      for (int b = 0; b < NBATCH; ++b) {
        if (b + 1 < NBATCH) load_batch(b + 1, v[(b + 1) & 1]);
        float4(&w)[RB][VPL] = v[b & 1];
        // the RB rows of a batch advance in lockstep: RB independent shuffle / arithmetic chains per step
        float s[RB], ss[RB], inv[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          s[j] = 0.f;
#pragma unroll
          for (int i = 0; i < VPL; ++i) s[j] += (w[j][i].x + w[j][i].y) + (w[j][i].z + w[j][i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < RB; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const float mean = s[j] * (1.f / D);
          ss[j] = 0.f;
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            w[j][i].x -= mean; w[j][i].y -= mean; w[j][i].z -= mean; w[j][i].w -= mean;
            ss[j] += (w[j][i].x * w[j][i].x + w[j][i].y * w[j][i].y) + (w[j][i].z * w[j][i].z + w[j][i].w * w[j][i].w);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < RB; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
#pragma unroll
        for (int j = 0; j < RB; ++j) inv[j] = 1.f / (sqrtf(ss[j] * (1.f / (D - 1))) + eps);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int c4 = lane + 32 * i;  // columns 4 c4 .. 4 c4 + 3
          const float4 a = __ldg(reinterpret_cast<const float4*>(a2) + c4);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b2) + c4);
          // k-block c4 / 16, 16-byte chunk (c4 % 16) / 2 of the 128-byte row, XOR-swizzled with row % 8
          const uint32_t kb = (uint32_t)c4 >> 4, chunk = ((uint32_t)c4 & 15u) >> 1;
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            const int r = ew + 8 * (b * RB + j);
            float4 o;
            o.x = a.x * w[j][i].x * inv[j] + bb.x;
            o.y = a.y * w[j][i].y * inv[j] + bb.y;
            o.z = a.z * w[j][i].z * inv[j] + bb.z;
            o.w = a.w * w[j][i].w * inv[j] + bb.w;
            uint2 pk = make_uint2(pack_f16x2_sat(o.x, o.y), pack_f16x2_sat(o.z, o.w));
            if (m0 + r >= M) pk = make_uint2(0u, 0u);  // rows past M: zero operand rows (their outputs are never stored)
            const uint32_t off = kb * L::A_KB_BYTES + (uint32_t)r * 128u + ((chunk ^ ((uint32_t)r & 7u)) << 4) +
                                 ((uint32_t)c4 & 1u) * 8u;
            *reinterpret_cast<uint2*>(smem + off) = pk;
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (stamp && ew == 0) ts[2] = clock64();
      if (lane == 0) mbar_arrive(bar_a_ready);
    }
    // ------------------------------------------------------------ workers: epilogue
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int half = ew >> 2;  // 0: chunks 0, 2   1: chunks 1, 3
    uint4* hp = reinterpret_cast<uint4*>(smem + L::XPOSE_OFF + ew * 2048);
    const int sub_r = lane >> 2, c8 = lane & 3;
    constexpr int NCHUNK = LG_BN / 32;
    uint32_t lt = 0;
    for (int t = first_tile; t < tiles_n; t += tile_stride, ++lt) {
      const int n0 = t * LG_BN;
      const uint32_t buf = lt & 1u;
      const int row0 = m0 + q * 32 + sub_r;
      size_t off16[4];
      bool row_ok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        row_ok[i] = row0 + i * 8 < M;
        off16[i] = (size_t)(row0 + i * 8) * ld16;
      }
      mbar_wait(bar_acc_full(buf), (lt >> 1) & 1);
      tc_fence_after();
      if (stamp && ew == 0 && lt == 0) ts[5] = clock64();
      const uint32_t t_acc = tmem_base + buf * LG_BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        const int cb = n0 + c * 32;
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias != nullptr && cb + 4 * j < N) bb[j] = __ldg(reinterpret_cast<const float4*>(bias + cb + 4 * j));
        }
        uint32_t acc[32];
        tc_ld32(t_acc + c * 32, acc);
        tc_wait_ld();
        if (c + 2 >= NCHUNK) {  // this warp's last read of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty(buf));
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v0 = __uint_as_float(acc[4 * j]) + bb[j].x, v1 = __uint_as_float(acc[4 * j + 1]) + bb[j].y;
          float v2 = __uint_as_float(acc[4 * j + 2]) + bb[j].z, v3 = __uint_as_float(acc[4 * j + 3]) + bb[j].w;
          if (act == MTN_ACT_RELU) {
            v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
          }
          pk[2 * j] = pack_f16x2_sat(v0, v1);
          pk[2 * j + 1] = pack_f16x2_sat(v2, v3);
        }
        // [32 rows x 64 B] tile, 16-B slots XOR-swizzled by (row >> 1) & 3 (as in gemm.cu's f16 epilogue)
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
          if (row_ok[i] && colh < N) *reinterpret_cast<uint4*>(out16 + off16[i] + colh) = hv[i];
      }
    }
    if (stamp && ew == 0) ts[6] = clock64();
    tc_fence_before();
  }
  __syncthreads();
  if (stamp && warp == 0) ts[7] = clock64();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static long long* g_ln_gemm_ts = nullptr;  // tools only: device buffer of 8 phase timestamps

template <int VPL>
static int launch_ln_gemm(const float* x, const float* a2, const float* b2, float eps, int M, const void* W, int ldw,
                          const float* bias, int N, int act, void* out16, int ld16, int num_sms, cudaStream_t st) {
  using L = LnGemmSmem<VPL>;
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(ln_gemm_f16_tc_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  CUtensorMap tmW;
  int rc = make_tmap_2d_f16(&tmW, W, (uint64_t)L::D, (uint64_t)N, (uint64_t)ldw, LG_BK, LG_BN, TM_SWZ_128);
  if (rc) return rc;
  const int tiles_m = (M + LG_BM - 1) / LG_BM, tiles_n = (N + LG_BN - 1) / LG_BN;
  int nsplit = num_sms / tiles_m;  // one wave of CTAs
  if (nsplit < 1) nsplit = 1;
  if (nsplit > tiles_n) nsplit = tiles_n;
  MTN_CHECK_CUDA(launch_kernel(ln_gemm_f16_tc_kernel<VPL>, dim3(tiles_m, nsplit), dim3(LG_THREADS), L::TOTAL, st, tmW, x, a2,
                               b2, eps, bias, act, reinterpret_cast<__half*>(out16), ld16, M, N, tiles_n, g_ln_gemm_ts));
  return MTN_OK;
}

}  // namespace mtn

extern "C" int mtn_ln_linear_debug_timestamps(void* dev_buf8) {  // tools only; NULL switches the stamps off
  mtn::g_ln_gemm_ts = static_cast<long long*>(dev_buf8);
  return MTN_OK;
}

extern "C" int mtn_ln_linear_supported(int d) { return d == 128 || d == 256 || d == 512; }

extern "C" int mtn_ln_linear_fwd(const float* x, const float* a_2, const float* b_2, float eps, int M, int d,
                                 const void* W, int ldw, const float* bias, int N, int act, void* out_f16, int ld16,
                                 void* stream) {
  using namespace mtn;
  MTN_REQUIRE(x && a_2 && b_2 && W && out_f16, MTN_E_ARG, "ln_linear: NULL pointer");
  MTN_REQUIRE(M > 0 && N > 0 && N % 8 == 0, MTN_E_SHAPE, "ln_linear: M=%d N=%d (N must be a multiple of 8)", M, N);
  MTN_REQUIRE(mtn_ln_linear_supported(d), MTN_E_SHAPE,
              "ln_linear: d=%d (the row block's operand panels fit shared memory for d in {128, 256, 512}; use "
              "mtn_layernorm_fwd + mtn_linear_fwd otherwise)", d);
  MTN_REQUIRE(act == MTN_ACT_NONE || act == MTN_ACT_RELU, MTN_E_ARG, "ln_linear: act=%d", act);
  MTN_REQUIRE(ldw >= d && ldw % 8 == 0 && ld16 >= N && ld16 % 8 == 0, MTN_E_ALIGN, "ln_linear: ldw=%d ld16=%d", ldw, ld16);
  MTN_REQUIRE(aligned16(x) && aligned16(a_2) && aligned16(b_2) && aligned16(W) && aligned16(out_f16) &&
                  (bias == nullptr || aligned16(bias)),
              MTN_E_ALIGN, "ln_linear: operands must be 16-byte aligned");
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0, n = 0;
    MTN_CHECK_CUDA(cudaGetDevice(&dev));
    MTN_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    num_sms = n;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d == 128) return launch_ln_gemm<1>(x, a_2, b_2, eps, M, W, ldw, bias, N, act, out_f16, ld16, num_sms, st);
  if (d == 256) return launch_ln_gemm<2>(x, a_2, b_2, eps, M, W, ldw, bias, N, act, out_f16, ld16, num_sms, st);
  return launch_ln_gemm<4>(x, a_2, b_2, eps, M, W, ldw, bias, N, act, out_f16, ld16, num_sms, st);
}
