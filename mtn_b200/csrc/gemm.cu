// Fused linear layer on tcgen05:  C = act(A W^T + bias) + addend
//
// Replaces the nn.Linear call sites of the MTN hot path (mtn.py:256-258 Q/K/V
// projections, :267 output projection, :280 FFN w_1/w_2, :35 video encoder).
// A [M,K] and W [N,K] are both K-major f16, so the GEMM is the "TN" form UMMA
// consumes directly: 128 x BN x 64 tiles are staged by TMA (128-byte swizzle) into
// a STAGES-deep shared-memory ring, one elected thread issues 128xBNx16
// tcgen05.mma (f32 accumulate in tensor memory) and four epilogue warps read the
// accumulator back with tcgen05.ld (one thread per output row) and apply bias /
// ReLU / residual-or-positional addend before storing f32 and/or f16.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31).
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
};

constexpr int BM = 128;
constexpr int BK = 64;  // 64 f16 = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;  // barriers + 1 KB alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
    gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmEpi epi, int M, int N, int K) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle-128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar_base = base + L::BAR_OFF;
  const uint32_t bar_tmem_full = bar_base + 8 * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8 * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int nkb = (K + BK - 1) / BK;  // K tail: TMA zero-fills columns >= K

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);             // full[s]:  producer arrive + tx bytes
      mbar_init(bar_base + 8 * (STAGES + s), 1);  // empty[s]: tcgen05.commit
    }
    mbar_init(bar_tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_base + 8 * (STAGES + s), ph ^ 1);
        const uint32_t full = bar_base + 8 * s;
        mbar_arrive_expect_tx(full, L::STAGE_BYTES);
        const uint32_t sA = base + s * L::STAGE_BYTES;
        tma_load_2d(sA, &tmA, full, kb * BK, m0);
        tma_load_2d(sA + L::A_BYTES, &tmB, full, kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(BM, BN, 0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(bar_base + 8 * s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sA = base + s * L::STAGE_BYTES;
        const uint64_t da = make_smem_desc(sA, 16, 1024, SWZ_128B);
        const uint64_t db = make_smem_desc(sA + L::A_BYTES, 16, 1024, SWZ_128B);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // +32 B along K inside the swizzled row = +2 encoded
          tc_mma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        tc_commit(bar_base + 8 * (STAGES + s));  // frees the stage when these MMAs retire
        if (kb == nkb - 1) tc_commit(bar_tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(bar_tmem_full, 0);
    tc_fence_after();
    const bool row_ok = row < M;
    const float* add_row = nullptr;
    if (epi.addend != nullptr && row_ok)
      add_row = epi.addend + (size_t)(epi.add_period > 0 ? row % epi.add_period : row) * epi.ld_add;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      tc_wait_ld();
      const int col = n0 + c * 32;
      const int nv = N - col;  // valid columns of this chunk (multiple of 8); >= 32 except in the N tail
      if (row_ok && nv > 0) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (epi.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j >= nv) break;
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (epi.act == MTN_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (add_row != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j >= nv) break;
            const float4 a4 = *reinterpret_cast<const float4*>(add_row + col + j);
            v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
          }
        }
        if (epi.out32 != nullptr) {
          float4* o = reinterpret_cast<float4*>(epi.out32 + (size_t)row * epi.ld32 + col);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nv) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (epi.out16 != nullptr) {
          uint4* o = reinterpret_cast<uint4*>(epi.out16 + (size_t)row * epi.ld16 + col);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (8 * j < nv) o[j] = make_uint4(pack_f16x2_sat(v[8 * j], v[8 * j + 1]), pack_f16x2_sat(v[8 * j + 2], v[8 * j + 3]),
                              pack_f16x2_sat(v[8 * j + 4], v[8 * j + 5]), pack_f16x2_sat(v[8 * j + 6], v[8 * j + 7]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, int STAGES>
static int launch_gemm(const MtnLinearArgs& a, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    MTN_CHECK_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<BN, STAGES>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_f16(&tmA, a.A, a.K, a.M, a.lda, BK, BM, TM_SWZ_128);
  if (rc) return rc;
  rc = make_tmap_2d_f16(&tmB, a.W, a.K, a.N, a.ldw, BK, BN, TM_SWZ_128);
  if (rc) return rc;
  GemmEpi epi{a.bias, a.act, a.addend, a.ld_add, a.add_period, a.out_f32, a.ld32,
              reinterpret_cast<__half*>(a.out_f16), a.ld16};
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
  gemm_f16_tc_kernel<BN, STAGES><<<grid, GEMM_THREADS, L::TOTAL, st>>>(tmA, tmB, epi, a.M, a.N, a.K);
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}

static int validate_linear(const MtnLinearArgs* a) {
  MTN_REQUIRE(a != nullptr && a->A != nullptr && a->W != nullptr, MTN_E_ARG, "linear: NULL operand");
  MTN_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, MTN_E_SHAPE, "linear: M=%d N=%d K=%d", a->M, a->N, a->K);
  MTN_REQUIRE(a->K % 8 == 0, MTN_E_SHAPE, "linear: K=%d must be a multiple of 8", a->K);
  MTN_REQUIRE(a->N % 8 == 0, MTN_E_SHAPE, "linear: N=%d must be a multiple of 8", a->N);
  MTN_REQUIRE(a->lda >= a->K && a->ldw >= a->K && a->lda % 8 == 0 && a->ldw % 8 == 0, MTN_E_ALIGN,
              "linear: lda=%d ldw=%d must be >= K and multiples of 8", a->lda, a->ldw);
  MTN_REQUIRE(aligned16(a->A) && aligned16(a->W), MTN_E_ALIGN, "linear: A/W not 16-byte aligned");
  MTN_REQUIRE(a->out_f32 != nullptr || a->out_f16 != nullptr, MTN_E_ARG, "linear: no output");
  MTN_REQUIRE(a->act == MTN_ACT_NONE || a->act == MTN_ACT_RELU, MTN_E_ARG, "linear: act=%d", a->act);
  if (a->out_f32)
    MTN_REQUIRE(aligned16(a->out_f32) && a->ld32 % 4 == 0 && a->ld32 >= a->N, MTN_E_ALIGN,
                "linear: out_f32 alignment / ld32=%d", a->ld32);
  if (a->out_f16)
    MTN_REQUIRE(aligned16(a->out_f16) && a->ld16 % 8 == 0 && a->ld16 >= a->N, MTN_E_ALIGN,
                "linear: out_f16 alignment / ld16=%d", a->ld16);
  if (a->addend)
    MTN_REQUIRE(aligned16(a->addend) && a->ld_add % 4 == 0 && a->ld_add >= a->N && a->add_period >= 0,
                MTN_E_ALIGN, "linear: addend alignment / ld_add=%d", a->ld_add);
  if (a->bias) MTN_REQUIRE(aligned16(a->bias), MTN_E_ALIGN, "linear: bias not 16-byte aligned");
  return MTN_OK;
}

// ----------------------------------------------------------------------------
// self-check kernel (tests only): one thread per output element, same arithmetic
// contract (f16 operands, f32 accumulate, same epilogue order).
// ----------------------------------------------------------------------------
__global__ void gemm_f16_check_kernel(const __half* A, int lda, const __half* W, int ldw, GemmEpi epi,
                                      int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc = fmaf(__half2float(A[(size_t)m * lda + k]), __half2float(W[(size_t)n * ldw + k]), acc);
  if (epi.bias) acc += epi.bias[n];
  if (epi.act == MTN_ACT_RELU) acc = fmaxf(acc, 0.f);
  if (epi.addend) acc += epi.addend[(size_t)(epi.add_period > 0 ? m % epi.add_period : m) * epi.ld_add + n];
  if (epi.out32) epi.out32[(size_t)m * epi.ld32 + n] = acc;
  if (epi.out16) {
    const uint32_t p = pack_f16x2_sat(acc, 0.f);
    epi.out16[(size_t)m * epi.ld16 + n] = __ushort_as_half((unsigned short)(p & 0xffff));
  }
}

}  // namespace mtn

extern "C" int mtn_linear_fwd(const MtnLinearArgs* a, void* stream) {
  int rc = mtn::validate_linear(a);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->N > 64) return mtn::launch_gemm<128, 3>(*a, st);
  return mtn::launch_gemm<64, 4>(*a, st);
}

extern "C" int mtn_check_linear_fwd(const MtnLinearArgs* a, void* stream) {
  int rc = mtn::validate_linear(a);
  if (rc) return rc;
  mtn::GemmEpi epi{a->bias, a->act, a->addend, a->ld_add, a->add_period, a->out_f32, a->ld32,
                   reinterpret_cast<__half*>(a->out_f16), a->ld16};
  dim3 grid((a->N + 127) / 128, a->M);
  mtn::gemm_f16_check_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(a->A), a->lda, reinterpret_cast<const __half*>(a->W), a->ldw, epi,
      a->M, a->N, a->K);
  MTN_CHECK_CUDA(cudaGetLastError());
  return MTN_OK;
}
