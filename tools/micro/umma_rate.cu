// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, kind::f16) as a function of N, operands in shared memory (SS)
// or A in tensor memory (TS), issued back to back by one thread / with a commit + mbarrier round per group of 4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mtn_b200/csrc -o gpurun_out/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include "common.cuh"
using namespace mtn;

template <int N, bool TS, int GROUP>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bars = base + 16384 + 32768;
  const uint32_t slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); mbar_fence_init(); }
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem_raw + (base - raw) + 16384 + 32768 + 64);
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(128, N, 0, 0);
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    if (lane == 0) {
      const uint64_t da = make_smem_desc(sA, 16, 1024, SWZ_128B), db = make_smem_desc(sB, 16, 1024, SWZ_128B);
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (TS) tc_mma_f16_ts(tm, tm + 256 + (r & 3) * 8, db + 2 * (r & 3), idesc, 1);
        else tc_mma_f16(tm, da + 2 * (r & 3), db + 2 * (r & 3), idesc, 1);
        if (GROUP > 0 && (r % (GROUP > 0 ? GROUP : 1)) == GROUP - 1) {   // commit + wait, like a ring stage
          tc_commit(bars + 8);
          if (GROUP >= 1000) { mbar_wait(bars + 8, ph); ph ^= 1; }
        }
      }
      tc_commit(bars);
      mbar_wait(bars, 0);
      t1 = clock64();
      out[0] = t1 - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, bool TS, int GROUP>
static void run(const char* what, long long* d_out) {
  const int reps = 256;
  cudaFuncSetAttribute(rate_kernel<N, TS, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  long long h = 0;
  for (int i = 0; i < 3; ++i) {
    rate_kernel<N, TS, GROUP><<<1, 64, 60000>>>(d_out, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d: %6.1f cycles per MMA (%lld cycles / %d)\n", what, N, (double)h / reps, h, reps);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<64, false, 0>("SS back to back", d_out);
  run<96, false, 0>("SS back to back", d_out);
  run<128, false, 0>("SS back to back", d_out);
  run<256, false, 0>("SS back to back", d_out);
  run<64, true, 0>("TS (A in TMEM) back to back", d_out);
  run<128, true, 0>("TS (A in TMEM) back to back", d_out);
  run<256, true, 0>("TS (A in TMEM) back to back", d_out);
  run<128, false, 4>("SS, commit every 4 MMAs", d_out);
  run<64, false, 4>("SS, commit every 4 MMAs", d_out);
  return 0;
}
