// Probe 2: M = 64 tcgen05.mma.kind::tf32 with the A operand in tensor memory (TS mode): which TMEM lane feeds
// accumulator row m, for accumulator / operand lane offsets 0 and 16; and the pace of a chain of such MMAs at N = 192.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nemo_cvpr2023_b200/csrc -o mma_m64_ts tools/ubench/mma_m64_ts.cu
// A[lane L][k = 0] = L + 1 in every one of the 128 lanes (columns 64..71), B[n = 0][k = 0] = 1 -> D[row][0] = source lane + 1.
#include <cstdio>
#include <cstdlib>
#include "sm100.cuh"
using namespace nb::sm100;

__global__ void __launch_bounds__(160, 1) k(int dlane, int alane, int pace_iters, float* out, unsigned long long* cyc) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + 100 * 1024, slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<512>(slot);
  for (int i = threadIdx.x; i < 100 * 1024 / 4; i += 160) sts32(sbase + 4 * i, 0.f);
  __syncthreads();
  if (threadIdx.x == 0) sts32(sbase, 1.f);  // B: row n = 0, k = 0
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = lds32u(slot);
  if (warp < 4) {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t row = tm + ((uint32_t)(warp * 32) << 16);
    tmem_st32(row, z);        // D columns 0..31
    z[0] = __float_as_uint((float)(threadIdx.x + 1));
    tmem_st32(row + 64, z);   // A columns 64..95: k = 0 carries lane + 1
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = umma_idesc_tf32(64, 16, 0, 0);
    const uint64_t b = umma_smem_desc(sbase, 16, 1024, LAYOUT_SW128);
    umma_tf32_ts(tm + ((uint32_t)dlane << 16), tm + ((uint32_t)alane << 16) + 64, b, idesc, 0u);
    umma_commit(bars);
    mbar_wait(bars, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    uint32_t t[4];
    tmem_ld4(tm + ((uint32_t)(warp * 32) << 16), t);
    tmem_ld_wait();
    out[threadIdx.x] = __uint_as_float(t[0]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0 && pace_iters) {  // pace: M = 64, N = 192, alternating accumulators at lane offsets 0 / 16
    const uint32_t idesc = umma_idesc_tf32(64, 192, 0, 0);
    const uint64_t b = umma_smem_desc(sbase, 16, 1024, LAYOUT_SW128);
    const long long t0 = clock64();
    for (int it = 0; it < pace_iters; ++it)
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        umma_tf32_ts(tm + 128 + ((uint32_t)((it & 1) * dlane) << 16), tm + ((uint32_t)((it & 1) * alane) << 16) + 64 + (ks & 3) * 8, umma_desc_advance(b, (ks & 3) * 32), idesc, 1u);
    umma_commit(bars + 8);
    mbar_wait(bars + 8, 0);
    cyc[0] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main(int argc, char** argv) {  // mma_m64_ts <accumulator lane offset> <A lane offset> <pace iterations>
  const int c[3] = {argc > 1 ? atoi(argv[1]) : 0, argc > 2 ? atoi(argv[2]) : 0, argc > 3 ? atoi(argv[3]) : 0};
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 128 * 4);
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 102 * 1024);
  cudaMemset(out, 0, 512);
  cudaMemset(cyc, 0, 8);
  k<<<1, 160, 102 * 1024>>>(c[0], c[1], c[2], out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("TS mode, M = 64: accumulator lane offset %d, A operand lane offset %d: %s\n", c[0], c[1],
         e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  if (e != cudaSuccess) return 0;
  float h[128];
  unsigned long long hc;
  cudaMemcpy(h, out, 512, cudaMemcpyDeviceToHost);
  cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  for (int l = 0; l < 128; ++l) printf("%s%4.0f", l % 32 == 0 ? "\n  D lanes " : "", h[l]);
  printf("\n");
  if (c[2]) printf("  pace M=64 N=192 TS, accumulators alternating between lane offsets 0 / %d: %.1f cycles/MMA\n", c[0], hc / (c[2] * 16.0));
  return 0;
}
