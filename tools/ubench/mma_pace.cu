// Micro-benchmark: pace of tcgen05.mma.kind::tf32 (M = 128, K = 8) for small N, as used inside the skinning kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nemo_cvpr2023_b200/csrc -o mma_pace tools/ubench/mma_pace.cu
// Variants: A from TMEM (TS) or shared memory (SS); B K-major (128-byte swizzle) or MN-major; one accumulator or two
// alternating; with / without four warps concurrently reading TMEM (tcgen05.ld) or writing shared memory.
#include <cstdio>
#include <cstdlib>
#include "sm100.cuh"
using namespace nb::sm100;

struct P { int N, ts, mn, alt, ldwarps, stswarps, iters, ldcols; };

__global__ void __launch_bounds__(320, 1) k(P p, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + 200 * 1024;
  const uint32_t slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bars, 1); fence_barrier_init(); }
  if (warp == 8) tmem_alloc<512>(slot);
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 320) sts128(sbase + 16 * i, make_float4(1.f, 1.f, 1.f, 1.f));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = lds32u(slot);
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, p.N, 0, p.mn);
      const uint64_t b0 = p.mn ? umma_smem_desc(sbase, 128 * 128, 512, LAYOUT_SW128_BASE32B)
                               : umma_smem_desc(sbase, 16, 1024, LAYOUT_SW128);
      const uint64_t a0 = umma_smem_desc(sbase + 96 * 1024, 16, 1024, LAYOUT_SW128);
      const long long t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t d = tm + ((p.alt && (ks & 1)) ? 256 : 0);
          const uint64_t b = umma_desc_advance(b0, p.mn ? ks * 1024 : (ks >> 2) * (p.N * 128) + (ks & 3) * 32);
          if (p.ts) umma_tf32_ts(d, tm + 384 + (ks & 7) * 8, b, idesc, 1u);
          else umma_tf32(d, umma_desc_advance(a0, (ks >> 2) * 16384 + (ks & 3) * 32), b, idesc, 1u);
        }
      }
      umma_commit(bars);
      mbar_wait(bars, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
      stop = 1;
    }
  } else if (warp < 4 && warp < p.ldwarps) {
    unsigned long long n = 0;
    uint32_t t[32];
    while (!stop) {
      const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + 128;
      if (p.ldcols == 32) tmem_ld32(ta, t); else tmem_ld16(ta, t);
      tmem_ld_wait();
      ++n;
      if (t[0] == 0x12345u) out[3] = 1;
    }
    if (blockIdx.x == 0 && lane == 0) atomicAdd(out + 1, n);
  } else if (warp >= 4 && warp < 8 && warp - 4 < p.stswarps) {
    unsigned long long n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sts32(sbase + 128 * 1024 + warp * 4096 + i * 128 + lane * 4, 1.f);
      n += 8;
    }
    if (blockIdx.x == 0 && lane == 0) atomicAdd(out + 2, n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  const int iters = 64;
  printf("   N  ts  mn alt ldw stw | cycles/MMA   tmem-ld/MMA  sts/MMA\n");
  const P cases[] = {
      {96, 1, 0, 0, 0, 0, iters, 16}, {96, 1, 0, 1, 0, 0, iters, 16}, {96, 1, 1, 0, 0, 0, iters, 16}, {96, 0, 0, 0, 0, 0, iters, 16},
      {48, 1, 0, 0, 0, 0, iters, 16}, {48, 0, 0, 0, 0, 0, iters, 16}, {32, 0, 0, 0, 0, 0, iters, 16},
      {192, 1, 0, 0, 0, 0, iters, 16}, {256, 1, 0, 0, 0, 0, iters, 16}, {72, 1, 0, 0, 0, 0, iters, 16},
      {96, 1, 0, 0, 1, 0, iters, 16}, {96, 1, 0, 0, 4, 0, iters, 16}, {96, 1, 0, 0, 4, 0, iters, 32}, {96, 1, 0, 0, 0, 4, iters, 16}, {96, 1, 0, 0, 4, 4, iters, 16},
      {96, 1, 0, 1, 4, 4, iters, 16},
  };
  for (const P& p : cases) {
    cudaMemset(out, 0, 64);
    k<<<148, 320, 202 * 1024>>>(p, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[4];
    cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    const double mm = 16.0 * p.iters;
    printf("%4d %3d %3d %3d %3d %3d | %8.1f   %8.2f (x%d cols x32 lanes)  %8.2f\n", p.N, p.ts, p.mn, p.alt, p.ldwarps, p.stswarps,
           h[0] / mm, h[1] / mm, p.ldcols, h[2] / mm);
  }
  return 0;
}
