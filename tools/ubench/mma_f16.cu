// Probe 3: tcgen05.mma.kind::f16 (fp16 operands, fp32 accumulation) in the shapes the second-generation skinning
// backward uses: A operand in tensor memory (two halves per 32-bit column), B operand in shared memory MN-major
// (128-byte swizzle) or K-major, M = 64 at accumulator / operand lane offsets 0 and 16, M = 128; and the pace of
// chains of such MMAs next to the tf32 ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nemo_cvpr2023_b200/csrc -o mma_f16 tools/ubench/mma_f16.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "sm100.cuh"
using namespace nb::sm100;

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, __half v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(__half_as_ushort(v)) : "memory");
}
__host__ __device__ inline int aval(int m, int k) { return ((m * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int bval(int k, int n) { return ((k * 5 + n * 11) % 9) - 4; }

constexpr int KT = 128;                  // k rows stored per n block (as in the kernel)
constexpr uint32_t LBO = KT * 128;       // stride between 64-wide n blocks of the MN-major operand
// byte offset of element (n, k) of the MN-major fp16 operand: [n block][k group of 8: 1024 B][k row: 128 B][16-byte
// chunk ((n % 64) / 8) ^ (k % 8)][n % 8]
__host__ __device__ inline uint32_t mn_off(int n, int k) {
  return (n >> 6) * LBO + (k >> 3) * 1024 + (k & 7) * 128 + ((((n & 63) >> 3) ^ (k & 7)) << 4) + (n & 7) * 2;
}
// K-major fp16 operand, 128-byte rows (64 halves of K), 128-byte swizzle
__host__ __device__ inline uint32_t kmaj_off(int n, int k) { return n * 128 + ((((k >> 3) ^ n) & 7) << 4) + (k & 7) * 2; }

struct P { int test, M, N, off, iters, kind, mn; };

__global__ void __launch_bounds__(160, 1) k(P p, float* out, unsigned long long* cyc) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + 128 * 1024, slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<512>(slot);
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += 160) sts32(sbase + 4 * i, 0.f);
  __syncthreads();
  const int K = 32;
  if (p.test == 1) {  // B MN-major
    for (int i = threadIdx.x; i < K * p.N; i += 160) { const int kk = i / p.N, n = i % p.N; sts16(sbase + mn_off(n, kk), __int2half_rn(bval(kk, n))); }
  } else if (p.test == 2) {  // B K-major
    for (int i = threadIdx.x; i < K * p.N; i += 160) { const int kk = i / p.N, n = i % p.N; sts16(sbase + kmaj_off(n, kk), __int2half_rn(bval(kk, n))); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = lds32u(slot);
  constexpr int CA = 256;  // A columns
  if (warp < 4 && p.test) {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t row = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 256; c += 32) tmem_st32(row + c, z);
    // M = 64: row m lives in lane 32 (m / 16) + m % 16, second copy 16 lanes up.  M = 128: row = lane.
    const int m = p.M == 64 ? 16 * warp + (lane & 15) : threadIdx.x;
    for (int c = 0; c < K / 2; ++c) {
      const __half2 h = __halves2half2(__int2half_rn(aval(m, 2 * c)), __int2half_rn(aval(m, 2 * c + 1)));  // .x = low half = even k
      z[c] = *reinterpret_cast<const uint32_t*>(&h);
    }
    tmem_st32(row + CA, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0 && p.test) {
    const uint32_t idesc = idesc_f16(p.M, p.N, 0, p.test == 1 ? 1 : 0);
    const uint64_t b = p.test == 1 ? umma_smem_desc(sbase, LBO, 1024, LAYOUT_SW128) : umma_smem_desc(sbase, 16, 1024, LAYOUT_SW128);
    const uint32_t lo = (uint32_t)p.off << 16;
    for (int ks = 0; ks < K / 16; ++ks)
      umma_f16_ts(tm + lo, tm + lo + CA + ks * 8, umma_desc_advance(b, p.test == 1 ? ks * 2048 : ks * 32), idesc, ks ? 1u : 0u);
    umma_commit(bars);
    mbar_wait(bars, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4 && p.test) {
    for (int c = 0; c < 256; c += 32) {
      uint32_t t[32];
      tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, t);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[threadIdx.x * 256 + c + i] = __uint_as_float(t[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0 && p.iters) {
    const uint32_t idesc = p.kind ? idesc_f16(p.M, p.N, 0, p.mn) : umma_idesc_tf32(p.M, p.N, 0, p.mn);
    const uint64_t b = p.mn ? (p.kind ? umma_smem_desc(sbase, LBO, 1024, LAYOUT_SW128) : umma_smem_desc(sbase, LBO, 512, LAYOUT_SW128_BASE32B))
                            : umma_smem_desc(sbase, 16, 1024, LAYOUT_SW128);
    const uint32_t step = p.mn ? (p.kind ? 2048 : 1024) : 32;
    const uint32_t alt = p.M == 64 ? (16u << 16) : 0u;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t lo = (it & 1) ? alt : 0u;
        const uint32_t d = tm + lo, a = tm + lo + CA + (ks & 3) * 8;
        const uint64_t bd = umma_desc_advance(b, (ks & 3) * step);
        if (p.kind) umma_f16_ts(d, a, bd, idesc, 1u); else umma_tf32_ts(d, a, bd, idesc, 1u);
      }
    umma_commit(bars + 8);
    mbar_wait(bars + 8, 0);
    if (blockIdx.x == 0) cyc[0] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 128 * 256 * 4);
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  static float h[128 * 256];
  const P tests[] = {{1, 64, 192, 0, 0, 1, 1}, {1, 64, 192, 16, 0, 1, 1}, {2, 128, 144, 0, 0, 1, 0}, {1, 128, 192, 0, 0, 1, 1}};
  for (const P& p : tests) {
    cudaMemset(out, 0, sizeof(h));
    k<<<1, 160, 130 * 1024>>>(p, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kind::f16 TS, B %s, M = %d, N = %d, lane offset %d: %s\n", p.test == 1 ? "MN-major" : "K-major", p.M, p.N, p.off,
           e == cudaSuccess ? "ran" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0, shown = 0;
    for (int m = 0; m < (p.M == 64 ? 64 : 128); ++m) {
      const int L = p.M == 64 ? 32 * (m / 16) + m % 16 + p.off : m;
      for (int n = 0; n < p.N; ++n) {
        int ref = 0;
        for (int kk = 0; kk < 32; ++kk) ref += aval(m, kk) * bval(kk, n);
        if (h[L * 256 + n] != (float)ref) {
          ++bad;
          if (shown++ < 6) printf("   D[m=%d (lane %d)][n=%d] = %g, expected %d\n", m, L, n, h[L * 256 + n], ref);
        }
      }
    }
    printf("   %s (%d mismatches)\n", bad ? "MISMATCH" : "exact", bad);
  }
  printf("pace (148 CTAs, chains of 8 MMAs, TS mode):  kind  M    N  B-major | cycles/MMA\n");
  const P pace[] = {{0, 64, 192, 0, 128, 1, 1}, {0, 64, 256, 0, 128, 1, 1}, {0, 64, 96, 0, 128, 1, 1}, {0, 128, 192, 0, 128, 1, 1},
                    {0, 128, 144, 0, 128, 1, 0}, {0, 128, 72, 0, 128, 1, 0}, {0, 128, 96, 0, 128, 1, 0}, {0, 128, 256, 0, 128, 1, 0},
                    {0, 64, 192, 0, 128, 0, 1}, {0, 128, 96, 0, 128, 0, 0}};
  for (const P& p : pace) {
    cudaMemset(cyc, 0, 8);
    k<<<148, 160, 130 * 1024>>>(p, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long hc;
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("   %s %4d %4d %s | %7.1f\n", p.kind ? "f16 (K=16)" : "tf32 (K=8)", p.M, p.N, p.mn ? "MN" : "K ", hc / (p.iters * 8.0));
  }
  return 0;
}
