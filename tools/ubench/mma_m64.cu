// Micro-benchmark / probe: where does tcgen05.mma.kind::tf32 with M = 64 (cta_group::1) put its accumulator rows in
// tensor memory, and may two M = 64 accumulators share the same columns at different lane offsets?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nemo_cvpr2023_b200/csrc -o mma_m64 tools/ubench/mma_m64.cu
// A (64 x 8, K-major, shared memory) = row index + 1 in column 0, zeros elsewhere; B (16 x 8) = identity in k = 0 ->
// D[m][0] = m + 1.  All 128 lanes of column 0 are dumped after the MMA was issued at lane offset `lane0`.
#include <cstdio>
#include "sm100.cuh"
using namespace nb::sm100;

__global__ void __launch_bounds__(160, 1) k(int lane0, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = sbase + 32 * 1024, slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bars, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<64>(slot);
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 160) sts32(sbase + 4 * i, 0.f);
  __syncthreads();
  // A: K-major, 128-byte rows (32 floats of K), 128-byte swizzle: element (row, k = 0) sits at row*128 + ((0 ^ row) & 7) * 16
  if (threadIdx.x < 64) sts32(sbase + threadIdx.x * 128 + ((threadIdx.x & 7) << 4), (float)(threadIdx.x + 1));
  // B at +16 KB: rows n = 0..15, B[n][k=0] = (n == 0)
  if (threadIdx.x == 0) sts32(sbase + 16384, 1.f);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = lds32u(slot);
  // clear the 64 columns in all lanes
  if (warp < 4) {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16), z);
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 32, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = umma_idesc_tf32(64, 16, 0, 0);
    const uint64_t a = umma_smem_desc(sbase, 16, 1024, LAYOUT_SW128), b = umma_smem_desc(sbase + 16384, 16, 1024, LAYOUT_SW128);
    umma_tf32(tm + ((uint32_t)lane0 << 16), a, b, idesc, 0u);
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
  if (warp == 4) { tc_fence_after(); tmem_dealloc<64>(tm); }
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  for (int lane0 : {0, 16, 64}) {
    cudaMemset(out, 0, 512);
    k<<<1, 160, 40 * 1024>>>(lane0, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M = 64 accumulator issued at lane offset %d: %s\n", lane0, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    float h[128];
    cudaMemcpy(h, out, 512, cudaMemcpyDeviceToHost);
    for (int l = 0; l < 128; ++l) printf("%s%3.0f", l % 32 == 0 ? "\n  lanes " : " ", h[l]);
    printf("\n");
  }
  return 0;
}
