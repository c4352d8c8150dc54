// Probe: which 2-D fp32 TMA boxes of the dV tensor (N, 6890, 3) -- frame stride 82 680 bytes, viewed as even / odd
// frame tensors -- complete, and what lands in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nemo_cvpr2023_b200/csrc -o tma_box tools/ubench/tma_box.cu nemo_cvpr2023_b200/csrc/tmap.cu nemo_cvpr2023_b200/csrc/capi.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.cuh"
using namespace nb::sm100;

__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int box_floats, float* out, int* status) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = sbase + 32768;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sts32(sbase + 4 * i, -1.f);
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, box_floats * 4);
    tma_load_2d(sbase, &tm, bar, c0, c1);
    // bounded wait
    uint32_t done = 0;
    for (int it = 0; it < 2000000 && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(0) : "memory");
    *status = done ? 1 : -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < box_floats; i += blockDim.x) out[i] = lds32(sbase + 4 * i);
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const int N = 6, NV = 6890;
  const size_t FRAME = (size_t)NV * 12;
  std::vector<float> h((size_t)N * NV * 3);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000);
  float *d, *out; int* st;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 32768); cudaMalloc(&st, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  struct T { const char* name; int odd, binner, bouter, c0, c1; };
  const T tests[] = {{"even box 192x4 col 0", 0, 192, 4, 0, 0}, {"even box 192x4 col 384*5+192", 0, 192, 4, 384 * 5 + 192, 0},
                     {"even box 128x4", 0, 128, 4, 384, 1}, {"even box 192x1", 0, 192, 1, 384, 1},
                     {"odd box 192x4 col 2", 1, 192, 4, 2, 0}, {"odd box 192x4 col 384*5+2", 1, 192, 4, 384 * 5 + 2, 0},
                     {"odd box 192x4 col 0 (16B-aligned start)", 1, 192, 4, 0, 0}, {"odd box 196x4 col 384 (16B-aligned start)", 1, 196, 4, 384, 0},
                     {"even box 192x4 col 2 (8B-aligned start)", 0, 192, 4, 2, 0}, {"odd box 196x4 last tile", 1, 196, 4, 53 * 384 + 192, 1},
                     {"even box 192x4 last tile", 0, 192, 4, 53 * 384 + 192, 1}, {"even rows out of range", 0, 192, 4, 0, 1 << 20}};
  int idx = -1;
  for (const T& t : tests) {
    if (++idx != only && only >= 0) continue;
    CUtensorMap tm;
    int rc = t.odd ? nb::make_tmap_2d(&tm, (const char*)d + FRAME - 8, 4, FRAME / 4 + 2, N / 2, 2 * FRAME, t.binner, t.bouter)
                   : nb::make_tmap_2d(&tm, d, 4, FRAME / 4, (N + 1) / 2, 2 * FRAME, t.binner, t.bouter);
    if (rc) { printf("%s: encode failed: %s\n", t.name, nemo_b200_last_error()); continue; }
    cudaMemset(st, 0, 4);
    k<<<1, 128, 40000>>>(tm, t.c0, t.c1, t.binner * t.bouter, out, st);
    cudaError_t e = cudaDeviceSynchronize();
    int hs = 0; cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
    std::vector<float> o(t.binner * t.bouter);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < t.bouter; ++r)
      for (int c = 0; c < t.binner; ++c) {
        const long long row = t.c1 + r, col = t.c0 + c;
        float ref = 0.f;
        const long long ext = FRAME / 4 + (t.odd ? 2 : 0), rows = t.odd ? N / 2 : (N + 1) / 2;
        if (row < rows && col < ext) { const long long fl = (t.odd ? (long long)(FRAME - 8) / 4 : 0) + row * (2 * (long long)FRAME / 4) + col; ref = (float)(fl % 100000); }
        if (o[r * t.binner + c] != ref) ++bad;
      }
    printf("%-44s: %s, arrival %d, %d mismatches, first %g %g %g\n", t.name, e == cudaSuccess ? "ok" : cudaGetErrorString(e), hs, bad, o[0], o[1], o[2]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
