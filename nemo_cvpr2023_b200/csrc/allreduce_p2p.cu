// Sum of the shared MotionNet + RBF gradient over the ranks of one NVSwitch node through peer memory.
// Reference semantics: the single-process reference back-propagates one loss over all video instances
// (nemo/neural_motion_model.py:3511-3598), so under data parallelism the shared gradient is the sum of the ranks'.
//
// NCCL's all-reduce of the 9 MB buffer costs 78 us at 2 GPUs and ~115 us at 8 between the backward and Adam
// (profiles/r02_dp_allreduce_sweep.txt: channel count and protocol do not move it), and it cannot overlap the persistent
// kernels of the step.  Every rank's buffer lives in symmetric memory (mapped into every peer: NVSwitch gives each GPU
// full bandwidth to each peer), so the reduction is three small kernels of this library, capturable in the step graph:
//   K1  copy the local gradient into the local symmetric buffer, bump the local epoch
//   K2  signal "my buffer is ready" to every peer, wait for theirs; rank r sums slice r of all the buffers (in rank
//       order: every rank ends up with bit-identical sums) and writes the sum into slice r of EVERY buffer
//   K3  signal "my slice is written everywhere", wait for the peers'; copy the local symmetric buffer out
// Traffic per rank: (W-1)/W of the buffer read from peers and written to peers; two signal round trips.
// Signals are monotonically increasing epochs in a small area behind the data (release / acquire at system scope), so
// nothing is ever reset; a wait that never completes traps after ~2 minutes instead of hanging the GPU for good.
#include <algorithm>
#include "common.cuh"

using namespace nb;

namespace {

constexpr int MAX_WORLD = 16;
constexpr int SIG_EPOCH = 0, SIG_READY = 16, SIG_DONE = 32;   // uint32 slots of the signal area

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer / peer-written data: never through the (incoherent) L1
__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

struct Peers {
  float* buf[MAX_WORLD];
};

// every block: block 0 posts this rank's epoch to the peers' slot, then all wait until every peer has posted
__device__ __forceinline__ void exchange(const Peers& pr, size_t sig_off, int slot, int rank, int world, uint32_t epoch) {
  uint32_t* mine = reinterpret_cast<uint32_t*>(pr.buf[rank] + sig_off);
  if (blockIdx.x == 0 && threadIdx.x < world)
    st_release_sys(reinterpret_cast<uint32_t*>(pr.buf[threadIdx.x] + sig_off) + slot + rank, epoch);
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    // epochs only grow: "not yet" is a signed difference below zero
    while ((int32_t)(ld_acquire_sys(mine + slot + threadIdx.x) - epoch) < 0) {
      __nanosleep(64);
      if (clock64() - t0 > 240000000000ll) __trap();   // ~2 minutes: ranks may be seconds apart at start-up
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_p2p_stage(const float* __restrict__ src, float* __restrict__ mine, size_t n4,
                                                   size_t sig_off) {
  const float4* s = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(mine);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    d[i] = s[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t* e = reinterpret_cast<uint32_t*>(mine + sig_off) + SIG_EPOCH;
    *e = *e + 1;   // read by the next two kernels of this stream
  }
}

__global__ void __launch_bounds__(256) k_p2p_reduce(const Peers pr, size_t n4, size_t sig_off, int rank, int world) {
  const uint32_t epoch = *(reinterpret_cast<const uint32_t*>(pr.buf[rank] + sig_off) + SIG_EPOCH);
  exchange(pr, sig_off, SIG_READY, rank, world, epoch);
  const size_t per = (n4 + world - 1) / world, lo = per * rank, hi = lo + per < n4 ? lo + per : n4;
  for (size_t i = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    float4 v[MAX_WORLD];
#pragma unroll
    for (int p = 0; p < MAX_WORLD; ++p)
      if (p < world) v[p] = ld_sys4(pr.buf[p] + 4 * i);
    float4 s = v[0];
#pragma unroll
    for (int p = 1; p < MAX_WORLD; ++p)
      if (p < world) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
#pragma unroll
    for (int p = 0; p < MAX_WORLD; ++p)
      if (p < world) st_sys4(pr.buf[p] + 4 * i, s);
  }
  __threadfence_system();   // the stores into the peers' buffers are performed before this kernel ends
}

__global__ void __launch_bounds__(256) k_p2p_finish(const Peers pr, float* __restrict__ dst, size_t n4, size_t sig_off,
                                                    int rank, int world) {
  const uint32_t epoch = *(reinterpret_cast<const uint32_t*>(pr.buf[rank] + sig_off) + SIG_EPOCH);
  exchange(pr, sig_off, SIG_DONE, rank, world, epoch);
  const float* mine = pr.buf[rank];
  float4* d = reinterpret_cast<float4*>(dst);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    d[i] = ld_sys4(mine + 4 * i);
}

}  // namespace

// bufs: HOST array of `world` device pointers, entry r = rank r's symmetric buffer as mapped into THIS process (at
// least n + sig_offset... floats: n data floats, then a 64-word signal area at float offset sig_off, zeroed once by
// the caller before the first call and followed by a barrier of the group).  src / dst: the local gradient (may alias).
extern "C" int nemo_b200_p2p_allreduce(const void* const* bufs, int rank, int world, const float* src, float* dst,
                                       size_t n, size_t sig_off, void* stream) {
  NB_REQUIRE(bufs && src && dst && world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world && n % 4 == 0 &&
                 sig_off >= n && sig_off % 4 == 0,
             "p2p_allreduce: bad args (at most 16 ranks, n and sig_off multiples of 4, signal area behind the data)");
  if (n == 0) return 0;
  Peers pr{};
  for (int r = 0; r < world; ++r) {
    NB_REQUIRE(bufs[r] && ((uintptr_t)bufs[r] & 15) == 0, "p2p_allreduce: peer buffers must be 16-byte aligned");
    pr.buf[r] = reinterpret_cast<float*>(const_cast<void*>(bufs[r]));
  }
  const size_t n4 = n / 4;
  const int sms = nb::num_sms();
  const int g_copy = (int)std::min<size_t>((n4 + 255) / 256, (size_t)sms * 4);
  const size_t per = (n4 + world - 1) / world;
  const int g_red = (int)std::min<size_t>((per + 255) / 256, (size_t)sms * 2);
  k_p2p_stage<<<g_copy, 256, 0, S(stream)>>>(src, pr.buf[rank], n4, sig_off);
  NB_LAUNCH_CHECK_CONT("p2p_stage");
  k_p2p_reduce<<<g_red, 256, 0, S(stream)>>>(pr, n4, sig_off, rank, world);
  NB_LAUNCH_CHECK_CONT("p2p_reduce");
  k_p2p_finish<<<g_copy, 256, 0, S(stream)>>>(pr, dst, n4, sig_off, rank, world);
  NB_LAUNCH_CHECK("p2p_finish");
}
