// All-reduce of the flat fp32 gradient buffer over NVLink / NVSwitch PEER MEMORY, as a single kernel that can live inside
// the captured CUDA graph of a training step (one sum of ~1.4 MB of parameter gradients per step; at
// that size a collective is LATENCY-bound, and an NCCL call sits outside the captured step).
//
// Every rank owns symmetric buffers  recv[2][world][n]  (two halves used alternately),  res[2][n]  and two flag arrays
// sig[2][ctas][world];  the peers' addresses of all of them are mapped into this process
// (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at start-up).  CTA c of every rank owns the same
// sub-range of every slice and runs, without any grid-wide or host synchronisation:
//
//  world <= kArOneShotMaxWorld ("one-shot"):
//   1. PUSH   its chunk of the local gradient into slot [half][rank] of EVERY rank's receive buffer (plain 16-byte stores to
//             peer addresses: posted writes over NVLink, nothing waits for a round trip),
//   2. SIGNAL fence.sys, then st.release.sys of the step's epoch number into flag [c][rank] on every peer,
//   3. WAIT   until its own flags [c][*] have all reached the epoch (ld.acquire.sys),
//   4. SUM    the `world` slots of the chunk from LOCAL memory in rank order and store the result over the local gradient.
//  larger worlds ("two-shot": reduce-scatter + all-gather, (world-1)/world x 2 n floats out per rank instead of (world-1) n --
//  measured at 8 GPUs: the one-shot form needs 44 us for 1.4 MB, NCCL 33 us):
//   1. PUSH   slice p of the local gradient into slot [half][rank] of rank p's receive buffer, for every p,
//   2. SIGNAL / WAIT on the first flag array,
//   3. SUM    the `world` slots of ITS OWN slice in rank order and push the sums into res[half][own slice] of EVERY rank,
//   4. SIGNAL / WAIT on the second flag array, then copy res[half] over the local gradient.
// Sums are taken in rank order by exactly one rank per element, so every rank ends with bit-identical values.  The epoch
// lives in device memory and is advanced by the kernel itself, so replays of a captured graph need no host input; `half =
// epoch & 1` alternates the buffers, which makes one flag exchange per phase sufficient: a rank overwrites half h at step
// s+2 only after passing the step-s+1 exchange, by which time every peer has finished reading half h of step s.
#include <cstdlib>

#include "common.cuh"

namespace pfn {
namespace {

constexpr int kArThreads = 512;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void copy_quads(float* __restrict__ dst, const float* __restrict__ src, long long q_lo, long long q_hi) {
  for (long long q = q_lo + threadIdx.x; q < q_hi; q += kArThreads)
    *reinterpret_cast<float4*>(dst + 4 * q) = *reinterpret_cast<const float4*>(src + 4 * q);
}
// signal every peer's flag [c][rank] with `epoch`, then wait until this rank's flags [c][*] have all reached it
__device__ __forceinline__ void flag_exchange(uint32_t* const* __restrict__ peer_sig, long long flag_off, int c, int rank, int world, uint32_t epoch) {
  // The CTA's stores to peer memory are ordered before the flags by ONE barrier + the release stores of `world` threads
  // (cumulativity: __syncthreads orders every thread's stores before them).  A __threadfence_system() in all 512 threads
  // instead cost 12 us per exchange at 8 GPUs -- system-scope fences of the warps of an SM drain one after the other.
  __syncthreads();
  if (threadIdx.x < world) {
    st_release_sys(peer_sig[threadIdx.x] + flag_off + (long long)c * world + rank, epoch);
    const uint32_t* mine = peer_sig[rank] + flag_off + (long long)c * world + threadIdx.x;
    while (int32_t(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
}

// n is a multiple of 4 * world (the host pads): slices of ns = n / world floats, quads of 4 floats
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define AR_STAMP(i)                                                                          \
  do {                                                                                       \
    if (stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0) stamps[i] = globaltimer_ns(); \
  } while (0)

__global__ void __launch_bounds__(kArThreads)
k_allreduce_peer(float* __restrict__ grad, float* const* __restrict__ peer_recv, float* const* __restrict__ peer_res,
                 uint32_t* const* __restrict__ peer_sig, uint32_t* __restrict__ epochs, int rank, int world, long long n, int two_shot,
                 unsigned long long* __restrict__ stamps) {
  AR_STAMP(0);
  pdl_wait();
  AR_STAMP(1);
  const int c = blockIdx.x, ctas = gridDim.x;
  const uint32_t epoch = epochs[c] + 1u;
  const long long half = (long long)(epoch & 1u);
  __syncthreads();  // every thread has read the epoch before thread 0 advances it at the end
  if (!two_shot) {
    const long long quads = n / 4, per = (quads + ctas - 1) / ctas;
    const long long q_lo = min(quads, per * c), q_hi = min(quads, q_lo + per);
    const long long half_off = half * world * n;
    for (int p = 0; p < world; ++p) copy_quads(peer_recv[p] + half_off + (long long)rank * n, grad, q_lo, q_hi);
    AR_STAMP(2);
    flag_exchange(peer_sig, 0, c, rank, world, epoch);
    AR_STAMP(3);
    const float* recv = peer_recv[rank] + half_off;
    for (long long q = q_lo + threadIdx.x; q < q_hi; q += kArThreads) {
      float4 acc = *reinterpret_cast<const float4*>(recv + 4 * q);
      for (int r = 1; r < world; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(recv + (long long)r * n + 4 * q);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(grad + 4 * q) = acc;
    }
  } else {
    const long long ns = n / world, sq = ns / 4, per = (sq + ctas - 1) / ctas;
    const long long q_lo = min(sq, per * c), q_hi = min(sq, q_lo + per);  // this CTA's quads inside every slice
    const long long recv_off = half * world * ns;                          // recv: [2][world][ns]
    // 1. reduce-scatter: slice p of the local gradient -> rank p, slot [rank]
    for (int p = 0; p < world; ++p) copy_quads(peer_recv[p] + recv_off + (long long)rank * ns, grad + (long long)p * ns, q_lo, q_hi);
    AR_STAMP(2);
    flag_exchange(peer_sig, 0, c, rank, world, epoch);
    AR_STAMP(3);
    // 2. sum the slots of the own slice in rank order; all-gather the sums into res[half][own slice] of every rank
    const float* recv = peer_recv[rank] + recv_off;
    const long long res_off = half * n + (long long)rank * ns;
    for (long long q = q_lo + threadIdx.x; q < q_hi; q += kArThreads) {
      float4 acc = *reinterpret_cast<const float4*>(recv + 4 * q);
      for (int r = 1; r < world; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(recv + (long long)r * ns + 4 * q);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      for (int p = 0; p < world; ++p) *reinterpret_cast<float4*>(peer_res[p] + res_off + 4 * q) = acc;
    }
    AR_STAMP(4);
    flag_exchange(peer_sig, (long long)ctas * world, c, rank, world, epoch);
    AR_STAMP(5);
    // 3. the reduced buffer -> the local gradient (this CTA's quads of every slice)
    const float* res = peer_res[rank] + half * n;
    for (int p = 0; p < world; ++p) copy_quads(grad + (long long)p * ns, res + (long long)p * ns, q_lo, q_hi);
  }
  if (threadIdx.x == 0) epochs[c] = epoch;
  AR_STAMP(6);
}

}  // namespace
}  // namespace pfn

using namespace pfn;

// PFN_AR_TIMING=1: CTA 0 leaves %globaltimer stamps of its phases in a device buffer the caller can read back
// (pfn_allreduce_debug_stamps); null otherwise
static unsigned long long* debug_stamps() {
  static unsigned long long* buf = [] {
    unsigned long long* p = nullptr;
    const char* e = std::getenv("PFN_AR_TIMING");
    if (e != nullptr && e[0] == '1' && cudaMalloc(&p, 8 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(p, 0, 8 * sizeof(unsigned long long));
    return p;
  }();
  return buf;
}
extern "C" int pfn_allreduce_debug_stamps(unsigned long long* host8) {
  unsigned long long* d = debug_stamps();
  if (d == nullptr || host8 == nullptr) return PFN_E_INVALID;
  return static_cast<int>(cudaMemcpy(host8, d, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
}

extern "C" int pfn_allreduce_peer(float* grad, void* const* peer_recv, void* const* peer_res, void* const* peer_sig, void* epochs, int rank,
                                  int world, int64_t n, int ctas, int two_shot, void* stream) {
  PFN_REQUIRE(grad && peer_recv && peer_res && peer_sig && epochs && world >= 1 && rank >= 0 && rank < world && n > 0 && ctas >= 1, PFN_E_INVALID,
              "pfn_allreduce_peer: bad arguments");
  PFN_REQUIRE(aligned16(grad) && n % (4 * int64_t(world)) == 0, PFN_E_INVALID,
              "pfn_allreduce_peer: the buffer must be 16-byte aligned and n a multiple of 4 * world");
  PFN_REQUIRE(world <= kArThreads, PFN_E_UNSUPPORTED, "pfn_allreduce_peer: world size too large");
  PFN_CUDA_OK(launch_kernel(k_allreduce_peer, dim3(static_cast<unsigned>(ctas)), dim3(kArThreads), 0, static_cast<cudaStream_t>(stream), grad,
                            reinterpret_cast<float* const*>(peer_recv), reinterpret_cast<float* const*>(peer_res),
                            reinterpret_cast<uint32_t* const*>(peer_sig), static_cast<uint32_t*>(epochs), rank, world,
                            static_cast<long long>(n), two_shot, debug_stamps()));
  PFN_LAUNCHED();
  return 0;
}
