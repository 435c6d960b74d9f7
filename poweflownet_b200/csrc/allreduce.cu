// One-shot all-reduce of the flat fp32 gradient buffer over NVLink / NVSwitch PEER MEMORY, as a single kernel that can
// live inside the captured CUDA graph of a training step (one sum of ~1.4 MB of parameter
// gradients per step; at that size a collective is LATENCY-bound, so NCCL's ring/tree protocols buy nothing and its
// launch sits outside the captured step).
//
// Every rank owns a symmetric receive buffer  recv[2][world][n]  (two halves used alternately, one slot per source rank)
// and a symmetric flag array  sig[ctas][world];  the peers' addresses of both are mapped into this process
// (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at start-up).  CTA c of every rank owns the same
// contiguous chunk of the buffer and runs, without any grid-wide or host synchronisation:
//   1. PUSH   its chunk of the local gradient into slot [half][rank] of EVERY rank's receive buffer (plain 16-byte
//             stores to peer addresses: posted writes over NVLink, nothing waits for a round trip),
//   2. SIGNAL fence.sys, then st.release.sys of the step's epoch number into flag [c][rank] on every peer,
//   3. WAIT   until its own flags [c][*] have all reached the epoch (ld.acquire.sys; every peer's chunk c has landed),
//   4. SUM    the `world` slots of the chunk from LOCAL memory in rank order (identical order on every rank, so every rank
//             obtains bit-identical sums) and store the result over the local gradient.
// The epoch lives in device memory and is advanced by the kernel itself, so replays of a captured graph need no host
// input; `half = epoch & 1` alternates the receive halves, which makes ONE flag exchange per step sufficient: a rank
// overwrites half h at step s+2 only after passing the step-s+1 exchange, by which time every peer has finished reading
// half h of step s (program order on the peer).  Traffic per rank: (world - 1) * n floats out -- fine for the ~1 MB
// gradient of configs/standard.json, wasteful for the 30-66 MB of the hidden-512 models, which keep using NCCL.
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

__global__ void __launch_bounds__(kArThreads)
k_allreduce_oneshot(float* __restrict__ grad, float* const* __restrict__ peer_recv, uint32_t* const* __restrict__ peer_sig,
                    uint32_t* __restrict__ epochs, int rank, int world, long long n) {
  pdl_wait();
  const int c = blockIdx.x, ctas = gridDim.x;
  // chunk of this CTA: multiples of 4 floats (the buffers are 16-byte aligned)
  const long long quads = (n + 3) / 4, per = (quads + ctas - 1) / ctas;
  const long long q_lo = min(quads, per * c), q_hi = min(quads, q_lo + per);
  const uint32_t epoch = epochs[c] + 1u;
  const long long half_off = (long long)(epoch & 1u) * world * n;
  __syncthreads();  // every thread has read the epoch before thread 0 advances it at the end
  // 1. push
  for (int p = 0; p < world; ++p) {
    float* dst = peer_recv[p] + half_off + (long long)rank * n;
    for (long long q = q_lo + threadIdx.x; q < q_hi; q += kArThreads) {
      const long long i = 4 * q;
      if (i + 3 < n) {
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(grad + i);
      } else {
        for (long long j = i; j < n; ++j) dst[j] = grad[j];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal, 3. wait (one thread per peer)
  if (threadIdx.x < world) {
    st_release_sys(peer_sig[threadIdx.x] + (long long)c * world + rank, epoch);
    const uint32_t* mine = peer_sig[rank] + (long long)c * world + threadIdx.x;
    while (int32_t(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
  // 4. sum the slots in rank order from local memory
  const float* recv = peer_recv[rank] + half_off;
  for (long long q = q_lo + threadIdx.x; q < q_hi; q += kArThreads) {
    const long long i = 4 * q;
    if (i + 3 < n) {
      float4 acc = *reinterpret_cast<const float4*>(recv + i);
      for (int r = 1; r < world; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(recv + (long long)r * n + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(grad + i) = acc;
    } else {
      for (long long j = i; j < n; ++j) {
        float acc = recv[j];
        for (int r = 1; r < world; ++r) acc += recv[(long long)r * n + j];
        grad[j] = acc;
      }
    }
  }
  if (threadIdx.x == 0) epochs[c] = epoch;
}

}  // namespace
}  // namespace pfn

using namespace pfn;

extern "C" int pfn_allreduce_oneshot(float* grad, void* const* peer_recv, void* const* peer_sig, void* epochs, int rank, int world,
                                     int64_t n, int ctas, void* stream) {
  PFN_REQUIRE(grad && peer_recv && peer_sig && epochs && world >= 1 && rank >= 0 && rank < world && n > 0 && ctas >= 1, PFN_E_INVALID,
              "pfn_allreduce_oneshot: bad arguments");
  PFN_REQUIRE(aligned16(grad) && n % 4 == 0, PFN_E_INVALID, "pfn_allreduce_oneshot: the buffer must be 16-byte aligned and n a multiple of 4");
  PFN_REQUIRE(world <= kArThreads, PFN_E_UNSUPPORTED, "pfn_allreduce_oneshot: world size too large");
  PFN_CUDA_OK(launch_kernel(k_allreduce_oneshot, dim3(static_cast<unsigned>(ctas)), dim3(kArThreads), 0, static_cast<cudaStream_t>(stream), grad,
                            reinterpret_cast<float* const*>(peer_recv), reinterpret_cast<uint32_t* const*>(peer_sig),
                            static_cast<uint32_t*>(epochs), rank, world, static_cast<long long>(n)));
  PFN_LAUNCHED();
  return 0;
}
