// Probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st) against the usual
// shared-memory A operand -- (1) is "lane = row, column = k" the layout, (2) cycles per instruction for both forms.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_a_probe tmem_a_probe.cu && ./tmem_a_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t saddr) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ uint64_t desc_mn128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {  // SWIZZLE_128B_BASE32B, MN-major TF32
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46) | (uint64_t(1) << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// element (row, k) of a K-major SWIZZLE_128B tile of 32 floats per row
__device__ __forceinline__ uint32_t sw128(int row, int k) {
  const int chunk = (k >> 2) ^ (row & 7);
  return uint32_t(row >> 3) * 1024u + uint32_t(row & 7) * 128u + uint32_t(chunk) * 16u + uint32_t(k & 3) * 4u;
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe(float* out_ss, float* out_ts, long long* cycles, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int stop_flag;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 16384u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto A = [](int m, int k) { return float(((m * 3 + k * 5) % 7) - 3); };
  auto B = [](int n, int k) { return float(((n + 2 * k) % 5) - 2); };
  for (int i = tid; i < 128 * 32; i += 128) {
    const int r = i / 32, k = i % 32;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_s + sw128(r, k)), "f"(A(r, k)));
    if (r < N) asm volatile("st.shared.f32 [%0], %1;" ::"r"(b_s + sw128(r, k)), "f"(B(r, k)));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t d_ss = tmem, d_ts = tmem + uint32_t(N), a_t = tmem + 2u * uint32_t(N);  // A in TMEM: 32 columns behind the accumulators
  // A -> TMEM: thread (warp q, lane l) owns row 32 q + l = TMEM lane 32 q + l; 32 columns = k 0..31
  {
    uint32_t v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(A(32 * warp + lane, k));
    const uint32_t taddr = a_t + (uint32_t(32 * warp) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
        "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
        "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  uint32_t phase = 0;
  if (warp == 0) {
    // correctness: 4 K steps (k = 0..31) into each accumulator
    if (elect_one()) {
      for (int j = 0; j < 4; ++j) mma_ss(d_ss, desc_k128(a_s) + uint64_t(2 * j), desc_k128(b_s) + uint64_t(2 * j), idesc, j > 0);
      for (int j = 0; j < 4; ++j) mma_ts(d_ts, a_t + uint32_t(8 * j), desc_k128(b_s) + uint64_t(2 * j), idesc, j > 0);
      commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), phase);
  phase ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    for (int which = 0; which < 2; ++which) {
      const uint32_t taddr = (which ? d_ts : d_ss) + (uint32_t(32 * warp) << 16) + uint32_t(c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* o = which ? out_ts : out_ss;
      for (int i = 0; i < 16; ++i) o[(32 * warp + lane) * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // timing: `reps` accumulating instructions back to back, one commit, wait.
  //   which 0: SS, same tiles every time      1: TS, same tiles
  //   which 2: SS in the GEMM kernel's pattern (three products per K step over hi / lo planes of 3 rotating stages, three
  //            accumulators)                  3: the same with the A planes in tensor memory
  //   which 4, 5: as 2, 3 while warps 1..3 stream through shared memory (converter-like ld.shared.v4 / st.shared.v4)
  for (int which = 0; which < 8; ++which) {
    long long t0 = 0;
    if (tid == 0) stop_flag = 0;
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
          const int j = i & 3;
          if (which == 0) mma_ss(d_ss, desc_k128(a_s) + uint64_t(2 * j), desc_k128(b_s) + uint64_t(2 * j), idesc, 1u);
          else if (which == 1) mma_ts(d_ts, a_t + uint32_t(8 * j), desc_k128(b_s) + uint64_t(2 * j), idesc, 1u);
          else {
            const int stage = (i >> 2) % 3;
            const uint32_t st = base + uint32_t(stage) * 65536u;  // A_hi | A_lo | B_hi | B_lo, 16 KB each
            const uint64_t ah = desc_k128(st) + uint64_t(2 * j), al = desc_k128(st + 16384u) + uint64_t(2 * j);
            const uint64_t bh = desc_k128(st + 32768u) + uint64_t(2 * j), bl = desc_k128(st + 49152u) + uint64_t(2 * j);
            const uint32_t dlo = tmem + 2u * uint32_t(N), dhi = tmem + uint32_t(((i >> 2) & 1) * N);
            if ((which & 1) == 0) {
              mma_ss(dlo, al, bh, idesc, 1u);
              mma_ss(dlo, ah, bl, idesc, 1u);
              mma_ss(dhi, ah, bh, idesc, 1u);
            } else {
              const uint32_t at = tmem + 3u * uint32_t(N) + uint32_t((stage & 1) * 64);
              mma_ts(dlo, at + 32u + uint32_t(8 * j), bh, idesc, 1u);
              mma_ts(dlo, at + uint32_t(8 * j), bl, idesc, 1u);
              mma_ts(dhi, at + uint32_t(8 * j), bh, idesc, 1u);
            }
          }
        }
        commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase);
      if (lane == 0) stop_flag = 1;
    } else if (which >= 6) {
      // accumulator read-out traffic (tcgen05.ld of a TMEM region the MMAs do not touch) while the MMAs run
      uint32_t r[16];
      const uint32_t taddr = tmem + (uint32_t(32 * warp) << 16) + 3u * uint32_t(N) + 64u;
      float sink = 0.f;
      while (*(volatile int*)&stop_flag == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                         "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                       : "r"(taddr + 16u * c)
                       : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          sink += __uint_as_float(r[c]);
        }
      }
      if (sink == 123.456f) out_ss[0] = sink;
    } else if (which >= 4) {
      // converter-like traffic on the stages' A planes until the MMAs are done
      uint32_t off = uint32_t(tid - 32) * 16u;
      while (*(volatile int*)&stop_flag == 0) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + 196608u + off));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + 196608u + ((off + 8192u) & 16383u)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        off = (off + 1536u) & 8191u;
      }
    }
    if (warp != 0) mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    if (tid == 0) cycles[which] = clock64();
    if (warp == 0 && t0 != 0) cycles[8 + which] = t0;
    __syncthreads();
  }
  // MN-major operands (the weight-gradient kernels' layout: [node][feature] tiles, 32-column boxes of 4 KB, K step = 1024 B):
  // which 0: A and B MN-major from shared memory; 1: A from tensor memory, B MN-major
  for (int which = 0; which < 2; ++which) {
    long long t0 = 0;
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc_mn = idesc | (1u << 16) | (which == 0 ? (1u << 15) : 0u);
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
          const int j = i & 3;
          const int stage = (i >> 2) % 3;
          const uint32_t st = base + uint32_t(stage) * 65536u;
          const uint32_t koff = uint32_t(j) * 1024u;
          const uint64_t ah = desc_mn128(st + koff, 4096u, 512u), al = desc_mn128(st + 16384u + koff, 4096u, 512u);
          const uint64_t bh = desc_mn128(st + 32768u + koff, 4096u, 512u), bl = desc_mn128(st + 49152u + koff, 4096u, 512u);
          const uint32_t dlo = tmem + 2u * uint32_t(N), dhi = tmem + uint32_t(((i >> 2) & 1) * N);
          if (which == 0) {
            mma_ss(dlo, al, bh, idesc_mn, 1u);
            mma_ss(dlo, ah, bl, idesc_mn, 1u);
            mma_ss(dhi, ah, bh, idesc_mn, 1u);
          } else {
            const uint32_t at = tmem + 3u * uint32_t(N) + uint32_t((stage & 1) * 64);
            mma_ts(dlo, at + 32u + uint32_t(8 * j), bh, idesc_mn, 1u);
            mma_ts(dlo, at + uint32_t(8 * j), bl, idesc_mn, 1u);
            mma_ts(dhi, at + uint32_t(8 * j), bh, idesc_mn, 1u);
          }
        }
        commit(smem_u32(&bar));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    if (tid == 0) cycles[17 + which] = clock64();
    if (warp == 0 && t0 != 0) cycles[19 + which] = t0;
    __syncthreads();
  }
  // latency of one K tile: 12 MMAs (GEMM pattern, A in TMEM), commit, wait -- repeated; the wait is done by ANOTHER warp
  // (warp 1, as a worker would) which then releases the issuer through a second barrier
  {
    __shared__ __align__(8) uint64_t bar2;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int rounds = 64;
    long long t0 = clock64();
    uint32_t ph2 = 0;
    for (int rd = 0; rd < rounds; ++rd) {
      if (warp == 0) {
        if (elect_one()) {
          for (int j = 0; j < 4; ++j) {
            const uint32_t st = base + uint32_t(rd % 3) * 65536u;
            const uint64_t bh = desc_k128(st + 32768u) + uint64_t(2 * j), bl = desc_k128(st + 49152u) + uint64_t(2 * j);
            const uint32_t at = tmem + 3u * uint32_t(N) + uint32_t((rd & 1) * 64);
            mma_ts(tmem + 2u * uint32_t(N), at + 32u + uint32_t(8 * j), bh, idesc, 1u);
            mma_ts(tmem + 2u * uint32_t(N), at + uint32_t(8 * j), bl, idesc, 1u);
            mma_ts(tmem + uint32_t((rd & 1) * N), at + uint32_t(8 * j), bh, idesc, 1u);
          }
          commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar2), ph2);  // released by warp 1 once it has seen the commit
      } else if (warp == 1) {
        mbar_wait(smem_u32(&bar), phase);
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      if (warp <= 1) { phase ^= (warp == 1); ph2 ^= (warp == 0); }
    }
    if (tid == 0) cycles[16] = (clock64() - t0) / rounds;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N>
void run(int reps) {
  float *ss, *ts;
  long long* cyc;
  cudaMalloc(&ss, 128 * N * 4);
  cudaMalloc(&ts, 128 * N * 4);
  cudaMalloc(&cyc, 24 * sizeof(long long));
  cudaMemset(cyc, 0, 192);
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  probe<N><<<1, 128, 216 * 1024>>>(ss, ts, cyc, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e));
    exit(1);
  }
  static float hs[128 * 256], ht[128 * 256];
  long long hc[24];
  cudaMemcpy(hs, ss, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ht, ts, 128 * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, 192, cudaMemcpyDeviceToHost);
  int bad_ss = 0, bad_ts = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0.f;
      for (int k = 0; k < 32; ++k) ref += float(((m * 3 + k * 5) % 7) - 3) * float(((n + 2 * k) % 5) - 2);
      bad_ss += hs[m * N + n] != ref;
      bad_ts += ht[m * N + n] != ref;
    }
  printf("N=%3d: mismatches SS %d, TS %d of %d | cycles per MMA: same tiles SS %.1f TS %.1f | GEMM pattern SS %.1f TS %.1f | + shared-memory traffic SS %.1f TS %.1f | + tcgen05.ld traffic SS %.1f TS %.1f | one K tile (12 MMAs) issue -> commit seen by another warp -> issuer released: %lld cycles | MN-major operands (weight-gradient layout): SS %.1f, A in TMEM %.1f cycles per MMA\n",
         N, bad_ss, bad_ts, 128 * N, double(hc[0] - hc[8]) / reps, double(hc[1] - hc[9]) / reps, double(hc[2] - hc[10]) / (3.0 * reps),
         double(hc[3] - hc[11]) / (3.0 * reps), double(hc[4] - hc[12]) / (3.0 * reps), double(hc[5] - hc[13]) / (3.0 * reps), double(hc[6] - hc[14]) / (3.0 * reps), double(hc[7] - hc[15]) / (3.0 * reps), hc[16], double(hc[17] - hc[19]) / (3.0 * reps), double(hc[18] - hc[20]) / (3.0 * reps));
}

int main() {
  run<16>(512);
  run<64>(512);
  run<128>(512);
  return 0;
}
