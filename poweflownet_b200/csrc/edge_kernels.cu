// Segmented gather -> combine -> sum kernels over the stable CSR built by graph_prep.cu.
//
//   k_ea_fwd  : EdgeAggregation message + aggregate   (networks/MPN.py:23-28 message, :53 propagate, aggr='add')
//   k_ea_bwd  : its backward (target-side pass: dHi + dWe ; source-side pass: dHj), ReLU mask recomputed
//   k_hop     : one TAGConv propagation  x <- A_hat x     (PyG TAGConv.propagate, call site MPN.py:545)
//
// Work decomposition (all three): one thread owns one (node row, float4 column chunk); a CTA is
// `cx` chunks wide and `rows` rows tall (cx*rows <= 256) and walks a CONTIGUOUS slab of rows, so
//  - every row read/write is a run of consecutive 16-byte accesses (coalesced, vectorised),
//  - a thread's column chunk is fixed => its slice of We / its dWe partial sums live in registers,
//  - neighbour rows gathered by one CTA come from the same graph of the batch (block-diagonal
//    adjacency) and hit L1/L2 instead of HBM,
//  - a row is summed by exactly one thread in ascending edge id: no atomics, deterministic.
// These kernels are HBM/L2-bandwidth work (a few flops per byte): no tensor cores by design.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace pfn {
namespace {

struct RowTiling {
  int c4, cx, rows, threads, nblocks, npb;
};

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e != nullptr ? std::atoi(e) : dflt;
}

RowTiling make_tiling(int64_t n_nodes, int64_t h, int blocks_per_sm) {
  static const int max_threads = std::min(1024, std::max(32, env_int("PFN_EDGE_THREADS", 256)));
  static const int bps_override = env_int("PFN_EDGE_BPS", 0);
  if (bps_override > 0) blocks_per_sm = bps_override;
  RowTiling t;
  t.c4 = static_cast<int>((h + 3) / 4);
  t.cx = std::min(t.c4, max_threads);
  t.rows = std::max(1, max_threads / t.cx);
  t.threads = t.cx * t.rows;
  int64_t want = ceil_div64(std::max<int64_t>(n_nodes, 1), t.rows);
  t.nblocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(sm_count()) * blocks_per_sm)));
  t.npb = static_cast<int>(ceil_div64(std::max<int64_t>(n_nodes, 1), t.nblocks));
  return t;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void load_we(const float* __restrict__ We, int64_t ldwe, int q, int h, float4& w0,
                                        float4& w1) {
  float a[4], b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int ch = 4 * q + c;
    a[c] = ch < h ? __ldg(We + ch * ldwe) : 0.f;
    b[c] = ch < h ? __ldg(We + ch * ldwe + 1) : 0.f;
  }
  w0 = make_float4(a[0], a[1], a[2], a[3]);
  w1 = make_float4(b[0], b[1], b[2], b[3]);
}

// pre-activation of one message:  Hi[tgt] + Hj[src] + ea0 * We[:,0] + ea1 * We[:,1]
__device__ __forceinline__ float4 preact(float4 hi, float4 hj, float2 a, float4 w0, float4 w1) {
  float4 p;
  p.x = fmaf(a.y, w1.x, fmaf(a.x, w0.x, hi.x + hj.x));
  p.y = fmaf(a.y, w1.y, fmaf(a.x, w0.y, hi.y + hj.y));
  p.z = fmaf(a.y, w1.z, fmaf(a.x, w0.z, hi.z + hj.z));
  p.w = fmaf(a.y, w1.w, fmaf(a.x, w0.w, hi.w + hj.w));
  return p;
}

__device__ __forceinline__ void add_relu(float4& acc, float4 p) {
  acc.x += fmaxf(p.x, 0.f);
  acc.y += fmaxf(p.y, 0.f);
  acc.z += fmaxf(p.z, 0.f);
  acc.w += fmaxf(p.w, 0.f);
}

// ---- CSR slab staging ---------------------------------------------------------------------------------
// A CTA owns a contiguous slab of rows, so its CSR entries (rowptr slice, neighbour ids, edge_attr) are one contiguous
// range: they are copied into shared memory once, cooperatively and coalesced.  Without this every thread walks a
// chain of three dependent global loads per row (rowptr -> neighbour id -> neighbour row); with it the per-row work
// is a single round of independent gathers.  Slabs larger than the staging buffers are walked in pieces; a piece
// whose edge count exceeds the buffer (a hub bus) falls back to reading the CSR arrays from global memory.
constexpr int kSlabRows = 64;
constexpr int kSlabEdges = 768;

struct SlabSmem {
  int rowptr[kSlabRows + 1];
  int nbr[kSlabEdges];
  float2 ea[kSlabEdges];
};

struct SlabView {
  const int* nbr;    // indexed by absolute edge id minus `e0` when staged, by absolute edge id otherwise
  const float2* ea;
  int e0;
};

__device__ __forceinline__ SlabView stage_slab(SlabSmem& sm, const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                               const float2* __restrict__ ea, int r0, int nr) {
  for (int i = threadIdx.x; i <= nr; i += blockDim.x) sm.rowptr[i] = rowptr[r0 + i];
  __syncthreads();
  const int e0 = sm.rowptr[0], ne = sm.rowptr[nr] - e0;
  SlabView v;
  if (ne <= kSlabEdges) {
    for (int i = threadIdx.x; i < ne; i += blockDim.x) {
      sm.nbr[i] = nbr[e0 + i];
      if (ea != nullptr) sm.ea[i] = ea[e0 + i];
    }
    v.nbr = sm.nbr;
    v.ea = sm.ea;
    v.e0 = e0;
  } else {
    v.nbr = nbr;
    v.ea = ea;
    v.e0 = 0;
  }
  __syncthreads();
  return v;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_ea_fwd(const float* __restrict__ Hi, const float* __restrict__ Hj, int64_t ldh, const int* __restrict__ rowptr,
         const int* __restrict__ nbr, const float2* __restrict__ ea, const float* __restrict__ We, int64_t ldwe,
         float* __restrict__ S, int64_t lds, int n_nodes, int h, int c4, int cx, int rows, int npb) {
  pdl_wait();
  __shared__ SlabSmem sm;
  const int x = threadIdx.x % cx, y = threadIdx.x / cx;
  const int start = blockIdx.x * npb, end = min(n_nodes, start + npb);
  for (int r0 = start; r0 < end; r0 += kSlabRows) {  // CTA-uniform
    const int nr = min(kSlabRows, end - r0);
    const SlabView sv = stage_slab(sm, rowptr, nbr, ea, r0, nr);
    for (int q = x; q < c4; q += cx) {
      float4 w0, w1;
      load_we(We, ldwe, q, h, w0, w1);
      for (int lr = y; lr < nr; lr += rows) {
        const int node = r0 + lr;
        const float4 hi = ld4(Hi + node * ldh + 4 * q);
        const int beg = sm.rowptr[lr] - sv.e0, fin = sm.rowptr[lr + 1] - sv.e0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = beg;
        for (; e + 3 < fin; e += 4) {  // four gathers in flight per thread
          const int s0 = sv.nbr[e], s1 = sv.nbr[e + 1], s2 = sv.nbr[e + 2], s3 = sv.nbr[e + 3];
          const float4 h0 = ldg4(Hj + s0 * ldh + 4 * q), h1 = ldg4(Hj + s1 * ldh + 4 * q);
          const float4 h2 = ldg4(Hj + s2 * ldh + 4 * q), h3 = ldg4(Hj + s3 * ldh + 4 * q);
          add_relu(acc, preact(hi, h0, sv.ea[e], w0, w1));
          add_relu(acc, preact(hi, h1, sv.ea[e + 1], w0, w1));
          add_relu(acc, preact(hi, h2, sv.ea[e + 2], w0, w1));
          add_relu(acc, preact(hi, h3, sv.ea[e + 3], w0, w1));
        }
        if (e + 1 < fin) {
          const int s0 = sv.nbr[e], s1 = sv.nbr[e + 1];
          const float4 h0 = ldg4(Hj + s0 * ldh + 4 * q), h1 = ldg4(Hj + s1 * ldh + 4 * q);
          add_relu(acc, preact(hi, h0, sv.ea[e], w0, w1));
          add_relu(acc, preact(hi, h1, sv.ea[e + 1], w0, w1));
          e += 2;
        }
        if (e < fin) add_relu(acc, preact(hi, ldg4(Hj + sv.nbr[e] * ldh + 4 * q), sv.ea[e], w0, w1));
        st4(S + node * lds + 4 * q, acc);
      }
    }
    __syncthreads();  // the staging buffers are rewritten by the next piece
  }
}

// ------------------------------------------------------------------------------------------------
// k_ea_fwd_pipe: the same message + aggregate as k_ea_fwd, restructured around ASYNCHRONOUS COPIES INTO SHARED MEMORY so
// that the bytes in flight are bounded by shared memory (~200 KB per SM) instead of by registers (k_ea_fwd: 61 registers
// x 924 threads, five 16-byte row loads per thread; ncu: long_scoreboard 50 %, barrier 23 %, SMs busy 0.67 of the time).
//
//  * One persistent CTA per SM, W autonomous warps.  A warp owns a contiguous range of rows and runs its own software
//    pipeline: no CTA-wide barrier after the prologue, no producer/consumer hand-off between warps.
//  * The warp's slice of the CSR (row pointers, neighbour ids, edge_attr) is staged in shared memory in chunks of up to
//    kPipeRows rows / kPipeEdges edges, so the dependent chain rowptr -> neighbour id -> neighbour row is paid once per
//    chunk, not once per row.
//  * Rows travel in BATCHES of whole rows: R consecutive Hi rows -- ONE bulk copy (cp.async.bulk, SASS UBLKCP; the rows
//    of a node matrix are contiguous) -- followed by the gathered Hj row of every incoming edge, into a ring of
//    n_stages buffers per warp.  Gathered rows of >= 1 KB go as one bulk copy per row issued by one lane each (32 rows
//    per instruction round); shorter rows (hidden 129: 528 B) go as 16-byte cp.async (LDGSTS) chunks, whose issue
//    cost is known and small (the TMA unit's rate for sub-kilobyte copies is not).  Either way completion lands on the
//    batch's mbarrier (complete_tx bytes / cp.async.mbarrier.arrive), and the warp computes batch i while batches
//    i+1 .. i+n_stages-1 are in flight.
//  * Compute: the (row, 16-byte column chunk) pairs of a batch are dealt to the lanes round-robin (pair t -> lane t mod
//    32), so widths that are not a multiple of 32 chunks (hidden 129 = 33 chunks) leave no lane idle; the pair's Hi chunk
//    sits at byte 16 t of the batch buffer.  Per pair the edges are summed in ascending edge id with the same FMAs as
//    k_ea_fwd: the two kernels are BIT-IDENTICAL (tests/test_gpu_kernels.py).  S rows leave as 16-byte stores at
//    consecutive addresses.
//  * Small problems (both node matrices within a fraction of L2): before griddepcontrol.wait every warp asks the L2 to
//    prefetch its slice of Hi, Hj and the CSR arrays (cp.async.bulk.prefetch.L2): the HBM reads start at once instead
//    of trickling in behind the two dependent metadata loads, and the gathers become L2 hits.
//  A row whose degree exceeds a batch buffer (a hub bus) is summed straight from global memory by the same warp.
constexpr int kPipeRows = 64;     // rows of CSR metadata staged per chunk and warp
constexpr int kPipeEdges = 256;   // edges of CSR metadata staged per chunk and warp
constexpr int kPipeMaxStages = 4;
constexpr int kPipeMaxWarps = 8;

struct PipeWarpMeta {
  int rp[kPipeRows + 4];
  int nb[kPipeEdges];
  float2 ea[kPipeEdges];
};

struct PipeArgs {
  const float* Hi;
  const float* Hj;
  const int* rowptr;
  const int* nbr;
  const float2* ea;
  const float* We;
  float* S;
  long long ldh, lds, ldwe;
  int n_nodes, h, c4, warps, n_stages, cap_slots, rows_per_warp, prefetch_l2, early_trigger, contiguous;
  unsigned stage_bytes;
  long long e_cap;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t pipe_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void pipe_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pipe_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pipe_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// prefetch this warp's 1/total share of [base, base + bytes) into L2 (one lane; 16-byte granules, pieces of <= 64 KB)
__device__ __forceinline__ void prefetch_share(const void* base, long long bytes, int part, int parts) {
  const long long gran = (bytes + 15) >> 4;  // 16-byte granules (the arrays are padded to 16 bytes)
  const long long per = (gran + parts - 1) / parts;
  long long lo = per * part, hi = min(gran, lo + per);
  const char* p = static_cast<const char*>(base);
  for (; lo < hi; lo += 4096) bulk_prefetch_l2(p + (lo << 4), static_cast<uint32_t>(min((long long)4096, hi - lo) << 4));
}

template <bool BULK>
__global__ void __launch_bounds__(32 * kPipeMaxWarps, 1) k_ea_fwd_pipe(const __grid_constant__ PipeArgs a) {
  extern __shared__ __align__(128) uint8_t pipe_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = a.c4, W = a.warps, NS = a.n_stages, cap = a.cap_slots;
  const uint32_t rowbytes = uint32_t(c4) * 16u;
  // shared-memory map: We table [c4][2] float4 | barriers [W][kPipeMaxStages] | per-warp metadata | per-warp stage ring
  float4* s_w = reinterpret_cast<float4*>(pipe_smem);
  const uint32_t off_bar = (uint32_t(c4) * 32u + 127u) & ~127u;
  const uint32_t off_meta = off_bar + ((uint32_t(W) * kPipeMaxStages * 8u + 127u) & ~127u);
  const uint32_t meta_bytes = (uint32_t(sizeof(PipeWarpMeta)) + 127u) & ~127u;
  const uint32_t off_data = off_meta + uint32_t(W) * meta_bytes;
  PipeWarpMeta& m = *reinterpret_cast<PipeWarpMeta*>(pipe_smem + off_meta + uint32_t(warp) * meta_bytes);
  const uint32_t smem0 = pipe_smem_u32(pipe_smem);
  const uint32_t bar0 = smem0 + off_bar + uint32_t(warp) * kPipeMaxStages * 8u;
  const uint32_t data0 = smem0 + off_data + uint32_t(warp) * uint32_t(NS) * a.stage_bytes;

  const int gw = blockIdx.x * W + warp, n_warps = gridDim.x * W;
  const int w_start = min(a.n_nodes, gw * a.rows_per_warp), w_end = min(a.n_nodes, w_start + a.rows_per_warp);

  // ---- prologue that may overlap the previous kernel's tail (nothing here reads data as a value) ----
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) pipe_mbar_init(bar0 + 8u * s, BULK ? 1u : 33u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (a.prefetch_l2) {
      if (a.contiguous && w_end > w_start) {
        // the warp's own Hi rows; its 1/n_warps share of Hj (some warp gathers every row of it exactly once from HBM)
        const long long lo = (long long)w_start * a.ldh * 4, hi = (long long)w_end * a.ldh * 4;
        for (long long o = lo; o < hi; o += 65536)
          bulk_prefetch_l2(reinterpret_cast<const char*>(a.Hi) + o, static_cast<uint32_t>(min((long long)65536, hi - o)));
      }
      if (a.contiguous) prefetch_share(a.Hj, (long long)a.n_nodes * a.ldh * 4, gw, n_warps);
      prefetch_share(a.nbr, a.e_cap * 4, gw, n_warps);
      prefetch_share(a.ea, a.e_cap * 8, gw, n_warps);
      prefetch_share(a.rowptr, ((long long)a.n_nodes + 1) * 4, gw, n_warps);
    }
  }
  pdl_wait();
  if (a.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = threadIdx.x; i < 2 * c4; i += blockDim.x) {
    const int q = i >> 1, k = i & 1;
    float w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ch = 4 * q + c;
      w[c] = ch < a.h ? __ldg(a.We + ch * a.ldwe + k) : 0.f;
    }
    s_w[2 * q + k] = make_float4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();  // the only CTA-wide barrier: We table + barrier initialisation
  if (warp >= W) return;

  // largest R in [0, min(32, n - r0)] with R + (edges of rows [r0, r0 + R)) <= cap slots; 0 = row r0 alone does not fit
  auto greedy = [&](int r0, int n) {
    const int c = lane + 1;
    bool ok = false;
    if (r0 + c <= n) ok = c + (m.rp[r0 + c] - m.rp[r0]) <= cap;
    return __popc(__ballot_sync(0xffffffffu, ok));
  };

  unsigned issued = 0, consumed = 0;  // batches over the whole kernel: stage = idx % NS, mbarrier parity = (idx / NS) & 1
  int row = w_start;
  while (row < w_end) {  // ---- one chunk of CSR metadata ----
    const int nr_try = min(kPipeRows, w_end - row);
    for (int i = lane; i <= nr_try; i += 32) m.rp[i] = a.rowptr[row + i];
    __syncwarp();
    const int e_base = m.rp[0];
    int n;
    {
      const int c1 = lane + 1, c2 = lane + 33;
      const bool ok1 = c1 <= nr_try && m.rp[c1] - e_base <= kPipeEdges;
      const bool ok2 = c2 <= nr_try && m.rp[c2] - e_base <= kPipeEdges;
      n = __popc(__ballot_sync(0xffffffffu, ok1)) + __popc(__ballot_sync(0xffffffffu, ok2));
    }
    const int ne = n > 0 ? m.rp[n] - e_base : 0;
    for (int i = lane; i < ne; i += 32) {
      m.nb[i] = a.nbr[e_base + i];
      m.ea[i] = a.ea[e_base + i];
    }
    __syncwarp();

    // a row summed straight from global memory (hub bus: more incident edges than a batch buffer / a metadata chunk holds)
    auto slow_row = [&](int node) {
      const int beg = a.rowptr[node], fin = a.rowptr[node + 1];
      for (int q = lane; q < c4; q += 32) {
        const float4 w0 = s_w[2 * q], w1 = s_w[2 * q + 1];
        const float4 hi = ld4(a.Hi + (size_t)node * a.ldh + 4 * q);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = beg; e < fin; ++e)
          add_relu(acc, preact(hi, ldg4(a.Hj + (size_t)a.nbr[e] * a.ldh + 4 * q), a.ea[e], w0, w1));
        st4(a.S + (size_t)node * a.lds + 4 * q, acc);
      }
    };
    if (n == 0) {  // the first row alone has more than kPipeEdges edges
      slow_row(row);
      ++row;
      __syncwarp();
      continue;
    }

    int issue_ptr = 0, cons_ptr = 0;  // local row indices inside the chunk
    while (true) {
      // ---- issue as many batches as the ring holds ----
      while (issued - consumed < unsigned(NS) && issue_ptr < n) {
        const int R = greedy(issue_ptr, n);
        if (R == 0) break;  // hub row: handled below once the ring has drained
        const uint32_t bar = bar0 + 8u * (issued % NS), buf = data0 + (issued % NS) * a.stage_bytes;
        const int e0 = m.rp[issue_ptr] - e_base, nE = m.rp[issue_ptr + R] - e_base - e0;
        const float* hi_src = a.Hi + (size_t)(row + issue_ptr) * a.ldh;
        __syncwarp();  // every lane has finished reading the buffer's previous contents
        if (lane == 0) {
          const uint32_t tx = (a.contiguous || BULK ? uint32_t(R) * rowbytes : 0u) + (BULK ? uint32_t(nE) * rowbytes : 0u);
          pipe_mbar_expect_tx(bar, tx);
          if (a.contiguous) bulk_g2s(buf, hi_src, uint32_t(R) * rowbytes, bar);  // R consecutive Hi rows: one bulk copy
        }
        __syncwarp();
        if (BULK) {
          if (!a.contiguous && lane < R) bulk_g2s(buf + uint32_t(lane) * rowbytes, hi_src + (size_t)lane * a.ldh, rowbytes, bar);
          for (int i = lane; i < nE; i += 32)
            bulk_g2s(buf + uint32_t(R + i) * rowbytes, a.Hj + (size_t)m.nb[e0 + i] * a.ldh, rowbytes, bar);
        } else {
          // 16-byte chunks, dealt to the lanes round-robin over the (row, chunk) pairs of the batch
          const int first = a.contiguous ? R : 0, total = (R + nE) * c4;
          int k = first, q = lane;
          while (q >= c4) { q -= c4; ++k; }
          for (int t = first * c4 + lane; t < total; t += 32) {
            const float* src = k < R ? a.Hi + (size_t)(row + issue_ptr + k) * a.ldh : a.Hj + (size_t)m.nb[e0 + k - R] * a.ldh;
            cp_async16(buf + uint32_t(t) * 16u, src + 4 * q);
            q += 32;
            while (q >= c4) { q -= c4; ++k; }
          }
          cp_async_arrive_noinc(bar);
        }
        ++issued;
        issue_ptr += R;
      }
      if (consumed == issued) {
        if (issue_ptr >= n) break;
        slow_row(row + issue_ptr);  // nothing in flight and the next row does not fit a batch
        ++issue_ptr;
        cons_ptr = issue_ptr;
        __syncwarp();
        continue;
      }
      // ---- consume the oldest batch ----
      const int R = greedy(cons_ptr, n);  // the same value as when the batch was issued
      const uint32_t bar = bar0 + 8u * (consumed % NS), buf = data0 + (consumed % NS) * a.stage_bytes;
      pipe_mbar_wait(bar, (consumed / NS) & 1u);
      const int e0 = m.rp[cons_ptr] - e_base;
      const uint32_t gat = buf + uint32_t(R) * rowbytes;  // gathered rows: edge e of the chunk at slot e - e0
      float* s_dst = a.S + (size_t)(row + cons_ptr) * a.lds;
      const int total = R * c4;
      int k = 0, q = lane;
      while (q >= c4) { q -= c4; ++k; }
      for (int t = lane; t < total; t += 32) {
        const float4 hi = lds4(buf + uint32_t(t) * 16u);
        const float4 w0 = s_w[2 * q], w1 = s_w[2 * q + 1];
        const int beg = m.rp[cons_ptr + k] - e_base, fin = m.rp[cons_ptr + k + 1] - e_base;
        const uint32_t col = gat + uint32_t(q) * 16u;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = beg;
        for (; e + 3 < fin; e += 4) {  // four shared-memory row reads in flight
          const float4 h0 = lds4(col + uint32_t(e - e0) * rowbytes), h1 = lds4(col + uint32_t(e + 1 - e0) * rowbytes);
          const float4 h2 = lds4(col + uint32_t(e + 2 - e0) * rowbytes), h3 = lds4(col + uint32_t(e + 3 - e0) * rowbytes);
          add_relu(acc, preact(hi, h0, m.ea[e], w0, w1));
          add_relu(acc, preact(hi, h1, m.ea[e + 1], w0, w1));
          add_relu(acc, preact(hi, h2, m.ea[e + 2], w0, w1));
          add_relu(acc, preact(hi, h3, m.ea[e + 3], w0, w1));
        }
        if (e + 1 < fin) {
          const float4 h0 = lds4(col + uint32_t(e - e0) * rowbytes), h1 = lds4(col + uint32_t(e + 1 - e0) * rowbytes);
          add_relu(acc, preact(hi, h0, m.ea[e], w0, w1));
          add_relu(acc, preact(hi, h1, m.ea[e + 1], w0, w1));
          e += 2;
        }
        if (e < fin) add_relu(acc, preact(hi, lds4(col + uint32_t(e - e0) * rowbytes), m.ea[e], w0, w1));
        st4(s_dst + (size_t)k * a.lds + 4 * q, acc);
        q += 32;
        while (q >= c4) { q -= c4; ++k; }
      }
      ++consumed;
      cons_ptr += R;
    }
    row += n;
    __syncwarp();  // the metadata of this chunk is overwritten next
  }
}

// ------------------------------------------------------------------------------------------------
// blockIdx.y == 0: target-side pass over the CSR by target:  dHi[i] = sum_{e in in(i)} dS[i] * 1[p_e > 0]
//                  and the per-CTA partial of dWe[c,k] = sum_e g_e[c] * ea_e[k]
// blockIdx.y == 1: source-side pass over the CSR by source:  dHj[j] = sum_{e in out(j)} dS[tgt e] * 1[p_e > 0]
__global__ void __launch_bounds__(1024)
k_ea_bwd(const float* __restrict__ dS, int64_t ldds, const float* __restrict__ Hi, const float* __restrict__ Hj,
         int64_t ldh, const int* __restrict__ rowptr_t, const int* __restrict__ nbr_t,
         const float2* __restrict__ ea_t, const int* __restrict__ rowptr_s, const int* __restrict__ nbr_s,
         const float2* __restrict__ ea_s, const float* __restrict__ We, int64_t ldwe, float* __restrict__ dHi,
         float* __restrict__ dHj, int64_t ldd, float* __restrict__ dwe_partial, int n_nodes, int h, int c4, int cx,
         int rows, int npb) {
  pdl_wait();
  __shared__ SlabSmem sm;
  __shared__ float red[8][1024];
  const int x = threadIdx.x % cx, y = threadIdx.x / cx;
  const int start = blockIdx.x * npb, end = min(n_nodes, start + npb);
  const bool target_side = blockIdx.y == 0;
  const int nq = (c4 + cx - 1) / cx;  // column passes (1 unless hidden_dim > 1024)
  float4 g0[1] = {make_float4(0.f, 0.f, 0.f, 0.f)}, g1[1] = {make_float4(0.f, 0.f, 0.f, 0.f)};
  for (int qi = 0; qi < nq; ++qi) {  // CTA-uniform loops: the body holds barriers
    const int q = qi * cx + x;
    const bool active = q < c4;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
    if (active) load_we(We, ldwe, q, h, w0, w1);
    g0[0] = g1[0] = make_float4(0.f, 0.f, 0.f, 0.f);  // dWe[:,0], dWe[:,1] partial sums of this chunk
    for (int r0 = start; r0 < end; r0 += kSlabRows) {
      const int nr = min(kSlabRows, end - r0);
      const SlabView sv = target_side ? stage_slab(sm, rowptr_t, nbr_t, ea_t, r0, nr) : stage_slab(sm, rowptr_s, nbr_s, ea_s, r0, nr);
      if (active) {
        for (int lr = y; lr < nr; lr += rows) {
          const int node = r0 + lr;
          const int beg = sm.rowptr[lr] - sv.e0, fin = sm.rowptr[lr + 1] - sv.e0;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (target_side) {
            const float4 hi = ld4(Hi + node * ldh + 4 * q);
            const float4 ds = ld4(dS + node * ldds + 4 * q);
            int e = beg;
            for (; e < fin; e += 2) {
              const bool two = e + 1 < fin;
              const int s0 = sv.nbr[e], s1 = two ? sv.nbr[e + 1] : s0;
              const float4 h0 = ldg4(Hj + s0 * ldh + 4 * q), h1 = ldg4(Hj + s1 * ldh + 4 * q);
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                const float2 a = sv.ea[e + u];
                const float4 p = preact(hi, u == 0 ? h0 : h1, a, w0, w1);
                float4 g;
                g.x = p.x > 0.f ? ds.x : 0.f;
                g.y = p.y > 0.f ? ds.y : 0.f;
                g.z = p.z > 0.f ? ds.z : 0.f;
                g.w = p.w > 0.f ? ds.w : 0.f;
                acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
                g0[0].x = fmaf(g.x, a.x, g0[0].x); g0[0].y = fmaf(g.y, a.x, g0[0].y); g0[0].z = fmaf(g.z, a.x, g0[0].z); g0[0].w = fmaf(g.w, a.x, g0[0].w);
                g1[0].x = fmaf(g.x, a.y, g1[0].x); g1[0].y = fmaf(g.y, a.y, g1[0].y); g1[0].z = fmaf(g.z, a.y, g1[0].z); g1[0].w = fmaf(g.w, a.y, g1[0].w);
              }
            }
            st4(dHi + node * ldd + 4 * q, acc);
          } else {
            const float4 hj = ld4(Hj + node * ldh + 4 * q);
            int e = beg;
            for (; e < fin; e += 2) {
              const bool two = e + 1 < fin;
              const int t0 = sv.nbr[e], t1 = two ? sv.nbr[e + 1] : t0;
              const float4 hi0 = ldg4(Hi + t0 * ldh + 4 * q), hi1 = ldg4(Hi + t1 * ldh + 4 * q);
              const float4 ds0 = ldg4(dS + t0 * ldds + 4 * q), ds1 = ldg4(dS + t1 * ldds + 4 * q);
              const float4 p0 = preact(hi0, hj, sv.ea[e], w0, w1);
              acc.x += p0.x > 0.f ? ds0.x : 0.f;
              acc.y += p0.y > 0.f ? ds0.y : 0.f;
              acc.z += p0.z > 0.f ? ds0.z : 0.f;
              acc.w += p0.w > 0.f ? ds0.w : 0.f;
              if (two) {
                const float4 p1 = preact(hi1, hj, sv.ea[e + 1], w0, w1);
                acc.x += p1.x > 0.f ? ds1.x : 0.f;
                acc.y += p1.y > 0.f ? ds1.y : 0.f;
                acc.z += p1.z > 0.f ? ds1.z : 0.f;
                acc.w += p1.w > 0.f ? ds1.w : 0.f;
              }
            }
            st4(dHj + node * ldd + 4 * q, acc);
          }
        }
      }
      __syncthreads();
    }
    if (target_side) {
      // fixed-order reduction over the CTA's rows, then one partial row per CTA (no atomics)
      red[0][threadIdx.x] = g0[0].x; red[1][threadIdx.x] = g0[0].y; red[2][threadIdx.x] = g0[0].z; red[3][threadIdx.x] = g0[0].w;
      red[4][threadIdx.x] = g1[0].x; red[5][threadIdx.x] = g1[0].y; red[6][threadIdx.x] = g1[0].z; red[7][threadIdx.x] = g1[0].w;
      __syncthreads();
      if (y == 0 && active) {
        float sum[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sum[k] = 0.f;
        for (int yy = 0; yy < rows; ++yy) {
#pragma unroll
          for (int k = 0; k < 8; ++k) sum[k] += red[k][yy * cx + x];
        }
        // partial layout [2][4*c4][nblocks]: the final reduction reads consecutive CTAs contiguously
        const size_t nb = gridDim.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dwe_partial[(size_t(0) * 4 * c4 + 4 * q + c) * nb + blockIdx.x] = sum[c];
          dwe_partial[(size_t(1) * 4 * c4 + 4 * q + c) * nb + blockIdx.x] = sum[4 + c];
        }
      }
      __syncthreads();
    }
  }
}

// dWe[c, k] = sum over CTAs of the partial rows: one warp per output element, lanes stride over the CTAs (contiguous
// in memory) and combine with a fixed shuffle tree (deterministic; no atomics)
__global__ void k_reduce_dwe(const float* __restrict__ partial, int nblocks, int c4, int h, float* __restrict__ dWe,
                             int64_t lddwe) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * h) return;
  const int k = warp / h, c = warp - k * h;
  const float* src = partial + (size_t(k) * 4 * c4 + c) * nblocks;
  float sum = 0.f;
  for (int b = lane; b < nblocks; b += 32) sum += src[b];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) dWe[c * lddwe + k] = sum;
}

// ------------------------------------------------------------------------------------------------
// Y[i] = dis[i] * sum_{e in row i} dis[nbr e] * X[nbr e]  (+ addend[i])  (* (ymask[i] > 0 ? scale : 0))
__global__ void __launch_bounds__(1024)
k_hop(const float* __restrict__ X, int64_t ldx, const int* __restrict__ rowptr, const int* __restrict__ nbr,
      const float* __restrict__ dis, const float* addend, int64_t ldadd, const float* __restrict__ ymask,
      int64_t ldym, float scale, float* Y, int64_t ldy, int n_nodes, int c4, int cx, int rows, int npb) {
  pdl_wait();
  __shared__ SlabSmem sm;
  const int x = threadIdx.x % cx, y = threadIdx.x / cx;
  const int start = blockIdx.x * npb, end = min(n_nodes, start + npb);
  for (int r0 = start; r0 < end; r0 += kSlabRows) {
    const int nr = min(kSlabRows, end - r0);
    const SlabView sv = stage_slab(sm, rowptr, nbr, nullptr, r0, nr);
    for (int q = x; q < c4; q += cx) {
      for (int lr = y; lr < nr; lr += rows) {
        const int node = r0 + lr;
        const int beg = sm.rowptr[lr] - sv.e0, fin = sm.rowptr[lr + 1] - sv.e0;
        const float di = dis[node];
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), m4 = a4;
        if (addend != nullptr) a4 = ld4(addend + node * ldadd + 4 * q);
        if (ymask != nullptr) m4 = ld4(ymask + node * ldym + 4 * q);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = beg;
        for (; e + 3 < fin; e += 4) {
          const int s0 = sv.nbr[e], s1 = sv.nbr[e + 1], s2 = sv.nbr[e + 2], s3 = sv.nbr[e + 3];
          const float d0 = __ldg(dis + s0), d1 = __ldg(dis + s1), d2 = __ldg(dis + s2), d3 = __ldg(dis + s3);
          const float4 v0 = ldg4(X + s0 * ldx + 4 * q), v1 = ldg4(X + s1 * ldx + 4 * q);
          const float4 v2 = ldg4(X + s2 * ldx + 4 * q), v3 = ldg4(X + s3 * ldx + 4 * q);
          acc.x = fmaf(d0, v0.x, acc.x); acc.y = fmaf(d0, v0.y, acc.y); acc.z = fmaf(d0, v0.z, acc.z); acc.w = fmaf(d0, v0.w, acc.w);
          acc.x = fmaf(d1, v1.x, acc.x); acc.y = fmaf(d1, v1.y, acc.y); acc.z = fmaf(d1, v1.z, acc.z); acc.w = fmaf(d1, v1.w, acc.w);
          acc.x = fmaf(d2, v2.x, acc.x); acc.y = fmaf(d2, v2.y, acc.y); acc.z = fmaf(d2, v2.z, acc.z); acc.w = fmaf(d2, v2.w, acc.w);
          acc.x = fmaf(d3, v3.x, acc.x); acc.y = fmaf(d3, v3.y, acc.y); acc.z = fmaf(d3, v3.z, acc.z); acc.w = fmaf(d3, v3.w, acc.w);
        }
        for (; e < fin; ++e) {
          const int s0 = sv.nbr[e];
          const float d0 = __ldg(dis + s0);
          const float4 v0 = ldg4(X + s0 * ldx + 4 * q);
          acc.x = fmaf(d0, v0.x, acc.x); acc.y = fmaf(d0, v0.y, acc.y); acc.z = fmaf(d0, v0.z, acc.z); acc.w = fmaf(d0, v0.w, acc.w);
        }
        float4 out = make_float4(fmaf(di, acc.x, a4.x), fmaf(di, acc.y, a4.y), fmaf(di, acc.z, a4.z), fmaf(di, acc.w, a4.w));
        if (ymask != nullptr) {
          out.x = m4.x > 0.f ? out.x * scale : 0.f;
          out.y = m4.y > 0.f ? out.y * scale : 0.f;
          out.z = m4.z > 0.f ? out.z * scale : 0.f;
          out.w = m4.w > 0.f ? out.w * scale : 0.f;
        }
        st4(Y + node * ldy + 4 * q, out);
      }
    }
    __syncthreads();
  }
}

// ---- host side of k_ea_fwd_pipe --------------------------------------------------------------------------------------
// PFN_EA_FWD=cta forces the CTA-slab kernel (k_ea_fwd), PFN_EA_FWD=pipe / unset takes the pipelined kernel whenever the
// width fits; read per call so that tests can compare the two.  Tuning knobs (experiments): PFN_EA_STAGES (2..4),
// PFN_EA_WARPS (1..8), PFN_EA_PREFETCH (0/1), PFN_EA_TRIGGER (0/1), PFN_EA_BULK (0/1: gathered rows as bulk copies).
constexpr uint32_t kPipeSmemLimit = 227 * 1024;
constexpr long long kPipePrefetchMaxBytes = 48ll << 20;  // L2-prefetch both node matrices only when they fit well inside L2

bool ea_fwd_use_pipe() {
  const char* e = std::getenv("PFN_EA_FWD");
  return !(e != nullptr && e[0] == 'c');
}

// 0 = launched, 1 = shape outside this kernel (caller uses k_ea_fwd)
int ea_fwd_pipe_launch(const float* Hi, const float* Hj, int64_t ldh, const GraphView& g, int64_t n_nodes, const float* We,
                       int64_t ldwe, float* S, int64_t lds, int64_t h, cudaStream_t stream) {
  const int c4 = static_cast<int>((h + 3) / 4);
  const uint32_t rowbytes = uint32_t(c4) * 16u;
  if (n_nodes >= (int64_t(1) << 31) - 64 || rowbytes > 16384u) return 1;
  PipeArgs a{};
  const int env_bulk = env_int("PFN_EA_BULK", -1);
  const bool bulk = env_bulk >= 0 ? env_bulk != 0 : rowbytes >= 1024u;
  int warps = rowbytes <= 1024u ? 8 : rowbytes <= 4096u ? 4 : 2;
  warps = std::max(1, std::min(kPipeMaxWarps, env_int("PFN_EA_WARPS", warps)));
  int stages = std::max(2, std::min(kPipeMaxStages, env_int("PFN_EA_STAGES", 3)));
  const uint32_t meta_bytes = (uint32_t(sizeof(PipeWarpMeta)) + 127u) & ~127u;
  auto fixed_bytes = [&](int w) {
    return ((uint32_t(c4) * 32u + 127u) & ~127u) + ((uint32_t(w) * kPipeMaxStages * 8u + 127u) & ~127u) + uint32_t(w) * meta_bytes;
  };
  // at least 8 slots (a row and seven incident edges) per batch buffer: fewer warps, then fewer stages
  uint32_t stage_bytes = 0;
  int cap = 0;
  for (;;) {
    const uint32_t per_warp = (kPipeSmemLimit - 128u - fixed_bytes(warps)) / uint32_t(warps);
    stage_bytes = (per_warp / uint32_t(stages)) & ~127u;
    cap = static_cast<int>(stage_bytes / rowbytes);
    if (cap >= 8) break;
    if (stages > 2) --stages;
    else if (warps > 1) warps = std::max(1, warps / 2);
    else break;
  }
  if (cap < 2) return 1;
  cap = std::min(cap, 32 + kPipeEdges);
  a.Hi = Hi;
  a.Hj = Hj;
  a.rowptr = g.rowptr_t;
  a.nbr = g.nbr_t;
  a.ea = reinterpret_cast<const float2*>(g.ea_t);
  a.We = We;
  a.S = S;
  a.ldh = ldh;
  a.lds = lds;
  a.ldwe = ldwe;
  a.n_nodes = static_cast<int>(n_nodes);
  a.h = static_cast<int>(h);
  a.c4 = c4;
  a.warps = warps;
  a.n_stages = stages;
  a.cap_slots = cap;
  a.stage_bytes = stage_bytes;
  a.contiguous = (ldh * 4 == int64_t(rowbytes)) ? 1 : 0;
  a.e_cap = g.e_cap;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(sm_count(), ceil_div64(n_nodes, warps))));
  a.rows_per_warp = static_cast<int>(ceil_div64(n_nodes, int64_t(grid) * warps));
  a.prefetch_l2 = env_int("PFN_EA_PREFETCH", 2 * n_nodes * int64_t(rowbytes) <= kPipePrefetchMaxBytes ? 1 : 0) != 0 ? 1 : 0;
  a.early_trigger = env_int("PFN_EA_TRIGGER", 0) != 0 ? 1 : 0;
  const uint32_t smem = fixed_bytes(warps) + uint32_t(warps) * uint32_t(stages) * stage_bytes + 128u;
  static SmemAttrOnce attr_once;
  PFN_CUDA_OK(ensure_dynamic_smem(attr_once, [] {
    const cudaError_t e = cudaFuncSetAttribute(k_ea_fwd_pipe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPipeSmemLimit));
    return e != cudaSuccess ? e : cudaFuncSetAttribute(k_ea_fwd_pipe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPipeSmemLimit));
  }));
  PFN_CUDA_OK(launch_kernel(bulk ? k_ea_fwd_pipe<true> : k_ea_fwd_pipe<false>, dim3(grid), dim3(32 * warps), smem, stream, a));
  PFN_LAUNCHED();
  return 0;
}

constexpr int kBlocksPerSm = 4;  // measured: 4 x 231-thread CTAs per SM beat 8 (and 1 x 1024) at case118 sizes

bool rows_ok(const void* p, int64_t ld) { return p != nullptr && aligned16(p) && ld % 4 == 0; }

}  // namespace

int ea_fwd_launch(const float* Hi, const float* Hj, int64_t ldh, const GraphView& g, int64_t n_nodes, const float* We,
                  int64_t ldwe, float* S, int64_t lds, int64_t h, cudaStream_t stream) {
  PFN_REQUIRE(rows_ok(Hi, ldh) && rows_ok(Hj, ldh) && rows_ok(S, lds) && We != nullptr, PFN_E_INVALID,
              "ea_fwd: node matrices must be 16-byte aligned with ld %% 4 == 0");
  PFN_REQUIRE(h > 0 && ldh >= round_up64(h, 4) && lds >= round_up64(h, 4), PFN_E_INVALID, "ea_fwd: ld < round_up(h,4)");
  if (n_nodes == 0) return 0;
  const RowTiling t = make_tiling(n_nodes, h, kBlocksPerSm);
  ProfScope prof(PFN_PROF_EA_FWD, stream);
  if (ea_fwd_use_pipe()) {
    const int rc = ea_fwd_pipe_launch(Hi, Hj, ldh, g, n_nodes, We, ldwe, S, lds, h, stream);
    if (rc != 1) return rc;
  }
  PFN_CUDA_OK(launch_kernel(k_ea_fwd, dim3(t.nblocks), dim3(t.threads), 0, stream, Hi, Hj, ldh, g.rowptr_t, g.nbr_t, reinterpret_cast<const float2*>(g.ea_t),
                                               We, ldwe, S, lds, static_cast<int>(n_nodes), static_cast<int>(h), t.c4,
                                               t.cx, t.rows, t.npb));
  PFN_LAUNCHED();
  return 0;
}

int ea_bwd_launch(const float* dS, int64_t ldds, const float* Hi, const float* Hj, int64_t ldh, const GraphView& g,
                  int64_t n_nodes, const float* We, int64_t ldwe, float* dHi, float* dHj, int64_t ldd, float* dWe,
                  int64_t lddwe, void* scratch, int64_t h, cudaStream_t stream) {
  PFN_REQUIRE(rows_ok(dS, ldds) && rows_ok(Hi, ldh) && rows_ok(Hj, ldh) && rows_ok(dHi, ldd) && rows_ok(dHj, ldd),
              PFN_E_INVALID, "ea_bwd: node matrices must be 16-byte aligned with ld %% 4 == 0");
  PFN_REQUIRE(We && dWe && scratch, PFN_E_INVALID, "ea_bwd: null argument");
  const RowTiling t = make_tiling(n_nodes, h, kBlocksPerSm);
  float* partial = static_cast<float*>(scratch);
  int nblocks = 0;
  ProfScope prof(PFN_PROF_EA_BWD, stream);
  if (n_nodes > 0) {
    nblocks = t.nblocks;
    dim3 grid(t.nblocks, 2);
    PFN_CUDA_OK(launch_kernel(k_ea_bwd, grid, dim3(t.threads), 0, stream, dS, ldds, Hi, Hj, ldh, g.rowptr_t, g.nbr_t,
                                            reinterpret_cast<const float2*>(g.ea_t), g.rowptr_s, g.nbr_s,
                                            reinterpret_cast<const float2*>(g.ea_s), We, ldwe, dHi, dHj, ldd, partial,
                                            static_cast<int>(n_nodes), static_cast<int>(h), t.c4, t.cx, t.rows, t.npb));
    PFN_LAUNCHED();
  }
  PFN_CUDA_OK(launch_kernel(k_reduce_dwe, dim3(static_cast<int>(ceil_div64(2 * h * 32, 256))), dim3(256), 0, stream, partial, nblocks, t.c4, static_cast<int>(h),
                                                                           dWe, lddwe));
  PFN_LAUNCHED();
  return 0;
}

// the dWe reduction for up to 8 layers in one launch (blockIdx.y = layer)
namespace {
struct DweMulti {
  const float* partial[8];
  float* dwe[8];
  int lddwe[8];
};
__global__ void k_reduce_dwe_multi(const __grid_constant__ DweMulti a, int nblocks, int c4, int h) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * h) return;
  const int k = warp / h, c = warp - k * h;
  const float* src = a.partial[blockIdx.y] + (size_t(k) * 4 * c4 + c) * nblocks;
  float sum = 0.f;
  for (int b = lane; b < nblocks; b += 32) sum += src[b];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (lane == 0) a.dwe[blockIdx.y][c * a.lddwe[blockIdx.y] + k] = sum;
}
}  // namespace
int reduce_dwe_multi_launch(const float* const* partial, float* const* dwe, const int* lddwe, int n, int nblocks, int64_t h,
                            cudaStream_t stream) {
  for (int base = 0; base < n; base += 8) {
    DweMulti a{};
    const int cnt = std::min(8, n - base);
    for (int i = 0; i < cnt; ++i) {
      a.partial[i] = partial[base + i];
      a.dwe[i] = dwe[base + i];
      a.lddwe[i] = lddwe[base + i];
    }
    PFN_CUDA_OK(launch_kernel(k_reduce_dwe_multi, dim3(static_cast<unsigned>(ceil_div64(2 * h * 32, 256)), static_cast<unsigned>(cnt)),
                              dim3(256), 0, stream, a, nblocks, static_cast<int>((h + 3) / 4), static_cast<int>(h)));
    PFN_LAUNCHED();
  }
  return 0;
}

// dWe[c, k] = sum over `nblocks` per-CTA partial rows laid out [2][4 * ceil(h/4)][nblocks] (k_ea_bwd, or the tiles of the
// graph-resident EdgeAggregation backward)
int reduce_dwe_launch(const float* partial, int nblocks, int64_t h, float* dWe, int64_t lddwe, cudaStream_t stream) {
  PFN_CUDA_OK(launch_kernel(k_reduce_dwe, dim3(static_cast<int>(ceil_div64(2 * h * 32, 256))), dim3(256), 0, stream, partial, nblocks,
                            static_cast<int>((h + 3) / 4), static_cast<int>(h), dWe, lddwe));
  PFN_LAUNCHED();
  return 0;
}

int hop_launch(const float* X, int64_t ldx, const GraphView& g, int64_t n_nodes, bool transpose, const float* addend,
               int64_t ldadd, const float* ymask, int64_t ldym, float scale, float* Y, int64_t ldy, int64_t h,
               cudaStream_t stream) {
  PFN_REQUIRE(rows_ok(X, ldx) && rows_ok(Y, ldy), PFN_E_INVALID, "hop: node matrices must be 16-byte aligned with ld %% 4 == 0");
  PFN_REQUIRE(addend == nullptr || rows_ok(addend, ldadd), PFN_E_INVALID, "hop: addend misaligned");
  PFN_REQUIRE(ymask == nullptr || rows_ok(ymask, ldym), PFN_E_INVALID, "hop: ymask misaligned");
  if (n_nodes == 0) return 0;
  const RowTiling t = make_tiling(n_nodes, h, kBlocksPerSm);
  ProfScope prof(PFN_PROF_HOP, stream);
  PFN_CUDA_OK(launch_kernel(k_hop, dim3(t.nblocks), dim3(t.threads), 0, stream, X, ldx, transpose ? g.rowptr_s : g.rowptr_t, transpose ? g.nbr_s : g.nbr_t,
                                            g.dis, addend, ldadd, ymask, ldym, scale, Y, ldy, static_cast<int>(n_nodes),
                                            t.c4, t.cx, t.rows, t.npb));
  PFN_LAUNCHED();
  return 0;
}

}  // namespace pfn

using namespace pfn;

extern "C" int pfn_ea_fwd(const float* Hi, const float* Hj, int64_t ldh, const void* graph_ws, int64_t n_nodes,
                          int64_t e_raw, const float* We, int64_t ldwe, float* S, int64_t lds, int64_t h, void* stream) {
  PFN_REQUIRE(graph_ws != nullptr, PFN_E_INVALID, "pfn_ea_fwd: null graph workspace");
  return ea_fwd_launch(Hi, Hj, ldh, graph_view(graph_ws, n_nodes, e_raw), n_nodes, We, ldwe, S, lds, h,
                       static_cast<cudaStream_t>(stream));
}

extern "C" size_t pfn_ea_bwd_scratch_bytes(int64_t h) {
  return size_t(sm_count()) * kBlocksPerSm * 8 * size_t((h + 3) / 4) * sizeof(float);
}

extern "C" int pfn_ea_bwd(const float* dS, int64_t ldds, const float* Hi, const float* Hj, int64_t ldh,
                          const void* graph_ws, int64_t n_nodes, int64_t e_raw, const float* We, int64_t ldwe,
                          float* dHi, float* dHj, int64_t ldd, float* dWe, int64_t lddwe, void* scratch, int64_t h,
                          void* stream) {
  PFN_REQUIRE(graph_ws != nullptr, PFN_E_INVALID, "pfn_ea_bwd: null graph workspace");
  return ea_bwd_launch(dS, ldds, Hi, Hj, ldh, graph_view(graph_ws, n_nodes, e_raw), n_nodes, We, ldwe, dHi, dHj, ldd,
                       dWe, lddwe, scratch, h, static_cast<cudaStream_t>(stream));
}

extern "C" int pfn_spmm_hop(const float* X, int64_t ldx, const void* graph_ws, int64_t n_nodes, int64_t e_raw,
                            int transpose, const float* addend, int64_t ldadd, const float* ymask, int64_t ldym,
                            float scale, float* Y, int64_t ldy, int64_t h, void* stream) {
  PFN_REQUIRE(graph_ws != nullptr, PFN_E_INVALID, "pfn_spmm_hop: null graph workspace");
  return hop_launch(X, ldx, graph_view(graph_ws, n_nodes, e_raw), n_nodes, transpose != 0, addend, ldadd, ymask, ldym,
                    scale, Y, ldy, h, static_cast<cudaStream_t>(stream));
}
