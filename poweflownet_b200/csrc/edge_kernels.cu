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

constexpr int kSlabRows = 64;    // rows of CSR metadata a CTA stages at once (see stage_slab)
constexpr int kSlabEdges = 768;

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
  // large batches: cyclic slabs (npb = 0, see SlabWalk) once every CTA gets at least four of them
  static const int cyclic_env = env_int("PFN_EDGE_CYCLIC", -1);
  const bool cyclic = cyclic_env >= 0 ? cyclic_env != 0 : ceil_div64(n_nodes, kSlabRows) >= int64_t(4) * t.nblocks;
  if (cyclic) t.npb = 0;
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

// ---- the same message arithmetic on packed pairs (Blackwell FADD2 / FFMA2: two IEEE fp32 operations per instruction,
// identical results): 12 instructions per 16-byte chunk and edge instead of 20 -- the bulk-copy kernel below is bound by
// instruction issue once its copies flow (5.4 bytes per warp instruction measured), not by arithmetic throughput.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
struct Pair4 {
  uint64_t lo, hi;  // (x, y), (z, w)
};
__device__ __forceinline__ Pair4 pair4(float4 v) { return Pair4{pack2(v.x, v.y), pack2(v.z, v.w)}; }
// acc += ReLU(hi + hj + a.x * w0 + a.y * w1), the operation order of preact() / add_relu()
__device__ __forceinline__ void add_relu_preact2(Pair4& acc, const Pair4& hi, float4 hj, float2 a, const Pair4& w0, const Pair4& w1) {
  const uint64_t ax = pack2(a.x, a.x), ay = pack2(a.y, a.y);
  const uint64_t p01 = fma2(ay, w1.lo, fma2(ax, w0.lo, add2(hi.lo, pack2(hj.x, hj.y))));
  const uint64_t p23 = fma2(ay, w1.hi, fma2(ax, w0.hi, add2(hi.hi, pack2(hj.z, hj.w))));
  float p0, p1, p2, p3;
  unpack2(p01, p0, p1);
  unpack2(p23, p2, p3);
  acc.lo = add2(acc.lo, pack2(fmaxf(p0, 0.f), fmaxf(p1, 0.f)));
  acc.hi = add2(acc.hi, pack2(fmaxf(p2, 0.f), fmaxf(p3, 0.f)));
}

// ---- CSR slab staging ---------------------------------------------------------------------------------
// A CTA owns a contiguous slab of rows, so its CSR entries (rowptr slice, neighbour ids, edge_attr) are one contiguous
// range: they are copied into shared memory once, cooperatively and coalesced.  Without this every thread walks a
// chain of three dependent global loads per row (rowptr -> neighbour id -> neighbour row); with it the per-row work
// is a single round of independent gathers.  Slabs larger than the staging buffers are walked in pieces; a piece
// whose edge count exceeds the buffer (a hub bus) falls back to reading the CSR arrays from global memory.

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

// How a CTA walks its slabs of kSlabRows rows.  npb > 0: one contiguous range of npb rows per CTA (small batches: every
// CTA gets work).  npb == 0: slab s belongs to CTA s mod gridDim (large batches).  With contiguous ranges the CTAs of a
// large batch work in ~150 distant regions at once, i.e. on every graph of the batch simultaneously, and the gathered
// neighbour rows -- each needed ~3 times, always from the same graph -- fall out of L2 between uses: measured at
// case6470rte x 32, hidden 512: 1.47 GB read from DRAM for 0.86 GB of operands, L2 hit rate 4.6 %.  Cyclic slabs keep the
// whole grid inside a window of gridDim x 64 rows (a graph or two), whose neighbour rows stay L2-resident.
struct SlabWalk {
  int start, end, step;
};
__device__ __forceinline__ SlabWalk slab_walk(int npb, int n_nodes) {
  SlabWalk w;
  if (npb > 0) {
    w.start = blockIdx.x * npb;
    w.end = min(n_nodes, w.start + npb);
    w.step = kSlabRows;
  } else {
    w.start = blockIdx.x * kSlabRows;
    w.end = n_nodes;
    w.step = gridDim.x * kSlabRows;
  }
  return w;
}

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
  const SlabWalk walk = slab_walk(npb, n_nodes);
  for (int r0 = walk.start; r0 < walk.end; r0 += walk.step) {  // CTA-uniform
    const int nr = min(kSlabRows, walk.end - r0);
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
// k_ea_fwd_tma: the same message + aggregate as k_ea_fwd, fed by ASYNCHRONOUS BULK COPIES INTO SHARED MEMORY, so that
// the bytes in flight are bounded by shared memory (~200 KB per SM) instead of by registers.  k_ea_fwd keeps five 16-byte
// row loads per thread in flight and walks "stage the CSR slab -> barrier -> one row per thread -> next row" in lock
// step (ncu: long_scoreboard 9.5 + barrier 5.0 stalled warps per issue, DRAM 13 % of peak); a first attempt with
// autonomous warps (each warp its own copy ring, 8 warps per SM) moved the bytes but was issue-bound: 7.0 M warp
// instructions at 0.4 eligible warps per cycle (profiles/r2/ncu_ea_fwd_earlier_designs.txt, profiles/r2_summary.md).  This kernel separates the two jobs:
//
//  * PRODUCER WARPS (8 by default; one persistent CTA per SM) walk the CSR and issue, per batch of R whole rows: ONE bulk
//    copy (cp.async.bulk, SASS UBLKCP) for the R consecutive Hi rows, and the gathered Hj rows as 16-byte cp.async chunks
//    (LDGSTS; every lane a fixed column chunk, four copies per trip) into a ring of n_stages batch buffers; completion is
//    counted on the batch's "full" mbarrier (complete_tx bytes for the bulk copy, cp.async.mbarrier.arrive.noinc for the
//    chunk copies).  A bulk copy per gathered row -- the first version -- is bound by the TMA unit's ~35-77 ns per
//    operation whatever its size (14.8 us of copies at case118v2 x 128); it is used for rows of >= 8 KB only
//    (PFN_EA_BULK8 = how many of every eight gathered rows go that way).  Producer 0 also writes the batch header (row
//    pointers relative to the batch, edge_attr in edge order), so the consumers never touch the CSR arrays.
//  * CONSUMER WARPS (16) run k_ea_fwd's inner loop -- thread (x, y) owns 16-byte column chunk x (its slice of We stays in
//    registers) and sums rows y, y + rows, ... of the batch in ascending edge id -- on packed pairs (fma.rn.f32x2 /
//    add.rn.f32x2, half the issue slots, same rounding), every operand from shared memory: the two kernels are
//    BIT-IDENTICAL (tests/test_gpu_kernels.py).  A consumer warp releases the buffer through the batch's "empty" mbarrier.
//  * The row pointers of the CTA's range and the first window of neighbour ids / edge_attr are fetched by ALL threads
//    before the roles split (two wide coalesced round trips instead of two narrow ones).
//  * Large problems: CTA c takes the row chunks c, c + grid, ... (kTmaChunkRows rows each), so that all CTAs work inside
//    one moving window whose gathered rows stay in L2 (DRAM reads 1.47 -> 0.93 GB at case6470rte x 32 x h512).
//  * Optional (PFN_EA_PREFETCH=1; measured ~10 % slower, off): L2 prefetch of the CTA's share of Hi, Hj and the CSR.
//  A row with more incident edges than a batch buffer holds (a hub bus) is summed from global memory by the consumers.
constexpr int kTmaMaxStages = 4;
constexpr int kTmaMetaRows = 2048;   // row pointers staged at once (a CTA's whole range unless the batch is huge)
constexpr int kTmaMetaEdges = 1024;  // producer-private window of neighbour ids / edge_attr
constexpr int kTmaMaxBatchRows = 64;
constexpr int kTmaChunkRows = 128;   // rows per cyclic chunk of a large batch

struct TmaBatchHeader {
  int R, row0, direct, pad;
  int rp[kTmaMaxBatchRows + 4];  // edge offsets relative to the batch's first edge; rp[R] = edges of the batch
  // float2 ea[cap_slots] follows
};

struct TmaArgs {
  const float* Hi;
  const float* Hj;
  const int* rowptr;
  const int* nbr;
  const float2* ea;
  const float* We;
  float* S;
  long long ldh, lds, ldwe, e_cap;
  int n_nodes, h, c4, rows, n_stages, cap_slots, rows_per_cta, cyclic, prefetch_l2, contiguous, cons_warps, prod_warps, bulk_eighths, round_rows,
      meta_rows, meta_edges;  // capacity of the staged row-pointer slice / neighbour window
  unsigned stage_bytes, hdr_bytes;
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
__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"  // suspend-time hint: sleep in hardware, do not spin
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  }
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// the CTA's 1/gridDim share of [base, base + bytes) -> L2 as ONE bulk prefetch (16-byte granules).  One operation, not many
// small ones: the TMA unit needs ~35 ns per operation whatever its size, and the batch copies queue behind these.
__device__ __forceinline__ void prefetch_share(const void* base, unsigned bytes) {
  const unsigned gran = (bytes + 15u) >> 4, per = (gran + gridDim.x - 1) / gridDim.x;
  const unsigned lo = min(gran, per * blockIdx.x), hi = min(gran, lo + per);
  if (hi > lo) bulk_prefetch_l2(static_cast<const char*>(base) + (size_t(lo) << 4), (hi - lo) << 4);
}

__global__ void __launch_bounds__(1024, 1) k_ea_fwd_tma(const __grid_constant__ TmaArgs a) {
  extern __shared__ __align__(128) uint8_t tma_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c4 = a.c4, NS = a.n_stages, cap = a.cap_slots;
  const uint32_t rowbytes = uint32_t(c4) * 16u;
  // shared-memory map: barriers full[4] empty[4] | rp [meta_rows + 4] | nb [meta_edges] | ea [meta_edges] | stages
  const int kMetaRows = a.meta_rows, kMetaEdges = a.meta_edges;
  int* m_rp = reinterpret_cast<int*>(tma_smem + 128);
  int* m_nb = m_rp + kMetaRows + 4;
  float2* m_ea = reinterpret_cast<float2*>(m_nb + kMetaEdges);
  const uint32_t smem0 = tma_smem_u32(tma_smem);
  const uint32_t off_stage = (128u + uint32_t(kMetaRows + 4) * 4u + uint32_t(kMetaEdges) * 12u + 127u) & ~127u;
  auto full_bar = [&](int s) { return smem0 + 8u * s; };
  auto empty_bar = [&](int s) { return smem0 + 8u * (kTmaMaxStages + s); };
  // Row ranges of this CTA: ONE contiguous range of rows_per_cta rows (small batches), or, cyclic, every gridDim-th chunk of
  // rows_per_cta rows (large batches: the whole grid then works inside a window of a graph or two whose neighbour rows
  // stay L2-resident -- see SlabWalk).
  const int row_start = min(a.n_nodes, int(blockIdx.x) * a.rows_per_cta), row_end = min(a.n_nodes, row_start + a.rows_per_cta);
  const long long range_step = a.cyclic ? (long long)gridDim.x * a.rows_per_cta : (long long)a.n_nodes + 1;

  // ---- prologue that may overlap the previous kernel's tail (nothing here reads data as a value) ----
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      tma_mbar_init(full_bar(s), 1u + (a.bulk_eighths >= 8 ? 0u : 32u * uint32_t(a.prod_warps)));  // expect_tx + one arrival per cp.async lane
      tma_mbar_init(empty_bar(s), uint32_t(a.cons_warps));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (a.prefetch_l2 && tid < 5) {
    // thread 0: the CTA's share of Hj (every row of it is fetched from HBM exactly once, by whoever comes first);
    // thread 1: its own Hi rows; threads 2..4: its shares of the CSR arrays
    if (tid == 0 && a.contiguous) prefetch_share(a.Hj, unsigned(a.n_nodes) * rowbytes);
    if (tid == 1 && a.contiguous && row_end > row_start)
      bulk_prefetch_l2(reinterpret_cast<const char*>(a.Hi) + size_t(row_start) * rowbytes, unsigned(row_end - row_start) * rowbytes);
    if (tid == 2) prefetch_share(a.nbr, unsigned(a.e_cap) * 4u);
    if (tid == 3) prefetch_share(a.ea, unsigned(a.e_cap) * 8u);
    if (tid == 4) prefetch_share(a.rowptr, (unsigned(a.n_nodes) + 1u) * 4u);
  }
  pdl_wait();
  // ---- all threads: row pointers of the CTA's range, then the first window of neighbour ids / edge_attr ----
  int meta_row0 = row_start, meta_rows = min(kMetaRows, row_end - row_start);
  for (int i = tid; i <= meta_rows; i += blockDim.x) m_rp[i] = a.rowptr[row_start + i];
  __syncthreads();
  int ew_lo = m_rp[0];
  long long ew_hi = min((long long)a.e_cap, (long long)ew_lo + kMetaEdges);
  for (int i = tid; i < int(ew_hi - ew_lo); i += blockDim.x) {
    m_nb[i] = a.nbr[ew_lo + i];
    m_ea[i] = a.ea[ew_lo + i];
  }
  __syncthreads();

  if (warp >= a.cons_warps) {
    // =================================== producer warps ===================================
    // All `P` producer warps walk the same batch sequence (the greedy row count is recomputed by each: a ballot over
    // shared row pointers).  Producer 0 writes the batch header and issues the Hi bulk copy; the gathered rows are
    // dealt to the 32 P producer lanes in 16-byte chunks (consecutive lanes = consecutive chunks of a row: coalesced)
    // or, for rows of >= kTmaBulkRowBytes, as one bulk copy per row.
    const int P = a.prod_warps, pw = warp - a.cons_warps, pt = pw * 32 + lane, stride = 32 * P;
    const int step_slot = stride / c4, step_q = stride % c4;
    auto producers_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(stride) : "memory"); };
    unsigned batch = 0;
    for (long long rs = row_start; rs < a.n_nodes; rs += range_step) {
    const int range_end = min(a.n_nodes, int(rs) + a.rows_per_cta);
    int r = int(rs);
    if (r != row_start) meta_rows = 0;  // a later range: its row pointers are fetched by the producers (first trip below)
    while (r < range_end) {
      if (r >= meta_row0 + meta_rows) {  // next slice of row pointers (later ranges; ranges longer than kTmaMetaRows rows)
        meta_row0 = r;
        meta_rows = min(kMetaRows, range_end - r);
        producers_sync();
        for (int i = pt; i <= meta_rows; i += stride) m_rp[i] = a.rowptr[r + i];
        producers_sync();
      }
      const int lr = r - meta_row0;
      // largest R <= 64 with R + (edges of rows [r, r + R)) <= cap slots; 0 = row r alone does not fit (hub bus)
      int R;
      {
        const int avail = meta_rows - lr, c1 = lane + 1, c2 = lane + 33;
        const bool ok1 = c1 <= avail && c1 + (m_rp[lr + c1] - m_rp[lr]) <= cap;
        const bool ok2 = c2 <= avail && c2 + (m_rp[lr + c2] - m_rp[lr]) <= cap;
        R = __popc(__ballot_sync(0xffffffffu, ok1)) + __popc(__ballot_sync(0xffffffffu, ok2));
      }
      if (a.round_rows && R > a.rows) R = R / a.rows * a.rows;  // whole passes of the consumers' row lanes
      const bool direct = R == 0;
      if (direct) R = 1;
      const int e0 = m_rp[lr], nE = direct ? 0 : m_rp[lr + R] - e0;
      if (!direct && (long long)e0 + nE > ew_hi) {  // slide the producers' window of neighbour ids / edge_attr
        ew_lo = e0;
        ew_hi = min((long long)a.e_cap, (long long)ew_lo + kMetaEdges);
        producers_sync();
        for (int i = pt; i < int(ew_hi - ew_lo); i += stride) {
          m_nb[i] = a.nbr[ew_lo + i];
          m_ea[i] = a.ea[ew_lo + i];
        }
        producers_sync();
      }
      const int s = int(batch % unsigned(NS));
      tma_mbar_wait(empty_bar(s), ((batch / unsigned(NS)) & 1u) ^ 1u);  // the consumers have released this buffer
      const uint32_t buf = smem0 + off_stage + uint32_t(s) * a.stage_bytes + a.hdr_bytes;
      const float* hi_src = a.Hi + (size_t)r * a.ldh;
      // Gathered rows travel through BOTH copy engines at once: the first b8/8 of a batch's slots go as one bulk copy each
      // (TMA unit: ~35-77 ns per copy whatever its size), the rest as 16-byte cp.async chunks (LSU / L1 miss queue:
      // bounded number of lines in flight).  Each engine alone topped out near 25 KB/us per SM on 2 KB rows.
      const int b8 = a.bulk_eighths;
      const int n_bulk = (nE * b8 + 4) >> 3, n_async = nE - n_bulk;
      if (pw == 0) {
        uint8_t* stage = tma_smem + off_stage + size_t(s) * a.stage_bytes;
        TmaBatchHeader* hdr = reinterpret_cast<TmaBatchHeader*>(stage);
        float2* hdr_ea = reinterpret_cast<float2*>(stage + sizeof(TmaBatchHeader));
        if (lane == 0) {
          hdr->R = R;
          hdr->row0 = r;
          hdr->direct = direct ? 1 : 0;
        }
        for (int i = lane; i <= R; i += 32) hdr->rp[i] = direct ? 0 : m_rp[lr + i] - e0;
        for (int j = lane; j < nE; j += 32) hdr_ea[j] = m_ea[e0 - ew_lo + j];
        __syncwarp();
        if (lane == 0) {
          // bytes that arrive by bulk copy: the Hi rows (always) and the bulk share of the gathered rows
          tma_mbar_expect_tx(full_bar(s), direct ? 0u : uint32_t(R + n_bulk) * rowbytes);
          if (!direct && a.contiguous) bulk_g2s(buf, hi_src, uint32_t(R) * rowbytes, full_bar(s));  // R consecutive rows: one copy
        }
        __syncwarp();
        if (!direct && !a.contiguous)
          for (int i = lane; i < R; i += 32) bulk_g2s(buf + uint32_t(i) * rowbytes, hi_src + (size_t)i * a.ldh, rowbytes, full_bar(s));
      }
      if (!direct) {
        const uint32_t gat = buf + uint32_t(R) * rowbytes;
        const int* nb = m_nb + (e0 - ew_lo);
        // bulk share: slots [0, n_bulk), dealt to all producer lanes
        for (int j = pt; j < n_bulk; j += stride)
          bulk_g2s(gat + uint32_t(j) * rowbytes, a.Hj + (size_t)nb[j] * a.ldh, rowbytes, full_bar(s));
        if (n_async > 0) {
          // cp.async share: slots [n_bulk, nE).  Four chunks per lane and trip: the neighbour-id reads and the address
          // arithmetic of a trip are independent, so the shared-memory latency is paid once per four copies (one chunk per
          // trip measured 88 cycles per copy and made the kernel producer-bound; so did a division in this loop)
          int k = n_bulk + pt / c4, q = pt % c4;
          if (step_q == 0) {
            // every lane keeps its chunk column (the producer lanes cover whole rows): one neighbour-id read, one 64-bit
            // multiply-add and the copy per chunk -- ~6 instructions instead of ~26 for the general walk below
            const char* col = reinterpret_cast<const char*>(a.Hj) + size_t(q) * 16u;
            const size_t pitch = size_t(a.ldh) * 4u;
            uint32_t dst = gat + uint32_t(k) * rowbytes + uint32_t(q) * 16u;
            const uint32_t dst_step = uint32_t(step_slot) * rowbytes;
            for (; k + 3 * step_slot < nE; k += 4 * step_slot, dst += 4u * dst_step) {
              const int n0 = nb[k], n1 = nb[k + step_slot], n2 = nb[k + 2 * step_slot], n3 = nb[k + 3 * step_slot];
              cp_async16(dst, col + size_t(n0) * pitch);
              cp_async16(dst + dst_step, col + size_t(n1) * pitch);
              cp_async16(dst + 2u * dst_step, col + size_t(n2) * pitch);
              cp_async16(dst + 3u * dst_step, col + size_t(n3) * pitch);
            }
            for (; k < nE; k += step_slot, dst += dst_step) cp_async16(dst, col + size_t(nb[k]) * pitch);
          }
          while (k < nE) {
            int sl[4], qq[4], nbv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              sl[u] = k;
              qq[u] = q;
              k += step_slot;
              q += step_q;
              if (q >= c4) {
                q -= c4;
                ++k;
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) nbv[u] = nb[min(sl[u], nE - 1)];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (sl[u] < nE)
                cp_async16(gat + uint32_t(sl[u]) * rowbytes + uint32_t(qq[u]) * 16u, a.Hj + (size_t)nbv[u] * a.ldh + 4 * qq[u]);
          }
        }
      }
      if (b8 < 8) cp_async_arrive_noinc(full_bar(s));  // fires when this lane's cp.async copies of the batch have landed
      r += R;
      ++batch;
    }
    }
  } else if (warp < a.cons_warps) {
    // =================================== consumer warps ===================================
    const int x = tid % c4, y = tid / c4;
    const bool active = y < a.rows;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
    if (active) load_we(a.We, a.ldwe, x, a.h, w0, w1);
    const Pair4 w0p = pair4(w0), w1p = pair4(w1);
    unsigned batch = 0;
    for (long long rs = row_start; rs < a.n_nodes; rs += range_step) {
    const int range_end = min(a.n_nodes, int(rs) + a.rows_per_cta);
    int r = int(rs);
    while (r < range_end) {
      const int s = int(batch % unsigned(NS));
      tma_mbar_wait(full_bar(s), (batch / unsigned(NS)) & 1u);
      const uint32_t stage = smem0 + off_stage + uint32_t(s) * a.stage_bytes;
      const TmaBatchHeader* hdr = reinterpret_cast<const TmaBatchHeader*>(tma_smem + off_stage + size_t(s) * a.stage_bytes);
      const int R = hdr->R;
      if (hdr->direct) {
        if (active && y == 0) {  // hub bus: the row's operands straight from global memory, same order of operations
          const int beg = a.rowptr[r], fin = a.rowptr[r + 1];
          const float4 hi = ld4(a.Hi + (size_t)r * a.ldh + 4 * x);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int e = beg; e < fin; ++e)
            add_relu(acc, preact(hi, ldg4(a.Hj + (size_t)a.nbr[e] * a.ldh + 4 * x), a.ea[e], w0, w1));
          st4(a.S + (size_t)r * a.lds + 4 * x, acc);
        }
      } else if (active) {
        const uint32_t ea_s = stage + uint32_t(sizeof(TmaBatchHeader));
        const uint32_t hi_s = stage + a.hdr_bytes + uint32_t(x) * 16u, gat = hi_s + uint32_t(R) * rowbytes;
        for (int lr = y; lr < R; lr += a.rows) {
          const Pair4 hi = pair4(lds4(hi_s + uint32_t(lr) * rowbytes));
          const int beg = hdr->rp[lr], fin = hdr->rp[lr + 1];
          Pair4 acc{0ull, 0ull};
          int e = beg;
          for (; e + 3 < fin; e += 4) {  // four shared-memory row reads in flight per thread
            const float4 h0 = lds4(gat + uint32_t(e) * rowbytes), h1 = lds4(gat + uint32_t(e + 1) * rowbytes);
            const float4 h2 = lds4(gat + uint32_t(e + 2) * rowbytes), h3 = lds4(gat + uint32_t(e + 3) * rowbytes);
            add_relu_preact2(acc, hi, h0, lds2(ea_s + 8u * e), w0p, w1p);
            add_relu_preact2(acc, hi, h1, lds2(ea_s + 8u * (e + 1)), w0p, w1p);
            add_relu_preact2(acc, hi, h2, lds2(ea_s + 8u * (e + 2)), w0p, w1p);
            add_relu_preact2(acc, hi, h3, lds2(ea_s + 8u * (e + 3)), w0p, w1p);
          }
          if (e + 1 < fin) {
            const float4 h0 = lds4(gat + uint32_t(e) * rowbytes), h1 = lds4(gat + uint32_t(e + 1) * rowbytes);
            add_relu_preact2(acc, hi, h0, lds2(ea_s + 8u * e), w0p, w1p);
            add_relu_preact2(acc, hi, h1, lds2(ea_s + 8u * (e + 1)), w0p, w1p);
            e += 2;
          }
          if (e < fin) add_relu_preact2(acc, hi, lds4(gat + uint32_t(e) * rowbytes), lds2(ea_s + 8u * e), w0p, w1p);
          float4 o;
          unpack2(acc.lo, o.x, o.y);
          unpack2(acc.hi, o.z, o.w);
          st4(a.S + (size_t)(r + lr) * a.lds + 4 * x, o);
        }
      }
      __syncwarp();  // every lane has finished reading the buffer
      if (lane == 0) tma_mbar_arrive(empty_bar(s));
      r += R;
      ++batch;
    }
    }
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
  const SlabWalk walk = slab_walk(npb, n_nodes);
  const bool target_side = blockIdx.y == 0;
  const int nq = (c4 + cx - 1) / cx;  // column passes (1 unless hidden_dim > 1024)
  float4 g0[1] = {make_float4(0.f, 0.f, 0.f, 0.f)}, g1[1] = {make_float4(0.f, 0.f, 0.f, 0.f)};
  for (int qi = 0; qi < nq; ++qi) {  // CTA-uniform loops: the body holds barriers
    const int q = qi * cx + x;
    const bool active = q < c4;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
    if (active) load_we(We, ldwe, q, h, w0, w1);
    g0[0] = g1[0] = make_float4(0.f, 0.f, 0.f, 0.f);  // dWe[:,0], dWe[:,1] partial sums of this chunk
    for (int r0 = walk.start; r0 < walk.end; r0 += walk.step) {
      const int nr = min(kSlabRows, walk.end - r0);
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
  const SlabWalk walk = slab_walk(npb, n_nodes);
  for (int r0 = walk.start; r0 < walk.end; r0 += walk.step) {
    const int nr = min(kSlabRows, walk.end - r0);
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

// ---- host side of k_ea_fwd_tma ----------------------------------------------------------------------------------------
// PFN_EA_FWD=cta forces the CTA-slab kernel (k_ea_fwd); unset / anything else takes the bulk-copy kernel whenever the
// width fits.  Read per call so that tests can compare the two.  Tuning knobs (experiments): PFN_EA_STAGES (2..4),
// PFN_EA_THREADS (consumer threads, default 512), PFN_EA_PRODUCERS (1..8), PFN_EA_BULK (0/1), PFN_EA_PREFETCH (0/1).
constexpr uint32_t kTmaSmemLimit = 227 * 1024;
constexpr long long kTmaPrefetchMaxBytes = 48ll << 20;  // L2-prefetch both node matrices only when they fit well inside L2

bool ea_fwd_use_tma() {
  const char* e = std::getenv("PFN_EA_FWD");
  return !(e != nullptr && e[0] == 'c');
}

// 0 = launched, 1 = shape outside this kernel (caller uses k_ea_fwd)
int ea_fwd_tma_launch(const float* Hi, const float* Hj, int64_t ldh, const GraphView& g, int64_t n_nodes, const float* We,
                      int64_t ldwe, float* S, int64_t lds, int64_t h, cudaStream_t stream) {
  const int c4 = static_cast<int>((h + 3) / 4);
  const uint32_t rowbytes = uint32_t(c4) * 16u;
  if (n_nodes >= (int64_t(1) << 31) - 64 || c4 > 31 * 32 || g.e_cap >= (int64_t(1) << 28)) return 1;
  constexpr uint32_t kTmaBulkRowBytes = 8192;
  TmaArgs a{};
  // gathered rows: how many of every eight go as one bulk copy each, the rest as 16-byte cp.async chunks (both engines run
  // concurrently; PFN_EA_BULK8 = 0..8 overrides, PFN_EA_BULK = 0/1 means 0/8)
  const int env_bulk = env_int("PFN_EA_BULK", -1);
  int b8 = rowbytes >= kTmaBulkRowBytes ? 8 : 0;  // (measured at 2 KB rows: 0/8 258 us, 2/8 269 us, 4/8 276 us, 8/8 357 us)
  if (env_bulk >= 0) b8 = env_bulk != 0 ? 8 : 0;
  a.bulk_eighths = std::max(0, std::min(8, env_int("PFN_EA_BULK8", b8)));
  // CTAs per SM: one CTA owning all of the SM's shared memory (default), or two half-sized CTAs per SM whose latency chains
  // (row pointers -> neighbour ids -> rows) overlap each other (PFN_EA_CTAS_PER_SM=2; measured at case118v2 x 128: 10.4 us
  // against 10.1 us -- no gain, it stays an experiment knob).
  const int per_sm = std::max(1, std::min(2, env_int("PFN_EA_CTAS_PER_SM", 1)));
  // producer warps: 8, or 12 for wide rows whose chunk count divides 12 warps' lanes (every lane then keeps ONE column chunk
  // for the whole kernel, the cheap addressing path; measured at 2 KB rows: 8 -> 254 us, 12 -> 247 us, 10 / 14 -> 282 us)
  const int prod_default = a.bulk_eighths >= 8 ? 1 : (per_sm == 2 ? 4 : ((c4 >= 128 && 384 % c4 == 0) ? 12 : 8));
  a.prod_warps = std::max(1, std::min(16, env_int("PFN_EA_PRODUCERS", prod_default)));
  const int max_cons = (32 / per_sm - a.prod_warps) * 32;
  if (c4 > max_cons) return 1;
  const int want_threads = std::max(c4, std::min(max_cons, env_int("PFN_EA_THREADS", per_sm == 2 ? 256 : 512)));
  a.rows = std::max(1, std::min(kTmaMaxBatchRows, want_threads / c4));
  a.cons_warps = (c4 * a.rows + 31) / 32;
  a.meta_rows = per_sm == 2 ? 512 : kTmaMetaRows;
  a.meta_edges = per_sm == 2 ? 512 : kTmaMetaEdges;
  const uint32_t smem_limit = per_sm == 2 ? 113u * 1024u : kTmaSmemLimit;  // 2 x (113 KB + 1 KB reserved) <= 228 KB per SM
  int stages = std::max(1, std::min(kTmaMaxStages, env_int("PFN_EA_STAGES", 2)));  // measured: two large batch buffers beat three or four
  const uint32_t off_stage = (128u + uint32_t(a.meta_rows + 4) * 4u + uint32_t(a.meta_edges) * 12u + 127u) & ~127u;
  uint32_t stage_bytes = 0, hdr_bytes = 0;
  int cap = 0;
  for (;; --stages) {
    stage_bytes = ((smem_limit - 128u - off_stage) / uint32_t(stages)) & ~127u;
    // header: fixed part + one float2 per slot (an upper bound of the edges of a batch), rounded to 128 bytes
    cap = static_cast<int>((stage_bytes - sizeof(TmaBatchHeader) - 128u) / (rowbytes + 8u));
    cap = std::min(cap, a.meta_edges);
    hdr_bytes = (uint32_t(sizeof(TmaBatchHeader)) + 8u * uint32_t(std::max(cap, 0)) + 127u) & ~127u;
    // a batch buffer should hold at least one pass of the consumers' row lanes with their edges (~4 slots per row)
    if (cap >= std::min(4 * a.rows, 64) || stages <= 2) break;
  }
  if (cap < 2) return 1;
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
  a.e_cap = g.e_cap;
  a.n_nodes = static_cast<int>(n_nodes);
  a.h = static_cast<int>(h);
  a.c4 = c4;
  a.n_stages = stages;
  a.cap_slots = cap;
  a.stage_bytes = stage_bytes;
  a.hdr_bytes = hdr_bytes;
  a.contiguous = (ldh * 4 == int64_t(rowbytes)) ? 1 : 0;
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(int64_t(sm_count()) * per_sm, ceil_div64(n_nodes, a.rows))));
  a.rows_per_cta = static_cast<int>(ceil_div64(n_nodes, grid));
  // large batches: cyclic chunks of kTmaChunkRows rows once every CTA gets at least four of them (L2 residency of the
  // gathered rows, see SlabWalk); PFN_EA_CHUNK = rows per chunk (0 = contiguous ranges)
  const int chunk = env_int("PFN_EA_CHUNK", n_nodes >= int64_t(8) * grid * kTmaChunkRows ? kTmaChunkRows : 0);
  if (chunk > 0) {
    a.rows_per_cta = chunk;
    a.cyclic = 1;
  }
  // (measured at case118v2 x 128: the prefetch makes the kernel ~10 % SLOWER -- it stays an opt-in experiment)
  a.round_rows = env_int("PFN_EA_ROUND", 0) != 0 ? 1 : 0;
  a.prefetch_l2 = (env_int("PFN_EA_PREFETCH", 0) != 0 && 2 * n_nodes * int64_t(rowbytes) <= kTmaPrefetchMaxBytes) ? 1 : 0;
  const uint32_t smem = off_stage + uint32_t(stages) * stage_bytes + 128u;
  static SmemAttrOnce attr_once;
  PFN_CUDA_OK(ensure_dynamic_smem(attr_once, [] { return cudaFuncSetAttribute(k_ea_fwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTmaSmemLimit)); }));
  PFN_CUDA_OK(launch_kernel(k_ea_fwd_tma, dim3(grid), dim3(32 * (a.cons_warps + a.prod_warps)), smem, stream, a));
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
  if (ea_fwd_use_tma()) {
    const int rc = ea_fwd_tma_launch(Hi, Hj, ldh, g, n_nodes, We, ldwe, S, lds, h, stream);
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
