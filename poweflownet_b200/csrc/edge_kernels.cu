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
// k_ea_fwd_warp: the same message + aggregate with a WARP owning two consecutive rows at a time (hidden widths of
// 24..33 float4 columns, i.e. hidden_dim 93..132: standard.json's 129).  k_ea_fwd above moves a CTA in lock step through
// "stage the CSR slab -> barrier -> one row per thread -> next row" and keeps ~5 row loads per thread in flight; its
// SMs are busy only two thirds of the kernel's duration (ncu: sm__cycles_active / elapsed = 0.67, 5 barrier-stalled
// warps per issue).  Here nothing is shared between warps after the first barrier (We in shared memory) and the warps
// are persistent and software-pipelined: while the neighbour rows of row pair i are in flight, the neighbour ids /
// edge_attr of pair i+1 and the row pointers of pair i+2 are already being fetched, so the dependent chain
// rowptr -> neighbour id -> neighbour row is paid once per warp, not once per row.  Lanes 0..2 read the row pointers;
// the lanes read the pair's neighbour ids / edge_attr in one coalesced load each and broadcast them by shuffle; every
// lane keeps its float4 column of BOTH rows and up to eight gathered neighbour rows in flight at once.  Column 32
// (floats 128..131, the odd 129th channel) would cost a second pass with one active lane: instead lane u gathers that
// chunk for edge u of the pair and the per-row sums are taken in edge order through shuffles.  Per-row results are
// bit-identical to k_ea_fwd (same FMAs, same ascending-edge summation order).
constexpr int kEaWarpChunk = 8;
constexpr int kEaWarpThreads = 128;
constexpr bool kEaFwdWarpDefault = false;  // measured on the B200: 19.5 us against 12.6 us for k_ea_fwd at case118v2 x 128
                                           // (profiles/r1_ea_fwd_warp_vs_cta.json) -- opt-in (PFN_EA_FWD=warp) until it wins

__device__ __forceinline__ float4 shfl4(float4 v, int src) {
  return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                     __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}

struct RowPair {
  int r0, nrow, e0, deg0, ne;  // first row, rows (1 or 2), first edge, edges of row 0, edges of the pair
};

// lanes 0..nrow of a live pair read its row pointers; 0 for pairs past the end
__device__ __forceinline__ int pair_rowptr(const int* __restrict__ rowptr, int pair, int npairs, int n_nodes, int lane) {
  if (pair >= npairs) return 0;
  const int r0 = 2 * pair, nrow = min(2, n_nodes - r0);
  return lane <= nrow ? __ldg(rowptr + r0 + lane) : 0;
}

__device__ __forceinline__ RowPair pair_derive(int rp, int pair, int n_nodes) {
  RowPair p;
  p.r0 = 2 * pair;
  p.nrow = min(2, n_nodes - p.r0);
  const int e0 = __shfl_sync(0xffffffffu, rp, 0), e1 = __shfl_sync(0xffffffffu, rp, 1), e2 = __shfl_sync(0xffffffffu, rp, p.nrow);
  p.e0 = e0;
  p.deg0 = e1 - e0;
  p.ne = e2 - e0;
  return p;
}

__global__ void __launch_bounds__(kEaWarpThreads)
k_ea_fwd_warp(const float* __restrict__ Hi, const float* __restrict__ Hj, int64_t ldh, const int* __restrict__ rowptr,
              const int* __restrict__ nbr, const float2* __restrict__ ea, const float* __restrict__ We, int64_t ldwe,
              float* __restrict__ S, int64_t lds, int n_nodes, int h, int c4) {
  pdl_wait();
  __shared__ float4 s_w[33][2];
  constexpr int kWarps = kEaWarpThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npairs = (n_nodes + 1) >> 1, stride = gridDim.x * kWarps;
  int pair = blockIdx.x * kWarps + warp;
  const int rp_cur = pair_rowptr(rowptr, pair, npairs, n_nodes, lane);  // in flight while We is staged
  int rp_nxt = pair_rowptr(rowptr, pair + stride, npairs, n_nodes, lane);
  if (threadIdx.x < 2 * c4) {
    const int q = threadIdx.x >> 1, k = threadIdx.x & 1;
    float w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ch = 4 * q + c;
      w[c] = ch < h ? __ldg(We + ch * ldwe + k) : 0.f;
    }
    s_w[q][k] = make_float4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  if (pair >= npairs) return;  // surplus warp
  const bool act = lane < min(c4, 32);
  const bool tail = c4 > 32;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w0 = zero, w1 = zero;
  if (act) {
    w0 = s_w[lane][0];
    w1 = s_w[lane][1];
  }
  RowPair cur = pair_derive(rp_cur, pair, n_nodes);
  int my_nbr = 0;  // neighbour id / edge_attr of edge `lane` of the current pair (its first 32 edges)
  float2 my_ea = make_float2(0.f, 0.f);
  if (lane < min(32, cur.ne)) {
    my_nbr = __ldg(nbr + cur.e0 + lane);
    my_ea = __ldg(ea + cur.e0 + lane);
  }
  while (true) {
    // ---- prefetch: row pointers two pairs ahead, neighbour ids / edge_attr one pair ahead ----
    const int nxt_pair = pair + stride;
    const bool has_next = nxt_pair < npairs;
    const int rp_nn = pair_rowptr(rowptr, nxt_pair + stride, npairs, n_nodes, lane);
    RowPair nxt = cur;
    int nbr_n = 0;
    float2 ea_n = make_float2(0.f, 0.f);
    if (has_next) {
      nxt = pair_derive(rp_nxt, nxt_pair, n_nodes);
      if (lane < min(32, nxt.ne)) {
        nbr_n = __ldg(nbr + nxt.e0 + lane);
        ea_n = __ldg(ea + nxt.e0 + lane);
      }
    }
    // ---- the current pair ----
    float4 hi0 = zero, hi1 = zero;
    if (act) {
      hi0 = ld4(Hi + cur.r0 * ldh + 4 * lane);
      if (cur.nrow == 2) hi1 = ld4(Hi + (cur.r0 + 1) * ldh + 4 * lane);
    }
    float4 acc0 = zero, acc1 = zero, tacc0 = zero, tacc1 = zero;
    for (int eb = 0; eb < cur.ne; eb += 32) {  // 32 edges of the pair per round (one round unless a hub bus is involved)
      const int cnt = min(32, cur.ne - eb);
      int nb = my_nbr;
      float2 a_l = my_ea;
      if (eb > 0) {
        nb = 0;
        a_l = make_float2(0.f, 0.f);
        if (lane < cnt) {
          nb = __ldg(nbr + cur.e0 + eb + lane);
          a_l = __ldg(ea + cur.e0 + eb + lane);
        }
      }
      float4 tg = zero, ht = zero;  // column 32: lane u works on edge u of the round
      if (tail && lane < cnt) {
        ht = ld4(Hi + (cur.r0 + (eb + lane < cur.deg0 ? 0 : 1)) * ldh + 128);
        tg = ldg4(Hj + nb * ldh + 128);
      }
      for (int cb = 0; cb < cnt; cb += kEaWarpChunk) {
        float4 g[kEaWarpChunk];
#pragma unroll
        for (int u = 0; u < kEaWarpChunk; ++u) {
          const int s = __shfl_sync(0xffffffffu, nb, (cb + u) & 31);
          g[u] = zero;
          if (cb + u < cnt && act) g[u] = ldg4(Hj + s * ldh + 4 * lane);
        }
#pragma unroll
        for (int u = 0; u < kEaWarpChunk; ++u) {
          if (cb + u < cnt) {  // warp-uniform
            const float2 a = make_float2(__shfl_sync(0xffffffffu, a_l.x, cb + u), __shfl_sync(0xffffffffu, a_l.y, cb + u));
            if (eb + cb + u < cur.deg0) {  // warp-uniform: which row of the pair the edge belongs to
              add_relu(acc0, preact(hi0, g[u], a, w0, w1));
            } else {
              add_relu(acc1, preact(hi1, g[u], a, w0, w1));
            }
          }
        }
      }
      if (tail) {
        float4 tr = zero;
        if (lane < cnt) {
          const float4 p = preact(ht, tg, a_l, s_w[32][0], s_w[32][1]);
          tr = make_float4(fmaxf(p.x, 0.f), fmaxf(p.y, 0.f), fmaxf(p.z, 0.f), fmaxf(p.w, 0.f));
        }
        for (int u = 0; u < cnt; ++u) {  // ascending edge order; every lane keeps a copy
          const float4 v = shfl4(tr, u);
          if (eb + u < cur.deg0) {
            tacc0.x += v.x; tacc0.y += v.y; tacc0.z += v.z; tacc0.w += v.w;
          } else {
            tacc1.x += v.x; tacc1.y += v.y; tacc1.z += v.z; tacc1.w += v.w;
          }
        }
      }
    }
    if (act) {
      st4(S + cur.r0 * lds + 4 * lane, acc0);
      if (cur.nrow == 2) st4(S + (cur.r0 + 1) * lds + 4 * lane, acc1);
    }
    if (tail && lane < cur.nrow) st4(S + (cur.r0 + lane) * lds + 128, lane == 0 ? tacc0 : tacc1);
    if (!has_next) break;
    pair = nxt_pair;
    cur = nxt;
    my_nbr = nbr_n;
    my_ea = ea_n;
    rp_nxt = rp_nn;
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

// which forward kernel: warp-owned row pairs for 24..33 float4 columns (PFN_EA_FWD=cta forces the CTA-slab kernel,
// PFN_EA_FWD=warp the warp kernel; read per call so that tests can compare the two)
bool ea_fwd_warp_rows(int c4) {
  const bool eligible = c4 >= 24 && c4 <= 33;
  const char* e = std::getenv("PFN_EA_FWD");
  if (e != nullptr && e[0] == 'c') return false;
  if (e != nullptr && e[0] == 'w') return eligible;
  return eligible && kEaFwdWarpDefault;
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
  if (ea_fwd_warp_rows(t.c4) && n_nodes * std::max(ldh, lds) < (int64_t(1) << 31)) {
    static const int resident = [] {
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ea_fwd_warp, kEaWarpThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
      return per_sm;
    }();
    const int64_t warps = kEaWarpThreads / 32, pairs = ceil_div64(n_nodes, 2);
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div64(pairs, warps), int64_t(sm_count()) * resident));
    PFN_CUDA_OK(launch_kernel(k_ea_fwd_warp, dim3(blocks), dim3(kEaWarpThreads), 0, stream, Hi, Hj, ldh, static_cast<const int*>(g.rowptr_t),
                              static_cast<const int*>(g.nbr_t), reinterpret_cast<const float2*>(g.ea_t), We, ldwe, S, lds,
                              static_cast<int>(n_nodes), static_cast<int>(h), t.c4));
    PFN_LAUNCHED();
    return 0;
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
