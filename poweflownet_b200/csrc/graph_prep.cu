// Graph preparation for one mini-batch, entirely on the device and without a host round trip.
//
// Replaces (reference, /root/reference):
//   networks/MPN.py:498-504  is_directed   -- first-edge reverse test (the reference syncs to the host)
//   networks/MPN.py:506-523  undirect_graph -- edge_index <- [ei | flip(ei)], edge_attr <- [ea ; ea]
//   networks/MPN.py:43-47 / PyG gcn_norm -- in-degree and deg^-1/2
// and produces what PyG's gather/scatter does implicitly: a CSR by target node (forward passes)
// and a CSR by source node (transposed passes of the backward), both STABLE (slots inside a row
// are in ascending directed-edge id, i.e. the order a sequential scatter_add visits them), with
// edge_attr permuted into each order so the edge kernels read it coalesced.
//
// Integer work only; everything here is bit-exact against the oracle (tests/test_graph_prep*.py).
#include "common.cuh"

namespace pfn {
namespace {

constexpr int kScanThreads = 1024;

__device__ __forceinline__ bool graph_directed(const int32_t* meta, int e_raw, int mode) {
  return mode == 1 && e_raw > 0 && meta[0] == 0;
}

__device__ __forceinline__ void edge_at(const int64_t* __restrict__ ei, int64_t stride, int e_raw, int e,
                                        int64_t& s, int64_t& t) {
  if (e < e_raw) {
    s = ei[e];
    t = ei[stride + e];
  } else {  // reversed copy appended after all originals (networks/MPN.py:508-515)
    s = ei[stride + (e - e_raw)];
    t = ei[e - e_raw];
  }
}

// Pass 1: does the reverse (b -> a) of the first edge (a -> b) exist anywhere?  (MPN.py:504)
__global__ void k_find_reverse(const int64_t* __restrict__ ei, int64_t stride, int e_raw, int32_t* meta) {
  pdl_wait();
  const int64_t a = ei[0], b = ei[stride];
  int found = 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < e_raw; e += gridDim.x * blockDim.x)
    found |= (ei[e] == b && ei[stride + e] == a);
  if (__syncthreads_or(found) && threadIdx.x == 0) atomicOr(&meta[0], 1);
}

// Pass 2: in/out degree histograms (integer atomics: result independent of order).
__global__ void k_count(const int64_t* __restrict__ ei, int64_t stride, int e_raw, int mode, int n_nodes,
                        int32_t* meta, int32_t* cnt_t, int32_t* cnt_s) {
  pdl_wait();
  const bool directed = graph_directed(meta, e_raw, mode);
  const int E = directed ? 2 * e_raw : e_raw;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    meta[1] = directed;
    meta[2] = E;
    meta[4] = e_raw;
    meta[5] = n_nodes;
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int64_t s, t;
    edge_at(ei, stride, e_raw, e, s, t);
    if (s < 0 || s >= n_nodes || t < 0 || t >= n_nodes) {
      meta[3] = 1;
      continue;
    }
    atomicAdd(&cnt_t[t], 1);
    atomicAdd(&cnt_s[s], 1);
  }
}

// Pass 3: exclusive scan of a histogram into rowptr (one CTA per CSR), degree and deg^-1/2, and
// reset of the histogram so pass 4 can reuse it as the per-row cursor.
// Shared-memory variant of k_scan for batches whose counts fit (n_nodes <= kScanSmemNodes): all global traffic is
// coalesced (thread t touches elements t, t + 1024, ...), the per-thread contiguous chunks are read from and written to
// shared memory.  The register variant below lets every thread store 16 scattered words per array from ONE SM, which
// made its load/store unit the bottleneck (36 us for 15104 nodes).
constexpr int kScanSmemNodes = 24 * 1024;
__global__ void __launch_bounds__(kScanThreads) k_scan_smem(int n_nodes, int32_t* cnt_t, int32_t* cnt_s,
                                                            int32_t* rowptr_t, int32_t* rowptr_s, float* deg, float* dis) {
  extern __shared__ int32_t sm_scan[];  // counts [n_pad] | exclusive sums [n_pad]
  pdl_wait();
  int32_t* cnt = blockIdx.x == 0 ? cnt_t : cnt_s;
  int32_t* rowptr = blockIdx.x == 0 ? rowptr_t : rowptr_s;
  const int per = (n_nodes + kScanThreads - 1) / kScanThreads;
  const int n_pad = per * kScanThreads;
  const int pad_words = n_pad + n_pad / 32 + 32;
  int32_t* s_cnt = sm_scan;
  int32_t* s_sum = sm_scan + pad_words;
  auto slot = [&](int i) { return i + (i >> 5); };  // pad one word per 32: chunk-strided accesses spread over the banks
  for (int i = threadIdx.x; i < n_pad; i += kScanThreads) s_cnt[slot(i)] = i < n_nodes ? cnt[i] : 0;
  __syncthreads();
  const int beg = threadIdx.x * per;
  int local = 0;
  for (int j = 0; j < per; ++j) local += s_cnt[slot(beg + j)];
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += v;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? warp_sums[warp - 1] : 0);
  for (int j = 0; j < per; ++j) {
    const int c = s_cnt[slot(beg + j)];
    s_sum[slot(beg + j)] = run;
    run += c;
  }
  if (threadIdx.x == kScanThreads - 1) rowptr[n_nodes] = run;
  __syncthreads();
  for (int i = threadIdx.x; i < n_nodes; i += kScanThreads) {
    const int c = s_cnt[slot(i)];
    rowptr[i] = s_sum[slot(i)];
    cnt[i] = 0;
    if (blockIdx.x == 0) {
      deg[i] = static_cast<float>(c);
      dis[i] = c > 0 ? 1.0f / sqrtf(static_cast<float>(c)) : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan(int n_nodes, int32_t* cnt_t, int32_t* cnt_s,
                                                       int32_t* rowptr_t, int32_t* rowptr_s, float* deg,
                                                       float* dis) {
  pdl_wait();
  int32_t* cnt = blockIdx.x == 0 ? cnt_t : cnt_s;
  int32_t* rowptr = blockIdx.x == 0 ? rowptr_t : rowptr_s;
  // each thread owns a contiguous chunk whose length is a multiple of 4 so it can be read with 16-byte loads; chunks
  // of up to kRegChunk counts stay in registers between the two passes (one memory round trip instead of two)
  constexpr int kRegChunk = 16;
  const int per = ((n_nodes + kScanThreads - 1) / kScanThreads + 3) & ~3;
  const int beg = min(n_nodes, (int)threadIdx.x * per), end = min(n_nodes, beg + per);
  const bool vec = (reinterpret_cast<uintptr_t>(cnt) & 15u) == 0;
  const bool in_regs = per <= kRegChunk;
  int vals[kRegChunk];
  int local = 0;
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kRegChunk; j += 4) {
      int4 v = make_int4(0, 0, 0, 0);
      if (j < per) {
        if (vec && beg + j + 3 < end) {
          v = *reinterpret_cast<const int4*>(cnt + beg + j);
        } else {
          if (beg + j + 0 < end) v.x = cnt[beg + j + 0];
          if (beg + j + 1 < end) v.y = cnt[beg + j + 1];
          if (beg + j + 2 < end) v.z = cnt[beg + j + 2];
          if (beg + j + 3 < end) v.w = cnt[beg + j + 3];
        }
      }
      vals[j] = v.x; vals[j + 1] = v.y; vals[j + 2] = v.z; vals[j + 3] = v.w;
      local += v.x + v.y + v.z + v.w;
    }
  } else {
    for (int i = beg; i < end; ++i) local += cnt[i];
  }
  // block-wide exclusive scan of `local`
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += v;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? warp_sums[warp - 1] : 0);
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < kRegChunk; ++j) {
      const int i = beg + j;
      if (j < per && i < end) {
        const int c = vals[j];
        rowptr[i] = run;
        run += c;
        cnt[i] = 0;
        if (blockIdx.x == 0) {
          deg[i] = static_cast<float>(c);
          dis[i] = c > 0 ? 1.0f / sqrtf(static_cast<float>(c)) : 0.0f;
        }
      }
    }
  } else {
    for (int i = beg; i < end; ++i) {
      const int c = cnt[i];
      rowptr[i] = run;
      run += c;
      cnt[i] = 0;
      if (blockIdx.x == 0) {
        deg[i] = static_cast<float>(c);
        dis[i] = c > 0 ? 1.0f / sqrtf(static_cast<float>(c)) : 0.0f;
      }
    }
  }
  if (threadIdx.x == kScanThreads - 1) rowptr[n_nodes] = run;
}

// Pass 4: drop every directed edge id into its row (arbitrary order inside the row for now).
__global__ void k_fill(const int64_t* __restrict__ ei, int64_t stride, int e_raw, int mode, int n_nodes,
                       const int32_t* meta, const int32_t* __restrict__ rowptr_t,
                       const int32_t* __restrict__ rowptr_s, int32_t* cur_t, int32_t* cur_s, int32_t* eid_t,
                       int32_t* eid_s) {
  pdl_wait();
  const int E = graph_directed(meta, e_raw, mode) ? 2 * e_raw : e_raw;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int64_t s, t;
    edge_at(ei, stride, e_raw, e, s, t);
    if (s < 0 || s >= n_nodes || t < 0 || t >= n_nodes) continue;
    eid_t[rowptr_t[t] + atomicAdd(&cur_t[t], 1)] = e;
    eid_s[rowptr_s[s] + atomicAdd(&cur_s[s], 1)] = e;
  }
}

// Pass 5: one thread per (CSR, row): sort the row's edge ids ascending (rows are short: a bus has a
// handful of branches) => stable CSR; then write the neighbour id and the permuted edge_attr.
__global__ void k_finalize(const int64_t* __restrict__ ei, int64_t stride, const float2* __restrict__ ea,
                           int e_raw, int n_nodes, const int32_t* __restrict__ rowptr_t,
                           const int32_t* __restrict__ rowptr_s, int32_t* eid_t, int32_t* eid_s, int32_t* nbr_t,
                           int32_t* nbr_s, float2* ea_t, float2* ea_s) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * n_nodes) return;
  const bool by_target = idx < n_nodes;
  const int row = by_target ? idx : idx - n_nodes;
  const int32_t* rowptr = by_target ? rowptr_t : rowptr_s;
  int32_t* eid = by_target ? eid_t : eid_s;
  int32_t* nbr = by_target ? nbr_t : nbr_s;
  float2* eao = by_target ? ea_t : ea_s;
  const int beg = rowptr[row], end = rowptr[row + 1];
  for (int i = beg + 1; i < end; ++i) {  // insertion sort
    const int key = eid[i];
    int j = i - 1;
    while (j >= beg && eid[j] > key) {
      eid[j + 1] = eid[j];
      --j;
    }
    eid[j + 1] = key;
  }
  for (int i = beg; i < end; ++i) {
    const int e = eid[i];
    int64_t s, t;
    edge_at(ei, stride, e_raw, e, s, t);
    nbr[i] = static_cast<int32_t>(by_target ? s : t);
    eao[i] = ea[e < e_raw ? e : e - e_raw];  // networks/MPN.py:516-519: attributes duplicated
  }
}

// ---- one-launch preparation for batches of small graphs laid out tile by tile ------------------------------------------
// The five passes above cost ~39 us per step at case118v2 x 128 (two of them are one- or two-CTA kernels), 5 % of the
// training step.  A PyG batch of equal-sized graphs is block diagonal: the nodes of tile t are rows [t R, (t+1) R) and --
// because the loader concatenates the graphs' edge lists in order -- its edges are columns [t P, (t+1) P) of edge_index
// with P = e_raw / n_tiles.  One CTA per tile then builds the tile's slice of BOTH stable CSRs in shared memory (<= 768
// directed edges, <= 128 rows) and writes it to the same workspace arrays, bit for bit what the general passes write:
//   * is_directed (MPN.py:504): the reverse of edge 0 can only sit among tile 0's edges, every CTA looks there itself;
//   * row pointers: the directed edges of the tiles before this one are t * (P or 2 P), no grid-wide scan;
//   * stable order: one warp per CSR walks the directed list 32 edges at a time, __match_any_sync ranks the lanes that
//     hit the same row (lane order = edge id order) behind the row's cursor.
// "Columns [t P, (t+1) P) belong to tile t" is a GUESS that the kernel VALIDATES: an edge of the range with an endpoint
// outside the tile raises meta[6] (the broken-tile flag) and makes the tile's neighbour ids -1, so that the graph-resident
// kernels poison the tile with NaN exactly as for any other broken promise; the caller then falls back to pfn_graph_prep.
// The ranges partition [0, e_raw) by construction, so if every tile validates, every edge has been placed.
constexpr int kTileThreads = 256;
constexpr int kTileEdgeCap = 768;  // = kFusedEdgeCap: directed edges of one tile
constexpr int kTileRows = 128;

__global__ void __launch_bounds__(kTileThreads) k_prep_tiled(const int64_t* __restrict__ ei, int64_t stride,
                                                             const float2* __restrict__ ea, int e_raw, int mode, int n_nodes,
                                                             int tile_rows, int per, GraphView g) {
  __shared__ int16_t s_src[kTileEdgeCap], s_dst[kTileEdgeCap];  // local endpoints of the ORIGINAL edges of the tile
  __shared__ float2 s_ea[kTileEdgeCap];
  __shared__ int s_cnt[2][kTileRows], s_ptr[2][kTileRows + 1], s_cur[2][kTileRows];
  __shared__ int16_t o_nbr[2][kTileEdgeCap];
  __shared__ int16_t o_loc[2][kTileEdgeCap];  // position in the tile's directed list
  __shared__ int s_bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = blockIdx.x, n_tiles = gridDim.x;
  const int n0 = t * tile_rows, n1 = min(n_nodes, n0 + tile_rows), nr = n1 - n0;
  const int lo = t * per;
  if (tid < kTileRows) s_cnt[0][tid] = s_cnt[1][tid] = 0;
  if (tid == 0) s_bad = 0;
  pdl_wait();
  // one round trip: this tile's edges, and the first-edge test over tile 0's edges
  const int64_t a0 = ei[0], b0 = ei[stride];
  int found = 0, bad = 0, oob = 0;
  for (int i = tid; i < per; i += kTileThreads) {
    const int64_t s = ei[lo + i], d = ei[stride + lo + i];
    const int64_t fs = ei[i], fd = ei[stride + i];
    found |= (fs == b0 && fd == a0);
    if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) oob = 1;
    const bool ok = s >= n0 && s < n1 && d >= n0 && d < n1;
    if (!ok) bad = 1;
    s_src[i] = static_cast<int16_t>(ok ? s - n0 : 0);
    s_dst[i] = static_cast<int16_t>(ok ? d - n0 : 0);
    s_ea[i] = ea[lo + i];
  }
  found = __syncthreads_or(found);
  const bool directed = mode == 1 && e_raw > 0 && found == 0;
  const int ne = directed ? 2 * per : per;  // directed edges of the tile; host guarantees per <= kTileEdgeCap
  if (ne > kTileEdgeCap || nr <= 0 || nr > kTileRows) bad = 1;
  if (bad) s_bad = 1;
  const int base = t * ne;
  const int ne_c = min(ne, kTileEdgeCap);
  // directed edge d of the tile: d < per -> original (src -> dst), else the reversed copy of original d - per
  auto row_of = [&](int which, int d) {  // which 0: CSR by target, 1: by source
    const bool rev = d >= per;
    const int o = rev ? d - per : d;
    return int((which == 0) != rev ? s_dst[o] : s_src[o]);
  };
  auto nbr_of = [&](int which, int d) {
    const bool rev = d >= per;
    const int o = rev ? d - per : d;
    return int((which == 0) != rev ? s_src[o] : s_dst[o]);
  };
  for (int d = tid; d < ne_c; d += kTileThreads) {
    atomicAdd(&s_cnt[0][row_of(0, d)], 1);
    atomicAdd(&s_cnt[1][row_of(1, d)], 1);
  }
  __syncthreads();
  if (warp < 2) {
    // exclusive scan of 128 counts: four per lane
    int c[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      c[j] = s_cnt[warp][4 * lane + j];
      sum += c[j];
    }
    int incl = sum;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, dlt);
      if (lane >= dlt) incl += v;
    }
    int run = incl - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_ptr[warp][4 * lane + j] = run;
      s_cur[warp][4 * lane + j] = run;
      run += c[j];
    }
    if (lane == 31) s_ptr[warp][kTileRows] = run;
    __syncwarp();
    // stable placement: lanes of a chunk that hit the same row take consecutive slots in lane (= edge id) order
    for (int d0 = 0; d0 < ne_c; d0 += 32) {
      const int d = d0 + lane;
      const bool valid = d < ne_c;
      const int row = valid ? row_of(warp, d) : kTileRows + lane;  // idle lanes: private keys
      const unsigned peers = __match_any_sync(0xffffffffu, row);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      const int pos = valid ? s_cur[warp][row] + rank : 0;
      __syncwarp();
      if (valid) {
        o_nbr[warp][pos] = static_cast<int16_t>(nbr_of(warp, d));
        o_loc[warp][pos] = static_cast<int16_t>(d);
        if ((peers >> lane) == 1u) s_cur[warp][row] = pos + 1;  // the last peer leaves the cursor behind the group
      }
      __syncwarp();
    }
  }
  __syncthreads();
  const bool tile_bad = s_bad != 0;
  for (int w = 0; w < 2; ++w) {
    int32_t* rowptr = w == 0 ? g.rowptr_t : g.rowptr_s;
    int32_t* nbr = w == 0 ? g.nbr_t : g.nbr_s;
    int32_t* eid = w == 0 ? g.eid_t : g.eid_s;
    float2* eao = reinterpret_cast<float2*>(w == 0 ? g.ea_t : g.ea_s);
    for (int p = tid; p < ne_c; p += kTileThreads) {
      const int d = o_loc[w][p];
      const int o = d >= per ? d - per : d;
      nbr[base + p] = tile_bad ? -1 : n0 + o_nbr[w][p];
      eid[base + p] = d >= per ? e_raw + lo + o : lo + o;
      eao[base + p] = s_ea[o];
    }
    for (int i = tid; i < nr; i += kTileThreads) rowptr[n0 + i] = base + s_ptr[w][i];
    if (t == n_tiles - 1 && tid == 0) rowptr[n_nodes] = n_tiles * ne;
  }
  for (int i = tid; i < nr; i += kTileThreads) {
    const int c = s_cnt[0][i];
    g.deg[n0 + i] = static_cast<float>(c);
    g.dis[n0 + i] = c > 0 ? 1.0f / sqrtf(static_cast<float>(c)) : 0.0f;
  }
  if (t == 0 && tid == 0) {
    g.meta[0] = (mode == 1 && found) ? 1 : 0;
    g.meta[1] = directed;
    g.meta[2] = n_tiles * ne;
    g.meta[4] = e_raw;
    g.meta[5] = n_nodes;
  }
  if (tid == 0 && tile_bad) g.meta[6] = 1;
  if (__syncthreads_or(oob) && tid == 0) g.meta[3] = 1;
}

__global__ void k_export(const int64_t* __restrict__ ei, int64_t stride, const float2* __restrict__ ea,
                         int e_raw, int e_out, int64_t* ei_out, float2* ea_out) {
  pdl_wait();
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < e_out; e += gridDim.x * blockDim.x) {
    int64_t s, t;
    edge_at(ei, stride, e_raw, e, s, t);
    ei_out[e] = s;
    ei_out[e_out + e] = t;
    ea_out[e] = ea[e < e_raw ? e : e - e_raw];
  }
}

inline int64_t align16(int64_t x) { return (x + 15) & ~int64_t(15); }

}  // namespace
}  // namespace pfn

using namespace pfn;

extern "C" int pfn_graph_layout_get(int64_t n_nodes, int64_t e_raw, pfn_graph_layout* out) {
  PFN_REQUIRE(out != nullptr && n_nodes >= 0 && e_raw >= 0, PFN_E_INVALID, "pfn_graph_layout_get: bad arguments");
  PFN_REQUIRE(n_nodes < (int64_t(1) << 30) && e_raw < (int64_t(1) << 29), PFN_E_UNSUPPORTED,
              "pfn_graph_layout_get: graph too large for int32 indices (N=%lld, E_raw=%lld)", (long long)n_nodes,
              (long long)e_raw);
  const int64_t cap = 2 * e_raw;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = pfn::align16(off + bytes);
    return o;
  };
  out->meta = take(8 * 4);
  out->rowptr_t = take((n_nodes + 1) * 4);
  out->nbr_t = take(cap * 4);
  out->eid_t = take(cap * 4);
  out->ea_t = take(cap * 8);
  out->rowptr_s = take((n_nodes + 1) * 4);
  out->nbr_s = take(cap * 4);
  out->eid_s = take(cap * 4);
  out->ea_s = take(cap * 8);
  out->deg = take(n_nodes * 4);
  out->dis = take(n_nodes * 4);
  out->cursor = take(2 * n_nodes * 4);
  out->total_bytes = off;
  out->e_cap = cap;
  return 0;
}

extern "C" int pfn_graph_prep(const int64_t* edge_index, int64_t ei_row_stride, const float* edge_attr,
                              int64_t n_nodes, int64_t e_raw, int undirect_mode, void* graph_ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_REQUIRE(graph_ws != nullptr && aligned16(graph_ws), PFN_E_INVALID, "pfn_graph_prep: graph_ws null or misaligned");
  PFN_REQUIRE(e_raw == 0 || (edge_index != nullptr && edge_attr != nullptr), PFN_E_INVALID,
              "pfn_graph_prep: null edge arrays");
  PFN_REQUIRE(e_raw == 0 || (reinterpret_cast<uintptr_t>(edge_attr) & 7u) == 0, PFN_E_INVALID,
              "pfn_graph_prep: edge_attr must be 8-byte aligned");
  PFN_REQUIRE(ei_row_stride >= e_raw, PFN_E_INVALID, "pfn_graph_prep: row stride < e_raw");
  pfn_graph_layout lay;
  PFN_TRY(pfn_graph_layout_get(n_nodes, e_raw, &lay));
  GraphView g = graph_view(graph_ws, n_nodes, e_raw);
  const int N = static_cast<int>(n_nodes), ER = static_cast<int>(e_raw);
  ProfScope prof(PFN_PROF_PREP, stream);
  // meta + cursors start from zero
  PFN_CUDA_OK(cudaMemsetAsync(g.meta, 0, 8 * sizeof(int32_t), stream));
  if (N > 0) PFN_CUDA_OK(cudaMemsetAsync(g.cursor, 0, size_t(2) * N * sizeof(int32_t), stream));
  int32_t* cnt_t = g.cursor;
  int32_t* cnt_s = g.cursor + N;
  const int threads = 256;
  const int cap = 2 * ER;
  const int edge_blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(std::max(cap, 1), threads), 148 * 8)));
  if (ER > 0 && undirect_mode == 1) {
    PFN_CUDA_OK(launch_kernel(k_find_reverse, dim3(edge_blocks), dim3(threads), 0, stream, edge_index, ei_row_stride, ER, g.meta));
    PFN_LAUNCHED();
  }
  PFN_CUDA_OK(launch_kernel(k_count, dim3(edge_blocks), dim3(threads), 0, stream, edge_index, ei_row_stride, ER, undirect_mode, N, g.meta, cnt_t, cnt_s));
  PFN_LAUNCHED();
  if (N <= kScanSmemNodes) {
    const int per = (N + kScanThreads - 1) / kScanThreads, n_pad = per * kScanThreads;
    const size_t smem = 2 * size_t(n_pad + n_pad / 32 + 32) * sizeof(int32_t);
    static SmemAttrOnce attr_once;
    PFN_CUDA_OK(ensure_dynamic_smem(attr_once, [] { return cudaFuncSetAttribute(k_scan_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); }));
    PFN_CUDA_OK(launch_kernel(k_scan_smem, dim3(2), dim3(kScanThreads), smem, stream, N, cnt_t, cnt_s, g.rowptr_t, g.rowptr_s, g.deg, g.dis));
  } else {
    PFN_CUDA_OK(launch_kernel(k_scan, dim3(2), dim3(kScanThreads), 0, stream, N, cnt_t, cnt_s, g.rowptr_t, g.rowptr_s, g.deg, g.dis));
  }
  PFN_LAUNCHED();
  if (ER > 0) {
    PFN_CUDA_OK(launch_kernel(k_fill, dim3(edge_blocks), dim3(threads), 0, stream, edge_index, ei_row_stride, ER, undirect_mode, N, g.meta, g.rowptr_t,
                                                g.rowptr_s, cnt_t, cnt_s, g.eid_t, g.eid_s));
    PFN_LAUNCHED();
    if (N > 0) {
      PFN_CUDA_OK(launch_kernel(k_finalize, dim3(static_cast<int>(ceil_div64(2 * n_nodes, threads))), dim3(threads), 0, stream, 
          edge_index, ei_row_stride, reinterpret_cast<const float2*>(edge_attr), ER, N, g.rowptr_t, g.rowptr_s,
          g.eid_t, g.eid_s, g.nbr_t, g.nbr_s, reinterpret_cast<float2*>(g.ea_t), reinterpret_cast<float2*>(g.ea_s)));
      PFN_LAUNCHED();
    }
  }
  return 0;
}

extern "C" int pfn_graph_prep_tiled(const int64_t* edge_index, int64_t ei_row_stride, const float* edge_attr,
                                    int64_t n_nodes, int64_t e_raw, int undirect_mode, int64_t tile_rows, void* graph_ws,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_REQUIRE(graph_ws != nullptr && aligned16(graph_ws), PFN_E_INVALID, "pfn_graph_prep_tiled: graph_ws null or misaligned");
  PFN_REQUIRE(edge_index != nullptr && edge_attr != nullptr && (reinterpret_cast<uintptr_t>(edge_attr) & 7u) == 0, PFN_E_INVALID,
              "pfn_graph_prep_tiled: null or misaligned edge arrays");
  PFN_REQUIRE(ei_row_stride >= e_raw, PFN_E_INVALID, "pfn_graph_prep_tiled: row stride < e_raw");
  if (pfn_graph_prep_tiled_supported(n_nodes, e_raw, tile_rows) == 0) return PFN_E_UNSUPPORTED;
  pfn_graph_layout lay;
  PFN_TRY(pfn_graph_layout_get(n_nodes, e_raw, &lay));
  GraphView g = graph_view(graph_ws, n_nodes, e_raw);
  const int64_t n_tiles = ceil_div64(n_nodes, tile_rows);
  ProfScope prof(PFN_PROF_PREP, stream);
  PFN_CUDA_OK(cudaMemsetAsync(g.meta, 0, 8 * sizeof(int32_t), stream));
  PFN_CUDA_OK(launch_kernel(k_prep_tiled, dim3(static_cast<unsigned>(n_tiles)), dim3(kTileThreads), 0, stream, edge_index, ei_row_stride,
                            reinterpret_cast<const float2*>(edge_attr), static_cast<int>(e_raw), undirect_mode, static_cast<int>(n_nodes),
                            static_cast<int>(tile_rows), static_cast<int>(e_raw / n_tiles), g));
  PFN_LAUNCHED();
  return 0;
}

extern "C" int pfn_graph_prep_tiled_supported(int64_t n_nodes, int64_t e_raw, int64_t tile_rows) {
  if (n_nodes <= 0 || e_raw <= 0 || tile_rows <= 0 || tile_rows > kTileRows) return 0;
  const int64_t n_tiles = ceil_div64(n_nodes, tile_rows);
  if (n_nodes % tile_rows != 0 || e_raw % n_tiles != 0) return 0;  // a short last tile holds fewer graphs, hence fewer columns
  return e_raw / n_tiles <= kTileEdgeCap ? 1 : 0;
}

extern "C" int pfn_graph_meta(const void* graph_ws, int32_t* host_meta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_REQUIRE(graph_ws != nullptr && host_meta != nullptr, PFN_E_INVALID, "pfn_graph_meta: null argument");
  int32_t tmp[8];
  PFN_CUDA_OK(cudaMemcpyAsync(tmp, graph_ws, sizeof(tmp), cudaMemcpyDeviceToHost, stream));
  PFN_CUDA_OK(cudaStreamSynchronize(stream));
  host_meta[0] = tmp[1];
  host_meta[1] = tmp[2];
  host_meta[2] = tmp[3];
  return 0;
}

extern "C" int pfn_graph_export(const int64_t* edge_index, int64_t ei_row_stride, const float* edge_attr,
                                int64_t e_raw, int64_t e_out, int64_t* ei_out, float* ea_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_REQUIRE(e_out == e_raw || e_out == 2 * e_raw, PFN_E_INVALID, "pfn_graph_export: e_out must be e_raw or 2*e_raw");
  if (e_out == 0) return 0;
  PFN_REQUIRE(edge_index && edge_attr && ei_out && ea_out, PFN_E_INVALID, "pfn_graph_export: null argument");
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div64(e_out, threads), 148 * 8));
  PFN_CUDA_OK(launch_kernel(k_export, dim3(blocks), dim3(threads), 0, stream, edge_index, ei_row_stride, reinterpret_cast<const float2*>(edge_attr),
                                           static_cast<int>(e_raw), static_cast<int>(e_out), ei_out,
                                           reinterpret_cast<float2*>(ea_out)));
  PFN_LAUNCHED();
  return 0;
}
