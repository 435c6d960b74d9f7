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
