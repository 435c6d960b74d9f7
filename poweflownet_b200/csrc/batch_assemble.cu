// libpfn_b200 -- mini-batch assembly from a device-resident dataset.
//
// The reference keeps every sample as a PyG `Data` object on the host (datasets/PowerFlowData.py:171-217), collates
// 128 of them per step with `Batch.from_data_list` (train.py:90, utils/training.py:55) and copies the result to the
// GPU.  Here the RAW arrays of the dataset stay in HBM in the reference's own file format --
//   node_features [S, n, 6] = (index, type, Vm, Va, P, Q),  edge_features [S, E, 4] = (from, to, r, x)
// (40 bytes per bus/branch, against 80 / 40 bytes of the processed + normalised tensors) -- and ONE launch turns a list
// of sample ids into the tensors of the PyG `Batch`:
//   y = (Vm, Va, P, Q)                                   :188
//   bus_type = long(type), pred_mask = table[bus_type]   :189-190, 71-74
//   x = y * (1 - pred_mask)                              :191
//   edge_index = long(from, to)^T + node offset, edge_attr = (r, x)         :198-199 + PyG collation
//   x, y <- (. - xymean) / (xystd + 1e-7),  edge_attr likewise              :132-139
//   batch [N], ptr [B + 1]                               PyG `Batch`
// Integer outputs are bit-exact; the float outputs too (IEEE subtract / divide in the reference's order).
// A dataset may concatenate several cases (`--case mixed`, :151-155): sample ids index the concatenation and every
// sample carries its own (n, E), so batches that mix graph sizes come out in the same variable-`ptr` layout.
#include "common.cuh"

namespace pfn {
namespace {

constexpr int kMaxCases = 8;
constexpr int kAsmBlock = 256;

struct AsmCase {
  const float* node_raw;
  const float* edge_raw;
  long long first_id;  // dataset index of this case's first sample
  int n_nodes, n_edges;
};
struct AsmArgs {
  AsmCase cases[kMaxCases];
  int n_cases, batch, normalize, pad_;
  float xm[4], xden[4], em[2], eden[2];
  const long long* ids;  // device [batch]
  int* off;              // device scratch: node_off [batch + 1] | edge_off [batch + 1] | case_of [batch] | bad flag
  float *x, *y, *edge_attr;
  long long *bus_type, *pred_mask, *edge_index, *batch_vec, *ptr;
  long long ei_stride;   // = E of the batch (row pitch of edge_index [2, E])
  long long n_total, e_total, n_samples;
  unsigned long long bus_seed;  // != 0: bus_type <- uniform {0, 1} (the `random_bus_type` transform, :36-40)
};
static_assert(sizeof(AsmArgs) <= 4000, "kernel parameter space");

__device__ __forceinline__ int case_of_sample(const AsmArgs& a, long long id) {
  int c = 0;
  for (int k = 1; k < a.n_cases; ++k)
    if (id >= a.cases[k].first_id) c = k;
  return c;
}

// node / edge offsets of the batch: one CTA, each thread scans a contiguous run of samples
__global__ void __launch_bounds__(kAsmBlock) k_batch_offsets(const __grid_constant__ AsmArgs a) {
  pdl_wait();
  __shared__ long long s_n[kAsmBlock], s_e[kAsmBlock];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const int B = a.batch, per = (B + kAsmBlock - 1) / kAsmBlock;
  const int beg = min(B, int(threadIdx.x) * per), end = min(B, beg + per);
  int* node_off = a.off;
  int* edge_off = a.off + (B + 1);
  int* case_of = a.off + 2 * (B + 1);
  long long sn = 0, se = 0;
  for (int b = beg; b < end; ++b) {
    const long long id = a.ids[b];
    if (id < 0 || id >= a.n_samples) s_bad = 1;
    const int c = case_of_sample(a, id < 0 ? 0 : id);
    case_of[b] = c;
    sn += a.cases[c].n_nodes;
    se += a.cases[c].n_edges;
  }
  s_n[threadIdx.x] = sn;
  s_e[threadIdx.x] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long an = 0, ae = 0;
    for (int t = 0; t < kAsmBlock; ++t) {
      const long long tn = s_n[t], te = s_e[t];
      s_n[t] = an;
      s_e[t] = ae;
      an += tn;
      ae += te;
    }
    node_off[B] = int(an);
    edge_off[B] = int(ae);
    a.ptr[B] = an;
    // totals the host did not expect (it sized the outputs), or a sample id outside the dataset: the fill kernel stops
    a.off[3 * B + 2] = (s_bad != 0 || an != a.n_total || ae != a.e_total) ? 1 : 0;
  }
  __syncthreads();
  long long on = s_n[threadIdx.x], oe = s_e[threadIdx.x];
  for (int b = beg; b < end; ++b) {
    node_off[b] = int(on);
    edge_off[b] = int(oe);
    a.ptr[b] = on;
    const int c = case_of[b];
    on += a.cases[c].n_nodes;
    oe += a.cases[c].n_edges;
  }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// blockIdx.y = graph of the batch; thread t handles bus t and branch t of that graph
__global__ void __launch_bounds__(kAsmBlock) k_batch_fill(const __grid_constant__ AsmArgs a) {
  pdl_wait();
  const int B = a.batch, b = blockIdx.y;
  if (a.off[3 * B + 2] != 0) return;
  const int c = a.off[2 * (B + 1) + b];
  const AsmCase cs = a.cases[c];
  const long long local = a.ids[b] - cs.first_id;
  const int t = blockIdx.x * kAsmBlock + threadIdx.x;
  const int node0 = a.off[b], edge0 = a.off[(B + 1) + b];
  if (t < cs.n_nodes) {
    const float2* __restrict__ row = reinterpret_cast<const float2*>(cs.node_raw + (local * cs.n_nodes + t) * 6);
    const float2 r0 = row[0], r1 = row[1], r2 = row[2];  // (index, type) (Vm, Va) (P, Q)
    long long bt = static_cast<long long>(r0.y);           // `.type(torch.long)` truncates
    const int m = bt == 0 ? 0x3 << 2 : (bt == 1 ? 0xA : 0x3);  // bit k = pred_mask[k]: (0,0,1,1) (0,1,0,1) (1,1,0,0)
    const float yv[4] = {r1.x, r1.y, r2.x, r2.y};
    float xo[4], yo[4];
    long long pm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pm[k] = (m >> k) & 1;
      const float xr = __fmul_rn(yv[k], pm[k] ? 0.f : 1.f);  // y.clone() * (1. - mask)
      xo[k] = a.normalize ? __fdiv_rn(__fsub_rn(xr, a.xm[k]), a.xden[k]) : xr;
      yo[k] = a.normalize ? __fdiv_rn(__fsub_rn(yv[k], a.xm[k]), a.xden[k]) : yv[k];
    }
    const long long g = node0 + t;
    *reinterpret_cast<float4*>(a.x + 4 * g) = make_float4(xo[0], xo[1], xo[2], xo[3]);
    *reinterpret_cast<float4*>(a.y + 4 * g) = make_float4(yo[0], yo[1], yo[2], yo[3]);
    *reinterpret_cast<longlong2*>(a.pred_mask + 4 * g) = make_longlong2(pm[0], pm[1]);
    *reinterpret_cast<longlong2*>(a.pred_mask + 4 * g + 2) = make_longlong2(pm[2], pm[3]);
    if (a.bus_seed != 0) bt = static_cast<long long>(mix64(a.bus_seed + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(g + 1)) >> 63);
    a.bus_type[g] = bt;
    a.batch_vec[g] = b;
  }
  if (t < cs.n_edges) {
    const float4 r = *reinterpret_cast<const float4*>(cs.edge_raw + (local * cs.n_edges + t) * 4);  // (from, to, r, x)
    const long long g = edge0 + t;
    a.edge_index[g] = static_cast<long long>(r.x) + node0;
    a.edge_index[a.ei_stride + g] = static_cast<long long>(r.y) + node0;
    float2 ea = make_float2(r.z, r.w);
    if (a.normalize) {
      ea.x = __fdiv_rn(__fsub_rn(ea.x, a.em[0]), a.eden[0]);
      ea.y = __fdiv_rn(__fsub_rn(ea.y, a.em[1]), a.eden[1]);
    }
    *reinterpret_cast<float2*>(a.edge_attr + 2 * g) = ea;
  }
}

}  // namespace
}  // namespace pfn

using namespace pfn;

extern "C" size_t pfn_batch_assemble_scratch_bytes(int64_t batch) { return size_t(3 * batch + 3) * sizeof(int32_t); }

extern "C" int pfn_batch_assemble(const pfn_dataset_case* cases, int n_cases, const int64_t* sample_ids, int64_t batch,
                                  int64_t n_total, int64_t e_total, const float* norm, uint64_t random_bus_type_seed,
                                  float* x, float* y, int64_t* bus_type, int64_t* pred_mask, int64_t* edge_index,
                                  float* edge_attr, int64_t* batch_vec, int64_t* ptr, void* scratch, void* stream_) {
  PFN_REQUIRE(cases && n_cases >= 1 && n_cases <= kMaxCases, PFN_E_INVALID, "pfn_batch_assemble: 1..%d cases", kMaxCases);
  PFN_REQUIRE(sample_ids && batch >= 1 && batch <= (int64_t(1) << 20), PFN_E_INVALID, "pfn_batch_assemble: bad batch size");
  PFN_REQUIRE(n_total >= 0 && e_total >= 0 && n_total < (int64_t(1) << 31) && e_total < (int64_t(1) << 31), PFN_E_INVALID,
              "pfn_batch_assemble: batch too large for 32-bit offsets");
  PFN_REQUIRE(x && y && bus_type && pred_mask && batch_vec && ptr && scratch && (e_total == 0 || (edge_index && edge_attr)),
              PFN_E_INVALID, "pfn_batch_assemble: null output");
  PFN_REQUIRE(aligned16(x) && aligned16(y) && aligned16(pred_mask) && (reinterpret_cast<uintptr_t>(edge_attr) & 7u) == 0,
              PFN_E_INVALID, "pfn_batch_assemble: outputs must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AsmArgs a{};
  long long first = 0;
  int max_items = 0;
  for (int c = 0; c < n_cases; ++c) {
    const pfn_dataset_case& s = cases[c];
    PFN_REQUIRE(s.n_samples >= 0 && s.n_nodes >= 1 && s.n_edges >= 0 && s.n_nodes < (1 << 30) && s.n_edges < (1 << 30), PFN_E_INVALID,
                "pfn_batch_assemble: case %d has a bad shape", c);
    PFN_REQUIRE(s.n_samples == 0 || (s.node_features && (s.n_edges == 0 || s.edge_features)), PFN_E_INVALID,
                "pfn_batch_assemble: case %d has null arrays", c);
    PFN_REQUIRE((reinterpret_cast<uintptr_t>(s.node_features) & 7u) == 0 && aligned16(s.edge_features), PFN_E_INVALID,
                "pfn_batch_assemble: case %d: node_features must be 8-byte and edge_features 16-byte aligned", c);
    a.cases[c].node_raw = s.node_features;
    a.cases[c].edge_raw = s.edge_features;
    a.cases[c].first_id = first;
    a.cases[c].n_nodes = static_cast<int>(s.n_nodes);
    a.cases[c].n_edges = static_cast<int>(s.n_edges);
    first += s.n_samples;
    const int items = static_cast<int>(s.n_nodes > s.n_edges ? s.n_nodes : s.n_edges);
    if (s.n_samples > 0 && items > max_items) max_items = items;
  }
  PFN_REQUIRE(max_items > 0, PFN_E_INVALID, "pfn_batch_assemble: empty dataset");
  a.n_cases = n_cases;
  a.batch = static_cast<int>(batch);
  a.normalize = norm != nullptr ? 1 : 0;
  if (norm != nullptr) {
    for (int k = 0; k < 4; ++k) a.xm[k] = norm[k], a.xden[k] = norm[4 + k];
    for (int k = 0; k < 2; ++k) a.em[k] = norm[8 + k], a.eden[k] = norm[10 + k];
  }
  a.ids = reinterpret_cast<const long long*>(sample_ids);
  a.off = static_cast<int*>(scratch);
  a.x = x;
  a.y = y;
  a.edge_attr = edge_attr;
  a.bus_type = reinterpret_cast<long long*>(bus_type);
  a.pred_mask = reinterpret_cast<long long*>(pred_mask);
  a.edge_index = reinterpret_cast<long long*>(edge_index);
  a.batch_vec = reinterpret_cast<long long*>(batch_vec);
  a.ptr = reinterpret_cast<long long*>(ptr);
  a.ei_stride = e_total;
  a.n_total = n_total;
  a.e_total = e_total;
  a.n_samples = first;
  a.bus_seed = random_bus_type_seed;
  PFN_CUDA_OK(launch_kernel(k_batch_offsets, dim3(1), dim3(kAsmBlock), 0, stream, a));
  PFN_LAUNCHED();
  const unsigned gx = static_cast<unsigned>(ceil_div64(max_items, kAsmBlock));
  PFN_REQUIRE(batch <= 65535, PFN_E_INVALID, "pfn_batch_assemble: at most 65535 graphs per batch");
  PFN_CUDA_OK(launch_kernel(k_batch_fill, dim3(gx, static_cast<unsigned>(batch)), dim3(kAsmBlock), 0, stream, a));
  PFN_LAUNCHED();
  return 0;
}

extern "C" int pfn_batch_assemble_status(const void* scratch, int64_t batch, int32_t* host_flag, void* stream_) {
  PFN_REQUIRE(scratch && host_flag && batch >= 1, PFN_E_INVALID, "pfn_batch_assemble_status: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_CUDA_OK(cudaMemcpyAsync(host_flag, static_cast<const int32_t*>(scratch) + 3 * batch + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  PFN_CUDA_OK(cudaStreamSynchronize(stream));
  return 0;
}
