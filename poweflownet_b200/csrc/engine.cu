// Whole-model orchestration: MaskEmbdMultiMPN.forward (networks/MPN.py:525-559) and the backward
// pass autograd would run for it (utils/training.py:74), as stream-ordered sequences of the kernels in
// graph_prep.cu / edge_kernels.cu / gemm.cu.  No host synchronisation, no allocation, no atomics.
//
// Algebra used (identical to the reference up to fp32 rounding, checked against the oracle):
//   EdgeAggregation (MPN.py:17-28,53)   W1 = [Wi | Wj | We]  (column blocks of edge_aggr.0.weight)
//       Hi = x Wi^T + b1 ,  Hj = x Wj^T                       per-node GEMMs instead of per-edge
//       S[i] = sum_{e in in(i)} ReLU(Hi[i] + Hj[src e] + We ea_e)           fused gather/sum kernel
//       out  = S W2^T + deg (.) b2                            second Linear hoisted after the sum
//   TAGConv (PyG; MPN.py:545)   x_k = A_hat x_{k-1} ;  out = sum_k x_k W_k^T + bias    one K-segmented GEMM
//   dropout + ReLU (MPN.py:546-547) fused in the producing GEMM's epilogue; backward needs only the
//   saved output: d pre = d post * (post > 0 ? 1/(1-p) : 0).
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fused.cuh"

namespace pfn {

// ---- library state -----------------------------------------------------------------------------
static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("PFN_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;  // B200
  }
  return cached;
}

// ---- optional kernel timing ------------------------------------------------------------------------
namespace {
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // 2 per slot
  std::vector<int> cat;
  size_t used = 0;
};
ProfState g_prof;
constexpr size_t kProfMaxSlots = 1 << 16;
}  // namespace

ProfScope::ProfScope(int category, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof.on || g_prof.used >= kProfMaxSlots) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  if (g_prof.used * 2 >= g_prof.ev.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    g_prof.ev.push_back(a);
    g_prof.ev.push_back(b);
    g_prof.cat.push_back(category);
  }
  slot = static_cast<int>(g_prof.used++);
  g_prof.cat[slot] = category;
  cudaEventRecord(g_prof.ev[2 * slot], s);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof.ev[2 * slot + 1], stream);
}

namespace {

// ---- small elementwise kernels -------------------------------------------------------------------
__global__ void k_i64_to_f32(const int64_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  pdl_wait();
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = static_cast<float>(in[i]);
}

constexpr int kMseBlock = 256;
__global__ void __launch_bounds__(kMseBlock)
k_mse_partial(const float* __restrict__ out, const float* __restrict__ y, int64_t count, float inv_count,
              float* __restrict__ dout, float* __restrict__ partial) {
  pdl_wait();
  float local = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = out[i] - y[i];
    local = fmaf(d, d, local);
    dout[i] = 2.f * d * inv_count;
  }
  __shared__ float red[kMseBlock];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = kMseBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(kMseBlock)
k_mse_final(const float* __restrict__ partial, int n, float inv_count, float* __restrict__ loss) {
  pdl_wait();
  __shared__ float red[kMseBlock];
  float local = 0.f;
  for (int i = threadIdx.x; i < n; i += kMseBlock) local += partial[i];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = kMseBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * inv_count;
}
// ---- Masked_L2_loss (utils/custom_loss_functions.py:10-46, the parser's default --train_loss_fn) ---------------------
//   loss = mean_{mask=1} (out-y)^2 + regcoeff * mean_{mask=0} (out-y)^2      (second term only if `regularize`)
// masked_select + MSELoss(mean) in the reference (dynamic shapes, two device->host syncs); here three small launches
// with device-resident element counts: partial sums, final (loss + the two gradient scales), scaled gradient.
__global__ void __launch_bounds__(kMseBlock)
k_ml2_partial(const float* __restrict__ out, const float* __restrict__ y, const int64_t* __restrict__ mask, int64_t count,
              float* __restrict__ partial) {
  pdl_wait();
  float s1 = 0.f, s0 = 0.f, c1 = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = out[i] - y[i];
    const bool m = mask[i] != 0;
    s1 += m ? d * d : 0.f;
    s0 += m ? 0.f : d * d;
    c1 += m ? 1.f : 0.f;
  }
  __shared__ float red[3][kMseBlock];
  red[0][threadIdx.x] = s1;
  red[1][threadIdx.x] = s0;
  red[2][threadIdx.x] = c1;
  __syncthreads();
  for (int s = kMseBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[threadIdx.x * gridDim.x + blockIdx.x] = red[threadIdx.x][0];
}
__global__ void __launch_bounds__(kMseBlock)
k_ml2_final(const float* __restrict__ partial, int n, int64_t count, int regularize, float regcoeff,
            const float* __restrict__ global_counts, float* __restrict__ loss, float* __restrict__ scales) {
  pdl_wait();
  __shared__ float red[3][kMseBlock];
  for (int k = 0; k < 3; ++k) {
    float local = 0.f;
    for (int i = threadIdx.x; i < n; i += kMseBlock) local += partial[k * n + i];
    red[k][threadIdx.x] = local;
  }
  __syncthreads();
  for (int s = kMseBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float n1 = red[2][0], n0 = static_cast<float>(count) - n1;
    if (global_counts != nullptr) {  // data parallel: means over the selections of ALL ranks
      n1 = global_counts[0];
      n0 = global_counts[1];
    }
    float l = red[0][0] / n1;  // mean over an empty selection is NaN, as torch.nn.MSELoss gives
    float sc0 = 0.f;
    if (regularize) {
      l += regcoeff * (red[1][0] / n0);
      sc0 = 2.f * regcoeff / n0;
    }
    loss[0] = l;
    scales[0] = 2.f / n1;
    scales[1] = sc0;
  }
}
__global__ void k_ml2_grad(const float* __restrict__ out, const float* __restrict__ y, const int64_t* __restrict__ mask,
                           int64_t count, const float* __restrict__ scales, float* __restrict__ dout) {
  pdl_wait();
  const float sc1 = scales[0], sc0 = scales[1];
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    dout[i] = (out[i] - y[i]) * (mask[i] != 0 ? sc1 : sc0);
}
// Variable-size tiles for batches that mix graph sizes: whole graphs are packed greedily, in order, into tiles of at
// most 128 rows (`ptr` = PyG Batch.ptr).  One thread: a batch holds a few hundred graphs.  meta[7] = number of tiles;
// a graph with more than 128 nodes cannot be tiled and raises meta[6] (the same flag a broken closed-tile promise raises).
__global__ void k_tile_table(const int64_t* __restrict__ ptr, int n_graphs, int n_nodes, int* __restrict__ tile_start,
                             int32_t* __restrict__ meta) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int t = 0, start = 0;
  tile_start[0] = 0;
  for (int g = 0; g < n_graphs; ++g) {
    const int end = static_cast<int>(ptr[g + 1]);
    if (end - static_cast<int>(ptr[g]) > 128) meta[6] = 1;
    if (end - start > 128) {  // graph g does not fit any more: close the tile in front of it
      const int cut = static_cast<int>(ptr[g]);
      if (cut > start) {
        tile_start[++t] = cut;
        start = cut;
      }
    }
  }
  if (n_nodes > start) tile_start[++t] = n_nodes;
  meta[7] = t;
}

int mse_blocks(int64_t count) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div64(count, kMseBlock * 4), 1024)));
}

// ---- model plan ------------------------------------------------------------------------------------
struct LayerPlan {
  bool is_ea;
  int fin, fout;
  int p0;       // index of the layer's first tensor in the params table
  int slot;     // index among the EA layers (is_ea) or among the TAG layers
  bool act;     // dropout+ReLU applied to the output (every layer but the last)
};

struct Plan {
  pfn_mpn_desc d;
  std::vector<LayerPlan> layers;
  int n_ea = 0, n_tag = 0;
  int p_mask = 0;  // index of mask_embd.0.weight
  int n_params = 0;
  int64_t N = 0, ldh = 0;
  // activation workspace offsets (floats)
  int64_t off_maskf = 0, off_t1 = 0, off_x0 = 0, off_ea = 0, off_tag = 0, off_tiles = 0, act_floats = 0;
  // scratch offsets (floats)
  int64_t off_dz = 0, off_ds = 0, off_dhi = 0, off_dhj = 0, off_dxcat = 0, off_dx0 = 0, off_part = 0, scratch_floats = 0;
  int64_t part_bytes = 0;
  // 16-byte-pitched (and transposed) weight copies for the TMA-fed tensor-core GEMMs; offsets in floats into act_ws
  struct EaPack { int64_t wi, wj, w2, wiT, wjT, w2T; };
  struct TagPack { int64_t w[kGemmMaxItems], wT[kGemmMaxItems]; };
  std::vector<EaPack> ea_pack;
  std::vector<TagPack> tag_pack;
  int64_t mask_w1 = 0, mask_w2 = 0, mask_w2T = 0;
  // every packed weight whose row pitch is ldh sits in one contiguous arena (one TMA tensor map covers them all)
  int64_t arena_off = 0, arena_rows = 0;

  int64_t ea_stride() const { return 3 * N * ldh; }                       // Hi, Hj, S
  int64_t xcat_ld() const { return int64_t(d.K + 1) * ldh; }
  int64_t tag_stride() const { return N * xcat_ld() + N * ldh; }           // Xcat, Y
};

int make_plan(const pfn_mpn_desc* desc, int64_t n_nodes, Plan& p) {
  PFN_REQUIRE(desc != nullptr, PFN_E_INVALID, "null model descriptor");
  p.d = *desc;
  const pfn_mpn_desc& d = p.d;
  PFN_REQUIRE(d.efeature_dim == 2, PFN_E_UNSUPPORTED, "efeature_dim must be 2 (got %d)", d.efeature_dim);
  PFN_REQUIRE(d.n_gnn_layers >= 2, PFN_E_UNSUPPORTED,
              "n_gnn_layers must be >= 2 (the reference's single-layer constructor, MPN.py:475-477,489, is "
              "shape-inconsistent)");
  PFN_REQUIRE(d.nfeature_dim > 0 && d.output_dim > 0 && d.hidden_dim > 0 && d.K >= 0 && d.K + 1 <= kGemmMaxItems,
              PFN_E_UNSUPPORTED, "unsupported dims (nfeat=%d out=%d hidden=%d K=%d)", d.nfeature_dim, d.output_dim,
              d.hidden_dim, d.K);
  PFN_REQUIRE(d.dropout_rate >= 0.f && d.dropout_rate < 1.f, PFN_E_INVALID, "dropout_rate out of [0,1)");
  const int h = d.hidden_dim;
  p.layers.clear();
  int pi = 0;
  auto add_ea = [&](int fin, int fout) {
    p.layers.push_back(LayerPlan{true, fin, fout, pi, p.n_ea++, true});
    pi += 4;
  };
  auto add_tag = [&](int fin, int fout) {
    p.layers.push_back(LayerPlan{false, fin, fout, pi, p.n_tag++, true});
    pi += d.K + 2;
  };
  add_ea(d.nfeature_dim, h);  // MPN.py:479-480
  add_tag(h, h);
  for (int l = 0; l < d.n_gnn_layers - 2; ++l) {  // MPN.py:482-484
    add_ea(h, h);
    add_tag(h, h);
  }
  add_ea(h, d.output_dim);  // MPN.py:489
  p.layers.back().act = false;
  p.p_mask = pi;
  p.n_params = pi + 4;
  p.N = n_nodes;
  p.ldh = round_up64(h, 4);
  // activation workspace
  int64_t off = 0;
  auto take = [&](int64_t floats) {
    int64_t o = off;
    off += round_up64(floats, 4);
    return o;
  };
  p.off_tiles = take(n_nodes + 8);  // int32 tile table of the graph-resident route (at most one tile per node)
  p.off_maskf = take(n_nodes * d.nfeature_dim);
  p.off_t1 = take(n_nodes * p.ldh);
  p.off_x0 = take(n_nodes * d.nfeature_dim);
  p.off_ea = take(p.n_ea * p.ea_stride());
  p.off_tag = take(p.n_tag * p.tag_stride());
  auto ld4 = [](int v) { return round_up64(v, 4); };
  p.ea_pack.assign(p.n_ea, Plan::EaPack{});
  p.tag_pack.assign(p.n_tag, Plan::TagPack{});
  // three planes per matrix: fp32 | tf32 hi | tf32 lo.  Two passes: pitch-ldh matrices first (the arena), then the rest.
  struct Req { int64_t* slot; int64_t rows, ld; };
  std::vector<Req> reqs;
  for (const LayerPlan& L : p.layers) {
    if (L.is_ea) {
      Plan::EaPack& e = p.ea_pack[L.slot];
      reqs.push_back(Req{&e.wi, h, ld4(L.fin)});
      reqs.push_back(Req{&e.wj, h, ld4(L.fin)});
      reqs.push_back(Req{&e.w2, L.fout, ld4(h)});
      reqs.push_back(Req{&e.wiT, L.fin, ld4(h)});
      reqs.push_back(Req{&e.wjT, L.fin, ld4(h)});
      reqs.push_back(Req{&e.w2T, h, ld4(L.fout)});
    } else {
      Plan::TagPack& t = p.tag_pack[L.slot];
      for (int k = 0; k <= d.K; ++k) {
        reqs.push_back(Req{&t.w[k], L.fout, ld4(L.fin)});
        reqs.push_back(Req{&t.wT[k], L.fin, ld4(L.fout)});
      }
    }
  }
  reqs.push_back(Req{&p.mask_w1, h, ld4(d.nfeature_dim)});
  reqs.push_back(Req{&p.mask_w2, d.nfeature_dim, ld4(h)});
  reqs.push_back(Req{&p.mask_w2T, h, ld4(d.nfeature_dim)});
  p.arena_off = off;
  for (const Req& r : reqs)
    if (r.ld == p.ldh) *r.slot = take(3 * r.rows * r.ld);
  p.arena_rows = (off - p.arena_off) / p.ldh;
  for (const Req& r : reqs)
    if (r.ld != p.ldh) *r.slot = take(3 * r.rows * r.ld);
  p.act_floats = off;
  // scratch.  Every layer keeps its own gradient buffers (d cur, dHi, dHj, d[x_0..x_K]) because the weight
  // gradients are deferred to ONE grouped launch at the end of the backward pass and read them all.
  off = 0;
  p.off_dx0 = take(n_nodes * d.nfeature_dim);  // first: d loss / d (mask_embd(mask) + x), read back by the Python layer
  p.off_ds = take(n_nodes * p.ldh);
  p.off_dz = take(p.n_ea * n_nodes * p.ldh);
  p.off_dhi = take(p.n_ea * n_nodes * p.ldh);
  p.off_dhj = take(p.n_ea * n_nodes * p.ldh);
  p.off_dxcat = take(p.n_tag * n_nodes * p.xcat_ld());
  size_t part = pfn_ea_bwd_scratch_bytes(h);
  // every split-K weight-gradient problem the backward launches: (rows of dW, cols of dW, problems per launch)
  const int wg[][3] = {{h, h, 1}, {d.output_dim, h, 1},                   // EA  dW2 = G^T S
                       {h, h, 2}, {h, d.nfeature_dim, 2},                 // EA  dWi | dWj
                       {h, h, d.K + 1},                                   // TAG dW_k
                       {d.nfeature_dim, h, 1}, {h, d.nfeature_dim, 1}};   // mask_embd
  for (const auto& w : wg) part = std::max(part, gemm_splitk_scratch_bytes(w[0], w[1], n_nodes, w[2]));
  {
    std::vector<WgradProblem> shapes;
    auto shape = [&](int mo, int ni) {
      WgradProblem w{};
      w.Mo = mo;
      w.Ni = ni;
      shapes.push_back(w);
    };
    for (const LayerPlan& L : p.layers) {
      if (L.is_ea) {
        shape(L.fout, h);
        shape(h, L.fin);
        shape(h, L.fin);
      } else {
        for (int k = 0; k <= d.K; ++k) shape(L.fout, L.fin);
      }
    }
    shape(d.nfeature_dim, h);
    shape(h, d.nfeature_dim);
    // the per-EdgeAggregation dWe partial rows share the buffer: they are consumed before the grouped launch starts
    part = std::max(part, wgrad_group_scratch_bytes(shapes.data(), static_cast<int>(shapes.size()), n_nodes));
  }
  p.part_bytes = static_cast<int64_t>(part);
  p.off_part = take(static_cast<int64_t>((part + 3) / 4));
  p.scratch_floats = off;
  return 0;
}

struct Ctx {
  const Plan& p;
  const float* const* params;
  GraphView g;
  float* act;
  float* scratch;
  cudaStream_t stream;
  bool training;
  float scale;  // 1/(1-p) in training, 1 otherwise
  const uint64_t* seed_device = nullptr;
  // graph-resident route: uniform tiles of `tile_rows` rows, or (n_graphs > 0) variable tiles packed from `graph_ptr`
  int64_t tile_rows = 0, n_graphs = 0;
  const int64_t* graph_ptr = nullptr;
  int* tile_table() const { return reinterpret_cast<int*>(act + p.off_tiles); }
  int64_t n_tiles() const { return n_graphs > 0 ? n_graphs : ceil_div64(p.N, std::max<int64_t>(tile_rows, 1)); }

  float* hi(int slot) const { return act + p.off_ea + slot * p.ea_stride(); }
  float* hj(int slot) const { return hi(slot) + p.N * p.ldh; }
  float* s(int slot) const { return hi(slot) + 2 * p.N * p.ldh; }
  float* xcat(int slot) const { return act + p.off_tag + slot * p.tag_stride(); }
  float* ytag(int slot) const { return xcat(slot) + p.N * p.xcat_ld(); }
};

GemmArgs base_args(int M, int N) {
  GemmArgs a{};
  a.M = M;
  a.N = N;
  a.n_items = 1;
  a.splitk = 1;
  a.scale = 1.f;
  return a;
}

// Y = X W^T view helpers -----------------------------------------------------------------------------
// W is a packed weight [n_out, ldw] followed by its TF32 hi / lo planes (k_pack_weights); w_rows = n_out
inline GemmItem fwd_item(const float* X, int64_t ldx, const float* W, int64_t ldw, int K, float* C, int64_t ldc,
                         const float* bias, int64_t w_rows) {
  return GemmItem{X, W, C, bias, nullptr, ldx, 1, 1, ldw, K, static_cast<int>(ldc), w_rows * ldw};
}
inline GemmItem wgrad_item(const float* dY, int64_t lddy, const float* X, int64_t ldx, int K, float* dW, int64_t lddw,
                           float* dbias) {
  return GemmItem{dY, X, dW, nullptr, dbias, 1, lddy, ldx, 1, K, static_cast<int>(lddw), 0};
}

void set_activation(GemmArgs& a, const Ctx& c, bool act, int layer_index, uint64_t seed, const float* inj, int64_t ld_inj) {
  if (!act) {
    a.act = PFN_ACT_NONE;
    return;
  }
  if (c.training && c.p.d.dropout_rate > 0.f) {
    a.act = PFN_ACT_DROPOUT_RELU;
    a.scale = c.scale;
    a.inj = inj;
    a.ld_inj = static_cast<int>(ld_inj);
    const uint64_t host_seed = c.seed_device != nullptr ? 0 : seed;  // the device seed is XORed in by the kernel
    a.seed_lo = static_cast<uint32_t>(host_seed) ^ (0x9E3779B9u * static_cast<uint32_t>(layer_index + 1));
    a.seed_hi = static_cast<uint32_t>(host_seed >> 32);
    a.seed_dev = reinterpret_cast<const uint32_t*>(c.seed_device);
    a.keep_thresh = keep_threshold(c.p.d.dropout_rate);
  } else {
    a.act = PFN_ACT_RELU;
  }
}

// ---- forward ------------------------------------------------------------------------------------------
// The whole forward as ONE launch of the graph-resident kernel (fused_fwd.cu); fills the same activation workspace the
// layer-wise backward reads.  The packed weights must already be in place (k_pack_weights).
int forward_fused(const Ctx& c, const float* x, const int64_t* pred_mask, uint64_t seed, const float* const* inj_masks,
                  float* out, int64_t tile_rows) {
  if (c.n_graphs > 0) {  // variable-size tiles: build the table once per forward (the backward reuses it)
    PFN_CUDA_OK(launch_kernel(k_tile_table, dim3(1), dim3(32), 0, c.stream, c.graph_ptr, static_cast<int>(c.n_graphs),
                              static_cast<int>(c.p.N), c.tile_table(), c.g.meta));
    PFN_LAUNCHED();
  }
  const Plan& p = c.p;
  const pfn_mpn_desc& d = p.d;
  const int h = d.hidden_dim;
  const int n_layers = static_cast<int>(p.layers.size());
  PFN_REQUIRE(n_layers <= kFusedMaxLayers, PFN_E_UNSUPPORTED, "fused forward: too many layers (%d)", n_layers);
  FusedArgs a;
  std::memset(&a, 0, sizeof(a));
  auto arena_row = [&](int64_t off) { return static_cast<int>((off - p.arena_off) / p.ldh); };
  const bool dropout = c.training && d.dropout_rate > 0.f;
  for (int li = 0; li < n_layers; ++li) {
    const LayerPlan& L = p.layers[li];
    const float* const* lp = c.params + L.p0;
    FLayer& f = a.layers[li];
    const bool last = li == n_layers - 1;
    f.last = last ? 1 : 0;
    f.act = L.act ? 1 : 0;
    f.fin = L.fin;
    f.w_rows = h;
    f.seed_xor = 0x9E3779B9u * static_cast<uint32_t>(li + 1);
    f.inj = (inj_masks != nullptr && L.act && dropout) ? inj_masks[li] : nullptr;
    if (L.is_ea) {
      const Plan::EaPack& pk = p.ea_pack[L.slot];
      f.type = L.fin == h ? kFusedEaTc : kFusedEaSimt;
      PFN_REQUIRE(f.type == kFusedEaTc || L.fin == d.nfeature_dim, PFN_E_UNSUPPORTED, "fused forward: unexpected layer width");
      if (f.type == kFusedEaTc) {
        f.w_row[0] = arena_row(pk.wi);
        f.w_row[1] = arena_row(pk.wj);
      }
      f.w_row[2] = arena_row(pk.w2);
      f.W1 = lp[0];
      f.b1 = lp[1];
      f.W2 = lp[2];
      f.b2 = lp[3];
      f.save0 = c.hi(L.slot);
      if (last) {
        f.dest = out;
        f.ld_dest = L.fout;
      } else {
        f.dest = c.xcat(p.layers[li + 1].slot);
        f.ld_dest = static_cast<int>(p.xcat_ld());
      }
    } else {
      PFN_REQUIRE(!last && L.fin == h && L.fout == h, PFN_E_UNSUPPORTED, "fused forward: unexpected TAGConv shape");
      const Plan::TagPack& tk = p.tag_pack[L.slot];
      f.type = kFusedTag;
      for (int k = 0; k <= d.K; ++k) f.w_row[k] = arena_row(tk.w[k]);
      f.bias = lp[d.K + 1];
      f.save0 = c.xcat(L.slot);
      f.dest = c.ytag(L.slot);
      f.ld_dest = static_cast<int>(p.ldh);
    }
  }
  a.n_layers = n_layers;
  a.n_nodes = static_cast<int>(p.N);
  a.tile_rows = static_cast<int>(tile_rows);
  a.n_tiles = static_cast<int>(c.n_tiles());
  a.tile_start = c.n_graphs > 0 ? c.tile_table() : nullptr;
  a.h = h;
  a.K = d.K;
  a.ldh = static_cast<int>(p.ldh);
  a.dropout = dropout ? 1 : 0;
  a.out_dim = d.output_dim;
  a.x = x;
  a.pred_mask = pred_mask;
  const float* const* mp = c.params + p.p_mask;
  a.mW1 = mp[0];
  a.mb1 = mp[1];
  a.mW2 = mp[2];
  a.mb2 = mp[3];
  a.maskf = c.act + p.off_maskf;
  a.t1 = c.act + p.off_t1;
  a.x0 = c.act + p.off_x0;
  a.rowptr = c.g.rowptr_t;
  a.nbr = c.g.nbr_t;
  a.ea = reinterpret_cast<const float2*>(c.g.ea_t);
  a.deg = c.g.deg;
  a.dis = c.g.dis;
  a.meta = c.g.meta;
  a.out = out;
  a.scale = c.scale;
  const uint64_t host_seed = c.seed_device != nullptr ? 0 : seed;
  a.seed_lo = static_cast<uint32_t>(host_seed);
  a.seed_hi = static_cast<uint32_t>(host_seed >> 32);
  a.keep_thresh = keep_threshold(d.dropout_rate);
  a.seed_dev = reinterpret_cast<const uint32_t*>(c.seed_device);
  return fused_fwd_launch(a, c.act + p.arena_off, p.arena_rows, c.stream);
}

int forward_impl(const Ctx& c, const float* x, const int64_t* pred_mask, uint64_t seed, const float* const* inj_masks,
                 float* out, int64_t tile_rows) {
  const Plan& p = c.p;
  const pfn_mpn_desc& d = p.d;
  const int N = static_cast<int>(p.N), h = d.hidden_dim, nf = d.nfeature_dim;
  const int64_t ldh = p.ldh;
  if (N == 0) return 0;
  float* maskf = c.act + p.off_maskf;
  float* t1 = c.act + p.off_t1;
  float* x0 = c.act + p.off_x0;
  const int ld_h = static_cast<int>(ldh), ld_nf = static_cast<int>(round_up64(nf, 4));
  // one launch: copy every weight into 16-byte-pitched K-major buffers (+ transposes for the data gradients)
  {
    std::vector<PackDesc> packs;
    for (const LayerPlan& L : p.layers) {
      const float* const* lp = c.params + L.p0;
      const int ld_fin = static_cast<int>(round_up64(L.fin, 4)), ld_fout = static_cast<int>(round_up64(L.fout, 4));
      if (L.is_ea) {
        const Plan::EaPack& e = p.ea_pack[L.slot];
        const int ldw1 = 2 * L.fin + 2;
        packs.push_back(PackDesc{lp[0], c.act + e.wi, c.act + e.wiT, ldw1, h, L.fin, ld_fin, ld_h});
        packs.push_back(PackDesc{lp[0] + L.fin, c.act + e.wj, c.act + e.wjT, ldw1, h, L.fin, ld_fin, ld_h});
        packs.push_back(PackDesc{lp[2], c.act + e.w2, c.act + e.w2T, h, L.fout, h, ld_h, ld_fout});
      } else {
        const Plan::TagPack& t = p.tag_pack[L.slot];
        for (int k = 0; k <= d.K; ++k)
          packs.push_back(PackDesc{lp[k], c.act + t.w[k], c.act + t.wT[k], L.fin, L.fout, L.fin, ld_fin, ld_fout});
      }
    }
    const float* const* mp = c.params + p.p_mask;
    packs.push_back(PackDesc{mp[0], c.act + p.mask_w1, nullptr, nf, h, nf, ld_nf, 0});
    packs.push_back(PackDesc{mp[2], c.act + p.mask_w2, c.act + p.mask_w2T, h, nf, h, ld_h, ld_nf});
    PFN_TRY(pack_weights_launch(packs.data(), static_cast<int>(packs.size()), c.stream));
  }
  if (tile_rows > 0) return forward_fused(c, x, pred_mask, seed, inj_masks, out, tile_rows);
  // mask_embd (MPN.py:533,537): x0 = Linear(ReLU(Linear(mask.float()))) + x
  {
    const int64_t n = int64_t(N) * nf;
    PFN_CUDA_OK(launch_kernel(k_i64_to_f32, dim3(static_cast<int>(std::min<int64_t>(ceil_div64(n, 256), 1184))), dim3(256), 0, c.stream, pred_mask, maskf, n));
    PFN_LAUNCHED();
    const float* const* mp = c.params + p.p_mask;
    GemmArgs a = base_args(N, h);
    a.it[0] = fwd_item(maskf, nf, c.act + p.mask_w1, ld_nf, nf, t1, ldh, mp[1], h);
    a.act = PFN_ACT_RELU;
    PFN_TRY(gemm_launch(a, true, true, c.stream));
    GemmArgs b = base_args(N, nf);
    b.it[0] = fwd_item(t1, ldh, c.act + p.mask_w2, ld_h, h, x0, nf, mp[3], nf);
    b.addend = x;
    b.ld_add = nf;
    PFN_TRY(gemm_launch(b, true, true, c.stream));
  }
  const float* cur = x0;
  int64_t ldcur = nf;
  const int n_layers = static_cast<int>(p.layers.size());
  for (int li = 0; li < n_layers; ++li) {
    const LayerPlan& L = p.layers[li];
    const float* const* lp = c.params + L.p0;
    const bool last = li == n_layers - 1;
    const float* inj = (inj_masks != nullptr && L.act) ? inj_masks[li] : nullptr;
    const int ld_fin = static_cast<int>(round_up64(L.fin, 4));
    if (L.is_ea) {
      const int ldw1 = 2 * L.fin + 2;
      const Plan::EaPack& pk = p.ea_pack[L.slot];
      // Hi = cur Wi^T + b1 ; Hj = cur Wj^T   (one launch, two problems)
      GemmArgs a = base_args(N, h);
      a.n_items = 2;
      a.batched = 1;
      a.it[0] = fwd_item(cur, ldcur, c.act + pk.wi, ld_fin, L.fin, c.hi(L.slot), ldh, lp[1], h);
      a.it[1] = fwd_item(cur, ldcur, c.act + pk.wj, ld_fin, L.fin, c.hj(L.slot), ldh, nullptr, h);
      PFN_TRY(gemm_launch(a, true, true, c.stream));
      PFN_TRY(ea_fwd_launch(c.hi(L.slot), c.hj(L.slot), ldh, c.g, N, lp[0] + 2 * L.fin, ldw1, c.s(L.slot), ldh, h,
                            c.stream));
      // out = S W2^T + deg (.) b2, then dropout+ReLU unless this is the last layer
      float* dest;
      int64_t lddest;
      if (last) {
        dest = out;
        lddest = L.fout;
      } else {
        dest = c.xcat(p.layers[li + 1].slot);  // block 0 of the next TAGConv's [x_0 | x_1 | ... | x_K]
        lddest = p.xcat_ld();
      }
      GemmArgs b = base_args(N, L.fout);
      b.it[0] = fwd_item(c.s(L.slot), ldh, c.act + pk.w2, ld_h, h, dest, lddest, lp[3], L.fout);
      b.rowscale = c.g.deg;
      set_activation(b, c, L.act, li, seed, inj, h);
      PFN_TRY(gemm_launch(b, true, true, c.stream));
      cur = dest;
      ldcur = lddest;
    } else {
      float* xc = c.xcat(L.slot);
      const int64_t ldx = p.xcat_ld();
      for (int k = 1; k <= d.K; ++k)
        PFN_TRY(hop_launch(xc + (k - 1) * ldh, ldx, c.g, N, false, nullptr, 0, nullptr, 0, 1.f, xc + k * ldh, ldx, L.fin,
                           c.stream));
      float* dest = last ? out : c.ytag(L.slot);
      const int64_t lddest = last ? L.fout : ldh;
      GemmArgs a = base_args(N, L.fout);
      a.n_items = d.K + 1;
      const Plan::TagPack& tk = p.tag_pack[L.slot];
      for (int k = 0; k <= d.K; ++k)
        a.it[k] = fwd_item(xc + k * ldh, ldx, c.act + tk.w[k], ld_fin, L.fin, dest, lddest, lp[d.K + 1], L.fout);
      set_activation(a, c, L.act, li, seed, inj, h);
      PFN_TRY(gemm_launch(a, true, true, c.stream));
      cur = dest;
      ldcur = lddest;
    }
  }
  return 0;
}

// ---- backward -----------------------------------------------------------------------------------------
// Backward of one TAGConv through the graph-resident kernel (fused_fwd.cu, mode 1): d x_0 = sum_k ((A_hat^T)^k G) W_k,
// masked by the layer input -- one launch instead of a 4-problem GEMM + K hop launches.
void fill_tag_backward_layer(const Ctx& c, const LayerPlan& L, const float* G, int64_t ldG, const float* xc, float* dest,
                             FLayer& f) {
  const Plan& p = c.p;
  const pfn_mpn_desc& d = p.d;
  const Plan::TagPack& tk = p.tag_pack[L.slot];
  f.type = kFusedTag;
  f.fin = L.fin;
  f.w_rows = d.hidden_dim;
  for (int k = 0; k <= d.K; ++k) f.w_row[k] = static_cast<int>((tk.wT[k] - p.arena_off) / p.ldh);
  f.dest = dest;
  f.ld_dest = static_cast<int>(p.xcat_ld());
  f.gin = G;
  f.ld_gin = static_cast<int>(ldG);
  f.ymask = xc;
  f.ld_ymask = static_cast<int>(p.xcat_ld());
}

void fill_backward_common(const Ctx& c, FusedArgs& a, int mode, int n_layers, int64_t tile_rows) {
  const Plan& p = c.p;
  const pfn_mpn_desc& d = p.d;
  a.mode = mode;
  a.n_layers = n_layers;
  a.n_nodes = static_cast<int>(p.N);
  a.tile_rows = static_cast<int>(tile_rows);
  a.n_tiles = static_cast<int>(c.n_tiles());
  a.tile_start = c.n_graphs > 0 ? c.tile_table() : nullptr;
  a.h = d.hidden_dim;
  a.K = d.K;
  a.ldh = static_cast<int>(p.ldh);
  a.out_dim = d.output_dim;
  a.rowptr = c.g.rowptr_s;
  a.nbr = c.g.nbr_s;
  a.ea = reinterpret_cast<const float2*>(c.g.ea_s);
  a.rowptr2 = c.g.rowptr_t;
  a.nbr2 = c.g.nbr_t;
  a.ea2 = reinterpret_cast<const float2*>(c.g.ea_t);
  a.deg = c.g.deg;
  a.dis = c.g.dis;
  a.meta = c.g.meta;
  a.scale = c.scale;
}

int tag_backward_fused(const Ctx& c, const LayerPlan& L, const float* G, int64_t ldG, const float* xc, float* dest,
                       int64_t tile_rows) {
  FusedArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_tag_backward_layer(c, L, G, ldG, xc, dest, a.layers[0]);
  fill_backward_common(c, a, kFusedModeTagBackward, 1, tile_rows);
  return fused_fwd_launch(a, c.act + c.p.arena_off, c.p.arena_rows, c.stream);
}

// Backward of one EdgeAggregation through the graph-resident kernel (fused_fwd.cu, mode 2): dS = G W2, both segmented
// passes (dHj by source, dHi + dWe by target, ReLU mask recomputed from the saved Hi / Hj) and d cur = dHj Wj + dHi Wi in
// ONE launch (+ the tiny dWe reduction) instead of two GEMMs and the two-pass edge kernel.
void fill_ea_backward_layer(const Ctx& c, const LayerPlan& L, bool last, const float* const* lp, const float* G, int64_t ldG,
                            const float* ymask, int64_t ld_ymask, float* dest, int64_t ld_dest, float* dhi, float* dhj,
                            float* part, FLayer& f) {
  const Plan& p = c.p;
  const int h = p.d.hidden_dim;
  const Plan::EaPack& pk = p.ea_pack[L.slot];
  auto arena_row = [&](int64_t off) { return static_cast<int>((off - p.arena_off) / p.ldh); };
  f.type = L.fin == h ? kFusedEaTc : kFusedEaSimt;
  f.last = last ? 1 : 0;
  f.fin = L.fin;
  f.w_rows = h;
  if (f.type == kFusedEaTc) {  // transposed packed weights: [fin, ldh] each
    f.w_row[0] = arena_row(pk.wiT);
    f.w_row[1] = arena_row(pk.wjT);
  }
  if (!last) f.w_row[2] = arena_row(pk.w2T);  // [h, ld(fout)] with fout = h
  f.W1 = lp[0];
  f.W2 = lp[2];
  f.save0 = c.hi(L.slot);
  f.dest = dest;
  f.ld_dest = static_cast<int>(ld_dest);
  f.gin = G;
  f.ld_gin = static_cast<int>(ldG);
  f.ymask = ymask;
  f.ld_ymask = static_cast<int>(ld_ymask);
  f.dhi = dhi;
  f.dhj = dhj;
  f.dwe_partial = part;
}

int ea_backward_fused(const Ctx& c, const LayerPlan& L, bool last, const float* const* lp, const float* G, int64_t ldG,
                      const float* ymask, int64_t ld_ymask, float* dest, int64_t ld_dest, float* dhi, float* dhj, float* part,
                      float* dWe, int64_t tile_rows) {
  const Plan& p = c.p;
  FusedArgs a;
  std::memset(&a, 0, sizeof(a));
  fill_ea_backward_layer(c, L, last, lp, G, ldG, ymask, ld_ymask, dest, ld_dest, dhi, dhj, part, a.layers[0]);
  fill_backward_common(c, a, kFusedModeEaBackward, 1, tile_rows);
  PFN_TRY(fused_fwd_launch(a, c.act + p.arena_off, p.arena_rows, c.stream));
  return reduce_dwe_launch(part, a.n_tiles, p.d.hidden_dim, dWe, 2 * L.fin + 2, c.stream);
}

int backward_impl(const Ctx& c, float* const* grads, const float* dout, int64_t tile_rows) {
  const Plan& p = c.p;
  const pfn_mpn_desc& d = p.d;
  const int N = static_cast<int>(p.N), h = d.hidden_dim, nf = d.nfeature_dim;
  const int64_t ldh = p.ldh;
  if (N == 0) return 0;
  float* ds = c.scratch + p.off_ds;
  float* dx0 = c.scratch + p.off_dx0;
  const int64_t nld = int64_t(N) * ldh;
  std::vector<GemmArgs> deferred;  // weight-gradient problems, launched together at the end
  // Whole backward data path in ONE launch of the graph-resident kernel (mode 3) when every layer fits it; PFN_BWD_CHAIN=0
  // keeps one launch per layer (modes 1 / 2).
  const char* chain_env = std::getenv("PFN_BWD_CHAIN");  // read per call: the tests compare both launch structures
  const bool chain_off = chain_env != nullptr && chain_env[0] == '0';
  bool chain = tile_rows > 0 && !chain_off && static_cast<int>(p.layers.size()) <= kFusedMaxLayers;
  for (const LayerPlan& L : p.layers) chain = chain && (!L.is_ea || L.fin == h || L.fin == nf);
  FusedArgs ch;
  std::memset(&ch, 0, sizeof(ch));
  int n_chain = 0;
  const int64_t dwe_stride = 2 * 4 * ((h + 3) / 4) * c.n_tiles();  // floats per EA layer
  float* part = c.scratch + p.off_part;
  const float* x0 = c.act + p.off_x0;
  const float* G = dout;  // gradient w.r.t. the current layer's (pre-activation) output
  int64_t ldG = d.output_dim;
  const int n_layers = static_cast<int>(p.layers.size());
  for (int li = n_layers - 1; li >= 0; --li) {
    const LayerPlan& L = p.layers[li];
    const float* const* lp = c.params + L.p0;
    float* const* lg = grads + L.p0;
    if (L.is_ea) {
      const int ldw1 = 2 * L.fin + 2;
      float* dz = c.scratch + p.off_dz + L.slot * nld;
      float* dhi = c.scratch + p.off_dhi + L.slot * nld;
      float* dhj = c.scratch + p.off_dhj + L.slot * nld;
      // input of this layer and whether it is the (post-activation) output of a previous layer
      const float* cur;
      int64_t ldcur;
      bool cur_has_act;
      if (li == 0) {
        cur = x0;
        ldcur = nf;
        cur_has_act = false;
      } else {
        cur = c.ytag(p.layers[li - 1].slot);
        ldcur = ldh;
        cur_has_act = true;
      }
      // dW2 = G^T S ; db2 = sum_i deg_i G[i]
      {
        GemmArgs a = base_args(L.fout, h);
        a.it[0] = wgrad_item(G, ldG, c.s(L.slot), ldh, N, lg[2], h, lg[3]);
        a.extra_col = 2;
        a.extra_vec = c.g.deg;
        a.partial = part;
        a.partial_bytes = static_cast<size_t>(p.part_bytes);
        gemm_plan_splitk(a, N, 1);
        deferred.push_back(a);
      }
      const bool ea_fused = chain || (tile_rows > 0 && ldG % 4 == 0 && (L.fin == h || L.fin == nf));
      const Plan::EaPack& pk = p.ea_pack[L.slot];
      if (chain) {
        float* dest = li == 0 ? dx0 : dz;
        const int64_t lddest = li == 0 ? nf : ldh;
        fill_ea_backward_layer(c, L, li == n_layers - 1, lp, G, ldG, cur_has_act ? cur : nullptr, ldcur, dest, lddest, dhi, dhj,
                               part + L.slot * dwe_stride, ch.layers[n_chain++]);
      } else if (ea_fused) {
        float* dest = li == 0 ? dx0 : dz;
        const int64_t lddest = li == 0 ? nf : ldh;
        PFN_TRY(ea_backward_fused(c, L, li == n_layers - 1, lp, G, ldG, cur_has_act ? cur : nullptr, ldcur, dest, lddest, dhi, dhj,
                                  part, lg[0] + 2 * L.fin, tile_rows));
      } else {
      // dS = G W2   (as G (W2^T)^T with the packed transpose: both operands K-major)
      {
        GemmArgs a = base_args(N, h);
        a.it[0] = fwd_item(G, ldG, c.act + pk.w2T, round_up64(L.fout, 4), L.fout, ds, ldh, nullptr, h);
        a.prof_cat = PFN_PROF_GEMM_DGRAD + 1;
        PFN_TRY(gemm_launch(a, true, true, c.stream));
      }
      // dHi, dHj, dWe
      PFN_TRY(ea_bwd_launch(ds, ldh, c.hi(L.slot), c.hj(L.slot), ldh, c.g, N, lp[0] + 2 * L.fin, ldw1, dhi, dhj, ldh,
                            lg[0] + 2 * L.fin, ldw1, part, h, c.stream));
      }
      // dWi = dHi^T cur (+ db1 = colsum dHi) ; dWj = dHj^T cur
      {
        GemmArgs a = base_args(h, L.fin);
        a.n_items = 2;
        a.batched = 1;
        a.it[0] = wgrad_item(dhi, ldh, cur, ldcur, N, lg[0], ldw1, lg[1]);
        a.it[1] = wgrad_item(dhj, ldh, cur, ldcur, N, lg[0] + L.fin, ldw1, nullptr);
        a.extra_col = 1;
        a.partial = part;
        a.partial_bytes = static_cast<size_t>(p.part_bytes);
        gemm_plan_splitk(a, N, 2);
        deferred.push_back(a);
      }
      // d cur = dHi Wi + dHj Wj, masked by the previous layer's activation
      if (!ea_fused) {
        float* dest = li == 0 ? dx0 : dz;
        const int64_t lddest = li == 0 ? nf : ldh;
        GemmArgs a = base_args(N, L.fin);
        a.n_items = 2;
        a.it[0] = fwd_item(dhi, ldh, c.act + pk.wiT, ldh, h, dest, lddest, nullptr, L.fin);
        a.it[1] = fwd_item(dhj, ldh, c.act + pk.wjT, ldh, h, dest, lddest, nullptr, L.fin);
        a.prof_cat = PFN_PROF_GEMM_DGRAD + 1;
        if (cur_has_act) {
          a.act = kActMaskByY;
          a.ymask = cur;
          a.ld_ym = static_cast<int>(ldcur);
          a.scale = c.scale;
        }
        PFN_TRY(gemm_launch(a, true, true, c.stream));
        G = dest;
        ldG = lddest;
      }
      if (ea_fused) {
        G = li == 0 ? dx0 : dz;
        ldG = li == 0 ? nf : ldh;
      }
    } else {
      const float* xc = c.xcat(L.slot);
      const int64_t ldx = p.xcat_ld();
      float* dxcat = c.scratch + p.off_dxcat + L.slot * int64_t(N) * ldx;
      // dW_k = G^T x_k (k = 0..K) ; dbias = colsum G
      {
        GemmArgs a = base_args(L.fout, L.fin);
        a.n_items = d.K + 1;
        a.batched = 1;
        for (int k = 0; k <= d.K; ++k)
          a.it[k] = wgrad_item(G, ldG, xc + k * ldh, ldx, N, lg[k], L.fin, k == 0 ? lg[d.K + 1] : nullptr);
        a.extra_col = 1;
        a.partial = part;
        a.partial_bytes = static_cast<size_t>(p.part_bytes);
        gemm_plan_splitk(a, N, d.K + 1);
        deferred.push_back(a);
      }
      if (chain) {
        fill_tag_backward_layer(c, L, G, ldG, xc, dxcat, ch.layers[n_chain++]);
      } else if (tile_rows > 0 && ldG % 4 == 0) {
        PFN_TRY(tag_backward_fused(c, L, G, ldG, xc, dxcat, tile_rows));
      } else {
        // d x_k = G W_k (k = 0..K), then the transposed hop chain d x_{k-1} += A_hat^T d x_k; the last hop
        // also applies the activation mask of the layer input (which is block 0 of xcat itself)
        {
          GemmArgs a = base_args(N, L.fin);
          a.n_items = d.K + 1;
          a.batched = 1;
          const Plan::TagPack& tk = p.tag_pack[L.slot];
          for (int k = 0; k <= d.K; ++k)
            a.it[k] = fwd_item(G, ldG, c.act + tk.wT[k], round_up64(L.fout, 4), L.fout, dxcat + k * ldh, ldx, nullptr, L.fin);
          a.prof_cat = PFN_PROF_GEMM_DGRAD + 1;
          if (d.K == 0) {
            a.act = kActMaskByY;
            a.ymask = xc;
            a.ld_ym = static_cast<int>(ldx);
            a.scale = c.scale;
          }
          PFN_TRY(gemm_launch(a, true, true, c.stream));
        }
        for (int k = d.K; k >= 1; --k) {
          const bool final_hop = k == 1;
          PFN_TRY(hop_launch(dxcat + k * ldh, ldx, c.g, N, true, dxcat + (k - 1) * ldh, ldx, final_hop ? xc : nullptr, ldx,
                             c.scale, dxcat + (k - 1) * ldh, ldx, L.fin, c.stream));
        }
      }
      G = dxcat;
      ldG = ldx;
    }
  }
  if (chain) {
    fill_backward_common(c, ch, kFusedModeBackward, n_chain, tile_rows);
    // the last step (first layer) also produces d t1 of mask_embd: d t1 = (d x0 W2m) * (t1 > 0)
    ch.mW2 = c.params[p.p_mask + 2];
    ch.t1 = c.act + p.off_t1;
    ch.dt1 = ds;
    PFN_TRY(fused_fwd_launch(ch, c.act + p.arena_off, p.arena_rows, c.stream));
    std::vector<const float*> parts;
    std::vector<float*> dwes;
    std::vector<int> lds;
    for (const LayerPlan& L : p.layers)
      if (L.is_ea) {
        parts.push_back(part + L.slot * dwe_stride);
        dwes.push_back(grads[L.p0] + 2 * L.fin);
        lds.push_back(2 * L.fin + 2);
      }
    PFN_TRY(reduce_dwe_multi_launch(parts.data(), dwes.data(), lds.data(), static_cast<int>(parts.size()), ch.n_tiles, h, c.stream));
  }
  // mask_embd backward: x0 = W2m relu(W1m mask + b1m) + b2m + x
  {
    float* const* mg = grads + p.p_mask;
    const float* maskf = c.act + p.off_maskf;
    const float* t1 = c.act + p.off_t1;
    GemmArgs a = base_args(nf, h);
    a.it[0] = wgrad_item(G, ldG, t1, ldh, N, mg[2], h, mg[3]);
    a.extra_col = 1;
    a.partial = part;
    a.partial_bytes = static_cast<size_t>(p.part_bytes);
    gemm_plan_splitk(a, N, 1);
    deferred.push_back(a);
    if (!chain) {  // (the chained tile kernel has already written d t1 into `ds`)
      GemmArgs b = base_args(N, h);
      b.it[0] = fwd_item(G, ldG, c.act + p.mask_w2T, round_up64(nf, 4), nf, ds, ldh, nullptr, h);
      b.prof_cat = PFN_PROF_GEMM_DGRAD + 1;
      b.act = kActMaskByY;
      b.ymask = t1;
      b.ld_ym = static_cast<int>(ldh);
      b.scale = 1.f;
      PFN_TRY(gemm_launch(b, true, true, c.stream));
    }
    GemmArgs w = base_args(h, nf);
    w.it[0] = wgrad_item(ds, ldh, maskf, nf, N, mg[0], nf, mg[1]);
    w.extra_col = 1;
    w.partial = part;
    w.partial_bytes = static_cast<size_t>(p.part_bytes);
    gemm_plan_splitk(w, N, 1);
    deferred.push_back(w);
  }
  // ---- every weight gradient of the step: one grouped tensor-core launch + one reduction (gemm_tc.cu) ----
  {
    std::vector<WgradProblem> probs;
    for (const GemmArgs& a : deferred) {
      const int count = a.batched ? a.n_items : 1;
      for (int i = 0; i < count; ++i) {
        const GemmItem& it = a.it[i];
        probs.push_back(WgradProblem{it.A, it.a_cs, it.B, it.b_rs, a.M, a.N, it.C, it.ldc, it.bias_out, a.extra_col, a.extra_vec});
      }
    }
    int rc;
    {
      ProfScope prof(PFN_PROF_GEMM_WGRAD, c.stream);
      rc = wgrad_group_launch(probs.data(), static_cast<int>(probs.size()), N, part, static_cast<size_t>(p.part_bytes), c.stream);
    }
    if (rc == 1) {  // outside the grouped kernel's range (e.g. hidden_dim > 255): one launch per problem, as before
      for (const GemmArgs& a : deferred) PFN_TRY(gemm_launch(a, false, false, c.stream));
    } else if (rc != 0) {
      return rc;
    }
  }
  return 0;
}

int check_tables(const Plan& p, const float* const* params, const char* what) {
  PFN_REQUIRE(params != nullptr, PFN_E_INVALID, "%s: null pointer table", what);
  for (int i = 0; i < p.n_params; ++i) PFN_REQUIRE(params[i] != nullptr, PFN_E_INVALID, "%s: entry %d is null", what, i);
  return 0;
}

}  // namespace
}  // namespace pfn

using namespace pfn;

extern "C" const char* pfn_version(void) { return "pfn_b200 0.1.0 (sm_100a)"; }
extern "C" const char* pfn_last_error(void) { return g_error; }
extern "C" uint64_t pfn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int pfn_profile_enable(int on) {
  g_prof.on = on != 0;
  if (on) g_prof.used = 0;
  return 0;
}

extern "C" int pfn_profile_read(int category, double* total_ms, int64_t* launches) {
  PFN_REQUIRE(total_ms && launches && category >= 0 && category < PFN_PROF_CATEGORIES, PFN_E_INVALID,
              "pfn_profile_read: bad arguments");
  double total = 0.0;
  int64_t n = 0;
  for (size_t i = 0; i < g_prof.used; ++i) {
    if (g_prof.cat[i] != category) continue;
    PFN_CUDA_OK(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
    float ms = 0.f;
    PFN_CUDA_OK(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    total += ms;
    ++n;
  }
  *total_ms = total;
  *launches = n;
  return 0;
}

extern "C" int pfn_mpn_num_params(const pfn_mpn_desc* desc) {
  Plan p;
  if (make_plan(desc, 0, p) != 0) return -1;
  return p.n_params;
}

extern "C" int pfn_mpn_workspace(const pfn_mpn_desc* desc, int64_t n_nodes, int64_t e_raw, size_t* act_bytes,
                                 size_t* scratch_bytes) {
  (void)e_raw;
  PFN_REQUIRE(act_bytes && scratch_bytes && n_nodes >= 0, PFN_E_INVALID, "pfn_mpn_workspace: bad arguments");
  Plan p;
  PFN_TRY(make_plan(desc, n_nodes, p));
  *act_bytes = size_t(p.act_floats) * sizeof(float) + 16;
  *scratch_bytes = size_t(p.scratch_floats) * sizeof(float) + 16;
  return 0;
}

extern "C" int pfn_mpn_fused_supported(const pfn_mpn_desc* desc, int64_t tile_rows) {
  if (desc == nullptr || desc->efeature_dim != 2 || desc->n_gnn_layers < 2) return 0;
  if (2 * desc->n_gnn_layers - 1 > kFusedMaxLayers) return 0;
  return fused_fwd_supported(desc->hidden_dim, desc->K, desc->nfeature_dim, desc->output_dim, tile_rows) ? 1 : 0;
}

extern "C" int pfn_graph_tile_status(const void* graph_ws, int32_t* violated, void* stream_) {
  PFN_REQUIRE(graph_ws != nullptr && violated != nullptr, PFN_E_INVALID, "pfn_graph_tile_status: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  pfn_graph_layout lay;
  pfn_graph_layout_get(0, 0, &lay);
  const int32_t* meta = reinterpret_cast<const int32_t*>(static_cast<const char*>(graph_ws) + lay.meta);
  int32_t v = 0;
  PFN_CUDA_OK(cudaMemcpyAsync(&v, meta + 6, sizeof(v), cudaMemcpyDeviceToHost, stream));
  PFN_CUDA_OK(cudaStreamSynchronize(stream));
  *violated = v;
  return 0;
}

static int mpn_forward_common(const pfn_mpn_desc* desc, const float* const* params, const float* x,
                              const int64_t* pred_mask, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                              void* act_ws, void* scratch_ws, int training, uint64_t seed, const uint64_t* seed_device,
                              const float* const* inj_masks, float* out, int64_t tile_rows, const int64_t* graph_ptr,
                              int64_t n_graphs, void* stream);

extern "C" int pfn_mpn_forward_tiled(const pfn_mpn_desc* desc, const float* const* params, const float* x,
                                     const int64_t* pred_mask, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                                     void* act_ws, void* scratch_ws, int training, uint64_t seed,
                                     const uint64_t* seed_device, const float* const* inj_masks, float* out,
                                     int64_t tile_rows, const int64_t* graph_ptr, int64_t n_graphs, void* stream) {
  PFN_REQUIRE(tile_rows > 0 && pfn_mpn_fused_supported(desc, tile_rows), PFN_E_UNSUPPORTED,
              "pfn_mpn_forward_tiled: configuration outside the graph-resident kernel (use pfn_mpn_forward)");
  PFN_REQUIRE(graph_ptr == nullptr || (n_graphs > 0 && n_graphs <= n_nodes), PFN_E_INVALID, "pfn_mpn_forward_tiled: bad graph_ptr / n_graphs");
  return mpn_forward_common(desc, params, x, pred_mask, n_nodes, e_raw, graph_ws, act_ws, scratch_ws, training, seed,
                            seed_device, inj_masks, out, tile_rows, graph_ptr, graph_ptr != nullptr ? n_graphs : 0, stream);
}

extern "C" int pfn_mpn_forward(const pfn_mpn_desc* desc, const float* const* params, const float* x,
                               const int64_t* pred_mask, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                               void* act_ws, void* scratch_ws, int training, uint64_t seed, const uint64_t* seed_device,
                               const float* const* inj_masks, float* out, void* stream) {
  return mpn_forward_common(desc, params, x, pred_mask, n_nodes, e_raw, graph_ws, act_ws, scratch_ws, training, seed,
                            seed_device, inj_masks, out, 0, nullptr, 0, stream);
}

static int mpn_forward_common(const pfn_mpn_desc* desc, const float* const* params, const float* x,
                              const int64_t* pred_mask, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                              void* act_ws, void* scratch_ws, int training, uint64_t seed, const uint64_t* seed_device,
                              const float* const* inj_masks, float* out, int64_t tile_rows, const int64_t* graph_ptr,
                              int64_t n_graphs, void* stream) {
  Plan p;
  PFN_TRY(make_plan(desc, n_nodes, p));
  PFN_TRY(check_tables(p, params, "pfn_mpn_forward(params)"));
  PFN_REQUIRE(n_nodes == 0 || (x && pred_mask && out), PFN_E_INVALID, "pfn_mpn_forward: null tensor");
  PFN_REQUIRE(graph_ws && act_ws && aligned16(act_ws), PFN_E_INVALID, "pfn_mpn_forward: workspace null or misaligned");
  (void)scratch_ws;
  Ctx c{p, params, graph_view(graph_ws, n_nodes, e_raw), static_cast<float*>(act_ws), static_cast<float*>(scratch_ws),
        static_cast<cudaStream_t>(stream), training != 0,
        (training != 0 && desc->dropout_rate > 0.f) ? 1.f / (1.f - desc->dropout_rate) : 1.f};
  c.seed_device = seed_device;
  c.tile_rows = tile_rows;
  c.graph_ptr = graph_ptr;
  c.n_graphs = n_graphs;
  return forward_impl(c, x, pred_mask, seed, inj_masks, out, tile_rows);
}

static int mpn_backward_common(const pfn_mpn_desc* desc, const float* const* params, float* const* grads,
                               const float* dout, int64_t n_nodes, int64_t e_raw, const void* graph_ws, void* act_ws,
                               void* scratch_ws, int training, int64_t tile_rows, int64_t n_graphs, void* stream);

extern "C" int pfn_mpn_backward(const pfn_mpn_desc* desc, const float* const* params, float* const* grads,
                                const float* dout, int64_t n_nodes, int64_t e_raw, const void* graph_ws, void* act_ws,
                                void* scratch_ws, int training, void* stream) {
  return mpn_backward_common(desc, params, grads, dout, n_nodes, e_raw, graph_ws, act_ws, scratch_ws, training, 0, 0, stream);
}

extern "C" int pfn_mpn_backward_tiled(const pfn_mpn_desc* desc, const float* const* params, float* const* grads,
                                      const float* dout, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                                      void* act_ws, void* scratch_ws, int training, int64_t tile_rows, int64_t n_graphs,
                                      void* stream) {
  PFN_REQUIRE(tile_rows > 0 && pfn_mpn_fused_supported(desc, tile_rows), PFN_E_UNSUPPORTED,
              "pfn_mpn_backward_tiled: configuration outside the graph-resident kernel (use pfn_mpn_backward)");
  return mpn_backward_common(desc, params, grads, dout, n_nodes, e_raw, graph_ws, act_ws, scratch_ws, training, tile_rows,
                             n_graphs, stream);
}

static int mpn_backward_common(const pfn_mpn_desc* desc, const float* const* params, float* const* grads,
                               const float* dout, int64_t n_nodes, int64_t e_raw, const void* graph_ws, void* act_ws,
                               void* scratch_ws, int training, int64_t tile_rows, int64_t n_graphs, void* stream) {
  Plan p;
  PFN_TRY(make_plan(desc, n_nodes, p));
  PFN_TRY(check_tables(p, params, "pfn_mpn_backward(params)"));
  PFN_TRY(check_tables(p, const_cast<const float* const*>(grads), "pfn_mpn_backward(grads)"));
  PFN_REQUIRE(n_nodes == 0 || dout, PFN_E_INVALID, "pfn_mpn_backward: null dout");
  PFN_REQUIRE(graph_ws && act_ws && scratch_ws && aligned16(act_ws) && aligned16(scratch_ws), PFN_E_INVALID,
              "pfn_mpn_backward: workspace null or misaligned");
  Ctx c{p, params, graph_view(graph_ws, n_nodes, e_raw), static_cast<float*>(act_ws), static_cast<float*>(scratch_ws),
        static_cast<cudaStream_t>(stream), training != 0,
        (training != 0 && desc->dropout_rate > 0.f) ? 1.f / (1.f - desc->dropout_rate) : 1.f};
  if (n_nodes == 0) {  // gradients of an empty batch are zero
    return 0;
  }
  c.tile_rows = tile_rows;
  c.n_graphs = n_graphs;  // > 0: the variable-size tile table the forward left in the activation workspace
  return backward_impl(c, grads, dout, tile_rows);
}

extern "C" size_t pfn_mse_scratch_bytes(int64_t count) { return size_t(mse_blocks(count)) * sizeof(float); }

extern "C" int pfn_mse_fwd_bwd(const float* out, const float* y, int64_t count, float inv_count, float* loss,
                               float* dout, void* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_REQUIRE(out && y && loss && dout && scratch && count > 0, PFN_E_INVALID, "pfn_mse_fwd_bwd: bad arguments");
  const int blocks = mse_blocks(count);
  PFN_CUDA_OK(launch_kernel(k_mse_partial, dim3(blocks), dim3(kMseBlock), 0, stream, out, y, count, inv_count, dout, static_cast<float*>(scratch)));
  PFN_LAUNCHED();
  PFN_CUDA_OK(launch_kernel(k_mse_final, dim3(1), dim3(kMseBlock), 0, stream, static_cast<const float*>(scratch), blocks, inv_count, loss));
  PFN_LAUNCHED();
  return 0;
}

extern "C" size_t pfn_masked_l2_scratch_bytes(int64_t count) { return size_t(3 * mse_blocks(count) + 2) * sizeof(float); }

extern "C" int pfn_masked_l2_fwd_bwd(const float* out, const float* y, const int64_t* mask, int64_t count, int regularize,
                                     float regcoeff, const float* global_counts, float* loss, float* dout, void* scratch,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PFN_REQUIRE(out && y && mask && loss && dout && scratch && count > 0, PFN_E_INVALID, "pfn_masked_l2_fwd_bwd: bad arguments");
  const int blocks = mse_blocks(count);
  float* partial = static_cast<float*>(scratch);
  float* scales = partial + 3 * blocks;
  PFN_CUDA_OK(launch_kernel(k_ml2_partial, dim3(blocks), dim3(kMseBlock), 0, stream, out, y, mask, count, partial));
  PFN_LAUNCHED();
  PFN_CUDA_OK(launch_kernel(k_ml2_final, dim3(1), dim3(kMseBlock), 0, stream, static_cast<const float*>(partial), blocks, count, regularize,
                            regcoeff, global_counts, loss, scales));
  PFN_LAUNCHED();
  PFN_CUDA_OK(launch_kernel(k_ml2_grad, dim3(blocks), dim3(kMseBlock), 0, stream, out, y, mask, count, static_cast<const float*>(scales), dout));
  PFN_LAUNCHED();
  return 0;
}
