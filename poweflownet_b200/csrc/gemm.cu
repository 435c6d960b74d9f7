// Dense per-node Linear stacks (nn.Linear / PyG Linear of networks/MPN.py:17-21,491-495 and
// TAGConv.lins) in fp32: forward  Y = X W^T (+bias, +epilogue), data gradient dX = dY W and
// weight gradient dW = dY^T X (split-K over the node dimension, deterministic two-pass reduction).
//
// Accuracy contract: outputs within 1e-5 relative of the fp32 reference (BASELINE.json north_star),
// which rules out single-pass TF32/BF16 tensor-core math; this first implementation therefore uses
// fp32 FFMA register tiles (16x16 threads, TM x TN accumulators each, strided so that any width --
// hidden_dim is 129 in configs/standard.json -- maps onto 16*TM x 16*TN tiles with little waste).
// A split-precision tcgen05 path is the planned replacement for the large shapes.
//
// One kernel covers every call site through element strides (no operand is ever repacked):
//   A(m,k) = A[m*a_rs + k*a_cs],  B(k,n) = B[k*b_rs + n*b_cs];
// template flags only say which index is contiguous so that global loads coalesce.
// Epilogue fusions: bias (optionally row-scaled: the deg (.) b2 term of the hoisted second Linear
// of EdgeAggregation), residual add (mask_embd, MPN.py:537), dropout+ReLU (MPN.py:546-547) and the
// ReLU/dropout backward mask.  Several K-segments can accumulate into one tile (TAGConv's
// sum_k lins[k](x_k); dX = dHi Wi + dHj Wj) and several independent problems can share a launch.
#include <algorithm>

#include "common.cuh"

namespace pfn {
namespace {

constexpr int kBK = 16;

template <int TM, int TN, bool AK, bool BKC>
__global__ void __launch_bounds__(256) k_gemm(const __grid_constant__ GemmArgs args) {
  pdl_wait();
  constexpr int BM = 16 * TM, BN = 16 * TN;
  __shared__ float As[kBK][BM + 1];
  __shared__ float Bs[kBK][BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = args.M, N = args.N;
  const int n_eff = N + (args.extra_col ? 1 : 0);
  int prob = blockIdx.z, split = 0;
  if (args.splitk > 1) {
    prob = blockIdx.z / args.splitk;
    split = blockIdx.z - prob * args.splitk;
  }
  const int seg_begin = args.batched ? prob : 0;
  const int seg_end = args.batched ? prob + 1 : args.n_items;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[TM], rb[TN];
  auto seg_range = [&](int s, int& kb, int& ke) {
    kb = 0;
    ke = args.it[s].K;
    if (args.splitk > 1) {
      kb = split * args.kchunk;
      ke = min(ke, kb + args.kchunk);
    }
  };
  auto load_tile = [&](int s, int k0, int kend) {
    const GemmItem& it = args.it[s];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int id = tid + 256 * i;
      const int m = AK ? id / kBK : id % BM, k = AK ? id % kBK : id / BM;
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < M && gk < kend) ? __ldg(it.A + gm * it.a_rs + gk * it.a_cs) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int id = tid + 256 * j;
      const int n = BKC ? id / kBK : id % BN, k = BKC ? id % kBK : id / BN;
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gk < kend) {
        if (gn < N)
          v = __ldg(it.B + gk * it.b_rs + gn * it.b_cs);
        else if (gn == N && args.extra_col)
          v = args.extra_col == 1 ? 1.f : __ldg(args.extra_vec + gk);
      }
      rb[j] = v;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int id = tid + 256 * i;
      const int m = AK ? id / kBK : id % BM, k = AK ? id % kBK : id / BM;
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int id = tid + 256 * j;
      const int n = BKC ? id / kBK : id % BN, k = BKC ? id % kBK : id / BN;
      Bs[k][n] = rb[j];
    }
  };

  int seg = seg_begin, k0 = 0, kend = 0;
  seg_range(seg, k0, kend);
  while (seg < seg_end && k0 >= kend) {
    if (++seg < seg_end) seg_range(seg, k0, kend);
  }
  bool have = seg < seg_end;
  if (have) load_tile(seg, k0, kend);
  while (have) {
    __syncthreads();
    store_tile();
    __syncthreads();
    k0 += kBK;
    while (seg < seg_end && k0 >= kend) {
      if (++seg < seg_end) seg_range(seg, k0, kend);
    }
    have = seg < seg_end;
    if (have) load_tile(seg, k0, kend);  // next tile's global loads fly while this one is multiplied
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  // ---- epilogue -----------------------------------------------------------------------------
  const GemmItem& out = args.it[args.batched ? prob : 0];
  const int count = args.batched ? args.n_items : 1;
  uint32_t seed_lo = args.seed_lo, seed_hi = args.seed_hi;
  if (args.act == PFN_ACT_DROPOUT_RELU && args.seed_dev != nullptr) {
    seed_lo ^= args.seed_dev[0];
    seed_hi ^= args.seed_dev[1];
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= n_eff) continue;
      float v = acc[i][j];
      if (args.splitk > 1) {
        args.partial[(size_t(split * count + prob) * M + m) * n_eff + n] = v;
        continue;
      }
      if (n == N) {  // virtual extra column = bias gradient
        if (out.bias_out != nullptr) out.bias_out[m] = v;
        continue;
      }
      if (out.bias != nullptr) v = fmaf(args.rowscale != nullptr ? args.rowscale[m] : 1.f, out.bias[n], v);
      if (args.addend != nullptr) v += args.addend[size_t(m) * args.ld_add + n];
      if (args.act == PFN_ACT_RELU) {
        v = fmaxf(v, 0.f);
      } else if (args.act == PFN_ACT_DROPOUT_RELU) {
        const bool keep = args.inj != nullptr ? args.inj[size_t(m) * args.ld_inj + n] != 0.f
                                              : dropout_hash(m, n, seed_lo, seed_hi) >= args.keep_thresh;
        v = keep ? fmaxf(v * args.scale, 0.f) : 0.f;
      } else if (args.act == kActMaskByY) {
        v = args.ymask[size_t(m) * args.ld_ym + n] > 0.f ? v * args.scale : 0.f;
      }
      out.C[size_t(m) * out.ldc + n] = v;
    }
  }
}

// Second pass of the split-K weight gradient: sum the partial tiles in a fixed order (deterministic).  Eight partials
// are loaded before the first add so the pass is bandwidth- rather than latency-bound.
__global__ void __launch_bounds__(256) k_splitk_reduce(const __grid_constant__ GemmArgs args) {
  pdl_wait();
  const int M = args.M, N = args.N, n_eff = N + (args.extra_col ? 1 : 0);
  const int count = args.batched ? args.n_items : 1;
  const size_t per = size_t(M) * n_eff;
  const size_t total = per * count;
  const int S = args.splitk;
  for (size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
    const int prob = static_cast<int>(idx / per);
    const size_t r = idx - size_t(prob) * per;
    const int m = static_cast<int>(r / n_eff), n = static_cast<int>(r - size_t(m) * n_eff);
    const float* __restrict__ src = args.partial + size_t(prob) * per + r;
    const size_t stride = size_t(count) * per;
    float sum = 0.f;
    int s = 0;
    for (; s + 8 <= S; s += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + size_t(s + u) * stride);
      sum += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    for (; s < S; ++s) sum += __ldg(src + size_t(s) * stride);
    const GemmItem& out = args.it[prob];
    if (n == N) {
      if (out.bias_out != nullptr) out.bias_out[m] = sum;
    } else {
      out.C[size_t(m) * out.ldc + n] = sum;
    }
  }
}

int pick_tile(int64_t dim) {
  if (dim <= 16) return 1;
  if (dim <= 32) return 2;
  if (dim <= 64) return 4;
  const int64_t w8 = ceil_div64(dim, 128) * 128 - dim, w9 = ceil_div64(dim, 144) * 144 - dim;
  return w9 < w8 ? 9 : 8;
}

template <int TM, int TN, bool AK, bool BKC>
int launch_t(const GemmArgs& args, dim3 grid, cudaStream_t stream) {
  PFN_CUDA_OK(launch_kernel(k_gemm<TM, TN, AK, BKC>, grid, dim3(256), 0, stream, args));
  PFN_LAUNCHED();
  return 0;
}

template <int TM, bool AK, bool BKC>
int launch_tn(int tn, const GemmArgs& args, dim3 grid, cudaStream_t stream) {
  switch (tn) {
    case 1: return launch_t<TM, 1, AK, BKC>(args, grid, stream);
    case 2: return launch_t<TM, 2, AK, BKC>(args, grid, stream);
    case 4: return launch_t<TM, 4, AK, BKC>(args, grid, stream);
    case 8: return launch_t<TM, 8, AK, BKC>(args, grid, stream);
    default: return launch_t<TM, 9, AK, BKC>(args, grid, stream);
  }
}

}  // namespace

size_t gemm_splitk_scratch_bytes(int64_t M, int64_t N, int64_t K, int count) {
  size_t worst = 0;
  for (int extra = 0; extra <= 1; ++extra) {  // the plan depends on whether a bias-gradient column rides along
    GemmArgs tmp{};
    tmp.M = static_cast<int>(M);
    tmp.N = static_cast<int>(N);
    tmp.extra_col = extra;
    gemm_plan_splitk(tmp, K, count);
    worst = std::max(worst, size_t(tmp.splitk) * count * size_t(M) * size_t(N + extra) * sizeof(float));
  }
  // the tensor-core weight-gradient path splits so that (splits x problems) <= SM count
  worst = std::max(worst, size_t(sm_count()) * size_t(M) * size_t(N + 1) * sizeof(float));
  return worst;
}

void gemm_plan_splitk(GemmArgs& args, int64_t K, int count) {
  const int tm = pick_tile(args.M), tn = pick_tile(args.N + (args.extra_col ? 1 : 0));
  const int64_t tiles = ceil_div64(args.M, 16 * tm) * ceil_div64(args.N + (args.extra_col ? 1 : 0), 16 * tn) * count;
  int64_t want = std::max<int64_t>(1, (2 * int64_t(sm_count())) / std::max<int64_t>(tiles, 1));
  want = std::min<int64_t>(want, 64);
  int64_t kchunk = round_up64(std::max<int64_t>(ceil_div64(std::max<int64_t>(K, 1), want), kBK), kBK);
  args.kchunk = static_cast<int>(kchunk);
  args.splitk = static_cast<int>(std::max<int64_t>(1, ceil_div64(std::max<int64_t>(K, 1), kchunk)));
}

int gemm_launch(const GemmArgs& args_in, bool a_kcontig, bool b_kcontig, cudaStream_t stream) {
  GemmArgs args = args_in;
  PFN_REQUIRE(args.n_items >= 1 && args.n_items <= kGemmMaxItems, PFN_E_INVALID, "gemm: bad item count %d", args.n_items);
  PFN_REQUIRE(!(args.splitk > 1) || args.partial != nullptr, PFN_E_INVALID, "gemm: split-K without scratch");
  PFN_REQUIRE(!(args.splitk > 1) || args.batched || args.n_items == 1, PFN_E_INVALID, "gemm: split-K with K-segments");
  if (args.M <= 0 || args.N <= 0) return 0;
  const int n_eff = args.N + (args.extra_col ? 1 : 0);
  // forward / data-gradient GEMMs have the (large) node dimension as M; weight gradients have M = n_out
  const int tm = (a_kcontig) ? 8 : pick_tile(args.M);
  const int tn = pick_tile(n_eff);
  const int count = args.batched ? args.n_items : 1;
  dim3 grid(static_cast<unsigned>(ceil_div64(n_eff, 16 * tn)), static_cast<unsigned>(ceil_div64(args.M, 16 * tm)),
            static_cast<unsigned>(count * std::max(args.splitk, 1)));
  ProfScope prof(args.prof_cat > 0 ? args.prof_cat - 1
                                   : (a_kcontig ? (b_kcontig ? PFN_PROF_GEMM_FWD : PFN_PROF_GEMM_DGRAD) : PFN_PROF_GEMM_WGRAD),
                 stream);
  int rc;
  if (a_kcontig && b_kcontig) {  // both operands K-major: tcgen05 path when the operands are TMA-addressable
    rc = gemm_tc_launch(args, stream);
    if (rc != 1) return rc;
  }
  if (!a_kcontig && !b_kcontig && args.partial != nullptr) {  // weight gradient: tcgen05 with MN-major operands
    rc = wgrad_tc_launch(args, stream);
    if (rc == 0) {
      const size_t total = size_t(args.M) * n_eff * count;
      const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, size_t(sm_count()) * 8));
      PFN_CUDA_OK(launch_kernel(k_splitk_reduce, dim3(blocks), dim3(256), 0, stream, args));
      PFN_LAUNCHED();
      return 0;
    }
    if (rc != 1) return rc;
  }
  if (a_kcontig && b_kcontig) {
    rc = launch_tn<8, true, true>(tn, args, grid, stream);
  } else if (a_kcontig && !b_kcontig) {
    rc = launch_tn<8, true, false>(tn, args, grid, stream);
  } else if (!a_kcontig && !b_kcontig) {
    switch (tm) {
      case 1: rc = launch_tn<1, false, false>(tn, args, grid, stream); break;
      case 2: rc = launch_tn<2, false, false>(tn, args, grid, stream); break;
      case 4: rc = launch_tn<4, false, false>(tn, args, grid, stream); break;
      case 8: rc = launch_tn<8, false, false>(tn, args, grid, stream); break;
      default: rc = launch_tn<9, false, false>(tn, args, grid, stream); break;
    }
  } else {
    set_error("gemm: operand layout (A m-contiguous, B k-contiguous) is not used by the path");
    return PFN_E_UNSUPPORTED;
  }
  PFN_TRY(rc);
  if (args.splitk > 1) {
    const size_t total = size_t(args.M) * n_eff * count;
    const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, size_t(sm_count()) * 8));
    PFN_CUDA_OK(launch_kernel(k_splitk_reduce, dim3(blocks), dim3(256), 0, stream, args));
    PFN_LAUNCHED();
  }
  return 0;
}

}  // namespace pfn

using namespace pfn;

// ---- C ABI wrappers ------------------------------------------------------------------------------
extern "C" int pfn_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                              const float* rowscale, const float* addend, int64_t ldadd, float* Y, int64_t ldy,
                              int64_t M, int64_t n_in, int64_t n_out, int act, float dropout_p, uint64_t seed,
                              const float* inj_mask, int64_t ld_inj, void* stream) {
  PFN_REQUIRE(X && W && Y, PFN_E_INVALID, "pfn_linear_fwd: null argument");
  PFN_REQUIRE(act >= PFN_ACT_NONE && act <= PFN_ACT_DROPOUT_RELU, PFN_E_INVALID, "pfn_linear_fwd: bad act %d", act);
  PFN_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, PFN_E_INVALID, "pfn_linear_fwd: dropout_p out of [0,1)");
  GemmArgs a{};
  a.n_items = 1;
  a.it[0] = GemmItem{X, W, Y, bias, nullptr, ldx, 1, 1, ldw, static_cast<int>(n_in), static_cast<int>(ldy), 0};
  a.M = static_cast<int>(M);
  a.N = static_cast<int>(n_out);
  a.rowscale = rowscale;
  a.addend = addend;
  a.ld_add = static_cast<int>(ldadd);
  a.act = act;
  a.scale = 1.f / (1.f - dropout_p);
  a.inj = inj_mask;
  a.ld_inj = static_cast<int>(ld_inj);
  a.seed_lo = static_cast<uint32_t>(seed);
  a.seed_hi = static_cast<uint32_t>(seed >> 32);
  a.keep_thresh = keep_threshold(dropout_p);
  a.splitk = 1;
  return gemm_launch(a, true, true, static_cast<cudaStream_t>(stream));
}

extern "C" int pfn_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, const float* ymask,
                                int64_t ldym, float scale, float* dX, int64_t lddx, int64_t M, int64_t n_in,
                                int64_t n_out, void* stream) {
  PFN_REQUIRE(dY && W && dX, PFN_E_INVALID, "pfn_linear_dgrad: null argument");
  GemmArgs a{};
  a.n_items = 1;
  // dX(m, i) = sum_o dY(m, o) W(o, i):  B(k=o, n=i) = W[o*ldw + i]
  a.it[0] = GemmItem{dY, W, dX, nullptr, nullptr, lddy, 1, ldw, 1, static_cast<int>(n_out), static_cast<int>(lddx), 0};
  a.M = static_cast<int>(M);
  a.N = static_cast<int>(n_in);
  a.act = ymask != nullptr ? kActMaskByY : PFN_ACT_NONE;
  a.ymask = ymask;
  a.ld_ym = static_cast<int>(ldym);
  a.scale = scale;
  a.splitk = 1;
  return gemm_launch(a, true, false, static_cast<cudaStream_t>(stream));
}

extern "C" size_t pfn_linear_wgrad_scratch_bytes(int64_t M, int64_t n_in, int64_t n_out) {
  return gemm_splitk_scratch_bytes(n_out, n_in, M, 1) + 256;
}

extern "C" int pfn_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* rowscale,
                                float* dW, int64_t lddw, float* dbias, int64_t M, int64_t n_in, int64_t n_out,
                                void* scratch, void* stream) {
  PFN_REQUIRE(dY && X && dW && scratch, PFN_E_INVALID, "pfn_linear_wgrad: null argument");
  GemmArgs a{};
  a.n_items = 1;
  // dW(o, i) = sum_m dY(m, o) X(m, i):  A(m'=o, k=m) = dY[m*lddy + o],  B(k=m, n=i) = X[m*ldx + i]
  a.it[0] = GemmItem{dY, X, dW, nullptr, dbias, 1, lddy, ldx, 1, static_cast<int>(M), static_cast<int>(lddw), 0};
  a.M = static_cast<int>(n_out);
  a.N = static_cast<int>(n_in);
  a.extra_col = dbias != nullptr ? (rowscale != nullptr ? 2 : 1) : 0;
  a.extra_vec = rowscale;
  a.partial = static_cast<float*>(scratch);
  a.partial_bytes = pfn_linear_wgrad_scratch_bytes(M, n_in, n_out);
  gemm_plan_splitk(a, M, 1);
  return gemm_launch(a, false, false, static_cast<cudaStream_t>(stream));
}
