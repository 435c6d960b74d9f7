// Shared helpers for libpfn_b200 (sm_100a).  Internal header; the public ABI is include/pfn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <atomic>

#include "../../include/pfn_b200.h"

namespace pfn {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

#define PFN_CUDA_OK(expr)                                                                       \
  do {                                                                                          \
    cudaError_t pfn_e_ = (expr);                                                                \
    if (pfn_e_ != cudaSuccess) {                                                                \
      ::pfn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(pfn_e_)); \
      return static_cast<int>(pfn_e_);                                                          \
    }                                                                                           \
  } while (0)

#define PFN_LAUNCHED()                   \
  do {                                   \
    ::pfn::count_launch();               \
    PFN_CUDA_OK(cudaGetLastError());     \
  } while (0)

#define PFN_REQUIRE(cond, code, ...)     \
  do {                                   \
    if (!(cond)) {                       \
      ::pfn::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define PFN_TRY(expr)                    \
  do {                                   \
    int pfn_rc_ = (expr);                \
    if (pfn_rc_ != 0) return pfn_rc_;    \
  } while (0)

// ---- launches: programmatic dependent launch (PDL) ------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with
// griddepcontrol.wait: the next kernel's launch latency and prologue (barrier init, TMEM allocation, tensor-map
// prefetch, index arithmetic) overlap the tail of the previous one, which matters when a step is 13-90 kernels of
// 5-20 us.  PFN_PDL=0 restores plain stream-ordered launches.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#if defined(__CUDACC__)
// wait until the preceding kernel(s) in the stream have completed and their writes are visible (no-op without PDL)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// The opt-in dynamic shared-memory limit (cudaFuncAttributeMaxDynamicSharedMemorySize) is a PER-DEVICE function
// attribute: a process that drives several GPUs must set it on each of them.  `done` holds one bit per device ordinal;
// the check is a cudaGetDevice + an atomic load on the hot path.  Thread-safe (setting the attribute twice is harmless).
struct SmemAttrOnce {
  std::atomic<uint64_t> done[4] = {};  // device ordinals 0..255
  bool needs(int dev) const { return (done[(dev >> 6) & 3].load(std::memory_order_acquire) & (1ull << (dev & 63))) == 0; }
  void mark(int dev) { done[(dev >> 6) & 3].fetch_or(1ull << (dev & 63), std::memory_order_release); }
};
template <typename F>
inline cudaError_t ensure_dynamic_smem(SmemAttrOnce& once, F set_all) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!once.needs(dev)) return cudaSuccess;
  e = set_all();
  if (e == cudaSuccess) once.mark(dev);
  return e;
}

// RAII bracket used by the launch sites; a no-op unless pfn_profile_enable(1) was called.
struct ProfScope {
  int slot;
  cudaStream_t stream;
  ProfScope(int category, cudaStream_t s);
  ~ProfScope();
};

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- graph workspace view -------------------------------------------------------------------
struct GraphView {
  int32_t* meta;
  int32_t* rowptr_t;
  int32_t* nbr_t;
  int32_t* eid_t;
  float* ea_t;
  int32_t* rowptr_s;
  int32_t* nbr_s;
  int32_t* eid_s;
  float* ea_s;
  float* deg;
  float* dis;
  int32_t* cursor;
  int64_t e_cap;
};

inline GraphView graph_view(const void* base, int64_t n_nodes, int64_t e_raw) {
  pfn_graph_layout lay;
  pfn_graph_layout_get(n_nodes, e_raw, &lay);
  char* b = const_cast<char*>(static_cast<const char*>(base));
  GraphView v;
  v.meta = reinterpret_cast<int32_t*>(b + lay.meta);
  v.rowptr_t = reinterpret_cast<int32_t*>(b + lay.rowptr_t);
  v.nbr_t = reinterpret_cast<int32_t*>(b + lay.nbr_t);
  v.eid_t = reinterpret_cast<int32_t*>(b + lay.eid_t);
  v.ea_t = reinterpret_cast<float*>(b + lay.ea_t);
  v.rowptr_s = reinterpret_cast<int32_t*>(b + lay.rowptr_s);
  v.nbr_s = reinterpret_cast<int32_t*>(b + lay.nbr_s);
  v.eid_s = reinterpret_cast<int32_t*>(b + lay.eid_s);
  v.ea_s = reinterpret_cast<float*>(b + lay.ea_s);
  v.deg = reinterpret_cast<float*>(b + lay.deg);
  v.dis = reinterpret_cast<float*>(b + lay.dis);
  v.cursor = reinterpret_cast<int32_t*>(b + lay.cursor);
  v.e_cap = lay.e_cap;
  return v;
}

// ---- dense GEMM (gemm.cu) -------------------------------------------------------------------
// C(m,n) = sum over items/k of A(m,k) * B(k,n) with fully general element strides:
//   A(m,k) = A[m*a_rs + k*a_cs],  B(k,n) = B[k*b_rs + n*b_cs].
constexpr int kGemmMaxItems = 8;

struct GemmItem {
  const float* A;
  const float* B;
  float* C;           // batched mode: per-problem output
  const float* bias;  // batched mode: per-problem bias (may be null)
  float* bias_out;    // split-K weight-gradient mode: where the extra column (bias gradient) goes
  long long a_rs, a_cs, b_rs, b_cs;
  int K;
  int ldc;
  long long b_plane;  // != 0: B is followed by its pre-split TF32 planes: B + b_plane = hi, B + 2*b_plane = lo
};

struct GemmArgs {
  GemmItem it[kGemmMaxItems];
  int n_items;
  int batched;  // 0: items are K-segments accumulated into it[0].C ; 1: independent problems (blockIdx.z)
  int M, N;
  // epilogue (shared)
  const float* rowscale;  // bias term becomes rowscale[m] * bias[n]
  const float* addend;
  int ld_add;
  int act;  // PFN_ACT_* or 3 = multiply by (ymask > 0 ? scale : 0)
  const float* ymask;
  int ld_ym;
  float scale;
  const float* inj;
  int ld_inj;
  uint32_t seed_lo, seed_hi, keep_thresh;  // keep iff hash >= keep_thresh
  const uint32_t* seed_dev;                 // optional device-resident 64-bit seed, XORed into (seed_lo, seed_hi)
  // split-K (weight gradients): partial[(split*count + z)][M][N + extra]
  int splitk;
  int kchunk;
  float* partial;
  size_t partial_bytes;    // capacity of `partial` (the tensor-core weight-gradient path picks its own split count)
  int extra_col;           // 0 none, 1 B gets a virtual extra column of ones, 2 extra column = extra_vec[k]
  const float* extra_vec;
  int prof_cat;            // PFN_PROF_* + 1 to attribute the launch to, 0 = derive from the operand layout
};

constexpr int kActMaskByY = 3;

int gemm_launch(const GemmArgs& args, bool a_kcontig, bool b_kcontig, cudaStream_t stream);
size_t gemm_splitk_scratch_bytes(int64_t M, int64_t N, int64_t K, int count);
// fills splitk/kchunk for a weight-gradient GEMM with reduction length K and `count` problems
void gemm_plan_splitk(GemmArgs& args, int64_t K, int count);

// ---- tensor-core path (gemm_tc.cu) ---------------------------------------------------------------
// 0 = launched on tcgen05 ; 1 = shape/alignment outside that path, caller falls back to the FFMA kernel ; else error
int gemm_tc_launch(const GemmArgs& args, cudaStream_t stream);
int wgrad_tc_launch(GemmArgs& args, cudaStream_t stream);  // may rewrite args.splitk / args.kchunk
bool tc_enabled();
// one weight-gradient problem of a backward pass: dW[Mo, Ni] = dY^T X over `nodes` rows (+ bias gradient as an extra
// column of X: extra_col 1 = ones, 2 = extra_vec[node]); see wgrad_group_launch in gemm_tc.cu
struct WgradProblem {
  const float* dY;  // [nodes, Mo] row-major, pitch lddy
  long long lddy;
  const float* X;   // [nodes, Ni] row-major, pitch ldx
  long long ldx;
  int Mo, Ni;
  float* dW;        // [Mo, Ni] with row pitch lddw
  int lddw;
  float* dbias;     // [Mo] or null
  int extra_col;
  const float* extra_vec;
};
size_t wgrad_group_scratch_bytes(const WgradProblem* probs, int n, int64_t nodes);
int wgrad_group_launch(const WgradProblem* probs, int n, int64_t nodes, float* partial, size_t partial_bytes, cudaStream_t stream);
struct PackDesc {
  const float* src;  // [rows, cols] with row pitch ld_src (a state_dict weight or a column block of one)
  float* dst;        // 3 planes of [rows, ld_dst]:   fp32 copy | rn_tf32(w) | rn_tf32(w - hi)      (may be null)
  float* dst_t;      // 3 planes of [cols, ld_dst_t]: the same for the transpose                      (may be null)
  int ld_src, rows, cols, ld_dst, ld_dst_t;
};
int pack_weights_launch(const PackDesc* items, int n, cudaStream_t stream);

// ---- hash-based dropout (shared by gemm.cu epilogue) ------------------------------------------
__host__ __device__ inline uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__host__ __device__ inline uint32_t dropout_hash(uint32_t row, uint32_t col, uint32_t seed_lo, uint32_t seed_hi) {
  return lowbias32(lowbias32(row * 0x9E3779B1U ^ seed_lo) ^ (col * 0x85EBCA77U + seed_hi));
}
inline uint32_t keep_threshold(float p) {
  if (p <= 0.f) return 0u;
  double t = static_cast<double>(p) * 4294967296.0;
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  return static_cast<uint32_t>(t);
}

// ---- edge kernels (edge_kernels.cu) -----------------------------------------------------------
int ea_fwd_launch(const float* Hi, const float* Hj, int64_t ldh, const GraphView& g, int64_t n_nodes,
                  const float* We, int64_t ldwe, float* S, int64_t lds, int64_t h, cudaStream_t stream);
int ea_bwd_launch(const float* dS, int64_t ldds, const float* Hi, const float* Hj, int64_t ldh,
                  const GraphView& g, int64_t n_nodes, const float* We, int64_t ldwe, float* dHi, float* dHj,
                  int64_t ldd, float* dWe, int64_t lddwe, void* scratch, int64_t h, cudaStream_t stream);
int reduce_dwe_launch(const float* partial, int nblocks, int64_t h, float* dWe, int64_t lddwe, cudaStream_t stream);
int reduce_dwe_multi_launch(const float* const* partial, float* const* dwe, const int* lddwe, int n, int nblocks, int64_t h,
                            cudaStream_t stream);
int hop_launch(const float* X, int64_t ldx, const GraphView& g, int64_t n_nodes, bool transpose,
               const float* addend, int64_t ldadd, const float* ymask, int64_t ldym, float scale, float* Y,
               int64_t ldy, int64_t h, cudaStream_t stream);

}  // namespace pfn
