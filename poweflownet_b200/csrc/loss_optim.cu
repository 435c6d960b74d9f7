// libpfn_b200 -- the two neighbours of the hot path that SURVEY.md section 8(f) ranks next:
//
//   PowerImbalance loss  (utils/custom_loss_functions.py:99-286; `--train_loss_fn power_imbalance`, train.py:95-101)
//     a second gather -> per-branch math -> segmented sum workload on the same CSR the model uses.  The reference runs
//     it as index_select x2, ~40 elementwise launches, scatter_add_ and an autograd tape of the same length; here the
//     loss AND its gradient w.r.t. the predictions come from three launches without atomics:
//       k_pi_node   one thread per bus i: walks its row of the CSR by source (flow='target_to_source': the aggregating
//                   bus is edge_index[0]), sums the branch injections in the reference's edge order, writes
//                   (dP_i, dQ_i) and a per-CTA partial of sum_i dP_i^2 + dQ_i^2
//       k_pi_final  loss = sum / N
//       k_pi_grad   one thread per bus v: the analytic gradient -- its own row of the CSR by source (v aggregates) plus
//                   its row of the CSR by target (v is the neighbour; needs (dP, dQ) of the aggregating bus)
//     The forward arithmetic mirrors the reference's operation order with explicit round-to-nearest mul/add (no FMA
//     contraction): the injections cancel catastrophically (e_i e_j - e_i^2 ...), so a fused multiply-add would move the
//     result by more than the 1e-5 contract allows on its own.
//
//   AdamW  (train.py:123 `torch.optim.AdamW`; torch/optim/adamw.py `_single_tensor_adamw`)
//     one launch over every parameter tensor (pointer table in the kernel parameter block) instead of torch's
//     per-dtype foreach chains.
#include <math.h>

#include "common.cuh"

namespace pfn {
namespace {

constexpr int kPiBlock = 128;
constexpr float kDegToRad = 0.017453292519943295f;  // float(1/180.*pi), custom_loss_functions.py:191

struct PiStats {
  float xm[4], xs[4], em[2], es[2];
};

struct Bus {
  float e, f, c, s;  // rectangular voltage (e, f) = vm (cos va, sin va); c, s kept for the gradient
};

// x * std + mean, vm e^{j va}: de_normalize (:124-129) then :190-197
__device__ __forceinline__ Bus bus_state(float4 xn, const PiStats& st) {
  const float vm = __fadd_rn(__fmul_rn(xn.x, st.xs[0]), st.xm[0]);
  const float va = __fmul_rn(kDegToRad, __fadd_rn(__fmul_rn(xn.y, st.xs[1]), st.xm[1]));
  Bus b;
  sincosf(va, &b.s, &b.c);
  b.e = __fmul_rn(vm, b.c);
  b.f = __fmul_rn(vm, b.s);
  return b;
}

// series admittance of a branch from its normalised (r, x): :180-181
__device__ __forceinline__ void branch_gb(float2 a, const PiStats& st, float& g, float& b) {
  const float r = __fadd_rn(__fmul_rn(a.x, st.es[0]), st.em[0]);
  const float x = __fadd_rn(__fmul_rn(a.y, st.es[1]), st.em[1]);
  const float den = __fadd_rn(__fmul_rn(r, r), __fmul_rn(x, x));
  g = __fdiv_rn(r, den);
  b = __fdiv_rn(-x, den);
}

// (Pji, Qji) of :216-217, evaluated left to right as torch does
__device__ __forceinline__ float2 injection(const Bus& i, const Bus& j, float g, float b) {
  float t = __fmul_rn(i.e, j.e);
  t = __fadd_rn(t, -__fmul_rn(i.e, i.e));
  t = __fadd_rn(t, __fmul_rn(i.f, j.f));
  t = __fadd_rn(t, -__fmul_rn(i.f, i.f));
  const float cross = __fadd_rn(__fmul_rn(i.f, j.e), -__fmul_rn(i.e, j.f));
  float u = __fmul_rn(-i.e, j.e);
  u = __fadd_rn(u, __fmul_rn(i.e, i.e));
  u = __fadd_rn(u, -__fmul_rn(i.f, j.f));
  u = __fadd_rn(u, __fmul_rn(i.f, i.f));
  float2 pq;
  pq.x = __fadd_rn(__fmul_rn(g, t), __fmul_rn(b, cross));
  pq.y = __fadd_rn(__fmul_rn(g, cross), __fmul_rn(b, u));
  return pq;
}

__device__ __forceinline__ float4 load_bus(const float* __restrict__ x, int64_t ldx, int node) {
  return *reinterpret_cast<const float4*>(x + int64_t(node) * ldx);
}

__global__ void __launch_bounds__(kPiBlock)
k_pi_node(const float* __restrict__ x, int64_t ldx, const int* __restrict__ rowptr_s, const int* __restrict__ nbr_s,
          const float2* __restrict__ ea_s, PiStats st, int n_nodes, float2* __restrict__ dpq, float* __restrict__ partial) {
  pdl_wait();
  const int i = blockIdx.x * kPiBlock + threadIdx.x;
  float local = 0.f;
  if (i < n_nodes) {
    const float4 xi = load_bus(x, ldx, i);
    const Bus bi = bus_state(xi, st);
    float agg_p = 0.f, agg_q = 0.f;
    const int beg = rowptr_s[i], fin = rowptr_s[i + 1];
    for (int e = beg; e < fin; ++e) {  // CSR order = edge order of the (doubled) list = the order scatter_add_ sums in
      const Bus bj = bus_state(load_bus(x, ldx, nbr_s[e]), st);
      float g, b;
      branch_gb(ea_s[e], st, g, b);
      const float2 pq = injection(bi, bj, g, b);
      agg_p = __fadd_rn(agg_p, pq.x);
      agg_q = __fadd_rn(agg_q, pq.y);
    }
    // update (:229-252): dP = -aggregated + P_i with the de-normalised P, Q of the bus
    const float dp = __fadd_rn(-agg_p, __fadd_rn(__fmul_rn(xi.z, st.xs[2]), st.xm[2]));
    const float dq = __fadd_rn(-agg_q, __fadd_rn(__fmul_rn(xi.w, st.xs[3]), st.xm[3]));
    dpq[i] = make_float2(dp, dq);
    local = __fadd_rn(__fmul_rn(dp, dp), __fmul_rn(dq, dq));
  }
  __shared__ float red[kPiBlock];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = kPiBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(kPiBlock)
k_pi_final(const float* __restrict__ partial, int n, float inv_nodes, float* __restrict__ loss) {
  pdl_wait();
  __shared__ float red[kPiBlock];
  float local = 0.f;
  for (int i = threadIdx.x; i < n; i += kPiBlock) local += partial[i];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = kPiBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * inv_nodes;
}

// d loss / d x (normalised predictions).  With a_i = d loss / d aggregated_i = -2 (dP_i, dQ_i) / N:
//   bus v as the aggregating end of branch (v, j):  d/d(e_v, f_v) of (P, Q)_vj
//   bus v as the far end of branch (u, v):          d/d(e_v, f_v) of (P, Q)_uv, weighted with a_u
// then (e, f) = vm (cos va, sin va) back to (vm, va) and through the de-normalisation.
__global__ void __launch_bounds__(kPiBlock)
k_pi_grad(const float* __restrict__ x, int64_t ldx, const int* __restrict__ rowptr_s, const int* __restrict__ nbr_s,
          const float2* __restrict__ ea_s, const int* __restrict__ rowptr_t, const int* __restrict__ nbr_t,
          const float2* __restrict__ ea_t, PiStats st, int n_nodes, float inv_nodes, const float2* __restrict__ dpq,
          float* __restrict__ dx, int64_t lddx) {
  pdl_wait();
  const int v = blockIdx.x * kPiBlock + threadIdx.x;
  if (v >= n_nodes) return;
  const Bus bv = bus_state(load_bus(x, ldx, v), st);
  const float2 dv = dpq[v];
  const float ap = -2.f * dv.x * inv_nodes, aq = -2.f * dv.y * inv_nodes;
  float de = 0.f, df = 0.f;
  for (int e = rowptr_s[v], fin = rowptr_s[v + 1]; e < fin; ++e) {
    const Bus bj = bus_state(load_bus(x, ldx, nbr_s[e]), st);
    float g, b;
    branch_gb(ea_s[e], st, g, b);
    const float p_e = g * (bj.e - 2.f * bv.e) - b * bj.f, p_f = g * (bj.f - 2.f * bv.f) + b * bj.e;
    const float q_e = -g * bj.f + b * (2.f * bv.e - bj.e), q_f = g * bj.e + b * (2.f * bv.f - bj.f);
    de += ap * p_e + aq * q_e;
    df += ap * p_f + aq * q_f;
  }
  for (int e = rowptr_t[v], fin = rowptr_t[v + 1]; e < fin; ++e) {
    const int u = nbr_t[e];
    const Bus bu = bus_state(load_bus(x, ldx, u), st);
    float g, b;
    branch_gb(ea_t[e], st, g, b);
    const float2 du = dpq[u];
    const float up = -2.f * du.x * inv_nodes, uq = -2.f * du.y * inv_nodes;
    const float p_e = g * bu.e + b * bu.f, p_f = g * bu.f - b * bu.e;  // d(P_uv)/d(e_v, f_v); d(Q_uv) = (p_f, -p_e)
    de += up * p_e + uq * p_f;
    df += up * p_f - uq * p_e;
  }
  const float dvm = de * bv.c + df * bv.s;
  const float dva = -de * bv.f + df * bv.e;
  float4 out;
  out.x = dvm * st.xs[0];
  out.y = dva * kDegToRad * st.xs[1];
  out.z = 2.f * dv.x * inv_nodes * st.xs[2];
  out.w = 2.f * dv.y * inv_nodes * st.xs[3];
  *reinterpret_cast<float4*>(dx + int64_t(v) * lddx) = out;
}

// ---- AdamW ------------------------------------------------------------------------------------------------
constexpr int kAdamMaxTensors = 64;
constexpr int kAdamBlock = 256;
struct AdamArgs {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long n[kAdamMaxTensors];
  float decay, beta1_w, beta2, beta2_w, bc2_sqrt, eps, step_size;
};
static_assert(sizeof(AdamArgs) <= 4000, "kernel parameter space");

__global__ void __launch_bounds__(kAdamBlock) k_adamw(const __grid_constant__ AdamArgs a) {
  pdl_wait();
  const int t = blockIdx.y;
  const long long n = a.n[t];
  float* __restrict__ const p = a.p[t];
  const float* __restrict__ const g = a.g[t];
  float* __restrict__ const m = a.m[t];
  float* __restrict__ const v = a.v[t];
  for (long long i = (long long)blockIdx.x * kAdamBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kAdamBlock) {
    const float gi = g[i];
    const float pi = p[i] * a.decay;                       // param.mul_(1 - lr * weight_decay)
    const float mi = m[i] + a.beta1_w * (gi - m[i]);       // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * a.beta2 + a.beta2_w * gi * gi;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(vi) / a.bc2_sqrt + a.eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - a.step_size * (mi / denom);                // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

}  // namespace
}  // namespace pfn

using namespace pfn;

extern "C" size_t pfn_power_imbalance_scratch_bytes(int64_t n_nodes) {
  const int64_t blocks = ceil_div64(n_nodes > 0 ? n_nodes : 1, kPiBlock);
  return size_t(2 * (n_nodes > 0 ? n_nodes : 1) + blocks) * sizeof(float);
}

extern "C" int pfn_power_imbalance_fwd_bwd(const float* x, int64_t ldx, const void* graph_ws, int64_t n_nodes, int64_t e_raw,
                                           const float* stats, float* loss, float* dx, int64_t lddx, void* scratch,
                                           void* stream_) {
  PFN_REQUIRE(x && graph_ws && stats && loss && scratch && n_nodes > 0, PFN_E_INVALID, "pfn_power_imbalance_fwd_bwd: bad arguments");
  PFN_REQUIRE(aligned16(x) && ldx % 4 == 0 && ldx >= 4, PFN_E_INVALID, "pfn_power_imbalance_fwd_bwd: x must be 16-byte aligned rows of >= 4 floats");
  PFN_REQUIRE(dx == nullptr || (aligned16(dx) && lddx % 4 == 0 && lddx >= 4), PFN_E_INVALID, "pfn_power_imbalance_fwd_bwd: dx misaligned");
  PFN_REQUIRE(n_nodes < (int64_t(1) << 31), PFN_E_INVALID, "pfn_power_imbalance_fwd_bwd: too many nodes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const GraphView g = graph_view(graph_ws, n_nodes, e_raw);
  PiStats st;
  for (int k = 0; k < 4; ++k) st.xm[k] = stats[k], st.xs[k] = stats[4 + k];
  for (int k = 0; k < 2; ++k) st.em[k] = stats[8 + k], st.es[k] = stats[10 + k];
  const int n = static_cast<int>(n_nodes);
  const int blocks = static_cast<int>(ceil_div64(n_nodes, kPiBlock));
  float2* dpq = static_cast<float2*>(scratch);
  float* partial = static_cast<float*>(scratch) + 2 * n_nodes;
  const float inv_nodes = 1.f / static_cast<float>(n_nodes);
  PFN_CUDA_OK(launch_kernel(k_pi_node, dim3(blocks), dim3(kPiBlock), 0, stream, x, ldx, static_cast<const int*>(g.rowptr_s),
                            static_cast<const int*>(g.nbr_s), reinterpret_cast<const float2*>(g.ea_s), st, n, dpq, partial));
  PFN_LAUNCHED();
  PFN_CUDA_OK(launch_kernel(k_pi_final, dim3(1), dim3(kPiBlock), 0, stream, static_cast<const float*>(partial), blocks, inv_nodes, loss));
  PFN_LAUNCHED();
  if (dx != nullptr) {
    PFN_CUDA_OK(launch_kernel(k_pi_grad, dim3(blocks), dim3(kPiBlock), 0, stream, x, ldx, static_cast<const int*>(g.rowptr_s),
                              static_cast<const int*>(g.nbr_s), reinterpret_cast<const float2*>(g.ea_s),
                              static_cast<const int*>(g.rowptr_t), static_cast<const int*>(g.nbr_t),
                              reinterpret_cast<const float2*>(g.ea_t), st, n, inv_nodes, static_cast<const float2*>(dpq), dx, lddx));
    PFN_LAUNCHED();
  }
  return 0;
}

extern "C" int pfn_adamw_step(int64_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2,
                              double eps, double weight_decay, int64_t step, void* stream_) {
  PFN_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), PFN_E_INVALID,
              "pfn_adamw_step: bad arguments");
  PFN_REQUIRE(step >= 1, PFN_E_INVALID, "pfn_adamw_step: step counts from 1 (the value AFTER torch's `step += 1`)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // bias corrections in double, as the Python scalars of torch/optim/adamw.py
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step)), bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  AdamArgs a{};
  a.decay = static_cast<float>(1.0 - lr * weight_decay);
  a.beta1_w = static_cast<float>(1.0 - beta1);
  a.beta2 = static_cast<float>(beta2);
  a.beta2_w = static_cast<float>(1.0 - beta2);
  a.bc2_sqrt = static_cast<float>(sqrt(bc2));
  a.eps = static_cast<float>(eps);
  a.step_size = static_cast<float>(lr / bc1);
  for (int64_t t0 = 0; t0 < n_tensors; t0 += kAdamMaxTensors) {
    const int cnt = static_cast<int>(n_tensors - t0 < kAdamMaxTensors ? n_tensors - t0 : kAdamMaxTensors);
    int64_t max_n = 0;
    for (int t = 0; t < cnt; ++t) {
      PFN_REQUIRE(numel[t0 + t] >= 0 && (numel[t0 + t] == 0 || (params[t0 + t] && grads[t0 + t] && exp_avg[t0 + t] && exp_avg_sq[t0 + t])),
                  PFN_E_INVALID, "pfn_adamw_step: null tensor %lld", static_cast<long long>(t0 + t));
      a.p[t] = params[t0 + t];
      a.g[t] = grads[t0 + t];
      a.m[t] = exp_avg[t0 + t];
      a.v[t] = exp_avg_sq[t0 + t];
      a.n[t] = numel[t0 + t];
      if (numel[t0 + t] > max_n) max_n = numel[t0 + t];
    }
    if (max_n == 0) continue;
    int64_t bx = ceil_div64(max_n, int64_t(kAdamBlock) * 4);
    const int64_t cap = int64_t(sm_count()) * 8;
    if (bx > cap) bx = cap;
    PFN_CUDA_OK(launch_kernel(k_adamw, dim3(static_cast<unsigned>(bx), static_cast<unsigned>(cnt)), dim3(kAdamBlock), 0, stream, a));
    PFN_LAUNCHED();
  }
  return 0;
}
