// Tensor-core path for the dense per-node Linear stacks: tcgen05.mma (kind::tf32) with TMEM accumulators,
// operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), fp32-grade accuracy through a
// 3-term split-precision scheme ("3xTF32"):
//
//     a = a_hi + a_lo (+ r),  a_hi = rn_tf32(a) (low 13 mantissa bits clear: exactly representable in TF32),
//                             a_lo = rn_tf32(a - a_hi);  |r| <= 2^-24 |a|
//     A B^T  ~=  A_lo B_hi^T + A_hi B_lo^T + A_hi B_hi^T            (accumulated in fp32 in TMEM)
//
// which keeps the result within ~1e-6 relative of an fp32 GEMM -- inside the path's 1e-5 contract -- at a third
// of the TF32 tensor rate instead of the FFMA rate.  The split is done ON CHIP: converter warps rewrite each
// TMA-landed tile in shared memory as (hi, lo) in place; because the transform is elementwise, the 128B swizzle
// pattern written by TMA and read by the UMMA descriptors is untouched.
//
// Shape: Y[M, N] (+)= sum over segments  A_seg[M, K_seg] * B_seg[N, K_seg]^T, both operands K-major (row-major
// activations x row-major [out, in] weights): forward Linear directly, data gradients through transposed weight
// copies (k_pack_weights).  CTA tile 128 x BN (BN = N rounded up to 16, <= 256), K tile 32 floats (= one 128B
// swizzle row), one CTA per output tile, 6 warps:
//     warp 0      TMA producer (one elected lane)          warp 1   tcgen05.mma issuer (one lane) + TMEM alloc
//     warps 2..9  hi/lo converters, then epilogue (tcgen05.ld -> shared-memory tile -> float4 rows with bias /
//                 row-scaled bias / residual / dropout+ReLU / backward mask -> global)
// Pipelines: full[s] (TMA -> converters), conv[s] (converters -> MMA), empty[s] (MMA commit -> TMA), accum (MMA -> epilogue).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace pfn {
namespace {

constexpr int kTcBM = 128;
constexpr int kTcBK = 32;            // floats per K tile = 128 bytes = one swizzle row
constexpr int kTcWorkers = 256;          // 8 converter / epilogue warps
constexpr int kTcThreads = 64 + kTcWorkers;  // + TMA producer warp + MMA issuer warp
constexpr int kTcMaxStages = 4;
constexpr uint32_t kSmemLimit = 227 * 1024;

struct TcItem {
  CUtensorMap a;     // [rows = M, cols = K]   box {32, 128}
  CUtensorMap b;     // [rows = N, cols = K]   box {32, BN}   (the hi plane when the weights are pre-split)
  CUtensorMap b_lo;  // lo plane (pre-split weights only)
  int K;
  int presplit;
  int pad_[14];
};
static_assert(sizeof(TcItem) % 64 == 0, "tensor maps must stay 64-byte aligned inside the parameter block");

struct TcArgs {
  TcItem it[kGemmMaxItems];
  float* C[kGemmMaxItems];
  const float* bias[kGemmMaxItems];
  int n_items, batched, M, N, BN, ldc, stages, tmem_cols, n_hi, drain_tiles, debias_ulps, a_tmem;  // a_tmem: first TMEM column of the two A-operand stages (ATMEM kernels)
  const float* rowscale;
  const float* addend;
  const float* ymask;
  const float* inj;
  int ld_add, ld_ym, ld_inj, act;
  float scale;
  uint32_t seed_lo, seed_hi, keep_thresh;
  const uint32_t* seed_dev;
  long long* timing;  // debug: CTA 0 writes clock64() at phase boundaries when non-null
};

using namespace tc;


// rewrite a TMA-landed tile as (hi in place, lo at +lo_off); n_vec float4 elements, kTcWorkers converter threads.
// Four elements per thread are loaded before the first store: the volatile shared-memory accesses are issued in program
// order, so a load-convert-store loop pays one shared-memory round trip per element (measured 154 cycles per float4,
// 1.2 k cycles per 32 KB of tiles -- the serial stage of the weight-gradient pipeline).
__device__ __forceinline__ void split_tile(uint32_t hi_addr, uint32_t lo_off, int n_vec, int tid_c) {
  constexpr int B = 4;
  for (int i0 = tid_c; i0 < n_vec; i0 += B * kTcWorkers) {
    float4 v[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int i = i0 + j * kTcWorkers;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n_vec)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w) : "r"(hi_addr + 16u * i));
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int i = i0 + j * kTcWorkers;
      if (i >= n_vec) continue;
      const uint32_t a = hi_addr + 16u * i;
      const float4 h = make_float4(tf32_rn(v[j].x), tf32_rn(v[j].y), tf32_rn(v[j].z), tf32_rn(v[j].w));
      const float4 l = make_float4(tf32_rn(v[j].x - h.x), tf32_rn(v[j].y - h.y), tf32_rn(v[j].z - h.z), tf32_rn(v[j].w - h.w));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + lo_off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
    }
  }
}

struct EpiCtx {
  uint32_t tile, tile_ld;
  int row0, m0, n0, ncols, ncols_vec, lane;
  float* C;
  const float* bias;
};

__device__ __forceinline__ float epi_apply(float x, int act, float ym, bool keep, float scale) {
  if (act == PFN_ACT_RELU) return fmaxf(x, 0.f);
  if (act == PFN_ACT_DROPOUT_RELU) return keep ? fmaxf(x * scale, 0.f) : 0.f;
  if (act == kActMaskByY) return ym > 0.f ? x * scale : 0.f;
  return x;
}

// 16 rows of the staged tile -> global.  Columns [0, ncols_vec) go out as float4 (lane = 4 columns, 8 rows per batch with
// the batch's loads issued before the first use); the remaining columns (alignment tail, or everything when a pointer is
// not 16-byte aligned / a test keep-mask is injected) take the scalar path with lanes across columns.
// Deliberately NOT fully unrolled: a CTA executes its epilogue once, so straight-line code is instruction-fetch bound
// (measured: 27k cycles unrolled vs the few thousand the data movement needs); small loop bodies stay in the I-cache.
template <int ACT, bool ADD>
__device__ __forceinline__ void epilogue_rows(const TcArgs& args, const EpiCtx& e) {
  const int M = args.M;
  float* __restrict__ const C = e.C;
  const float* __restrict__ const bias = e.bias;
  const float* __restrict__ const addend = args.addend;
  const float* __restrict__ const ymask = args.ymask;
  const float* __restrict__ const rowscale = args.rowscale;
  const float scale = args.scale;
  uint32_t seed_lo = args.seed_lo, seed_hi = args.seed_hi;
  if (ACT == PFN_ACT_DROPOUT_RELU && args.seed_dev != nullptr) {
    seed_lo ^= args.seed_dev[0];
    seed_hi ^= args.seed_dev[1];
  }
  constexpr int RB = 8;
#define PFN_ESTAMP(slot)                                                                                        \
  do {                                                                                                          \
    if (args.timing != nullptr && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)   \
      args.timing[slot] = clock64();                                                                            \
  } while (0)
  PFN_ESTAMP(32);
#pragma unroll 1
  for (int cb = 0; cb < e.ncols_vec; cb += 128) {
    const int c = cb + 4 * e.lane;
    const bool cok = c < e.ncols_vec;
    const int n = e.n0 + c;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr && cok) b4 = make_float4(bias[n], bias[n + 1], bias[n + 2], bias[n + 3]);
    PFN_ESTAMP(33);
#pragma unroll 1
    for (int rb = 0; rb < 16; rb += RB) {
      float4 v[RB], ad[RB], ym[RB];
      float rs[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int row = e.row0 + rb + i, m = e.m0 + row;
        const bool ok = cok && m < M;
        v[i] = ad[i] = ym[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cok)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                       : "r"(e.tile + (uint32_t(row) * e.tile_ld + uint32_t(c)) * 4u));
        rs[i] = (rowscale != nullptr && m < M) ? rowscale[m] : 1.f;
        if (ADD && ok) ad[i] = *reinterpret_cast<const float4*>(addend + size_t(m) * args.ld_add + n);
        if (ACT == kActMaskByY && ok) ym[i] = *reinterpret_cast<const float4*>(ymask + size_t(m) * args.ld_ym + n);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int m = e.m0 + e.row0 + rb + i;
        if (!(cok && m < M)) continue;
        float x[4] = {fmaf(rs[i], b4.x, v[i].x) + ad[i].x, fmaf(rs[i], b4.y, v[i].y) + ad[i].y,
                      fmaf(rs[i], b4.z, v[i].z) + ad[i].z, fmaf(rs[i], b4.w, v[i].w) + ad[i].w};
        const float y4[4] = {ym[i].x, ym[i].y, ym[i].z, ym[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool keep = true;
          if (ACT == PFN_ACT_DROPOUT_RELU) keep = dropout_hash(m, n + j, seed_lo, seed_hi) >= args.keep_thresh;
          x[j] = epi_apply(x[j], ACT, y4[j], keep, scale);
        }
        *reinterpret_cast<float4*>(C + size_t(m) * args.ldc + n) = make_float4(x[0], x[1], x[2], x[3]);
      }
      PFN_ESTAMP(34 + rb / RB);
    }
  }
  PFN_ESTAMP(38);
  // scalar columns [ncols_vec, ncols): one (row, column) pair per lane and iteration, rows fastest, so the few tail
  // columns of all 16 rows are in flight together (a lane walking 16 rows in sequence cost more than the float4 part)
  const float* __restrict__ const inj = args.inj;
  const int n_tail = e.ncols - e.ncols_vec;
#pragma unroll 1
  for (int idx = e.lane; idx < 16 * n_tail; idx += 32) {
    const int r = idx & 15, c = e.ncols_vec + (idx >> 4);
    const int row = e.row0 + r, m = e.m0 + row, n = e.n0 + c;
    if (m >= M) continue;
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(e.tile + (uint32_t(row) * e.tile_ld + uint32_t(c)) * 4u));
    float x = v;
    if (bias != nullptr) x = fmaf(rowscale != nullptr ? rowscale[m] : 1.f, bias[n], v);
    if (ADD) x += addend[size_t(m) * args.ld_add + n];
    bool keep = true;
    float ym = 0.f;
    if (ACT == PFN_ACT_DROPOUT_RELU)
      keep = inj != nullptr ? inj[size_t(m) * args.ld_inj + n] != 0.f : dropout_hash(m, n, seed_lo, seed_hi) >= args.keep_thresh;
    if (ACT == kActMaskByY) ym = ymask[size_t(m) * args.ld_ym + n];
    C[size_t(m) * args.ldc + n] = epi_apply(x, ACT, ym, keep, scale);
  }
  PFN_ESTAMP(39);
}

#define PFN_TSTAMP(slot)                                                                   \
  do {                                                                                     \
    if (args.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)   \
      args.timing[slot] = clock64();                                                       \
  } while (0)

// DRAIN = true (reductions longer than kTcDrainMinTiles K tiles: hidden 512, the K-segmented TAGConv GEMMs): the hi*hi
// accumulator is FLUSHED every drain_tiles K tiles.  The tensor core adds into its fp32 TMEM accumulator with truncation,
// so a sum that stays in TMEM for c accumulating instructions is off by ~c * 2^-25 relative -- and always towards zero,
// a BIAS that compounds through a 10-19 layer forward + backward instead of averaging out (measured: gradients of the
// hidden-512 configurations 1.5e-5 from the fp64 twin).  Here phase p (kTcDrainTiles tiles) accumulates into TMEM slot
// p mod n_hi starting from zero; when its MMAs have completed the worker warps read the slot (tcgen05.ld) and add it to
// per-thread fp32 REGISTER accumulators in round-to-nearest, while the tensor core is already filling the next slot.
// No TMEM sum is ever longer than 4 * drain_tiles instructions, whatever K is (drain_tiles = 1 or 2 K tiles per phase).
constexpr int kTcDrainTilesDefault = 1;  // measured (case6470rte x 2, hidden 512, 5 layers): gradients 1.0-1.3x the fp32 reference's own distance to fp64 (2 tiles: 2-3x), forward GEMMs +8 %
constexpr int kTcDrainMinTiles = 2;
constexpr int kTcDrainChunks = 5;  // 16-column chunks per worker thread: ceil(160 / 32)

// ATMEM = true: the A operand of every MMA lives in TENSOR MEMORY.  Measured (ncu + cycle stamps at case6470rte x 32,
// hidden 512): the mainloop is bound by the shared-memory pipe -- per K tile the tensor core reads 96 KB of operands, the
// converters read 16 KB and write 32 KB, TMA writes 48 KB: ~1.5 k wavefronts against 768 cycles of MMA work -- and the MMAs
// execute at ~170 cycles each instead of the 64 they take alone (scripts/probes/tmem_a_probe.cu).  With ATMEM a worker thread
// reads its row's share of the TMA-landed fp32 tile, splits it and writes (hi, lo) with tcgen05.st into one of two TMEM
// operand stages (lane = row, column = k); the MMAs take A from there (no A descriptor reads, no hi/lo planes in shared
// memory, a stage shrinks from 64 to 48 KB -> four stages in flight).  B is unchanged (weights pre-split by k_pack_weights).
template <bool DRAIN, bool ATMEM>
__global__ void __launch_bounds__(kTcThreads, 1) k_gemm_tc(const __grid_constant__ TcArgs args) {
  extern __shared__ uint8_t smem_dyn[];
  if (threadIdx.x == 0) PFN_TSTAMP(0);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const int BN = args.BN, S = args.stages;
  const uint32_t a_bytes = kTcBM * 128u, b_bytes = uint32_t(BN) * 128u;
  const uint32_t b_off = ATMEM ? a_bytes : 2u * a_bytes;          // ATMEM: A (fp32 as landed) | B_hi | B_lo
  const uint32_t stage_bytes = b_off + 2u * b_bytes;              // else:  A_hi | A_lo | B_hi | B_lo
  const uint32_t bar_base = base + uint32_t(S) * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (kTcMaxStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kTcMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * kTcMaxStages);
  const uint32_t tmem_slot = accum_bar + 8u;
  auto ready_bar = [&](int a) { return accum_bar + 16u + 8u * a; };         // DRAIN: MMAs of a phase complete -> workers
  auto drained_bar = [&](int a) { return accum_bar + 16u + 8u * (3 + a); };  // DRAIN: slot read out -> MMA issuer

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kTcBM, n0 = blockIdx.y * BN;
  const int prob = blockIdx.z;
  const int seg_begin = args.batched ? prob : 0, seg_end = args.batched ? prob + 1 : args.n_items;
  int n_tiles_total = 0;
  for (int seg = seg_begin; seg < seg_end; ++seg) n_tiles_total += (args.it[seg].K + kTcBK - 1) / kTcBK;

  // TMA producer state (warp 0, lane 0): the first S tiles are requested BEFORE the CTA-wide setup barrier so the
  // first load's latency (~2 us: descriptor fetch + 272 row segments from L2/HBM) overlaps TMEM allocation
  int p_it = 0, p_seg = seg_begin, p_k0 = 0;
  auto produce = [&](int limit) {
    while (p_seg < seg_end && p_it < limit) {
      const TcItem& item = args.it[p_seg];
      if (p_k0 >= item.K) {
        ++p_seg;
        p_k0 = 0;
        continue;
      }
      const int s = p_it % S;
      const uint32_t ph = (p_it / S) & 1;
      mbar_wait(empty_bar(s), ph ^ 1u);
      mbar_arrive_expect_tx(full_bar(s), a_bytes + (item.presplit ? 2u : 1u) * b_bytes);
      const uint32_t st = base + uint32_t(s) * stage_bytes;
      tma_load_2d(st, &item.a, p_k0, m0, full_bar(s));
      tma_load_2d(st + b_off, &item.b, p_k0, n0, full_bar(s));
      if (item.presplit) tma_load_2d(st + b_off + b_bytes, &item.b_lo, p_k0, n0, full_bar(s));
      p_k0 += kTcBK;
      ++p_it;
    }
  };
  if (warp == 0 && elect_one()) {  // (warp-uniform first operand: all of warp 0 reaches the election)
    for (int seg = seg_begin; seg < seg_end; ++seg) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.it[seg].a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.it[seg].b)) : "memory");
      if (args.it[seg].presplit) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.it[seg].b_lo)) : "memory");
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), kTcWorkers / 32);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    if (DRAIN) {
      for (int a = 0; a < 3; ++a) {
        mbar_init(ready_bar(a), 1);
        mbar_init(drained_bar(a), kTcWorkers / 32);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_wait();  // everything above overlaps the previous kernel's tail; the operands may only be read from here on
    produce(S);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(uint32_t(args.tmem_cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) PFN_TSTAMP(1);

  if (warp == 0) {
    // ===== TMA producer: remaining tiles =====
    if (elect_one()) produce(1 << 30);
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(kTcBM >> 4) << 24);
    // The tensor core adds into its fp32 accumulator with truncation, so the error grows with the number of
    // accumulating instructions.  The hi*hi products therefore rotate over n_hi accumulators and the (2^-11 smaller)
    // lo terms go to their own accumulator; the epilogue adds them in round-to-nearest fp32.
    const int n_hi = args.n_hi;
    const uint32_t d_lo = tmem_base + uint32_t(n_hi * BN);
    int it = 0, kk = 0;
    for (int seg = seg_begin; seg < seg_end; ++seg) {
      const int K = args.it[seg].K;
      for (int k0 = 0; k0 < K; k0 += kTcBK, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        if (lane == 0 && it == 5) PFN_TSTAMP(61);
        mbar_wait(conv_bar(s), ph);
        if (lane == 0 && it == 5) PFN_TSTAMP(62);
        const int dt = args.drain_tiles;
        const int phase = it / dt, slot = phase % n_hi;  // DRAIN only
        const bool phase_first = it % dt == 0;
        if (DRAIN && phase_first && phase >= n_hi) mbar_wait(drained_bar(slot), uint32_t(phase / n_hi - 1) & 1u);
        tc_fence_after();
        const int nk = min(kTcBK / 8, (K - k0 + 7) / 8);  // 8 TF32 elements (32 bytes) per UMMA K step
        if (lane == 0 && it < 8) PFN_TSTAMP(18 + it);
        if (elect_one()) {
          const uint32_t st = base + uint32_t(s) * stage_bytes;
          const uint64_t a_hi = umma_desc_k128(st), a_lo = umma_desc_k128(st + a_bytes);
          const uint64_t b_hi = umma_desc_k128(st + b_off), b_lo = umma_desc_k128(st + b_off + b_bytes);
          const uint32_t at_hi = tmem_base + uint32_t(args.a_tmem + 64 * (it & 1)), at_lo = at_hi + 32u;  // ATMEM: this tile's operand stage
          if (DRAIN && nk == kTcBK / 8) {
            // the common case straight-line: four K steps with constant offsets and constant accumulate flags (only the
            // first step of a phase / of the whole reduction starts from zero).  The issuing thread's instruction stream
            // is the pipeline's pace-maker: ~100 cycles of descriptor arithmetic per MMA in the general loop below
            const uint32_t d_hi = tmem_base + uint32_t(slot * BN);
            const uint32_t acc_hi0 = phase_first ? 0u : 1u, acc_lo0 = kk > 0 ? 1u : 0u;
#pragma unroll
            for (int j = 0; j < kTcBK / 8; ++j) {
              const uint64_t adv = uint64_t(j * 2);
              if (ATMEM) {
                umma_tf32_ts(d_lo, at_lo + uint32_t(8 * j), b_hi + adv, idesc, j == 0 ? acc_lo0 : 1u);
                umma_tf32_ts(d_lo, at_hi + uint32_t(8 * j), b_lo + adv, idesc, 1u);
                umma_tf32_ts(d_hi, at_hi + uint32_t(8 * j), b_hi + adv, idesc, j == 0 ? acc_hi0 : 1u);
              } else {
                umma_tf32(d_lo, a_lo + adv, b_hi + adv, idesc, j == 0 ? acc_lo0 : 1u);
                umma_tf32(d_lo, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_tf32(d_hi, a_hi + adv, b_hi + adv, idesc, j == 0 ? acc_hi0 : 1u);
              }
            }
          } else
          for (int j = 0; j < nk; ++j) {
            const uint64_t adv = uint64_t(j * 2);  // +32 bytes in the 16-byte-granular start-address field
            const int k_idx = kk + j;
            uint32_t d_hi, acc_hi;
            if (DRAIN) {
              d_hi = tmem_base + uint32_t(slot * BN);
              acc_hi = (phase_first && j == 0) ? 0u : 1u;
            } else {
              d_hi = tmem_base + uint32_t((k_idx % n_hi) * BN);
              acc_hi = k_idx >= n_hi ? 1u : 0u;
            }
            if (ATMEM) {
              umma_tf32_ts(d_lo, at_lo + uint32_t(8 * j), b_hi + adv, idesc, k_idx > 0 ? 1u : 0u);
              umma_tf32_ts(d_lo, at_hi + uint32_t(8 * j), b_lo + adv, idesc, 1u);
              umma_tf32_ts(d_hi, at_hi + uint32_t(8 * j), b_hi + adv, idesc, acc_hi);
            } else {
              umma_tf32(d_lo, a_lo + adv, b_hi + adv, idesc, k_idx > 0 ? 1u : 0u);
              umma_tf32(d_lo, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_tf32(d_hi, a_hi + adv, b_hi + adv, idesc, acc_hi);
            }
          }
          umma_commit(empty_bar(s));  // implies tcgen05.fence::before_thread_sync
          if (DRAIN && (it % dt == dt - 1 || it == n_tiles_total - 1)) umma_commit(ready_bar(slot));
        }
        kk += nk;
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(accum_bar);
    if (lane == 0) PFN_TSTAMP(26);
    __syncwarp();
  } else {
    // ===== converters (hi/lo split in shared memory), then epilogue =====
    const int tid_c = threadIdx.x - 64;
    const int wk = warp - 2;                 // 0..7
    const int q = warp & 3, half = wk >> 2;  // TMEM lane quarter this warp may read / which 16-column chunks it drains
    const uint32_t lane_base = tmem_base + (uint32_t(32 * q) << 16);
    // DRAIN: this thread's share of the output tile (row 32 q + lane, columns 16 half + 32 i + [0, 16)) lives in registers
    // (kept as packed pairs: the flush below adds two values per instruction)
    uint64_t racc2[DRAIN ? kTcDrainChunks : 1][8];
#pragma unroll
    for (int i = 0; i < (DRAIN ? kTcDrainChunks : 1); ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) racc2[i][j] = 0ull;
    // racc += TMEM columns [col0, col0 + BN) of this thread's row.  `ulps` > 0 (the hi*hi slots): every value read is first
    // moved `ulps` units in the last place AWAY from zero.  A phase sum has been truncated towards zero once per
    // accumulating instruction (4 per K tile), i.e. it comes out ~1.5 ulp too small in magnitude on average (measured:
    // mean signed relative error -1.26e-7 on same-sign data); left alone, that uniform shrink of every GEMM output
    // compounds linearly through a 10-20 layer forward + backward (measured 1-2e-5 on configs/wide.json gradients where
    // fp32 FFMA GEMMs give 2.6e-6).  Adding back the expected loss turns the bias into zero-mean rounding noise.
    auto drain_slot = [&](uint32_t col0, uint32_t ulps) {
#pragma unroll
      for (int i = 0; i < (DRAIN ? kTcDrainChunks : 1); ++i) {
        const int c0 = 16 * half + 32 * i;
        if (c0 < BN) {  // warp-uniform
          uint32_t r[16];
          tmem_ld16(lane_base + col0 + uint32_t(c0), r);
          // two instructions per PAIR of values (this loop is the mainloop's instruction budget: 64 values per thread and
          // K tile): the integer add moves the magnitude `ulps` up -- a +-0 becomes a denormal, which the flush-to-zero
          // add then drops again, so exact zeros stay exact -- and one packed add accumulates both
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            uint64_t v;
            asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(r[j] + ulps), "r"(r[j + 1] + ulps));
            asm("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(racc2[i][j >> 1]) : "l"(v));
          }
        }
      }
    };
    int next_drain = 0;  // first phase whose slot has not been read out yet
    int it = 0;
    for (int seg = seg_begin; seg < seg_end; ++seg) {
      const int K = args.it[seg].K;
      for (int k0 = 0; k0 < K; k0 += kTcBK, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        if (tid_c == 0 && it == 5) PFN_TSTAMP(56);
        mbar_wait(full_bar(s), ph);
        if (tid_c == 0 && it == 5) PFN_TSTAMP(57);
        if (tid_c == 0 && it < 8) PFN_TSTAMP(2 + it);
        const uint32_t st = base + uint32_t(s) * stage_bytes;
        if (ATMEM) {
          // this thread: row 32 q + lane of the tile, K elements [16 half, 16 half + 16) -- four 16-byte chunks of the
          // 128-byte swizzled row (chunk c of row r sits at position c ^ (r mod 8))
          const int row = 32 * q + lane;
          const uint32_t row_addr = st + uint32_t(row) * 128u;
          float v[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t pos = uint32_t((4 * half + c) ^ (row & 7));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3]) : "r"(row_addr + 16u * pos));
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float h = tf32_rn(v[i]);
            hi[i] = __float_as_uint(h);
            lo[i] = __float_as_uint(tf32_rn(v[i] - h));
          }
          // the operand stage was last read by the MMAs of tile it - 2: their completion is what frees that tile's
          // shared-memory stage (empty_bar)
          if (it >= 2) {
            mbar_wait(empty_bar((it - 2) % S), uint32_t((it - 2) / S) & 1u);
            tc_fence_after();
          }
          const uint32_t at = lane_base + uint32_t(args.a_tmem + 64 * (it & 1) + 16 * half);
          tmem_st16(at, hi);
          tmem_st16(at + 32u, lo);
          tmem_st_wait();
          if (!args.it[seg].presplit) {
            split_tile(st + b_off, b_bytes, BN * 8, tid_c);
            proxy_fence_async();
          }
          tc_fence_before();  // the MMA issuer orders the tcgen05.st above through conv_bar + fence::after_thread_sync
        } else {
          split_tile(st, a_bytes, kTcBM * 8, tid_c);
          if (!args.it[seg].presplit) split_tile(st + 2u * a_bytes, b_bytes, BN * 8, tid_c);  // weights arrive pre-split
          proxy_fence_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        }
        if (tid_c == 0 && it < 8) PFN_TSTAMP(10 + it);
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar(s));
        if (tid_c == 0 && it == 5) PFN_TSTAMP(58);
        if (DRAIN && (it + 1) % args.drain_tiles == 0 && it + 1 >= 2 * args.drain_tiles) {
          // every tile of phase p has been handed to the tensor core: read out phase p - 1 (issued a phase ago, so its
          // MMAs have normally completed) while phase p is being multiplied
          const int pd = (it + 1) / args.drain_tiles - 2, slot = pd % args.n_hi;
          mbar_wait(ready_bar(slot), uint32_t(pd / args.n_hi) & 1u);
          if (tid_c == 0 && it == 5) PFN_TSTAMP(59);
          tc_fence_after();
          drain_slot(uint32_t(slot * BN), uint32_t(args.debias_ulps));
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(drained_bar(slot));
          if (tid_c == 0 && it == 5) PFN_TSTAMP(60);
          next_drain = pd + 1;
        }
      }
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // The staging tile written below overlays pipeline stage 0, last written by split_tile.  The mbarrier chain (conv_bar
    // arrivals of ALL converter warps -> MMA -> accum_bar) already orders the two; this named barrier restates the
    // ordering among the worker warps in a form compute-sanitizer's racecheck models (once per CTA: free).
    asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
    if (threadIdx.x == 64) PFN_TSTAMP(27);
    // Epilogue.  The pipeline stages are free now (every MMA has completed), so the accumulator tile is staged through
    // shared memory: phase 1, each warp drains its TMEM lane quarter (two warps per quarter, alternating 16-column
    // chunks) and adds the accumulators in fp32; phase 2, the eight warps take 16 rows each and stream them out as
    // float4 with all global loads of a row batch in flight together.
    const uint32_t tile_ld = uint32_t(BN) + 4u;  // floats; +4 keeps the 16-byte row-chunk stores conflict-free
    const int ncols = min(BN, args.N - n0);
    if (DRAIN) {
      // the phases not read out yet (every MMA has completed), then the lo terms; registers -> staging tile
      const int n_phases = (n_tiles_total + args.drain_tiles - 1) / args.drain_tiles;
      for (int pd = next_drain; pd < n_phases; ++pd) drain_slot(uint32_t((pd % args.n_hi) * BN), uint32_t(args.debias_ulps));
      drain_slot(uint32_t(args.n_hi * BN), 0u);  // the (2^-11 smaller) lo terms: no correction
      const uint32_t row_addr = base + (uint32_t(32 * q + lane) * tile_ld) * 4u;
#pragma unroll
      for (int i = 0; i < kTcDrainChunks; ++i) {
        const int c0 = 16 * half + 32 * i;
        if (c0 < ncols) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(row_addr + 4u * (c0 + j)), "l"(racc2[DRAIN ? i : 0][j >> 1]),
                         "l"(racc2[DRAIN ? i : 0][(j >> 1) + 1]) : "memory");
        }
      }
    } else {
      int total_k = 0;
      for (int seg = seg_begin; seg < seg_end; ++seg) total_k += (args.it[seg].K + 7) / 8;
      const int n_act = min(args.n_hi, total_k);  // hi accumulators that were actually written
      const uint32_t row_addr = base + (uint32_t(32 * q + lane) * tile_ld) * 4u;
      for (int c0 = 16 * half; c0 < ncols; c0 += 32) {
        uint32_t r[4][16];  // lo terms + up to three hi accumulators, all loads in flight before one wait
        tmem_ld16_issue(lane_base + uint32_t(args.n_hi * BN + c0), r[0]);
        if (n_act > 0) tmem_ld16_issue(lane_base + uint32_t(0 * BN + c0), r[1]);
        if (n_act > 1) tmem_ld16_issue(lane_base + uint32_t(1 * BN + c0), r[2]);
        if (n_act > 2) tmem_ld16_issue(lane_base + uint32_t(2 * BN + c0), r[3]);
        tmem_ld_wait();
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = __uint_as_float(r[0][i]);  // the small (lo) terms first
          if (n_act > 0) a += __uint_as_float(r[1][i]);
          if (n_act > 1) a += __uint_as_float(r[2][i]);
          if (n_act > 2) a += __uint_as_float(r[3][i]);
          acc[i] = a;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + 4u * (c0 + i)), "f"(acc[i]), "f"(acc[i + 1]),
                       "f"(acc[i + 2]), "f"(acc[i + 3]) : "memory");
      }
    }
    if (threadIdx.x == 64) PFN_TSTAMP(28);
    asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
    if (threadIdx.x == 64) PFN_TSTAMP(29);
    EpiCtx e;
    e.tile = base;
    e.tile_ld = tile_ld;
    e.row0 = 16 * wk;
    e.m0 = m0;
    e.n0 = n0;
    e.ncols = ncols;
    e.lane = lane;
    e.C = args.C[args.batched ? prob : 0];
    e.bias = args.bias[args.batched ? prob : 0];
    const bool vec = args.inj == nullptr && (args.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(e.C) & 15u) == 0 &&
                     (args.addend == nullptr || ((args.ld_add & 3) == 0 && (reinterpret_cast<uintptr_t>(args.addend) & 15u) == 0)) &&
                     (args.ymask == nullptr || ((args.ld_ym & 3) == 0 && (reinterpret_cast<uintptr_t>(args.ymask) & 15u) == 0));
    e.ncols_vec = vec ? (ncols & ~3) : 0;
    const bool add = args.addend != nullptr;
    switch (args.act) {
      case PFN_ACT_NONE: add ? epilogue_rows<PFN_ACT_NONE, true>(args, e) : epilogue_rows<PFN_ACT_NONE, false>(args, e); break;
      case PFN_ACT_RELU: add ? epilogue_rows<PFN_ACT_RELU, true>(args, e) : epilogue_rows<PFN_ACT_RELU, false>(args, e); break;
      case PFN_ACT_DROPOUT_RELU: epilogue_rows<PFN_ACT_DROPOUT_RELU, false>(args, e); break;
      default: epilogue_rows<kActMaskByY, false>(args, e); break;
    }
    if (threadIdx.x == 64) PFN_TSTAMP(30);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(args.tmem_cols)) : "memory");
  }
}

// ---- weight gradient on the tensor cores ---------------------------------------------------------------
// dW[Mo, Ni] = dY^T X = sum over nodes m of dY[m, :Mo]^T X[m, :Ni]: the reduction runs over the node dimension, so both
// operands are "MN-major" for the tensor core (the M/N index is the contiguous one).  For TF32 the only MN-major
// shared-memory layout the tensor core accepts is SWIZZLE_128B_BASE32B (128-byte rows whose 32-byte chunks are XORed with
// row mod 4), which is what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: boxes of 32 nodes x 32 floats land as
// atoms of 4 K-rows x 128 B, K-groups 512 B apart (SBO), 32-float M/N blocks one box (4096 B) apart (LBO).  Split-K over node chunks fills the
// GPU; every CTA writes a partial tile that k_splitk_reduce sums in fixed order.  A CTA can own two 128-row M tiles
// sharing the X tile (hidden_dim 129 = 128 + 1), and the bias gradient rides along as one extra column of X that the
// converter warps overwrite with ones (or the in-degree, for the deg (.) b2 term) before the hi/lo split.
struct WgItem {
  CUtensorMap a;  // dY  [rows = nodes, cols = Mo]   box {32, 32}
  CUtensorMap b;  // X   [rows = nodes, cols = Ni]   box {32, 32}
};
struct WgArgs {
  WgItem it[kGemmMaxItems];
  int count, Mo, N, n_eff, BN, mt, nb, K, kchunk, splitk, stages, tmem_cols, extra_col;
  const float* extra_vec;
  float* partial;
  uint32_t lbo, sbo, kstep_bytes;  // descriptor strides (bytes)
  int mn_major, debug;
  int acc_hi[2], acc_lo[2];        // TMEM column of each M tile's hi*hi / lo-term accumulator (equal = shared)
};

__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // layout type 1 = SWIZZLE_128B_BASE32B: the only shared-memory layout the tensor core accepts for MN-major TF32
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) |
         (uint64_t(1) << 46) | (uint64_t(1) << 61);
}

__global__ void __launch_bounds__(kTcThreads, 1) k_wgrad_tc(const __grid_constant__ WgArgs args) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int BN = args.BN, S = args.stages, mt = args.mt, nb = args.nb;
  const uint32_t a_bytes = uint32_t(mt) * 4u * 4096u, b_bytes = uint32_t(nb) * 4096u;
  const uint32_t stage_bytes = 2u * (a_bytes + b_bytes);  // A_hi | A_lo | B_hi | B_lo
  const uint32_t bar_base = base + uint32_t(S) * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (kTcMaxStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kTcMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * kTcMaxStages);
  const uint32_t tmem_slot = accum_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, prob = blockIdx.y, mgrp = blockIdx.z;  // mgrp: group of `mt` M tiles
  const int row0 = mgrp * mt * kTcBM;                                   // first dW row of this CTA
  const int k_beg = split * args.kchunk, k_end = min(args.K, k_beg + args.kchunk);
  const int n_tiles = (k_end > k_beg) ? (k_end - k_beg + kTcBK - 1) / kTcBK : 0;
  const WgItem& item = args.it[prob];

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), kTcWorkers / 32);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(uint32_t(args.tmem_cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        const int k0 = k_beg + it * kTcBK;
        mbar_wait_sleep(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), a_bytes + b_bytes);
        const uint32_t st = base + uint32_t(s) * stage_bytes;
        for (int b = 0; b < mt * 4; ++b) tma_load_2d(st + uint32_t(b) * 4096u, &item.a, row0 + 32 * b, k0, full_bar(s));
        for (int b = 0; b < nb; ++b) tma_load_2d(st + 2u * a_bytes + uint32_t(b) * 4096u, &item.b, 32 * b, k0, full_bar(s));
      }
    }
  } else if (warp == 1) {
    // D = F32, A = B = TF32, A and B MN-major (bits 15, 16), N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(args.mn_major) << 15) | (uint32_t(args.mn_major) << 16) |
                           (uint32_t(BN >> 3) << 17) | (uint32_t(kTcBM >> 4) << 24);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait_sleep(conv_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = base + uint32_t(s) * stage_bytes;
        const uint32_t b_hi = st + 2u * a_bytes, b_lo = b_hi + b_bytes;
        for (int j = 0; j < kTcBK / 8; ++j) {  // 8 nodes per UMMA K step = one 1024-byte K group
          const uint32_t koff = uint32_t(j) * args.kstep_bytes;
          const uint64_t dbh = umma_desc_mn128(b_hi + koff, args.lbo, args.sbo), dbl = umma_desc_mn128(b_lo + koff, args.lbo, args.sbo);
          for (int t = 0; t < mt; ++t) {
            const uint32_t a_hi = st + uint32_t(t) * 16384u, a_lo = a_hi + a_bytes;
            const uint64_t dah = umma_desc_mn128(a_hi + koff, args.lbo, args.sbo), dal = umma_desc_mn128(a_lo + koff, args.lbo, args.sbo);
            const uint32_t d_hi = tmem_base + uint32_t(args.acc_hi[t]), d_lo = tmem_base + uint32_t(args.acc_lo[t]);
            const uint32_t first = (it > 0 || j > 0) ? 1u : 0u;
            umma_tf32(d_lo, dal, dbh, idesc, first);
            umma_tf32(d_lo, dah, dbl, idesc, 1u);
            umma_tf32(d_hi, dah, dbh, idesc, (d_hi == d_lo) ? 1u : first);
          }
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accum_bar);
    __syncwarp();
  } else {
    const int tid_c = threadIdx.x - 64;
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      const int k0 = k_beg + it * kTcBK;
      mbar_wait_sleep(full_bar(s), ph);
      const uint32_t st = base + uint32_t(s) * stage_bytes;
      if (args.extra_col != 0) {
        // virtual column N of X := 1 (or extra_vec[node]) -> its dot products with dY are the bias gradient
        if (tid_c < kTcBK) {
          const int r = tid_c, cN = args.N, bb = cN >> 5, cc = cN & 31;
          const int node = k0 + r;
          float v = 0.f;
          if (node < k_end) v = args.extra_col == 1 ? 1.f : args.extra_vec[node];
          // 128B rows, 32-byte chunks XOR-swizzled with (row mod 4)  (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
          const uint32_t addr = st + 2u * a_bytes + uint32_t(bb) * 4096u + uint32_t(r) * 128u +
                                (uint32_t((cc >> 3) ^ (r & 3)) << 5) + uint32_t(cc & 7) * 4u;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcWorkers) : "memory");
      }
      split_tile(st, a_bytes, int(a_bytes / 16u), tid_c);
      split_tile(st + 2u * a_bytes, b_bytes, int(b_bytes / 16u), tid_c);
      proxy_fence_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar(s));
    }
    // epilogue: accumulators -> padded smem tile -> coalesced rows of the split-K partial buffer
    const int wk = warp - 2, q = warp & 3, half = wk >> 2;
    const int n_eff = args.n_eff, Mo = args.Mo;
    float* __restrict__ const part = args.partial + size_t(split * args.count + prob) * size_t(Mo) * n_eff;
    const uint32_t tile_ld = uint32_t(BN) + 4u;
    if (n_tiles > 0) {
      mbar_wait_sleep(accum_bar, 0);
      tc_fence_after();
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");  // (see k_gemm_tc: orders split_tile's stores before the staging tile's)
    for (int t = 0; t < mt; ++t) {
      if (row0 + t * kTcBM >= Mo) break;  // CTA-uniform
      if (n_tiles > 0) {
        const uint32_t row_addr = base + (uint32_t(32 * q + lane) * tile_ld) * 4u;
        const uint32_t lane_base = tmem_base + (uint32_t(32 * q) << 16);
        for (int c0 = 16 * half; c0 < n_eff; c0 += 32) {
          uint32_t r[16];
          float acc[16];
          tmem_ld16(lane_base + uint32_t(args.acc_lo[t] + c0), r);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(r[i]);
          if (args.acc_hi[t] != args.acc_lo[t]) {
            tmem_ld16(lane_base + uint32_t(args.acc_hi[t] + c0), r);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + 4u * (c0 + i)), "f"(acc[i]), "f"(acc[i + 1]),
                         "f"(acc[i + 2]), "f"(acc[i + 3]) : "memory");
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
      for (int rr = 0; rr < 16; ++rr) {
        const int row = 16 * wk + rr, m = row0 + t * kTcBM + row;
        if (m >= Mo) break;
        const uint32_t row_addr = base + (uint32_t(row) * tile_ld) * 4u;
        for (int c = lane; c < n_eff; c += 32) {
          float v = 0.f;
          if (n_tiles > 0) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(row_addr + 4u * c));
          part[size_t(m) * n_eff + c] = v;
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(args.tmem_cols)) : "memory");
  }
}


// ---- grouped weight gradient: every dW of a backward pass in ONE launch -------------------------------------------
// A step has ~20 weight-gradient problems (dW2, dWi, dWj per EdgeAggregation; dW_0..dW_K per TAGConv; mask_embd), each
// a [<=129, <=130] output reduced over all nodes.  Launched one by one, each is a 118-way split-K with a 118 x 66 KB
// partial buffer and its own reduction pass (26 launches, ~160 MB of partial traffic per step).  Here the work items
// (problem, M group, node chunk) of ALL problems form one grid of about one CTA per SM, so a CTA reduces over ~1/7 of
// the nodes in TMEM and the partial traffic drops ~17x; one more launch sums the few partials in fixed order.
// Same mainloop as k_wgrad_tc (MN-major TF32 operands, SWIZZLE_128B_BASE32B, converter warps, 3xTF32).
constexpr int kWgMaxProb = 32;
// The tensor core adds into its fp32 accumulator with truncation, so the error of a TMEM-resident sum grows linearly
// with the number of accumulating instructions (measured: 1.8e-5 relative at 2176 nodes per CTA, above the 1e-5
// contract).  A CTA therefore reduces at most this many nodes; the partials are summed in round-to-nearest fp32.
constexpr int kWgMaxChunkDefault = 512;
static int wg_max_chunk() {
  static const int v = [] {
    const char* e = std::getenv("PFN_WG_CHUNK");  // experiments only: larger chunks exceed the 1e-5 contract
    const int c = e != nullptr ? std::atoi(e) : 0;
    return c >= 32 ? c / 32 * 32 : kWgMaxChunkDefault;
  }();
  return v;
}
struct WgGroupProb {
  CUtensorMap a;  // dY  [rows = nodes, cols = Mo]   box {32, 32}
  CUtensorMap b;  // X   [rows = nodes, cols = Ni]   box {32, 32}
  float* dW;
  float* dbias;
  const float* extra_vec;
  long long part_off;  // float offset of this problem's partials: [split][Mo][n_eff]
  int Mo, N, n_eff, BN, mt, nb, extra_col, lddw, m_groups, tmem_cols, stages, item0;
  int acc_hi[2], acc_lo[2];
  int odd, oddn;
  int n_hi;  // mt == 1: hi*hi products rotate over n_hi accumulators at columns 0, BN, ... (lo terms at n_hi * BN)
  int atmem, a_tmem;  // A operand in tensor memory: first column of its two 64-column stages
};
struct WgGroupArgs {
  WgGroupProb p[kWgMaxProb];
  int n_prob, K, kchunk, splitk;
  float* partial;
  long long* timing;  // debug (PFN_WG_TIMING): CTA 0 writes clock64() at phase boundaries
};

__global__ void __launch_bounds__(kTcThreads, 1) k_wgrad_group(const __grid_constant__ WgGroupArgs args) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
#define WSTAMP(slot) do { if (args.timing != nullptr && blockIdx.x == 0) args.timing[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) WSTAMP(0);
  int prob = 0;
  for (int i = 1; i < args.n_prob; ++i)
    if (int(blockIdx.x) >= args.p[i].item0) prob = i;
  const WgGroupProb& P = args.p[prob];
  const int item = int(blockIdx.x) - P.item0;
  const int mgrp = item / args.splitk, split = item - mgrp * args.splitk;
  const int BN = P.BN, S = P.stages, mt = P.mt, nb = P.nb;
  const uint32_t a_bytes = uint32_t(mt) * 4u * 4096u, b_bytes = uint32_t(nb) * 4096u;
  // P.odd: Mo = 128 + 1 (hidden_dim 129).  A second 128-row M tile for ONE row would double the MMA and converter work,
  // so row 128 of dW is accumulated by the converter threads instead (thread n: sum over nodes of dY[node][128] X[node][n],
  // fp32 FMAs on the TMA-landed tiles); its dY column arrives as a fifth 32-column box behind the operand planes.
  // P.oddn: Ni = 128 + 1 likewise: X column 128 and the virtual bias column are multiplied by the converter threads too
  // (thread m: sums over nodes of dY[node][m] X[node][128] and dY[node][m] e[node]), so the B operand is exactly four
  // 32-column boxes and a third pipeline stage fits in shared memory.
  // P.atmem: the A operand (dY^T) lives in TENSOR MEMORY.  A converter thread owns one output feature m (= TMEM lane) and
  // 16 of the tile's 32 nodes: it reads dY[node][m] from the TMA-landed tile (column reads: one wavefront per node and
  // warp), splits and writes (hi, lo) with tcgen05.st into one of two 64-column operand stages -- a transposed copy for
  // free.  No hi / lo planes of A in shared memory (a stage shrinks by 16 KB: a fourth stage fits at hidden 129), no
  // in-place rewrite of the A tile and no A descriptor reads by the MMAs: ~640 of the ~2650 shared-memory wavefronts per
  // tile that bound this kernel (profiles/r2_summary.md).  Costs one rotating accumulator (two instead of three).
  const bool atmem = P.atmem != 0;
  const uint32_t b_off = atmem ? a_bytes : 2u * a_bytes;                     // A (as landed) | B_hi | B_lo | odd boxes
  const uint32_t oddy_off = b_off + 2u * b_bytes;                            // dY column 128..159 (P.odd)
  const uint32_t oddx_off = oddy_off + (P.odd ? 4096u : 0u);                 // X columns 128..159 (P.oddn)
  const uint32_t odd_bytes = (P.odd ? 4096u : 0u) + (P.oddn ? 4096u : 0u);
  __shared__ float evec[kTcMaxStages][kTcBK];  // the bias column's entries (ones / extra_vec) of the tile in each stage
  const uint32_t stage_bytes = oddy_off + odd_bytes;  // (A_hi | A_lo without atmem)
  const uint32_t bar_base = base + uint32_t(S) * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (kTcMaxStages + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * kTcMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * kTcMaxStages);
  const uint32_t tmem_slot = accum_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = mgrp * mt * kTcBM;  // first dW row of this CTA
  const int k_beg = split * args.kchunk, k_end = min(args.K, k_beg + args.kchunk);
  const int n_tiles = (k_end > k_beg) ? (k_end - k_beg + kTcBK - 1) / kTcBK : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&P.a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&P.b)) : "memory");
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), kTcWorkers / 32);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(uint32_t(P.tmem_cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) WSTAMP(1);

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        const int k0 = k_beg + it * kTcBK;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), a_bytes + b_bytes + odd_bytes);
        const uint32_t st = base + uint32_t(s) * stage_bytes;
        for (int b = 0; b < mt * 4; ++b) tma_load_2d(st + uint32_t(b) * 4096u, &P.a, row0 + 32 * b, k0, full_bar(s));
        if (P.odd) tma_load_2d(st + oddy_off, &P.a, kTcBM, k0, full_bar(s));
        if (P.oddn) tma_load_2d(st + oddx_off, &P.b, 128, k0, full_bar(s));
        for (int b = 0; b < nb; ++b) tma_load_2d(st + b_off + uint32_t(b) * 4096u, &P.b, 32 * b, k0, full_bar(s));
      }
    }
  } else if (warp == 1) {
    // D = F32, A = B = TF32, A and B MN-major (bits 15, 16), N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(BN >> 3) << 17) |
                           (uint32_t(kTcBM >> 4) << 24);
    const uint32_t lbo = 4096u, sbo = 512u, kstep = 1024u;
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(conv_bar(s), ph);
      tc_fence_after();
      if (lane == 0 && it < 12) WSTAMP(30 + it);
      if (elect_one()) {
        const uint32_t st = base + uint32_t(s) * stage_bytes;
        const uint32_t b_hi = st + b_off, b_lo = b_hi + b_bytes;
        for (int j = 0; j < kTcBK / 8; ++j) {  // 8 nodes per UMMA K step = one 1024-byte K group
          const uint32_t koff = uint32_t(j) * kstep;
          const uint64_t dbh = umma_desc_mn128(b_hi + koff, lbo, sbo), dbl = umma_desc_mn128(b_lo + koff, lbo, sbo);
          const int k_idx = it * (kTcBK / 8) + j;
          if (atmem) {  // (mt == 1) A from the tile's operand stage in tensor memory: K-major there, so without the A-major bit
            const uint32_t at_hi = tmem_base + uint32_t(P.a_tmem + 64 * (it & 1) + 8 * j), at_lo = at_hi + 32u;
            const uint32_t idesc_ts = idesc & ~(1u << 15);
            const uint32_t d_lo = tmem_base + uint32_t(P.acc_lo[0]), d_hi = tmem_base + uint32_t((k_idx % P.n_hi) * BN);
            umma_tf32_ts(d_lo, at_lo, dbh, idesc_ts, k_idx > 0 ? 1u : 0u);
            umma_tf32_ts(d_lo, at_hi, dbl, idesc_ts, 1u);
            umma_tf32_ts(d_hi, at_hi, dbh, idesc_ts, k_idx >= P.n_hi ? 1u : 0u);
            continue;
          }
          for (int t = 0; t < mt; ++t) {
            const uint32_t a_hi = st + uint32_t(t) * 16384u, a_lo = a_hi + a_bytes;
            const uint64_t dah = umma_desc_mn128(a_hi + koff, lbo, sbo), dal = umma_desc_mn128(a_lo + koff, lbo, sbo);
            const uint32_t d_lo = tmem_base + uint32_t(P.acc_lo[t]);
            const uint32_t first = k_idx > 0 ? 1u : 0u;
            // one M tile per CTA: the hi*hi products rotate over P.n_hi accumulators, so each one takes 1/n_hi of the
            // truncating fp32 adds (the tensor core's accumulate is not round-to-nearest; see kWgMaxChunkDefault)
            const uint32_t d_hi = tmem_base + uint32_t(mt == 1 ? (k_idx % P.n_hi) * BN : P.acc_hi[t]);
            const uint32_t acc_hi = mt == 1 ? (k_idx >= P.n_hi ? 1u : 0u) : ((d_hi == d_lo) ? 1u : first);
            umma_tf32(d_lo, dal, dbh, idesc, first);
            umma_tf32(d_lo, dah, dbl, idesc, 1u);
            umma_tf32(d_hi, dah, dbh, idesc, acc_hi);
          }
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accum_bar);
    if (lane == 0) WSTAMP(44);
    __syncwarp();
  } else {
    const int tid_c = threadIdx.x - 64;
    float odd_acc = 0.f, col_acc = 0.f, bias_acc = 0.f, corner_acc = 0.f;
    float evec_next = 0.f;
    if (P.oddn && P.extra_col == 2 && tid_c < kTcBK && k_beg + tid_c < k_end) evec_next = P.extra_vec[k_beg + tid_c];
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      const int k0 = k_beg + it * kTcBK;
      mbar_wait(full_bar(s), ph);
      if (tid_c == 0 && it < 12) WSTAMP(2 + it);
      const uint32_t st = base + uint32_t(s) * stage_bytes;
      if (P.oddn) {
        if (tid_c < kTcBK) {
          // the entry of this tile was fetched one tile ahead (a global load here sat on the pipeline's critical path)
          const int node = k0 + tid_c;
          float ev = (P.extra_col != 0 && node < k_end) ? 1.f : 0.f;
          if (P.extra_col == 2) ev = node < k_end ? evec_next : 0.f;
          evec[s][tid_c] = ev;
          const int nn = node + kTcBK;
          if (P.extra_col == 2) evec_next = nn < k_end ? P.extra_vec[nn] : 0.f;
        }
        if (tid_c == 0 && it == 4) WSTAMP(50);
        asm volatile("bar.sync 1, %0;" ::"n"(kTcWorkers) : "memory");
        if (tid_c == 0 && it == 4) WSTAMP(51);
        const uint32_t ycol = st + oddy_off, xodd = st + oddx_off;
        // all eight converter warps: thread -> (m = tid mod 128, rows [16 (tid / 128), +16)); the halves meet in the epilogue
        {
          // column 128 and the bias column of dW rows 0..127:  sum_r dY[r][m] * X[r][128],  sum_r dY[r][m] * e[r]
          const int mrow = tid_c & 127, rbeg = (tid_c >> 7) * 16;
          const int bb = mrow >> 5, cc = mrow & 31;
          const uint32_t ym = st + uint32_t(bb) * 4096u + uint32_t(cc & 7) * 4u;
#pragma unroll
          for (int r8 = 0; r8 < 16; r8 += 8) {  // all loads of eight rows before the first use (see split_tile)
            float y[8], x[8], e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = rbeg + r8 + u;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[u]) : "r"(ym + uint32_t(r) * 128u + (uint32_t((cc >> 3) ^ (r & 3)) << 5)));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[u]) : "r"(xodd + uint32_t(r) * 128u + (uint32_t(r & 3) << 5)));
              e[u] = evec[s][r];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              col_acc = fmaf(y[u], x[u], col_acc);
              bias_acc = fmaf(y[u], e[u], bias_acc);
            }
          }
          if (P.odd && mrow < 2) {
            // the corner: row 128 of dW times column 128 (m = 0) / the bias column (m = 1)
#pragma unroll 8
            for (int r = rbeg; r < rbeg + 16; ++r) {
              float y, x;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y) : "r"(ycol + uint32_t(r) * 128u + (uint32_t(r & 3) << 5)));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(xodd + uint32_t(r) * 128u + (uint32_t(r & 3) << 5)));
              corner_acc = fmaf(y, mrow == 0 ? x : evec[s][r], corner_acc);
            }
          }
        }
      } else if (P.extra_col != 0) {
        // virtual column N of X := 1 (or extra_vec[node]) -> its dot products with dY are the bias gradient
        if (tid_c < kTcBK) {
          const int r = tid_c, cN = P.N, bb = cN >> 5, cc = cN & 31;
          const int node = k0 + r;
          float v = 0.f;
          if (node < k_end) v = P.extra_col == 1 ? 1.f : P.extra_vec[node];
          const uint32_t addr = st + b_off + uint32_t(bb) * 4096u + uint32_t(r) * 128u +
                                (uint32_t((cc >> 3) ^ (r & 3)) << 5) + uint32_t(cc & 7) * 4u;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTcWorkers) : "memory");
      }
      if (P.odd) {
        // row 128 of dW: sum_r dY[r][128] * X[r][n].  With P.oddn the 128 columns are split over all 256 threads by rows.
        const int ncol = P.oddn ? (tid_c & 127) : tid_c;
        const int rbeg = P.oddn ? (tid_c >> 7) * 16 : 0, rend = P.oddn ? rbeg + 16 : kTcBK;
        if (P.oddn || tid_c < P.n_eff) {
          const int bb = ncol >> 5, cc = ncol & 31;
          const uint32_t xcol = st + b_off + uint32_t(bb) * 4096u + uint32_t(cc & 7) * 4u;
          const uint32_t ycol = st + oddy_off;
          for (int r8 = rbeg; r8 < rend; r8 += 8) {  // 128-byte rows, 32-byte chunks XOR-swizzled with (row mod 4)
            float x[8], y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = r8 + u;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[u]) : "r"(xcol + uint32_t(r) * 128u + (uint32_t((cc >> 3) ^ (r & 3)) << 5)));
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y[u]) : "r"(ycol + uint32_t(r) * 128u + (uint32_t(r & 3) << 5)));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) odd_acc = fmaf(y[u], x[u], odd_acc);
          }
        }
      }
      if (tid_c == 0 && it == 4) WSTAMP(52);
      if (P.odd || P.oddn) asm volatile("bar.sync 1, %0;" ::"n"(kTcWorkers) : "memory");  // the split below rewrites the tiles in place
      if (tid_c == 0 && it == 4) WSTAMP(53);
      if (atmem) {
        // feature m = 32 (warp mod 4) + lane (the TMEM lanes this warp may access), nodes [16 nh, 16 nh + 16) of the tile
        const int m = 32 * (warp & 3) + lane, nh = (warp - 2) >> 2;
        const uint32_t ym = st + uint32_t(m >> 5) * 4096u + uint32_t(m & 7) * 4u;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int r = 16 * nh + u;
          float v;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ym + uint32_t(r) * 128u + (uint32_t(((m & 31) >> 3) ^ (r & 3)) << 5)));
          const float h = tf32_rn(v);
          hi[u] = __float_as_uint(h);
          lo[u] = __float_as_uint(tf32_rn(v - h));
        }
        if (it >= 2) {  // the operand stage was last read by the MMAs of tile it - 2
          mbar_wait(empty_bar((it - 2) % S), uint32_t((it - 2) / S) & 1u);
          tc_fence_after();
        }
        const uint32_t at = tmem_base + (uint32_t(32 * (warp & 3)) << 16) + uint32_t(P.a_tmem + 64 * (it & 1) + 16 * nh);
        tmem_st16(at, hi);
        tmem_st16(at + 32u, lo);
        tmem_st_wait();
        tc_fence_before();
      } else {
        split_tile(st, a_bytes, int(a_bytes / 16u), tid_c);
      }
      split_tile(st + b_off, b_bytes, int(b_bytes / 16u), tid_c);
      if (tid_c == 0 && it == 4) WSTAMP(54);
      proxy_fence_async();
      if (tid_c == 0 && it < 12) WSTAMP(16 + it);
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar(s));
    }
    // epilogue: accumulators -> padded smem tile -> coalesced rows of the split-K partial buffer
    const int wk = warp - 2, q = warp & 3, half = wk >> 2;
    const int n_eff = P.n_eff, Mo = P.Mo;
    const int n_main = P.oddn ? kTcBM : n_eff;  // columns that come from the tensor core
    float* __restrict__ const part = args.partial + P.part_off + size_t(split) * size_t(Mo) * n_eff;
    const uint32_t tile_ld = uint32_t(BN) + 4u;
    if (n_tiles > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");  // (see k_gemm_tc: orders split_tile's stores before the staging tile's)
    if (tid_c == 0) WSTAMP(45);
    for (int t = 0; t < mt; ++t) {
      if (row0 + t * kTcBM >= Mo) break;  // CTA-uniform
      if (n_tiles > 0) {
        const uint32_t row_addr = base + (uint32_t(32 * q + lane) * tile_ld) * 4u;
        const uint32_t lane_base = tmem_base + (uint32_t(32 * q) << 16);
        const int n_rot = mt == 1 ? min(P.n_hi, n_tiles * (kTcBK / 8)) : 0;  // rotating accumulators actually written
        for (int c0 = 16 * half; c0 < n_main; c0 += 32) {
          uint32_t r[16];
          float acc[16];
          tmem_ld16(lane_base + uint32_t(P.acc_lo[t] + c0), r);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(r[i]);
          if (mt == 1) {
            for (int a = 0; a < n_rot; ++a) {  // round-to-nearest fp32 adds of the partial sums
              tmem_ld16(lane_base + uint32_t(a * BN + c0), r);
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
            }
          } else if (P.acc_hi[t] != P.acc_lo[t]) {
            tmem_ld16(lane_base + uint32_t(P.acc_hi[t] + c0), r);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + 4u * (c0 + i)), "f"(acc[i]), "f"(acc[i + 1]),
                         "f"(acc[i + 2]), "f"(acc[i + 3]) : "memory");
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
      for (int rr = 0; rr < 16; ++rr) {
        const int row = 16 * wk + rr, m = row0 + t * kTcBM + row;
        if (m >= Mo) break;
        const uint32_t row_addr = base + (uint32_t(row) * tile_ld) * 4u;
        for (int c = lane; c < n_main; c += 32) {
          float v = 0.f;
          if (n_tiles > 0) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(row_addr + 4u * c));
          part[size_t(m) * n_eff + c] = v;
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
    }
    if (P.oddn) {
      // the two row halves of the border sums meet here (threads t and t + 128 hold the same m / n)
      const uint32_t scr = base + uint32_t(tid_c) * 16u;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(scr), "f"(odd_acc), "f"(col_acc), "f"(bias_acc), "f"(corner_acc) : "memory");
      asm volatile("bar.sync 2, %0;" ::"n"(kTcWorkers) : "memory");
      if (tid_c < kTcBM) {
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(scr + 128u * 16u));
        odd_acc += o.x;
        col_acc += o.y;
        bias_acc += o.z;
        corner_acc += o.w;
        if (P.odd) {
          part[size_t(kTcBM) * n_eff + tid_c] = odd_acc;
          if (tid_c == 0) part[size_t(kTcBM) * n_eff + kTcBM] = corner_acc;
          if (tid_c == 1 && n_eff > kTcBM + 1) part[size_t(kTcBM) * n_eff + kTcBM + 1] = corner_acc;
        }
        if (tid_c < Mo) {
          part[size_t(tid_c) * n_eff + kTcBM] = col_acc;
          if (n_eff > kTcBM + 1) part[size_t(tid_c) * n_eff + kTcBM + 1] = bias_acc;
        }
      }
    } else if (P.odd && tid_c < n_eff) {
      part[size_t(kTcBM) * n_eff + tid_c] = odd_acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) WSTAMP(46);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(P.tmem_cols)) : "memory");
  }
}

// sum the node-chunk partials of every problem in fixed order (deterministic) and scatter into dW / dbias
__global__ void __launch_bounds__(256) k_wgrad_group_reduce(const __grid_constant__ WgGroupArgs args) {
  pdl_wait();
  const WgGroupProb& P = args.p[blockIdx.y];
  const int n_eff = P.n_eff, total = P.Mo * n_eff, S = args.splitk;
  const float* __restrict__ src0 = args.partial + P.part_off;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    float sum = 0.f;
    int s = 0;
    for (; s + 4 <= S; s += 4) {
      const float v0 = __ldg(src0 + size_t(s) * total + idx), v1 = __ldg(src0 + size_t(s + 1) * total + idx);
      const float v2 = __ldg(src0 + size_t(s + 2) * total + idx), v3 = __ldg(src0 + size_t(s + 3) * total + idx);
      sum += (v0 + v1) + (v2 + v3);
    }
    for (; s < S; ++s) sum += __ldg(src0 + size_t(s) * total + idx);
    const int m = idx / n_eff, n = idx - m * n_eff;
    if (n == P.N) {
      if (P.dbias != nullptr) P.dbias[m] = sum;
    } else {
      P.dW[size_t(m) * P.lddw + n] = sum;
    }
  }
}

// ---- weight packing ------------------------------------------------------------------------------------
// state_dict weights are [out, in] with arbitrary row pitch (129 floats = 516 B is not a legal TMA stride) and the
// data-gradient GEMMs need them transposed: one small kernel per step copies every matrix into 16-byte-pitched
// K-major buffers (dst = W, dstT = W^T), each as three planes: the fp32 value, its TF32 hi part and its TF32 lo part,
// so the GEMMs TMA-load hi and lo directly and only activations are split on chip.
struct PackItem {
  const float* src;
  float* dst;
  float* dst_t;
  int ld_src, rows, cols, ld_dst, ld_dst_t, pad_;
};
constexpr int kPackMax = 64;
struct PackArgs {
  PackItem it[kPackMax];
};
static_assert(sizeof(PackArgs) <= 4000, "kernel parameter space");

__global__ void k_pack_weights(const __grid_constant__ PackArgs args) {
  pdl_wait();
  const PackItem& p = args.it[blockIdx.y];
  const int total = p.rows * p.cols;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / p.cols, c = idx - r * p.cols;
    const float v = p.src[size_t(r) * p.ld_src + c];
    const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
    if (p.dst != nullptr) {
      const size_t plane = size_t(p.rows) * p.ld_dst, o = size_t(r) * p.ld_dst + c;
      p.dst[o] = v;
      p.dst[plane + o] = hi;
      p.dst[2 * plane + o] = lo;
    }
    if (p.dst_t != nullptr) {
      const size_t plane = size_t(p.cols) * p.ld_dst_t, o = size_t(c) * p.ld_dst_t + r;
      p.dst_t[o] = v;
      p.dst_t[plane + o] = hi;
      p.dst_t[2 * plane + o] = lo;
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr || rows <= 0 || cols <= 0) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  cuuint32_t box[2] = {kTcBK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  auto encode = [&]() {
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = encode();
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    // a driver-API call on a thread that has not touched the runtime yet (autograd's backward thread): bind the
    // primary context of the current device to this thread, then retry
    cudaFree(nullptr);
    r = encode();
  }
  return r == CUDA_SUCCESS;
}

bool tma_ok(const float* p, int64_t ld) { return p != nullptr && aligned16(p) && ld % 4 == 0; }

}  // namespace

bool tc_make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  return make_map(map, ptr, rows, cols, ld, box_rows);
}

bool tc_enabled() {
  static int state = -1;
  if (state < 0) {
    const char* e = std::getenv("PFN_GEMM");
    state = (e != nullptr && std::strcmp(e, "ffma") == 0) ? 0 : 1;
    if (state == 1 && encode_fn() == nullptr) state = 0;
  }
  return state == 1;
}

// Tensor-core launch of a GemmArgs problem whose operands are K-major (A row-major [M,K], B row-major [N,K]).
// Returns 1 if the shapes/alignments do not fit this path (caller falls back to the FFMA kernel), 0 on success.
int gemm_tc_launch(const GemmArgs& g, cudaStream_t stream) {
  if (!tc_enabled() || g.splitk > 1 || g.extra_col != 0 || g.M <= 0 || g.N <= 0) return 1;
  const int count = g.batched ? g.n_items : 1;
  for (int i = 0; i < g.n_items; ++i) {
    const GemmItem& it = g.it[i];
    if (it.a_cs != 1 || it.b_rs != 1 || !tma_ok(it.A, it.a_rs) || !tma_ok(it.B, it.b_cs) || it.K <= 0) return 1;
  }
  TcArgs a;
  std::memset(&a, 0, sizeof(a));
  // N tile: the whole width while it leaves room for >= 2 rotating hi*hi accumulators (N <= 160), else 128-wide tiles
  // (three rotating accumulators): a 256-wide tile has ONE, and the truncating accumulate then costs ~4e-6 relative at
  // K = 512 -- enough to flip ReLU masks downstream and push wide models' gradients outside the 1e-5 contract.
  const int64_t n16 = round_up64(g.N, 16);
  const int bn = static_cast<int>(n16 <= 160 ? n16 : 128);
  a.BN = bn;
  a.n_items = g.n_items;
  a.batched = g.batched;
  a.M = g.M;
  a.N = g.N;
  // A operand in tensor memory (default; PFN_TC_ATMEM=0 keeps the hi / lo planes in shared memory): two operand stages of
  // 64 columns behind the accumulators, which leaves (512 - 128) / BN accumulators
  static const bool atmem_env = [] {
    const char* e = std::getenv("PFN_TC_ATMEM");
    return !(e != nullptr && e[0] == '0');
  }();
  const bool atmem = atmem_env && (512 - 128) / bn >= 3;  // (at least two rotating hi*hi accumulators + lo must remain)
  const uint32_t stage_bytes = (atmem ? (kTcBM + 2u * bn) : 2u * (kTcBM + bn)) * 128u;
  a.stages = static_cast<int>(std::min<uint32_t>(kTcMaxStages, (kSmemLimit - 2048u) / stage_bytes));
  if (a.stages < 1) return 1;
  if (uint32_t(a.stages) * stage_bytes < kTcBM * uint32_t(bn + 4) * 4u) return 1;  // (the epilogue's staging tile overlays the stages)
  a.n_hi = std::max(1, std::min(3, (atmem ? 384 : 512) / bn - 1));
  a.a_tmem = (a.n_hi + 1) * bn;
  int cols = 32;
  while (cols < (a.n_hi + 1) * bn + (atmem ? 128 : 0)) cols <<= 1;
  a.tmem_cols = cols;
  for (int i = 0; i < g.n_items; ++i) {
    const GemmItem& it = g.it[i];
    const bool presplit = it.b_plane != 0;
    if (!make_map(&a.it[i].a, it.A, g.M, it.K, it.a_rs, kTcBM) ||
        !make_map(&a.it[i].b, it.B + (presplit ? it.b_plane : 0), g.N, it.K, it.b_cs, bn))
      return 1;
    if (presplit && !make_map(&a.it[i].b_lo, it.B + 2 * it.b_plane, g.N, it.K, it.b_cs, bn)) return 1;
    a.it[i].presplit = presplit ? 1 : 0;
    a.it[i].K = it.K;
    a.C[i] = it.C;
    a.bias[i] = it.bias;
  }
  a.ldc = g.it[0].ldc;
  for (int i = 1; i < count; ++i)
    if (g.it[i].ldc != a.ldc) return 1;
  a.rowscale = g.rowscale;
  a.addend = g.addend;
  a.ymask = g.ymask;
  a.inj = g.inj;
  a.ld_add = g.ld_add;
  a.ld_ym = g.ld_ym;
  a.ld_inj = g.ld_inj;
  a.act = g.act;
  a.scale = g.scale;
  a.seed_lo = g.seed_lo;
  a.seed_hi = g.seed_hi;
  a.keep_thresh = g.keep_thresh;
  a.seed_dev = g.seed_dev;
  const uint32_t smem = uint32_t(a.stages) * stage_bytes + 1024u + 8u * (3 * kTcMaxStages + 2 + 6);
  // long reductions flush the TMEM accumulators into registers every drain_tiles K tiles (see the kernel);
  // PFN_TC_DRAIN=0 / 1 forces the choice (accuracy experiments)
  int max_tiles = 0;
  for (int z = 0; z < count; ++z) {
    int tiles = 0;
    for (int i = (g.batched ? z : 0); i < (g.batched ? z + 1 : g.n_items); ++i) tiles += (g.it[i].K + kTcBK - 1) / kTcBK;
    max_tiles = std::max(max_tiles, tiles);
  }
  static const int drain_env = [] {
    const char* e = std::getenv("PFN_TC_DRAIN");
    return e == nullptr ? -1 : (e[0] == '0' ? 0 : 1);
  }();
  static const int drain_tiles_env = [] {
    const char* e = std::getenv("PFN_TC_DRAIN_TILES");
    return e != nullptr && e[0] == '1' ? 1 : (e != nullptr && e[0] == '2' ? 2 : kTcDrainTilesDefault);
  }();
  a.drain_tiles = drain_tiles_env;
  static const int debias_env = [] {
    const char* e = std::getenv("PFN_TC_DEBIAS");  // ulps added back per flushed phase sum (0 = off; experiments)
    return e != nullptr ? std::max(0, std::min(8, std::atoi(e))) : 1;
  }();
  a.debias_ulps = a.drain_tiles == 1 ? debias_env : 2 * debias_env;
  const bool drain = a.n_hi >= 2 && bn <= 32 * kTcDrainChunks && (drain_env >= 0 ? drain_env == 1 : max_tiles >= kTcDrainMinTiles);
  static SmemAttrOnce attr_once;
  PFN_CUDA_OK(ensure_dynamic_smem(attr_once, [] {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit));
    return e;
  }));
  dim3 grid(static_cast<unsigned>(ceil_div64(g.M, kTcBM)), static_cast<unsigned>(ceil_div64(g.N, bn)), static_cast<unsigned>(count));
  static const bool timing_on = std::getenv("PFN_TC_TIMING") != nullptr;  // debug aid: phase timestamps of CTA 0
  static long long* timing_dev = nullptr;
  if (timing_on) {
    if (timing_dev == nullptr) cudaMalloc(&timing_dev, 64 * sizeof(long long));
    cudaMemsetAsync(timing_dev, 0, 64 * sizeof(long long), stream);
    a.timing = timing_dev;
  }
  auto kernel = drain ? (atmem ? k_gemm_tc<true, true> : k_gemm_tc<true, false>) : (atmem ? k_gemm_tc<false, true> : k_gemm_tc<false, false>);
  PFN_CUDA_OK(launch_kernel(kernel, grid, dim3(kTcThreads), smem, stream, a));
  PFN_LAUNCHED();
  if (timing_on) {
    long long t[64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(t, timing_dev, sizeof(t), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc-timing] M=%d N=%d K0=%d items=%d batched=%d act=%d | setup %lld |", g.M, g.N, g.it[0].K, g.n_items, g.batched,
            g.act, t[1] - t[0]);
    for (int i = 0; i < 8 && t[2 + i]; ++i)
      fprintf(stderr, " tile%d: full@%lld conv@%lld mma@%lld |", i, t[2 + i] - t[0], t[10 + i] - t[0], t[18 + i] - t[0]);
    fprintf(stderr, " worker it5: top@%lld full@%lld conv@%lld ready@%lld drained@%lld | issuer it5: top@%lld conv@%lld n_hi=%d S=%d |", t[56] - t[0], t[57] - t[0], t[58] - t[0],
            t[59] - t[0], t[60] - t[0], t[61] - t[0], t[62] - t[0], a.n_hi, a.stages);
    fprintf(stderr, " mma_done@%lld accum@%lld epi1@%lld bar@%lld end@%lld || epi: entry@%lld bias@%lld b0@%lld b1@%lld b2@%lld b3@%lld vec@%lld tail@%lld\n",
            t[26] - t[0], t[27] - t[0], t[28] - t[0], t[29] - t[0], t[30] - t[0], t[32] - t[0], t[33] - t[0], t[34] - t[0], t[35] - t[0],
            t[36] - t[0], t[37] - t[0], t[38] - t[0], t[39] - t[0]);
  }
  return 0;
}

// Tensor-core weight gradient for a split-K GemmArgs (A = dY viewed [Mo, nodes], B = X viewed [nodes, Ni], both with the
// M/N index contiguous).  Fills the same partial buffer layout as the FFMA kernel; the caller runs k_splitk_reduce.
// Returns 1 when the problem does not fit (caller falls back), 0 on success; *splitk_out receives the split count used.
int wgrad_tc_launch(GemmArgs& g, cudaStream_t stream) {
  if (!tc_enabled() || !g.batched && g.n_items != 1) return 1;
  const int count = g.batched ? g.n_items : 1;
  const int Mo = g.M, Ni = g.N, n_eff = Ni + (g.extra_col ? 1 : 0);
  if (Mo <= 0 || Ni <= 0 || n_eff > 256 || g.partial == nullptr) return 1;
  const int K = g.it[0].K;
  for (int i = 0; i < count; ++i) {
    const GemmItem& it = g.it[i];
    if (it.a_rs != 1 || it.b_cs != 1 || it.K != K || !tma_ok(it.A, it.a_cs) || !tma_ok(it.B, it.b_rs)) return 1;
  }
  WgArgs a;
  std::memset(&a, 0, sizeof(a));
  a.count = count;
  a.Mo = Mo;
  a.N = Ni;
  a.n_eff = n_eff;
  a.BN = static_cast<int>(round_up64(n_eff, 16));
  a.nb = (a.BN + 31) / 32;
  const int m_tiles = static_cast<int>(ceil_div64(Mo, kTcBM));
  a.mt = (m_tiles >= 2 && 2 * a.BN <= 512) ? 2 : 1;
  const int m_groups = static_cast<int>(ceil_div64(m_tiles, a.mt));
  a.K = K;
  int used;
  if (a.mt == 1) {
    a.acc_hi[0] = 0; a.acc_lo[0] = a.BN; used = 2 * a.BN;
  } else if (3 * a.BN <= 512) {
    a.acc_hi[0] = 0; a.acc_lo[0] = a.BN; a.acc_hi[1] = a.acc_lo[1] = 2 * a.BN; used = 3 * a.BN;
  } else {
    a.acc_hi[0] = a.acc_lo[0] = 0; a.acc_hi[1] = a.acc_lo[1] = a.BN; used = 2 * a.BN;
  }
  int cols = 32;
  while (cols < used) cols <<= 1;
  a.tmem_cols = cols;
  const uint32_t stage_bytes = 2u * (uint32_t(a.mt) * 16384u + uint32_t(a.nb) * 4096u);
  a.stages = static_cast<int>(std::min<uint32_t>(kTcMaxStages, (kSmemLimit - 2048u) / stage_bytes));
  if (a.stages < 2) return 1;
  // split the node dimension so that about one CTA per SM is in flight
  int64_t want = std::max<int64_t>(1, int64_t(sm_count()) / std::max(1, count * m_groups));
  int64_t kchunk = round_up64(std::max<int64_t>(ceil_div64(K, want), kTcBK), kTcBK);
  a.kchunk = static_cast<int>(kchunk);
  a.splitk = static_cast<int>(std::max<int64_t>(1, ceil_div64(K, kchunk)));
  if (size_t(a.splitk) * count * size_t(Mo) * n_eff * sizeof(float) > g.partial_bytes) return 1;
  for (int i = 0; i < count; ++i) {
    const GemmItem& it = g.it[i];
    if (!make_map(&a.it[i].a, it.A, K, Mo, it.a_cs, kTcBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map(&a.it[i].b, it.B, K, Ni, it.b_rs, kTcBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return 1;
  }
  a.extra_col = g.extra_col;
  a.extra_vec = g.extra_vec;
  a.partial = g.partial;
  {
    a.lbo = 4096u;        // one TMA box (32 floats of M/N x 32 nodes) to the next
    a.sbo = 512u;         // 4 K-rows of 128 B: one SWIZZLE_128B_BASE32B atom to the next along K
    a.kstep_bytes = 1024u;  // 8 nodes per UMMA K step
    a.mn_major = 1;
  }
  g.splitk = a.splitk;  // the reduction pass must know how many partials were written
  g.kchunk = a.kchunk;
  const uint32_t smem = uint32_t(a.stages) * stage_bytes + 1024u + 8u * (3 * kTcMaxStages + 2);
  static SmemAttrOnce attr_once;
  PFN_CUDA_OK(ensure_dynamic_smem(attr_once, [] { return cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit)); }));
  dim3 grid(static_cast<unsigned>(a.splitk), static_cast<unsigned>(count), static_cast<unsigned>(m_groups));
  PFN_CUDA_OK(launch_kernel(k_wgrad_tc, grid, dim3(kTcThreads), smem, stream, a));
  PFN_LAUNCHED();
  return 0;
}


// Node-chunk count for the grouped weight gradient: chunks of at most wg_max_chunk() nodes (accuracy, see above), and among
// those the count that minimises (waves of one CTA per SM) x (32-node tiles per CTA).
static void wg_plan_chunks(int64_t nodes, int64_t items_per_split, int64_t* kchunk_out, int64_t* splitk_out) {
  const int64_t tiles = ceil_div64(std::max<int64_t>(nodes, 1), kTcBK);
  const int64_t max_tiles = std::max<int64_t>(1, wg_max_chunk() / kTcBK);
  const int64_t s_min = ceil_div64(tiles, max_tiles);
  int64_t best_s = s_min, best_cost = -1;
  for (int64_t sp = s_min; sp <= std::min<int64_t>(tiles, s_min + 24); ++sp) {
    const int64_t per = ceil_div64(tiles, sp);
    const int64_t real = ceil_div64(tiles, per);  // chunks actually needed with `per` tiles each
    const int64_t cost = ceil_div64(real * items_per_split, sm_count()) * (per + 3);  // + ~3 tiles of prologue/epilogue per CTA
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_s = real;
    }
  }
  const int64_t per = ceil_div64(tiles, best_s);
  *kchunk_out = per * kTcBK;
  *splitk_out = ceil_div64(tiles, per);
}

static size_t wgrad_batch_scratch_bytes(const WgradProblem* probs, int n, int64_t nodes) {
  if (n <= 0) return 0;
  int64_t groups = 0;
  for (int i = 0; i < n; ++i)
    groups += probs[i].Mo == kTcBM + 1 ? 1 : std::max<int64_t>(1, ceil_div64(ceil_div64(probs[i].Mo, kTcBM), 2));
  int64_t kchunk, splitk;
  wg_plan_chunks(nodes, groups, &kchunk, &splitk);
  size_t floats = 0;
  for (int i = 0; i < n; ++i) floats += size_t(splitk) * size_t(probs[i].Mo) * size_t(probs[i].Ni + 1);
  return floats * sizeof(float) + 256;
}

// All weight gradients of a backward pass in one launch (+ one reduction launch).  Returns 1 when some problem does not
// fit the tensor-core path (caller launches them one by one), 0 on success.
static int wgrad_batch_launch(const WgradProblem* probs, int n, int64_t nodes, float* partial, size_t partial_bytes, cudaStream_t stream) {
  if (!tc_enabled() || n <= 0 || n > kWgMaxProb || nodes <= 0 || partial == nullptr) return 1;
  static const bool disabled = std::getenv("PFN_WGRAD_GROUP") != nullptr && std::getenv("PFN_WGRAD_GROUP")[0] == '0';
  if (disabled) return 1;
  WgGroupArgs a;
  std::memset(&a, 0, sizeof(a));
  int items_per_split = 0;
  uint32_t smem_max = 0;
  for (int i = 0; i < n; ++i) {
    const WgradProblem& w = probs[i];
    WgGroupProb& P = a.p[i];
    const int n_eff = w.Ni + (w.extra_col ? 1 : 0);
    if (w.Mo <= 0 || w.Ni <= 0 || n_eff > 256 || !tma_ok(w.dY, w.lddy) || !tma_ok(w.X, w.ldx)) return 1;
    P.Mo = w.Mo;
    P.N = w.Ni;
    P.n_eff = n_eff;
    P.odd = (w.Mo == kTcBM + 1) ? 1 : 0;  // row 128 of dW is summed by the converter threads (see the kernel)
    P.oddn = (w.Ni == kTcBM + 1 && w.Mo <= kTcBM + 1) ? 1 : 0;  // and so are column 128 and the bias column
    P.BN = P.oddn ? kTcBM : static_cast<int>(round_up64(n_eff, 16));
    P.nb = (P.BN + 31) / 32;
    const int m_tiles = P.odd ? 1 : static_cast<int>(ceil_div64(w.Mo, kTcBM));
    // two M tiles per CTA only for the 129..256-row case; wider dW (hidden 512) takes one tile per CTA so that every tile
    // keeps its own hi*hi and lo accumulators
    P.mt = (m_tiles == 2 && 2 * P.BN <= 512) ? 2 : 1;
    P.m_groups = static_cast<int>(ceil_div64(m_tiles, P.mt));
    int used;
    P.n_hi = 1;
    static const bool atmem_env = [] {
      const char* e = std::getenv("PFN_WG_ATMEM");
      return !(e != nullptr && e[0] == '0');
    }();
    P.atmem = (atmem_env && P.mt == 1 && 384 / P.BN - 1 >= 2) ? 1 : 0;  // (two rotating accumulators + lo + 128 operand columns)
    if (P.mt == 1) {
      P.n_hi = std::max(1, std::min(3, (P.atmem ? 384 : 512) / P.BN - 1));
      P.acc_hi[0] = 0; P.acc_lo[0] = P.n_hi * P.BN; used = (P.n_hi + 1) * P.BN;
      P.a_tmem = used;
      if (P.atmem) used += 128;
    } else if (3 * P.BN <= 512) {
      P.acc_hi[0] = 0; P.acc_lo[0] = P.BN; P.acc_hi[1] = P.acc_lo[1] = 2 * P.BN; used = 3 * P.BN;
    } else {
      P.acc_hi[0] = P.acc_lo[0] = 0; P.acc_hi[1] = P.acc_lo[1] = P.BN; used = 2 * P.BN;
    }
    int cols = 32;
    while (cols < used) cols <<= 1;
    P.tmem_cols = cols;
    const uint32_t stage_bytes = (P.atmem ? 1u : 2u) * uint32_t(P.mt) * 16384u + 2u * uint32_t(P.nb) * 4096u + (P.odd ? 4096u : 0u) + (P.oddn ? 4096u : 0u);
    P.stages = static_cast<int>(std::min<uint32_t>(kTcMaxStages, (kSmemLimit - 3072u) / stage_bytes));
    if (P.stages < 2) return 1;
    smem_max = std::max(smem_max, uint32_t(P.stages) * stage_bytes + 1024u + 8u * (3 * kTcMaxStages + 2));
    if (!make_map(&P.a, w.dY, nodes, w.Mo, w.lddy, kTcBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
        !make_map(&P.b, w.X, nodes, w.Ni, w.ldx, kTcBK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return 1;
    P.dW = w.dW;
    P.dbias = w.dbias;
    P.extra_col = w.extra_col;
    P.extra_vec = w.extra_vec;
    P.lddw = w.lddw;
    items_per_split += P.m_groups;
  }
  int64_t kchunk, splitk_plan;
  wg_plan_chunks(nodes, items_per_split, &kchunk, &splitk_plan);
  a.K = static_cast<int>(nodes);
  a.kchunk = static_cast<int>(kchunk);
  a.splitk = static_cast<int>(splitk_plan);
  a.n_prob = n;
  a.partial = partial;
  int item = 0;
  size_t floats = 0;
  for (int i = 0; i < n; ++i) {
    a.p[i].item0 = item;
    item += a.p[i].m_groups * a.splitk;
    a.p[i].part_off = static_cast<long long>(floats);
    floats += size_t(a.splitk) * size_t(a.p[i].Mo) * size_t(a.p[i].n_eff);
  }
  if (floats * sizeof(float) > partial_bytes) return 1;
  static SmemAttrOnce attr_once;
  // (the kernel also has 512 bytes of static shared memory: the opt-in maximum is for the sum)
  PFN_CUDA_OK(ensure_dynamic_smem(attr_once, [] { return cudaFuncSetAttribute(k_wgrad_group, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit - 1024u)); }));
  static const bool timing_on = std::getenv("PFN_WG_TIMING") != nullptr;
  static long long* timing_dev = nullptr;
  if (timing_on) {
    if (timing_dev == nullptr) cudaMalloc(&timing_dev, 64 * sizeof(long long));
    cudaMemsetAsync(timing_dev, 0, 64 * sizeof(long long), stream);
    a.timing = timing_dev;
  }
  PFN_CUDA_OK(launch_kernel(k_wgrad_group, dim3(static_cast<unsigned>(item)), dim3(kTcThreads), smem_max, stream, a));
  PFN_LAUNCHED();
  if (timing_on) {
    long long t[64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(t, timing_dev, sizeof(t), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[wg-timing] setup %lld |", t[1] - t[0]);
    for (int i = 0; i < 12 && t[2 + i]; ++i) fprintf(stderr, " t%d full@%lld conv@%lld mma@%lld |", i, t[2 + i] - t[0], t[16 + i] - t[0], t[30 + i] - t[0]);
    fprintf(stderr, " mma_done@%lld accum@%lld end@%lld || tile4: full@%lld evec@%lld bar@%lld dots@%lld bar@%lld split@%lld fence@%lld\n", t[44] - t[0], t[45] - t[0], t[46] - t[0],
            t[6] - t[0], t[50] - t[0], t[51] - t[0], t[52] - t[0], t[53] - t[0], t[54] - t[0], t[20] - t[0]);
  }
  int max_total = 1;
  for (int i = 0; i < n; ++i) max_total = std::max(max_total, a.p[i].Mo * a.p[i].n_eff);
  // one output element per thread: each sums ~30 partials, so the pass is latency-bound and wants many threads in flight
  PFN_CUDA_OK(launch_kernel(k_wgrad_group_reduce, dim3(static_cast<unsigned>(ceil_div64(max_total, 256)), static_cast<unsigned>(n)), dim3(256), 0, stream, a));
  PFN_LAUNCHED();
  return 0;
}


// Problems wider than one N tile (n_eff > 256: hidden 512) are cut into 128-column chunks of X -- independent problems
// that share dY; the bias column rides on the last chunk.
static std::vector<WgradProblem> wgrad_split_wide(const WgradProblem* probs, int n) {
  std::vector<WgradProblem> out;
  for (int i = 0; i < n; ++i) {
    const WgradProblem& w = probs[i];
    const int n_eff = w.Ni + (w.extra_col ? 1 : 0);
    if (n_eff <= 256) {
      out.push_back(w);
      continue;
    }
    for (int n0 = 0; n0 < w.Ni; n0 += 128) {
      WgradProblem c = w;
      c.X = w.X + n0;
      c.Ni = std::min(128, w.Ni - n0);
      c.dW = w.dW + n0;
      const bool last = n0 + 128 >= w.Ni;
      if (!last) {
        c.extra_col = 0;
        c.extra_vec = nullptr;
        c.dbias = nullptr;
      }
      out.push_back(c);
    }
  }
  return out;
}

size_t wgrad_group_scratch_bytes(const WgradProblem* probs, int n, int64_t nodes) {
  const std::vector<WgradProblem> work = wgrad_split_wide(probs, n);
  size_t worst = 0;
  for (size_t b = 0; b < work.size(); b += kWgMaxProb)
    worst = std::max(worst, wgrad_batch_scratch_bytes(work.data() + b, static_cast<int>(std::min<size_t>(kWgMaxProb, work.size() - b)), nodes));
  return worst;
}

// All weight gradients of a backward pass: one launch (+ one reduction launch) per batch of up to kWgMaxProb problems.
// Returns 1 when some problem does not fit the tensor-core path (nothing has been launched: the caller takes the
// per-problem route), 0 on success.
int wgrad_group_launch(const WgradProblem* probs, int n, int64_t nodes, float* partial, size_t partial_bytes, cudaStream_t stream) {
  if (!tc_enabled() || n <= 0 || nodes <= 0 || partial == nullptr) return 1;
  const std::vector<WgradProblem> work = wgrad_split_wide(probs, n);
  for (const WgradProblem& w : work)  // applicability of every problem is checked before the first launch
    if (w.Mo <= 0 || w.Ni <= 0 || w.Ni + (w.extra_col ? 1 : 0) > 256 || !tma_ok(w.dY, w.lddy) || !tma_ok(w.X, w.ldx)) return 1;
  if (wgrad_group_scratch_bytes(probs, n, nodes) > partial_bytes) return 1;
  for (size_t b = 0; b < work.size(); b += kWgMaxProb) {
    const int rc = wgrad_batch_launch(work.data() + b, static_cast<int>(std::min<size_t>(kWgMaxProb, work.size() - b)), nodes, partial,
                                      partial_bytes, stream);
    if (rc != 0) return (rc == 1 && b > 0) ? PFN_E_UNSUPPORTED : rc;  // (no falling back once a batch has been launched)
  }
  return 0;
}

int pack_weights_launch(const PackDesc* items, int n, cudaStream_t stream) {
  for (int base = 0; base < n; base += kPackMax) {
    PackArgs args;
    std::memset(&args, 0, sizeof(args));
    const int cnt = std::min(kPackMax, n - base);
    int max_elems = 1;
    for (int i = 0; i < cnt; ++i) {
      const PackDesc& d = items[base + i];
      args.it[i] = PackItem{d.src, d.dst, d.dst_t, d.ld_src, d.rows, d.cols, d.ld_dst, d.ld_dst_t, 0};
      max_elems = std::max(max_elems, d.rows * d.cols);
    }
    dim3 grid(static_cast<unsigned>(std::min<int64_t>(ceil_div64(max_elems, 256), 64)), static_cast<unsigned>(cnt));
    PFN_CUDA_OK(launch_kernel(k_pack_weights, grid, dim3(256), 0, stream, args));
    PFN_LAUNCHED();
  }
  return 0;
}

}  // namespace pfn
