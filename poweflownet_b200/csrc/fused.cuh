// Argument blocks of the graph-resident (tile-per-CTA) kernels: fused_fwd.cu.  Internal header.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfn {

constexpr int kFusedMaxLayers = 16;
constexpr int kFusedMaxSeg = 4;      // K + 1 <= 4 TAGConv segments
constexpr int kFusedEdgeCap = 768;   // directed edges of one tile

constexpr int kFusedEaSimt = 0;  // EdgeAggregation whose input width is nfeature_dim (first layer): input Linear by FMAs
constexpr int kFusedEaTc = 1;    // EdgeAggregation with hidden-width input: Hi | Hj on the tensor cores
constexpr int kFusedTag = 2;     // TAGConv

struct FLayer {
  int type, last, act, fin;
  int w_row[kFusedMaxSeg];  // first arena row of each packed weight: EA {Wi, Wj, W2, -} ; TAG {W_0 .. W_K}
  int w_rows;               // rows of those matrices (= hidden_dim); plane p of a weight starts at w_row + p * w_rows
  uint32_t seed_xor;        // per-layer salt of the dropout hash (same value as the layer-wise path)
  int ld_dest, pad_;
  const float* W1;    // EA: edge_aggr.0.weight [h, 2 fin + 2] as stored
  const float* b1;
  const float* W2;    // EA: edge_aggr.2.weight [fout, h] as stored
  const float* b2;
  const float* bias;  // TAG
  const float* inj;   // optional injected keep mask [N, h] (tests)
  float* save0;       // EA: Hi [N, ldh] (Hj and S follow, N*ldh apart) ; TAG: [x_0 | .. | x_K]  [N, (K+1) ldh]
  float* dest;        // layer output: block 0 of the next TAGConv's buffer / the TAGConv's Y / the model output
  // backward programs (modes 1..3); `dest` is then the gradient w.r.t. the layer input
  const float* gin;    // incoming gradient G [N, ld_gin] (read from global memory by the first step of a launch only)
  const float* ymask;  // saved layer input: d input *= (ymask > 0 ? 1/(1-p) : 0); null = no activation in front of the layer
  float* dhi;          // EA: dHi / dHj [N, ldh] out (the weight gradients read them)
  float* dhj;
  float* dwe_partial;  // EA: [2][4 * ceil(h/4)][n_tiles] per-tile partial sums of dWe, reduced by k_reduce_dwe
  int ld_gin, ld_ymask;
};

struct FusedArgs {
  CUtensorMap wmap;  // the packed-weight arena as one 2-D tensor [rows, h] with pitch ldh, box {32, 128}
  FLayer layers[kFusedMaxLayers];
  int n_layers, n_nodes, tile_rows, h, K, ldh, dropout, out_dim;
  const float* arena;
  const float* x;
  const int64_t* pred_mask;
  const float *mW1, *mb1, *mW2, *mb2;  // mask_embd.{0,2}.{weight,bias} as stored
  float *maskf, *t1, *x0;              // saved for the backward pass
  float* dt1;                          // mode 3: d loss / d (mask_embd hidden pre-activation) [N, ldh], written by the last step
  const int* rowptr;
  const int* nbr;
  const float2* ea;
  const float* deg;
  const float* dis;
  int* meta;
  float* out;
  float scale;
  uint32_t seed_lo, seed_hi, keep_thresh;
  const uint32_t* seed_dev;
  // mode 1 (TAGConv backward, one layer per launch): d x_0 = sum_k ((A_hat^T)^k G) W_k, masked by the layer input.
  // Hops and GEMMs commute (one acts on rows, the other on columns), so this is the forward TAGConv program run on G
  // with the CSR by source and the transposed weights; `gin` is G, `ymask` the saved layer input x_0.
  // mode 2 (EdgeAggregation backward, one layer per launch): dS = G W2, the two segmented passes (by source: dHj, by
  // target: dHi and dWe) with the ReLU mask recomputed from the saved Hi / Hj, d cur = dHj Wj + dHi Wi masked by the
  // layer input.  rowptr/nbr/ea = CSR by source, rowptr2/nbr2/ea2 = CSR by target.
  // mode 3 (whole backward data path): the steps of modes 2 and 1 for every layer in reverse order in ONE launch; the
  // gradient stays in the planes between layers exactly as the activation does in the forward.
  int mode, n_tiles, pad1_, pad2_;
  const int* rowptr2;
  const int* nbr2;
  const float2* ea2;
  // variable-size tiles (batches that mix graph sizes): tile t = rows [tile_start[t], tile_start[t+1]), t < meta[7]; the
  // grid is the number of graphs (an upper bound known on the host) and the surplus CTAs leave at once.  Null: uniform
  // tiles of `tile_rows` rows.
  const int* tile_start;
  long long* timing;  // debug (PFN_FUSED_TIMING): worker 0 of CTA 0 writes clock64() at phase boundaries
};
static_assert(sizeof(FusedArgs) <= 4000, "kernel parameter space");

bool fused_fwd_supported(int h, int K, int nfeature_dim, int output_dim, int64_t tile_rows);
constexpr int kFusedModeForward = 0, kFusedModeTagBackward = 1, kFusedModeEaBackward = 2, kFusedModeBackward = 3;
int fused_fwd_launch(FusedArgs& args, const float* arena, int64_t arena_rows, cudaStream_t stream);

}  // namespace pfn
