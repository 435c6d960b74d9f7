/*
 * pfn_b200.h -- C ABI of libpfn_b200.so: the sm_100a implementation of PowerFlowNet's hot path
 * (MaskEmbdMultiMPN forward + backward).
 *
 * The reference (StavrosOrf/PoweFlowNet) has no FFI/plugin API for this path: its only seam is the
 * Python nn.Module surface (networks/MPN.py:456-559, constructed at train.py:109-117, called at
 * utils/training.py:58,74).  The entry points below are therefore what a ctypes binding on the
 * reference side would bind; each cites the reference lines it replaces.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers, sizes and leading dimensions; no torch types; `stream` is a
 *    `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *  - every function returns 0 on success, a `cudaError_t` value or a negative PFN_E_* code otherwise;
 *    `pfn_last_error()` describes the most recent failure on the calling thread.  Nothing throws,
 *    nothing allocates device memory: all workspaces are owned by the caller.
 *  - nothing here synchronises the device (everything is stream-ordered) except `pfn_graph_meta`,
 *    which copies three integers to the host.
 *  - node-feature matrices are row-major fp32 with a leading dimension that is a multiple of 4 floats
 *    (16-byte rows); base pointers are 16-byte aligned.  Weights are in the reference's state_dict
 *    layout ([out, in] row-major) and are never repacked by the caller.
 *  - efeature_dim must be 2 (the dataset's real edge width, datasets/PowerFlowData.py:112-113).
 */
#ifndef PFN_B200_H_
#define PFN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PFN_API __attribute__((visibility("default")))
#else
#define PFN_API
#endif

#define PFN_E_INVALID (-1)   /* bad argument (null pointer, misaligned pointer, bad size) */
#define PFN_E_UNSUPPORTED (-2) /* configuration outside the implemented path                */
#define PFN_E_WORKSPACE (-3) /* caller-provided workspace too small                        */

/* ---- library ------------------------------------------------------------------------------- */
PFN_API const char* pfn_version(void);
PFN_API const char* pfn_last_error(void);
/* number of kernels this library has launched on the calling process since load (bench "gpu_launches") */
PFN_API uint64_t pfn_launch_count(void);

/* ---- in-library kernel timing (bench.py's roofline figure): while enabled, every launch site brackets its
 *      kernel(s) with CUDA events on the launching stream; pfn_profile_read synchronises on the recorded
 *      events and returns the summed device time and the number of launches of one category. ----------- */
#define PFN_PROF_EA_FWD 0     /* fused EdgeAggregation message+aggregate forward kernel */
#define PFN_PROF_EA_BWD 1     /* its backward (both passes + the dWe reduction)          */
#define PFN_PROF_HOP 2        /* TAGConv hop (forward and transposed)                    */
#define PFN_PROF_GEMM_FWD 3   /* dense Linear forward                                    */
#define PFN_PROF_GEMM_DGRAD 4 /* dense Linear data gradient                              */
#define PFN_PROF_GEMM_WGRAD 5 /* dense Linear weight gradient (incl. split-K reduction)  */
#define PFN_PROF_PREP 6       /* graph preparation (all passes)                          */
#define PFN_PROF_FUSED_FWD 7 /* graph-resident whole-forward kernel (pfn_mpn_forward_tiled)   */
#define PFN_PROF_FUSED_BWD 8 /* graph-resident backward kernel(s) (pfn_mpn_backward_tiled)     */
#define PFN_PROF_CATEGORIES 9
PFN_API int pfn_profile_enable(int on);
PFN_API int pfn_profile_read(int category, double* total_ms, int64_t* launches);

/* ---- graph preparation: networks/MPN.py:498-523 (is_directed / undirect_graph) + the CSR the
 *      message passing needs (PyG MessagePassing.propagate gathers, `degree`, `gcn_norm`) ---------- */
typedef struct pfn_graph_layout {
  /* byte offsets of each array inside the caller-allocated graph workspace */
  int64_t meta;      /* int32[8]: [0] reverse-of-first-edge found, [1] directed, [2] E (directed edges in
                        use), [3] error flag (node id out of range), [4] e_raw, [5] n_nodes            */
  int64_t rowptr_t;  /* int32[N+1]  CSR by TARGET node (edge_index[1])                                  */
  int64_t nbr_t;     /* int32[Ecap] source node of each in-edge, CSR order                              */
  int64_t eid_t;     /* int32[Ecap] directed-edge id of each slot (ascending inside a row => stable)    */
  int64_t ea_t;      /* float[Ecap*2] edge_attr in CSR-by-target order                                  */
  int64_t rowptr_s;  /* int32[N+1]  CSR by SOURCE node (edge_index[0]), for the transposed passes       */
  int64_t nbr_s;     /* int32[Ecap] target node of each out-edge                                        */
  int64_t eid_s;     /* int32[Ecap]                                                                     */
  int64_t ea_s;      /* float[Ecap*2]                                                                   */
  int64_t deg;       /* float[N]  in-degree (count of edges whose target is the node)                   */
  int64_t dis;       /* float[N]  deg^-1/2, 0 where deg == 0  (gcn_norm)                                */
  int64_t cursor;    /* int32[2N] scratch                                                               */
  int64_t total_bytes;
  int64_t e_cap;     /* 2*e_raw: capacity in directed edges                                             */
} pfn_graph_layout;

PFN_API int pfn_graph_layout_get(int64_t n_nodes, int64_t e_raw, pfn_graph_layout* out);

/* undirect_mode 1: reference semantics -- look at the first edge only, append all reversed edges when
 * its reverse is absent (networks/MPN.py:498-523); 0: take edge_index as it is (stand-alone layers).
 * edge_index: int64 [2, e_raw] with row stride `ei_row_stride` elements; edge_attr: float [e_raw, 2]. */
PFN_API int pfn_graph_prep(const int64_t* edge_index, int64_t ei_row_stride, const float* edge_attr,
                   int64_t n_nodes, int64_t e_raw, int undirect_mode, void* graph_ws, void* stream);

/* One-launch variant of pfn_graph_prep for batches laid out TILE BY TILE, the shape PyG's loader gives a batch of
 * equal-sized small graphs (utils/training.py:55-58 iterating a DataLoader over datasets/PowerFlowData.py): rows
 * [t*tile_rows, (t+1)*tile_rows) are closed under edge_index AND their edges are columns [t*P, (t+1)*P) of edge_index,
 * P = e_raw / n_tiles.  Fills the workspace with exactly the bytes pfn_graph_prep writes.  The layout is validated on
 * the device: a column whose endpoints leave its tile raises the same flag pfn_graph_tile_status reports (and the
 * graph-resident kernels then return NaN for that tile); call pfn_graph_prep instead for such batches.
 * pfn_graph_prep_tiled_supported: 1 when the shape qualifies (tile_rows <= 128 divides n_nodes, n_tiles divides e_raw, P <= 768);
 * pfn_graph_prep_tiled returns PFN_E_UNSUPPORTED otherwise. */
PFN_API int pfn_graph_prep_tiled_supported(int64_t n_nodes, int64_t e_raw, int64_t tile_rows);
PFN_API int pfn_graph_prep_tiled(const int64_t* edge_index, int64_t ei_row_stride, const float* edge_attr,
                                 int64_t n_nodes, int64_t e_raw, int undirect_mode, int64_t tile_rows, void* graph_ws,
                                 void* stream);
/* host_meta[0]=directed, [1]=E, [2]=error flag.  Synchronises `stream`. */
PFN_API int pfn_graph_meta(const void* graph_ws, int32_t* host_meta, void* stream);
/* materialise the undirected lists (what `undirect_graph` returns): ei_out int64 [2, E], ea_out float [E, 2] */
PFN_API int pfn_graph_export(const int64_t* edge_index, int64_t ei_row_stride, const float* edge_attr,
                     int64_t e_raw, int64_t e_out, int64_t* ei_out, float* ea_out, void* stream);

/* ---- fused EdgeAggregation message + aggregate: networks/MPN.py:23-28 (message) + :53 (propagate,
 *      aggr='add').  With W1 = [Wi | Wj | We] (columns of edge_aggr.0.weight), Hi = x Wi^T + b1,
 *      Hj = x Wj^T:   S[i] = sum_{e: tgt(e)=i} ReLU(Hi[i] + Hj[src(e)] + We ea[e]);
 *      the second Linear is applied after the sum by pfn_linear (sum aggregation is linear). ------- */
PFN_API int pfn_ea_fwd(const float* Hi, const float* Hj, int64_t ldh, const void* graph_ws, int64_t n_nodes,
               int64_t e_raw, const float* We, int64_t ldwe, float* S, int64_t lds, int64_t h, void* stream);
/* backward of the above: dHi, dHj [N, ldd]; dWe [h, 2] written with row stride lddwe (it is a column
 * block of the gradient of edge_aggr.0.weight).  scratch: at least pfn_ea_bwd_scratch_bytes(h). */
PFN_API size_t pfn_ea_bwd_scratch_bytes(int64_t h);
PFN_API int pfn_ea_bwd(const float* dS, int64_t ldds, const float* Hi, const float* Hj, int64_t ldh,
               const void* graph_ws, int64_t n_nodes, int64_t e_raw, const float* We, int64_t ldwe,
               float* dHi, float* dHj, int64_t ldd, float* dWe, int64_t lddwe, void* scratch,
               int64_t h, void* stream);

/* ---- TAGConv hop (PyG TAGConv.propagate with gcn_norm weights; call site networks/MPN.py:545):
 *      Y[i] = dis[i] * sum_{e: tgt(e)=i} dis[src(e)] X[src(e)]   (+ addend[i]) ;
 *      transpose != 0 walks the CSR by source instead (A_hat^T, used by the backward pass);
 *      ymask != NULL multiplies the result by (ymask > 0 ? scale : 0) (ReLU/dropout backward). ------ */
PFN_API int pfn_spmm_hop(const float* X, int64_t ldx, const void* graph_ws, int64_t n_nodes, int64_t e_raw,
                 int transpose, const float* addend, int64_t ldadd, const float* ymask, int64_t ldym,
                 float scale, float* Y, int64_t ldy, int64_t h, void* stream);

/* ---- dense per-node Linear stacks (nn.Linear / PyG Linear; networks/MPN.py:17-21,491-495 and
 *      TAGConv.lins).  W is [n_out, n_in] row-major with leading dimension ldw (state_dict layout). -- */
#define PFN_ACT_NONE 0
#define PFN_ACT_RELU 1
#define PFN_ACT_DROPOUT_RELU 2 /* y = relu(keep ? v / (1-p) : 0): networks/MPN.py:546-547 */
/* Y[M,n_out] = X[M,n_in] W^T + rowscale (.) bias + addend, then activation.
 * rowscale NULL => 1.  dropout keep-mask: `inj_mask` (float 0/1, [M, ld_inj]) when non-NULL, else the
 * library's counter-based generator keyed by `seed`. */
PFN_API int pfn_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                   const float* rowscale, const float* addend, int64_t ldadd, float* Y, int64_t ldy,
                   int64_t M, int64_t n_in, int64_t n_out, int act, float dropout_p, uint64_t seed,
                   const float* inj_mask, int64_t ld_inj, void* stream);
/* dX[M,n_in] = dY[M,n_out] W ; if ymask != NULL: dX *= (ymask > 0 ? scale : 0). */
PFN_API int pfn_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, const float* ymask,
                     int64_t ldym, float scale, float* dX, int64_t lddx, int64_t M, int64_t n_in,
                     int64_t n_out, void* stream);
/* dW[n_out,n_in] = dY^T X (row stride lddw); dbias[n_out] = sum_m rowscale[m] dY[m,:] when dbias != NULL.
 * scratch: at least pfn_linear_wgrad_scratch_bytes(M, n_in, n_out). */
PFN_API size_t pfn_linear_wgrad_scratch_bytes(int64_t M, int64_t n_in, int64_t n_out);
PFN_API int pfn_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* rowscale,
                     float* dW, int64_t lddw, float* dbias, int64_t M, int64_t n_in, int64_t n_out,
                     void* scratch, void* stream);

/* ---- whole model: MaskEmbdMultiMPN.forward (networks/MPN.py:525-559) and its backward
 *      (autograd of the same, utils/training.py:74) ------------------------------------------------ */
typedef struct pfn_mpn_desc {
  int32_t nfeature_dim; /* 4 */
  int32_t efeature_dim; /* 2 */
  int32_t output_dim;   /* 4 */
  int32_t hidden_dim;
  int32_t n_gnn_layers; /* >= 2 */
  int32_t K;
  float dropout_rate;
  int32_t reserved;
} pfn_mpn_desc;

/* Number of parameter tensors and the order `params` / `grads` tables use:
 * for each entry of `layers` in order: EdgeAggregation -> edge_aggr.0.weight, edge_aggr.0.bias,
 * edge_aggr.2.weight, edge_aggr.2.bias; TAGConv -> lins.0.weight ... lins.K.weight, bias;
 * then mask_embd.0.weight, mask_embd.0.bias, mask_embd.2.weight, mask_embd.2.bias. */
PFN_API int pfn_mpn_num_params(const pfn_mpn_desc* desc);
/* activation workspace (lives from forward to backward) and scratch (transient) sizes in bytes */
PFN_API int pfn_mpn_workspace(const pfn_mpn_desc* desc, int64_t n_nodes, int64_t e_raw, size_t* act_bytes,
                      size_t* scratch_bytes);
/* x float [N,4]; pred_mask int64 [N,4]; graph_ws prepared by pfn_graph_prep(undirect_mode=1);
 * training != 0 applies dropout (keep masks from `inj_masks[i]` [N, hidden] when inj_masks != NULL,
 * else from the generator keyed by `seed`, or -- when seed_device != NULL -- by the 64-bit value the
 * kernels read from that DEVICE address at run time, so a captured CUDA graph draws fresh masks on every
 * replay); out float [N, output_dim]. */
PFN_API int pfn_mpn_forward(const pfn_mpn_desc* desc, const float* const* params, const float* x,
                    const int64_t* pred_mask, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                    void* act_ws, void* scratch_ws, int training, uint64_t seed,
                    const uint64_t* seed_device, const float* const* inj_masks, float* out, void* stream);
/* Graph-resident variant of pfn_mpn_forward: ONE kernel runs the whole layer stack for tiles of whole graphs with
 * the tile's activations kept in shared / tensor memory (small-graph regime: case14, case118).  The caller promises
 * that rows [t*tile_rows, (t+1)*tile_rows) are closed under edge_index for every t (tile_rows = nodes per graph x
 * graphs per tile, <= 128).  The kernel validates the promise: a tile with an edge leaving it (or more than 768
 * directed edges) gets NaN outputs and raises a flag that pfn_graph_tile_status reports (synchronises the stream).
 * Same workspaces, same saved activations and same results (to fp32 rounding) as pfn_mpn_forward, so
 * pfn_mpn_backward follows either.  pfn_mpn_fused_supported: 1 when (desc, tile_rows) is inside the kernel's range
 * (hidden_dim 129 or a multiple of 16 in [32,128], K <= 3, nfeature_dim 4, output_dim <= 4).
 * graph_ptr != NULL (device int64 [n_graphs + 1], PyG `Batch.ptr`): batches that MIX graph sizes (the reference's
 * `--case mixed`, datasets/PowerFlowData.py:67-70) -- whole graphs are packed greedily into tiles of at most 128 rows by
 * a device-side pass (tile_rows is then only the 128-row bound); a graph of more than 128 nodes raises the same flag. */
PFN_API int pfn_mpn_fused_supported(const pfn_mpn_desc* desc, int64_t tile_rows);
PFN_API int pfn_mpn_forward_tiled(const pfn_mpn_desc* desc, const float* const* params, const float* x,
                    const int64_t* pred_mask, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                    void* act_ws, void* scratch_ws, int training, uint64_t seed,
                    const uint64_t* seed_device, const float* const* inj_masks, float* out,
                    int64_t tile_rows, const int64_t* graph_ptr, int64_t n_graphs, void* stream);
PFN_API int pfn_graph_tile_status(const void* graph_ws, int32_t* violated, void* stream);
/* backward after pfn_mpn_forward_tiled (same tile_rows / n_graphs -- 0 when graph_ptr was NULL --, same promise): the TAGConv layers' data gradients run through
 * the graph-resident kernel too (hops and weight products commute, so d x_0 = sum_k ((A_hat^T)^k G) W_k is the forward
 * program on G with the CSR by source and the transposed weights); everything else as pfn_mpn_backward. */
PFN_API int pfn_mpn_backward_tiled(const pfn_mpn_desc* desc, const float* const* params, float* const* grads,
                     const float* dout, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                     void* act_ws, void* scratch_ws, int training, int64_t tile_rows, int64_t n_graphs,
                     void* stream);
/* dout float [N, output_dim]; grads[i] receives d loss / d params[i] (overwritten, same shapes) */
PFN_API int pfn_mpn_backward(const pfn_mpn_desc* desc, const float* const* params, float* const* grads,
                     const float* dout, int64_t n_nodes, int64_t e_raw, const void* graph_ws,
                     void* act_ws, void* scratch_ws, int training, void* stream);

/* ---- loss head (utils/training.py:61-72 with torch.nn.MSELoss, train.py:103): fused forward +
 *      gradient.  loss[0] = inv_count * sum (out-y)^2  (the mean when inv_count = 1/count; under data
 *      parallelism inv_count = 1/global_count gives this rank's share), dout = 2 (out - y) * inv_count.
 *      Deterministic two-pass reduction; scratch: at least pfn_mse_scratch_bytes(count). ------------ */
PFN_API size_t pfn_mse_scratch_bytes(int64_t count);
PFN_API int pfn_mse_fwd_bwd(const float* out, const float* y, int64_t count, float inv_count, float* loss,
                    float* dout, void* scratch, void* stream);

/* ---- Masked_L2_loss head (utils/custom_loss_functions.py:10-46, the default --train_loss_fn of
 *      utils/argument_parser.py:36-37; dispatched at utils/training.py:61-62): fused value + gradient.
 *      loss = mean over {mask != 0} of (out-y)^2 + regcoeff * mean over {mask == 0} of (out-y)^2 (second term only when
 *      `regularize`); the element counts stay on the device (the reference's masked_select costs two host syncs).
 *      mask: int64 [count] (data.pred_mask).  scratch: at least pfn_masked_l2_scratch_bytes(count).
 *      global_counts: NULL, or a DEVICE array of 2 floats = (number of mask != 0 entries, number of mask == 0 entries)
 *      over ALL ranks of a data-parallel step: the two means are then taken over the global selections, so that the SUM
 *      over ranks of `loss` is the loss of the whole batch and the SUM all-reduce of the parameter gradients equals the
 *      single-process gradient (the reference has no multi-GPU mode; 1-GPU results are the parity oracle). --------- */
PFN_API size_t pfn_masked_l2_scratch_bytes(int64_t count);
PFN_API int pfn_masked_l2_fwd_bwd(const float* out, const float* y, const int64_t* mask, int64_t count, int regularize,
                          float regcoeff, const float* global_counts, float* loss, float* dout, void* scratch,
                          void* stream);

/* ---- PowerImbalance loss + gradient: utils/custom_loss_functions.py:99-286 (`PowerImbalance.forward`; selected by
 *      train.py:95-101, called at utils/training.py:63-68).  x: normalised predictions [N, >=4] (Vm, Va[deg], P, Q),
 *      row pitch ldx floats (multiple of 4, 16-byte aligned).  graph_ws: a workspace filled by pfn_graph_prep with
 *      undirect_mode = 1 from the batch's edge_index / edge_attr (the loss doubles the branch list by the same first-edge
 *      rule as the model, :131-150).  stats: HOST array of 12 floats = xymean[4], xystd[4], edgemean[2], edgestd[2]
 *      (PowerFlowData.get_data_means_stds, datasets/PowerFlowData.py:115-117).  loss: device scalar = mean over buses
 *      of dP^2 + dQ^2.  dx: d loss / d x [N, 4] with row pitch lddx, or NULL for the forward value only.
 *      scratch: at least pfn_power_imbalance_scratch_bytes(N).  No atomics: sums follow the edge order. ------------- */
PFN_API size_t pfn_power_imbalance_scratch_bytes(int64_t n_nodes);
PFN_API int pfn_power_imbalance_fwd_bwd(const float* x, int64_t ldx, const void* graph_ws, int64_t n_nodes, int64_t e_raw,
                                const float* stats, float* loss, float* dx, int64_t lddx, void* scratch, void* stream);

/* ---- AdamW over a list of fp32 tensors in one launch: train.py:123 `torch.optim.AdamW` (torch/optim/adamw.py
 *      `_single_tensor_adamw`, amsgrad = maximize = False).  params / grads / exp_avg / exp_avg_sq / numel are HOST arrays
 *      of n_tensors device pointers / element counts; `step` is the 1-based count of this update (bias corrections are
 *      taken in double precision on the host, like torch's Python scalars).  Updates params and both moments in place. */
PFN_API int pfn_adamw_step(int64_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, void* stream);

/* ---- mini-batch assembly from a device-resident dataset: datasets/PowerFlowData.py:171-217 (`process`) + :119-140
 *      (`_normalize_dataset`) + PyG `Batch.from_data_list` (train.py:90, utils/training.py:55-56), in one pass.
 *      A dataset is the concatenation of up to 8 cases (`--case mixed`, :151-155); each case keeps the reference's RAW
 *      arrays in device memory: node_features [S, n, 6] = (index, type, Vm, Va, P, Q), edge_features [S, E, 4] =
 *      (from, to, r, x), fp32.  sample_ids: DEVICE int64 [batch], indices into the concatenation.  n_total / e_total:
 *      node / branch totals of the batch (the caller sized the outputs with them; checked on the device).
 *      norm: HOST array of 12 floats = xymean[4], xystd[4] + 1e-7, edgemean[2], edgestd[2] + 1e-7, or NULL for
 *      normalize=False.  random_bus_type_seed != 0 applies the `random_bus_type` transform (:36-40; bus_type is unused
 *      by the model).  Outputs (device): x, y [N, 4] f32; bus_type [N], pred_mask [N, 4], batch_vec [N], ptr [batch + 1]
 *      int64; edge_index [2, E] int64 (node offsets added); edge_attr [E, 2] f32.
 *      scratch: at least pfn_batch_assemble_scratch_bytes(batch).  pfn_batch_assemble_status reads back (synchronising)
 *      whether a sample id was out of range or the totals disagreed (flag != 0: outputs were not written). ---------- */
typedef struct pfn_dataset_case {
  const float* node_features;
  const float* edge_features;
  int64_t n_samples, n_nodes, n_edges;
} pfn_dataset_case;
PFN_API size_t pfn_batch_assemble_scratch_bytes(int64_t batch);
PFN_API int pfn_batch_assemble(const pfn_dataset_case* cases, int n_cases, const int64_t* sample_ids, int64_t batch,
                       int64_t n_total, int64_t e_total, const float* norm, uint64_t random_bus_type_seed,
                       float* x, float* y, int64_t* bus_type, int64_t* pred_mask, int64_t* edge_index,
                       float* edge_attr, int64_t* batch_vec, int64_t* ptr, void* scratch, void* stream);
PFN_API int pfn_batch_assemble_status(const void* scratch, int64_t batch, int32_t* host_flag, void* stream);

/* ---- all-reduce (SUM, fp32, in place) of the flat gradient buffer over NVLink peer memory: the data-parallel exchange of
 *      SURVEY.md section 8e as ONE kernel that can be captured inside the step's CUDA graph (the reference has no multi-GPU
 *      mode; nothing to mirror).  grad: n floats (n % (4 * world) == 0, 16-byte aligned), replaced by the sum over ranks.
 *      peer_recv / peer_res / peer_sig: DEVICE arrays of `world` pointers -- rank r's symmetric buffers as mapped into this
 *      process: recv = 2 * world * n floats (one-shot) or 2 * n floats (two-shot), res = 2 * n floats, sig = 2 * ctas * world
 *      uint32 (zero-initialised once); epochs: device array of `ctas` uint32 (zero-initialised once, local).  two_shot = 0:
 *      every rank pushes its whole buffer to every peer and sums locally; 1: reduce-scatter + all-gather (world > 4).
 *      Every rank must call with the same n / ctas / two_shot.  poweflownet_b200.parallel.OneShotAllReduce owns the buffers. */
PFN_API int pfn_allreduce_peer(float* grad, void* const* peer_recv, void* const* peer_res, void* const* peer_sig, void* epochs, int rank,
                       int world, int64_t n, int ctas, int two_shot, void* stream);
/* debug aid (PFN_AR_TIMING=1 in the environment): %globaltimer stamps (ns) CTA 0 of the LAST pfn_allreduce_peer call left at
 * its phase boundaries: start, after griddepcontrol.wait, after the pushes, after exchange 1, after sum + second pushes,
 * after exchange 2, end.  Synchronises the device. */
PFN_API int pfn_allreduce_debug_stamps(unsigned long long* host8);

#ifdef __cplusplus
}
#endif
#endif /* PFN_B200_H_ */
