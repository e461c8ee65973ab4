/*
 * igcn_b200 -- C ABI of the B200-native IG-GCN graph-convolution hot path.
 *
 * Boundary contract:
 *   - every pointer is a DEVICE pointer unless the parameter is named host_*;
 *   - the caller (PyTorch) owns every buffer: inputs, outputs and workspaces.  The library never
 *     allocates, frees or retains device memory;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); no internal
 *     synchronisation, no host read-back;
 *   - return value 0 = success, otherwise one of IGCN_ERR_*; igcn_last_error() gives the text.
 *     Nothing throws, nothing calls exit().  There is no CPU fallback: an unsupported shape is an error;
 *   - outputs are fully overwritten unless the parameter comment says "accumulates".
 *
 * Each entry point names the reference interface it replaces (file:line in Houliang-Zhou/IG-GCN).
 */
#ifndef IGCN_B200_H_
#define IGCN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGCN_OK 0
#define IGCN_ERR_BAD_ARG 1     /* null pointer, negative size, misaligned buffer             */
#define IGCN_ERR_UNSUPPORTED 2 /* shape outside what the sm_100a kernels are built for        */
#define IGCN_ERR_LAUNCH 3      /* cudaGetLastError() != cudaSuccess after the launch          */

const char* igcn_last_error(void);
int igcn_version(void);
/* number of SMs of the current device (used by the host side to size partial-sum workspaces) */
int igcn_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Collation: replaces Batch.from_data_list (batch.py:24-123) as driven by DataLoader.collate
 * (dataloader.py:26-29), plus the per-call index juggling inside PyG's gcn_norm / propagate.
 *
 * in : graph_ptr (B+1) i64  prefix sums of edges per graph of this batch
 *      loc_src/loc_dst (E) i32  LOCAL node ids (0..R-1) of every edge, graph after graph, in the
 *                               reference's order (row-major COO, util_gdc.py:84-86)
 *      w (E) f32  edge_attr in the same order
 * out: edge_index (2,E) i64  global ids = local + g*R            -- bit exact with batch.py:54-55
 *      batch (B*R) i64       graph id per node                   -- bit exact with batch.py:96-99
 *      rowptr_t (B*R+1) i32, csr_src (E) i32 (global ids), csr_perm (E) i32 (original edge id of
 *        every CSR slot), csr_w (E) f32: in-edges grouped by TARGET, stable in original edge order
 *      rowptr_s (B*R+1) i32, csc_pos (E) i32: out-edges grouped by SOURCE (stable); csc_pos is the
 *        CSR slot of that edge (the transposed operator of the backward pass)
 * One CTA per graph, counting sort in shared memory, no global atomics.  max_eg = max edges of
 * any graph in the batch (host knows graph_ptr).
 */
int igcn_collate_csr(const int64_t* graph_ptr, const int32_t* loc_src, const int32_t* loc_dst, const float* w,
                     int64_t B, int64_t R, int64_t E, int64_t max_eg,
                     int64_t* edge_index, int64_t* batch,
                     int32_t* rowptr_t, int32_t* csr_src, int32_t* csr_perm, float* csr_w,
                     int32_t* rowptr_s, int32_t* csc_pos, void* stream);

/* Same CSR products from an already-collated global edge_index (2,E) i64 whose graphs each own R
 * consecutive node ids (what a reference Batch moved to the device looks like). */
int igcn_csr_from_edge_index(const int64_t* edge_index, const float* w, int64_t B, int64_t R, int64_t E,
                             int64_t max_eg, int32_t* graph_eptr /* (B+1) i32 out */,
                             int32_t* rowptr_t, int32_t* csr_src, int32_t* csr_perm, float* csr_w,
                             int32_t* rowptr_s, int32_t* csc_pos, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused SGCN encoder: replaces cal_probability (kernel/sgcn_img_snp.py:133-151), L x PyG GCNConv
 * (gcn_norm + Linear + propagate + bias; call sites kernel/sgcn_img_snp.py:218-221, kernel/sgcn.py:363-366),
 * relu, torch.cat and to_dense_batch (kernel/sgcn_img_snp.py:223-228) in ONE kernel.
 *
 *   x (B*R,F0) f32; CSR by target from igcn_collate_csr; prob (R,F0) / prob_bias (2*F0) f32 or NULL
 *   for the plain (isExplain=False) pass; wb = packed layer parameters
 *   [W_1 (H,F0) | b_1 (H) | W_2 (H,H) | b_2 (H) | ...] f32.
 *   out (B,R,L*H) f32: relu(conv_l) in concat layout;  p_e (E) f32 in CSR-slot order or NULL.
 * L == 0 computes the masks only (p_e), which is what loss_probability needs.
 * relu = 1: ReLU after every layer (the SGCN encoder); relu = 0: raw GCNConv output (single-layer operator).
 */
int igcn_sgcn_encoder_fwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* csr_w,
                          const float* prob, const float* prob_bias, const float* wb,
                          int64_t B, int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg, int64_t relu,
                          float* out, float* p_e, void* stream);

/* Backward of the above (the autograd graph of cal_probability + gcn_norm + GCNConv x L + relu + cat).
 *   g_out (B,R,L*H) f32 = dLoss/d out;  g_pe (E) f32 = extra dLoss/d p_e in CSR-slot order or NULL.
 *   dx (B*R,F0) f32 out.
 *   partials (n_cta, P) f32 workspace, P = igcn_sgcn_param_count(...): per-CTA partial parameter
 *   gradients in the order [wb layout | dprob (R*F0) | dprob_bias (2*F0)]; reduced in a fixed order
 *   into grads (P) f32 by the same call (deterministic, no float atomics).
 */
int64_t igcn_sgcn_param_count(int64_t R, int64_t F0, int64_t H, int64_t L);
int64_t igcn_sgcn_bwd_ctas(int64_t B, int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg);
int igcn_sgcn_encoder_bwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* csr_w,
                          const int32_t* rowptr_s, const int32_t* csc_pos,
                          const float* prob, const float* prob_bias, const float* wb,
                          const float* out, const float* g_out, const float* g_pe,
                          int64_t B, int64_t R, int64_t F0, int64_t H, int64_t L, int64_t max_eg, int64_t relu,
                          float* dx, float* partials, int64_t n_cta, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------
 * GO-hierarchy encoder (reference: kernel/go_model.py).  The DAG is static: the host builds, once per
 * model (replacing go_model.py:42-74,161-168), a CSR by row (rowptr,col,row_of) and a CSC by column
 * (colptr,crow,cpos: row and CSR slot of every column entry) for each layer's sub-adjacency.
 *
 * igcn_go_spmm_*: SNP->GO encode (go_model.py:208-215; channels=2, values t[0],t[1]) and GO->SNP decode
 * (go_model.py:281-282; channels=1, values t_D[0]) with learnable per-nnz values:
 *     out[b,r,c] = sum_{k in row r} vals[c*nnz+k] * in[b, col[k]]
 *   in (B,n_in) f32, vals (channels,nnz) f32, out (B,n_row,channels) f32.
 *   bwd: d_in (B,n_in) or NULL, d_vals (channels,nnz) (summed over the batch in subject order).  workspace: NULL, or
 *   (n_row*channels + n_in) * B floats -- then g_out and in are transposed once and every d_vals entry is a dot product of
 *   two contiguous batch vectors (same summation order; the path for large hierarchies, nnz >= a few thousand).
 */
int igcn_go_spmm_fwd(const float* in, const int32_t* rowptr, const int32_t* col, const float* vals,
                     int64_t B, int64_t n_in, int64_t n_row, int64_t nnz, int64_t channels, float* out, void* stream);
int igcn_go_spmm_bwd(const float* g_out, const float* in, const int32_t* row_of, const int32_t* col,
                     const int32_t* colptr, const int32_t* crow, const int32_t* cpos, const float* vals,
                     int64_t B, int64_t n_in, int64_t n_row, int64_t nnz, int64_t channels,
                     float* d_in, float* d_vals, float* workspace, void* stream);

/* igcn_go_layer_*: one hierarchy layer, fused per subject.
 *   attn=1 (encoder, go_model.py:219-251): x_in = X Wa^T, x_s = X Ws^T, a_e = exp(tanh(u.[x_in_row|x_in_col])),
 *          row-normalise, aggregate, + x_s*sigmoid(v.x_s); square (m_in == m_row), self_off = 0.
 *   attn=0 (decoder, go_model.py:258-275): uniform 1/|row| weights, self term x_s[i-self_off] on rows >= self_off.
 *   then LayerNorm over the node axis per (subject, channel) with gamma/beta (m_row), ReLU, optional dropout
 *   scale mask (B,m_row) (Dropout2d drops whole nodes per subject), output rows [keep_from, m_row).
 *   x (B,m_in,din), Wa/Ws (dout,din), u (2*dout), v (dout), y (B,m_row-keep_from,dout), stats (B,2*dout) = mean|rstd.
 *   bwd: dx (B,m_in,din); grads (P) = [dWa | dWs | du | dv | dgamma | dbeta], P = igcn_go_layer_param_count;
 *   partials (n_cta,P) workspace, n_cta = igcn_go_layer_bwd_ctas(...).  Deterministic (no float atomics).
 *   Instantiated (din,dout,attn): (2,5,1) (5,5,1) (5,5,0) (5,2,0) -- the reference's f_dim=[5,5], in_f_dim=2.
 */
int64_t igcn_go_layer_param_count(int64_t din, int64_t dout, int64_t m_row);
int64_t igcn_go_layer_bwd_ctas(int64_t B, int64_t din, int64_t dout, int64_t m_in, int64_t m_row, int64_t nnz, int64_t attn);
int igcn_go_layer_fwd(const float* x, const float* Wa, const float* Ws, const float* u, const float* v,
                      const float* gamma, const float* beta, const float* mask,
                      const int32_t* rowptr, const int32_t* col, const int32_t* colptr, const int32_t* crow, const int32_t* cpos,
                      int64_t B, int64_t m_in, int64_t m_row, int64_t nnz, int64_t din, int64_t dout, int64_t attn,
                      int64_t self_off, int64_t keep_from, float* y, float* stats, void* stream);
int igcn_go_layer_bwd(const float* x, const float* Wa, const float* Ws, const float* u, const float* v,
                      const float* gamma, const float* beta, const float* mask,
                      const int32_t* rowptr, const int32_t* col, const int32_t* colptr, const int32_t* crow, const int32_t* cpos,
                      int64_t B, int64_t m_in, int64_t m_row, int64_t nnz, int64_t din, int64_t dout, int64_t attn,
                      int64_t self_off, int64_t keep_from, const float* stats, const float* g_y,
                      float* dx, float* partials, int64_t n_cta, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused GATConv(in, out, heads=1, edge_dim=1) layer (PyG 2.0.2 semantics; reference call sites kernel/sgcn.py:163-166):
 * h = x W^T; self loops removed and re-added with the MEAN incoming edge attribute; LeakyReLU(negative_slope) edge
 * scores a_src.h_s + a_dst.h_t + ea*<lin_edge, att_edge>; per-target softmax; aggregate; + bias.  Graphs own R
 * consecutive nodes (CSR from igcn_collate_csr); edge_attr (E) is given per CSR slot.
 *   x (B*R,Fin), W (H,Fin), att_src/att_dst/lin_edge/att_edge/bias (H), out (B*R,H).
 *   bwd: dx (B*R,Fin), d_edge_attr (E) per CSR slot, grads (P) = [dW | datt_src | datt_dst | dlin_edge | datt_edge | dbias],
 *   P = igcn_gat_param_count; partials (n_cta,P) workspace, n_cta = igcn_gat_bwd_ctas(...).  Deterministic.
 */
int64_t igcn_gat_param_count(int64_t Fin, int64_t H);
int64_t igcn_gat_bwd_ctas(int64_t B, int64_t R, int64_t Fin, int64_t H, int64_t max_eg);
int igcn_gat_layer_fwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* edge_attr,
                       const float* W, const float* att_src, const float* att_dst, const float* lin_edge, const float* att_edge,
                       const float* bias, int64_t B, int64_t R, int64_t Fin, int64_t H, int64_t max_eg, double negative_slope,
                       float* out, void* stream);
int igcn_gat_layer_bwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const float* edge_attr,
                       const int32_t* rowptr_s, const int32_t* csc_pos,
                       const float* W, const float* att_src, const float* att_dst, const float* lin_edge, const float* att_edge,
                       const float* bias, const float* g_out, int64_t B, int64_t R, int64_t Fin, int64_t H, int64_t max_eg,
                       double negative_slope, float* dx, float* d_edge_attr, float* partials, int64_t n_cta, float* grads,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused cross attention (replaces nn.MultiheadAttention(E, heads, batch_first=True)(q, kv, kv)[0] + relu at
 * kernel/sgcn_img_snp.py:46,239-241): in-projections, per-head scaled scores, row softmax, P.V, out-projection, ReLU.
 *   q_in (B,R,E), kv_in (B,M,E), in_proj_weight (3E,E) = [Wq;Wk;Wv], in_proj_bias (3E), out_proj_weight (E,E),
 *   out_proj_bias (E); out (B,R,E).  relu=1 fuses the ReLU that follows the attention in the reference.
 *   bwd: `out` = forward output (ReLU mask); d_q_in (B,R,E), d_kv_in (B,M,E); grads (P) =
 *   [d in_proj_weight | d in_proj_bias | d out_proj_weight | d out_proj_bias], P = igcn_cross_attn_param_count(E);
 *   partials (n_cta,P) workspace, n_cta = igcn_cross_attn_bwd_ctas(...).  Deterministic.
 */
int64_t igcn_cross_attn_param_count(int64_t E);
/* 1 when relu = 2 is available for this shape: out = (q_in + relu(attn)) / 2, the fusion average of kernel/sgcn_img_snp.py
 * (out_z = (img_out + out_cross) / 2) folded into the attention epilogue; the backward then takes that tensor as `out` and
 * returns d q_in including the direct half. */
int64_t igcn_cross_attn_fused_average(int64_t R, int64_t M, int64_t E, int64_t heads);
int64_t igcn_cross_attn_bwd_ctas(int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads);
int igcn_cross_attn_fwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                        const float* out_proj_weight, const float* out_proj_bias,
                        int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu, float* out, void* stream);
int igcn_cross_attn_bwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                        const float* out_proj_weight, const float* out_proj_bias, const float* out, const float* g_out,
                        int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu,
                        float* d_q_in, float* d_kv_in, float* partials, int64_t n_cta, float* grads, void* stream);

/* Table-driven variant of the same operator (csrc/cross_attn_mma2.cuh), used when igcn_cross_attn_v2_supported(...) = 1 (the
 * reference's shape: E = 32, 2 heads, M <= 32 tokens, R <= 288): the forward first writes a per-graph record
 *   tab (B, igcn_cross_attn_v2_tab_floats(M, heads)) = K' | V' | c | K | V        (the query-side projections folded into key tables)
 * which the forward row kernel and the backward read (the caller keeps it with `out` for the backward).  The backward is a row kernel
 * over (graph, row chunk) items that writes per-item gradient records into `work` (igcn_cross_attn_v2_work_floats(B,R,M,heads)
 * floats) and a chain kernel that turns them into d_kv_in and the parameter gradients: partials (n_cta, P) with
 * n_cta = igcn_cross_attn_v2_bwd_ctas(B).  Same results as igcn_cross_attn_fwd/bwd to rounding; relu = 0, 1, 2 as above. */
int64_t igcn_cross_attn_v2_supported(int64_t R, int64_t M, int64_t E, int64_t heads);
int64_t igcn_cross_attn_v2_tab_floats(int64_t M, int64_t heads);
int64_t igcn_cross_attn_v2_work_floats(int64_t B, int64_t R, int64_t M, int64_t heads);
int64_t igcn_cross_attn_v2_bwd_ctas(int64_t B);
int igcn_cross_attn_v2_fwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                           const float* out_proj_weight, const float* out_proj_bias,
                           int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu, float* out, float* tab, void* stream);
int igcn_cross_attn_v2_bwd(const float* q_in, const float* kv_in, const float* in_proj_weight, const float* in_proj_bias,
                           const float* out_proj_weight, const float* out_proj_bias, const float* out, const float* g_out,
                           const float* tab, int64_t B, int64_t R, int64_t M, int64_t E, int64_t heads, int64_t relu,
                           float* d_q_in, float* d_kv_in, float* work, float* partials, int64_t n_cta, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fusion heads: out = act([X0 | X1 | X2] W^T + b) without materialising the concatenation (reference:
 * kernel/sgcn_img_snp.py:287-301 -- cat(out_z, latent) -> lin1 -> relu ; cat(out_lin, img_feat) -> lin1_regr -> relu).
 *   X_i (M, host_widths[i]) f32 with row stride host_strides[i] (width 0 = absent); W (N,K) row-major, K = sum of widths;
 *   bias (N); relu = 1 applies ReLU.  Split-K over S = igcn_cat_linear_splits(M,N,K) chunks: partials (S,M,N) workspace,
 *   reduced in a fixed order (deterministic).  host_* arrays are 3-element HOST arrays.
 *   bwd: g_out = dLoss/d out, `out` = the forward output (ReLU mask); dW (N,K), db (N), dx_i (M, width_i) with row
 *   stride host_dstrides[i] (NULL = not needed); dW = db = NULL skips the weight gradient; outputs fully overwritten.
 */
int64_t igcn_cat_linear_splits(int64_t M, int64_t N, int64_t K);
int igcn_cat_linear_fwd(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                        const float* W, const float* bias, int64_t M, int64_t N, int64_t K, int64_t relu,
                        float* partials, int64_t S, float* out, void* stream);
int igcn_cat_linear_bwd(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                        const float* W, const float* out, const float* g_out, int64_t M, int64_t N, int64_t K, int64_t relu,
                        float* dx0, float* dx1, float* dx2, const int64_t* host_dstrides, float* dW, float* db, void* stream);

/* The same operator on warp-level tensor cores (csrc/catlin_mma.cu; mma.sync TF32 split in three, fp32 accumulate) for the reference's
 * batch sizes: igcn_catlin_mma_supported(M,N,K) = 1 for M <= 2048, N <= 64 (a multiple of 16).  One launch per product, operands read
 * in place.  host_rows[i] = number of rows of source i: a source with fewer rows than M is read cyclically (row m -> m % rows_i; the
 * stacked plain / explain passes share img_feat).  bwd_dx writes dx_i (M, width_i) for every non-NULL dx_i (a cyclic source gets one
 * gradient row per product row; the caller adds the repetitions); bwd_dw writes dW (N,K) and db (N, may be NULL).  The two backward
 * products are independent and may run on different streams.  Deterministic. */
int64_t igcn_catlin_mma_supported(int64_t M, int64_t N, int64_t K);
int igcn_catlin_mma_fwd(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                        const int64_t* host_rows, const float* W, const float* bias, int64_t M, int64_t N, int64_t K, int64_t relu,
                        float* out, void* stream);
int igcn_catlin_mma_bwd_dx(const int64_t* host_widths, const float* W, const float* out, const float* g_out, int64_t M, int64_t N, int64_t K,
                           int64_t relu, float* dx0, float* dx1, float* dx2, const int64_t* host_dx_strides, void* stream);
int igcn_catlin_mma_bwd_dw(const float* x0, const float* x1, const float* x2, const int64_t* host_widths, const int64_t* host_strides,
                           const int64_t* host_rows, const float* out, const float* g_out, int64_t M, int64_t N, int64_t K, int64_t relu,
                           float* dW, float* db, void* stream);

/* ------------------------------------------------------------------------------------------
 * All dropout masks of one forward pass in one launch (the reference draws nine per pass: Dropout2d(0.4) x4 and
 * Dropout(0.5) x3 in kernel/go_model.py:104,113,127,135,142; F.dropout 0.5 / 0.3 at kernel/sgcn_img_snp.py:290,300).
 *   out: flat f32 buffer; segment i covers [host_seg_end[i-1], host_seg_end[i]) and is filled with 0 or 1/keep_i,
 *   keep_i = host_seg_keep[i].  Philox4x32-10, seed + a DEVICE call counter (u64, incremented by the call), so a
 *   captured CUDA graph draws fresh masks on every replay.  host_* are host arrays of nseg (<= 32) entries.
 */
int igcn_dropout_masks(float* out, const int64_t* host_seg_end, const float* host_seg_keep, int64_t nseg, uint64_t seed,
                       unsigned long long* counter, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused flat-buffer Adam (replaces torch.optim.Adam.step as called at kernel/train_eval_sgcn_img_snps.py:547;
 * lr 1e-3, betas (0.9,0.999), eps 1e-8, weight_decay 0 -- :108).  One launch for the whole model.
 *   params / grads / exp_avg / exp_avg_sq: (n) f32, 16-byte aligned; updates params and both moments in place.
 *   step (1) f32 device scalar = the 1-based update count (the caller increments it before the call);
 *   lr (1) f32 device scalar.  grad_scale multiplies the gradient on the fly (1/world after a SUM all-reduce).
 */
int igcn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* step, const float* lr,
                   double beta1, double beta2, double eps, double grad_scale, int64_t n, void* stream);

/* Gathers `count` separately allocated gradient tensors into their slots of the flat gradient buffer igcn_adam_step /
 * igcn_dp_allreduce_adam read (the reference's optimizer walks p.grad of every parameter, torch.optim.Adam.step called at
 * kernel/train_eval_sgcn_img_snps.py:547; here the per-parameter walk is one copy launch per 96 tensors).
 *   host_src_ptrs / host_offsets / host_sizes: HOST arrays of `count` entries -- device address of a contiguous f32 gradient
 *   (0: the slot is zero-filled, a parameter that received no gradient), element offset of its slot in `flat` (multiple of 4)
 *   and element count.  The table is passed to the kernel by value: the host arrays may be freed on return and the launch can be
 *   captured in a CUDA graph (the captured addresses are the graph's own).  step_counter (device f32 scalar, may be NULL) is
 *   incremented by 1: the optimizer's step count that igcn_adam_step / igcn_dp_allreduce_adam read next, so that the increment is
 *   not a launch of its own between the gather and the update.  igcn_gather_flat_launches = kernels launched. */
int igcn_gather_flat(const int64_t* host_src_ptrs, const int64_t* host_offsets, const int64_t* host_sizes, int64_t count,
                     float* flat, int64_t flat_n, float* step_counter, void* stream);
int64_t igcn_gather_flat_launches(int64_t count);

/* ------------------------------------------------------------------------------------------
 * Read-out heads of the GO network: mask * relu(BatchNorm1d(z)) in TRAINING mode as one launch (forward) and one
 * launch (backward).  Replaces nn.BatchNorm1d + nn.ReLU + nn.Dropout as chained in kernel/go_model.py:117-146
 * (conc_for_attention[1:3], B, B_D, latent[1:4], latent[5:7]).
 *   z (N, C, L) f32 contiguous (L = 1 for a 2-D input); statistics per channel c over the N/groups * L values of each
 *   of `groups` consecutive slices of the batch, visited in order: running_mean / running_var (C, updated in place,
 *   may be NULL) and num_batches_tracked (i64 device scalar, may be NULL) receive exactly the updates of `groups`
 *   successive module calls.  gamma / beta / mask may be NULL; mask has z's shape (multiplicative, 0 or 1/keep).
 *   stats (groups, C, 2) f32 receives (mean, rstd) for the backward.  relu != 0 applies max(.,0) before the mask.
 *   bwd: dz (N,C,L), dgamma (C), dbeta (C) (either may be NULL) are fully overwritten.
 */
int igcn_bn_act_fwd(const float* z, const float* gamma, const float* beta, const float* mask, int64_t N, int64_t C, int64_t L,
                    int64_t groups, double eps, double momentum, int64_t relu, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* y, float* stats, void* stream);
int igcn_bn_act_bwd(const float* z, const float* gamma, const float* beta, const float* mask, const float* stats, const float* g_y,
                    int64_t N, int64_t C, int64_t L, int64_t groups, int64_t relu, float* dz, float* dgamma, float* dbeta,
                    void* stream);
/* Eval-mode form of the same heads (model.eval(): running statistics, no dropout) -- the inference path that eval_acc / eval_loss /
 * eval_scores of kernel/train_eval_sgcn_img_snps.py:551-671 run five times per epoch:
 *   y = act((z - running_mean) / sqrt(running_var + eps) * gamma + beta), z (N, C, L);  with g_y != NULL, y receives d z instead. */
int igcn_bn_eval_act(const float* z, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     int64_t N, int64_t C, int64_t L, double eps, int64_t relu, const float* g_y, float* y, void* stream);

/* The per-node read-out Linear (bias-free, K <= 8 inputs, L = 1 or 32 outputs) fused with the training-mode BatchNorm1d + ReLU +
 * dropout scale that follows it (kernel/go_model.py:117-131: conc_for_attention, conc -> B, conc_D -> B_D):
 *   y = mask * relu(BatchNorm_c(x W^T)),  x (N, C, K), W (L, K), y / mask (N, C, L), statistics per channel c over (n, l) and per
 *   stacked pass (groups = 2 only).  z = x W^T is never materialised; the backward recomputes it.  igcn_lin_bn_act_supported(...) = 1
 *   when the shape fits (otherwise use igcn_skinny_linear_* + igcn_bn_act_*).  bwd: dx (N, C, K), dW (L, K) via partials
 *   (igcn_lin_bn_act_partial_rows(N, C, L, K), L*K), dgamma / dbeta (C).  Deterministic.  For L = 32 a channel's rows are split over
 *   a thread-block cluster of up to 8 CTAs that combine their statistics through distributed shared memory. */
int64_t igcn_lin_bn_act_supported(int64_t N, int64_t C, int64_t L, int64_t K, int64_t groups);
int64_t igcn_lin_bn_act_partial_rows(int64_t N, int64_t C, int64_t L, int64_t K);
int igcn_lin_bn_act_fwd(const float* x, const float* W, const float* gamma, const float* beta, const float* mask, int64_t N, int64_t C,
                        int64_t L, int64_t K, int64_t groups, double eps, double momentum, int64_t relu, float* running_mean,
                        float* running_var, long long* num_batches_tracked, float* y, float* stats, void* stream);
int igcn_lin_bn_act_bwd(const float* x, const float* W, const float* gamma, const float* beta, const float* mask, const float* stats,
                        const float* g_y, int64_t N, int64_t C, int64_t L, int64_t K, int64_t groups, int64_t relu, float* dx,
                        float* partials, float* dW, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------
 * loss_probability (kernel/sgcn_img_snp.py:153-181): for p in {sigmoid(prob) (n_prob), p_e (n_e), sigmoid(snps_prob)
 * (n_snps)}:  c_l1 * mean|p| + c_ent * mean(-(p log(p+eps) + (1-p) log(1-p+eps))), summed.
 *   host_coef = {lamda_x_l1, lamda_e_l1, lamda_x_ent, lamda_e_ent} (sgcn_hyperparameters.py:18-21; the SNP mask uses
 *   the x coefficients, :176-179).  prob / snps_prob are the RAW parameters (sigmoid applied inside), p_e is a
 *   probability.  partials: workspace of n_partials = igcn_reduce_blocks(max n) floats; loss (1).  Fixed summation
 *   order (two launches, no atomics).  bwd: g_loss (1) device scalar; d_* may be NULL.
 */
int64_t igcn_reduce_blocks(int64_t n);
int igcn_mask_loss_fwd(const float* prob, int64_t n_prob, const float* p_e, int64_t n_e, const float* snps_prob, int64_t n_snps,
                       const float* host_coef, double eps, float* partials, int64_t n_partials, float* loss, void* stream);
int igcn_mask_loss_bwd(const float* prob, int64_t n_prob, const float* p_e, int64_t n_e, const float* snps_prob, int64_t n_snps,
                       const float* host_coef, double eps, const float* g_loss, float* d_prob, float* d_pe, float* d_snps_prob,
                       void* stream);

/* out[0] = scale * <a, b> over n f32 values in a fixed order (partials: igcn_reduce_blocks(n) floats); and
 * out[i] = a[i] * s[0] * scale with s a device scalar.  Together with one product T = Lsym S they are the consistency
 * loss tr(s^T (D - W) s) / B^2 of kernel/sgcn_img_snp.py:183-196 and its gradient 2 T g / B^2. */
int igcn_dot(const float* a, const float* b, int64_t n, double scale, float* partials, int64_t n_partials, float* out, void* stream);
int igcn_scale_by_scalar(const float* a, const float* s, double scale, int64_t n, float* out, void* stream);

/* out[i] = (a[i] + b[i]) + c[i] over n f32 values (c may be NULL; all 16-byte aligned): the gradient of a tensor read by up to three
 * consumers, summed when the last one has arrived.  The reference leaves this to autograd (one add per extra consumer, e.g. the
 * GO encoder output read by conc_for_attention, conc and the decoder, kernel/go_model.py:232-262); see ops.fan_out. */
int igcn_sum3(const float* a, const float* b, const float* c, int64_t n, float* out, void* stream);

/* The consistency loss without its cancellation (csrc/laplacian.cu): with L = D - W and L 1 = 0,
 *   <s, L s> = <s', T>,  T = L s = d .* s' - W s',  s' = s - column means,  d = row sums of W.
 * igcn_rbf_similarity: W (B,B) = exp(-gamma ||t_i - t_j||^2) (util/image_cluster.py:15-31 via kernel/sgcn_img_snp.py:188) and d (B).
 * igcn_col_mean: m (groups,D) = column means of each group of B rows of s (groups*B, D).
 * igcn_laplacian_finish: T = d .* (s - m) - U (U = W (s - m), e.g. from igcn_tc_gemm; NULL = 0, the all-ones similarity) and
 *   out[0] = scale * <s - m, T>; d NULL = the constant d_const; partials: n_partials = igcn_reduce_blocks(groups*B*D) floats. */
int igcn_rbf_similarity(const float* t, int64_t B, int64_t R, double gamma, float* W, float* d, void* stream);
int igcn_col_mean(const float* s, int64_t B, int64_t D, int64_t groups, float* m, void* stream);
int igcn_laplacian_finish(const float* s, const float* m, const float* d, const float* U, int64_t B, int64_t D, int64_t groups,
                          double d_const, double scale, float* T, float* partials, int64_t n_partials, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense products on the tcgen05 tensor cores, fp32-accurate ("3xTF32", accumulators in TMEM, operands by TMA).
 * Used for the fusion heads lin1 / lin1_regr (kernel/sgcn_img_snp.py:286-305) forward and backward, and for the
 * (B x B)(B x D) Laplacian product of consist_loss (kernel/sgcn_img_snp.py:183-196).
 *
 * igcn_tc_split: operand preparation, up to 8 jobs in one launch.  Job = 12 int64 in host memory:
 *   {src, mask, hi, lo, rows, cols, ld_src, ld_dst, row_off, col_off, transpose, sub}
 *   sub (cols f32, may be NULL): subtracted per source column before the split (column centring);
 *   src (rows, cols) f32 with row pitch ld_src (NULL = the constant 1.0); mask (same geometry, may be NULL): elements
 *   whose mask value is not > 0 are taken as 0 (the ReLU mask of a backward pass); every element x is written as the
 *   pair hi = rn_tf32(x), lo = rn_tf32(x - hi) to hi/lo[(row_off + r) * ld_dst + col_off + c], or transposed to
 *   hi/lo[(row_off + c) * ld_dst + col_off + r].  This is also where cat(...) happens (column offsets).
 * igcn_tc_gemm:  C[m][n] = act(sum_k A[m][k] * B[n][k] + bias[n]) with A = a_hi + a_lo (M, K; row pitch lda) and
 *   B = b_hi + b_lo (N, K; row pitch ldb), pitches multiples of 4 floats, bases 16-byte aligned.  Column n of C goes to
 *   up to three destination segments d0 | d1 | d2 (host_dst_widths add up to N, host_dst_strides = row pitches; a NULL
 *   destination drops that segment).  S = igcn_tc_gemm_splits(M, N, K) K-splits; for S > 1 `partials` (S*M*N floats)
 *   is required and the splits are summed in a fixed order.  bias may be NULL.
 */
int igcn_tc_split(const int64_t* host_jobs, int64_t njobs, void* stream);
int64_t igcn_tc_gemm_splits(int64_t M, int64_t N, int64_t K);
int igcn_tc_gemm(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo, int64_t ldb, int64_t M,
                 int64_t N, int64_t K, const float* bias, int64_t relu, float* d0, float* d1, float* d2,
                 const int64_t* host_dst_widths, const int64_t* host_dst_strides, float* partials, int64_t S, void* stream);

/* ------------------------------------------------------------------------------------------
 * Skinny bias-free linear layers of the GO read-outs (kernel/go_model.py:117-131: Linear(5 -> dim_snps_atten), Linear(5 -> 1),
 * Linear(2 -> 1) applied to every (subject, GO term) row, and the two bias-free Linears of `latent`, :138-146): z (rows, Lout) = x (rows, Kin) W (Lout, Kin)^T with Kin <= 32,
 * Lout <= 64.  bwd: dx (rows, Kin; may be NULL) and dW (Lout, Kin); partials = n_cta * Lout * Kin floats with
 * n_cta = igcn_skinny_linear_bwd_ctas(rows); summed in CTA order (deterministic).
 */
int igcn_skinny_linear_fwd(const float* x, const float* W, int64_t rows, int64_t Kin, int64_t Lout, float* z, void* stream);
int64_t igcn_skinny_linear_bwd_ctas(int64_t rows);
int igcn_skinny_linear_bwd(const float* x, const float* W, const float* dz, int64_t rows, int64_t Kin, int64_t Lout, float* dx,
                           float* partials, int64_t n_cta, float* dW, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tail of the training step (kernel/sgcn_img_snp.py:147-148,290-291,300-301; kernel/train_eval_sgcn_img_snps.py:524-544).
 *
 * igcn_snp_mask_pair_*: the SNP input of the two stacked passes, out (2B, S): rows [0,B) = snps, rows [B,2B) =
 *   snps * sigmoid(snps_prob) (cal_probability, :147-148).  bwd: d_snps_prob (S) from the gradient of out.
 * igcn_heads_*: logp (rows, C1) = log_softmax(lin2(h1 * m1)), reg (rows, C2) = lin2_regr(h2 * m2) with h (rows, K <= 64),
 *   dropout masks m (may be NULL), W (C, K), b (C), C <= 8.  bwd: g_logp / g_reg may be NULL (zero); dh1, dh2 (rows, K; may be NULL);
 *   grads = [dW1 | db1 | dW2 | db2]; partials = igcn_heads_bwd_ctas(rows) * len(grads) floats (summed in CTA order).
 * igcn_step_loss_*: c_reg * mean((reg - target)^2) + c_rec * sum((xhat - snps)^2) + c_prob * loss_prob + c_clu * quad, where
 *   reg is (2, n_reg) and xhat (2, n_rec) -- the plain and explain passes -- with target (n_reg) / snps (n_rec) shared by both;
 *   loss_prob / quad are device scalars (may be NULL).  One launch each way; d_loss_prob / d_quad receive the scalar gradients.
 */
int igcn_snp_mask_pair_fwd(const float* snps, const float* snps_prob, int64_t B, int64_t S, float* out, void* stream);
int igcn_snp_mask_pair_bwd(const float* snps, const float* snps_prob, const float* g_out, int64_t B, int64_t S, float* d_snps_prob,
                           void* stream);
int64_t igcn_heads_bwd_ctas(int64_t rows);
int igcn_heads_fwd(const float* h1, const float* m1, const float* h2, const float* m2, const float* W1, const float* b1,
                   const float* W2, const float* b2, int64_t rows, int64_t K, int64_t C1, int64_t C2, float* logp, float* reg,
                   void* stream);
int igcn_heads_bwd(const float* h1, const float* m1, const float* h2, const float* m2, const float* W1, const float* b1,
                   const float* W2, const float* b2, const float* logp, const float* g_logp, const float* g_reg, int64_t rows,
                   int64_t K, int64_t C1, int64_t C2, float* dh1, float* dh2, float* partials, int64_t n_cta, float* grads,
                   void* stream);
int igcn_step_loss_fwd(const float* reg, const float* target, int64_t n_reg, const float* xhat, const float* snps, int64_t n_rec,
                       const float* loss_prob, const float* quad, double c_reg, double c_rec, double c_prob, double c_clu,
                       float* out, void* stream);
int igcn_step_loss_bwd(const float* reg, const float* target, int64_t n_reg, const float* xhat, const float* snps, int64_t n_rec,
                       const float* g_loss, double c_reg, double c_rec, double c_prob, double c_clu, float* d_reg, float* d_xhat,
                       float* d_loss_prob, float* d_quad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel step: gradient all-reduce + Adam as ONE kernel over NVLink peer memory (replaces ncclAllReduce on the flat
 * gradient buffer followed by igcn_adam_step; reference: one optimizer.step() per batch, kernel/train_eval_sgcn_img_snps.py:547,
 * made data parallel over graphs).
 *   host_grad_ptrs / host_signal_ptrs: HOST arrays of `world` device pointers -- the flat gradient buffer (n f32, 16-byte aligned)
 *   and the signal pad (signal_pad_bytes, zero-initialised once) of every rank, all peer-mapped into this process (e.g. from
 *   torch.distributed._symmetric_memory: buffer_ptrs / signal_pad_ptrs).  Every rank launches the same call on its own stream;
 *   the kernel exchanges flags with all peers before and after reading their gradients.  Waits are bounded by `timeout_ms` per
 *   phase (choose minutes: ranks may be seconds apart); on expiry the kernel stores ((phase << 8) | (peer + 1)) into `error_flag`
 *   (a device or mapped-host int the caller polls; the step's result is then invalid) or, when error_flag is NULL, traps.  All
 *   ranks must issue their steps in lock-step and should meet at a host barrier before the first fused launch.  The kernel
 *   sums the `world` gradients in rank order -- each chunk once, by its owner rank, which stores the sum back into EVERY
 *   rank's gradient buffer (the buffers hold the un-scaled sum afterwards; bit-identical replicas; n bytes per rank each way over
 *   NVLink whatever the world size) -- scales by 1/world and updates params / exp_avg / exp_avg_sq exactly as igcn_adam_step does.  n must be a multiple of 4.  igcn_dp_adam_blocks = CTAs used (0: signal pad too small).
 */
int64_t igcn_dp_adam_blocks(int64_t n, int64_t world, int64_t signal_pad_bytes);
int igcn_dp_allreduce_adam(const int64_t* host_grad_ptrs, const int64_t* host_signal_ptrs, int64_t rank, int64_t world,
                           int64_t signal_pad_bytes, float* params, float* exp_avg, float* exp_avg_sq, const float* step,
                           const float* lr, double beta1, double beta2, double eps, int64_t n, int64_t timeout_ms, int* error_flag,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * GCNConv for ARBITRARY graphs: the operator the reference's model files call,
 * torch_geometric.nn.GCNConv(in, out)(x, edge_index, edge_weight) (kernel/sgcn_img_snp.py:218-221, kernel/sgcn.py:281-284; PyG 2.0.2:
 * add_remaining_self_loops with fill 1, D^-1/2 A D^-1/2, X W^T, scatter-add over targets, + bias), differentiable with respect
 * to x, edge_weight, weight and bias.  One graph of N nodes in global memory (no size limit, no equal-size requirement).
 *   igcn_graph_csr: int64 COO (2,E) -> in-edges grouped by target (rowptr_t, csr_src, csr_perm = original edge id per slot) and
 *     out-edges grouped by source (rowptr_s, csc_pos = CSR slot of the q-th out-edge); stable order.  work: igcn_graph_csr_work_ints
 *     int32s; its LAST entry is set to 1 when an endpoint lies outside [0, N).
 *   igcn_gcn_conv_fwd: edge_weight (E, original edge order) or NULL (= ones); weight (O,C); bias (O) or NULL; out (N,O);
 *     saved: igcn_gcn_conv_saved_floats floats the backward needs.
 *   igcn_gcn_conv_bwd: dx (N,C) or NULL; d_edge_weight (E, original order) or NULL; grads (O*C + O) = [d weight | d bias];
 *     work: igcn_gcn_conv_bwd_work_floats floats; partials (n_cta, O*C+O), n_cta = igcn_gcn_conv_bwd_ctas(N).  Deterministic.
 */
int64_t igcn_graph_csr_work_ints(int64_t N, int64_t E);
int igcn_graph_csr(const int64_t* edge_index, int64_t N, int64_t E, int32_t* rowptr_t, int32_t* csr_src, int32_t* csr_perm,
                   int32_t* rowptr_s, int32_t* csc_pos, int32_t* work, void* stream);
int64_t igcn_gcn_conv_saved_floats(int64_t N, int64_t E, int64_t O);
int igcn_gcn_conv_fwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const int32_t* csr_perm,
                      const float* edge_weight, const float* weight, const float* bias, int64_t N, int64_t E, int64_t C, int64_t O,
                      float* saved, float* out, void* stream);
int64_t igcn_gcn_conv_bwd_ctas(int64_t N);
int64_t igcn_gcn_conv_bwd_work_floats(int64_t N, int64_t E, int64_t O);
int igcn_gcn_conv_bwd(const float* x, const int32_t* rowptr_t, const int32_t* csr_src, const int32_t* csr_perm,
                      const int32_t* rowptr_s, const int32_t* csc_pos, const float* weight, const float* saved, const float* g_out,
                      int64_t N, int64_t E, int64_t C, int64_t O, float* work, float* dx, float* d_edge_weight, float* partials,
                      int64_t n_cta, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------
 * On-device preprocessing: the sparsification step of graph diffusion convolution (util_gdc.py:25-31 get_top_k_matrix + :84-101
 * dense -> COO; applied per subject as a pre_transform, sgcn_data.py:332-338).  ppr (B,R,R) f64 = the dense PPR matrices
 * alpha (I - (1-alpha) D^-1/2 A D^-1/2)^-1 (util_gdc.py:7-14).  Per subject and column: keep the k largest entries, divide by their
 * sum, cast to f32, emit the non-zeros in row-major order: edge_src / edge_dst (B*R*k) i32 LOCAL ids, edge_attr (B*R*k) f32.
 * Bit-identical edge lists to the numpy restatement for the same input. */
int igcn_gdc_topk_emit(const double* ppr, int64_t B, int64_t R, int64_t k, int32_t* edge_src, int32_t* edge_dst, float* edge_attr,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IGCN_B200_H_ */
