/* whisprrec_b200.h -- C-ABI of libwhisprrec_b200.so (sm_100a).
 *
 * The reference (HeyWeCome/WhisprRec) has no FFI: its hot path is a chain of PyTorch calls made from
 * Python classes.  Each entry point below replaces the chain of torch ops at the cited reference lines
 * (paths relative to the reference checkout, src/...).  INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add at each site.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `host_`;
 *   - tables are row-major fp32 [rows, D]; ids are int64 (what `collate_batch` hands to the model,
 *     models/BaseModel.py:96-127); CSR column / history indices are int32, row pointers int64;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and never syncs the
 *     host (wr_status is the one exception and says so);
 *   - `ws` is a caller-owned device scratch block of wr_workspace_bytes() bytes, zeroed once with
 *     wr_workspace_init; one workspace per stream;
 *   - return value: 0 ok, <0 argument error (WR_E_*), >0 a cudaError_t from the launch;
 *   - D must be a multiple of 4 (128-bit row accesses); other sizes return WR_E_DIM.
 *   - there is no CPU path: on a machine without an sm_100 device every compute call returns the CUDA error.
 */
#ifndef WHISPRREC_B200_H
#define WHISPRREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WR_VERSION 100

#define WR_OK 0
#define WR_E_NULL (-1)   /* a required pointer is NULL */
#define WR_E_SIZE (-2)   /* a negative / inconsistent size */
#define WR_E_DIM (-3)    /* embedding size not supported (D % 4 != 0 or D > 512) */
#define WR_E_TOPK (-4)   /* k outside [1, 32] */
#define WR_E_ALIGN (-5)  /* a table pointer is not 16-byte aligned */
#define WR_E_PRECISION (-6) /* unknown / unavailable scoring precision */

/* bits of the device-side status word (wr_status) */
#define WR_STATUS_INDEX_OUT_OF_RANGE 1u /* an id outside its table: the row was skipped (torch raises IndexError) */

int wr_version(void);
const char *wr_error_string(int code);

size_t wr_workspace_bytes(void);
int wr_workspace_init(void *ws, void *stream);
/* Copies the status word to *host_status and clears it.  SYNCHRONISES the stream. */
int wr_status(void *ws, uint32_t *host_status, void *stream);

/* ---- training ---------------------------------------------------------------------------------------
 * wr_bpr_fwd_bwd: models/general/BPRMF.py:69-80 (gather, row dots), utils/loss.py:33-39 (BPRLoss) and the
 * autograd backward BaseRunner.py:198 triggers, in one launch.
 *   loss = -mean_b log(gamma + sigmoid(<U[u_b],I[p_b]> - <U[u_b],I[n_b]>))
 *   c_b  = grad_scale * dloss/ds+_b ;  gU[u_b] += c_b (I[p_b]-I[n_b]);  gI[p_b] += c_b U[u_b];  gI[n_b] -= c_b U[u_b]
 * gU/gI are ACCUMULATED into (they must hold zeros or an earlier partial gradient; wr_adam_l2_sweep leaves
 * them zeroed).  U/I may be the ego tables (BPRMF) or the pooled tables (LightGCN.py:155-162, grad_scale =
 * 1/(L+1) folds the mean-pool adjoint in).  loss_out[0] is overwritten (accumulate_loss = 0) or added to.
 */
int wr_bpr_fwd_bwd(const float *U, const float *I, const int64_t *user, const int64_t *pos, const int64_t *neg,
                   int64_t B, int D, int64_t n_users, int64_t n_items, float gamma, float grad_scale,
                   float *gU, float *gI, float *loss_out, int accumulate_loss, void *ws, void *stream);

/* wr_embloss_fwd_bwd: utils/loss.py:83-98 (EmbLoss, require_pow=False) as called at LightGCN.py:165-175.
 *   reg = (||U0[user]||_F + ||I0[pos]||_F + ||I0[neg]||_F) / B ;  loss_out[0] += reg_weight * reg
 *   gU0[u_b] += reg_weight/B * U0[u_b]/||U0[user]||_F   (and likewise for pos / neg rows)
 * Two launches: the norms are batch-global and must exist before the gradient.
 */
int wr_embloss_fwd_bwd(const float *U0, const float *I0, const int64_t *user, const int64_t *pos,
                       const int64_t *neg, int64_t B, int D, int64_t n_users, int64_t n_items, float reg_weight,
                       float *gU0, float *gI0, float *loss_out, void *ws, void *stream);

/* wr_adam_l2_sweep: torch.optim.Adam(weight_decay=l2).step() as built at helpers/BaseRunner.py:120-124 and
 * called at :199 (torch/optim/adam.py _single_tensor_adam), plus the zero_grad of :196, fused in one sweep
 * over ALL rows (dense Adam: every row moves every step).
 *   g += l2 p;  m += (1-b1)(g-m);  v = b2 v + (1-b2) g g;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps);  g = 0
 * step_size = lr/(1-b1^t) and bc2_sqrt = sqrt(1-b2^t) are evaluated in double by the caller as torch does;
 * the betas are doubles because torch forms 1-beta in double before rounding it for the fp32 kernels.
 * If dev_scalars != NULL, {step_size, bc2_sqrt} are read from that device array instead (lets a captured
 * CUDA graph be replayed for later steps).
 */
int wr_adam_l2_sweep(float *P, float *M, float *V, float *G, int64_t n_elems, float l2, double beta1, double beta2,
                     float eps, float step_size, float bc2_sqrt, const float *dev_scalars, void *stream);

/* wr_bprmf_step: one whole iteration of BaseRunner.fit for BPRMF (BaseRunner.py:196-199: zero_grad, predict,
 * backward, Adam.step) on the fused table P = [U; I] ([n_users + n_items, D], M / V / G alike).  Tables of up to
 * 8 Mi elements take ONE cooperative launch (parameter loads issued first, BPR forward+backward, grid barrier,
 * Adam+L2 with the gradient re-zeroed); larger ones are wr_bpr_fwd_bwd followed by wr_adam_l2_sweep.  Same
 * arithmetic either way.  loss_out[0] is overwritten.
 */
int wr_bprmf_step(float *P, float *M, float *V, float *G, const int64_t *user, const int64_t *pos,
                  const int64_t *neg, int64_t B, int D, int64_t n_users, int64_t n_items, float gamma, float l2,
                  double beta1, double beta2, float eps, float step_size, float bc2_sqrt, const float *dev_scalars,
                  float *loss_out, void *ws, void *stream);

/* wr_bprmf_step_host: the same iteration fed from the HOST, i.e. utils.batch_to_gpu (utils/utils.py:33-37) +
 * the step + `loss.detach().cpu()` (BaseRunner.py:200).  host_ids: pinned [3, B] int64 (user, pos, neg rows);
 * dev_ids: device staging of the same shape; host_loss: pinned float that receives the batch loss.
 * sync != 0 waits for the stream before returning (the reference syncs on every step).
 */
int wr_bprmf_step_host(const int64_t *host_ids, int64_t *dev_ids, float *host_loss, float *P, float *M, float *V,
                       float *G, int64_t B, int D, int64_t n_users, int64_t n_items, float gamma, float l2,
                       double beta1, double beta2, float eps, float step_size, float bc2_sqrt, float *loss_out,
                       void *ws, void *stream, int sync);

/* ---- LightGCN propagation ------------------------------------------------------------------------------
 * wr_csr_norm_weights: the value recipe of LightGCN.py:89-97: val[e] = fl32(fl32(dinv[row] * 1) * dinv[col]).
 * dinv = np.power(fp32(deg) + 1e-10, -0.5) comes from the caller (NumPy's fp32 pow is not correctly rounded,
 * so only the same NumPy call reproduces the reference bit for bit).
 */
int wr_csr_norm_weights(const int64_t *rowptr, const int32_t *col, const float *dinv, int64_t N, float *val,
                        void *stream);

/* wr_csr_spmm: one `torch.sparse.mm(norm_adj, E)` of LightGCN.py:139 fused with the stack/mean of :142-143.
 *   y[r]       = sum_e val[e] * X[col[e]]  (+ add[r] if add != NULL)
 *   Y[r]       = y[r]                                   if Y != NULL
 *   acc_out[r] = (acc_in[r] + y[r]) / acc_div           if acc_out != NULL
 * The backward pass of the propagation is the same call (the adjacency is symmetric): H <- A H + G'.
 * zero_add != 0 clears add[r] after reading it (recycles the pooled-gradient buffer for the next step).
 * X must not alias Y / acc_out.
 */
typedef struct wr_spmm_plan {
    /* HOST struct of DEVICE pointers: how rows with more than long_threshold non-zeros are cut into slices so
     * that no warp walks more than one slice (power-law graphs: a 10^6-edge item row would otherwise be the tail
     * of every launch).  Built once per graph by the caller (whisprrec_b200/_lib.py SpmmPlan does it in NumPy). */
    int64_t long_threshold;       /* rows with more non-zeros than this are split */
    int64_t n_chunks, n_long;     /* number of slices, number of split rows */
    const int32_t *chunk_row;     /* [n_chunks] row a slice belongs to */
    const int64_t *chunk_beg;     /* [n_chunks] first edge of the slice */
    const int32_t *chunk_len;     /* [n_chunks] edges in the slice */
    const int32_t *chunk_slot;    /* [n_chunks] index of the split row in the slot arrays */
    const int32_t *slot_chunks;   /* [n_long] slices per split row */
    int32_t *slot_arrivals;       /* [n_long] zeroed; left zeroed by every call */
    float *slot_partial;          /* [n_long, D] zeroed; left zeroed by every call */
} wr_spmm_plan;

int wr_csr_spmm(const int64_t *rowptr, const int32_t *col, const float *val, int64_t N, int D, const float *X,
                float *Y, float *add, int zero_add, const float *acc_in, float *acc_out, float acc_div,
                const wr_spmm_plan *host_plan /* nullable */, void *stream);

/* ---- full-ranking evaluation ---------------------------------------------------------------------------
 * wr_eval_rank_topk: BPRMF.py:82-91 / LightGCN.py:177-187 (S = U[user] I^T), BaseRunner.py:238 (target
 * gather), :246-255 (history mask), :72-73 (rank) without materialising S.
 *   target[r] = <U[user_r], I[pos_r]>
 *   rank[r]   = 1 + #{ j not in hist(user_r) : <U[user_r], I[j]> > target[r] }
 *   topk_idx[r,:k], topk_val[r,:k] = the k best unmasked items, best first, ties to the lower id
 *                                    (unfilled slots: idx -1, val -inf); pass NULL to skip.
 * hist_ptr/hist_idx: CSR of the sorted union train_clicked_set U residual_clicked_set per user.
 * scores_out: optional dense fp32 [R, n_items] copy of S, unmasked -- what full_predict returns; NULL on the
 *             fast path (the point of the fusion is not to write it).
 * precision 0: fp32 FMA chains over d = 0..D-1 for every score (target included), so comparisons are
 * consistent.  precision 1: bf16 operands on the tcgen05 tensor cores (TMA-fed, TMEM accumulators), fp32
 * accumulation; D in {64, 128}; ranks only (topk_idx must be NULL); needs `scratch`; looser parity (operands are
 * rounded to bf16, the target's own column is excluded explicitly).
 */
int wr_eval_rank_topk(const float *Uemb, const float *Iemb, const int64_t *user, const int64_t *pos, int64_t R,
                      int64_t n_users, int64_t n_items, int D, const int64_t *hist_ptr, const int32_t *hist_idx,
                      int k, int precision, int32_t *topk_idx, float *topk_val, int32_t *rank, float *target,
                      float *scores_out, void *scratch, void *ws, void *stream);

/* Bytes of 1024-byte-aligned device scratch wr_eval_rank_topk needs for `precision` (0 for precision 0): the bf16
 * copies of the gathered user rows and of the item table that the TMA descriptors point at. */
size_t wr_eval_scratch_bytes(int64_t R, int64_t n_items, int D, int precision);

/* wr_metrics: BaseRunner.evaluate_method (BaseRunner.py:76-88) from the ranks; float64 means.
 *   hr[i] = mean(rank <= ks[i]);  ndcg[i] = mean((rank <= ks[i]) / log2(rank + 1))
 * host_ks: HOST array of nk cut-offs (nk <= 8); out: DEVICE double [2*nk] = hr[0..nk), ndcg[0..nk).
 */
int wr_metrics(const int32_t *rank, int64_t R, const int *host_ks, int nk, double *out, void *ws, void *stream);

/* ---- row movement for the row-sharded tables (multi-GPU) ---------------------------
 * wr_gather_rows:      out[b] = T[idx[b]]            (owner side of the all-to-all of embedding rows)
 * wr_scatter_add_rows: G[idx[b]] += rows[b]          (owner side of the all-to-all of gradient rows)
 */
int wr_gather_rows(const float *T, const int64_t *idx, int64_t B, int D, int64_t n_rows, float *out, void *ws,
                   void *stream);
int wr_scatter_add_rows(float *G, const int64_t *idx, int64_t B, int D, int64_t n_rows, const float *rows,
                        void *ws, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WHISPRREC_B200_H */
