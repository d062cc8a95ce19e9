/*
 * laplace_b200.h -- C ABI of the B200-native graph-propagation library (liblaplace_b200.so).
 *
 * The reference (dream-faster/laplace-gnn-recommendation) is pure Python and has no FFI of its own;
 * its hot path bottoms out in torch_sparse / torch_scatter / ATen kernels.  Every entry point below
 * replaces one such call site (cited as reference file:line) and is what a ctypes binding in the
 * reference would bind.  Conventions:
 *
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory (outputs and workspaces are caller-allocated; *_ws_bytes tells how much);
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work and never synchronise
 *     unless documented;
 *   - indices handed in by the reference are int64; the library's own CSR arrays are int32
 *     (n_rows, nnz < 2^31 is checked);
 *   - dense operands are row-major fp32 with row stride == d;
 *   - return value 0 = success, otherwise an LGB_E* code; lgb_last_error() gives the message
 *     (thread-local).  There is no CPU fallback anywhere: without a CUDA device every compute call
 *     returns LGB_ECUDA.
 */
#ifndef LAPLACE_B200_H
#define LAPLACE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGB_ABI_VERSION 4

enum {
  LGB_OK = 0,
  LGB_EINVAL = 1,   /* bad argument (null pointer, negative size, d == 0, ...) */
  LGB_ERANGE = 2,   /* size does not fit the int32 index space */
  LGB_EWS = 3,      /* workspace too small */
  LGB_ECUDA = 4     /* CUDA runtime error (message has the cudaError string) */
};

int lgb_abi_version(void);
const char* lgb_last_error(void);
/* number of SMs of the current device (host-side query). */
int lgb_sm_count(int* out_host);

/* ---------------------------------------------------------------------------------------------
 * CSR / CSC construction -- replaces torch_sparse.SparseTensor(row=, col=, sparse_sizes=) as called
 * at data/lightgcn_loader.py:65-79: entries ordered by key row*n_cols+col, duplicates
 * kept, rowptr[i] = #{e: row[e] < i}.  perm (optional) receives the stable sorting permutation.
 * Also used per mini-batch for the hetero edge types (row = destination, col = source).
 * ------------------------------------------------------------------------------------------- */
int lgb_csr_build_ws_bytes(int64_t nnz, int64_t n_rows, size_t* bytes_host);
int lgb_csr_build(const int64_t* row, const int64_t* col, int64_t nnz, int64_t n_rows, int64_t n_cols,
                  int32_t* rowptr, int32_t* colidx, int64_t* perm, void* ws, size_t ws_bytes, void* stream);
/* torch_sparse's SparseTensor asserts row.max() < M and col.max() < N (the check that catches item ids which were not
 * shifted by both_indexes_from_zero, data/lightgcn_loader.py:39-43).  lgb_csr_build counts such entries on the device
 * (and clamps them, so nothing is written out of bounds); this call -- the ONE entry point that synchronises `stream` --
 * reads the count back from the workspace of that build and returns LGB_ERANGE when it is not zero. */
int lgb_csr_build_check(const void* ws, void* stream);

/* CSR -> CSC: colptr[n_cols+1], rowidx[nnz], csr2csc[nnz] = argsort(col*n_rows+row) (stable).  The
 * transposed operand of the SpMM backward (torch_sparse SPMMSum::backward; model/lightgcn.py:85-87). */
int lgb_csr_transpose_ws_bytes(int64_t nnz, int64_t n_cols, size_t* bytes_host);
int lgb_csr_transpose(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, int64_t n_cols,
                      int64_t nnz, int32_t* colptr, int32_t* rowidx, int32_t* csr2csc, void* ws,
                      size_t ws_bytes, void* stream);

/* dst[i] = src[perm[i]]  (values of the transposed matrix: valT = val[csr2csc]). */
int lgb_gather_f32(const float* src, const int32_t* perm, int64_t n, float* dst, void* stream);

/* gcn_norm(adj, add_self_loops=False) -- model/lightgcn.py:56 (PyG):
 * deg = row counts, dinv = deg^-1/2 (0 where deg == 0), val[e] = (1*dinv[row])*dinv[col]. Square matrix. */
int lgb_gcn_norm(const int32_t* rowptr, const int32_t* colidx, int64_t n, int64_t nnz, float* dinv,
                 float* val, void* stream);
/* Same weights from a caller-supplied dinv[n] (multi-GPU: a rank's local block needs the GLOBAL degrees):
 * val[e] = (1*dinv[row[e]])*dinv[colidx[e]]. */
int lgb_gcn_values(const int32_t* rowptr, const int32_t* colidx, int64_t n, int64_t nnz, const float* dinv,
                   float* val, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SpMM plan: rows longer than `chunk` non-zeros are split into fixed-size tasks whose partial sums
 * are reduced in a fixed order by a second stage (deterministic, no atomics).
 *   lgb_spmm_plan_count : counts_host[0] = #long rows, counts_host[1] = #tasks  (SYNCHRONISES the stream)
 *   lgb_spmm_plan_fill  : long_rows[n_long] ascending, long_ptr[n_long+1] (task range per long row),
 *                         task_row[n_tasks], task_start[n_tasks], task_end[n_tasks]
 * ------------------------------------------------------------------------------------------- */
int lgb_spmm_plan_count(const int32_t* rowptr, int64_t n_rows, int32_t chunk, int64_t* counts_host,
                        void* ws, size_t ws_bytes, void* stream);
int lgb_spmm_plan_ws_bytes(int64_t n_rows, size_t* bytes_host);
int lgb_spmm_plan_fill(const int32_t* rowptr, int64_t n_rows, int32_t chunk, int64_t n_long,
                       int64_t n_tasks, int32_t* long_rows, int32_t* long_ptr, int32_t* task_row,
                       int32_t* task_start, int32_t* task_end, void* ws, size_t ws_bytes, void* stream);

/* Optional degree-bucketed row order: rows sorted by descending ceil(log2(deg)) bucket, stable. */
int lgb_degree_order_ws_bytes(int64_t n_rows, size_t* bytes_host);
int lgb_degree_order(const int32_t* rowptr, int64_t n_rows, int32_t* row_order, void* ws, size_t ws_bytes,
                     void* stream);

typedef struct lgb_csr {
  int64_t n_rows;
  int64_t n_cols;
  int64_t nnz;
  const int32_t* rowptr;     /* [n_rows+1] */
  const int32_t* colidx;     /* [nnz] */
  const float* val;          /* [nnz] or NULL (all ones) */
  const int32_t* row_order;  /* [n_rows] or NULL (natural order) */
  int32_t chunk;             /* split threshold used to build the plan (0 = no plan: every row by one warp) */
  int32_t _pad;
  int64_t n_long;
  int64_t n_tasks;
  const int32_t* long_rows;
  const int32_t* long_ptr;
  const int32_t* task_row;
  const int32_t* task_start;
  const int32_t* task_end;   /* [n_tasks] exclusive end offset of each slice */
  /* hot-column plan (optional; LGB_SPMM variant 30): the n_hot most-referenced columns of this matrix are kept in shared
   * memory by every CTA.  colidx_hot is colidx with a hot column c replaced by ~slot (negative), slot = its position in
   * hot_cols (descending column degree).  NULL / 0 = no plan. */
  const int32_t* colidx_hot; /* [nnz] */
  const int32_t* hot_cols;   /* [n_hot] */
  int32_t n_hot;
  int32_t _pad2;
  /* stage-2 tree plan (optional): the partial rows [long_ptr[L], long_ptr[L+1]) of every long row cut into segments of at
   * most 32; seg_row[s] = L, [seg_t0[s], seg_t1[s]) = its partial rows, row_seg0[L] = first segment of long row L
   * ([n_long + 1]).  Used when the call passes LGB_SPMM_TREE_WS (scratch sized for it, see lgb_spmm). */
  const int32_t* seg_row;
  const int32_t* seg_t0;
  const int32_t* seg_t1;
  const int32_t* row_seg0;
  int64_t n_seg;
  /* execution order of the long-row slices (optional): the CTA / warp with slice slot i runs slice task_exec[i] (a permutation of
   * 0..n_tasks-1); NULL = plan order (row by row).  Partial sums stay indexed by slice id, so stage 2 and the summation order of
   * every row -- hence the result, bit for bit -- do not depend on it.  The host plan sorts the slices by their first column
   * ("column sweep"): slices of different long rows that gather the same operand rows run next to each other and find them in L2. */
  const int32_t* task_exec;  /* [n_tasks] */
  const int32_t* task_seg;   /* [n_tasks] segment of the stage-2 tree plan that slice t belongs to (LGB_SPMM_FUSED_STAGE2), or NULL */
} lgb_csr;

/* ---------------------------------------------------------------------------------------------
 * Fused SpMM -- replaces torch_sparse.matmul(adj_t, x) (model/lightgcn.py:85-87), its autograd
 * backward (same kernel on the CSC arrays), the stack/mean read-out (model/lightgcn.py:67-68) and,
 * with val == NULL, the per-edge-type gather + scatter-{add,mean} of PyG SAGEConv
 * (model/layers.py:9-24 -> MessagePassing.propagate -> torch_scatter).
 *
 *   t[r,:]  = sum_{e in row r} val[e] * X[colidx[e],:]            (val == NULL: 1)
 *   if (flags & LGB_SPMM_MEAN)  t[r,:] /= max(deg(r), 1)
 *   y       = t (+ resid[r,:] if resid)
 *   if (Y)        Y[r,:]       = y
 *   if (acc_out)  acc_out[r,:] = ((acc_in ? acc_in[r,:] : 0) + y) / acc_div
 *
 * X/Y/resid/acc_* are [n, d] row-major.  partial_ws must hold n_tasks*d floats when g->n_tasks > 0; the hot-column
 * variants (30 / 31) need n_tasks*d + 64 floats whose last 64 are ZERO before the first launch (their work counters; every
 * launch leaves them zero) and a buffer that no concurrent launch shares.  With LGB_SPMM_TREE_WS in flags the caller
 * declares a scratch of n_tasks*d + 64 + n_seg*d + n_long floats, everything behind the first n_tasks*d ZERO before the
 * first launch: stage 2 then runs as a tree over the plan's segments (one warp per 32 partial rows, self-resetting
 * tickets) instead of one CTA per long row.
 * ------------------------------------------------------------------------------------------- */
#define LGB_SPMM_MEAN 1
#define LGB_SPMM_TREE_WS 2     /* partial_ws is sized and zeroed for the stage-2 tree (see above) */
/* With LGB_SPMM_FUSED_STAGE2 (and LGB_SPMM_TREE_WS, a plan that carries task_seg, and a scratch of n_seg more ZERO ints behind
 * the tree's: n_tasks*d + 64 + n_seg*d + n_long + n_seg) the sub-warp kernels run the tree INSIDE the main launch: the warp that
 * stores the last partial row of a segment (a ticket per segment) adds the segment, the warp that finishes the last segment of
 * a row adds the level-2 rows and runs the epilogue.  Same sums in the same fixed order as the separate stage-2 kernel (who does
 * them depends on timing, what they compute does not); one launch less per call, and the reductions overlap with the short rows
 * that follow the slices in the grid.  Kernel families without the fused path ignore the flag and launch stage 2 as before. */
#define LGB_SPMM_FUSED_STAGE2 4
/* bits 4..11 of flags pick a kernel variant for A/B measurements (all of them for d in 33..64; 0, 1 and 16 for d <= 32):
 * 0 = tuned default (sub-warp rows: one lane group per short row, 64 resident warps per SM), 1 = first version (warp per
 * row, unroll 8), 2/3 = software-pipelined persistent warps, 4..6, 12 = warp per row at other unroll / occupancy points,
 * 7..11 = cp.async rings, 13..15 = sub-warp rows at other unroll depths, 16 = sub-warp rows + one CTA per slice of a long
 * row, 17 = the 64-bit-index family that tables with n_cols*d/4 >= 2^31 elements get automatically,
 * 18 = sub-warp rows + L2 prefetch of the epilogue operands, double-buffered (col,val) batches and an L2 evict_first
 * policy on those streamed loads, 19 = 16 + 18,
 * 20..22 (d in 33..64) = four rows per warp: 8 lanes x 2 float4 per row, with CTA-wide slices (22: + the prefetches of 18),
 * 23..25 (d = 64) = the same with one 256-bit load per lane and non-zero (LDG.E.256), 26/27 (d = 128) = warp per row with
 * 256-bit gathers (16 lanes x 32 bytes per row; unroll 1 at 64 warps/SM, unroll 2 at 40), 30/31 = hot-column cache (needs
 * the plan's colidx_hot / hot_cols; two / one persistent 1024-thread CTAs per SM).  Every variant computes the same operator (rtol 1e-5); summation order inside a row differs between families. */
#define LGB_SPMM_VARIANT_SHIFT 4
int lgb_spmm(const lgb_csr* g, const float* X, int32_t d, float* Y, const float* resid,
             const float* acc_in, float* acc_out, float acc_div, int32_t flags, float* partial_ws,
             void* stream);

/* Same kernel with a split epilogue: rows >= split_row skip the epilogue and store their raw sums to
 * y_tail[(r - split_row), :] -- multi-GPU: one launch computes the owned user rows (fused epilogue) AND the partial
 * item rows that go to the exchange buffer.  d % 4 == 0; variants 2/3 are not available. */
int lgb_spmm_split(const lgb_csr* g, const float* X, int32_t d, float* Y, const float* resid, const float* acc_in,
                   float* acc_out, float acc_div, int32_t flags, float* partial_ws, int64_t split_row, float* y_tail,
                   void* stream);
/* lgb_spmm for an operand X that is ZERO outside a known set of rows: bit c of x_row_bitmap ([ceil(n_cols / 32)] words) set
 * <=> row c of X may be non-zero.  Same result as lgb_spmm (entries that multiply a zero row contribute nothing), but such
 * entries are never gathered: the kernel streams colidx, tests the bitmap and fetches the flagged rows only.  Use: the FIRST
 * layer of the LightGCN backward -- dE_f is non-zero on the <= 3*B rows of the BPR batch (utils/metrics_lightgcn.py:9-45 through
 * the six gathers of run_pipeline_lightgcn.py:133-144), so A^T dE_f touches 3*B of the N operand rows (B = 128: 384 of 1.48 M).
 * resid_row_bitmap (optional, [ceil(n_rows / 32)] words): the same promise for resid -- rows that are not flagged are not read
 * (in that backward layer resid is dE_f itself: 351 MB of zeros not fetched, one dependent load less per output row).
 * d <= 64 with d % 4 == 0 uses the filtered kernels; other widths run lgb_spmm's dense kernels.  lgb_rows_bitmap builds the bitmap. */
int lgb_spmm_rowsparse(const lgb_csr* g, const float* X, const uint32_t* x_row_bitmap, int32_t d, float* Y,
                       const float* resid, const uint32_t* resid_row_bitmap, const float* acc_in, float* acc_out,
                       float acc_div, int32_t flags, float* partial_ws, void* stream);
/* bitmap[(idx[i] + offset) / 32] |= 1 << ((idx[i] + offset) % 32) for i < n (atomicOr; the caller zeroes the bitmap once per
 * batch, lgb_zero).  Indices outside [0, n_bits) are ignored. */
int lgb_rows_bitmap(const int64_t* idx, int64_t n, int64_t offset, int64_t n_bits, uint32_t* bitmap, void* stream);
/* out = scale * a ([n_rows, d]) and, in the same pass, WHICH rows of a hold a non-zero: bit row_offset + r of bitmap is set and
 * count[0] incremented for each such row r (caller zeroes bitmap and count).  The autograd backward of LightGCN.forward receives dE_f as
 * a dense tensor (model/lightgcn.py:46-80 under torch autograd: the six row gathers of run_pipeline_lightgcn.py:133-144 scatter
 * into zeros); this finds its batch rows for lgb_spmm_rowsparse without another pass.  d in {4, 8, 16, 32, 64, 128}. */
int lgb_scale_rows_nonzero(const float* a, int64_t n_rows, int32_t d, float scale, float* out, int64_t row_offset,
                           uint32_t* bitmap, int32_t* count, void* stream);

/* scatter-max per destination (PyG aggr="max"): Y[r,:] = max_e X[colidx[e],:] (0 for empty rows),
 * argmax[r,:] = the winning source row (or -1), used by the backward. */
int lgb_segment_max(const lgb_csr* g, const float* X, int32_t d, float* Y, int32_t* argmax, void* stream);
int lgb_segment_max_bwd(const float* gY, const int32_t* argmax, int64_t n_rows, int32_t d, float* gX_zeroed,
                        void* stream);

/* X[r,:] /= max(deg(r),1)  into out (mean-aggregation backward pre-scale). */
int lgb_row_div_by_degree(const float* X, const int32_t* rowptr, int64_t n_rows, int32_t d, float* out,
                          void* stream);

/* cudaMemsetAsync(p, 0, bytes) on `stream` (gradient buffers that the atomic scatters accumulate into). */
int lgb_zero(void* p, size_t bytes, void* stream);

/* out[i] = ((acc ? acc[i] : 0) + y[i] + (resid ? resid[i] : 0)) / div over n floats -- the SpMM epilogue as a
 * stand-alone pass, for rows whose sum arrives from a collective (multi-GPU item rows). out may alias acc or y. */
int lgb_accumulate(const float* y, const float* acc, const float* resid, int64_t n, float div, float* out,
                   void* stream);
/* out = (((s0 + s1) + s2) + ...) / div over `count` (<= 8) equally-shaped fp32 buffers (srcs_host: HOST array of device
 * pointers): mean(stack(embs, dim=1), dim=1) of model/lightgcn.py:67-68 in one pass, same left-to-right order as the
 * fused accumulate epilogue of lgb_spmm.  out must not alias a source. */
int lgb_mean_rows(const float* const* srcs_host, int32_t count, int64_t n, float div, float* out, void* stream);

/* out = cat(a[na,d], b[nb,d]) * scale   (one pass; cat/split/mean backward of model/lightgcn.py:58,67-72). */
int lgb_scale_concat(const float* a, int64_t na, const float* b, int64_t nb, int32_t d, float scale,
                     float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused BPR -- replaces the six row gathers (run_pipeline_lightgcn.py:133-144) + bpr_loss
 * (utils/metrics_lightgcn.py:9-45) + their autograd backward (index_put accumulate).
 *
 * Operand k in {u_f,u_0,p_f,p_0,n_f,n_0}: row b lives at base_k + (idx ? idx[b] : b)*d, where idx is
 * iu for the two user operands, ip for the positives and in for the negatives.
 *   x_b   = <u_f,p_f> - <u_f,n_f>
 *   loss  = -(1/B) sum_b softplus(x_b) + lambda * sum_b (|u_0|^2 + |p_0|^2 + |n_0|^2)
 * Backward (when any grad pointer is non-NULL), with g = *gout (device scalar; NULL = 1):
 *   d u_f += -g*sigma(x_b)/B*(p_f-n_f) * gscale;  d p_f += -g*sigma/B*u_f * gscale;  d n_f += +g*sigma/B*u_f * gscale
 *   d k_0 += 2*lambda*g*k_0
 * With index arrays the gradient rows are accumulated with red.global.add.v4.f32 into caller-ZEROED
 * [N,d] buffers (duplicates add up); without indices they are plain stores into [B,d] buffers.
 * ws: 2*lgb_bpr_blocks(B) floats.  loss: device float[1].
 * ------------------------------------------------------------------------------------------- */
#define LGB_BPR_FILTER_USER_ROWS_ONLY 1   /* the owned-user filter skips only the USER operand of a foreign triple (its item rows are
                                            still processed): the layer-0 regulariser gradients of a user-sharded table whose
                                            item block is replicated.  Only with du0 / dp0 / dn0 outputs. */
typedef struct lgb_bpr_args {
  const float* uf; const float* u0; const float* pf; const float* p0; const float* nf; const float* n0;
  const int64_t* iu; const int64_t* ip; const int64_t* in;   /* all three NULL, or all three set */
  int64_t B;
  int64_t B_norm;          /* the 1/B of the mean; 0 = use B.  A rank that holds a shard of a global batch passes the global size */
  int32_t d;
  float lambda;
  float gscale;            /* extra factor folded into the *_f gradients, e.g. 1/(K+1) */
  int32_t flags;           /* LGB_BPR_* */
  int64_t user_lo;         /* with index arrays: only triples with user_lo <= iu[b] < user_hi are processed and the   */
  int64_t user_hi;         /* user row becomes iu[b]-user_lo (a rank's shard of a global batch); user_hi == 0 = no filter */
  const float* gout;       /* device scalar upstream gradient or NULL */
  float* duf; float* du0; float* dpf; float* dp0; float* dnf; float* dn0;   /* each may be NULL */
  float* loss;             /* device float[1], may be NULL when only gradients are wanted */
  float* ws;
} lgb_bpr_args;
int64_t lgb_bpr_blocks(int64_t B);
int lgb_bpr(const lgb_bpr_args* a, void* stream);
/* A global batch against a user-sharded table (multi-GPU BPR; the reference gathers from one table, run_pipeline_lightgcn.py:133-144):
 *   gather : dst[b, :] = lo <= idx[b] < hi ? src[idx[b]-lo, :] : 0      (summed over ranks = the gathered batch rows)
 *   scatter: dst[idx[b]-lo, :] += src[b, :] for the owned b             (atomic; dst zeroed or holding earlier terms) */
int lgb_gather_rows_owned(const float* src, const int64_t* idx, int64_t B, int32_t d, int64_t lo, int64_t hi,
                          float* dst, void* stream);
int lgb_scatter_add_rows_owned(const float* src, const int64_t* idx, int64_t B, int32_t d, int64_t lo, int64_t hi,
                               float* dst, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Edge decoder -- model/encoder_decoder.py:55-72.
 *   concat: out[e, :] = [zu[row[e], :du] , zi[col[e], :di]]   (the reference's gather + cat; MLP stays cuBLAS)
 *   dot   : out[e]    = <zu[row[e]], zi[col[e]]>              (the decoder BASELINE.json's north_star names)
 * Backward accumulates into caller-zeroed dzu/dzi with vector atomics.
 * ------------------------------------------------------------------------------------------- */
int lgb_edge_concat_fwd(const float* zu, const float* zi, const int64_t* row, const int64_t* col, int64_t L,
                        int32_t du, int32_t di, float* out, void* stream);
int lgb_edge_concat_bwd(const float* gout, const int64_t* row, const int64_t* col, int64_t L, int32_t du,
                        int32_t di, float* dzu_zeroed, float* dzi_zeroed, void* stream);
int lgb_edge_dot_fwd(const float* zu, const float* zi, const int64_t* row, const int64_t* col, int64_t L,
                     int32_t d, float* out, void* stream);
int lgb_edge_dot_bwd(const float* zu, const float* zi, const int64_t* row, const int64_t* col,
                     const float* gout, int64_t L, int32_t d, float* dzu_zeroed, float* dzi_zeroed,
                     void* stream);
/* Inference form of the reference's two-layer decoder (model/encoder_decoder.py:55-72 with the default layers of
 * model/layers.py:35-56: Linear(du+di -> H), relu, Linear(H -> 1)) when dropout is off: because the first Linear acts on a
 * concatenation, W1 [zu[r] ; zi[c]] + b1 = pu[r] + pi[c] with the per-NODE projections pu = zu W1[:, :du]^T + b1 and
 * pi = zi W1[:, du:]^T (two library GEMMs over the nodes instead of one over the label edges), and
 *   out[e] = b2 + sum_h w2[h] * relu(pu[row[e], h] + pi[col[e], h]).
 * Forward only (training-mode dropout acts on the per-edge concatenation, so the training step keeps the concat form). */
int lgb_edge_mlp2_fwd(const float* pu, const float* pi, const int64_t* row, const int64_t* col, int64_t L, int32_t H,
                      const float* w2, const float* b2, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight / bias gradient of a Linear layer over many rows and few features -- the AddmmBackward of SAGEConv's
 * lin_l / lin_r (model/layers.py:9-24) and of the decoder's Linear layers (model/encoder_decoder.py:55-72):
 *   dW[o, i] = sum_n dY[n, o] * X[n, i]        db[o] = sum_n dY[n, o]   (db may be NULL)
 * X [N, in], dY [N, out], dW [out, in] row-major fp32.  Split-K over the rows (fp32 FMA, fixed-order reduction of the
 * partial tiles: deterministic); ws: lgb_linear_wgrad_ws_bytes(N, in, out).
 * ------------------------------------------------------------------------------------------- */
int lgb_linear_wgrad_ws_bytes(int64_t N, int32_t in, int32_t out, size_t* bytes_host);
int lgb_linear_wgrad(const float* X, const float* dY, int64_t N, int32_t in, int32_t out, float* dW, float* db,
                     void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Candidate generation -- make_predictions_for_user (utils/metrics_lightgcn.py:125-142) for a block of
 * users: scores = Wu[u] . Wi^T (fp32 FMA, ascending-d order), seen items masked, top-k by (score desc,
 * id asc).  seen CSR (seen_ptr[n_users_total+1], seen_idx) is indexed by user id (NULL = nothing seen).
 * k <= 1024, d <= 512.  score_ws: n_users*n_items floats.  out_ids[u*k + j] = -1 past the last unseen item.
 * ------------------------------------------------------------------------------------------- */
int lgb_topk_exclude(const float* Wu, const float* Wi, const int64_t* users, int64_t n_users, int64_t n_items,
                     int32_t d, const int32_t* seen_ptr, const int32_t* seen_idx, int32_t k, int64_t* out_ids,
                     float* out_scores, float* score_ws, void* stream);

/* Same result (bit-identical scores and ids), scoring tiled over 8 users per CTA: every item row is loaded once per CTA
 * instead of once per user, which turns the U x I x d contraction from L2-bound into FMA-bound; selection per user as above. */
int lgb_topk_exclude_tiled(const float* Wu, const float* Wi, const int64_t* users, int64_t n_users, int64_t n_items,
                     int32_t d, const int32_t* seen_ptr, const int32_t* seen_idx, int32_t k, int64_t* out_ids,
                     float* out_scores, float* score_ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Negative-sample rejection test -- the np.isin step of PyG structured_negative_sampling
 * (data/lightgcn_loader.py:105-107): mask[j] = 1 iff key(row[j], cand[j]) is in the
 * positive key set {row*num_nodes + col} (plus self-loop keys i*(num_nodes+1), i < num_nodes, when
 * with_self_loops != 0).  pos_keys_sorted is the ascending int64 key array.  The random draws
 * themselves stay on torch's CPU generator so the indices are bit-identical to the reference.
 * ------------------------------------------------------------------------------------------- */
int lgb_neg_reject_mask(const int64_t* row, const int64_t* cand, int64_t n, int64_t num_nodes,
                        const int64_t* pos_keys_sorted, int64_t n_pos, int32_t with_self_loops,
                        uint8_t* mask, void* stream);
int lgb_sort_keys_ws_bytes(int64_t n, size_t* bytes_host);
/* keys_out = sort(row*num_nodes + col) ascending. */
int lgb_edge_keys_sorted(const int64_t* row, const int64_t* col, int64_t n, int64_t num_nodes,
                         int64_t* keys_out, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sub-graph batch assembly -- the index-heavy steps of GraphDataset.__getitem__ / fetch_n_hop_neighbourhood
 * / remap_edges_to_start_from_zero (data/dataset.py:39-182,258-300) for a whole batch of root users at once.
 *   lgb_segment_expand:      adjacency (ptr[N+1], idx) in int64; frontier nodes[n]; out_off[n+1] = exclusive scan of their
 *                            degrees; for every adjacency entry j: out_pos[j] = frontier position, out_nbr[j] = neighbour
 *                            (adjacency order kept).
 *   lgb_bucketize_segmented: out[j] = torch.bucketize(values[j], buckets[bucket_ptr[seg[j]] : bucket_ptr[seg[j]+1]]).
 * ------------------------------------------------------------------------------------------- */
int lgb_segment_expand(const int64_t* ptr, const int64_t* idx, const int64_t* nodes, const int64_t* out_off, int64_t n,
                       int64_t total, int64_t* out_pos, int64_t* out_nbr, void* stream);
int lgb_bucketize_segmented(const int64_t* values, const int64_t* seg, int64_t n, const int64_t* buckets,
                            const int64_t* bucket_ptr, int64_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Adam step -- optim.Adam(model.parameters(), lr) + optimizer.step() of run_pipeline_lightgcn.py:103,159
 * (betas/eps as given, no weight decay, no amsgrad), torch's single-tensor arithmetic in one streaming pass.
 * `step` is the 1-based step count AFTER the increment (torch's state["step"]).
 * ------------------------------------------------------------------------------------------- */
int lgb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, int32_t step, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU item-block exchange over CUDA symmetric memory (new: the reference is single-process).
 * Sum-all-reduce, in place, of an fp32 buffer that every rank allocated symmetrically; rank g reduces and
 * republishes slice g.  The caller brackets the call with symmetric-memory barriers on the same stream.
 *   lgb_multimem_allreduce_f32 : NVSwitch multicast address (multimem.ld_reduce / multimem.st, NVLS)
 *   lgb_peer_allreduce_f32     : per-peer UVA pointers (peer_ptrs_host[world], HOST array of device addresses)
 * ------------------------------------------------------------------------------------------- */
int lgb_multimem_allreduce_f32(void* multicast_ptr, int64_t n_floats, int32_t rank, int32_t world, void* stream);
/* The exchange the sharded engine ships: the two barriers and the reduce + republish loop in ONE launch, no host
 * synchronisation, replayable from a CUDA graph.  `x` describes ONE symmetric arena (same size on every rank): its
 * multicast address (NULL when the fabric has none: peer loads / stores are used), the arena's address on every rank
 * and a symmetric, zero-initialised array of n_channels * lgb_exchange_pad_words(world) uint32 signal slots on every
 * rank.  The call reduces the n_floats at byte_offset of the arena in place.  Exchanges that may be in flight at the
 * same time (different streams) must use different channels; every rank must issue the exchanges of one channel in the
 * same order and with the same flags.  LGB_EXCHANGE_NO_BARRIER skips both barriers (single-process tests only).  A barrier
 * that waits longer than ~4 s traps. */
#define LGB_EXCHANGE_MAX_WORLD 16
#define LGB_EXCHANGE_NO_BARRIER 1
#define LGB_EXCHANGE_PEER 2            /* peer loads / stores even when a multicast address is given: (G-1)/G of the buffer per
                                          link direction instead of (G+1)/G, at the price of G loads per element in the SM */
#define LGB_EXCHANGE_BLOCKS_SHIFT 8    /* bits 8..15: CTAs of the launch (0 = 64; at most 128) -- the same value on every rank */
typedef struct lgb_exchange {
  void* multicast_base;                          /* NVSwitch multicast mapping of the arena, or NULL */
  void* peer_base[LGB_EXCHANGE_MAX_WORLD];       /* the arena on rank r (UVA / symmetric-memory pointer) */
  void* pad_base[LGB_EXCHANGE_MAX_WORLD];        /* the signal-slot array on rank r */
  int32_t rank, world, n_channels, _pad;
} lgb_exchange;
int lgb_exchange_pad_words(int32_t world);
int lgb_exchange_allreduce_f32(const lgb_exchange* x, int64_t byte_offset, int64_t n_floats, int32_t channel,
                               int32_t flags, void* stream);
int lgb_peer_allreduce_f32(const uint64_t* peer_ptrs_host, int64_t n_floats, int32_t rank, int32_t world,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAPLACE_B200_H */
