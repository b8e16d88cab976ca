/* parsenet_b200.h — C ABI of libparsenet_b200.so (hand-written sm_100a kernels for the ParSeNet hot path).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Conventions for every entry point
 *   - all pointers are DEVICE pointers (fp32 / int32 / fp64 as typed); the caller owns every buffer, the library
 *     allocates nothing and keeps no state besides an error string and a launch counter;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and never synchronise;
 *   - return value 0 = ok, non-zero = error (pn_last_error() gives the message) — callers raise, nothing falls back;
 *   - activations are point-major: (shape b, point n, channel c) at base[(b*Np + n)*pitch + c]; `ld*` arguments are
 *     row pitches in elements, so channel slices of a wider buffer can be passed without copies;
 *   - buffers documented "zeroed" must be zero-filled by the caller (they are accumulated with atomics).
 * Each declaration cites the reference code (relative to Hippogriff/parsenet-codebase) that it replaces.
 * The reference has no native layer: its FFI for this path is "call a torch op"; INTEGRATION.md shows the
 * ctypes binding (parsenet-codebase_b200/pnb200/cabi.py) a maintainer would add to call these from src/*.py.
 */
#ifndef PARSENET_B200_H
#define PARSENET_B200_H

#ifdef __cplusplus
extern "C" {
#endif


/* ---- cabi.cu ---- */
/* replaces: (error string of the last failing call on this thread; the reference drops into ipdb instead) */
const char* pn_last_error();
/* replaces: (number of kernel launches issued through this library; bench.py gpu_launches) */
unsigned long long pn_launch_count();
void pn_reset_launch_count();
int pn_abi_version();

/* ---- chamfer.cu ---- */
/* replaces: chamfer_distance / _one_side / _single_shape: src/utils.py:273-358 */
int pn_chamfer_nn_fwd(const float* A, int Na, const float* Bs, int Nb, int B, float* mind, int* arg, void* stream);
/* replaces: autograd of the min over the broadcast difference tensor */
int pn_chamfer_nn_bwd(const float* A, int Na, const float* Bs, int Nb, int B, const int* arg, const float* g, float* dA_accum, float* dB_accum, void* stream);

/* ---- edgeconv.cu ---- */
/* replaces: get_graph_feature[_with_normals] + conv + norm + LeakyReLU + max over k: src/PointNet.py:72-140,176-192; src/model.py:25-53,140-154 */
int pn_edge_gather_fwd(const float* PQ, long long ldpq, const int* idx, int B, int N, int k, int Cout, const float* gamma, float* esel, int* jsel, float* esum, double* stats, int G, int stats_per_shape, void* stream);
/* replaces: same (activation of the selected extreme) */
int pn_edge_apply(const float* esel, const float* scale, const float* shift, float* out, long long ldo, int B, int N, int Cout, void* stream);
/* replaces: autograd of the edge-conv block */
int pn_edge_bwd_prep(const float* g, long long ldg, const float* esel, const float* scale, const float* shift, const float* mean_rstd, const float* gamma, int B, int N, int Cout, int G, int stats_per_shape, float* dy, double* gsum, float* dgamma, float* dbeta, void* stream);
/* replaces: autograd of feature[idx] gather (index_put in the reference): src/PointNet.py:93 */
int pn_knn_csr_transpose(const int* idx, int B, int N, int k, int* cnt, int* off, int* cursor, int* rev, void* stream);
/* replaces: autograd of the edge-conv block */
int pn_edge_bwd(const float* PQ, long long ldpq, const float* dy, const float* esum, const int* jsel, const int* off, const int* rev, const float* mean_rstd, const double* gsum, const float* scale, int B, int N, int k, int Cout, int G, int stats_per_shape, double count, int dense, float* dPQ, long long lddpq, void* stream);

/* ---- fit.cu ---- */
/* replaces: Fit.fit_{plane,sphere,cylinder,cone}_torch reductions: src/primitive_forward.py:708-843; LeastSquares.lstsq src/fitting_utils.py:36-65; CustomSVD :420-455 */
int pn_fit_moments_fwd(const float* P, const float* Nr, const float* W, long long ldw, int S, int start, int step, int m, float eps, double* mom_zeroed, void* stream);
/* replaces: autograd of the weighted reductions w.r.t. the membership weights */
int pn_fit_moments_bwd(const float* P, const float* Nr, const float* W, long long ldw, int S, int start, int step, int m, float eps, const float* gmom, float* gW, long long ldg, void* stream);
/* the same two kernels over all shapes of a step in one launch (gridDim.y = shape): P / Nr [B][N][3], W [B][N][ldw] with
   column = segment slot, mom [B][S][55]; replaces the per-shape python loop of Evaluation.fitting_loss src/residual_utils.py:86-208 */
int pn_fit_moments_fwd_batched(const float* P, const float* Nr, const float* W, long long ldw, int B, int N, int S, int start, int step, int m, float eps, double* mom_zeroed, void* stream);
int pn_fit_moments_bwd_batched(const float* P, const float* Nr, const float* W, long long ldw, int B, int N, int S, int start, int step, int m, float eps, const float* gmom, float* gW, long long ldg, void* stream);

/* ---- fitsolve.cu ---- */
/* replaces: the per-segment 3x3 algebra of Fit.fit_{plane,sphere,cylinder,cone}_torch src/primitive_forward.py:708-831 +
   LeastSquares.lstsq / best_lambda src/fitting_utils.py:36-85 + CustomSVD and its backward rule :385-455, for every segment
   slot of every shape in ONE launch: mom [S][55] -> par [S][8] (residual-kernel layout) and jac [S][8][55] = d par / d mom
   (forward mode; the backward is grad_mom = grad_par . jac); kind -1 none, 0 plane, 1 sphere, 2 cylinder, 3 cone;
   bad[s] = 1 for a degenerate cone (cond > 1e5, :818-823) */
int pn_fit_solve(const double* mom, const int* kind, int S, int rows, double* par, double* jac, float* bad, void* stream);

/* ---- weights.cu ---- */
/* replaces: weights_normalize src/fitting_utils.py:306-325 (exp(clamp(w / bw^2 / 2)), normalise over the clusters of a point,
   min-max over the points of a cluster; single-cluster early return :318-319) for all shapes of a step: raw / out [B][N][64]
   (slots >= K[b] are padding), bw2 [B], K [B]; keys [2][B][64] u64 workspace (first half all-ones, second half zero on entry)
   holding the packed (value, point) extrema for the backward */
int pn_weights_normalize_fwd(const float* raw, const float* bw2, const int* K, int B, int N, int S, float* out, unsigned long long* keys, void* stream);
/* replaces: autograd of the same expression (gradient of min / max goes to one point per cluster: lowest index among ties);
   red [B][64][2] doubles zero-initialised workspace */
int pn_weights_normalize_bwd(const float* raw, const float* g, const float* bw2, const int* K, int B, int N, int S, const unsigned long long* keys, double* red_zeroed, float* graw, void* stream);

/* ---- gridloss.cu ---- */
/* replaces: control_points_permute_reg_loss src/loss.py:76-97 (mode 0: 8 dihedral re-orderings of the target grid, all_permutations
   :21) and control_points_permute_closed_reg_loss :100-124 (mode 1: g cyclic shifts along u (roll :60) x 4 flips,
   all_permutations_half :41): out, gt [B][g][g][3] -> loss_b [B] = min_p sum (out - cand_p(gt))^2, pick [B], best [B][g][g][3] */
int pn_grid_perm_fwd(const float* out, const float* gt, int B, int g, int mode, float* diff_ws, float* loss_b, int* pick, float* best, void* stream);
/* replaces: autograd of the same (dout = 2 (out - best) * gscale[0] * inv) */
int pn_grid_perm_bwd(const float* out, const float* best, long long n, const float* gscale, float inv, float* dout, void* stream);
/* replaces: laplacian_loss src/loss.py:213-239 (two depthwise 3x3 convolutions with zero padding): l_ws [B][g][g][3] = lap(out) - lap(gt),
   part [B] = sum l^2 (l1 = 0) or sum |l| (l1 = 1) */
int pn_grid_laplacian_fwd(const float* out, const float* gt, int B, int g, int l1, float* l_ws, float* part, void* stream);
/* replaces: autograd of the same (the zero-padded stencil is self-adjoint); dgt may be null */
int pn_grid_laplacian_bwd(const float* l_ws, int B, int g, int l1, const float* gscale, float inv, float* dout, float* dgt, void* stream);

/* ---- knn.cu ---- */
/* replaces: src/PointNet.py:9-26 (knn), :29-69 (knn_points_normals); src/model.py:9-22 */
int pn_knn(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64, float* dist_out, float* ws_norms, void* stream);
/* the same graph build for the 32-channel-multiple feature spaces with TMA-staged point blocks (knn_tma.cu: cp.async.bulk.tensor
   ring, row-major swizzled tiles, admission straight from registers); identical results, identical arguments */
/* the same graph for the tiles of query rows that contain a row with flags[b][row] != 0 only (other tiles exit at once and leave
   their rows of idx_out / dist_out untouched): exact fall-back of pn_knn_lowdim */
int pn_knn_flagged(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64, float* dist_out, float* ws_norms, const int* flags, void* stream);
/* replaces: knn_points_normals src/PointNet.py:29-69 (metric 1, C = 6), knn on raw positions src/PointNet.py:9-26 / src/model.py:9-22
   (metric 0, C = 3), the 5-nearest search of up_sample_points_torch src/fitting_utils.py:150-163 (metric 2) — same arguments and
   the same bit-exact graph as pn_knn with a one-pass bracketed selection instead of the streaming top-k: exact costs of every
   row to a strided column sample give a per-row bracket, one pass appends the costs below it to per-row lists, a warp per row
   selects the k-th smallest, sorts the entries up to it by (cost, index) and writes the first k.  Workspaces: ws_norms [B*N],
   ws_T [B*N], ws_val [B*N][cap] u32, ws_col [B*N][cap] u16, ws_cnt [B*N][2], cap = 1024 or 2048 (a list may use cap / 2 entries).
   flags [B*N] is written: rows with flag 1 are NOT written — run pn_knn_flagged with the same flags next; knn_lowdim.cu */
int pn_knn_lowdim_supported(int N, int C, int k, int metric);
int pn_knn_lowdim(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64, float* dist_out, int stride, int b_sample, float* ws_norms, float* ws_T, unsigned* ws_val, unsigned short* ws_col, int* ws_cnt, int cap, int* flags, void* stream);
int pn_knn_tma_supported(const float* x, int N, int C, int ld, int k, int metric);
/* the same graph for the 64-row tiles that contain a row with flags[b][row] != 0 only (other tiles exit at once and leave their
   rows of idx_out / dist_out untouched): exact fall-back of pn_knn_tc */
int pn_knn_tma_flagged(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64, float* dist_out, float* ws_norms, const int* flags, void* stream);
/* replaces: knn: src/PointNet.py:9-26, src/model.py:9-22 for C = 64 / 128 — same arguments and the same bit-exact graph as pn_knn /
   pn_knn_tma, computed with the tensor cores as a FILTER: a split-TF32 tcgen05 pass gives every pair's cost with a proven error
   interval, a sampled bracket keeps ~6 % of a row in per-row lists, the k-th smallest upper bound bounds the exact k-th cost,
   and only the survivors (k plus the few inside the interval) get the reference-order fp32 fmaf chain and the final sort.
   stride / b_sample: column sample {0, stride, ...} (<= 1024 columns) and the order statistic used as bracket.  Workspaces:
   ws_norms [B*N], ws_xs [B*N*C], ws_colc [B*(Np + mp)] (Np, mp = N and the sample size rounded up to 64), ws_T [B*N], ws_val
   [B*N][cap] u32, ws_col [B*N][cap] u16, ws_cnt [B*N][2], cap = 1024 or 2048.  flags [B*N] is written: rows with flag 1 could not
   be decided and are NOT written — run pn_knn_tma_flagged with the same flags next (no host round trip); knn_tc.cu */
int pn_knn_tc_supported(const float* x, int N, int C, int ld, int k, int metric);
int pn_knn_tc(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64, float* dist_out, int stride, int b_sample, float* ws_norms, float* ws_xs, float* ws_colc, float* ws_T, unsigned* ws_val, unsigned short* ws_col, int* ws_cnt, int cap, int* flags, void* stream);
int pn_knn_tma(const float* x, int B, int N, int C, int ld, int k, int metric, void* idx_out, int idx_is_i64, float* dist_out, float* ws_norms, void* stream);

/* ---- linear.cu ---- */
/* replaces: Conv1d/Conv2d(k=1) + GroupNorm/BatchNorm + ReLU chains: src/PointNet.py:157-165,194-196,274-284; src/model.py:74-99,155-176 */
int pn_linear_fwd(const float* A, long long lda, const float* W, long long ldw, const float* bias, const float* sbias, const float* in_scale, const float* in_shift, int in_act, float* Y, long long ldy, double* stats, int B, int Np, int K, int Nout, int G, int stats_per_shape, void* stream);
/* replaces: same as pn_linear_fwd, on the tcgen05 tensor cores (split-TF32, fp32-level accuracy); linear_tc.cu.  Needs 16-byte aligned rows, K % 4 == 0, Nout % 32 == 0 and (with stats) channels-per-group % 32 == 0 */
int pn_linear_fwd_tc(const float* A, long long lda, const float* W, long long ldw, const float* bias, const float* sbias, const float* in_scale, const float* in_shift, int in_act, float* Y, long long ldy, double* stats, int B, int Np, int K, int Nout, int G, int stats_per_shape, void* stream);
/* replaces: (query: 1 if pn_linear_fwd_tc accepts the problem, else the caller uses pn_linear_fwd) */
int pn_linear_fwd_tc_supported(const float* A, long long lda, const float* W, long long ldw, const float* Y, long long ldy, int Np, int K, int Nout, int G, int has_stats);
/* replaces: autograd of the same chains (torch built-in in the reference) */
int pn_linear_bwd_data(const float* dY, long long lddy, const float* W, long long ldw, float* dZ, long long lddz, int accumulate, int finalize, const float* A, long long lda, const float* in_scale, const float* in_shift, int in_act, const float* gamma, const float* mean_rstd, double* gsum, int B, int Np, int K, int Nout, int G, int stats_per_shape, void* stream);
/* the same backward on the tcgen05 tensor cores (split-TF32): the forward GEMM kernel with A = dY and the TRANSPOSED weight
   Wt [K][Nout] as its row-major weight operand, finalize epilogue (activation mask + norm-backward sums) in the store phase */
int pn_linear_bwd_data_tc_supported(const float* dY, long long lddy, const float* Wt, long long ldwt, const float* dZ, long long lddz, const float* A, long long lda, int Np, int K, int Nout, int G, int has_gamma);
int pn_linear_bwd_data_tc(const float* dY, long long lddy, const float* Wt, long long ldwt, float* dZ, long long lddz, int accumulate, int finalize, const float* A, long long lda, const float* in_scale, const float* in_shift, int in_act, const float* gamma, const float* mean_rstd, double* gsum, int B, int Np, int K, int Nout, int G, int stats_per_shape, void* stream);
/* replaces: autograd of the same chains (torch built-in in the reference) */
int pn_linear_bwd_weight(const float* dY, long long lddy, const float* A, long long lda, const float* in_scale, const float* in_shift, int in_act, float* dW, long long lddw, float* db, float* dsb, int B, int Np, int K, int Nout, void* stream);
/* the same weight gradient on the tcgen05 tensor cores (split-TF32): D[n][k] += (dY tile)^T (act(norm(A)) tile), the loader warps
   transpose 4 x 4 blocks in registers into the K-major operand layout; rows split over the grid, fp32 atomics into dW */
int pn_linear_bwd_weight_tc_supported(const float* dY, long long lddy, const float* A, long long lda, int Np, int K, int Nout);
int pn_linear_bwd_weight_tc(const float* dY, long long lddy, const float* A, long long lda, const float* in_scale, const float* in_shift, int in_act, float* dW, long long lddw, float* db, float* dsb, int B, int Np, int K, int Nout, void* stream);
/* replaces: nn.GroupNorm / nn.BatchNorm statistics: src/PointNet.py:151-155,166-169; src/model.py:69-73,98-99 */
int pn_norm_finalize(const double* stats, const float* gamma, const float* beta, int S, int G, int C, double count, float eps, float* mean_rstd, float* scale, float* shift, void* stream);
/* replaces: autograd of nn.GroupNorm / nn.BatchNorm */
int pn_norm_bwd_apply(float* dZ, long long lddz, const float* A, long long lda, const float* gamma, const float* mean_rstd, const double* gsum, int B, int Np, int C, int G, int stats_per_shape, double count, float* dgamma, float* dbeta, void* stream);

/* ---- meanshift.cu ---- */
/* replaces: MeanShift.mean_shift_ (one iteration): src/mean_shift.py:58-77 */
int pn_ms_iter_fwd(const float* Y, const float* X, int B, int N, int d, const float* cinv, float* Ynew, float* den, float* unorm, void* stream);
/* replaces: autograd of one mean-shift iteration */
int pn_ms_iter_bwd(const float* gout, const float* Ynew, const float* Yprev, const float* X, const float* den, const float* unorm, int B, int N, int d, const float* cinv, float* ws_Gn, float* ws_gd, float* gYprev, float* gX, int accumulate_gX, void* stream);
/* replaces: MeanShift.compute_bandwidth: src/mean_shift.py:130-135 (2 - 2 X X^T, topk(K, largest=False)[:, -1]) */
int pn_ms_kth_dist(const float* X, const int* rows, int B, int S, long long shape_stride, int d, int K, float* kth, void* stream);
/* replaces: MeanShift.nms: src/mean_shift.py:146-149 (mode 0), :163-171 (mode 1), :177-178 (mode 2) */
int pn_ms_argsel(int mode, const float* A, long long a_stride, int Ma, const float* Bm, long long b_stride, int Nb, int B, int d, const float* cnt, const float* thr, int* out, void* stream);

/* Default in Evaluation.fitting_loss since round 2 (PN_MS_SPARSE_BWD=0: dense backward): backward of one mean-shift iteration restricted to a compact
   set of R = 64 rows per shape.  In Evaluation.fitting_loss the loss sees the shifted points only through the <= 49 cluster
   centres (src/mean_shift.py:41 `center = new_X[indices]`, src/residual_utils.py:118), row i of Y_t depends on row i of
   Y_{t-1} alone, so every other row of the autograd pass of src/mean_shift.py:58-77 contributes exactly zero.
   gout / Ynew_R / Yprev_R [B][64][d], den_R / unorm_R [B][64], ws_part [B][ceil(N/64)][64][d]; gX is accumulated. */
int pn_ms_rows_bwd(const float* gout, const float* Ynew_R, const float* Yprev_R, const float* den_R, const float* unorm_R, const float* X, int B, int R, int N, int d, const float* cinv, float* ws_Gn, float* ws_gd, float* ws_part, float* gYprev_R, float* gX, void* stream);

/* ---- meanshift_tc.cu / meanshift_tc_bwd.cu / meanshift_tc_kth.cu (tcgen05 split-TF32 versions, d must be 128) ---- */
/* replaces: MeanShift.mean_shift_ (one iteration): src/mean_shift.py:58-77 — same contract as pn_ms_iter_fwd */
int pn_ms_iter_fwd_tc(const float* Y, const float* X, int B, int N, int d, const float* cinv, float* Ynew, float* den, float* unorm, void* stream);
/* replaces: autograd of one mean-shift iteration — same contract as pn_ms_iter_bwd */
int pn_ms_iter_bwd_tc(const float* gout, const float* Ynew, const float* Yprev, const float* X, const float* den, const float* unorm, int B, int N, int d, const float* cinv, float* ws_Gn, float* ws_gd, float* gYprev, float* gX, int accumulate_gX, void* stream);
/* replaces: MeanShift.compute_bandwidth: src/mean_shift.py:130-135 — same contract as pn_ms_kth_dist */
int pn_ms_kth_dist_tc(const float* X, const int* rows, int B, int S, long long shape_stride, int d, int K, float* kth, void* stream);
/* the same selection for the 128-row blocks that contain a row with flags[b][row] != 0 only (other blocks exit at once, their
   kth entries stay untouched): exact fall-back of pn_ms_kth_dist_tma */
int pn_ms_kth_dist_tc_flagged(const float* X, const int* rows, int B, int S, long long shape_stride, int d, int K, const int* flags, float* kth, void* stream);
/* replaces: MeanShift.compute_bandwidth: src/mean_shift.py:115-137 with num_samples >= N (all rows against all N points of
   their shape) — ONE pass over the N x N distance tiles instead of the four radix passes of pn_ms_kth_dist_tc: the b_sample-th
   smallest distance to the column sample {0, stride, 2 stride, ...} brackets the K-th smallest from above, one TMA-fed tcgen05
   pass appends every distance below the bracket to per-row lists, a warp per row selects the K-th and recomputes that pair in
   fp32.  Xs = X - tf32_hi(X) (pn_ms_prepare_operands).  ws_key [B*N][cap] u32, ws_col [B*N][cap] u16, ws_cnt [B*N][2] i32
   (one list per half of a tile's columns), ws_hi [B*N] f32, cap = 1024.  flags [B*N] is written: 1 = bracket failed for that row, kth not written (run
   pn_ms_kth_dist_tc_flagged with the same flags next; no host round trip); meanshift_tma.cu */
int pn_ms_kth_dist_tma(const float* X, const float* Xs, int B, int N, int d, int K, int stride, int b_sample, unsigned* ws_key, unsigned short* ws_col, int* ws_cnt, float* ws_hi, int cap, int* flags, float* kth, void* stream);
/* replaces: MeanShift.nms arg-selects: src/mean_shift.py:146-149 (mode 0), :163-171 (mode 1) — same contract as pn_ms_argsel for modes 0 and 1; meanshift_tc_argsel.cu */
int pn_ms_argsel_tc(int mode, const float* A, long long a_stride, int Ma, const float* Bm, long long b_stride, int Nb, int B, int d, const float* cnt, const float* thr, int* out, void* stream);
/* the same arg-selects with the column tiles fetched by TMA (64-column tiles, N = 64 MMAs); Bm contiguous [B][Nb][128], 16-byte
   aligned; ws_Bms [B][Nb][128] receives its small split part.  Identical picks; meanshift_tma.cu */
int pn_ms_argsel_tma_supported(const float* Bm, long long b_stride, int Nb, int d);
int pn_ms_argsel_tma(int mode, const float* A, long long a_stride, int Ma, const float* Bm, long long b_stride, int Nb, int B, int d, const float* cnt, const float* thr, float* ws_Bms, int* out, void* stream);
/* the two halves of pn_ms_iter_bwd_tc on their own: prep pass (Gn, gd) and the cols kernel (gX) */
int pn_ms_bwd_prep_tc(const float* gout, const float* Ynew, const float* den, const float* unorm, int B, int N, int d, float* ws_Gn, float* ws_gd, void* stream);
int pn_ms_bwd_cols_tc(const float* Yprev, const float* X, int B, int N, int d, const float* cinv, const float* ws_Gn, const float* ws_gd, float* gX, int accumulate_gX, void* stream);

/* ---- meanshift_tma.cu (default since round 2; PN_MS_TMA=0 selects the loader-warp kernels: streamed operand tiles fetched by TMA) ---- */
/* once per MeanShift.mean_shift_ call (X is constant over the iterations, src/mean_shift.py:58-77):
   Xs = X - tf32_hi(X) [B][N][128], Xt / Xst = transposes [B][128][Np], Np = N rounded up to a multiple of 32 */
int pn_ms_prepare_operands(const float* X, int B, int N, int d, int Np, float* Xs, float* Xt, float* Xst, void* stream);
/* replaces: MeanShift.mean_shift_ (one iteration): src/mean_shift.py:58-77 — contract of pn_ms_iter_fwd_tc + operand forms */
int pn_ms_iter_fwd_tma(const float* Y, const float* X, const float* Xs, const float* Xt, const float* Xst, int B, int N, int d, int Np, const float* cinv, float* Ynew, float* den, float* unorm, void* stream);
/* replaces: autograd of one mean-shift iteration — contract of pn_ms_iter_bwd_tc + operand forms of X + workspace ws_C
   (4 * B * 2 Nq * 128 floats, Nq = N rounded up to 16) for the forms of the interleaved [Yprev; Gn] tiles */
int pn_ms_iter_bwd_tma(const float* gout, const float* Ynew, const float* Yprev, const float* X, const float* Xs, const float* Xt, const float* Xst, const float* den, const float* unorm, int B, int N, int d, int Np, const float* cinv, float* ws_Gn, float* ws_gd, float* ws_C, float* gYprev, float* gX, int accumulate_gX, void* stream);
/* replaces: (debug aid, no reference counterpart: host-mapped progress words written by the tcgen05 pipelines; NULL disables) */
int pn_debug_set_progress(int* host_mapped_words);

/* ---- pointwise.cu ---- */
/* replaces: F.relu(bnmlp1(mlp1(x))).max(dim=2): src/PointNet.py:194-196; x*weights + adaptive_max_pool1d: src/model.py:165-169 */
int pn_colmax_norm(const float* Y, long long ldy, int B, int N, int C, const float* scale, const float* shift, int act, const float* wts, float* out, int* arg, void* stream);
/* replaces: autograd of the above */
int pn_colmax_bwd_fill(const float* Y, long long ldy, const float* gt, const int* arg, const float* gamma, const float* mean_rstd, const double* gsum, int B, int N, int C, int G, int stats_per_shape, double count, int dense, float* dY, long long lddy, void* stream);
/* replaces: LogSoftmax(dim=1): src/PointNet.py:284 */
int pn_logsoftmax_fwd(const float* logits, long long ldl, int B, int N, int P, float* logp, void* stream);
/* replaces: autograd of LogSoftmax */
int pn_logsoftmax_bwd(const float* logp, const float* dlp, int B, int N, int P, float* dlogits, long long ldd, void* stream);
/* replaces: primitive_loss = NLLLoss: src/segment_loss.py:151 */
int pn_nll_fwd(const float* logp, const long long* target, int B, int N, int P, float* loss_zeroed, void* stream);
/* replaces: autograd of NLLLoss */
int pn_nll_bwd(const long long* target, const float* gout, int B, int N, int P, float* dlp_zeroed, void* stream);
/* replaces: torch.nn.functional.normalize: src/segment_loss.py:45, src/residual_utils.py:108 */
int pn_l2norm_fwd(const float* x, long long ldx, long long rows, int D, float eps, float* y, long long ldy, float* norms, void* stream);
/* replaces: autograd of normalize */
int pn_l2norm_bwd(const float* y, long long ldy, const float* dy, long long lddy, const float* norms, long long rows, int D, float* dx, long long lddx, int accumulate, void* stream);
/* replaces: EmbeddingLoss.triplet_loss inner loop: src/segment_loss.py:98-119 */
int pn_triplet_fwd(const float* E, long long lde, int D, const int* a_idx, const int* n_idx, int T, int S, float margin, float* pair_loss, float* pair_sat, void* stream);
/* replaces: autograd of the triplet loss */
int pn_triplet_bwd(const float* E, long long lde, int D, const int* a_idx, const int* n_idx, int T, int S, float margin, const float* pair_sat, const float* pair_w, float* dE, long long ldde, void* stream);

/* ---- primitives.cu ---- */
/* replaces: ComputePrimitiveDistance.distance_from_{plane,sphere,cylinder,cone}: src/primitives.py:100-195 */
int pn_residual_fwd(const float* P, const int* seg, int N, const int* type, const float* par, int S, float* sumf_zeroed, float* jac_zeroed, float* cnt_zeroed, void* stream);
/* all shapes of a step in one launch: P [B][N][3], seg [B][N] = slot of the point's segment or -1, tables [B][S] by slot */
int pn_residual_fwd_batched(const float* P, const int* seg, int B, int N, const int* type, const float* par, int S, float* sumf_zeroed, float* jac_zeroed, float* cnt_zeroed, void* stream);

/* ---- small3.cu ---- */
/* replaces: CustomSVD (torch.svd of the weighted (m,3) matrix, via the eigen-decomposition of its 3x3 Gram matrix): src/fitting_utils.py:420-455; torch.eig in pca_torch :585 */
int pn_sym3_eigh(const double* G, int S, double* w_asc, double* V, void* stream);
/* replaces: LeastSquares.lstsq + best_lambda (matrix_rank / qr / inverse, Tikhonov retry): src/fitting_utils.py:36-85 */
int pn_lstsq3(const double* AtA, const double* AtY, int S, int rows, double eps32, double* x, double* Minv, double* lam, void* stream);

/* ---- spline.cu ---- */
/* replaces: sample_points_from_control_points_: src/fitting_utils.py:609-622; src/loss.py:161-165,179-183 */
int pn_spline_eval_fwd(const float* Nu, const float* Nv, const float* P, int B, int gu, int gv, int cu, int cv, float* out, void* stream);
/* replaces: autograd of Nu P Nv^T */
int pn_spline_eval_bwd(const float* Nu, const float* Nv, const float* g, int B, int gu, int gv, int cu, int cv, float* dP, void* stream);
/* float64 instances of the same tensor-product kernels.  Run with the pseudo-inverse basis matrices Nu^+ (cu x gu), Nv^+ they
   replace the gridded control-point solve approximation.fit_bezier_surface: src/approximation.py:308-334 (numpy float64) */
int pn_spline_eval_fwd_f64(const double* Nu, const double* Nv, const double* P, int B, int gu, int gv, int cu, int cv, double* out, void* stream);
int pn_spline_eval_bwd_f64(const double* Nu, const double* Nv, const double* g, int B, int gu, int gv, int cu, int cv, double* dP, void* stream);

/* ---- kronfit.cu (post-fit control-point optimisation of the inference path) ---- */
/* replaces: fit_bezier_surface_fit_kronecker: src/approximation.py:338-364 (host Kronecker matrix + numpy lstsq per coordinate).
   U [S][M][n], V [S][M][m] per-sample basis rows, P [S][M][3] samples (float64) -> ctrl [S][n][m][3] least-squares control
   points: normal equations assembled and factored (Cholesky) in shared memory, one CTA per surface, n m <= 128.
   flag [S]: 1 = rank-deficient sampling (pivot below 1e-12 of the largest diagonal entry), ctrl[s] not written */
int pn_kron_fit(const double* U, const double* V, const double* P, int S, int M, int n, int m, double* ctrl, int* flag, void* stream);
/* replaces: geomdl surface evaluation at scattered parameters: src/primitive_forward.py:186,258 (evaluate_list) —
   out [S][M][3] = sum_ab U[s][i][a] V[s][i][b] C[s][a][b][:]; C advances by c_stride doubles per surface (0 = shared) */
int pn_kron_eval(const double* U, const double* V, const double* C, long long c_stride, int S, int M, int n, int m, double* out, void* stream);

/* ---- assign.cu (segment matching on the device: SURVEY a32 / 8f-2) ---- */
/* replaces: relaxed_iou_fast(to_one_hot(pred), to_one_hot(gt)) and `1 - cost`: src/segment_utils.py:356-373, src/fitting_utils.py:362-371.
   pred, gt [B][N] int32 labels in [0, K), K <= 64 -> cost [B][K][K] = 1 - inter / (|p| + |g| - inter + 1e-7) (the reference's fp32
   expression; counts are exact).  *bad is set to 1 when a label lies outside [0, K) */
int pn_iou_cost(const int* pred, const int* gt, int B, int N, int K, float* cost, int* bad, void* stream);
/* replaces: lapsolver.solve_dense(cost): src/fitting_utils.py:372, src/segment_utils.py:173 — optimal assignment of the n x n cost
   matrices (n <= 64), one warp per matrix (Kuhn-Munkres with potentials in float64); col_of_row [B][n] (rows are 0..n-1 in order) */
int pn_hungarian(const float* cost, int B, int n, int* col_of_row, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PARSENET_B200_H */
