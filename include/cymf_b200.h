/* cymf_b200.h -- C ABI of the B200-native factor-update hot path of CyMF.
 *
 * This is the drop-in boundary.  The reference (minatosato/cymf) has no FFI layer of its own: its hot
 * loops are the bodies of four typed Cython methods.  Each group of entry points below replaces one of
 * them and cites it (paths relative to the reference tree):
 *
 *   cymf_rng_*          <- cymf/math.pyx:12-18      UniformGenerator (std::mt19937 + uniform_int_distribution)
 *   cymf_bpr_*          <- cymf/bpr.pyx:117-171     BPR._fit_bpr      (+ cymf/model.pyx:47-87, cymf/optimizer.pyx)
 *   cymf_relmf_*        <- cymf/relmf.pyx:107-148   RelMF._fit_relmf  (+ cymf/model.pyx:99-142, cymf/optimizer.pyx)
 *   cymf_glove_*        <- cymf/glove.pyx:117-156   GloVe._fit_glove  (+ cymf/model.pyx:166-204, optimizer.pyx:85-123)
 *   cymf_als_*, cymf_gram_* <- cymf/wmf.pyx:136-174 WMF._als          (+ cymf/linalg.pyx:144-163 solvep)
 *   cymf_cooc_*         <- cymf/glove.pyx:183-241   read_text (co-occurrence counting loop, :218-221)
 *   cymf_eval_*         <- cymf/evaluator.pyx:57-139 Evaluator.evaluate (+ cymf/metrics.pyx:24-125)
 *
 * Conventions
 *   - Plain C types only.  `*_dev` functions take DEVICE pointers and a `cudaStream_t` passed as void*
 *     (NULL = default stream); they enqueue work and return without synchronising.  `*_host` functions
 *     take HOST pointers, own their device memory for the duration of the call, and return after the
 *     results are back in the caller's buffers -- they are what a cgo / JNI / ctypes binding of the
 *     reference method would call.
 *   - Return value: 0 on success, a positive cudaError_t value, or a negative CYMF_E* code.
 *     cymf_last_error() returns a thread-local message for the last non-zero status.  Nothing aborts.
 *   - Factor matrices are row-major with a row stride `ld` (elements) that must be a multiple of 4 and
 *     >= K; pad columns must be zero (they stay zero under every update rule here).  Element type is
 *     selected by `dtype`: CYMF_F32 or CYMF_F64.  The reference computes in f64.
 *   - CSR `indptr` is int64 on the device (N up to 1e9 for config C5), `indices` int32, rows sorted.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point returns an error.
 */
#ifndef CYMF_B200_H
#define CYMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYMF_ABI_VERSION 1

enum { CYMF_F32 = 0, CYMF_F64 = 1 };
enum { CYMF_SGD = 0, CYMF_ADAGRAD = 1, CYMF_ADAM = 2 };
enum { CYMF_EINVAL = -1, CYMF_ENOMEM = -2, CYMF_EUNSUPPORTED = -3 };

int cymf_abi_version(void);
const char *cymf_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches counter) */
int64_t cymf_launch_count(void);

/* ---- host RNG: the reference's negative / candidate stream (cymf/math.pyx:12-18) ------------------- */
typedef struct cymf_rng cymf_rng;
cymf_rng *cymf_rng_create(uint32_t seed);
void cymf_rng_destroy(cymf_rng *g);
/* `count` draws of uniform_int_distribution<long>(0, n-1)(mt19937), libstdc++ >= 11 semantics */
int cymf_rng_fill_below(cymf_rng *g, uint32_t n, int32_t *out, int64_t count);
/* the same distribution for any n >= 1, including ranges wider than the 32-bit engine (RelMF draws cells from
 * [0, U*I), cymf/relmf.pyx:127: 3.7e9 at the ml-20m shape) */
int cymf_rng_fill_below64(cymf_rng *g, uint64_t n, int64_t *out, int64_t count);

/* ---- layout helpers (device pointers) ------------------------------------------------------------------
 * The reference keeps W/H as dense f64 [rows, K] NumPy arrays (cymf/bpr.pyx:99-101).  On the device a
 * factor matrix is [rows, ld] of `dtype` with ld = K rounded up to 4; these convert between the two
 * (src/dst of the f64 side are dense DEVICE staging buffers) and fill optimizer state. */
int cymf_pack_rows_dev(const double *src, void *dst, int dtype, int64_t rows, int32_t K, int32_t ld, void *stream);
int cymf_unpack_rows_dev(const void *src, double *dst, int dtype, int64_t rows, int32_t K, int32_t ld, void *stream);
int cymf_fill_dev(void *dst, int dtype, int64_t n, double value, void *stream);
/* dense f64 device vector -> device vector of `dtype` (propensities, X.data) */
int cymf_convert_dev(const double *src, void *dst, int dtype, int64_t n, void *stream);

/* ---- BPR (cymf/bpr.pyx:160-171) -------------------------------------------------------------------- */
/* Optimizer state: AdaGrad uses s1 (accumulators, init 1); Adam uses s1 = M, s2 = V (init 0); SGD none. */
typedef struct {
    void *W;            /* [U, ld]  user factors                      */
    void *H;            /* [I, ld]  item factors                      */
    void *s1W, *s1H;    /* optimizer state, same shape / dtype, or NULL */
    void *s2W, *s2H;
} cymf_factors;

/* One Hogwild epoch over the N shuffled (user, positive) pairs (bpr.pyx:162-169): one lane group per
 * triplet, negative j ~ U[0, I) from Philox4x32-10 keyed by (seed, epoch, l), skipped (not resampled)
 * when j is a positive of u (bpr.pyx:166-167), lock-free read-modify-write of the three rows.
 * `scatter`: 0 = plain vector stores (races lose updates, as in the reference), 1 = the parameter step and
 * the optimizer-state increments (AdaGrad: g^2; Adam: the m and v deltas) are applied with red.global.add, so
 * no concurrent triplet of a row is lost.  `max_inflight` > 0 caps the number of triplets processed
 * concurrently (bounds Hogwild staleness on small matrices; 0 = fill the machine).
 * `applied` (device uint64, may be NULL) += accepted triplets. */
int cymf_bpr_hogwild_epoch_dev(const cymf_factors *f, int dtype, int optimizer, int scatter,
                               const int32_t *users, const int32_t *positives, int64_t N,
                               const int64_t *indptr, const int32_t *indices,
                               int32_t U, int32_t I, int32_t K, int32_t ld,
                               double learning_rate, double weight_decay,
                               uint64_t seed, uint32_t epoch, int64_t max_inflight,
                               unsigned long long *applied, void *stream);

/* The same pass over a RANGE of the epoch's pair list: users / positives point at pair `first`, N pairs follow; the
 * negative of pair l is still keyed by (seed, epoch, first + l), so an epoch issued as several ranges -- e.g. while
 * later ranges are still in flight over PCIe -- draws exactly the negatives of the single-launch epoch. */
int cymf_bpr_hogwild_range_dev(const cymf_factors *f, int dtype, int optimizer, int scatter,
                               const int32_t *users, const int32_t *positives, int64_t N,
                               const int64_t *indptr, const int32_t *indices,
                               int32_t U, int32_t I, int32_t K, int32_t ld,
                               double learning_rate, double weight_decay,
                               uint64_t seed, uint32_t epoch, int64_t max_inflight,
                               unsigned long long *applied, int64_t first, void *stream);

/* The negative the Hogwild kernels draw for triplets l = first .. first+count-1 of `epoch` (host-side
 * evaluation of the same Philox4x32-10 code), so that a run can be audited against the CPU oracle. */
int cymf_bpr_negatives_host(uint64_t seed, uint32_t epoch, int64_t first, int64_t count, uint32_t n, int32_t *out);

/* Serialized f64 replay of one epoch with a caller-supplied negative stream (device int32[N], e.g. from
 * cymf_rng_fill_below): triplets are applied strictly in order l = 0..N-1 with the reference's own
 * operation order (sequential-k dot product, no FMA contraction), i.e. num_threads = 1 semantics. */
int cymf_bpr_replay_epoch_dev(const cymf_factors *f, int optimizer,
                              const int32_t *users, const int32_t *positives, const int32_t *negatives,
                              int64_t N, const int64_t *indptr, const int32_t *indices,
                              int32_t U, int32_t I, int32_t K, int32_t ld,
                              double learning_rate, double weight_decay,
                              unsigned long long *applied, void *stream);

/* Host-buffer form of BPR._fit_bpr(users, positives, X, num_epochs, learning_rate, weight_decay, ...):
 * W [U,K], H [I,K] are dense row-major f64 HOST arrays updated in place; X is given as host CSR
 * (int32 indptr / indices, as scipy hands them over).  mode: 0 = Hogwild f32, 1 = Hogwild f64,
 * 2 = serialized f64 replay of the reference stream (mt19937 seed `seed`).  `applied_out` (may be NULL)
 * receives the number of accepted triplets summed over epochs. */
int cymf_bpr_fit_host(double *W, double *H, int32_t U, int32_t I, int32_t K,
                      const int32_t *users, const int32_t *positives, int64_t N,
                      const int32_t *indptr, const int32_t *indices,
                      int32_t num_epochs, double learning_rate, double weight_decay,
                      int optimizer, int mode, uint64_t seed, int64_t *applied_out);

/* ---- RelMF (cymf/relmf.pyx:143-148, cymf/model.pyx:99-142) ----------------------------------------------- */
/* One Hogwild epoch of `n_samples` uniformly drawn cells (the reference draws U*I per epoch, relmf.pyx:121,143):
 * sample l touches (u, i) ~ U[0,U) x U[0,I) from Philox4x32-10 keyed by (seed, epoch, l) -- the distribution of
 * the reference's r ~ U[0, U*I), u = r / I, i = r % I.  The label X[u,i] is looked up in the CSR (`values` in
 * CSR order with the element type of `dtype`, NULL = every stored cell is 1; absent cell = 0), the reference
 * reads its densified copy (relmf.pyx:79-80).  propensities: [I] of `dtype` (relmf.pyx:90).  `factors` and
 * `max_inflight` as for cymf_bpr_hogwild_epoch_dev.  `scatter` = 1 applies the parameter step and the
 * optimizer-state increments (AdaGrad: g^2; Adam: the m and v deltas) with red.global.add for every optimizer,
 * so no concurrent sample of a row is lost and Adam's (m, v) pair cannot be torn; 0 = plain stores. */
int cymf_relmf_hogwild_epoch_dev(const cymf_factors *f, int dtype, int optimizer, int scatter,
                                 const int64_t *indptr, const int32_t *indices, const void *values,
                                 const void *propensities, int32_t U, int32_t I, int32_t K, int32_t ld,
                                 int64_t n_samples, double learning_rate, double weight_decay, double clip_value,
                                 uint64_t seed, uint32_t epoch, int64_t max_inflight, void *stream);

/* The cells r = u * I + i the Hogwild kernel draws for samples first .. first+count-1 of `epoch` (host-side
 * evaluation of the same code), so that a run can be audited against the CPU oracle. */
int cymf_relmf_cells_host(uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t U, int32_t I,
                          int64_t *out);

/* Serialized f64 replay of one epoch with a caller-supplied cell stream (device int64[n_samples], r = u*I + i,
 * e.g. from cymf_rng_fill_below64): samples applied strictly in order with the reference's operation order. */
int cymf_relmf_replay_epoch_dev(const cymf_factors *f, int optimizer, const int64_t *cells, int64_t n_samples,
                                const int64_t *indptr, const int32_t *indices, const double *values,
                                const double *propensities, int32_t U, int32_t I, int32_t K, int32_t ld,
                                double learning_rate, double weight_decay, double clip_value, void *stream);

/* Host-buffer form of RelMF._fit_relmf(X, propensities, num_epochs, ...) (relmf.pyx:107-113) with X handed over as
 * host CSR (int32 indptr / indices, f64 values or NULL) instead of the reference's dense U x I array; W, H dense
 * f64 HOST arrays updated in place; U*I samples per epoch; mode as in cymf_bpr_fit_host. */
int cymf_relmf_fit_host(double *W, double *H, int32_t U, int32_t I, int32_t K,
                        const int32_t *indptr, const int32_t *indices, const double *values,
                        const double *propensities, int32_t num_epochs, double learning_rate,
                        double weight_decay, double clip_value, int optimizer, int mode, uint64_t seed);

/* ---- GloVe (cymf/glove.pyx:149-156, cymf/model.pyx:166-204, cymf/optimizer.pyx:85-123) ---------------- */
typedef struct {
    void *W, *H;        /* [Vw, ld] central and [Vh, ld] context word vectors                       */
    void *bW, *bH;      /* [Vw], [Vh] biases                                                        */
    void *aW, *aH;      /* AdaGrad accumulators of W, H (same shapes; the reference starts them at 1) */
    void *abW, *abH;    /* AdaGrad accumulators of the biases                                       */
} cymf_glove_params;

/* One Hogwild pass over the N shuffled (central, context, count) samples (glove.pyx:151-153).  `counts` has
 * the element type selected by `dtype`.  scatter: 0 = vector stores, 1 = red.global.add of the AdaGrad
 * increments and steps.  The reference's K-fold bias update (model.pyx:199-204) is kept.
 * `loss_sum` (device double, may be NULL) += sum_l 0.5 f(n) d^2 (model.pyx:180). */
int cymf_glove_hogwild_epoch_dev(const cymf_glove_params *p, int dtype, int scatter,
                                 const int32_t *central, const int32_t *context, const void *counts,
                                 int64_t N, int32_t K, int32_t ld, double learning_rate, double x_max,
                                 double alpha, int64_t max_inflight, double *loss_sum, void *stream);

/* Serialized f64 replay, samples applied in order with the reference's operation order.
 * `loss` (device double[N], may be NULL) receives loss[l] (glove.pyx:152). */
int cymf_glove_replay_epoch_dev(const cymf_glove_params *p, const int32_t *central, const int32_t *context,
                                const double *counts, int64_t N, int32_t K, int32_t ld,
                                double learning_rate, double x_max, double alpha, double *loss, void *stream);

/* Host-buffer form of GloVe._fit_glove(central_words, context_words, counts, central_W, central_bias,
 * context_W, context_bias, num_epochs, learning_rate, x_max, alpha, ...): dense f64 HOST arrays updated in
 * place.  mode as in cymf_bpr_fit_host.  loss_out (num_epochs, may be NULL) = mean loss per epoch. */
int cymf_glove_fit_host(const int32_t *central, const int32_t *context, const double *counts, int64_t N,
                        double *central_W, double *central_bias, double *context_W, double *context_bias,
                        int64_t Vw, int64_t Vh, int32_t K, int32_t num_epochs,
                        double learning_rate, double x_max, double alpha, int mode, double *loss_out);

/* ---- sparse-matrix preparation on the device -------------------------------------------
 * What the reference does on the host per epoch / per fit: `X.T.tocsr()` twice per epoch (cymf/wmf.pyx:112) and the
 * per-user positive sets (cymf/bpr.pyx:146-147; the sorted CSR row is the set here), plus the nnz-balanced row deal
 * of the sharded ALS.  All results are exact and deterministic.  Every function takes a caller-provided device
 * `workspace` of at least cymf_*_workspace_bytes(...) bytes. */
int64_t cymf_scan_workspace_bytes(int64_t n);
/* out[0..n] = exclusive prefix sums of in[0..n-1] (out[n] = total) */
int cymf_exclusive_scan_u32_dev(const uint32_t *in, int64_t *out, int64_t n, void *workspace, void *stream);
int64_t cymf_sort_workspace_bytes(int64_t n);
/* stable LSD radix sort of (key, value) pairs by the low `key_bits` bits of the key, in place; values may be NULL */
int cymf_sort_pairs_dev(uint32_t *keys, uint32_t *values, int64_t n, int32_t key_bits, void *workspace, void *stream);
int64_t cymf_csr_transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz);
/* CSR of X^T (t_indptr int64[cols+1], t_indices int32[nnz], rows sorted) from the CSR of X: scipy's X.T.tocsr() */
int cymf_csr_transpose_dev(const int64_t *indptr, const int32_t *indices, int64_t rows, int64_t cols, int64_t nnz,
                           int64_t *t_indptr, int32_t *t_indices, void *workspace, void *stream);
int64_t cymf_deal_workspace_bytes(int64_t rows);
/* Heaviest-first round-robin deal of rows over `world` ranks: rows by decreasing degree (ties: ascending id), the
 * q-th goes to rank q % world at position q / world.  slot_row[world*per_rank]: slot -> row id (-1 = phantom);
 * row_slot[rows]: row id -> slot.  per_rank >= ceil(rows / world). */
int cymf_deal_rows_dev(const int64_t *indptr, int64_t rows, int32_t world, int64_t per_rank,
                       int64_t *slot_row, int64_t *row_slot, void *workspace, void *stream);
int64_t cymf_csr_block_workspace_bytes(int64_t count);
/* Row block in dealt order: block row q = source row slot_row[q] (empty when -1) with columns renamed through
 * col_slot (NULL = keep).  Call once with blk_indices == NULL to get blk_indptr[count+1] (last entry = block nnz),
 * then again with blk_indices allocated to copy the entries. */
int cymf_csr_block_dev(const int64_t *indptr, const int32_t *indices, const int64_t *slot_row, int64_t count,
                       const int64_t *col_slot, int64_t *blk_indptr, int32_t *blk_indices, void *workspace,
                       void *stream);

/* ---- GloVe.save_word2vec_format (cymf/glove.pyx:164-177) ------------------------------------------------------
 * Writes "<V> <K>\n" and one line "<word> <w_0> ... <w_{K-1}>\n" per row, every value formatted exactly as the
 * reference's str(np.float64) (shortest round-trip digits, CPython repr layout), so the file is byte-identical.
 * W: dense f64 HOST [V, K]; words: V NUL-terminated, already encoded byte strings. */
int cymf_word2vec_write_host(const char *path, const double *W, int64_t V, int32_t K, const char *const *words);

/* ---- co-occurrence counting of read_text (cymf/glove.pyx:183-241, loop at :218-221) --------
 * For every kept token j of a line and every earlier kept token k of the same line with j - k <= window:
 * M[x_j, x_k] += 1.0 / (j - k).  tokens[n_tokens] = kept-word ids, lines concatenated; pos_in_line[n_tokens] = index
 * of the token inside its line (both DEVICE int32).  Cells come out sorted by (row, col) in rows / cols / vals
 * (`capacity` entries each; n_tokens * window always suffices), *nnz_out (DEVICE int64) = number of cells.  Every
 * cell is summed in f64 in the reference's own order, so the counts are bit-identical to its unordered_map's.
 * n_tokens * window must stay below 2^32. */
int64_t cymf_cooc_workspace_bytes(int64_t n_tokens, int32_t window);
int cymf_cooc_count_dev(const int32_t *tokens, const int32_t *pos_in_line, int64_t n_tokens, int32_t vocab,
                        int32_t window, int32_t *rows, int32_t *cols, double *vals, int64_t capacity,
                        int64_t *nnz_out, void *workspace, void *stream);

/* ---- WMF ALS (cymf/wmf.pyx:136-174, cymf/linalg.pyx:144-163) ------------------------------------------- */
/* Size limits of this section: K = num_components <= 256 for cymf_gram_*, cymf_als_cg_dev and cymf_als_half_host
 * (the reference has no limit, wmf.pyx:44); the Cholesky change of variables (cymf_chol_transforms_dev,
 * cymf_rows_times_matrix_*), cymf_spd_inverse_dev and the tensor-core kernels (cymf_als_rows_tc_dev,
 * cymf_als_heavy_rows_dev) stop at K <= 128 -- above that the row systems are solved by plain CG (G passed to
 * cymf_als_cg_dev). */
/* G = Y^T Y (+ weight_decay * I when add_weight_decay != 0), wmf.pyx:142-143.  Y is [n, ld] of `dtype`.
 * `workspace` needs cymf_gram_workspace_doubles(n, K) doubles (per-slab partials, reduced in a fixed order so
 * the result is deterministic).  out_f64 (dense K*K doubles) and/or out_native ([ld, ld] of `dtype`, zero
 * padded -- the layout cymf_als_cg_dev reads) receive the result.
 * Multi-GPU: each rank calls this on its own row block with add_weight_decay = 0, all-reduces out_f64 and
 * finishes with cymf_gram_finalize_dev. */
int64_t cymf_gram_workspace_doubles(int64_t n, int32_t K);
int cymf_gram_dev(const void *Y, int dtype, int64_t n, int32_t K, int32_t ld, double weight_decay,
                  int add_weight_decay, double *workspace, int64_t workspace_doubles,
                  double *out_f64, void *out_native, void *stream);
int cymf_gram_finalize_dev(const double *in_f64, int dtype, int32_t K, int32_t ld, double weight_decay,
                           void *out_native, void *stream);

/* Multi-GPU Gram all-reduce without NCCL: out_f64 [K,K] = sum over the n_parts (1..8) dense [K,K] f64 partials
 * (host array of device pointers; peers' symmetric buffers over NVLink) in index order, + weight_decay on the diagonal
 * (cymf/wmf.pyx:142-143 for a factor matrix whose rows are sharded over ranks).  The caller orders the ranks' writes
 * before this kernel with a cross-rank barrier. */
int cymf_gram_sum_dev(const double *const *parts, int32_t n_parts, int32_t K, double weight_decay, double *out_f64,
                      void *stream);

/* Solves the rows listed in `order` (n_solve row ids, heaviest first) of
 *     (G + (weight-1) sum_{i in row} y_i y_i^T) x = weight sum_{i in row} y_i            (wmf.pyx:158-168)
 * by conjugate gradient, warm-started from the current content of X, until |residual| <= cg_tol * |rhs| or
 * cg_max_iter iterations; rows without entries are zeroed (wmf.pyx:154-156).  The K x K matrix the reference
 * builds and LU-factorises is never formed.  stage_rows: item vectors per row kept in shared memory
 * (0 = auto).  queue: device int32 scratch (work-queue head).  stats (device uint64[2], may be NULL):
 * [0] += CG iterations over all rows, [1] += rows that stopped at cg_max_iter (the reference drops dgesv's
 * `info`; this is the equivalent health signal). */
int cymf_als_cg_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve,
                    void *X, const void *Y, const void *G, const void *Ginv, int dtype, int32_t K, int32_t ld,
                    double weight, double cg_tol, int32_t cg_max_iter, int32_t warps_per_row,
                    int32_t stage_rows, int32_t *queue, unsigned long long *stats, void *stream);

/* Ginv of cymf_als_cg_dev (may be NULL): [ld, ld] inverse of G, used as the CG preconditioner.  The
 * preconditioned operator is I + (w-1) G^-1 Y_r^T Y_r, whose spectrum is clustered near 1, so rows converge in
 * roughly half the iterations.  cymf_spd_inverse_dev computes it on the device from the dense f64 K x K matrix
 * A (+ add_diag on the diagonal) by in-place Gauss-Jordan elimination in f64. */
int cymf_spd_inverse_dev(const double *A, int32_t K, int32_t ld, double add_diag, int dtype, void *out_native,
                         void *stream);

/* G of cymf_als_cg_dev may be NULL: then the factors are taken to be in the coordinates y~ = L^-1 y,
 * x~ = L^T x (G = L L^T), where every row system reads (I + (w-1) sum y~ y~^T) x~ = w sum y~ and the CG
 * iteration has no dense K x K product at all (same Krylov space as G^-1-preconditioned CG).
 * cymf_chol_transforms_dev factorises the dense f64 K x K matrix A (+ add_diag on the diagonal) in f64 on the
 * device and writes the three [ld, ld] zero-padded matrices of `dtype` for cymf_rows_times_matrix_dev:
 *     Y~ = Y * By (By = L^-T),   X~ = X * Bfwd (Bfwd = L),   X = X~ * Bbwd (Bbwd = L^-1).
 * info (device int32, may be NULL) is set to k+1 if pivot k is not positive.
 * cymf_rows_times_matrix_dev: out[r,:] = in[r,:] * B for [rows, ld] matrices (in place allowed). */
int cymf_chol_transforms_dev(const double *A, int32_t K, int32_t ld, double add_diag, int dtype,
                             void *By, void *Bfwd, void *Bbwd, int32_t *info, void *stream);
int cymf_rows_times_matrix_dev(const void *in, void *out, const void *B, int dtype, int64_t rows, int32_t ld,
                               void *stream);
/* Same product, result written to n_outs (1..8) destinations (host array of device pointers).  Multi-GPU WMF
 * passes the address of this rank's row block inside EVERY rank's replica of the factor matrix (NVLink peer
 * memory, e.g. torch symmetric memory): the back-transform X = X~ L^-1 and the all-gather of the solved block
 * (cymf/wmf.pyx has no counterpart) are then one kernel -- rows reach all replicas from the
 * GEMM epilogue.  The caller must run a cross-rank barrier before any rank reads the gathered matrix. */
int cymf_rows_times_matrix_multi_dev(const void *in, void *const *outs, int32_t n_outs, const void *B, int dtype,
                                     int64_t rows, int32_t ld, void *stream);

/* Direct solve of the block's heaviest rows order[0 .. n_heavy) (rows are sorted heaviest first): the K x K matrix
 * G + (weight-1) sum_{c in row} y_c y_c^T the reference materialises per row (cymf/wmf.pyx:161-166) is accumulated
 * once on the tensor cores, 512 entries per CTA, and solved in f64 (wmf.pyx:168) -- instead of one CTA streaming a
 * 700 k-entry row ~7 times while the GPU idles.  first_slab: device int32[n_heavy+1], slab counts ceil(len/512)
 * prefix-summed.  G64 dense [K,K] doubles (+ add_diag on the diagonal), or NULL = identity (Y in the transformed
 * coordinates of cymf_chol_transforms_dev).  f32, ld in {32,64,96,128}; otherwise CYMF_EUNSUPPORTED. */
int64_t cymf_als_heavy_workspace_doubles(int64_t n_slabs, int32_t K, int32_t ld);
int cymf_als_heavy_rows_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_heavy,
                            const int32_t *first_slab, int32_t n_slabs, void *X, const void *Y, const double *G64,
                            double add_diag, int dtype, int32_t K, int32_t ld, double weight, double *workspace,
                            int64_t workspace_doubles, void *stream);

/* One-pass row solver (f32, ld in {32,64,96,128}, factors in the transformed coordinates of
 * cymf_chol_transforms_dev): per row the K x K matrix S = sum_{i in row} y~_i y~_i^T that the reference accumulates
 * entry by entry (cymf/wmf.pyx:161-166) is built ONCE on the tensor cores (tcgen05 3xTF32, TMEM accumulators) from a
 * single gather of the row's item vectors, then (I + (weight-1) S) x~ = weight sum y~_i (wmf.pyx:163,168) is solved
 * by conjugate gradient out of registers (same stopping rule and statistics as cymf_als_cg_dev).  Rows without
 * entries are zeroed (wmf.pyx:154-156).  Other dtypes / strides: CYMF_EUNSUPPORTED (use cymf_als_cg_dev). */
int cymf_als_rows_tc_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve,
                         void *X, const void *Y, int dtype, int32_t K, int32_t ld, double weight, double cg_tol,
                         int32_t cg_max_iter, int32_t *queue, unsigned long long *stats, void *stream);

/* Warp-specialised persistent form of cymf_als_rows_tc_dev (same row system, cymf/wmf.pyx:161-168, same stopping rule
 * and statistics): ONE 512-thread CTA per SM in which gather warps (cp.async into MN-major operand stages, four chunks
 * ahead), an MMA-issuing warp (tcgen05 3xTF32 into four TMEM accumulators) and two solver groups (fold + CG out of
 * registers, one thread per row of S) work on different rows at the same time.  Rows are assigned to CTAs beforehand:
 * cymf_als_ws_schedule_host (HOST arrays; longest-processing-time-first over nnz + row_cost) fills cta_ptr[n_ctas+1]
 * and rowinfo[4 n] = (row, nnz, indptr low word, indptr high word) per row, CTA after CTA, which the caller copies to
 * the device for cymf_als_rows_ws_dev.  cymf_als_ws_ctas() = the CTA count to schedule for (one per SM).
 * y_rows = rows of Y: with it the gather runs through the TMA unit (cp.async.bulk.tensor tile::gather4 over a tensor map
 * of Y, four item vectors per copy) when Y is larger than 96 MB or CYMF_ALS_WS_TMA=1; y_rows <= 0 or
 * CYMF_ALS_WS_TMA=0 selects the cp.async (LDGSTS) gather.
 * debug (DEVICE, 32 x u64, zeroed by the caller; may be NULL; [8..24) are progress markers of CTA 0's roles): a hand-over between the roles that does not complete
 * within ~0.5 s is recorded there (count, CTA, wait site, parity, three counters, thread) and 2^40 is added to
 * stats[1]; the kernel then runs to its end instead of hanging. */
int32_t cymf_als_ws_ctas(void);
int cymf_als_ws_schedule_host(const int64_t *indptr, const int32_t *rows, int32_t n, int32_t n_ctas, int32_t row_cost,
                              int32_t *cta_ptr, int32_t *rowinfo);
int cymf_als_rows_ws_dev(const int32_t *rowinfo, const int32_t *cta_ptr, int32_t n_ctas, const int32_t *indices, void *X,
                         const void *Y, int64_t y_rows, int dtype, int32_t K, int32_t ld, double weight, double cg_tol,
                         int32_t cg_max_iter, unsigned long long *stats, unsigned long long *debug, void *stream);

/* Short-row solver (f32, ld in {32,64,96,128}, transformed coordinates; rows of at most 128 entries): the same row
 * system (cymf/wmf.pyx:161-168) in its dual form.  With Y~ the n x K matrix of the row's item vectors,
 * x~ = weight Y~^T z where (I_n + (weight-1) Y~ Y~^T) z = 1 (push-through identity), an n x n system: Y~ Y~^T comes
 * from ONE gather of the row (cp.async into a K-major tcgen05 operand tile, 3xTF32, TMEM accumulator), CG on it runs
 * out of registers, and x~ is formed from the tile still in shared memory.  order[0 .. n128) are rows of 65..128
 * entries (ld = 128 only), the next n64 rows have 33..64 entries, the last n32 rows have 0..32 (rows without entries
 * are zeroed, wmf.pyx:154-156); one / two / four rows share a 128-row tile.  X rows are overwritten (no warm
 * start).  queue: three int32 work-queue heads.  stats as cymf_als_cg_dev. */
int cymf_als_rows_dual_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n128,
                           int32_t n64, int32_t n32, void *X, const void *Y, int dtype, int32_t K, int32_t ld,
                           double weight, double cg_tol, int32_t cg_max_iter, int32_t *queue,
                           unsigned long long *stats, void *stream);

/* warps_per_row of cymf_als_cg_dev is 4, 8 or 16 (CTA width per row; each warp brings 8 KB of shared-memory
 * staging for the row's item vectors).  Given the row lengths in `order` (decreasing), this returns how many
 * leading rows should be solved with 16 warps and how many following ones with 8; the rest take 4. */
int cymf_als_row_classes(const int64_t *sorted_lengths_desc, int64_t n, int dtype, int32_t ld,
                         int64_t *n_wide16, int64_t *n_wide8);

/* Host-buffer form of WMF._als(indptr, indices, X, Y, num_threads): X [rows,K], Y [n,K] dense f64 HOST arrays,
 * host CSR (int32), X solved in place.  dtype selects the device arithmetic (CYMF_F32 / CYMF_F64). */
int cymf_als_half_host(const int32_t *indptr, const int32_t *indices, double *X, const double *Y,
                       int64_t rows, int64_t n, int32_t K, double weight_decay, double weight,
                       int dtype, double cg_tol, int32_t cg_max_iter, int64_t *cg_iterations_out);

/* ---- Evaluator (cymf/evaluator.pyx:57-139, cymf/metrics.pyx:24-125; unbiased=False path) ---------------- */
/* Candidate lists of Evaluator.evaluate (evaluator.pyx:91-111), HOST arrays in and out: per user with test
 * items, its test positives in CSR order followed by `num_negatives` draws of the reference's sequential
 * mt19937 stream (seed), rejecting test+train positives (`all_*`: CSR of test+train, sorted rows).
 * cand_ptr [U+1] receives offsets into cand_items (`capacity` entries available; the exact need is
 * nnz(test) + #users_with_test_items * num_negatives). */
int cymf_eval_candidates_host(int32_t U, int32_t I, const int32_t *test_indptr, const int32_t *test_indices,
                              const int32_t *all_indptr, const int32_t *all_indices,
                              int32_t num_negatives, uint32_t seed, int64_t *cand_ptr, int32_t *cand_items,
                              int64_t capacity);

/* Scoring + ranking + metrics on the device (evaluator.pyx:113-133).  W [U,K], H [I,K] dense f64 (the
 * reference casts to f64, evaluator.pyx:58-59).  ks [nk] (device) are the cut-offs, log2_table[i] =
 * log2(i+1) for i < kmax = max(ks) (device; computed by the host's libm as metrics.pyx:38 does).
 * per_user [U, nk, 3] receives DCG@k, Recall@k, MAP@k of every user (0 for users without test items); the
 * caller averages over ALL U users (evaluator.pyx:135-137).  order (may be NULL): for every candidate slot
 * of cand_items, the candidate position holding that rank (descending score, ties by descending position). */
int cymf_eval_rank_dev(const double *W, const double *H, int32_t U, int32_t K,
                       const int32_t *test_indptr, const int64_t *cand_ptr, const int32_t *cand_items,
                       int32_t max_candidates, const int32_t *ks, int32_t nk, const double *log2_table,
                       int32_t kmax, double *per_user, int32_t *order, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CYMF_B200_H */
