/*
 * bear_b200.h -- C-ABI of libbear_b200.so, the B200 (sm_100a) implementation of BEAR's
 * data-parallel hot path (per-k-mer Dirichlet-multinomial marginal log-likelihood + gradient).
 *
 * The reference (debbiemarkslab/BEAR) is pure Python/TensorFlow and has NO native boundary; the
 * entry points below are what a ctypes binding inside the reference's modules would call instead
 * of building the TF graph.  Each entry cites the reference code it replaces (paths relative to
 * bear_model/ in the reference).  INTEGRATION.md shows the reference-side ctypes stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch/TF types.  Pointers named d_* are DEVICE pointers
 *     (caller-allocated, e.g. torch tensors), h_* are HOST pointers.  `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).  All device entry points are asynchronous
 *     on `stream`; the library never allocates or frees device memory.
 *   - return 0 on success, negative bear_status on failure; bear_last_error() gives the message
 *     (thread-local).
 *   - float64 arithmetic throughout (the reference default, config_files/bear_lin_bear.cfg:6-7).
 *
 * Packed table layout ("device-side packed k-mer batches", replaces dataloader.py:36-46 +
 * core.py:156-174):
 *   kmers  : uint64 [K].  DNA/RNA (alphabet 0/1): bits [0,2L) hold the lag-L context, 2 bits per
 *            symbol (A,C,G,T/U = 0..3), leftmost symbol in the most significant pair, so numeric
 *            order = lexicographic order; bits [58,64) hold n_start, the number of leading start
 *            symbols '[' (they only occur as a prefix run, summarize.py:441-443); payload bits
 *            under the start run are zero.  L <= 29.
 *            Protein (alphabet 2): 5 bits per symbol (0..19 = ARNDCEQGHILKMFPSTWYV, 20 = '[',
 *            31 = unknown -> all-zero one-hot row), L <= 12.
 *   counts : uint32, group-planar [G][A1][stride] (A1 = alphabet_size + 1, last plane = stop ']');
 *            stride >= K is the plane pitch in elements (multiple of 4 for 128-bit loads).
 *            KMC caps counts at ~4e9 (summarize.py:66) so uint32 is lossless.
 */
#ifndef BEAR_B200_H
#define BEAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    BEAR_OK = 0,
    BEAR_ERR_ARG = -1,      /* invalid argument */
    BEAR_ERR_IO = -2,       /* file could not be read */
    BEAR_ERR_PARSE = -3,    /* malformed row / symbol outside the alphabet */
    BEAR_ERR_CUDA = -4,     /* CUDA runtime error (message has cudaGetErrorString) */
    BEAR_ERR_RANGE = -5     /* value does not fit the packed layout (count > 2^32-1, lag too long) */
} bear_status;

enum { BEAR_ALPHABET_DNA = 0, BEAR_ALPHABET_RNA = 1, BEAR_ALPHABET_PROT = 2 };

/* heads for the fused kernels */
enum {
    BEAR_HEAD_NONE = 0,    /* no AR function (BMM / vanilla): f == 0          bear_net.py:38-42  */
    BEAR_HEAD_LINEAR = 1,  /* softmax(sum_j mat[j, s_j, :])                   ar_funcs.py:23-46  */
    BEAR_HEAD_EXPLICIT = 2,/* f given as a dense [rows, A1] float64 array (CNN / plugin heads)   */
    BEAR_HEAD_STOP = 3     /* constant [0,..,0,1]                             ar_funcs.py:102-127*/
};

const char* bear_last_error(void);
int bear_version(void);                       /* ABI version, currently 1 */
const char* bear_build_digest(void);          /* sha256 of the sources + compiler flags this binary was built from */
int bear_alphabet_size(int alphabet);         /* 4 / 4 / 20 */
int bear_max_lag(int alphabet);               /* 29 / 29 / 12 */

/* ------------------------------------------------------------------------------------------
 * Host ingest: text -> packed table.  Replaces dataloader.dataloader (dataloader.py:6-50) and
 * dataloader.sparse_dataloader (dataloader.py:52-109) and the `wc -l` row count of
 * models/train_bear_net.py:54-55.
 * ---------------------------------------------------------------------------------------- */
/* number of data rows (non-empty lines, minus the header line if header != 0); <0 on error */
int64_t bear_count_rows(const char* path, int header);

/* Dense TSV `kmer \t [[g0 counts],[g1 counts],...]`.  Writes up to max_rows rows starting at file
 * row `first_row` into h_kmers[max_rows] and h_counts[G][A1][stride].  *rows_out = rows written,
 * *lag_out = k-mer length.  Rows with a different length, symbols outside the alphabet, '[' after a
 * non-'[' symbol (DNA/RNA) or counts that are negative / non-integer / > 2^32-1 are errors. */
int bear_pack_tsv(const char* path, int header, int alphabet, int num_ds,
                  int64_t first_row, int64_t max_rows,
                  uint64_t* h_kmers, uint32_t* h_counts, int64_t stride,
                  int64_t* rows_out, int* lag_out);

/* Sparse `kmer; [[g,b],...]; [v,...]` (';'-separated, dataloader.py:82-105), same outputs. */
int bear_pack_sparse(const char* path, int header, int alphabet, int num_ds,
                     int64_t first_row, int64_t max_rows,
                     uint64_t* h_kmers, uint32_t* h_counts, int64_t stride,
                     int64_t* rows_out, int* lag_out);

/* DNA / RNA k-mers with a symbol outside the alphabet (e.g. 'N'; the reference one-hots such symbols to a row of
 * zeros, core.py:162, 2-bit codes cannot): policy 0 (default) = bear_pack_* fail with BEAR_ERR_PARSE naming the row;
 * policy 1 = the row's code is BEAR_INVALID_KMER and the caller drops it (dataloader.KmerTable.from_file(on_invalid=
 * 'skip')).  Process-wide; returns the previous policy. */
#define BEAR_INVALID_KMER 0xffffffffffffffffull
int bear_pack_set_invalid_policy(int policy);

/* Rank-sharded ingest for data-parallel training (the reference splits every global batch over the replicas inside
 * one process, bear_net.py:273; here every rank parses only its own rows): the dataset -- this file's rows at index
 * `row_offset` of `dataset_rows` rows in all (several files of one dataset, models/train_bear_net.py:78-86; one file:
 * row_offset = 0, dataset_rows = -1) -- is cut into global batches of `batch_rows` consecutive rows, and of every batch
 * rank keeps the contiguous slice [rank * per, (rank + 1) * per), per = ceil(rows of the batch / world).  The rank's rows
 * of THIS file are written densely in order.  sparse = 0: the dense TSV format, 1: the sparse format.  *rows_out = rows
 * written (needs max_rows and stride >= that). */
int bear_pack_shard(const char* path, int sparse, int header, int alphabet, int num_ds, int64_t batch_rows,
                    int world, int rank, int64_t row_offset, int64_t dataset_rows, int64_t max_rows,
                    uint64_t* h_kmers, uint32_t* h_counts, int64_t stride, int64_t* rows_out, int* lag_out);

/* n k-mer strings of length `lag`, concatenated without separators -> packed codes, and back. */
int bear_encode_kmers(const char* h_text, int64_t n, int lag, int alphabet, uint64_t* h_kmers);
int bear_decode_kmers(const uint64_t* h_kmers, int64_t n, int lag, int alphabet, char* h_text);

/* ------------------------------------------------------------------------------------------
 * Compact transfer format.  Host-to-device copies bound the end-to-end rate (28 B per row at PCIe speed), so a
 * table can cross the bus as byte planes: ceil(kbits/8) planes for the k-mer (kbits = 2*lag + 6 for DNA/RNA --
 * the 6 bits hold n_start --, 5*lag for protein) and one byte plane per count column and letter; a count
 * >= 255 is stored as 255 plus an escape entry {plane, row, value}.  Lossless for any table.
 * `wire` selects the variant: wire & 15 = bits per count, 8 or 4 (4: two rows per byte, low nibble = even row; a
 * count >= 15 is stored as 15 plus an escape), or 12: ONE 12-bit code per (row, group) -- the rank of the row's whole
 * count vector among the vectors whose counts sum to at most 10 (DNA/RNA, 3003 vectors; protein: sum <= 3), low byte in a
 * byte plane and high nibble in a nibble plane per group; rows with larger sums carry code 4095 and send their non-zero
 * counts as escapes (1.5 B per row and group instead of 2.5 B for the five nibble planes); wire & BEAR_WIRE_START_ESC (DNA/RNA): the k-mer planes hold the 2*lag
 * payload bits only and n_start of the start-padded rows travels as escape entries {0xffffffff, row, n_start} after
 * the count escapes -- one plane less at lag 20.  7.6 instead of 11 B per row for a one-column DNA table at lag 20
 * with sparse counts and 1 % start-padded rows.  bear_compact_choose_wire (host) scans the rows and returns the
 * variant with the fewest bytes on the wire, escapes (12 B each) included.
 * Plane pitch = n rounded up to 16 bytes (count planes: pitch * bits / 8); bear_compact_bytes() = the size of all
 * planes of n rows.
 * bear_compact_table (host, multi-threaded) writes rows [row0, row0+n) of a packed table into h_out and the
 * escapes (sorted by plane, row) into h_esc[esc_cap][3]; *n_esc_out = entries needed (re-call with a larger
 * capacity if it exceeds esc_cap).  bear_expand_table (device) restores rows [dst_row0, dst_row0+n) of the packed
 * table d_kmers / d_counts (plane pitch `stride`) bit-exactly.  Replaces nothing in the reference (its loader
 * re-parses text every epoch, dataloader.py:36-46); it is the wire format of dataloader.KmerTable uploads.
 * ---------------------------------------------------------------------------------------- */
#define BEAR_WIRE_START_ESC 16
int64_t bear_compact_bytes(int64_t n, int lag, int alphabet, int G, int wire);
int bear_compact_choose_wire(const uint64_t* h_kmers, const uint32_t* h_counts, int64_t stride, int64_t row0, int64_t n,
                             int lag, int alphabet, int G);
int bear_compact_table(const uint64_t* h_kmers, const uint32_t* h_counts, int64_t stride, int64_t row0, int64_t n,
                       int lag, int alphabet, int G, int wire, uint8_t* h_out, uint32_t* h_esc, int64_t esc_cap,
                       int64_t* n_esc_out);
int bear_expand_table(const uint8_t* d_compact, const uint32_t* d_esc, int64_t n_esc, int64_t n, int lag,
                      int alphabet, int G, int wire, uint64_t* d_kmers, uint32_t* d_counts, int64_t stride,
                      int64_t dst_row0, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: unpacking to the reference's dense tensors (core.tf_one_hot core.py:156-174; the counts
 * tensor of dataloader.py:44-46).  Used by plugin AR heads and by API-compat iteration.
 * ---------------------------------------------------------------------------------------- */
int bear_decode_onehot(const uint64_t* d_kmers, int64_t n, int lag, int alphabet,
                       double* d_onehot /* [n, lag, A1] */, void* stream);
int bear_decode_symbols(const uint64_t* d_kmers, int64_t n, int lag, int alphabet,
                        uint8_t* d_sym /* [n, lag], value A1 = all-zero row */, void* stream);
int bear_unpack_counts(const uint32_t* d_counts, int64_t stride, int64_t row0, int64_t n,
                       int G, int A1, double* d_out /* [n, G, A1] */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: generic dense distributions (core.tfpDirichletMultinomialPerm / tfpMultinomialPerm,
 * core.py:11-139).  `conc` / `probs` have conc_rows rows and broadcast over value rows with
 * period conc_rows (n % conc_rows == 0): row i uses conc[i % conc_rows].  Real-valued counts
 * are allowed (lgamma, not the integer fast path).
 * ---------------------------------------------------------------------------------------- */
int bear_dm_logprob(const double* d_conc, int64_t conc_rows, const double* d_value, int64_t n,
                    int A1, double* d_out /* [n] */, void* stream);
/* d out / d conc for an upstream gradient d_gout[n] (NULL = ones): d_gconc[n, A1] (not reduced over
 * the broadcast).  digamma form of bear_net.py:193's tape gradient. */
int bear_dm_logprob_bwd(const double* d_conc, int64_t conc_rows, const double* d_value, int64_t n,
                        int A1, const double* d_gout, double* d_gconc, void* stream);
int bear_mn_logprob(const double* d_probs, int64_t probs_rows, const double* d_value, int64_t n,
                    int A1, double* d_out, void* stream);
int bear_mn_logprob_bwd(const double* d_probs, int64_t probs_rows, const double* d_value, int64_t n,
                        int A1, const double* d_gout, double* d_gprobs, void* stream);
/* ml_output (core.py:69-71,134-136): argmax over the last axis of x[n, A1] + sigma*N(0,1) noise.
 * Noise is counter-based (seed, row, letter); seed < 0 disables it (plain first-max argmax). */
int bear_ml_output(const double* d_x, int64_t n, int A1, double sigma, int64_t seed,
                   double* d_out /* [n] index as float, like the reference */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: fused packed-path kernels (DNA/RNA, A1 = 5).
 * ---------------------------------------------------------------------------------------- */
/* Workspace (float64 elements) the fused entry points need for `n` rows; callers allocate once. */
int64_t bear_workspace_doubles(int64_t n, int lag, int nparams);

/* bear_net._train_step (bear_net.py:146-197) with the linear head (ar_funcs.py:23-46), fused
 * forward + backward over rows [row0, row0+n) of one count column:
 *   ll_k   = DM (train_ar=0, bear_net.py:43-44) or multinomial (train_ar=1, bear_net.py:68-69)
 *   loss   = -scale * sum_k ll_k,   scale = num_kmers / batch_rows   (bear_net.py:187-191)
 * Adds into d_flat[0] the loss, d_flat[1] d loss/d h_signed, d_flat[2 + ...] d loss/d mat
 * ([lag, A1, A1] row-major) -- the flat [loss, h, params...] buffer that is allreduced once per
 * step.  d_h_signed and d_mat are device scalars/arrays so the launch is CUDA-graph capturable.
 * d_ll_out (optional, may be NULL) receives the per-k-mer log-likelihoods [n]. */
int bear_linear_train_step(const uint64_t* d_kmers, const uint32_t* d_col /* plane 0 of the column */,
                           int64_t stride, int64_t row0, int64_t n, int lag,
                           const double* d_mat, const double* d_h_signed, double scale, int train_ar,
                           double* d_flat, double* d_ll_out, double* d_workspace, void* stream);

/* Same contract for an arbitrary head evaluated by the caller: f[n, A1] in, d loss / d f out
 * (d_gf[n, A1]); adds loss and d loss/d h_signed into d_flat[0..1].  This is the entry the CNN /
 * bear_ref / plugin heads use (bear_ref.py:207-259 shares it). */
int bear_dm_train_step_explicit(const uint32_t* d_col, int64_t stride, int64_t row0, int64_t n,
                                const double* d_f, const double* d_h_signed, double scale,
                                int train_ar, double* d_flat, double* d_gf, double* d_ll_out,
                                double* d_workspace, void* stream);

/* bear_net._evaluation_step + the accumulation of bear_net.evaluation / h_scan
 * (bear_net.py:323-371,439-463,516-531) over rows [row0,row0+n):
 *   ear: conc = f/h_i + train + eps for each of H h-values   (H <= BEAR_MAX_MODELS)
 *   arm: p = f + eps
 *   van: conc = train + van_j + eps for each of V van_reg values (V <= BEAR_MAX_MODELS)
 * d_train_col may be NULL (ds_loc_train = -1: no conditioning, bear_net.py:329-331).
 * Adds into d_acc (float64 [2H + 2 + 2V + 1]):
 *   [ll_ear[H], ll_arm, ll_van[V], correct_ear[H], correct_arm, correct_van[V], total_len].
 * head = BEAR_HEAD_LINEAR (d_head = mat), BEAR_HEAD_EXPLICIT (d_head = f[n, A1], row i = table row row0 + i),
 * BEAR_HEAD_STOP or BEAR_HEAD_NONE (f = 0: ear degenerates to conc = train + eps).
 * Argmax tie-breaking noise: sigma = 100*eps (ear, van) / eps (arm) as in core.py:70,135; seed < 0
 * disables the noise.  The noise of a row is a pure function of (seed, row_id0 + i, model), where row_id0 is the
 * index of row `row0` in the whole (unsharded) table: results do not depend on how rows are spread over GPUs. */
#define BEAR_MAX_MODELS 8
int bear_eval_step(const uint64_t* d_kmers, const uint32_t* d_test_col, const uint32_t* d_train_col,
                   int64_t stride, int64_t row0, int64_t n, int lag,
                   int head, const double* d_head,
                   const double* d_h /* [H] */, int H, const double* d_van /* [V] */, int V,
                   int64_t seed, int64_t row_id0, double* d_acc, double* d_workspace, void* stream);

/* One fused training step of a reference-based model with the stop net (bear_ref._train_step, bear_ref.py:207-259, with
 * make_ar_func_stop: the configs bear_stop_bear.cfg / bear_stop_ar.cfg): head f = (nw stop + JC(ref, tau)) / (nw + 1)
 * (bear_ref.py:9-69), Dirichlet-multinomial (train_ar = 0) or multinomial (1) log-likelihood of the data column and the
 * analytic backward in ONE pass over the two columns (40 B per row, no k-mers, no f array in memory):
 *   d_flat[0] += loss = -scale * sum_k ll_k;  d_flat[1..3] += d loss / d [h_signed, tau_signed, net_weight_signed].
 * d_ll_out (nullable, [n]) receives the per-k-mer log-likelihoods. */
int bear_ref_train_step(const uint32_t* d_col, const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n,
                        const double* d_h_signed, const double* d_tau_signed, const double* d_nw_signed, double scale,
                        int train_ar, double* d_flat, double* d_ll_out, double* d_workspace, void* stream);

/* The same evaluation for a reference-based model (bear_ref._evaluation_step / evaluation, bear_ref.py:391-539) in ONE
 * kernel: the head f = (nw g(k) + JC(ref, tau)) / (nw + 1) of bear_ref.py:9-69 (Jukes-Cantor mix of the reference
 * column's counts; nw = exp(*d_nw_signed), tau = exp(*d_tau_signed)) is evaluated in registers, no f array exists in
 * memory.  net = BEAR_HEAD_STOP (g = stop vector, d_mat and d_kmers unused) or BEAR_HEAD_LINEAR (g = linear head of
 * d_mat[lag, A1, A1] on the k-mers).  Other arguments as bear_eval_step. */
int bear_ref_eval_step(const uint64_t* d_kmers, const uint32_t* d_test_col, const uint32_t* d_train_col,
                       const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n, int lag, int net,
                       const double* d_mat, const double* d_tau_signed, const double* d_nw_signed,
                       const double* d_h /* [H] */, int H, const double* d_van /* [V] */, int V,
                       int64_t seed, int64_t row_id0, double* d_acc, double* d_workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: the CNN autoregressive head (ar_funcs.make_ar_func_cnn, ar_funcs.py:49-99) fused with the
 * loss: conv1d as a gather -> layer norm -> elu -> dense (FP64 tensor cores) -> layer norm -> elu ->
 * dense -> softmax, and its whole backward pass, in one kernel (DNA/RNA, A1 = 5).
 * d_params: the eight parameter arrays concatenated in the reference's list order (ar_funcs.py:98-99):
 *   filters[W,5,F], intercept0[P,F], weights1[P,F,H1], intercept1[H1], weights2[H1,5], intercept2[5],
 *   scale0[P,F], scale1[H1]   with P = lag - W + 1;  bear_cnn_num_params() doubles in total.
 * bear_cnn_supported() is 1 when the dimensions fit the fused kernels (F <= 32, H1 <= 16, tile in
 * shared memory); otherwise the entry points return BEAR_ERR_RANGE and the head has to be evaluated by
 * the caller (BEAR_HEAD_EXPLICIT / bear_dm_train_step_explicit).
 * ---------------------------------------------------------------------------------------- */
int bear_cnn_supported(int lag, int filter_width, int num_filters, int layer1_width);
int64_t bear_cnn_num_params(int lag, int filter_width, int num_filters, int layer1_width);
/* f[n, 5] = ar_func(k-mers row0..row0+n) (ar_funcs.py:91-97); row i of d_f = table row row0 + i */
int bear_cnn_head_forward(const uint64_t* d_kmers, int64_t row0, int64_t n, int lag, int filter_width,
                          int num_filters, int layer1_width, const double* d_params, double* d_f, void* stream);
/* bear_net._train_step (bear_net.py:146-197) with the CNN head: same contract as bear_linear_train_step;
 * adds [loss, d loss/d h_signed, d loss/d params (order of d_params)] into d_flat[0 .. 2+num_params). */
int bear_cnn_train_step(const uint64_t* d_kmers, const uint32_t* d_col, int64_t stride, int64_t row0, int64_t n,
                        int lag, int filter_width, int num_filters, int layer1_width, const double* d_params,
                        const double* d_h_signed, double scale, int train_ar, double* d_flat, double* d_ll_out,
                        double* d_workspace, void* stream);
/* Backward of bear_cnn_head_forward for an upstream gradient d_gf[n, 5] (w.r.t. f): adds the parameter
 * gradients into d_gparams[num_params] (used when the CNN is the embedded net of bear_ref, bear_ref.py:63-68). */
int bear_cnn_head_backward(const uint64_t* d_kmers, int64_t row0, int64_t n, int lag, int filter_width,
                           int num_filters, int layer1_width, const double* d_params, const double* d_gf,
                           double* d_gparams, double* d_workspace, void* stream);

/* dataloader._marginal_step / bmm_likelihood (dataloader.py:111-147) on the packed table:
 * adds sum_k lbeta(c + a_v) - lbeta(a_v) for every group and alpha into d_out[G, V]. */
int bear_bmm_likelihood(const uint32_t* d_counts, int64_t stride, int64_t row0, int64_t n,
                        int G, int A1, const double* d_alpha, int V, double* d_out,
                        double* d_workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: posterior sampling (log_gamma.py:17-76, get_var_probs.py:172-175).
 * ---------------------------------------------------------------------------------------- */
/* out[s, i] = log X, X ~ Gamma(conc[i], 1), for s < n_samples; counter-based RNG (seed, s, i). */
int bear_loggamma_sample(const double* d_conc, int64_t n, int64_t n_samples, int64_t seed,
                         double* d_out /* [n_samples, n] */, void* stream);
/* log-normalise groups of A1 consecutive values in place: x -= logsumexp(x) (get_var_probs.py:175) */
int bear_log_normalize(double* d_x, int64_t n_groups, int A1, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: reference-genome head (bear_ref.py:9-69) and the optimizer update.
 * ---------------------------------------------------------------------------------------- */
/* f = (nw*g + JC(ref, tau)) / (nw + 1), nw = exp(net_weight_signed), tau = exp(tau_signed)
 * (bear_ref.py:63-68) with ref = (counts[:, ds_loc_ref] + eps) * not_stop (bear_ref.py:332-337)
 * read from the packed reference column.  d_g[n, A1] is the embedded net's output, NULL = the stop
 * head (ar_funcs.py:102-127).  Writes d_f[n, A1] (row i of the output = table row row0 + i). */
int bear_ref_head(const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n, const double* d_g,
                  const double* d_tau_signed, const double* d_nw_signed, double* d_f, void* stream);
/* Backward of bear_ref_head for an upstream d_gf[n, A1]: d_gg[n, A1] (may be NULL) and
 * d_flat2[0..1] += [d/d tau_signed, d/d net_weight_signed]. d_workspace: bear_workspace_doubles. */
int bear_ref_head_bwd(const uint32_t* d_ref_col, int64_t stride, int64_t row0, int64_t n, const double* d_g,
                      const double* d_tau_signed, const double* d_nw_signed, const double* d_gf, double* d_gg,
                      double* d_flat2, double* d_workspace, void* stream);

/* tf.keras.optimizers.Adam step (bear_net.py:264-265,277-282) on a flat float64 buffer:
 * t = *d_step + 1; lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMAs; p -= lr_t*m/(sqrt(v)+eps); ++*d_step.
 * The step counter lives on the device so the update is CUDA-graph capturable. */
int bear_adam_update(double* d_params, const double* d_grads, double* d_m, double* d_v, int64_t n, double lr,
                     double beta1, double beta2, double eps, int64_t* d_step, void* stream);

/* One optimizer step of bear_net.train's loop (bear_net.py:300-313) in a single launch, for the flat buffers of the
 * training entry points: d_flat = [loss, d h_signed, d params...] (1 + n doubles, after the allreduce); the Adam update
 * of bear_adam_update is applied to d_params[n] with d_flat[1..n]; if d_loss_out is not NULL it receives
 * loss_scale * d_flat[0] (the recorded "elbo" is -loss / acc_steps); with zero_flat != 0 d_flat is cleared for the next
 * accumulation; ++*d_step.  n <= 2^22. */
int bear_adam_step(double* d_params, double* d_flat, double* d_m, double* d_v, int64_t n, double lr, double beta1,
                   double beta2, double eps, int64_t* d_step, double* d_loss_out, double loss_scale, int zero_flat,
                   void* stream);

/* Deterministic synthetic packed table for benchmarks and large-scale parity checks: rows
 * [row_begin, row_begin+n) of the table defined by (seed, lag, G, regime) -- see DESIGN.md.
 * regime bit 0: 0 = sparse counts (N = 1 + Poisson(2)), 1 = dense counts (N ~ LogNormal(ln 300, 1.5));
 * regime bit 1: 0 = pseudo-random row order (shuffled table), 1 = rows sorted by k-mer (KMC order).
 * d_kmers[n], d_counts[G][5][stride] (row i of the output = table row row_begin + i). */
int bear_synth_table(uint64_t* d_kmers, uint32_t* d_counts, int64_t stride, int64_t row_begin, int64_t n,
                     int lag, int G, int64_t seed, int regime, int start_permille, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device: building the count table from sequences (the table summarize.py emits; its definition is
 * the brute-force count over '[' * lag + seq + ']' of tests/test_summarize.py:96-114).
 * ---------------------------------------------------------------------------------------- */
/* d_seq: all sequences concatenated (bytes, ACGT/U in either case); d_offsets[nseq+1]: start of each
 * sequence; d_toff[nseq]: exclusive prefix sum of (len_i + 1), the transitions of each sequence;
 * d_groups[nseq]: dataset group of each sequence; ntrans = sum(len_i + 1).  With reverse_complement
 * the reverse complement of every sequence is counted too (summarize.py -r).  The hash table
 * d_keys[cap] (initialised to all-ones) / d_counts[cap][G][5] (zeroed) has cap a power of two
 * > distinct k-mers.  d_stats[3] += {distinct k-mers, skipped transitions (non-ACGT), count overflows}. */
int bear_count_transitions(const uint8_t* d_seq, const int64_t* d_offsets, const int64_t* d_toff,
                           const int32_t* d_groups, int64_t nseq, int64_t ntrans, int lag, int G,
                           int reverse_complement, uint64_t* d_keys, uint32_t* d_counts, int64_t cap,
                           uint64_t* d_stats, void* stream);
/* Rows d_rows[n] (occupied slots of the hash table) -> packed table d_out_kmers[n], d_out_counts[G][5][stride]. */
int bear_gather_table(const uint64_t* d_keys, const uint32_t* d_counts, const int64_t* d_rows, int64_t n, int G,
                      int64_t stride, uint64_t* d_out_kmers, uint32_t* d_out_counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEAR_B200_H */
