/* strotss_b200 -- C ABI of the B200-native STROTSS loss hot path.
 *
 * Drop-in boundary for the per-iteration loss (+gradient) evaluation of
 * interaction-lab-uh/STROTSS-tensorflow.  Every entry point names the reference interface it
 * replaces (file:line are relative to the reference tree).  Plain pointers and sizes only; all
 * matrix arguments are DEVICE pointers to dense row-major fp32 unless the name ends in _host.
 * Work is enqueued on the CUDA stream passed as `stream` (a cudaStream_t cast to void*); nothing
 * synchronises with the host except the *_host entry points.  The caller owns every input/output
 * buffer; the handle owns a grow-only device workspace (no N x M / N x N / full-size temporaries
 * other than one L2-sized row panel).  There is no CPU fallback: every call fails with
 * STROTSS_ERR_CUDA if no sm_100 device is usable.
 *
 * Return value: 0 on success, negative STROTSS_ERR_* otherwise; strotss_last_error() gives text.
 */
#ifndef STROTSS_B200_H
#define STROTSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct strotss_ctx* strotss_handle;

enum {
    STROTSS_OK = 0,
    STROTSS_ERR_ARG = -1,          /* bad argument (null pointer, non-positive size, ...)          */
    STROTSS_ERR_DISTANCE = -2,     /* unknown distance name: wrappers raise KeyError (losses.py:74) */
    STROTSS_ERR_CUDA = -3,         /* CUDA runtime / driver failure, or no sm_100 device           */
    STROTSS_ERR_STATE = -4,        /* call order violated (e.g. no style target set)               */
    STROTSS_ERR_UNSUPPORTED = -5   /* combination outside the hot path (l2/both with D != 3)       */
};

/* distance codes == keys of dist_metrics (nn/losses.py:27-28) */
enum { STROTSS_DIST_COSINE = 0, STROTSS_DIST_L2 = 1, STROTSS_DIST_BOTH = 2 };

/* slots of the scalar block written by strotss_eval* (device or host float[16]) */
enum {
    STROTSS_S_TOTAL = 0,        /* (alpha*loss_c + loss_s)/loss_denom        run_strotss.py:140 */
    STROTSS_S_LOSS_C = 1,       /* ContentLoss                                run_strotss.py:21-24 */
    STROTSS_S_LOSS_S = 2,       /* StyleLoss                                  run_strotss.py:33-40 */
    STROTSS_S_L_M = 3,          /* moment_matching                            nn/losses.py:39-52 */
    STROTSS_S_L_REMD = 4,       /* relaxed_emd cosine                         nn/losses.py:69-80 */
    STROTSS_S_L_PALETTE = 5,    /* relaxed_emd 'both' on YUV                  run_strotss.py:37-39 */
    STROTSS_S_REMD_RX = 6, STROTSS_S_REMD_RY = 7,
    STROTSS_S_L_COV = 8, STROTSS_S_L_MEAN = 9,
    STROTSS_S_PAL_RX = 10, STROTSS_S_PAL_RY = 11,
    STROTSS_S_REMD_BRANCH = 12, /* 1.0 if R_X >= R_Y (gradient flows through the target->pred minima) */
    STROTSS_S_PAL_BRANCH = 13,
    STROTSS_NUM_SCALARS = 16
};

/* Library / handle life cycle.  Replaces nothing in the reference (TensorFlow owns its runtime);
 * `device` plays the role of --gpu_id (run_strotss.py:176,179, nn/utils.py:73-85). */
int strotss_create(int device, strotss_handle* out);
void strotss_destroy(strotss_handle h);
const char* strotss_last_error(strotss_handle h);
const char* strotss_version(void);
/* bytes of device workspace currently held by the handle */
size_t strotss_workspace_bytes(strotss_handle h);
/* Output buffers for bindings whose tensors are immutable (TensorFlow: INTEGRATION.md): device memory on the handle's GPU,
 * owned by the caller until strotss_device_free (which may be called after the handle is gone: `h` is ignored there).  The tf.custom_gradient adapter wraps such a buffer in a DLPack capsule
 * (tf.experimental.dlpack.from_dlpack) instead of writing into an EagerTensor.  No reference counterpart (TensorFlow's
 * allocator does this inside every op, e.g. for the result of tf.matmul at nn/losses.py:15). */
int strotss_device_alloc(strotss_handle h, size_t bytes, void** out);
int strotss_device_free(strotss_handle h, void* ptr);

/* kernels launched by this handle since creation (the bench's gpu_launches counter) */
long long strotss_launch_count(strotss_handle h);

/* Optional per-phase timing (the reference has only a whole-run wall clock, nn/utils.py:97-114):
 * when enabled, every launch group is bracketed by CUDA events on the launching stream.
 * strotss_profile_read synchronises on them, fills ms_sum[i] / counts[i] for
 * i < strotss_profile_num_phases() and clears the records. */
int strotss_profile_enable(strotss_handle h, int on);
int strotss_profile_num_phases(void);
const char* strotss_profile_phase_name(int i);
int strotss_profile_read(strotss_handle h, double* ms_sum, long long* counts);

/* Multi-GPU (one process per GPU, like --gpu_id): the evaluation shards by prediction-sample rows.
 * Rank 0 creates an id with strotss_comm_unique_id (ncclGetUniqueId) and ships the 128 bytes to the
 * other ranks by any means (bench.py uses torch.distributed); every rank then calls strotss_comm_init.
 * Afterwards strotss_eval / strotss_eval_host compute rows [row_begin, row_end) of the prediction on
 * this rank (strotss_shard_rows), write identical scalars on every rank and only this rank's rows of
 * grad_pred.  Inputs are replicated.  Per evaluation the ranks exchange the packed (value, index) minima
 * of the M style rows (max), N + 16 + D floats (sum) and -- for the symmetric self-similarity matrices
 * (nn/losses.py:56-68) and the covariance (nn/losses.py:43-50), whose tiles are dealt out over the ranks --
 * bf16 sign blocks and partial covariance tiles; with CUDA IPC available all of it moves through peer
 * windows over NVLink (strotss_comm_transport), otherwise through NCCL allreduce / send / recv.
 * NCCL is loaded with dlopen("libnccl.so.2") at first use; the library itself does not link it. */
int strotss_comm_unique_id(char* out128);
int strotss_comm_init(strotss_handle h, int rank, int world, const char* id128);
int strotss_shard_rows(strotss_handle h, int N, int* row_begin, int* row_end);
/* How the self-similarity term of a sharded evaluation crosses ranks (nn/losses.py:56-68: Xd, Yd are symmetric, so every
 * 256 x 256 tile is computed by ONE rank and stands for its mirror image as well):
 *   1  CUDA-IPC peer windows -- the bf16 sign blocks of the mirrored tiles are pushed by copy engine into the owner's
 *      window while the GEMMs run, and the owner multiplies them itself (no data-path NCCL kernel besides the small allreduces);
 *  -1  peer windows unavailable (IPC refused, or STROTSS_PEER_WINDOW=0): the fp32 products travel by ncclSend / ncclRecv;
 *   0  not decided yet (no sharded evaluation has run) or no communicator. */
int strotss_comm_transport(strotss_handle h);

/* StyleLoss.__init__(target, alpha) (run_strotss.py:28-31): fix the style target for a scale.
 * Caches what the reference recomputes every iteration (nn/losses.py:43,49; run_strotss.py:37):
 * normalised bf16 operand, column mean, covariance, YUV records.  style: M x D, row stride ld. */
int strotss_set_style_target(strotss_handle h, const float* style, int M, int D, long long ld, void* stream);

/* One train_step loss evaluation (run_strotss.py:136-141 without VGG/sampling):
 *   loss_c = self_similarity(pred, content)                     run_strotss.py:24
 *   loss_s = moment_matching(style, pred) + relaxed_emd(style, pred)
 *            + relaxed_emd(yuv(style), yuv(pred), 'both') / max(alpha, 1)      run_strotss.py:33-40
 *   loss   = (alpha*loss_c + loss_s) / (2 + alpha + 1/max(alpha,1))            run_strotss.py:92,140
 * pred, content: N x D (D as given to strotss_set_style_target), row strides ld_pred / ld_content.
 * scalars: device float[STROTSS_NUM_SCALARS].  grad_pred: device N x D (row stride ld_grad) or NULL.
 * remd_row_argmin (M int32) / remd_col_argmin (N int32): optional device outputs (may be NULL):
 * the prediction index matched to each style sample and the style index matched to each
 * prediction sample. */
int strotss_eval(strotss_handle h, const float* pred, long long ld_pred, const float* content, long long ld_content,
                 int N, float alpha, float* scalars, float* grad_pred, long long ld_grad,
                 int32_t* remd_row_argmin, int32_t* remd_col_argmin, void* stream);

/* Masked (region-guided) transfer, run_strotss.py:97-125 (--content_mask / --style_mask): R regions, each with its
 * own StyleLoss target drawn under the style mask (:99-101) and its own content/prediction samples drawn under the
 * content mask (:115); the per-region losses are averaged (:118-124).  Region r owns rows
 * [offsets[r], offsets[r+1]) of the concatenated matrices; offsets are HOST arrays of R+1 ints starting at 0, every
 * region non-empty (N_r is dynamic per iteration, nn/strotss_utils.py:113,120; M_r is fixed per scale).
 * strotss_set_style_targets_grouped replaces the StyleLoss constructions of :99-101 (once per scale).
 * strotss_eval_grouped replaces the loop of :114-121:
 *     loss = (1/R) * sum_r (alpha * self_similarity(pred_r, content_r) + StyleLoss_r(pred_r)) / loss_denom
 * scalars: device float[STROTSS_NUM_SCALARS], every slot the mean over regions (TOTAL, LOSS_C, LOSS_S are the three
 * values train_step returns, :123-125).  region_scalars: optional device float[R][STROTSS_NUM_SCALARS].
 * grad_pred: device (sum_r N_r) x D or NULL; rows of region r receive d loss / d pred_r (including the 1/R).
 * Region 0 is enqueued on `stream` by the calling thread, regions 1..R-1 by library-owned launching threads on per-region
 * streams forked from / joined to `stream` (the first call after new targets, and any call during stream capture, stay on the
 * calling thread); every region has its own workspace, so N_r may change freely between calls.  Single GPU only. */
int strotss_set_style_targets_grouped(strotss_handle h, const float* style, long long ld, const int* offsets_M, int R, int D,
                                      void* stream);
int strotss_eval_grouped(strotss_handle h, const float* pred, long long ld_pred, const float* content, long long ld_content,
                         const int* offsets_N, int R, float alpha, float* scalars, float* region_scalars, float* grad_pred,
                         long long ld_grad, void* stream);

/* Same evaluation with HOST buffers (pinned memory recommended): copies pred and content to the
 * device, runs strotss_eval, copies scalars (and grad if non-NULL) back and synchronises the stream.
 * This is the call bench.py times as `e2e`. */
int strotss_eval_host(strotss_handle h, const float* pred_host, const float* content_host, int N, float alpha,
                      float* scalars_host, float* grad_host, void* stream);

/* Pipelined form of strotss_eval_host for throughput over independent evaluations (BASELINE configs[4], "throughput
 * mode"): submit enqueues the input copies, the evaluation and the read-back on three internal streams and returns
 * a ticket at once; up to two evaluations are in flight, so the PCIe transfers of neighbouring evaluations overlap
 * the kernels (submitting a third collects the oldest first).  wait blocks until that evaluation's scalars_host /
 * grad_host are complete.  Host buffers must stay valid (and should be pinned) until the ticket is collected. */
int strotss_eval_host_submit(strotss_handle h, const float* pred_host, const float* content_host, int N, float alpha,
                             float* scalars_host, float* grad_host, long long* ticket);
int strotss_eval_host_wait(strotss_handle h, long long ticket);

/* StyleLoss.__call__(prediction) alone (run_strotss.py:33-40): scalars slots L_M, L_REMD, L_PALETTE,
 * LOSS_S are written; grad (if non-NULL) is d loss_s / d pred. */
int strotss_style_loss(strotss_handle h, const float* pred, long long ld_pred, int N, float alpha,
                       float* scalars, float* grad_pred, long long ld_grad, void* stream);

/* relaxed_emd(x, y, distance) (nn/losses.py:69-80).  x: M x D target, y: N x D prediction.
 * loss: device float[4] = {loss, R_X, R_Y, branch}.  grad_y: device N x D or NULL (d loss / d y).
 * D == 3 runs the CUDA-core palette kernel for all three distances; D > 3 supports 'cosine' on the
 * tcgen05 path (the only combination on the reference's hot path, run_strotss.py:36,39). */
int strotss_relaxed_emd(strotss_handle h, const float* x, long long ldx, int M, const float* y, long long ldy, int N,
                        int D, int distance, float* loss, float* grad_y, long long ld_grad,
                        int32_t* row_argmin, int32_t* col_argmin, void* stream);

/* moment_matching(x, y) (nn/losses.py:39-52).  loss: device float[3] = {loss, l_cov, l_mean};
 * grad_y: device N x D or NULL. */
int strotss_moment_matching(strotss_handle h, const float* x, long long ldx, int M, const float* y, long long ldy, int N,
                            int D, float* loss, float* grad_y, long long ld_grad, void* stream);

/* self_similarity(x, y) (nn/losses.py:55-66).  x, y: N x D.  loss: device float[1];
 * grad_x: device N x D or NULL (the reference calls it with x = prediction, run_strotss.py:24). */
int strotss_self_similarity(strotss_handle h, const float* x, long long ldx, const float* y, long long ldy, int N,
                            int D, float* loss, float* grad_x, long long ld_grad, void* stream);

/* convert_rgb_to_yuv(x) (nn/strotss_utils.py:166-167): out[n][3] = x[n][0:3] . K_yuv. */
int strotss_convert_rgb_to_yuv(strotss_handle h, const float* x, long long ldx, int n, float* out, void* stream);

/* Sampling._sample(xs, indices, bilinear_sampling) (nn/strotss_utils.py:25-81): gather the hypercolumns of n
 * sample positions from nmaps NHWC feature maps (maps[k]: device pointer to h[k] x w[k] x c[k] fp32) into
 * out (n x sum_k c[k], row stride ld_out).  indices: device (n, 2) fp32 (row, column) in the resolution of
 * map 0; the per-map rescaling of :33-37 is derived from the shapes.  bilinear != 0 selects the 4-tap path.
 * `maps`, `hs`, `ws`, `cs` are HOST arrays.  The result is bit-identical to an fp32 evaluation of the
 * reference op sequence. */
int strotss_sample(strotss_handle h, int nmaps, const float* const* maps, const int* hs, const int* ws, const int* cs,
                   const float* indices, int n, int bilinear, float* out, long long ld_out, void* stream);
/* Its backward (what tape.gradient does for the tf.gather calls at :67-70,75): scatter-ADD grad_out through the
 * same taps into grad_maps[k] (device buffers shaped like the maps, zeroed by the caller; NULL entries are skipped). */
int strotss_sample_backward(strotss_handle h, int nmaps, float* const* grad_maps, const int* hs, const int* ws,
                            const int* cs, const float* indices, int n, int bilinear, const float* grad_out,
                            long long ld, void* stream);

/* ---- pixel-side step (SURVEY 8f "next #3"): images are NHWC fp32 with batch 1, element (y, x, ch) at (y*w + x)*c + ch ----
 *
 * tf.image.resize(x, (oh, ow)) with the default bilinear method (TF2: half-pixel centres, no antialiasing), as used by
 * utils.resize / utils.resize_like (nn/utils.py:32-41) and by the pyramid functions below. */
int strotss_resize_bilinear(strotss_handle h, const float* src, int sh, int sw, int c, float* out, int oh, int ow, void* stream);
/* make_laplacian(x, return_downscale=True) (nn/strotss_utils.py:139-146): down = resize(x, max(hw//2, 1)) (device buffer of
 * max(h//2,1) x max(w//2,1) x c), pyr = x - resize(down, hw).  make_laplacian_pyramid (:149-156) is `levels` such calls. */
int strotss_make_laplacian(strotss_handle h, const float* x, int hh, int ww, int c, float* pyr, float* down, void* stream);
/* fold_laplacian_pyramid(xs) (nn/strotss_utils.py:159-163): ret = xs[-1]; for x in reversed(xs[:-1]): ret = x + resize(ret,
 * shape(x)).  xs / hs / ws are HOST arrays over the nlev (<= 8) levels, finest first; out is hs[0] x ws[0] x c. */
int strotss_pyramid_fold(strotss_handle h, int nlev, const float* const* xs, const int* hs, const int* ws, int c, float* out,
                         void* stream);
/* What tape.gradient does for the fold (run_strotss.py:122,141): grad_xs[k] (device, shaped like xs[k]) = d loss / d xs[k]
 * given grad_out = d loss / d image.  Deterministic (gather form of the transposed resize, no atomics). */
int strotss_pyramid_fold_backward(strotss_handle h, int nlev, const int* hs, const int* ws, int c, const float* grad_out,
                                  float* const* grad_xs, void* stream);
/* opt.apply_gradients(zip(grads, st_variables)) with tf.keras.optimizers.RMSprop(rho, epsilon, learning_rate)
 * (run_strotss.py:63,148; momentum 0, not centred): rms = rho*rms + (1-rho)*g^2; var -= lr*g/(sqrt(rms) + eps), for all
 * nvars (<= 8) variables in one launch.  vars / rms / grads / counts are HOST arrays; counts = elements per variable. */
int strotss_rmsprop_step(strotss_handle h, int nvars, float* const* vars, float* const* rms, const float* const* grads,
                         const long long* counts, float lr, float rho, float eps, void* stream);

/* Test hook: C[m][n] = alpha * sum_k bf16(A[m][k]) * bf16(B[n][k]) through the tcgen05 GEMM core
 * (fp32 in, fp32 out; tile_n is 128 or 256).  Not part of the reference interface. */
int strotss_debug_gemm(strotss_handle h, const float* A, int m, const float* B, int n, int k, float alpha,
                       float* C, int tile_n, void* stream);

/* Test hook for the MN-major (transposed-A) operand path: At is k x m; C (+)= alpha * At^T . B^T.
 * `accumulate` is a bit set: 1 = add to C; 2 = run the CTA-pair kernel with the MN-major A operand (how stage 2 of the
 * self-similarity reads the row-major x^); 4 = CTA-pair kernel, Gram matrix C = alpha * At^T At with BOTH operands
 * MN-major from the one matrix (the covariance; needs n == m, B is not read). */
int strotss_debug_gemm_ta(strotss_handle h, const float* At, int m, const float* B, int n, int k, float alpha,
                          float* C, int accumulate, void* stream);

/* Test hooks without GPU work (CPU tests of the host logic).  strotss_debug_tile_walk replays the tile order of the
 * persistent kernels -- walk 0: rectangular raster in groups of group_n column tiles; 1: upper block triangle
 * (covariance); 2: block trapezoid tn >= tm of a symmetric row panel (self-similarity stage 1) -- returns the tile count
 * and writes up to `capacity` (tm, tn) pairs in visiting order.  strotss_debug_couples_pay returns 1 where the host
 * sends a GEMM of tiles_m128 x tiles_n256 tiles and `kblocks` K blocks to skewed tile couples.  Not part of the
 * reference interface. */
int strotss_debug_tile_walk(int walk, int tiles_m, int tiles_n, int group_n, int* tm_out, int* tn_out, int capacity);
int strotss_debug_couples_pay(int num_sms, int tiles_m128, int tiles_n256, int kblocks, int skew);
/* Work split of the row-sharded symmetric self-similarity (csrc/ss_jobs.h): returns 1 and the rank's jobs
 * (r0, r1, c0, c1, diag, kind each; at most 8), the row ranges of mirrored stage-2 products it sends / receives
 * (peer, r0, r1 each; at most 8) and the three counts -- or 0 if (N, world, panel) falls back to rectangular sharding. */
int strotss_debug_ss_jobs(int N, int world, int rank, int panel, int* jobs6, int* sends3, int* recvs3, int* counts3);

/* Sign-block copies of the same work split when the bf16 blocks of the mirrored tiles travel to the ranks that own their rows
 * (CUDA-IPC peer windows, csrc/ss_jobs.h): returns the number of copies of `rank` (-1: the split falls back) and, per copy,
 * (job, peer, element offset in the peer's window, row length there, i0, i1, j0, j1); *window_elems = bf16 elements of the
 * rank's own window. */
int strotss_debug_ss_copies(int N, int world, int rank, int panel, long long* copies8, int capacity, long long* window_elems);

#ifdef __cplusplus
}
#endif
#endif /* STROTSS_B200_H */
