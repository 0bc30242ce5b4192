/* calciumgan_b200 — C ABI of the B200-native WGAN-GP training step.
 *
 * This is the boundary that replaces TensorFlow 2.3.1 underneath the reference's
 * `gan/algorithms/wgan_gp.py` + `gan/models/calciumgan.py` (reference paths relative to
 * /root/reference).  Plain pointers and sizes only; no torch / DLPack types in signatures
 * (the Python host resolves DLPack capsules to device pointers before calling).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; cg_last_error() returns a
 *     thread-local message.  A CUDA error is never swallowed and there is no CPU fallback.
 *   - one context per process/GPU; a context is not thread-safe.
 *   - "dev" pointers are device pointers on the context's device, "host" pointers are host.
 *   - all kernels are launched on the stream set by cg_set_stream (default: stream 0).
 *   - signals cross as fp32 NWC (batch, seq_len, channels), exactly what `gan.train(signal)`
 *     receives at main.py:49.
 *   - weights cross as one flat fp32 array per model in Keras get_weights() order and layout
 *     (gan/utils/utils.py:116-152): generator = [dense k(nd, w*nd), b] + 5 x [convT k(K,1,Cout,Cin),
 *     b, (LN gamma, beta)] + [dense k(C,C), b];  critic = 5 x [conv k(K,Cin,Cout), b] + [dense k(w5*5nu,1), b].
 */
#ifndef CALCIUMGAN_B200_H_
#define CALCIUMGAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CG_VERSION 100

enum { CG_GENERATOR = 0, CG_DISCRIMINATOR = 1 };
enum { CG_FP32 = 0, CG_BF16 = 1 };

/* Hyper-parameters: the hparams fields read by gan/models/calciumgan.py:23-27,38-45,96-98,142-151,
 * gan/algorithms/wgan_gp.py:15-17, gan/algorithms/gan.py:18-22 and gan/algorithms/optimizer.py:8-9. */
typedef struct cg_config {
  int32_t seq_len;        /* hparams.signal_shape[0]                          */
  int32_t channels;       /* hparams.num_channels                              */
  int32_t noise_dim;      /* --noise_dim                                       */
  int32_t num_units;      /* --num_units                                       */
  int32_t kernel_size;    /* --kernel_size                                     */
  int32_t strides;        /* --strides (only 2 is implemented)                 */
  int32_t phase_m;        /* --m, PhaseShuffle range                           */
  int32_t layer_norm;     /* --layer_norm                                      */
  int32_t normalize;      /* hparams.normalize -> sigmoid head + denormalised metrics */
  int32_t max_batch;      /* largest per-call batch (smaller batches accepted) */
  int32_t n_critic;       /* --n_critic                                        */
  int32_t precision;      /* CG_FP32 | CG_BF16 (--mixed_precision)             */
  float gp_lambda;        /* --gradient_penalty                                */
  float learning_rate;    /* --learning_rate                                   */
  float signals_min;      /* hparams.signals_min                               */
  float signals_max;      /* hparams.signals_max                               */
  int32_t world_size;     /* data-parallel ranks (gradients are scaled 1/world_size in cg_apply_update) */
  int32_t rank;
  int32_t force_simt;     /* debug: run the CUDA-core kernels even in bf16 mode */
  int32_t debug_flags;    /* CG_DEBUG_* bits: run the unfused reference form of a fused kernel (parity tests) */
  int32_t reserved[6];
} cg_config;

/* cg_config.debug_flags (the environment variables of the same name, read once in cg_create, set the same bits) */
enum {
  CG_DEBUG_NO_PS_FUSE = 1,      /* CG_NO_PS_FUSE: PhaseShuffle gather as its own kernel after the conv GEMM            */
  CG_DEBUG_NO_PS_BWD_FUSE = 2,  /* CG_NO_PS_BWD_FUSE: PhaseShuffle adjoint + slope mask as its own kernel               */
  CG_DEBUG_NO_GHEAD = 4,        /* CG_NO_GHEAD: generic GEMM + interpolation kernel instead of the generator-head kernel */
  CG_DEBUG_NO_ADAM_FUSE = 8     /* CG_NO_ADAM_FUSE: Adam and the bf16 weight re-pack as two kernels                       */
};

typedef struct cg_ctx cg_ctx;

/* scalars written by the step functions (host float arrays) */
enum {
  CG_S_DIS_LOSS = 0, CG_S_GP = 1, CG_S_REAL_LOSS = 2, CG_S_FAKE_LOSS = 3,   /* critic step  */
  CG_S_GEN_LOSS = 4,                                                         /* generator step */
  CG_S_MET_MIN = 5, CG_S_MET_MAX = 6, CG_S_MET_MEAN = 7, CG_S_MET_STD = 8,  /* gan.py:36-41 */
  CG_NUM_SCALARS = 16
};

/* flags for the step functions */
enum {
  CG_FLAG_NO_UPDATE = 1,   /* leave gradients in the flat buffer, do not run Adam (DP host allreduces first) */
  CG_FLAG_NO_SYNC = 2,     /* do not copy scalars back / synchronise (scalars_host may be NULL) */
  CG_FLAG_SAME_REAL = 4,   /* real_dev holds the same batch as in the previous cg_critic_step (wgan_gp.py:85-86):
                              skip its conversion to the compute type */
  CG_FLAG_NO_FAKE32 = 8,   /* cg_critic_step: do not materialise the fp32 generator output (cg_fake_ptr is then stale);
                              the critic sub-steps of a train step only need the compute-type copies */
  CG_FLAG_GEN_PREFETCHED = 16  /* the generator forward of this sub-step (and its noise / alpha draws) already ran in
                              cg_prefetch_generator; noise_dev / alpha_dev are ignored */
};

int cg_version(void);
const char* cg_last_error(void);

int cg_create(const cg_config* cfg, cg_ctx** out);
void cg_destroy(cg_ctx* ctx);
int cg_set_stream(cg_ctx* ctx, void* cuda_stream);
int cg_synchronize(cg_ctx* ctx);

/* ---- parameters: Keras get_weights()/set_weights() (gan/utils/utils.py:124-125,146-147) ---- */
int64_t cg_num_params(cg_ctx* ctx, int which);
int cg_num_tensors(cg_ctx* ctx, int which);
/* shape (up to 4 dims, unused = 0) and flat offset of tensor `idx` in checkpoint order */
int cg_tensor_info(cg_ctx* ctx, int which, int idx, int64_t shape[4], int* ndim, int64_t* offset);
int cg_set_weights(cg_ctx* ctx, int which, const float* host_flat);
int cg_get_weights(cg_ctx* ctx, int which, float* host_flat);
int cg_init_weights(cg_ctx* ctx, uint64_t seed);            /* glorot-uniform / zeros / LN (1, 0) */
/* per-parameter gradients of the last step, before Adam (north_star parity check) */
int cg_get_grads(cg_ctx* ctx, int which, float* host_flat);
/* device pointer of the flat fp32 gradient buffer (the DP host all-reduces it in place) */
void* cg_grad_ptr(cg_ctx* ctx, int which);
/* Gradient buckets for the overlapped data-parallel all-reduce (no reference counterpart). The flat
 * gradient buffer of each model is cut into cg_num_buckets contiguous ranges in the order the backward pass
 * completes them (last layers first). After a step with CG_FLAG_NO_UPDATE, cg_stream_wait_bucket makes `cuda_stream`
 * wait for the last kernel that writes bucket `bucket`, so its all-reduce can overlap the remaining wgrad kernels. */
int cg_num_buckets(cg_ctx* ctx, int which);
int cg_bucket_info(cg_ctx* ctx, int which, int bucket, int64_t* offset, int64_t* count);
int cg_stream_wait_bucket(cg_ctx* ctx, int which, int bucket, void* cuda_stream);
/* Data parallel over NVLink peer memory (no reference counterpart): instead of an NCCL all-reduce, every rank
 * accumulates its gradients into a buffer all peers have mapped (cg_set_grad_buffer: caller-provided symmetric memory,
 * num_params floats, 16-byte aligned; NULL = the library's own buffer again), and after a cross-rank barrier
 * cg_reduce_peer_grads pulls all `world` (2, 4 or 8) buffers with 16-byte peer loads on `cuda_stream` and writes their
 * sum, added in rank order on every rank (bit-identical replicas), into a library-owned buffer that
 * cg_apply_update_reduced feeds to Adam (scaled by 1 / world_size). The kernel is sized to sit beside a resident tensor-
 * core CTA on every SM (128 threads, no shared memory), so it overlaps the next sub-step's generator GEMMs. */
int cg_set_grad_buffer(cg_ctx* ctx, int which, float* grad_dev);
int cg_reduce_peer_grads(cg_ctx* ctx, int which, const void* const* peer_ptrs_host, int world, void* cuda_stream);
/* Two-phase form for 4 / 8 ranks: the reduced buffer is peer-mapped as well (cg_set_reduced_buffer: num_params + 4
 * floats; NULL = the library's own); cg_peer_reduce_scatter sums this rank's slice (num_params / world, rounded up to 4)
 * of every peer's gradients into it, cg_peer_all_gather copies the other slices from their owners. Per rank that pulls
 * 2 (world - 1) / world gradient sizes instead of (world - 1). Barriers: before the first phase and between the phases. */
int cg_set_reduced_buffer(cg_ctx* ctx, int which, float* reduced_dev);
int cg_peer_reduce_scatter(cg_ctx* ctx, int which, const void* const* grad_peer_ptrs_host, int world, int rank, void* cuda_stream);
int cg_peer_all_gather(cg_ctx* ctx, int which, const void* const* reduced_peer_ptrs_host, int world, int rank, void* cuda_stream);
int cg_apply_update_reduced(cg_ctx* ctx, int which);
void* cg_reduced_grad_ptr(cg_ctx* ctx, int which);
/* overwrite the flat gradient buffer (parity test of cg_apply_update alone; Keras get_weights() order) */
int cg_set_grads(cg_ctx* ctx, int which, const float* host_flat);
/* updates skipped because a gradient was not finite (the reference's LossScaleOptimizer skips such steps,
 * optimizer.py:10-12; here the step is skipped and `iterations` is not advanced) */
int64_t cg_skipped_updates(cg_ctx* ctx, int which);
/* Adam moments + iteration counter (optimizer.py:15-21; extra checkpoint keys) */
int cg_get_opt_state(cg_ctx* ctx, int which, float* host_m, float* host_v, int64_t* step);
int cg_set_opt_state(cg_ctx* ctx, int which, const float* host_m, const float* host_v, int64_t step);
int cg_seed(cg_ctx* ctx, uint64_t seed);                    /* noise / alpha / shift streams */

/* ---- the hot path ----
 * real_dev  : (batch, seq_len, channels) fp32 device pointer
 * noise_dev : (batch, noise_dim) fp32 or NULL (drawn on device, gan.py:29-30)
 * alpha_dev : (batch) fp32 or NULL (drawn on device, wgan_gp.py:40)
 * shifts    : host int32 PhaseShuffle draws in call order or NULL (drawn on host, calciumgan.py:121-124)
 */

/* wgan_gp.py:64-80 `_train_discriminator`. shifts_host[12]: D(real), D(fake), D(xhat). */
int cg_critic_step(cg_ctx* ctx, const float* real_dev, int batch, const float* noise_dev,
                   const float* alpha_dev, const int32_t* shifts_host, int flags, float* scalars_host);

/* wgan_gp.py:22-36 `_train_generator` (+ gan.py:32-41 metrics). shifts_host[4]. */
int cg_generator_step(cg_ctx* ctx, const float* real_dev, int batch, const float* noise_dev,
                      const int32_t* shifts_host, int flags, float* scalars_host);

/* Data-parallel overlap (no reference counterpart): runs the generator part of the NEXT sub-step now --
 * for_generator_step = 0: wgan_gp.py:65-66 + the interpolation of :38-41 (fake and x_hat into the critic's input slots);
 * for_generator_step = 1: wgan_gp.py:23-26 (generator forward keeping what its backward needs). It reads no critic
 * weight, so the host enqueues it before waiting for the all-reduce of the current critic update's gradients and calls
 * the next cg_critic_step / cg_generator_step with CG_FLAG_GEN_PREFETCHED. Random streams advance exactly as in the
 * unsplit call. flags: CG_FLAG_NO_FAKE32. */
int cg_prefetch_generator(cg_ctx* ctx, const float* real_dev, int batch, const float* noise_dev, const float* alpha_dev,
                          int for_generator_step, int flags);

/* optimizer.py:31-34 `Optimizer.update` tail: Adam on the flat gradient buffer (scaled by
 * 1/world_size) and refresh of the packed low-precision weight copies. */
int cg_apply_update(cg_ctx* ctx, int which);

/* wgan_gp.py:82-95 `train`: n_critic critic updates on the same batch + one generator update.
 * noise_dev (n_critic+1, batch, noise_dim) | NULL, alpha_dev (n_critic, batch) | NULL,
 * shifts_host[12*n_critic+4] | NULL.  scalars: gen_loss, mean dis_loss, mean gp, 4 metrics. */
int cg_train_step(cg_ctx* ctx, const float* real_dev, int batch, const float* noise_dev,
                  const float* alpha_dev, const int32_t* shifts_host, float* scalars_host);

/* gan.py:87-90 `validate` (= _step(training=False) with WGAN-GP losses). shifts_host[12].
 * fake_out_dev (batch, seq_len, channels) fp32 or NULL. */
int cg_validate(cg_ctx* ctx, const float* real_dev, int batch, const float* noise_dev,
                const float* alpha_dev, const int32_t* shifts_host, float* fake_out_dev,
                float* scalars_host);

/* Batch assembly from a device-resident dataset cache (the reference caches its dataset, dataset_helper.py:171
 * `train_ds.cache()`, and draws shuffled batches from it): dst_dev[i, :] = src_dev[idx_dev[i], :] for i < n, rows of
 * row_elems floats (row_elems % 4 == 0, 16-byte aligned pointers), idx_dev int64 on the device, 0 <= idx < n_src.
 * Asynchronous on the context stream. */
int cg_gather_rows(cg_ctx* ctx, const float* src_dev, int64_t n_src, const int64_t* idx_dev, int n, int64_t row_elems,
                   float* dst_dev);

/* gan.py:32-41 `metrics(real, fake)` on caller tensors (batch, seq_len, channels) fp32: out_host[4] = mean squared
 * difference of the per-timestep min / max / mean / std over neurons (signals_metrics.py:9-28), after de-normalisation. */
int cg_metrics(cg_ctx* ctx, const float* real_dev, const float* fake_dev, int batch, float* out_host);

/* gan.py:92-97 `generate(noise, denorm)`. out_dev (batch, seq_len, channels) fp32. */
int cg_generate(cg_ctx* ctx, const float* noise_dev, int batch, int denorm, float* out_dev);

/* ---- parity / debug taps ---- */
/* critic forward on arbitrary input: scores_dev (batch) fp32. shifts_host[4]. */
int cg_debug_critic_forward(cg_ctx* ctx, const float* x_dev, int batch, const int32_t* shifts_host,
                            float* scores_dev);
/* gradient penalty of the critic at xhat: grad_dev (batch, seq_len, channels) fp32 | NULL,
 * norms_dev (batch) fp32 | NULL. shifts_host[4]. */
int cg_debug_gp(cg_ctx* ctx, const float* xhat_dev, int batch, const int32_t* shifts_host,
                float* grad_dev, float* norms_dev);
/* The gradient penalty alone (BASELINE.json configs[4]; wgan_gp.py:43-50 under the outer tape of optimizer.py:32): critic
 * forward at xhat, g = dD/dxhat, GP = mean_b (||g_b|| - 1)^2 and gp_lambda * dGP/dW for every critic tensor through the
 * double backward -- forward, data-gradient chain, linearised forward, weight gradients; the second-order graph is never
 * built. Gradients land in the critic's flat gradient buffer (cg_get_grads), GP in scalars_host[CG_S_GP]. shifts_host[4].
 * flags: CG_FLAG_NO_SYNC. */
int cg_gp_gradient(cg_ctx* ctx, const float* xhat_dev, int batch, const int32_t* shifts_host, int flags,
                   float* scalars_host);
/* One conv layer in isolation, fp32 in / fp32 out (cast to the compute type inside), for kernel-level
 * parity: which = CG_DISCRIMINATOR (Conv1D layer 1..5, calciumgan.py:145-185; forward includes bias +
 * LeakyReLU) or CG_GENERATOR (Conv1DTranspose layer 1..5, models/utils.py:79-89; forward includes bias).
 *   pass 0 forward : x (batch, Lin, Cin)            -> out (batch, Lout, Cout)
 *   pass 1 dgrad   : dy (batch, Lout, Cout)         -> out (batch, Lin, Cin)
 *   pass 2 wgrad   : x and dy                       -> out = kernel gradient in its Keras layout
 * Critic layers accept batch <= 3*max_batch, generator layers batch <= max_batch. */
int cg_debug_layer(cg_ctx* ctx, int which, int layer, int pass, const float* x_dev, const float* dy_dev,
                   int batch, float* out_dev);
/* PhaseShuffle alone (calciumgan.py:117-138) on a fp32 (batch, w, ch) device tensor, ch % 4 == 0;
 * bit-exact index arithmetic. cg_phase_shuffle_index fills host idx[w] with the source row of each
 * output row (no GPU needed). */
int cg_debug_phase_shuffle(cg_ctx* ctx, const float* x_dev, int batch, int w, int ch, int shift, float* out_dev);
/* Internal activation buffer of the last step as an unpadded fp32 tensor (bf16 -> fp32 is exact, so bit-level
 * comparisons of the stored values are possible). Critic buffers hold `batch` <= 3*max_batch samples in group order
 * [real; fake; xhat]; layer 0 = network input.
 *   CG_BUF_X   critic layer input  X[l]  (batch, L/2^l, C_l)   l = 0..5  (l >= 1: PhaseShuffle output, calciumgan.py:151)
 *   CG_BUF_H   critic activation   H[l]  (batch, L/2^l, C_l)   l = 1..5  (LeakyReLU output; its sign is the slope mask)
 *   CG_BUF_DA  critic d/d(pre-activation) of layer l           l = 1..5
 *   CG_BUF_HG  generator activation HG[i] (batch, w*2^i, C_i)  i = 0..5  (LeakyReLU output)
 *   CG_BUF_AG  generator pre-norm conv-transpose output        i = 1..5
 *   CG_BUF_DAG generator d/d(conv-transpose output)            i = 1..5  */
enum { CG_BUF_X = 0, CG_BUF_H = 1, CG_BUF_DA = 2, CG_BUF_HG = 3, CG_BUF_AG = 4, CG_BUF_DAG = 5 };
int cg_debug_read(cg_ctx* ctx, int buffer, int layer, int batch, float* out_dev);
/* shape of that tensor without the batch dimension */
int cg_debug_buffer_shape(cg_ctx* ctx, int buffer, int layer, int64_t* rows, int64_t* channels);
/* The random draws of the last step function, whether injected or drawn by the library: noise (n_noise floats,
 * device), alpha (n_alpha floats, device), shifts (n_shifts int32, host). Any pointer may be NULL. Returns an error
 * if more values are requested than the last step used. */
int cg_debug_last_draws(cg_ctx* ctx, float* noise_out_dev, int64_t n_noise, float* alpha_out_dev, int64_t n_alpha,
                        int32_t* shifts_out_host, int n_shifts);
/* Data gradient of critic conv layer `layer` (2..5) followed by the PhaseShuffle adjoint and the LeakyReLU slope of the
 * layer below (the step between DA[l] and DA[l-1] of the backward chain), in isolation:
 *   dy (batch, L/2^l, C_l) fp32, h (batch, L/2^(l-1), C_(l-1)) fp32 (sign = slope source), group_b samples per shift,
 *   shifts_host[ceil(batch / group_b)] -> out (batch, L/2^(l-1), C_(l-1)) fp32.
 * Runs the fused epilogue (EPI_PS_MASK) or, with CG_DEBUG_NO_PS_BWD_FUSE, the two-kernel form. */
int cg_debug_dgrad_ps(cg_ctx* ctx, int layer, const float* dy_dev, const float* h_dev, int batch, int group_b,
                      const int32_t* shifts_host, float* out_dev);
int cg_phase_shuffle_index(int w, int shift, int32_t* idx_host);
/* The same map in the two forms the fused tensor-core epilogues use (host copies of the device functions, so the
 * index arithmetic of the hot path can be checked bit for bit without a GPU):
 *   scatter form (forward): source row q is output row t1[q] and, when reflected, also t2[q] (-1 = none);
 *   adjoint plan (data gradient): accumulator row t is stored to output row dest[t] (-1: pushed over an edge), deposits
 *   its value in exchange slot src_slot[t] and adds exchange slot par_slot[t] (slots are per row parity; -1 = none);
 *   zero[t] = 1 when no row maps to output row t. */
int cg_phase_shuffle_scatter_index(int w, int shift, int32_t* t1_host, int32_t* t2_host);
int cg_phase_shuffle_adjoint_plan(int w, int shift, int32_t* dest_host, int32_t* src_slot_host, int32_t* par_slot_host,
                                  int32_t* zero_host);
/* device pointer to the generator output (batch, seq_len, channels) fp32 of the last step */
void* cg_fake_ptr(cg_ctx* ctx);
/* device pointer to the critic scores of the last critic forward (groups x batch) fp32 */
void* cg_scores_ptr(cg_ctx* ctx);
/* device pointer to the CG_NUM_SCALARS floats the last step wrote (valid with CG_FLAG_NO_SYNC) */
void* cg_scalars_ptr(cg_ctx* ctx);
/* number of kernels this library launched since creation (bench "gpu_launches") */
int64_t cg_launch_count(cg_ctx* ctx);
/* how many of those were tcgen05 tensor-core kernels */
int64_t cg_tc_launch_count(cg_ctx* ctx);
/* bytes of device memory owned by the context */
int64_t cg_device_bytes(cg_ctx* ctx);
/* Live kernel timing: while enabled, every implicit-GEMM launch is bracketed by CUDA events on the
 * context stream. cg_profile_report synchronises and fills out[12] =
 * {conv/dgrad GEMM ms, its algorithmic FLOPs, its launches, wgrad GEMM ms, FLOPs, launches,
 *  generator-head kernel ms, FLOPs, launches, 0, 0, 0}
 * accumulated since cg_profile(ctx, 1), then clears the accumulators. */
int cg_profile(cg_ctx* ctx, int enable);
int cg_profile_report(cg_ctx* ctx, double out[12]);
/* The same accumulators as text, one line per kernel -- the tensor-core kernels and every memory-bound kernel of the
 * step (layer norm, PhaseShuffle adjoint, bias column sums, Adam, metrics, ...):
 *   name <tab> tensor|hbm <tab> launches <tab> milliseconds <tab> algorithmic FLOPs <tab> algorithmic bytes <newline>
 * Writes at most cap - 1 characters + NUL, then clears the accumulators (use either report, not both). */
int cg_profile_report_text(cg_ctx* ctx, char* out, int cap);
/* microbenchmark hook: time `iters` launches of one conv layer kernel with CUDA events on the
 * context stream. which/layer: CG_DISCRIMINATOR conv 1..5 or CG_GENERATOR convT 1..5; pass: 0 fwd,
 * 1 dgrad, 2 wgrad. Writes avg milliseconds and the algorithmic FLOPs of one launch. */
int cg_bench_layer(cg_ctx* ctx, int which, int layer, int pass, int batch, int iters,
                   float* ms_out, double* flops_out);

#ifdef __cplusplus
}
#endif
#endif /* CALCIUMGAN_B200_H_ */
