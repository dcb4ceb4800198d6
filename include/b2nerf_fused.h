/*
 * b2nerf_fused.h — C ABI of the fused fast path of libb2nerf.so.  No reference counterpart: these entry points run the
 * same math as nerf_triplane/network.py:252-311 (NeRFNetwork.forward/density) and nerf_triplane/renderer.py:442-561
 * (run_cuda_for_inference) in a handful of launches instead of ~45 per loop iteration.  Conventions as in b2nerf.h.
 */
#ifndef B2NERF_FUSED_H_
#define B2NERF_FUSED_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Head-model geometry, fixed by network.py:129-152: three D=2 L=12 C=1 hash grids (xy, yz, xz), audio dim 32,
 * eye dim 1, individual-code dim 4, SH degree 4.  Weights are nn.Linear layout [out,in], fp32, no bias. */
typedef struct {
    /* tri-plane tables: each [offsets[12], 1] fp32 (network.py:129-133); offsets shared (grid.py:111-123) */
    const float   *table_xy, *table_yz, *table_xz;
    const int32_t *offsets;          /* [13], device */
    float          S;                /* log2(per_level_scale), grid.py:32 */
    uint32_t       H;                /* base resolution (64) */
    float          bound;            /* scene bound (1) */
    /* MLPs (network.py:141-152) */
    const float *aud_att_w0, *aud_att_w1;        /* aud_ch_att_net 36->64->32 */
    const float *eye_att_w0, *eye_att_w1;        /* eye_att_net    36->16->1  */
    const float *sigma_w0, *sigma_w1, *sigma_w2; /* sigma_net      69->64->64->65 */
    const float *color_w0, *color_w1;            /* color_net      84->64->3  */
    const float *unc_w0, *unc_w1;                /* unc_net        36->32->1 (both NULL => testing: unc = log 2) */
} b2n_head_weights;

typedef struct b2n_model b2n_model;   /* opaque: fp16 weight image in the tensor-core operand layout */

int  b2n_model_create(b2n_model **out, void *stream);
/* (re)packs the weights AND the tables on `stream` (four small kernels): the MLPs into the fp16 operand images, the three tri-plane tables into
 * the corner-quad image the fused gather reads (one 16-byte entry per cell = its four corner values, hashed levels de-hashed; 39.6 MB for the
 * reference geometry).  Call it after every change to weights or tables.  The table / offsets pointers are also kept (not copied) for the backward
 * kernels.  Environment B2N_HEAD_QUADS=0 (read at b2n_model_create) keeps the gather on the reference-format tables instead. */
int  b2n_model_update(b2n_model *m, const b2n_head_weights *w, void *stream);
void b2n_model_destroy(b2n_model *m);

/* Per-sample network forward = NeRFNetwork.forward (network.py:252-281) on M samples, autocast(fp16) numerics.
 * xyzs/dirs [M,3]; enc_a [32]; ind_code [4] or NULL; eye [1] or NULL (device).  Outputs (any may be NULL): sigmas [M],
 * rgbs [M,3], amb_aud [M] (L2 norm of the audio channel attention), amb_eye [M], unc [M] (softplus; log 2 when testing).
 * n_valid: optional device int32 — only the first min(M, *n_valid) samples are evaluated. */
int b2n_head_forward(const b2n_model *m, const float *xyzs, const float *dirs, uint32_t M,
                     const float *enc_a, const float *ind_code, const float *eye, const int32_t *n_valid,
                     float *sigmas, float *rgbs, float *amb_aud, float *amb_eye, float *unc, void *stream);

/* Audio prologue = NeRFNetwork.encode_audio with att > 0 (network.py:226-240): AudioNet (network.py:40-70) on each of the 8 frames of the
 * window, then AudioAttNet (network.py:9-36).  One kernel, one thread-block cluster of 8 CTAs, autocast(fp16) numerics.
 * Weights are the torch parameters in place (fp32, nn.Conv1d [out,in,3] / nn.Linear [out,in] layouts).
 * auds [8, dim_in, L] fp32 (HuBERT: dim_in 1024, L 2; DeepSpeech: dim_in 29, L 16); enc_a [32] fp32. */
typedef struct {
    const float *conv_w[4], *conv_b[4];          /* audio_net.encoder_conv.{0,2,4,6}   dim_in->32->32->64->64, k3 s2 p1 */
    const float *fc_w[2], *fc_b[2];              /* audio_net.encoder_fc1.{0,2}        64->64->32 */
    const float *att_conv_w[5], *att_conv_b[5];  /* audio_att_net.attentionConvNet.{0,2,4,6,8}  32->16->8->4->2->1, k3 s1 p1 */
    const float *att_fc_w, *att_fc_b;            /* audio_att_net.attentionNet.0       Linear(8, 8) */
    uint32_t dim_in;
} b2n_audio_weights;
int b2n_audio_encode(const b2n_audio_weights *w, const float *auds, uint32_t L, float *enc_a, void *stream);
/* The same with the reference's `smooth_lips` option (renderer.py:456-460; the serving config turns it on): lips_state is a device float[33] carried
 * across consecutive frames — [0..31] the previous frame's smoothed code, [32] a valid flag (zero it to start a sequence).  When valid,
 * enc_a = lambda * state + (1 - lambda) * enc_a (lambda = 0.35 in the reference); the result is written to enc_a AND back to the state.
 * lips_state NULL = b2n_audio_encode. */
int b2n_audio_encode_smooth(const b2n_audio_weights *w, const float *auds, uint32_t L, float *enc_a, float *lips_state, float lambda, void *stream);
/* Backward of b2n_audio_encode (training): d_enc_a [32] -> the gradient of every parameter, accumulated into `g` (same layouts as the weights; zero
 * them first).  One cluster kernel: recomputes the forward, back-propagates through the attention net on CTA 0 and through the eight AudioNet copies. */
typedef struct {
    float *conv_w[4], *conv_b[4];
    float *fc_w[2], *fc_b[2];
    float *att_conv_w[5], *att_conv_b[5];
    float *att_fc_w, *att_fc_b;
} b2n_audio_grads;
int b2n_audio_backward(const b2n_audio_weights *w, const float *auds, uint32_t L, const float *d_enc_a, const b2n_audio_grads *g, void *stream);

/* One whole inference frame = renderer.py:442 (near/far) + :480-545 (march / network / composite loop with compaction)
 * + :559-561 (background blend, clamp), with NO host synchronisation: alive-ray counts and n_step live in a device-side
 * control block; the host enqueues a fixed launch sequence (capturable in a CUDA graph).  perturb is off (inference).
 * rays_o/rays_d [N,3]; bitfield [cascade*grid_size^3/8]; bg_color [N,3] or NULL (=> white); image_out [N,3] fp32;
 * weights_sum_out [N] or NULL; depth_out [N] or NULL.  workspace: 256-byte aligned device scratch of
 * b2n_render_frame_workspace_bytes(N) bytes, owned by the caller. */
typedef struct {
    float    bound, dt_gamma, min_near, T_thresh, density_scale;
    uint32_t max_steps, cascade, grid_size;
    float    aabb[6];
    uint32_t head_ctas;   /* 0: the frame has the GPU to itself — each network launch uses every SM (lowest latency).  > 0: frames from other streams
                           * are in flight beside this one — a network launch takes at most head_ctas SMs (about half of them is best on B200), so two
                           * frames' network launches run side by side and the small march / composite launches never queue behind a full-GPU kernel. */
    uint32_t image_width; /* 0: no assumption about the rays.  W > 0: the N rays are the pixels of an image of width W in row-major order (N = H * W; W % 16 == 0,
                           * H % 8 == 0, else ignored) — the frame then walks them in 8 x 16 pixel tiles, which keeps the samples of a network tile close
                           * together in space.  Per-ray results do not depend on the order. */
} b2n_render_cfg;
uint64_t b2n_render_frame_workspace_bytes(uint32_t N);
int b2n_render_frame(const b2n_model *m, const b2n_render_cfg *cfg, const float *rays_o, const float *rays_d,
                     uint32_t N, const uint8_t *bitfield, const float *enc_a, const float *ind_code,
                     const float *eye, const float *bg_color, void *workspace,
                     float *image_out, float *weights_sum_out, float *depth_out, void *stream);

/* The same frame as ONE CUDA graph whose loop is a WHILE conditional node (init -> while (!done) { march, head, composite, compact, advance }
 * -> finish): only the iterations the reference's host loop would run are executed; one graph launch per frame.  All pointers are baked in
 * (static buffers, as with any CUDA graph).  If `audio` is non-NULL the graph starts with b2n_audio_encode(audio, auds, audio_L) -> enc_a;
 * otherwise enc_a is read as is.  Needs CUDA 12.4+ graph conditional nodes; on failure nothing is created and the plain b2n_render_frame
 * sequence (capturable into an ordinary graph) remains available. */
typedef struct b2n_frame_graph b2n_frame_graph;
int b2n_frame_graph_create(b2n_frame_graph **out, const b2n_model *m, const b2n_render_cfg *cfg, const b2n_audio_weights *audio, const float *auds,
                           uint32_t audio_L, float *enc_a, const float *rays_o, const float *rays_d, uint32_t N, const uint8_t *bitfield,
                           const float *ind_code, const float *eye, const float *bg_color, void *workspace, float *image_out, float *weights_sum_out,
                           float *depth_out);
/* Optional device-side frame prologue / epilogue, so that only a pose + the audio window go up and one RGB24 frame comes down:
 *   pose   : device float[16], row-major 4x4 camera-to-world; when non-NULL the graph starts by filling rays_o / rays_d (which must then be
 *            writable [H*W,3] buffers, N == H*W) with get_rays' all-pixel branch (nerf_triplane/utils.py:227-312: pixel centres + 0.5,
 *            direction ((i-cx)/fx, (j-cy)/fy, 1) normalised, rotated by pose[:3,:3], origin pose[:3,3]);
 *   rgb8_out: device uint8[N*3]; when non-NULL the epilogue also writes (image * 255) truncated to uint8 — the bytes the reference pushes to its
 *            frame queue (TrainerUtil.py:668). */
typedef struct {
    const float *pose;
    float fx, fy, cx, cy;
    uint32_t H, W;
    uint8_t *rgb8_out;
} b2n_frame_io;
int b2n_frame_graph_create_io(b2n_frame_graph **out, const b2n_model *m, const b2n_render_cfg *cfg, const b2n_audio_weights *audio, const float *auds,
                              uint32_t audio_L, float *enc_a, float *rays_o, float *rays_d, uint32_t N, const uint8_t *bitfield,
                              const float *ind_code, const float *eye, const float *bg_color, void *workspace, float *image_out, float *weights_sum_out,
                              float *depth_out, const b2n_frame_io *io);
/* the two stages as plain calls (also usable outside a graph) */
int b2n_get_rays(const float *pose, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, float *rays_o, float *rays_d, void *stream);
int b2n_image_to_rgb8(const float *image, uint32_t n_pixels, uint8_t *rgb8_out, void *stream);
int b2n_frame_graph_launch(b2n_frame_graph *fg, void *stream);
int b2n_frame_graph_info(const b2n_frame_graph *fg, uint64_t *kernels_fixed, uint64_t *kernels_per_iteration);
void b2n_frame_graph_destroy(b2n_frame_graph *fg);
/* number of loop iterations executed by the last frame rendered into `workspace` (4-byte read-back; synchronises `stream`) */
int b2n_frame_iterations(const void *workspace, uint32_t N, int32_t *iterations, void *stream);

/* ---- training: fused head forward that also keeps the activations the backward needs ------------------------------------------------
 * All buffers fp16, row-major, row pitches padded to a multiple of 8 halves (16 bytes) with zero padding, so they are consumed directly as
 * operands of b2n_linear_wgrad and of b2n_head_backward.  Every row 0..M-1 is written. */
typedef struct {
    void *x36;    /* [M,40]  tri-plane features enc_x (36) + 4 zeros                      (network.py:215-223) */
    void *ha;     /* [M,64]  relu(aud_ch_att_net.net.0(enc_x)) */
    void *he;     /* [M,16]  relu(eye_att_net.net.0(enc_x)) */
    void *hu;     /* [M,32]  relu(unc_net.net.0(enc_x)); NULL when unc_net is not evaluated */
    void *att;    /* [M,32]  aud_ch_att_net output */
    void *s_in;   /* [M,80]  sigma_net input [enc_x 36 | 0 x4 | enc_a * att 32 | eye * eye_att 1 | 0 x7]   (network.py:293-298) */
    void *h1;     /* [M,64]  relu(sigma_net.net.0) */
    void *h2;     /* [M,64]  relu(sigma_net.net.1) */
    void *c_in;   /* [M,88]  color_net input [sh 16 | geo_feat 64 | ind_code 4 | 0 0 0 0]            (network.py:267-270) */
    void *hc;     /* [M,64]  relu(color_net.net.0) */
    void *misc;   /* [M,8]   sigmoid(rgb logits) x3, eye_att, unc logit, 1.0 (ones column: column sums via b2n_linear_wgrad), 0 0 */
} b2n_head_saved;

/* b2n_head_forward on all M rows + the saved activations (training forward of NeRFNetwork.forward, network.py:252-311). */
int b2n_head_forward_train(const b2n_model *m, const float *xyzs, const float *dirs, uint32_t M, const float *enc_a, const float *ind_code,
                           const float *eye, float *sigmas, float *rgbs, float *amb_aud, float *amb_eye, float *unc,
                           const b2n_head_saved *saved, void *stream);

/* fp16 per-sample gradients produced by b2n_head_backward (same row pitches / zero padding as b2n_head_saved): each is the "dY" operand of one
 * b2n_linear_wgrad call whose "X" operand is a saved activation. */
typedef struct {
    void *d_rl;       /* [M,8]   d rgb logits (3)                                   dW color_net.1  = d_rl^T  hc   */
    void *d_hc;       /* [M,64]  d color hidden (pre-ReLU)                          dW color_net.0  = d_hc^T  c_in */
    void *d_o;        /* [M,72]  d sigma_net output, ROTATED [geo_feat 64 | density logit | 0 x7]   dW sigma_net.2 (rows rotated) = d_o^T h2 */
    void *d_h2;       /* [M,64]                                                     dW sigma_net.1  = d_h2^T  h1   */
    void *d_h1;       /* [M,64]                                                     dW sigma_net.0  = d_h1^T  s_in */
    void *d_ew;       /* [M,32]  d (enc_a * att): d enc_a = sum_m d_ew * att */
    void *d_att;      /* [M,32]                                                     dW aud_att.1    = d_att^T ha   */
    void *d_ha;       /* [M,64]                                                     dW aud_att.0    = d_ha^T  x36  */
    void *d_el;       /* [M,8]   d eye logit (col 0), d e (col 1)                    dW eye_att.1    = d_el^T  he   */
    void *d_he;       /* [M,16]                                                     dW eye_att.0    = d_he^T  x36  */
    void *d_ul;       /* [M,8]   d unc logit (col 0); NULL without unc_net           dW unc_net.1    = d_ul^T  hu   */
    void *d_hu;       /* [M,32]  NULL without unc_net                                dW unc_net.0    = d_hu^T  x36  */
    void *d_ci;       /* [M,8]   d of the ind-code columns of the color input (4): d ind_code = sum_m d_ci */
    float *d_planes;  /* fp32 [3][12][M]: d enc_x in the layout of grid_encode_backward's `grad` ([L,B,1] per plane: xy, yz, xz) */
} b2n_head_grads;

/* Backward-data pass of the head network (network.py:252-311 through autograd) in one kernel: consumes the gradients of the five outputs
 * (any may be NULL = zero) and the activations kept by b2n_head_forward_train, produces b2n_head_grads.  `sigmas` / `amb_aud` are the
 * forward's outputs. */
int b2n_head_backward(const b2n_model *m, uint32_t M, const float *enc_a, const float *eye, const b2n_head_saved *saved, const float *sigmas,
                      const float *amb_aud, const float *g_sigma, const float *g_rgb, const float *g_aud, const float *g_eye, const float *g_unc,
                      const b2n_head_grads *grads, void *stream);

/* Table gradients of the three tri-plane encoders (network.py:208-223) in one launch: grad_planes = b2n_head_grads.d_planes ([3][L][M] fp32),
 * xyz [M,3] the sample positions in [-bound, bound]; grad_xy / grad_yz / grad_xz [sO] fp32 are accumulated into (zero them first).  Same sums as
 * three b2n_grid_encode_backward calls on the xy / yz / xz slices (D = 2, C = 1, hash grid). */
int b2n_triplane_grid_backward(const float *grad_planes, const float *xyz, const int32_t *offsets, float *grad_xy, float *grad_yz, float *grad_xz,
                               uint32_t M, uint32_t L, float S, uint32_t H, float bound, void *stream);

/* Weight gradient of a bias-free Linear over a tall activation matrix:  dw[out,in] += dy[M,out]^T x[M,in]  (fp16 operands, row-major,
 * fp32 accumulation; dw is accumulated into, zero it first).  1 <= out, in <= 128.  Replaces the weight-gradient GEMM that autograd's
 * LinearBackward runs for every MLP layer of nerf_triplane/network.py:73-94 in a training step (csrc/wgrad.cu). */
int b2n_linear_wgrad(const void *dy_f16, const void *x_f16, uint32_t M, uint32_t out_dim, uint32_t in_dim, float *dw, void *stream);
/* Same product accumulated into `replicas` copies of the result, copy r at dw + r * replica_stride floats (0 = out_dim * in_dim); CTA b adds its
 * partial to copy b mod replicas and the caller sums the copies: hundreds of CTAs reducing into ONE small matrix serialise in the L2. */
int b2n_linear_wgrad_replicated(const void *dy_f16, const void *x_f16, uint32_t M, uint32_t out_dim, uint32_t in_dim, float *dw, uint32_t replicas,
                                uint32_t replica_stride, void *stream);
/* Up to 16 such products over the same M rows in ONE launch (all weight gradients of a backward): every CTA streams the jobs back to back through one
 * cp.async ring, two TMEM accumulators alternate so the reduction of job j overlaps the multiplication of job j + 1.  Operands must be 16-byte aligned
 * with widths that are multiples of 8 (the layouts of b2n_head_saved / b2n_head_grads); dw of job j is accumulated into replica (cta mod replicas) at
 * dw + r * replica_stride floats. */
typedef struct {
    const void *dy;      /* f16 [M, out_dim] */
    const void *x;       /* f16 [M, in_dim] */
    float *dw;           /* f32 [replicas][out_dim, in_dim] with replica stride */
    uint32_t out_dim, in_dim;
} b2n_wgrad_job;
int b2n_linear_wgrad_batch(const b2n_wgrad_job *jobs, uint32_t n_jobs, uint32_t M, uint32_t replicas, uint32_t replica_stride, void *stream);

/* Finishing pass for replicated weight gradients (b2n_linear_wgrad_batch): for every block, dst[i * dst_ld + c] += sum over replicas r of
 * src[r * replica_stride + src_off + i * src_ld + c], i < rows, c < cols — ONE launch that sums the replicas and accumulates each (padded, possibly
 * row-rotated or column-sliced) product straight into the storage of the parameter's gradient (what autograd's AccumulateGrad + the slicing / cat
 * kernels of a LinearBackward graph would do, TrainerUtil.py:1040-1046).  A diagonal is a block with cols = 1 and src_ld = ld + 1. */
typedef struct {
    uint32_t src_off, src_ld;     /* first element and row pitch inside one replica (floats) */
    uint32_t rows, cols;
    float   *dst;                 /* accumulated into */
    uint32_t dst_ld;
    uint32_t reserved;
} b2n_wgrad_block;
int b2n_wgrad_scatter(const float *src, uint32_t replicas, uint32_t replica_stride, const b2n_wgrad_block *blocks, uint32_t n_blocks, void *stream);

/* AdamW over one flat fp32 parameter buffer (parameters, gradients and both moments contiguous; two hyper-parameter groups split at
 * n_group0).  Replaces torch.optim.AdamW + the GradScaler unscale pass of the reference's optimizer step (TrainerUtil.py:1040-1056,
 * train.py:274).  `step` is a 1-element device counter (incremented unless found_inf), `grad_scale` / `found_inf` are the GradScaler's
 * device scalars (NULL = no scaling / never skip): the gradient is divided by grad_scale and the whole step is skipped when found_inf != 0. */
int b2n_adamw_flat(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, uint32_t n_group0, float lr0,
                   float weight_decay0, float lr1, float weight_decay1, float beta1, float beta2, float eps, float *step,
                   const float *grad_scale, const float *found_inf, void *stream);

/* The same with up to 8 contiguous hyper-parameter groups (network.py:332-356 get_params: tables lr / AdamW's default weight decay 0.01; networks lr_net / wd;
 * audio_att_net 5 * lr_net / 1e-4), group k covering flat indices [end[k-1], end[k]); the per-step learning rates are whatever the LR scheduler set
 * (train.py:284-288, LambdaLR stepped every iteration, TrainerUtil.py:1048-1049).  ema: optional fp32 [n] shadow copy updated in the same pass,
 * shadow -= (1 - ema_decay) * (shadow - param) — torch_ema's ExponentialMovingAverage.update (TrainerUtil.py:98-99, 1055-1056); NULL = no EMA this step. */
typedef struct {
    uint32_t n_groups;
    uint32_t end[8];
    float lr[8], weight_decay[8];
} b2n_adam_groups;
int b2n_adamw_flat_groups(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, const b2n_adam_groups *groups, float beta1,
                          float beta2, float eps, float *step, const float *grad_scale, const float *found_inf, float *ema, float ema_decay, void *stream);

/* The same step with the GradScaler protocol on the device (torch.amp.GradScaler.step + update, TrainerUtil.py:1046-1047): one non-finite check over the flat
 * gradients, the unscale / overflow skip inside the AdamW kernel, and the scale update behind it.  scaler_state: device float[4] = {scale, growth tracker,
 * found_inf (written by the call), unused}; `grads` 16-byte aligned. */
int b2n_adamw_flat_scaled(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, uint32_t n, const b2n_adam_groups *groups, float beta1,
                          float beta2, float eps, float *step, float *scaler_state, float growth_factor, float backoff_factor, uint32_t growth_interval,
                          float *ema, float ema_decay, void *stream);

/* Training loss of the head branch and its gradient (TrainerUtil.py:238-334 with the background blend of renderer.py:559-561), sf = *step_factor:
 *   img = clamp(image + (1 - weights_sum) * bg, 0, 1);   mse_n = mean_c (img - gt)^2
 *   loss = mean_n [ mse_n * (0.2 + 0.8 clamp((1 - sf) + sf N softmax(unc_sum)_n, 0, 10)) + sf face_n (|img - gt| / (2 (unc+1)^2) + log(unc+1)^2 / 2)
 *                   + 1e-3 sf (1 - face_n) unc_n ]                                          (the three uncertainty terms only when unc_sum != NULL)
 *          + lambda_ent mean H2(clamp(weights_sum, 1e-5, 1 - 1e-5)) + sf lambda_amb mean(aud_sum (1 - face)) + sf lambda_amb mean(eye_sum inv_max_steps aud_sum face)
 * image [N,3], weights_sum / aud_sum / eye_sum / unc_sum [N] are composite_rays_train_triplane's outputs; bg_color [3] or [N,3] (bg_per_ray); face_mask [N] bytes
 * (bool) or NULL (= every ray on the face); step_factor: device scalar (a replayed CUDA graph sees it ramp) or NULL (= step_factor_host).
 * stats: device float[256] — [0] the loss (overwritten), [4..131] per-CTA softmax statistics the backward re-uses.  The backward multiplies by the device scalar
 * *grad_loss (autograd's upstream gradient, i.e. the loss scale); grad_unc_sum may be NULL. */
typedef struct {
    const float *image, *weights_sum, *aud_sum, *eye_sum, *unc_sum;
    const float *gt_rgb, *bg_color;
    const uint8_t *face_mask;
    const float *step_factor;
    float step_factor_host, lambda_ent, lambda_amb, inv_max_steps;
    int bg_per_ray, amb_aud_loss, amb_eye_loss;
} b2n_loss_args;
int b2n_head_loss_forward(const b2n_loss_args *a, uint32_t N, float *stats, void *stream);
int b2n_head_loss_backward(const b2n_loss_args *a, uint32_t N, const float *stats, const float *grad_loss, float *grad_image, float *grad_weights_sum,
                           float *grad_aud_sum, float *grad_eye_sum, float *grad_unc_sum, void *stream);

/* ---- torso branch of a frame = NeRFRenderer.run_torso (renderer.py:572-631) + NeRFNetwork.forward_torso (network.py:170-205) as ONE
 * kernel, inference, autocast(fp16) numerics: 2-D occupancy test (F.grid_sample of density_grid_torso, align_corners=True, > density_thresh), frequency
 * encoding of the pixel coordinate (degree 8), deform MLP 84->32->32->2, tiled grid D=2 L=16 C=2 (half arithmetic), torso MLP 116->32->32->4,
 * sigmoid * 1.002 - 0.001, and the blend over the background.  Weights are the torch parameters in place (fp32, nn.Linear [out,in], no bias).
 *   bg_coords [N,2] in [-1,1] (utils.py:218-223); h_const [50] = [anchor_encoder(wrapped anchors) 42 | individual_codes_torso row 8] (the per-frame
 *   constant inputs of both MLPs, network.py:179-190); bg_color NULL (white), [3] or [N,3] (bg_per_ray);
 *   bg_out [N,3] = torso_color * torso_alpha + bg_color * (1 - torso_alpha) — the bg_color the head composite takes (renderer.py:620,559);
 *   alpha_out [N] / deform_out [N,2] optional (0 outside the occupancy mask). */
typedef struct {
    const float *deform_w0, *deform_w1, *deform_w2;   /* torso_deform_net.net.{0,1,2}.weight  [32,84] [32,32] [2,32] */
    const float *torso_w0, *torso_w1, *torso_w2;      /* torso_net.net.{0,1,2}.weight         [32,116] [32,32] [4,32] */
    const float *table;                               /* torso_encoder.embeddings [offsets[16], 2] fp32 */
    const int32_t *offsets;                           /* [17], device */
    float S;                                          /* log2(per_level_scale) */
    uint32_t H;                                       /* base resolution (16) */
    float torso_shrink;                               /* opt.torso_shrink (0.8) */
} b2n_torso_weights;
int b2n_torso_forward(const b2n_torso_weights *w, const float *bg_coords, uint32_t N, const float *density_grid_torso, uint32_t grid_size,
                      float density_thresh, const float *h_const, const float *bg_color, int bg_per_ray, float *bg_out, float *alpha_out,
                      float *deform_out, void *workspace, void *stream);
/* Training backward of the same branch (network.py:170-205 through autograd) in one kernel: recomputes the forward per pixel, back-propagates through
 * the blend, the sigmoids, torso_net, the tiled grid (table gradients accumulated in fp32 into g_table [sO,2]; gradient of the deformed coordinate through the
 * grid's dy_dx and the clamp) and torso_deform_net, and writes the fp16 operands of the six weight gradients dW = dY^T X (all N rows; zero rows outside the torso
 * mask) for b2n_linear_wgrad_batch: torso_net.{2,1,0} = (dy_t2 [N,8], x_t2 [N,32]) (dy_t1, x_t1 [N,32]) (dy_t0 [N,32], x_t0 [N,120] = [grid 32 | enc_x 34 | h_const 50 | 0 x4]);
 * torso_deform_net.{2,1,0} = (dy_d2 [N,8], x_d2) (dy_d1, x_d1) (dy_d0 [N,32], x_d0 [N,88] = [enc_x 34 | h_const 50 | 0 x4]).  The gradient of h_const is
 * W_t0[:, 66:116]^T colsum(dy_t0) + W_d0[:, 34:84]^T colsum(dy_d0).  g_out [N,3] = gradient of bg_out; g_alpha [N] or NULL = gradient of alpha_out. */
typedef struct {
    void *x_t0, *x_t1, *x_t2, *x_d0, *x_d1, *x_d2;
    void *dy_t2, *dy_t1, *dy_t0, *dy_d2, *dy_d1, *dy_d0;
} b2n_torso_operands;
int b2n_torso_backward(const b2n_torso_weights *w, const float *bg_coords, uint32_t N, const float *density_grid_torso, uint32_t grid_size,
                       float density_thresh, const float *h_const, const float *bg_color, int bg_per_ray, const float *g_out, const float *g_alpha,
                       float *g_table, const b2n_torso_operands *ops, void *workspace, void *stream);
/* workspace: 16-byte aligned device scratch of b2n_torso_workspace_bytes() bytes owned by the caller (packed weight image + tile counter); one per
 * concurrently running call */
uint64_t b2n_torso_workspace_bytes(void);

/* ---- data-parallel gradient exchange over NVLink peer memory (csrc/peer_allreduce.cu) ---------------------------------------
 * One process per GPU on ONE node.  Every rank allocates its flat gradient buffer with b2n_peer_alloc (a cudaMalloc block followed by a flag block, zeroed;
 * `ipc_handle_out` receives the 64-byte CUDA IPC handle), the ranks exchange the handles (torch.distributed all_gather), and b2n_peer_comm_create opens all of
 * them (all_handles = world x 64 bytes in rank order).  b2n_peer_allreduce_mean then replaces `ncclAllReduce(sum) / world` on the buffer with one kernel per
 * rank: barrier, each rank reduces 1/world of the elements over all ranks and writes the mean back into every rank's buffer, barrier — in place, capturable in
 * a CUDA graph (epochs live in device memory).  Every rank must make the same sequence of calls.  A barrier that sees no progress for ~2 s sets the error word
 * (b2n_peer_error) instead of hanging. */
typedef struct b2n_peer_comm b2n_peer_comm;
int  b2n_peer_alloc(uint64_t bytes, void **dev_ptr, void *ipc_handle_out);
int  b2n_peer_free(void *dev_ptr);
int  b2n_peer_comm_create(b2n_peer_comm **out, int rank, int world, void *local_ptr, const void *all_handles, uint64_t bytes);
/* ranks as streams of ONE process: ptrs[world] are this process's own b2n_peer_alloc blocks (no IPC) */
int  b2n_peer_comm_create_local(b2n_peer_comm **out, int rank, int world, void *const *ptrs, uint64_t bytes);
void b2n_peer_comm_destroy(b2n_peer_comm *c);
int  b2n_peer_allreduce_mean(b2n_peer_comm *c, uint64_t n_floats, void *stream);
int  b2n_peer_error(b2n_peer_comm *c, int32_t *host_out, void *stream);

/* ---- measurement probes (csrc/peaks.cu; not on the product path): the rate of random L2-resident gathers / reductions, the ceilings the grid encoder's
 * forward / backward are reported against (profiles/kernel_rooflines.py).  table: n_floats (a power of two) device floats; every thread of `blocks` x 256
 * issues loads_per_thread (multiple of 8) independent accesses of `vec` floats at pseudo-random aligned positions. */
int b2n_probe_gather(const float *table, uint32_t n_floats, uint32_t vec, uint32_t loads_per_thread, uint32_t blocks, float *sink, void *stream);
int b2n_probe_red(float *table, uint32_t n_floats, uint32_t reds_per_thread, uint32_t blocks, void *stream);

/* ---- density-grid maintenance of the head model (csrc/occupancy.cu) ------------------------------------------------
 * NeRFRenderer.mark_untrained_grid (renderer.py:633-697), called once before training (TrainerUtil.py:475): every cell of density_grid
 * [cascade, grid_size^3] (Morton order) that none of the B camera poses (device float [B,4,4], camera-to-world) sees gets -1; other cells are left
 * untouched.  intrinsics as in the reference (fx, fy, cx, cy). */
int b2n_mark_untrained_grid(const float *poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t cascade, uint32_t grid_size, float bound,
                            float *density_grid, void *stream);
/* update_extra_state's query points of cascade `cas` (renderer.py:729-747), written in MORTON order: row m of xyzs [grid_size^3, 3] is the jittered centre
 * of the cell with Morton index m, so the sigma the head kernel returns for row m IS tmp_grid[cas, m].  rand01 [grid_size^3, 3] in [0,1) is read in the
 * reference's own point order ((x * G + y) * G + z) — pass torch.rand_like's output to reproduce its jitter; NULL = cell centres. */
int b2n_density_grid_points(const float *rand01, uint32_t grid_size, uint32_t cas, float bound, float *xyzs, void *stream);
/* update_extra_state's tail (renderer.py:752-766) in two launches without a host read-back: tmp = morton3D_dilation(sigma * density_scale);
 * density_grid = max(density_grid * decay, tmp) where both are >= 0; mean_density = mean(clamp(density_grid, 0)); bitfield = packbits(density_grid,
 * min(mean_density, density_thresh)).  sigma / density_grid [cascade, grid_size^3] fp32 Morton order; stats: 16 bytes of 8-byte aligned device scratch —
 * after the call ((float *)stats)[2] holds mean_density. */
int b2n_density_grid_update(const float *sigma, float density_scale, float *density_grid, uint32_t cascade, uint32_t grid_size, float decay,
                            float density_thresh, uint8_t *bitfield, void *stats, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2NERF_FUSED_H_ */
