/*
 * b2nerf.h — C ABI of libb2nerf.so: B200-native (sm_100a) kernels for the NeRF render/train hot path of
 * GithinjiHans/LZZX-NeRF (tri-plane hash-grid encode -> occupancy-grid ray march -> fused small MLPs ->
 * alpha composite).  One entry point per function of the reference's four pybind modules, same argument
 * order and meaning, with tensors flattened to (device pointer[, dtype]) and an explicit CUDA stream:
 *
 *   reference module (file:line)                             this header
 *   _raymarching_face  raymarching/src/raymarching.h:7-38    b2n_<same name>            (22 functions)
 *   _gridencoder       gridencoder/src/gridencoder.h:12-13   b2n_grid_encode_{forward,backward}
 *   _shencoder         shencoder/src/shencoder.h:9-10        b2n_sh_encode_{forward,backward}
 *   _freqencoder       freqencoder/src/freqencoder.h:7,10    b2n_freq_encode_{forward,backward}
 *
 * plus the fused fast path (b2n_model_*, b2n_render_*, b2n_train_*) that has no reference counterpart:
 * it runs the same math as renderer.py:406-570 / :279-304 + network.py:252-311 in a few launches.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; all float tensors are fp32 unless a
 *     b2n_dtype argument says otherwise; layouts are the reference's (row-major, contiguous).
 *   - ownership follows the reference: the caller allocates every output; buffers the reference expects
 *     pre-zeroed (xyzs/dirs/deltas of the marchers, all grad_* outputs) must be pre-zeroed by the caller.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, what the reference uses).
 *   - return value: 0 on success, non-zero on error; b2n_last_error() returns a thread-local message.
 *     Argument validation mirrors gridencoder.cu:425-441 (the raymarching entry points of the reference
 *     validate nothing; here null pointers and unsupported shapes are rejected).
 *   - no entry point allocates device memory visible to the caller, synchronises the device, or touches
 *     the host.  The two marchers need a little scratch (the occupied box of the bitfield, per-CTA totals, a per-sample
 *     t cache): the entry points with the reference's argument lists take it from a library-internal grow-only block per
 *     device (calls on one device must then be stream-ordered, as the reference's legacy-stream kernels are); the *_ws
 *     variants take a caller-owned workspace and are safe on concurrent streams / inside CUDA graphs.
 */
#ifndef B2NERF_H_
#define B2NERF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2N_VERSION 100

typedef enum { B2N_F32 = 0, B2N_F16 = 1 } b2n_dtype;

/* ---- library ------------------------------------------------------------------------------------ */
int         b2n_version(void);
const char *b2n_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
uint64_t    b2n_launch_count(void);

/* ---- raymarching: utils (raymarching.h:7-12) ------------------------------------------------------ */
int b2n_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                           float min_near, float *nears, float *fars, void *stream);
int b2n_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream);
int b2n_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, void *stream);
int b2n_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, void *stream);
int b2n_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream);
int b2n_morton3D_dilation(const float *grid, uint32_t C, uint32_t H, float *grid_dilation, void *stream);

/* ---- raymarching: train (raymarching.h:14-17) ----------------------------------------------------- */
/* rays[N,3] = (ray id, sample offset, sample count).  Allocation is DETERMINISTIC: row n describes ray n
 * and offsets are the exclusive prefix sum of counts in ray order (a legal ordering of the reference's
 * atomicAdd allocation, raymarching.cu:446-454).  counter[0] += total samples, counter[1] += N. */
int b2n_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                         float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                         const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                         int32_t *rays, int32_t *counter, const float *noises, void *stream);
/* the same with a caller-owned 16-byte aligned device workspace of b2n_march_rays_train_workspace_bytes(N, max_steps) bytes */
uint64_t b2n_march_rays_train_workspace_bytes(uint32_t N, uint32_t max_steps);
int b2n_march_rays_train_ws(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                            float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                            const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                            int32_t *rays, int32_t *counter, const float *noises, void *workspace, void *stream);
int b2n_march_rays_train_backward(const float *grad_xyzs, const float *grad_dirs, const int32_t *rays,
                                  const float *deltas, uint32_t N, uint32_t M, float *grad_rays_o,
                                  float *grad_rays_d, void *stream);

/* The four training composites (raymarching.h:16-17, 23-24, 30-31, 35-36).  `ambient` semantics:
 *   plain / uncertainty / triplane: ambient_sum = sum of ambient over composited samples (unweighted)
 *   sigma:                           ambient_sum = sum of weight * ambient */
int b2n_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *ambient,
        const float *deltas, const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
        float *weights_sum, float *ambient_sum, float *depth, float *image, void *stream);
int b2n_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_ambient_sum,
        const float *grad_image, const float *sigmas, const float *rgbs, const float *ambient,
        const float *deltas, const int32_t *rays, const float *weights_sum, const float *ambient_sum,
        const float *image, uint32_t M, uint32_t N, float T_thresh,
        float *grad_sigmas, float *grad_rgbs, float *grad_ambient, void *stream);
int b2n_composite_rays_train_sigma_forward(const float *sigmas, const float *rgbs, const float *ambient,
        const float *deltas, const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
        float *weights_sum, float *ambient_sum, float *depth, float *image, void *stream);
int b2n_composite_rays_train_sigma_backward(const float *grad_weights_sum, const float *grad_ambient_sum,
        const float *grad_image, const float *sigmas, const float *rgbs, const float *ambient,
        const float *deltas, const int32_t *rays, const float *weights_sum, const float *ambient_sum,
        const float *image, uint32_t M, uint32_t N, float T_thresh,
        float *grad_sigmas, float *grad_rgbs, float *grad_ambient, void *stream);
int b2n_composite_rays_train_uncertainty_forward(const float *sigmas, const float *rgbs, const float *ambient,
        const float *uncertainty, const float *deltas, const int32_t *rays, uint32_t M, uint32_t N,
        float T_thresh, float *weights_sum, float *ambient_sum, float *uncertainty_sum, float *depth,
        float *image, void *stream);
int b2n_composite_rays_train_uncertainty_backward(const float *grad_weights_sum, const float *grad_ambient_sum,
        const float *grad_uncertainty_sum, const float *grad_image, const float *sigmas, const float *rgbs,
        const float *ambient, const float *uncertainty, const float *deltas, const int32_t *rays,
        const float *weights_sum, const float *ambient_sum, const float *uncertainty_sum, const float *image,
        uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas, float *grad_rgbs, float *grad_ambient,
        float *grad_uncertainty, void *stream);
int b2n_composite_rays_train_triplane_forward(const float *sigmas, const float *rgbs, const float *amb_aud,
        const float *amb_eye, const float *uncertainty, const float *deltas, const int32_t *rays,
        uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *amb_aud_sum, float *amb_eye_sum,
        float *uncertainty_sum, float *depth, float *image, void *stream);
int b2n_composite_rays_train_triplane_backward(const float *grad_weights_sum, const float *grad_amb_aud_sum,
        const float *grad_amb_eye_sum, const float *grad_uncertainty_sum, const float *grad_image,
        const float *sigmas, const float *rgbs, const float *amb_aud, const float *amb_eye,
        const float *uncertainty, const float *deltas, const int32_t *rays, const float *weights_sum,
        const float *amb_aud_sum, const float *amb_eye_sum, const float *uncertainty_sum, const float *image,
        uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas, float *grad_rgbs, float *grad_amb_aud,
        float *grad_amb_eye, float *grad_uncertainty, void *stream);

/* ---- raymarching: inference (raymarching.h:19-21, 26, 32, 38) -------------------------------------- */
int b2n_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                   const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                   uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                   float *xyzs, float *dirs, float *deltas, const float *noises, void *stream);
/* the same with a caller-owned 16-byte aligned device workspace of b2n_march_rays_workspace_bytes() bytes (NULL = no empty-space clipping) */
uint64_t b2n_march_rays_workspace_bytes(void);
int b2n_march_rays_ws(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                      const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                      uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                      float *xyzs, float *dirs, float *deltas, const float *noises, void *workspace, void *stream);
int b2n_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
        const float *sigmas, const float *rgbs, const float *deltas,
        float *weights_sum, float *depth, float *image, void *stream);
int b2n_composite_rays_ambient(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
        float *rays_t, const float *sigmas, const float *rgbs, const float *deltas, const float *ambients,
        float *weights_sum, float *depth, float *image, float *ambient_sum, void *stream);
int b2n_composite_rays_ambient_sigma(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
        float *rays_t, const float *sigmas, const float *rgbs, const float *deltas, const float *ambients,
        float *weights_sum, float *depth, float *image, float *ambient_sum, void *stream);
int b2n_composite_rays_uncertainty(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
        float *rays_t, const float *sigmas, const float *rgbs, const float *deltas, const float *ambients,
        const float *uncertainties, float *weights_sum, float *depth, float *image, float *ambient_sum,
        float *uncertainty_sum, void *stream);
int b2n_composite_rays_triplane(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
        float *rays_t, const float *sigmas, const float *rgbs, const float *deltas, const float *ambs_aud,
        const float *ambs_eye, const float *uncertainties, float *weights_sum, float *depth, float *image,
        float *amb_aud_sum, float *amb_eye_sum, float *uncertainty_sum, void *stream);

/* ---- grid encoder (gridencoder.h:12-13) ----------------------------------------------------------- */
/* inputs [B,D] fp32 in [0,1]; embeddings [sO,C] of `dtype`; offsets [L+1] int32; outputs [L,B,C] of `dtype`;
 * dy_dx [B,L*D*C] of `dtype` or NULL; gridtype 0 = hash, 1 = tiled.  D in 1..5, C in {1,2,4,8}. */
int b2n_grid_encode_forward(const float *inputs, const void *embeddings, const int32_t *offsets, void *outputs,
        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void *dy_dx,
        uint32_t gridtype, int align_corners, b2n_dtype dtype, void *stream);
/* the same values written ROW-MAJOR, outputs [B, L*C] — the layout GridEncoder.forward returns (grid.py:52 gets it from the [L,B,C] result with a transposing
 * copy): one thread evaluates all levels of a sample, rows leave through a shared-memory tile as contiguous lines.  No dy_dx.  D in 1..3, C in {1,2,4,8}. */
int b2n_grid_encode_forward_rows(const float *inputs, const void *embeddings, const int32_t *offsets, void *outputs,
        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
        b2n_dtype dtype, void *stream);
/* grad [L,B,C]; grad_embeddings [sO,C] pre-zeroed, accumulated with atomics; grad_inputs [B,D] or NULL */
int b2n_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings,
        const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
        uint32_t H, const void *dy_dx, void *grad_inputs, uint32_t gridtype, int align_corners,
        b2n_dtype dtype, void *stream);

/* scales_out[l] = exp2f(l*S)*H - 1 for l < L, computed with the kernels' own arithmetic (ex2.approx + fma, as the reference's
 * gridencoder.cu:125 compiles to): a diagnostic that lets a CPU checker reproduce fp32 encodings bit for bit */
int b2n_grid_level_scales(float S, uint32_t H, uint32_t L, float *scales_out, void *stream);

/* ---- spherical harmonics (shencoder.h:9-10) ------------------------------------------------------- */
/* inputs [B,3]; outputs [B,degree^2]; dy_dx [B,3*degree^2] or NULL; degree in 1..8 */
int b2n_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t degree,
                          float *dy_dx, void *stream);
int b2n_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t degree,
                           const float *dy_dx, float *grad_inputs, void *stream);

/* ---- frequency encoder (freqencoder.h:7,10) ------------------------------------------------------- */
/* outputs [B,C], C = D + 2*D*deg: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] */
int b2n_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                            float *outputs, void *stream);
int b2n_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg,
                             uint32_t C, float *grad_inputs, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2NERF_H_ */
