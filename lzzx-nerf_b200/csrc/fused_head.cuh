// fused_head.cuh — shared constants / argument block of the fused head kernel (fused_head.cu) and its callers.
#pragma once
#include "common.cuh"
#include "../../include/b2nerf_fused.h"

namespace b2n {

constexpr uint32_t HG_TILE = 128;                  // samples per tile = UMMA M
constexpr uint32_t HG_WGS = 3;                     // warpgroups (tiles in flight) per CTA; each owns two X tiles (double-buffered gather) and one H tile
constexpr uint32_t HG_TMEM_COLS = 160;             // TMEM columns per warpgroup (3 x 160 of the 512 allocated)
// accumulator columns inside a warpgroup's TMEM slice
constexpr uint32_t TC_A = 0;                       // 64: aud hidden (P1) | att 32 + unc hidden 32 (P2) | sigma layer 1 (P4) | geo 64 + logit (P5, 80 wide) | rgb (P7)
constexpr uint32_t TC_EYE = 64;                    // 16: eye hidden (P1)
constexpr uint32_t TC_S = 80;                      // 64: sigma hidden, accumulated by P1 and P3 | color hidden (P6)
constexpr uint32_t HG_THREADS = HG_WGS * 128;
constexpr uint32_t HG_TILE_BYTES = HG_TILE * 128;  // one SWIZZLE_128B K-atom of an A operand: 128 rows x 128 B

// fp16 weight image, every region a SWIZZLE_128B K-major atom [rows x 128 B]; offsets are multiples of 1024
constexpr uint32_t HW_A = 0;                        // 144 rows: aud_att0 (64) | eye_att0 (16) | sigma0[:, :36] (64)  — one N = 144 MMA
constexpr uint32_t HW_U = HW_A + 144 * 128;         //  32 rows: unc0
constexpr uint32_t HW_B = HW_U + 32 * 128;          //  32 rows: aud_att1
constexpr uint32_t HW_C = HW_B + 32 * 128;          //  64 rows: sigma0[:, 36:69]  (enc_w 32, eye 1)
constexpr uint32_t HW_D = HW_C + 64 * 128;          //  64 rows: sigma1
constexpr uint32_t HW_E = HW_D + 64 * 128;          //  80 rows: sigma2, rows rotated: 0..63 geo_feat, 64 density logit
constexpr uint32_t HW_F0 = HW_E + 80 * 128;         //  64 rows: color0[:, 16:80] (geo_feat)
constexpr uint32_t HW_F1 = HW_F0 + 64 * 128;        //  64 rows: color0[:, 0:16]  (SH)
constexpr uint32_t HW_G = HW_F1 + 64 * 128;         //  16 rows: color1 (3 valid)
constexpr uint32_t HW_BYTES = HW_G + 16 * 128;      // 71 680 B
static_assert(HW_U % 1024 == 0 && HW_B % 1024 == 0 && HW_C % 1024 == 0 && HW_D % 1024 == 0 && HW_E % 1024 == 0 && HW_F0 % 1024 == 0 && HW_F1 % 1024 == 0 && HW_G % 1024 == 0, "atoms must be 1024 B aligned");

// small fp32 vectors (values pre-rounded to fp16)
constexpr uint32_t HS_EYE_W1 = 0, HS_UNC_W1 = 16, HS_IND_W = 48, HS_C1W = 48 + 256, HS_FLOATS = 48 + 256 + 192;     // HS_C1W: color_net.net.1 [3][64] (backward)

// TRANSPOSED fp16 weight image for the backward-data kernel (fused_head_bwd.cu): dX = dY W needs B = W^T as [N = fan-in rows] x [K = fan-out]
// K-major SWIZZLE_128B atoms
constexpr uint32_t HT_C0G = 0;                      //  64 rows: color0[:, 16:80]^T   (d geo_feat    <- d color hidden)
constexpr uint32_t HT_C0I = HT_C0G + 64 * 128;      //  16 rows: color0[:, 80:84]^T   (d ind_code part, 4 valid)
constexpr uint32_t HT_S2A = HT_C0I + 16 * 128;      //  64 rows: sigma2[1:65, :]^T    (K = geo_feat index)
constexpr uint32_t HT_S2B = HT_S2A + 64 * 128;      //  64 rows: sigma2[0:1, :]^T     (K = density logit, 1 valid column)
constexpr uint32_t HT_S1 = HT_S2B + 64 * 128;       //  64 rows: sigma1^T
constexpr uint32_t HT_S0X = HT_S1 + 64 * 128;       //  48 rows: sigma0[:, 0:36]^T    (d enc_x, 36 valid)
constexpr uint32_t HT_S0W = HT_S0X + 48 * 128;      //  48 rows: sigma0[:, 36:69]^T   (d [enc_w, e], 33 valid)
constexpr uint32_t HT_A1 = HT_S0W + 48 * 128;       //  64 rows: aud_att1^T           (K = 32)
constexpr uint32_t HT_A0 = HT_A1 + 64 * 128;        //  48 rows: aud_att0^T           (36 valid)
constexpr uint32_t HT_E0 = HT_A0 + 48 * 128;        //  48 rows: eye_att0^T           (36 valid, K = 16)
constexpr uint32_t HT_BYTES = HT_E0 + 48 * 128;     // 67 584 B
static_assert(HT_C0I % 1024 == 0 && HT_S2A % 1024 == 0 && HT_S2B % 1024 == 0 && HT_S1 % 1024 == 0 && HT_S0X % 1024 == 0 && HT_S0W % 1024 == 0 &&
              HT_A1 % 1024 == 0 && HT_A0 % 1024 == 0 && HT_E0 % 1024 == 0, "atoms must be 1024 B aligned");

// per-level constants of the (shared) tri-plane geometry, precomputed at b2n_model_update
struct HeadLvl {
    float scale;        // exp2f(level*S)*H - 1, computed ON THE DEVICE (ex2.approx) so it carries the reference's bits
    uint32_t mul;       // dense level: row stride (resolution + 1); hashed level: the hash prime 2654435761
    uint32_t mask;      // hashed level: size - 1 (size is a power of two); dense level: 0xffffffff
    uint32_t off;       // first table entry of the level
};

struct HeadArgs {
    const float *xyzs, *dirs;
    uint32_t M;
    const int32_t *n_valid;
    const float *live_deltas;   // optional deltas[M,2]: rows with deltas[m,0] == 0 are unproduced march slots; tiles made only of such rows are skipped
    const float *tab[3];
    HeadLvl lvl[12];
    float bound;
    float inv_two_bound;        // 1 / (2 bound) when 2 bound is a power of two (exact), else 0 -> divide
    const uint8_t *wimg;
    const float *wsmall;
    const float *enc_a, *ind_code, *eye;
    float *sigmas, *rgbs, *amb_aud, *amb_eye, *unc;
    int has_unc;
    b2n_head_saved sv;          // training forward: where the activations go (used by the SAVE instantiation only)
    float density_scale;
    uint32_t max_ctas;          // 0 = one CTA per SM
    int sched;                  // tile -> CTA mapping, A/B switch (B2N_HEAD_SCHED): 0 balanced contiguous shares, short launches packed; 1 spread; 2 grid-strided
};

}  // namespace b2n

// the packed model behind the opaque C handle
struct b2n_model {
    uint8_t *wimg = nullptr;       // HW_BYTES   forward operand image
    uint8_t *wimg_t = nullptr;     // HT_BYTES   transposed image (backward-data)
    float *wsmall = nullptr;       // HS_FLOATS (+ 16 floats of scratch for the device-computed level scales)
    b2n_head_weights w = {};
    b2n::HeadLvl lvl[12] = {};
    // corner-quad image of the tables (fused_head.cu:k_pack_quads): [3][quad_cells] float4, qlvl[l] = {scale, res, -, first cell of the level}
    float4 *quads = nullptr;
    uint32_t quad_cells = 0;
    b2n::HeadLvl qlvl[12] = {};
    bool use_quads = true;
    // geometry the cached lvl[] was derived from (re-derived only when it changes: one small D2H read + sync)
    const int32_t *geo_offsets = nullptr;
    float geo_S = 0.0f;
    uint32_t geo_H = 0;
    bool ready = false;
};

namespace b2n {

size_t head_smem_bytes();
int launch_head_forward(const HeadArgs &a, cudaStream_t st, bool save = false, bool quad = false);

}  // namespace b2n
