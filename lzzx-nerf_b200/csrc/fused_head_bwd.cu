// fused_head_bwd.cu — backward-DATA pass of the per-sample head network as one persistent tcgen05 kernel (training).
//
// Autograd runs ~25 GEMMs and ~150 elementwise kernels for the backward of nerf_triplane/network.py:252-311 (LinearBackward, ReluBackward,
// CatBackward, the autocast casts ...).  Here every per-sample gradient of the graph is produced in ONE pass over the samples:
//   inputs : the gradients of the five network outputs (sigma, rgb, ambient_aud = ||att||, ambient_eye = eye_att, uncertainty) and the
//            fp16 activations kept by the training forward (b2n_head_saved);
//   outputs: the gradient of every pre-activation (fp16, the "dY" operands of the weight-gradient kernel b2n_linear_wgrad, whose "X" operands
//            are the saved activations), the gradient of the tri-plane features written straight in the grid backward's [plane][level][sample]
//            layout, and the per-sample gradients of enc_a * att and of the ind-code inputs (summed by the caller).
// dX = dY W is a tensor-core layer like the forward's: dY tile (thread = sample row, fp16, K-major SWIZZLE_128B) x W^T (the transposed weight
// image packed by b2n_model_update) -> TMEM; the epilogue applies the ReLU mask read from the saved activation and writes the next dY tile
// in place.  Six MMA rounds per 128-sample tile:
//   B1 (CUDA cores) d rgb logits -> d color hidden                        B2  d geo_feat | d ind-code part = d hc . color0
//   B3  d sigma hidden 2 = d [geo, logit] . sigma2                         B4  d sigma hidden 1 = . sigma1
//   B5  d enc_x (kept in TMEM) | d [enc_w, e] = d h1 . sigma0              -> d att, d eye logit, d eye hidden (CUDA cores)
//   B6  d aud hidden = d att . aud_att1 ;  d enc_x += d eye hidden . eye_att0
//   B7  d enc_x += d aud hidden . aud_att0                                 -> grid-gradient planes
// Numerics follow autograd under autocast(fp16): fp16 gradients between layers, fp32 accumulation inside a layer (TMEM); the three
// contributions to d enc_x are summed in fp32 (autograd sums them in fp16).
#include "common.cuh"
#include "tc5.cuh"
#include "fused_head.cuh"

namespace b2n {
using namespace tc5;

constexpr uint32_t BW_WGS = 3;
constexpr uint32_t BW_THREADS = BW_WGS * 128;
constexpr uint32_t BW_TMEM_COLS = 160;        // per warpgroup: scratch accumulators at 0..95, d enc_x at 96..143
constexpr uint32_t TB_S = 0, TB_X = 96;

struct BwdArgs {
    uint32_t M;
    const uint8_t *wimg_t;
    const float *wsmall, *enc_a, *eye;
    b2n_head_saved sv;
    b2n_head_grads g;
    const float *sigmas, *amb_aud, *g_sigma, *g_rgb, *g_aud, *g_eye, *g_unc;
    int has_unc;
};

struct BwdSmem {
    float wc1[192];                 // color_net.net.1 [3][64]
    float eye_w1[16], unc_w1[32], enc_a_h[32];
    float eye_val;
    uint32_t tmem_base;
    uint64_t bar_w;
    uint64_t bar_mma[BW_WGS];
};

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float lo_h(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float hi_h(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
// ReLU backward on a pair: the saved activation is the ReLU output (>= 0), so "was active" == "is not zero"
__device__ __forceinline__ uint32_t pk2_masked(float lo, float hi, uint32_t act) {
    return pk2((act & 0x7fffu) ? lo : 0.0f, (act & 0x7fff0000u) ? hi : 0.0f);
}
__device__ __forceinline__ void issue_mma_b(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, uint32_t ksteps, uint32_t N, bool accumulate) {
    const uint32_t idesc = idesc_f16(128, N);
    uint64_t da = smem_desc_sw128(a_saddr), db = smem_desc_sw128(b_saddr);
#pragma unroll 1
    for (uint32_t k = 0; k < ksteps; k++, da += 2, db += 2) mma_f16_ss(d_tmem, da, db, idesc, accumulate || k > 0);
}
// 64 accumulator columns -> masked fp16 row: written into the operand tile (next layer's dY) and to global memory (wgrad operand)
__device__ __forceinline__ void masked_epilogue(uint32_t taddr, uint8_t *tile, uint32_t row, const uint4 (&act)[8], uint4 *save) {
#pragma unroll
    for (uint32_t cb = 0; cb < 64; cb += 32) {
        uint32_t acc[32];
        ld32(taddr + cb, acc);
        wait_ld();
        uint4 q[4];
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            const uint4 a4 = act[(cb >> 3) + c];
            const uint32_t am[4] = {a4.x, a4.y, a4.z, a4.w};
            uint32_t w[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) w[j] = pk2_masked(__uint_as_float(acc[c * 8 + 2 * j]), __uint_as_float(acc[c * 8 + 2 * j + 1]), am[j]);
            q[c] = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(tile + sw128_offset(row, (cb >> 3) + c)) = q[c];
        }
        if (save) { st256(save + (cb >> 3), q[0], q[1]); st256(save + (cb >> 3) + 2, q[2], q[3]); }      // [M,64] fp16 rows are 128-byte aligned
    }
}
// this warp's 32 rows (128 B each) of a SWIZZLE_128B tile -> row-major global array, 8 lanes per row (see fused_head.cu:warp_rows_out)
__device__ __forceinline__ void warp_rows_out(const uint8_t *tile, uint32_t warp_row0, uint8_t *gtile, uint32_t pitch, uint32_t rows_valid) {
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) {
        const uint32_t p = i * 32u + lane, r = warp_row0 + (p >> 3), c = p & 7u;
        const uint4 v = *reinterpret_cast<const uint4 *>(tile + sw128_offset(r, c));
        if (r < rows_valid) __stcs(reinterpret_cast<uint4 *>(gtile + (size_t)r * pitch + c * 16u), v);
    }
}
__device__ __forceinline__ void load_row8(uint4 (&r)[8], const void *base, size_t m, bool live) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base) + m * 8;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        if (live) ld256(p + i, r[i], r[i + 1]);
        else { r[i] = make_uint4(0, 0, 0, 0); r[i + 1] = make_uint4(0, 0, 0, 0); }
    }
}

__global__ void __launch_bounds__(BW_THREADS, 1) k_head_backward(const __grid_constant__ BwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_w = base;
    uint8_t *s_tiles = base + HT_BYTES;
    BwdSmem &S = *reinterpret_cast<BwdSmem *>(s_tiles + BW_WGS * 2 * HG_TILE_BYTES);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u, warp = tid >> 5;
    uint8_t *sP = s_tiles + wg * 2 * HG_TILE_BYTES, *sQ = sP + HG_TILE_BYTES;

    if (tid == 0) {
        mbar_init(&S.bar_w, 1);
        for (int g = 0; g < (int)BW_WGS; g++) mbar_init(&S.bar_mma[g], 1);
        fence_mbar_init();
        mbar_expect_tx(&S.bar_w, HT_BYTES);
        bulk_g2s(s_w, a.wimg_t, HT_BYTES, &S.bar_w);
        S.eye_val = a.eye ? a.eye[0] : 0.0f;
    }
    if (warp == 1) tmem_alloc(&S.tmem_base, 512);
    if (tid >= 128 && tid < 320) S.wc1[tid - 128] = a.wsmall[HS_C1W + tid - 128];
    if (tid >= 320 && tid < 336) S.eye_w1[tid - 320] = a.wsmall[HS_EYE_W1 + tid - 320];
    if (tid >= 336 && tid < 368) S.unc_w1[tid - 336] = a.wsmall[HS_UNC_W1 + tid - 336];
    if (tid < 32) S.enc_a_h[tid] = rh(a.enc_a[tid]);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    mbar_wait(&S.bar_w, 0);

    const uint32_t n_tiles = (a.M + HG_TILE - 1) / HG_TILE;
    const uint32_t tmem_wg = S.tmem_base + wg * BW_TMEM_COLS;
    const uint32_t tmem_ld = tmem_wg + (((warp & 3u) * 32u) << 16);
    const uint32_t sP_a = smem_u32(sP), sQ_a = smem_u32(sQ), sW_a = smem_u32(s_w);
    uint64_t *bar = &S.bar_mma[wg];
    uint32_t phase = 0;
    auto sync_wg = [&]() { bar_sync(1 + wg, 128); };
    auto mma_done = [&]() { mbar_wait(bar, phase); phase ^= 1u; fence_after_sync(); };
    auto publish = [&]() { fence_before_sync(); fence_proxy_async(); sync_wg(); };

    for (uint32_t tile = blockIdx.x * BW_WGS + wg; tile < n_tiles; tile += gridDim.x * BW_WGS) {
        const size_t m = (size_t)tile * HG_TILE + t;
        const bool live = m < a.M;
        const size_t tile_row0 = (size_t)tile * HG_TILE;
        const uint32_t rows_valid = (uint32_t)min((size_t)HG_TILE, (size_t)a.M - tile_row0), wrow0 = (warp & 3u) * 32u;
        // ---- per-sample scalars -------------------------------------------------------------------------------------------------
        float g_sig = 0, sig = 0, g_r[3] = {0, 0, 0}, g_aud = 0, nrm = 0, g_eye = 0, g_unc = 0;
        uint4 misc = make_uint4(0, 0, 0, 0);
        if (live) {
            if (a.g_sigma) g_sig = __ldcs(a.g_sigma + m);
            sig = __ldcs(a.sigmas + m);
            if (a.g_rgb) { g_r[0] = __ldcs(a.g_rgb + 3 * m); g_r[1] = __ldcs(a.g_rgb + 3 * m + 1); g_r[2] = __ldcs(a.g_rgb + 3 * m + 2); }
            if (a.g_aud) g_aud = __ldcs(a.g_aud + m);
            nrm = __ldcs(a.amb_aud + m);
            if (a.g_eye) g_eye = __ldcs(a.g_eye + m);
            if (a.g_unc) g_unc = __ldcs(a.g_unc + m);
            misc = __ldcs(reinterpret_cast<const uint4 *>(a.sv.misc) + m);
        }
        uint4 act[8];
        load_row8(act, a.sv.hc, m, live);
        // ---- B1: rgb = sigmoid(logit) * 1.002 - 0.001 (network.py:275); d hc = d logit . color1, masked --------------------------------
        {
            const float s3[3] = {lo_h(misc.x), hi_h(misc.x), lo_h(misc.y)};
            float dl[3];
#pragma unroll
            for (int c = 0; c < 3; c++) dl[c] = rh(rh(g_r[c] * 1.002f) * rh(s3[c] * (1.0f - s3[c])));
            if (live) reinterpret_cast<uint4 *>(a.g.d_rl)[m] = make_uint4(pk2(dl[0], dl[1]), pk2(dl[2], 0.0f), 0u, 0u);
#pragma unroll
            for (uint32_t c = 0; c < 8; c++) {
                const uint32_t am[4] = {act[c].x, act[c].y, act[c].z, act[c].w};
                uint32_t w[4];
#pragma unroll
                for (uint32_t j = 0; j < 4; j++) {
                    const uint32_t k = c * 8 + 2 * j;
                    const float v0 = fmaf(dl[2], S.wc1[128 + k], fmaf(dl[1], S.wc1[64 + k], dl[0] * S.wc1[k]));
                    const float v1 = fmaf(dl[2], S.wc1[128 + k + 1], fmaf(dl[1], S.wc1[64 + k + 1], dl[0] * S.wc1[k + 1]));
                    w[j] = pk2_masked(v0, v1, am[j]);
                }
                const uint4 q = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4 *>(sP + sw128_offset(t, c)) = q;
            }
            warp_rows_out(sP, wrow0, reinterpret_cast<uint8_t *>(a.g.d_hc) + tile_row0 * 128, 128, rows_valid);
        }
        publish();
        // ---- B2: d geo_feat (64) | d ind-code inputs (4) = d hc . color0[:, 16:84] ------------------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma_b(tmem_wg + TB_S, sP_a, sW_a + HT_C0G, 4, 64, false);
            issue_mma_b(tmem_wg + TB_S + 64, sP_a, sW_a + HT_C0I, 4, 16, false);
            mma_commit(bar);
        }
        load_row8(act, a.sv.h2, m, live);                                  // ReLU mask of B3, in flight under the MMA
        mma_done();
        {
            // d o = [d geo (64) | d logit | 0 ...]: K = 80 operand = atom sP (geo) + atom sQ (logit in column 0); global copy in the same order
            uint4 *save = live ? reinterpret_cast<uint4 *>(a.g.d_o) + m * 9 : nullptr;
#pragma unroll
            for (uint32_t cb = 0; cb < 64; cb += 32) {
                uint32_t acc[32];
                ld32(tmem_ld + TB_S + cb, acc);
                wait_ld();
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    uint32_t w[4];
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++) w[j] = pk2(__uint_as_float(acc[c * 8 + 2 * j]), __uint_as_float(acc[c * 8 + 2 * j + 1]));
                    const uint4 q = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4 *>(sP + sw128_offset(t, (cb >> 3) + c)) = q;
                }
            }
            warp_rows_out(sP, wrow0, reinterpret_cast<uint8_t *>(a.g.d_o) + tile_row0 * 144, 144, rows_valid);
            uint32_t i16[16];
            ld16(tmem_ld + TB_S + 64, i16);
            wait_ld();
            const float d_logit = rh(g_sig * sig);                           // sigma = exp(h0) in fp32 (network.py:301)
            const uint4 q0 = make_uint4(pk2(d_logit, 0.0f), 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(sQ + sw128_offset(t, 0)) = q0;
            *reinterpret_cast<uint4 *>(sQ + sw128_offset(t, 1)) = make_uint4(0u, 0u, 0u, 0u);
            if (live) {
                save[8] = q0;
                reinterpret_cast<uint4 *>(a.g.d_ci)[m] = make_uint4(pk2(__uint_as_float(i16[0]), __uint_as_float(i16[1])), pk2(__uint_as_float(i16[2]), __uint_as_float(i16[3])), 0u, 0u);
            }
        }
        publish();
        // ---- B3: d h2 = d o . sigma2, masked ------------------------------------------------------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma_b(tmem_wg + TB_S, sP_a, sW_a + HT_S2A, 4, 64, false);
            issue_mma_b(tmem_wg + TB_S, sQ_a, sW_a + HT_S2B, 1, 64, true);
            mma_commit(bar);
        }
        mma_done();
        masked_epilogue(tmem_ld + TB_S, sP, t, act, nullptr);
        warp_rows_out(sP, wrow0, reinterpret_cast<uint8_t *>(a.g.d_h2) + tile_row0 * 128, 128, rows_valid);
        load_row8(act, a.sv.h1, m, live);
        publish();
        // ---- B4: d h1 = d h2 . sigma1, masked -----------------------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma_b(tmem_wg + TB_S, sP_a, sW_a + HT_S1, 4, 64, false); mma_commit(bar); }
        mma_done();
        masked_epilogue(tmem_ld + TB_S, sP, t, act, nullptr);
        warp_rows_out(sP, wrow0, reinterpret_cast<uint8_t *>(a.g.d_h1) + tile_row0 * 128, 128, rows_valid);
        publish();
        // ---- B5: d enc_x (TMEM, kept) | d [enc_w (32), e] = d h1 . sigma0 ------------------------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma_b(tmem_wg + TB_X, sP_a, sW_a + HT_S0X, 4, 48, false);
            issue_mma_b(tmem_wg + TB_S, sP_a, sW_a + HT_S0W, 4, 48, false);
            mma_commit(bar);
        }
        uint4 attq[4], heq[2];
        {
            const uint4 *pa = reinterpret_cast<const uint4 *>(a.sv.att) + m * 4;
            const uint4 *ph = reinterpret_cast<const uint4 *>(a.sv.he) + m * 2;
#pragma unroll
            for (int i = 0; i < 4; i++) attq[i] = live ? __ldcs(pa + i) : make_uint4(0, 0, 0, 0);
            heq[0] = live ? __ldcs(ph) : make_uint4(0, 0, 0, 0); heq[1] = live ? __ldcs(ph + 1) : make_uint4(0, 0, 0, 0);
        }
        mma_done();
        {
            uint32_t acc[32];
            ld32(tmem_ld + TB_S, acc);
            wait_ld();
            uint32_t e16[16];
            ld16(tmem_ld + TB_S + 32, e16);
            wait_ld();
            const float inv_n = nrm > 0.0f ? 1.0f / nrm : 0.0f;
            const uint32_t aw[16] = {attq[0].x, attq[0].y, attq[0].z, attq[0].w, attq[1].x, attq[1].y, attq[1].z, attq[1].w,
                                     attq[2].x, attq[2].y, attq[2].z, attq[2].w, attq[3].x, attq[3].y, attq[3].z, attq[3].w};
            uint32_t wew[16], wat[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const float d0 = rh(__uint_as_float(acc[2 * j])), d1 = rh(__uint_as_float(acc[2 * j + 1]));      // d enc_w (fp16)
                const float a0 = lo_h(aw[j]), a1 = hi_h(aw[j]);
                wew[j] = pk2(d0, d1);
                // enc_w = enc_a * att (network.py:285); ambient_aud = ||att|| in fp32 (network.py:308)
                wat[j] = pk2(rh(d0 * S.enc_a_h[2 * j]) + rh(g_aud * a0 * inv_n), rh(d1 * S.enc_a_h[2 * j + 1]) + rh(g_aud * a1 * inv_n));
            }
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                const uint4 q = make_uint4(wat[4 * c], wat[4 * c + 1], wat[4 * c + 2], wat[4 * c + 3]);
                *reinterpret_cast<uint4 *>(sP + sw128_offset(t, c)) = q;                                          // d att: K = 32 operand of B6
                if (live) {
                    reinterpret_cast<uint4 *>(a.g.d_att)[m * 4 + c] = q;
                    reinterpret_cast<uint4 *>(a.g.d_ew)[m * 4 + c] = make_uint4(wew[4 * c], wew[4 * c + 1], wew[4 * c + 2], wew[4 * c + 3]);
                }
            }
            // e = eye * eye_att, eye_att = sigmoid(logit) on half tensors (network.py:288-291); ambient_eye = eye_att
            const float ea = hi_h(misc.y);
            const float d_e = rh(__uint_as_float(e16[0]));
            const float d_ea = rh(d_e * S.eye_val) + rh(g_eye);
            const float d_el = rh(d_ea * rh(ea * (1.0f - ea)));
            const uint32_t hw[8] = {heq[0].x, heq[0].y, heq[0].z, heq[0].w, heq[1].x, heq[1].y, heq[1].z, heq[1].w};
            uint32_t whe[8];
#pragma unroll
            for (int j = 0; j < 8; j++) whe[j] = pk2_masked(d_el * S.eye_w1[2 * j], d_el * S.eye_w1[2 * j + 1], hw[j]);
            const uint4 h0 = make_uint4(whe[0], whe[1], whe[2], whe[3]), h1q = make_uint4(whe[4], whe[5], whe[6], whe[7]);
            *reinterpret_cast<uint4 *>(sQ + sw128_offset(t, 0)) = h0;                                              // d eye hidden: K = 16 operand
            *reinterpret_cast<uint4 *>(sQ + sw128_offset(t, 1)) = h1q;
            if (live) {
                reinterpret_cast<uint4 *>(a.g.d_he)[m * 2] = h0; reinterpret_cast<uint4 *>(a.g.d_he)[m * 2 + 1] = h1q;
                reinterpret_cast<uint4 *>(a.g.d_el)[m] = make_uint4(pk2(d_el, d_e), 0u, 0u, 0u);
            }
            // unc_net on enc_x.detach(): uncertainty = log(1 + exp(u)) (network.py:276-278) -> d u = g * sigmoid(u); d hidden = d u * w1, masked
            if (a.has_unc && live) {
                const float ul = lo_h(misc.z);
                const float d_ul = rh(g_unc * (1.0f / (1.0f + expf(-ul))));
                reinterpret_cast<uint4 *>(a.g.d_ul)[m] = make_uint4(pk2(d_ul, 0.0f), 0u, 0u, 0u);
                const uint4 *pu = reinterpret_cast<const uint4 *>(a.sv.hu) + m * 4;
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    const uint4 hq = __ldcs(pu + c);
                    const uint32_t hm[4] = {hq.x, hq.y, hq.z, hq.w};
                    uint32_t w[4];
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++) w[j] = pk2_masked(d_ul * S.unc_w1[c * 8 + 2 * j], d_ul * S.unc_w1[c * 8 + 2 * j + 1], hm[j]);
                    reinterpret_cast<uint4 *>(a.g.d_hu)[m * 4 + c] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        publish();
        // ---- B6: d ha = d att . aud_att1 (masked);  d enc_x += d he . eye_att0 -----------------------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma_b(tmem_wg + TB_S, sP_a, sW_a + HT_A1, 2, 64, false);
            issue_mma_b(tmem_wg + TB_X, sQ_a, sW_a + HT_E0, 1, 48, true);
            mma_commit(bar);
        }
        load_row8(act, a.sv.ha, m, live);
        mma_done();
        masked_epilogue(tmem_ld + TB_S, sP, t, act, nullptr);
        warp_rows_out(sP, wrow0, reinterpret_cast<uint8_t *>(a.g.d_ha) + tile_row0 * 128, 128, rows_valid);
        publish();
        // ---- B7: d enc_x += d ha . aud_att0 -> the grid backward's [plane][level][sample] planes ---------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma_b(tmem_wg + TB_X, sP_a, sW_a + HT_A0, 4, 48, true); mma_commit(bar); }
        mma_done();
        {
            uint32_t x32[32], x16[16];
            ld32(tmem_ld + TB_X, x32);
            ld16(tmem_ld + TB_X + 32, x16);
            wait_ld();
            if (live) {
#pragma unroll
                for (uint32_t f = 0; f < 32; f++) __stcs(a.g.d_planes + (size_t)f * a.M + m, __uint_as_float(x32[f]));       // feature f = plane * 12 + level
#pragma unroll
                for (uint32_t f = 0; f < 4; f++) __stcs(a.g.d_planes + (size_t)(32 + f) * a.M + m, __uint_as_float(x16[f]));
            }
        }
        // the next tile's first publish() orders this tile's TMEM reads before its first MMA
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(S.tmem_base, 512);
}

}  // namespace b2n

using namespace b2n;

extern "C" int b2n_head_backward(const b2n_model *m, uint32_t M, const float *enc_a, const float *eye, const b2n_head_saved *saved, const float *sigmas,
                                 const float *amb_aud, const float *g_sigma, const float *g_rgb, const float *g_aud, const float *g_eye, const float *g_unc,
                                 const b2n_head_grads *grads, void *stream) {
    B2N_REQUIRE(m && m->ready, "head_backward: model has no weights (call b2n_model_update)");
    B2N_REQUIRE(enc_a && saved && sigmas && amb_aud && grads, "head_backward: null pointer");
    B2N_REQUIRE(saved->ha && saved->he && saved->att && saved->h1 && saved->h2 && saved->hc && saved->misc, "head_backward: null activation buffer");
    B2N_REQUIRE(grads->d_rl && grads->d_hc && grads->d_o && grads->d_h2 && grads->d_h1 && grads->d_ew && grads->d_att && grads->d_ha && grads->d_el &&
                grads->d_he && grads->d_ci && grads->d_planes, "head_backward: null gradient buffer");
    const bool has_unc = m->w.unc_w0 != nullptr;
    B2N_REQUIRE(!has_unc || (saved->hu && grads->d_ul && grads->d_hu), "head_backward: unc_net is packed but its buffers are NULL");
    if (M == 0) return 0;
    BwdArgs a = {};
    a.M = M; a.wimg_t = m->wimg_t; a.wsmall = m->wsmall; a.enc_a = enc_a; a.eye = eye; a.sv = *saved; a.g = *grads;
    a.sigmas = sigmas; a.amb_aud = amb_aud; a.g_sigma = g_sigma; a.g_rgb = g_rgb; a.g_aud = g_aud; a.g_eye = g_eye; a.g_unc = g_unc;
    a.has_unc = has_unc;
    const size_t smem = 1024 + HT_BYTES + (size_t)BW_WGS * 2 * HG_TILE_BYTES + sizeof(BwdSmem);
    B2N_SMEM(k_head_backward, smem);
    uint32_t ctas = ceil_div<uint32_t>(ceil_div<uint32_t>(M, HG_TILE), BW_WGS);
    const uint32_t sms = (uint32_t)sm_count();
    if (ctas > sms) ctas = sms;
    k_head_backward<<<ctas, BW_THREADS, smem, as_stream(stream)>>>(a);
    return check_launch("head_backward");
}
