// fused_head.cu — the per-sample head network of the talking-head NeRF as ONE persistent tcgen05 kernel.
//
// Replaces the ~45-launch chain of nerf_triplane/network.py:252-311 (NeRFNetwork.forward / density):
//   tri-plane encode (3 x GridEncoder D=2 L=12 C=1, network.py:215-223)  -> 36 features
//   aud_ch_att_net 36->64->32, eye_att_net 36->16->1, [unc_net 36->32->1]            (network.py:141-152, 283-292)
//   sigma_net [36 + 32 + 1 = 69] -> 64 -> 64 -> 65, sigma = exp(h0)                   (network.py:298-301)
//   SH degree 4 of the view direction, color_net [16 + 64 + 4 = 84] -> 64 -> 3         (network.py:267-275)
// Numerics follow the reference under autocast(fp16): GEMM operands and activations are fp16, accumulation fp32
// (TMEM), exp / norm / softplus in fp32, sigmoid in fp16.
//
// Organisation (B200) — two kernels share the gather, the weight image and the layer arithmetic:
//   * k_head_forward<SAVE, QUAD> (training forward with kept activations; inference when unc_net is evaluated): persistent CTA per SM, 384 threads =
//     3 warpgroups; a warpgroup owns one 128-sample tile at a time, 160 TMEM columns and three 16 KB SWIZZLE_128B operand tiles (two feature tiles,
//     double-buffered, and one hidden-activation tile); epilogues read TMEM with tcgen05.ld (thread == sample row), apply the activation and write
//     the next layer's fp16 operand to shared memory.
//   * k_head_infer4<QUAD> (inference): 512 threads = 4 warpgroups, 128 TMEM columns each; the hidden activations never leave tensor memory — the
//     epilogue writes the next layer's A operand back with tcgen05.st and the layer is a TMEM-A tcgen05.mma (see the comment at the kernel).
//   * all weights (fp16, pre-swizzled K-major SWIZZLE_128B image, 70 KB) are brought in once per CTA by one TMA bulk copy and stay resident; every
//     layer is a tcgen05.mma (M=128 samples, N=16..144, K=16 per instruction) issued by one thread per warpgroup, accumulator in TMEM, completion
//     through tcgen05.commit -> mbarrier.
//   * the gather reads ONE 16-byte corner quad per (level, plane) cell from the quad image (36 LDG.128 per sample), six trips of the NEXT tile's
//     gather in flight underneath the CURRENT tile's MMA phases.
#include <stdlib.h>
#include "common.cuh"
#include "tc5.cuh"
#include "fused_head.cuh"

namespace b2n {
using namespace tc5;

// ---------------------------------------------------------------------------------------------------
// weight packing: fp32 nn.Linear weights [out,in] -> fp16 SWIZZLE_128B K-major image (+ small fp32 vectors)
// ---------------------------------------------------------------------------------------------------
struct PackRegion { const float *src; uint32_t byte_off, rows, src_rows, ld, col0, kvalid, row_shift, tr, k0, col1, k1_off, k1_valid; };
// (col1, k1_off, k1_valid): a second run of source columns placed at K positions [k1_off, k1_off + k1_valid) of the same rows (untransposed regions)
// tr = 1: transposed region, value(row n, k) = src[(k0 + k) * ld + col0 + n] for n < src_rows, k < kvalid
struct PackArgs { PackRegion r[12]; uint32_t n; };

__global__ void __launch_bounds__(256) k_pack_head(const __grid_constant__ PackArgs pa, uint8_t *__restrict__ img) {
    // one thread per 16-byte chunk (8 halves) of the image
    uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = 0; i < pa.n; i++) {
        const PackRegion &g = pa.r[i];
        const uint32_t nchunks = g.rows * 8u;
        if (chunk < nchunks) {
            const uint32_t row = chunk >> 3, c = chunk & 7u;
            // row_shift rotates source rows (used to move sigma_net's density logit from row 0 to row 64)
            const uint32_t srow = g.row_shift ? (row < g.src_rows ? (row + g.row_shift) % g.src_rows : row) : row;
            __half h[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const uint32_t kk = c * 8u + k;
                float v = 0.0f;
                if (g.src && row < g.src_rows && kk < g.kvalid) v = g.tr ? g.src[(size_t)(g.k0 + kk) * g.ld + g.col0 + row] : g.src[(size_t)srow * g.ld + g.col0 + kk];
                else if (g.src && !g.tr && row < g.src_rows && kk >= g.k1_off && kk < g.k1_off + g.k1_valid) v = g.src[(size_t)srow * g.ld + g.col1 + (kk - g.k1_off)];
                h[k] = __float2half_rn(v);
            }
            *reinterpret_cast<uint4 *>(img + g.byte_off + sw128_offset(row, c)) = *reinterpret_cast<uint4 *>(h);
            return;
        }
        chunk -= nchunks;
    }
}

// small vectors: eye_w1[16], unc_w1[32], color ind part [64][4] — stored as fp32 values already rounded to fp16
__global__ void k_pack_small(const float *__restrict__ eye_w1, const float *__restrict__ unc_w1, const float *__restrict__ color_w0,
                             const float *__restrict__ color_w1, float *__restrict__ out) {
    const uint32_t t = threadIdx.x;
    auto rh = [](float v) { return __half2float(__float2half_rn(v)); };
    if (t < 16) out[HS_EYE_W1 + t] = rh(eye_w1[t]);
    if (t < 32) out[HS_UNC_W1 + t] = unc_w1 ? rh(unc_w1[t]) : 0.0f;
    for (uint32_t i = t; i < 256; i += blockDim.x) out[HS_IND_W + i] = rh(color_w0[(i >> 2) * 84 + 80 + (i & 3)]);
    for (uint32_t i = t; i < 192; i += blockDim.x) out[HS_C1W + i] = rh(color_w1[i]);
}

// ---------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
// two fp32 -> packed fp16 with ReLU folded into the conversion (cvt.rn.relu: relu(round(x)) == round(relu(x))); lo goes to bits [0,16)
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// TMEM -> fp16 operand tile: reads 64 accumulator columns of this thread's row starting at TMEM address `taddr`, applies
// (+bias, ReLU) and writes them as halves 0..63 of row `row` of a SWIZZLE_128B tile.  Inlined (the table loads of a gather trip stay in
// flight across it) but rolled over 32-column blocks to keep the tile loop's code size down.  (Training copies the finished rows out with
// warp_rows_out.)
template <bool RELU, bool BIAS>
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr, uint8_t *tile, uint32_t row, const float *bias) {
#pragma unroll 1
    for (uint32_t cb = 0; cb < 64; cb += 32) {
        uint32_t acc[32];
        ld32(taddr + cb, acc);
        wait_ld();
        uint4 q[4];
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            uint32_t w[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                float x0 = __uint_as_float(acc[c * 8 + 2 * j]), x1 = __uint_as_float(acc[c * 8 + 2 * j + 1]);
                if (BIAS) { x0 += bias[cb + c * 8 + 2 * j]; x1 += bias[cb + c * 8 + 2 * j + 1]; }
                w[j] = RELU ? pack2_relu(x0, x1) : pack2(x0, x1);
            }
            q[c] = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(tile + sw128_offset(row, (cb >> 3) + c)) = q[c];
        }
    }
}

// Training: copy this warp's 32 rows (128 B each) of a SWIZZLE_128B tile to a row-major global array, warp-cooperatively: 8 lanes write the 8
// chunks of one row, so a store instruction covers 4 full rows (4 cache lines) — a thread storing its own row costs 32 lines per instruction, and
// with ~1 KB of activations per sample that is what bound the first version of this kernel (lg_throttle / mio_throttle).  The rows were written
// by this same warp, so a __syncwarp() orders them; the tensor pipe only READS the tile meanwhile.
template <uint32_t NCH>      // NCH 16-byte chunks per row, starting at chunk `chunk0` of the tile row
__device__ __forceinline__ void warp_chunks_out(const uint8_t *tile, uint32_t chunk0, uint32_t warp_row0, uint8_t *gtile, uint32_t pitch, uint32_t rows_valid) {
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (uint32_t i = 0; i < NCH; i++) {
        const uint32_t p = i * 32u + lane, r = warp_row0 + p / NCH, c = p % NCH;
        const uint4 v = *reinterpret_cast<const uint4 *>(tile + sw128_offset(r, chunk0 + c));
        if (r < rows_valid) __stcs(reinterpret_cast<uint4 *>(gtile + (size_t)r * pitch + c * 16u), v);
    }
}
__device__ __forceinline__ void warp_rows_out(const uint8_t *tile, uint32_t warp_row0, uint8_t *gtile, uint32_t pitch, uint32_t rows_valid) {
    warp_chunks_out<8>(tile, 0, warp_row0, gtile, pitch, rows_valid);
}

// ---- tri-plane gather, split in two halves so the table reads of one trip (2 levels x 3 planes x 4 corners = 24 loads per sample) stay
// in flight underneath an MMA completion wait and its epilogue:  gather_issue() computes the cells and issues the loads, gather_finish()
// blends and stores.  Arithmetic identical to gridenc.cu:k_grid_fwd<float,2,1>: position fma(u, scale, 0.5), weights
// (1-fx|fx)*(1-fy|fy), four fmas in corner order (0,0),(1,0),(0,1),(1,1); index i + j*stride on dense levels, (i ^ j*2654435761) & (size-1)
// on hashed levels (gridencoder.cu:54-72 with D = 2; the level kind is uniform over the grid).
struct SampleCoord { float ux, uy, uz; uint32_t ok; };      // normalised coordinates; ok bit p = plane p in range (and the row is live)
struct GatherTrip { float4 v[2][3]; float fx[2], fy[2], fz[2]; };      // v = corners (0,0),(1,0),(0,1),(1,1) of the cell

// QUAD = false: four 4-byte reads per (level, plane) straight from the reference-format tables.
// QUAD = true : ONE 16-byte read per (level, plane) from the model's corner-quad image (k_pack_quads below): entry (i, j) of a level holds the four
//               corner values of cell (i, j), hashed levels de-hashed, so the index is i + j * res on every level.  Same values, same blend —
//               the features are bit-identical; a sample costs 36 LDG.128 instead of 144 LDG.32 (+ their 64-bit address arithmetic and hashing).
template <bool QUAD>
__device__ __forceinline__ void gather_issue(GatherTrip &G, const float *__restrict__ t_xy, const float *__restrict__ t_yz, const float *__restrict__ t_xz,
                                             const HeadLvl *lv, const SampleCoord &c) {
#pragma unroll
    for (uint32_t q = 0; q < 2; q++) {
        const HeadLvl g = lv[q];                                   // warp-uniform shared-memory read
        const float qx = __fmaf_rn(c.ux, g.scale, 0.5f), qy = __fmaf_rn(c.uy, g.scale, 0.5f), qz = __fmaf_rn(c.uz, g.scale, 0.5f);
        const uint32_t ix = (uint32_t)floorf(qx), iy = (uint32_t)floorf(qy), iz = (uint32_t)floorf(qz);
        G.fx[q] = __fsub_rn(qx, (float)ix); G.fy[q] = __fsub_rn(qy, (float)iy); G.fz[q] = __fsub_rn(qz, (float)iz);
        // planes (i, j): xy = (ix, iy), yz = (iy, iz), xz = (ix, iz)   (split_xyz, network.py:208-212)
        if (QUAD) {
            const uint32_t my = iy * g.mul + g.off, mz = iz * g.mul + g.off;
            G.v[q][0] = __ldg(reinterpret_cast<const float4 *>(t_xy) + (ix + my));
            G.v[q][1] = __ldg(reinterpret_cast<const float4 *>(t_yz) + (iy + mz));
            G.v[q][2] = __ldg(reinterpret_cast<const float4 *>(t_xz) + (ix + mz));
            continue;
        }
        const uint32_t my0 = iy * g.mul, my1 = my0 + g.mul, mz0 = iz * g.mul, mz1 = mz0 + g.mul, ix1 = ix + 1u, iy1 = iy + 1u;
        uint32_t e[3][4];
        if (g.mask == 0xffffffffu) {
            e[0][0] = ix + my0; e[0][1] = ix1 + my0; e[0][2] = ix + my1; e[0][3] = ix1 + my1;
            e[1][0] = iy + mz0; e[1][1] = iy1 + mz0; e[1][2] = iy + mz1; e[1][3] = iy1 + mz1;
            e[2][0] = ix + mz0; e[2][1] = ix1 + mz0; e[2][2] = ix + mz1; e[2][3] = ix1 + mz1;
        } else {
            e[0][0] = (ix ^ my0) & g.mask; e[0][1] = (ix1 ^ my0) & g.mask; e[0][2] = (ix ^ my1) & g.mask; e[0][3] = (ix1 ^ my1) & g.mask;
            e[1][0] = (iy ^ mz0) & g.mask; e[1][1] = (iy1 ^ mz0) & g.mask; e[1][2] = (iy ^ mz1) & g.mask; e[1][3] = (iy1 ^ mz1) & g.mask;
            e[2][0] = (ix ^ mz0) & g.mask; e[2][1] = (ix1 ^ mz0) & g.mask; e[2][2] = (ix ^ mz1) & g.mask; e[2][3] = (ix1 ^ mz1) & g.mask;
        }
        const float *b0 = t_xy + g.off, *b1 = t_yz + g.off, *b2 = t_xz + g.off;
        G.v[q][0] = make_float4(__ldg(b0 + e[0][0]), __ldg(b0 + e[0][1]), __ldg(b0 + e[0][2]), __ldg(b0 + e[0][3]));
        G.v[q][1] = make_float4(__ldg(b1 + e[1][0]), __ldg(b1 + e[1][1]), __ldg(b1 + e[1][2]), __ldg(b1 + e[1][3]));
        G.v[q][2] = make_float4(__ldg(b2 + e[2][0]), __ldg(b2 + e[2][1]), __ldg(b2 + e[2][2]), __ldg(b2 + e[2][3]));
    }
}
__device__ __forceinline__ float blend4(const float4 &v, float fx, float gx, float fy, float gy) {
    float r = __fmaf_rn(__fmul_rn(gx, gy), v.x, 0.0f);
    r = __fmaf_rn(__fmul_rn(fx, gy), v.y, r);
    r = __fmaf_rn(__fmul_rn(gx, fy), v.z, r);
    r = __fmaf_rn(__fmul_rn(fx, fy), v.w, r);
    return r;
}
// blend + store: levels (2k, 2k+1) of plane p are one packed half2 word (word p*6 + k) of the sample's row; out-of-range planes store 0
// (gridencoder.cu:98-122).  `row` = the sample's row base inside the X tile, r7 = row index & 7 (SWIZZLE_128B chunk permutation).
__device__ __forceinline__ void gather_finish(const GatherTrip &G, uint32_t ok, uint8_t *row, uint32_t r7, uint32_t k) {
    float f[2][3];
#pragma unroll
    for (uint32_t q = 0; q < 2; q++) {
        const float fx = G.fx[q], fy = G.fy[q], fz = G.fz[q];
        const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy), gz = __fsub_rn(1.0f, fz);
        f[q][0] = blend4(G.v[q][0], fx, gx, fy, gy);
        f[q][1] = blend4(G.v[q][1], fy, gy, fz, gz);
        f[q][2] = blend4(G.v[q][2], fx, gx, fz, gz);
    }
#pragma unroll
    for (uint32_t p = 0; p < 3; p++) {
        const uint32_t word = p * 6u + k;
        const uint32_t v = ((ok >> p) & 1u) ? pack2(f[0][p], f[1][p]) : 0u;
        *reinterpret_cast<uint32_t *>(row + (((word >> 2) ^ r7) << 4) + (word & 3u) * 4u) = v;
    }
}

// real spherical harmonics, degree 4 (16 terms), fp32 — same polynomials as shencoder.cu:44-67
__device__ __forceinline__ void sh4(float x, float y, float z, float (&o)[16]) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y; o[2] = 0.48860251190291987f * z; o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy; o[5] = -1.0925484305920792f * yz; o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz; o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2); o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2); o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2); o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// One layer = `ksteps` K=16 MMAs into the same accumulator.  A K step advances both K-major SWIZZLE_128B operands by 32 bytes, i.e. +2 in the
// descriptors' start-address field (the operands are < 256 KB into shared memory, so the field never carries).
__device__ __forceinline__ void issue_mma(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, uint32_t ksteps, uint32_t N, bool accumulate) {
    const uint32_t idesc = idesc_f16(128, N);
    uint64_t da = smem_desc_sw128(a_saddr), db = smem_desc_sw128(b_saddr);
#pragma unroll 1
    for (uint32_t k = 0; k < ksteps; k++, da += 2, db += 2) mma_f16_ss(d_tmem, da, db, idesc, accumulate || k > 0);
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
// Tiles of one launch -> CTAs: every CTA takes a CONTIGUOUS, balanced share [lo, hi) of the tiles (shares differ by at most one tile) and its warpgroups
// walk it round-robin.  (A grid-strided walk hands the last round to the first CTAs only, four tiles each: SMs finished between 36 k and 52 k cycles
// in the ncu capture of a 1700-tile launch.)  Neighbouring tiles — neighbouring samples of the same rays — stay on one SM and share its L1.
// A launch with few tiles (the grid is sized for the buffer's capacity, the live count is only known on the device) is packed onto ceil(tiles / WGS) CTAs,
// one tile per warpgroup, and the other CTAs leave at once: their SMs stay free for the kernels of the other frames in flight.
__device__ __forceinline__ void cta_tile_share(uint32_t n_tiles, uint32_t wgs, int sched, uint32_t wg, uint32_t &first, uint32_t &end, uint32_t &stride) {
    if (sched & 2) {            // grid-strided walk (A/B: B2N_HEAD_SCHED=2)
        first = blockIdx.x * wgs + wg; end = n_tiles; stride = gridDim.x * wgs;
        if (blockIdx.x * wgs >= n_tiles) end = 0;
        return;
    }
    const uint32_t ctas = (sched & 1) ? gridDim.x : min(gridDim.x, (n_tiles + wgs - 1) / wgs);
    first = end = 0; stride = wgs;
    if (blockIdx.x >= ctas) return;
    first = (uint32_t)(((uint64_t)blockIdx.x * n_tiles) / ctas) + wg;
    end = (uint32_t)(((uint64_t)(blockIdx.x + 1) * n_tiles) / ctas);
}

struct HeadSmem {                       // lives after the 1024-aligned weight image and operand tiles
    HeadLvl lvl[12];
    __half2 enc_a_h2[16];               // the audio code in fp16, packed in pairs
    float eye_w1[16], unc_w1[32];
    uint32_t ind_p[2];                  // fp16-rounded individual code, packed (zeros without a code)
    float ind_h[4];                     // fp16-rounded individual code (training: part of the saved color_net input)
    float eye_val;
    uint32_t n_valid;
    uint32_t tmem_base;
    uint64_t bar_w;                     // weight image landed
    uint64_t bar_mma[HG_WGS];           // per-warpgroup MMA completion
};

// Register budget: the inference instantiation is held to 128 registers (it used 153 when allowed 170, with no spills at 112 either): next to a head CTA
// an SM then has 16 384 free registers instead of 4 096, so the small march / composite / init CTAs of the OTHER frames in flight run beside the head kernel
// instead of waiting for its tail — 2896 -> 3022 frames/s with the head kernel's own time unchanged (0.285 ms per frame).  The training instantiation keeps
// the full budget (168 registers; its kept-activation stores need them).
template <bool SAVE, bool QUAD>
__global__ void __maxnreg__(SAVE ? 168 : 128) k_head_forward(const __grid_constant__ HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // pointer arithmetic keeps the shared address space visible to the compiler
    uint8_t *s_w = base;
    uint8_t *s_tiles = base + HW_BYTES;
    HeadSmem &S = *reinterpret_cast<HeadSmem *>(s_tiles + HG_WGS * 3 * HG_TILE_BYTES);

    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u, warp = tid >> 5;
    // nothing to do (a loop iteration after the frame finished): leave before touching TMEM / the weight image
    const uint32_t n_valid_early = a.n_valid ? (uint32_t)max(0, min((int)a.M, __ldg(a.n_valid))) : a.M;
    uint32_t tile_first, tile_hi, tile_stride;
    cta_tile_share((n_valid_early + HG_TILE - 1) / HG_TILE, HG_WGS, a.sched, wg, tile_first, tile_hi, tile_stride);
    if (tile_hi == 0) return;            // nothing for this CTA (uniform over the CTA)
    uint8_t *sXb = s_tiles + wg * 3 * HG_TILE_BYTES, *sH = sXb + 2 * HG_TILE_BYTES;      // X[0], X[1], H

    // ---- one-time setup ------------------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(&S.bar_w, 1);
        for (int g = 0; g < (int)HG_WGS; g++) mbar_init(&S.bar_mma[g], 1);
        fence_mbar_init();
        mbar_expect_tx(&S.bar_w, HW_BYTES);
        bulk_g2s(s_w, a.wimg, HW_BYTES, &S.bar_w);
        S.n_valid = a.n_valid ? (uint32_t)max(0, min((int)a.M, *a.n_valid)) : a.M;
        S.eye_val = a.eye ? a.eye[0] : 0.0f;
    }
    if (warp == 1) tmem_alloc(&S.tmem_base, 512);
    if (tid >= 128 && tid < 144) S.enc_a_h2[tid - 128] = __floats2half2_rn(a.enc_a[2 * (tid - 128)], a.enc_a[2 * (tid - 128) + 1]);
    if (tid >= 160 && tid < 176) S.eye_w1[tid - 160] = a.wsmall[HS_EYE_W1 + tid - 160];
    if (tid >= 192 && tid < 224) S.unc_w1[tid - 192] = a.wsmall[HS_UNC_W1 + tid - 192];
    if (tid == 256) {                   // individual code -> two packed half2 words (K columns 16..19 of P6's second operand)
        S.ind_p[0] = a.ind_code ? pack2(a.ind_code[0], a.ind_code[1]) : 0u;
        S.ind_p[1] = a.ind_code ? pack2(a.ind_code[2], a.ind_code[3]) : 0u;
    }
    if (tid >= 320 && tid < 332) S.lvl[tid - 320] = a.lvl[tid - 320];
    if (tid >= 332 && tid < 336) S.ind_h[tid - 332] = a.ind_code ? round_h(a.ind_code[tid - 332]) : 0.0f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // (the weight image is only read by the tensor pipe: the issuing thread waits for it before its first MMA, the first gather runs meanwhile)

    const uint32_t n_valid = S.n_valid;
    const uint32_t n_tiles = tile_hi;                                        // this CTA's share ends here
    const uint32_t tmem_wg = S.tmem_base + wg * HG_TMEM_COLS;               // this warpgroup's columns
    const uint32_t tmem_ld = tmem_wg + (((warp & 3u) * 32u) << 16);         // + this warp's lane quarter
    const uint32_t sH_a = smem_u32(sH), sW_a = smem_u32(s_w);
    uint64_t *bar = &S.bar_mma[wg];
    uint32_t phase = 0;
    const float *t_xy = a.tab[0], *t_yz = a.tab[1], *t_xz = a.tab[2];
    const uint32_t r7 = t & 7u;
    const uint32_t row_off = (t >> 3) * 1024u + r7 * 128u;                  // this sample's row inside a tile

    auto sync_wg = [&]() { bar_sync(1 + wg, 128); };
    auto mma_done = [&]() { mbar_wait(bar, phase); phase ^= 1u; fence_after_sync(); };
    // operands written by this warpgroup's threads -> visible to the tensor pipe, then one thread issues
    auto publish = [&]() { fence_before_sync(); fence_proxy_async(); sync_wg(); };
    auto row_live = [&](uint32_t m) {
        bool live = m < n_valid;
        // frame mode: a row whose march slot was not produced (delta == 0) is ignored by the composite
        if (live && a.live_deltas) live = __ldg(a.live_deltas + 2 * (size_t)m) != 0.0f;
        return live;
    };
    // (x + bound) / (2 bound), fp32, like GridEncoder.forward (grid.py:143); out-of-range coordinates give zero features
    // (gridencoder.cu:98-122) — they are clamped for addressing and masked when stored
    auto make_coord = [&](float px, float py, float pz, bool live) {
        float ux, uy, uz;
        if (a.inv_two_bound != 0.0f) {       // 2 * bound is a power of two: the division is an exact scaling
            ux = __fmul_rn(__fadd_rn(px, a.bound), a.inv_two_bound); uy = __fmul_rn(__fadd_rn(py, a.bound), a.inv_two_bound); uz = __fmul_rn(__fadd_rn(pz, a.bound), a.inv_two_bound);
        } else {
            const float two_b = __fmul_rn(2.0f, a.bound);
            ux = __fdiv_rn(__fadd_rn(px, a.bound), two_b); uy = __fdiv_rn(__fadd_rn(py, a.bound), two_b); uz = __fdiv_rn(__fadd_rn(pz, a.bound), two_b);
        }
        const bool okx = !(ux < 0.0f || ux > 1.0f), oky = !(uy < 0.0f || uy > 1.0f), okz = !(uz < 0.0f || uz > 1.0f);
        SampleCoord c;
        c.ux = okx ? ux : 0.0f; c.uy = oky ? uy : 0.0f; c.uz = okz ? uz : 0.0f;
        c.ok = live ? ((okx && oky ? 1u : 0u) | (oky && okz ? 2u : 0u) | (okx && okz ? 4u : 0u)) : 0u;
        return c;
    };
    auto zero_k_padding = [&](uint8_t *tile) {        // words 18..23 (features 36..47) of the row
        *reinterpret_cast<uint2 *>(tile + row_off + ((4u ^ r7) << 4) + 8u) = make_uint2(0u, 0u);
        *reinterpret_cast<uint4 *>(tile + row_off + ((5u ^ r7) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    };

    uint32_t tile = tile_first;
    uint32_t buf = 0;
    // ---- pipeline prologue: the first tile's features (later tiles are gathered underneath the previous tile's MMA phases) -------
    if (tile < n_tiles) {
        const uint32_t m0 = tile * HG_TILE + t;
        const bool live0 = row_live(m0);
        float px = 0, py = 0, pz = 0;
        if (live0) { px = __ldcs(a.xyzs + 3 * (size_t)m0); py = __ldcs(a.xyzs + 3 * (size_t)m0 + 1); pz = __ldcs(a.xyzs + 3 * (size_t)m0 + 2); }
        const SampleCoord c = make_coord(px, py, pz, live0);
        GatherTrip GA, GB;                  // two trips in flight
        gather_issue<QUAD>(GA, t_xy, t_yz, t_xz, &S.lvl[0], c);
#pragma unroll 1
        for (uint32_t k = 0; k < 6; k += 2) {
            gather_issue<QUAD>(GB, t_xy, t_yz, t_xz, &S.lvl[2 * k + 2], c);
            gather_finish(GA, c.ok, sXb + row_off, r7, k);
            if (k + 2 < 6) gather_issue<QUAD>(GA, t_xy, t_yz, t_xz, &S.lvl[2 * k + 4], c);
            gather_finish(GB, c.ok, sXb + row_off, r7, k + 1);
        }
        zero_k_padding(sXb);
        if (SAVE) {      // enc_x (36 halves + zero padding = chunks 0..4 of the feature tile) -> x36 rows and the first 80 bytes of the sigma-input rows
            const size_t tr0 = (size_t)tile * HG_TILE;
            const uint32_t rv = (uint32_t)min((size_t)HG_TILE, (size_t)a.M - tr0);
            warp_chunks_out<5>(sXb, 0, (warp & 3u) * 32u, reinterpret_cast<uint8_t *>(a.sv.x36) + tr0 * 80, 80, rv);
            warp_chunks_out<5>(sXb, 0, (warp & 3u) * 32u, reinterpret_cast<uint8_t *>(a.sv.s_in) + tr0 * 160, 160, rv);
        }
    }
    for (; tile < n_tiles; tile += tile_stride) {
        uint8_t *sX = sXb + buf * HG_TILE_BYTES, *sXn = sXb + (buf ^ 1u) * HG_TILE_BYTES;
        const uint32_t sX_a = smem_u32(sX);
        const uint32_t m = tile * HG_TILE + t;
        const bool live = row_live(m);
        float dxv = 0, dyv = 0, dzv = 1;
        if (live) { dxv = __ldcs(a.dirs + 3 * (size_t)m); dyv = __ldcs(a.dirs + 3 * (size_t)m + 1); dzv = __ldcs(a.dirs + 3 * (size_t)m + 2); }
        const bool has_next = tile + tile_stride < n_tiles;                  // uniform over the warpgroup
        float npx = 0, npy = 0, npz = 0;
        bool nlive = false;
        if (has_next) {
            const uint32_t mn = (tile + tile_stride) * HG_TILE + t;
            nlive = row_live(mn);
            if (nlive) { npx = __ldcs(a.xyzs + 3 * (size_t)mn); npy = __ldcs(a.xyzs + 3 * (size_t)mn + 1); npz = __ldcs(a.xyzs + 3 * (size_t)mn + 2); }
        }
        uint8_t *rown = sXn + row_off;
        GatherTrip G;
        // training: rows of the saved activations for this tile's sample (m) and for the next tile's sample (features only)
        const bool sv_on = SAVE && live;
        const size_t tile_row0 = (size_t)tile * HG_TILE;
        const uint32_t rows_valid = SAVE ? (uint32_t)min((size_t)HG_TILE, (size_t)a.M - tile_row0) : 0u, wrow0 = (warp & 3u) * 32u;

        uint4 *c_in_q = sv_on ? reinterpret_cast<uint4 *>(a.sv.c_in) + (size_t)m * 11 : nullptr;
        float unc_logit = 0.0f;
        publish();
        // ---- P1: [aud hidden | eye hidden | sigma hidden (enc_x part)] = X * WA -------------------------------------------
        if (t == 0) { mbar_wait(&S.bar_w, 0); fence_after_sync(); issue_mma(tmem_wg + TC_A, sX_a, sW_a + HW_A, 3, 144, false); mma_commit(bar); }
        const SampleCoord cn = make_coord(npx, npy, npz, nlive);
        if (has_next) { zero_k_padding(sXn); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[0], cn); }      // trip 0 in flight under P1
        mma_done();
        float eye_att, unc_out = 0.6931471805599453f;       // testing: log(1 + e^0) (network.py:245,278)
        {
            hidden_epilogue<true, false>(tmem_ld + TC_A, sH, t, nullptr);
            if (SAVE) warp_rows_out(sH, wrow0, reinterpret_cast<uint8_t *>(a.sv.ha) + tile_row0 * 128, 128, rows_valid);
            uint32_t e16[16];
            ld16(tmem_ld + TC_EYE, e16); wait_ld();
            float dot = 0.0f;
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {      // relu(round_h(x)) of a pair in one cvt.rn.relu.f16x2, widened back for the fp32 dot product
                w[j] = pack2_relu(__uint_as_float(e16[2 * j]), __uint_as_float(e16[2 * j + 1]));
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[j]));
                dot = fmaf(f.x, S.eye_w1[2 * j], dot); dot = fmaf(f.y, S.eye_w1[2 * j + 1], dot);
            }
            if (sv_on) {
                *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 6)) = make_uint4(w[0], w[1], w[2], w[3]);      // chunks 6, 7 of the feature tile are not operands (K = 48)
                *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 7)) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            if (SAVE) warp_chunks_out<2>(sX, 6, wrow0, reinterpret_cast<uint8_t *>(a.sv.he) + tile_row0 * 32, 32, rows_valid);
            // sigmoid evaluated on the fp16 logit, result rounded to fp16 (torch.sigmoid on a half tensor)
            eye_att = round_h(1.0f / (1.0f + expf(-round_h(dot))));
        }
        publish();
        // ---- P2: att = H * WB ; [unc hidden = X * WU] --------------------------------------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma(tmem_wg + TC_A, sH_a, sW_a + HW_B, 4, 32, false);
            if (a.has_unc) issue_mma(tmem_wg + TC_A + 32, sX_a, sW_a + HW_U, 3, 32, false);
            mma_commit(bar);
        }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 0); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[2], cn); }      // trip 0 lands (issued a phase ago), trip 1 leaves
        mma_done();
        float amb_aud;
        {
            uint32_t acc[32];
            if (a.has_unc) {
                ld32(tmem_ld + TC_A + 32, acc); wait_ld();
                float du = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; j++) du = fmaf(fmaxf(round_h(__uint_as_float(acc[j])), 0.0f), S.unc_w1[j], du);
                du = round_h(du);
                unc_logit = du;
                unc_out = logf(1.0f + expf(du));                        // torch.log(1 + torch.exp(.)) in fp32 (network.py:278)
                if (SAVE && a.sv.hu) {
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        *reinterpret_cast<uint4 *>(sH + sw128_offset(t, 4 + c)) = make_uint4(pack2_relu(__uint_as_float(acc[8 * c]), __uint_as_float(acc[8 * c + 1])), pack2_relu(__uint_as_float(acc[8 * c + 2]), __uint_as_float(acc[8 * c + 3])),
                                          pack2_relu(__uint_as_float(acc[8 * c + 4]), __uint_as_float(acc[8 * c + 5])), pack2_relu(__uint_as_float(acc[8 * c + 6]), __uint_as_float(acc[8 * c + 7])));
                }
            }
            ld32(tmem_ld + TC_A, acc); wait_ld();
            float n2 = 0.0f;
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t ah = pack2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));      // att pair in fp16 (the Linear's output dtype)
                const __half2 a2 = *reinterpret_cast<const __half2 *>(&ah);
                const float2 f = __half22float2(a2);
                n2 = fmaf(f.x, f.x, n2); n2 = fmaf(f.y, f.y, n2);
                // enc_w = enc_a * att on half tensors (network.py:285): the product of two halves is exact in fp32, so one HMUL2 rounds exactly like
                // multiplying in fp32 and converting
                const __half2 p = __hmul2(a2, S.enc_a_h2[j]);
                w[j] = *reinterpret_cast<const uint32_t *>(&p);
            }
            amb_aud = sqrtf(n2);                                                    // .norm(dim=-1) in fp32 (network.py:308)
            // EW operand into X (the features are consumed): chunks 0..3 = enc_w, chunk 4 = [e, 0..], chunk 5 = 0   (K = 33 padded to 48)
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(sX + sw128_offset(t, c)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
            const float e = a.eye ? S.eye_val * eye_att : 0.0f;                     // e = e * eye_att (network.py:291)
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 4)) = make_uint4(pack2(e, 0.0f), 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 5)) = make_uint4(0u, 0u, 0u, 0u);
            if (SAVE) {
                // the aud hidden tile H is dead (P2 read it): stage att in its chunks 0..3 (unc hidden sits in 4..7), then everything leaves coalesced
#pragma unroll
                for (int c = 0; c < 4; c++)
                    *reinterpret_cast<uint4 *>(sH + sw128_offset(t, c)) =
                        make_uint4(pack2(__uint_as_float(acc[8 * c]), __uint_as_float(acc[8 * c + 1])), pack2(__uint_as_float(acc[8 * c + 2]), __uint_as_float(acc[8 * c + 3])),
                                   pack2(__uint_as_float(acc[8 * c + 4]), __uint_as_float(acc[8 * c + 5])), pack2(__uint_as_float(acc[8 * c + 6]), __uint_as_float(acc[8 * c + 7])));
                warp_chunks_out<4>(sH, 0, wrow0, reinterpret_cast<uint8_t *>(a.sv.att) + tile_row0 * 64, 64, rows_valid);
                if (a.sv.hu) warp_chunks_out<4>(sH, 4, wrow0, reinterpret_cast<uint8_t *>(a.sv.hu) + tile_row0 * 64, 64, rows_valid);
                // [enc_w 32 | e | 0 x7] = chunks 0..4 of X -> bytes 80..159 of the sigma-input rows
                warp_chunks_out<5>(sX, 0, wrow0, reinterpret_cast<uint8_t *>(a.sv.s_in) + tile_row0 * 160 + 80, 160, rows_valid);
            }
        }
        publish();
        // ---- P3: sigma hidden += [enc_w, e] * WC -------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma(tmem_wg + TC_S, sX_a, sW_a + HW_C, 3, 64, true); mma_commit(bar); }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 1); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[4], cn); }      // trip 1 lands (issued a phase ago), trip 2 leaves
        mma_done();
        hidden_epilogue<true, false>(tmem_ld + TC_S, sH, t, nullptr);
        if (SAVE) warp_rows_out(sH, wrow0, reinterpret_cast<uint8_t *>(a.sv.h1) + tile_row0 * 128, 128, rows_valid);
        publish();
        // ---- P4: sigma layer 1 --------------------------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma(tmem_wg + TC_A, sH_a, sW_a + HW_D, 4, 64, false); mma_commit(bar); }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 2); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[6], cn); }      // trip 2 lands (issued a phase ago), trip 3 leaves
        mma_done();
        hidden_epilogue<true, false>(tmem_ld + TC_A, sH, t, nullptr);
        if (SAVE) warp_rows_out(sH, wrow0, reinterpret_cast<uint8_t *>(a.sv.h2) + tile_row0 * 128, 128, rows_valid);
        publish();
        // ---- P5: sigma layer 2: cols 0..63 = geo_feat, col 64 = density logit (rows rotated at pack time) -----------------
        if (t == 0) { fence_after_sync(); issue_mma(tmem_wg + TC_A, sH_a, sW_a + HW_E, 4, 80, false); mma_commit(bar); }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 3); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[8], cn); }      // trip 3 lands (issued a phase ago), trip 4 leaves
        mma_done();
        float sigma;
        {
            hidden_epilogue<false, false>(tmem_ld + TC_A, sH, t, nullptr);
            if (SAVE) warp_rows_out(sH, wrow0, reinterpret_cast<uint8_t *>(a.sv.c_in) + tile_row0 * 176 + 32, 176, rows_valid);      // geo_feat = halves 16..79 of the color input
            uint32_t s16[16];
            ld16(tmem_ld + TC_A + 64, s16); wait_ld();
            sigma = expf(round_h(__uint_as_float(s16[0]))) * a.density_scale;       // torch.exp(h[..., 0]) in fp32 (network.py:301)
            // view-direction SH into X chunks 0,1 (the 16-wide K step of color layer 0)
            float shv[16];
            sh4(dxv, dyv, dzv, shv);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = pack2(shv[2 * j], shv[2 * j + 1]);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 0)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 1)) = make_uint4(w[4], w[5], w[6], w[7]);
            // the individual code as K columns 16..19 of the same operand (c.repeat(N, 1) concatenated behind the SH terms, network.py:270): its part of the
            // layer rides on the tensor core as a second K step instead of a 64-term bias add in the epilogue
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 2)) = make_uint4(S.ind_p[0], S.ind_p[1], 0u, 0u);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 3)) = make_uint4(0u, 0u, 0u, 0u);
            if (SAVE) warp_chunks_out<2>(sX, 0, wrow0, reinterpret_cast<uint8_t *>(a.sv.c_in) + tile_row0 * 176, 176, rows_valid);
            if (sv_on) c_in_q[10] = make_uint4(pack2(S.ind_h[0], S.ind_h[1]), pack2(S.ind_h[2], S.ind_h[3]), 0u, 0u);
        }
        publish();
        // ---- P6: color layer 0 = geo * WF0 + sh * WF1 (+ ind-code bias in the epilogue) -------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma(tmem_wg + TC_S, sH_a, sW_a + HW_F0, 4, 64, false);
            issue_mma(tmem_wg + TC_S, sX_a, sW_a + HW_F1, 2, 64, true);
            mma_commit(bar);
        }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 4); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[10], cn); }      // trip 4 lands (issued a phase ago), trip 5 leaves
        mma_done();
        hidden_epilogue<true, false>(tmem_ld + TC_S, sH, t, nullptr);
        if (SAVE) warp_rows_out(sH, wrow0, reinterpret_cast<uint8_t *>(a.sv.hc) + tile_row0 * 128, 128, rows_valid);
        publish();
        // ---- P7: color layer 1 (N padded 3 -> 16) -----------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma(tmem_wg + TC_A, sH_a, sW_a + HW_G, 4, 16, false); mma_commit(bar); }
        if (has_next) gather_finish(G, cn.ok, rown, r7, 5);
        mma_done();
        {
            uint32_t c16[16];
            ld16(tmem_ld + TC_A, c16); wait_ld();
            if (live) {
                float rgb[3], sg[3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    // torch.sigmoid(h_color) * (1 + 2*0.001) - 0.001 evaluated on half tensors (network.py:275)
                    const float s = round_h(1.0f / (1.0f + expf(-round_h(__uint_as_float(c16[j])))));
                    sg[j] = s;
                    rgb[j] = round_h(round_h(s * 1.002f) - 0.001f);
                }
                if (SAVE) reinterpret_cast<uint4 *>(a.sv.misc)[m] = make_uint4(pack2(sg[0], sg[1]), pack2(sg[2], eye_att), pack2(unc_logit, 1.0f), 0u);      // column 5 = 1: a ones column for column sums through the wgrad kernel
                if (a.sigmas) __stcs(a.sigmas + m, sigma);
                if (a.rgbs) { __stcs(a.rgbs + 3 * (size_t)m, rgb[0]); __stcs(a.rgbs + 3 * (size_t)m + 1, rgb[1]); __stcs(a.rgbs + 3 * (size_t)m + 2, rgb[2]); }
                if (a.amb_aud) __stcs(a.amb_aud + m, amb_aud);
                if (a.amb_eye) __stcs(a.amb_eye + m, eye_att);
                if (a.unc) __stcs(a.unc + m, unc_out);
            }
        }
        if (SAVE && has_next) {      // the next tile's feature rows are complete (this warp gathered its own 32 rows)
            const size_t tr0 = (size_t)(tile + tile_stride) * HG_TILE;
            const uint32_t rv = (uint32_t)min((size_t)HG_TILE, (size_t)a.M - tr0);
            warp_chunks_out<5>(sXn, 0, wrow0, reinterpret_cast<uint8_t *>(a.sv.x36) + tr0 * 80, 80, rv);
            warp_chunks_out<5>(sXn, 0, wrow0, reinterpret_cast<uint8_t *>(a.sv.s_in) + tr0 * 160, 160, rv);
        }
        buf ^= 1u;
        // the next tile's publish() orders this tile's TMEM reads (fence::before_thread_sync + warpgroup barrier) before its first MMA
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(S.tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// inference kernel, FOUR tiles in flight per SM: hidden activations live in TENSOR MEMORY
// ---------------------------------------------------------------------------------------------------
// k_head_forward above keeps three tiles per SM in flight: each warpgroup needs two feature tiles and one hidden-activation tile in shared memory (48 KB), and
// 3 x 48 + 70 KB of weights fill the SM.  Its ncu profile is latency-bound with 12 warps (issue slots 39 % busy, tensor pipe 15 %): more tiles in flight is
// the lever.  Here the hidden activations never touch shared memory: an epilogue reads the fp32 accumulator columns (tcgen05.ld), applies the activation,
// packs to fp16 and writes the next layer's A operand straight back to tensor memory (tcgen05.st, two halves per 32-bit column, row = lane), and the next
// layer is a tcgen05.mma whose A operand is read from TMEM.  A warpgroup then owns 32 KB of shared memory (the double-buffered feature tile) and 128 TMEM
// columns: 4 warpgroups x (32 KB, 128 columns) = 128 KB + 70 KB of weights and all 512 columns.  No swizzled st.shared / address arithmetic in the
// epilogues, no generic->async proxy fence for the activations.
// TMEM columns of a warpgroup: A operand [0, 32) | accumulators inside [32, 128):
//   P1  [aud hidden 64 | eye hidden 16] -> [32, 112)          (features from shared memory)
//   P2  att 32 -> [32, 64)  = A(aud hidden) x WB ;  sigma hidden (enc_x part) 64 -> [64, 128) = features x WA[80:144]
//   P3  [64, 128) += A([enc_a * att | e]) x WC                 P4  [32, 96)  = A x WD            P5  [32, 112) = A x WE  (geo 64 | logit)
//   P6  [32, 96)  = A(geo) x WF0 + [SH | ind code] x WF1       P7  [32, 48)  = A x WG
// Same arithmetic as k_head_forward<false, QUAD> (same MMAs on the same operands, same epilogue math) — tests compare the two bit for bit.
constexpr uint32_t H4_WGS = 4, H4_THREADS = H4_WGS * 128, H4_COLS = 128, H4_ACC = 32;
struct Head4Smem {
    HeadLvl lvl[12];
    __half2 enc_a_h2[16];
    float eye_w1[16];
    uint32_t ind_p[2];
    float eye_val;
    uint32_t n_valid, tmem_base;
    uint64_t bar_w, bar_mma[H4_WGS];
};

// 64 accumulator columns of this thread's row -> (ReLU) -> 32 packed fp16 columns of the A operand region
template <bool RELU>
__device__ __forceinline__ void hidden_to_tmem(uint32_t t_acc, uint32_t t_a) {
#pragma unroll 1
    for (uint32_t cb = 0; cb < 64; cb += 32) {
        uint32_t acc[32];
        ld32(t_acc + cb, acc);
        wait_ld();
        uint32_t w[16];
#pragma unroll
        for (uint32_t j = 0; j < 16; j++) {
            const float x0 = __uint_as_float(acc[2 * j]), x1 = __uint_as_float(acc[2 * j + 1]);
            w[j] = RELU ? pack2_relu(x0, x1) : pack2(x0, x1);
        }
        st16(t_a + (cb >> 1), w);
    }
}
__device__ __forceinline__ void issue_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_saddr, uint32_t ksteps, uint32_t N, bool accumulate) {
    const uint32_t idesc = idesc_f16(128, N);
    uint64_t db = smem_desc_sw128(b_saddr);
#pragma unroll 1
    for (uint32_t k = 0; k < ksteps; k++, a_tmem += 8, db += 2) mma_f16_ts(d_tmem, a_tmem, db, idesc, accumulate || k > 0);      // 16 halves = 8 columns per K step
}

template <bool QUAD>
__global__ void __launch_bounds__(H4_THREADS, 1) k_head_infer4(const __grid_constant__ HeadArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_w = base;
    uint8_t *s_tiles = base + HW_BYTES;
    Head4Smem &S = *reinterpret_cast<Head4Smem *>(s_tiles + H4_WGS * 2 * HG_TILE_BYTES);
    const uint32_t tid = threadIdx.x, wg = tid >> 7, t = tid & 127u, warp = tid >> 5;
    const uint32_t n_valid_early = a.n_valid ? (uint32_t)max(0, min((int)a.M, __ldg(a.n_valid))) : a.M;
    uint32_t tile_first, tile_hi, tile_stride;
    cta_tile_share((n_valid_early + HG_TILE - 1) / HG_TILE, H4_WGS, a.sched, wg, tile_first, tile_hi, tile_stride);
    if (tile_hi == 0) return;            // nothing for this CTA (uniform over the CTA)
    uint8_t *sXb = s_tiles + wg * 2 * HG_TILE_BYTES;

    if (tid == 0) {
        mbar_init(&S.bar_w, 1);
        for (int g = 0; g < (int)H4_WGS; g++) mbar_init(&S.bar_mma[g], 1);
        fence_mbar_init();
        mbar_expect_tx(&S.bar_w, HW_BYTES);
        bulk_g2s(s_w, a.wimg, HW_BYTES, &S.bar_w);
        S.n_valid = a.n_valid ? (uint32_t)max(0, min((int)a.M, *a.n_valid)) : a.M;
        S.eye_val = a.eye ? a.eye[0] : 0.0f;
    }
    if (warp == 1) tmem_alloc(&S.tmem_base, 512);
    if (tid >= 128 && tid < 144) S.enc_a_h2[tid - 128] = __floats2half2_rn(a.enc_a[2 * (tid - 128)], a.enc_a[2 * (tid - 128) + 1]);
    if (tid >= 160 && tid < 176) S.eye_w1[tid - 160] = a.wsmall[HS_EYE_W1 + tid - 160];
    if (tid == 256) {
        S.ind_p[0] = a.ind_code ? pack2(a.ind_code[0], a.ind_code[1]) : 0u;
        S.ind_p[1] = a.ind_code ? pack2(a.ind_code[2], a.ind_code[3]) : 0u;
    }
    if (tid >= 320 && tid < 332) S.lvl[tid - 320] = a.lvl[tid - 320];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // (the weight image is only read by the tensor pipe: the issuing thread waits for it before its first MMA, the first gather runs meanwhile)

    const uint32_t n_valid = S.n_valid;
    const uint32_t n_tiles = tile_hi;                                        // this CTA's share ends here
    const uint32_t tm = S.tmem_base + wg * H4_COLS;                          // this warpgroup's columns: A operand at +0, accumulators from +32
    const uint32_t lane_off = ((warp & 3u) * 32u) << 16;                    // this warp's lane quarter (tcgen05.ld / st)
    const uint32_t tA = tm + lane_off, tD = tm + H4_ACC + lane_off;         // per-thread addresses
    const uint32_t mA = tm, mD = tm + H4_ACC;                               // MMA operand / accumulator addresses (lane 0)
    const uint32_t sW_a = smem_u32(s_w);
    uint64_t *bar = &S.bar_mma[wg];
    uint32_t phase = 0;
    const float *t_xy = a.tab[0], *t_yz = a.tab[1], *t_xz = a.tab[2];
    const uint32_t r7 = t & 7u;
    const uint32_t row_off = (t >> 3) * 1024u + r7 * 128u;

    auto sync_wg = [&]() { bar_sync(1 + wg, 128); };
    auto mma_done = [&]() { mbar_wait(bar, phase); phase ^= 1u; fence_after_sync(); };
    // operands written by this warpgroup (shared memory: generic -> async proxy; tensor memory: tcgen05.st completion) -> visible to the tensor pipe
    auto publish = [&]() { wait_st(); fence_before_sync(); fence_proxy_async(); sync_wg(); };
    auto row_live = [&](uint32_t m) {
        bool live = m < n_valid;
        if (live && a.live_deltas) live = __ldg(a.live_deltas + 2 * (size_t)m) != 0.0f;
        return live;
    };
    auto make_coord = [&](float px, float py, float pz, bool live) {
        float ux, uy, uz;
        if (a.inv_two_bound != 0.0f) {
            ux = __fmul_rn(__fadd_rn(px, a.bound), a.inv_two_bound); uy = __fmul_rn(__fadd_rn(py, a.bound), a.inv_two_bound); uz = __fmul_rn(__fadd_rn(pz, a.bound), a.inv_two_bound);
        } else {
            const float two_b = __fmul_rn(2.0f, a.bound);
            ux = __fdiv_rn(__fadd_rn(px, a.bound), two_b); uy = __fdiv_rn(__fadd_rn(py, a.bound), two_b); uz = __fdiv_rn(__fadd_rn(pz, a.bound), two_b);
        }
        const bool okx = !(ux < 0.0f || ux > 1.0f), oky = !(uy < 0.0f || uy > 1.0f), okz = !(uz < 0.0f || uz > 1.0f);
        SampleCoord c;
        c.ux = okx ? ux : 0.0f; c.uy = oky ? uy : 0.0f; c.uz = okz ? uz : 0.0f;
        c.ok = live ? ((okx && oky ? 1u : 0u) | (oky && okz ? 2u : 0u) | (okx && okz ? 4u : 0u)) : 0u;
        return c;
    };
    auto zero_k_padding = [&](uint8_t *tile) {
        *reinterpret_cast<uint2 *>(tile + row_off + ((4u ^ r7) << 4) + 8u) = make_uint2(0u, 0u);
        *reinterpret_cast<uint4 *>(tile + row_off + ((5u ^ r7) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    };

    uint32_t tile = tile_first;
    uint32_t buf = 0;
    if (tile < n_tiles) {
        const uint32_t m0 = tile * HG_TILE + t;
        const bool live0 = row_live(m0);
        float px = 0, py = 0, pz = 0;
        if (live0) { px = __ldcs(a.xyzs + 3 * (size_t)m0); py = __ldcs(a.xyzs + 3 * (size_t)m0 + 1); pz = __ldcs(a.xyzs + 3 * (size_t)m0 + 2); }
        const SampleCoord c = make_coord(px, py, pz, live0);
        GatherTrip GA, GB;
        gather_issue<QUAD>(GA, t_xy, t_yz, t_xz, &S.lvl[0], c);
#pragma unroll 1
        for (uint32_t k = 0; k < 6; k += 2) {
            gather_issue<QUAD>(GB, t_xy, t_yz, t_xz, &S.lvl[2 * k + 2], c);
            gather_finish(GA, c.ok, sXb + row_off, r7, k);
            if (k + 2 < 6) gather_issue<QUAD>(GA, t_xy, t_yz, t_xz, &S.lvl[2 * k + 4], c);
            gather_finish(GB, c.ok, sXb + row_off, r7, k + 1);
        }
        zero_k_padding(sXb);
    }
    for (; tile < n_tiles; tile += tile_stride) {
        uint8_t *sX = sXb + buf * HG_TILE_BYTES, *sXn = sXb + (buf ^ 1u) * HG_TILE_BYTES;
        const uint32_t sX_a = smem_u32(sX);
        const uint32_t m = tile * HG_TILE + t;
        const bool live = row_live(m);
        float dxv = 0, dyv = 0, dzv = 1;
        if (live) { dxv = __ldcs(a.dirs + 3 * (size_t)m); dyv = __ldcs(a.dirs + 3 * (size_t)m + 1); dzv = __ldcs(a.dirs + 3 * (size_t)m + 2); }
        const bool has_next = tile + tile_stride < n_tiles;
        float npx = 0, npy = 0, npz = 0;
        bool nlive = false;
        if (has_next) {
            const uint32_t mn = (tile + tile_stride) * HG_TILE + t;
            nlive = row_live(mn);
            if (nlive) { npx = __ldcs(a.xyzs + 3 * (size_t)mn); npy = __ldcs(a.xyzs + 3 * (size_t)mn + 1); npz = __ldcs(a.xyzs + 3 * (size_t)mn + 2); }
        }
        uint8_t *rown = sXn + row_off;
        GatherTrip G;
        publish();
        // ---- P1: [aud hidden 64 | eye hidden 16] = X * WA[0:80] ---------------------------------------------------------------------
        if (t == 0) { mbar_wait(&S.bar_w, 0); fence_after_sync(); issue_mma(mD, sX_a, sW_a + HW_A, 3, 80, false); mma_commit(bar); }
        const SampleCoord cn = make_coord(npx, npy, npz, nlive);
        if (has_next) { zero_k_padding(sXn); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[0], cn); }
        mma_done();
        float eye_att;
        {
            hidden_to_tmem<true>(tD, tA);
            uint32_t e16[16];
            ld16(tD + 64, e16); wait_ld();
            float dot = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t w = pack2_relu(__uint_as_float(e16[2 * j]), __uint_as_float(e16[2 * j + 1]));
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
                dot = fmaf(f.x, S.eye_w1[2 * j], dot); dot = fmaf(f.y, S.eye_w1[2 * j + 1], dot);
            }
            eye_att = round_h(1.0f / (1.0f + expf(-round_h(dot))));
        }
        publish();
        // ---- P2: att = A(aud hidden) * WB -> [32, 64) ; sigma hidden (enc_x part) = X * WA[80:144] -> [64, 128) --------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma_ts(mD, mA, sW_a + HW_B, 4, 32, false);
            issue_mma(mD + 32, sX_a, sW_a + HW_A + 80 * 128, 3, 64, false);
            mma_commit(bar);
        }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 0); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[2], cn); }
        mma_done();
        float amb_aud;
        {
            uint32_t acc[32];
            ld32(tD, acc); wait_ld();
            float n2 = 0.0f;
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t ah = pack2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                const __half2 a2 = *reinterpret_cast<const __half2 *>(&ah);
                const float2 f = __half22float2(a2);
                n2 = fmaf(f.x, f.x, n2); n2 = fmaf(f.y, f.y, n2);
                const __half2 p = __hmul2(a2, S.enc_a_h2[j]);
                w[j] = *reinterpret_cast<const uint32_t *>(&p);
            }
            amb_aud = sqrtf(n2);
            // A operand of P3: [enc_w 32 | e | 0 x15] = 24 columns
            st16(tA, w);
            const float e = a.eye ? S.eye_val * eye_att : 0.0f;
            const uint32_t w8[8] = {pack2(e, 0.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            st8(tA + 16, w8);
        }
        publish();
        // ---- P3: sigma hidden += A([enc_w, e]) * WC ------------------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma_ts(mD + 32, mA, sW_a + HW_C, 3, 64, true); mma_commit(bar); }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 1); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[4], cn); }
        mma_done();
        hidden_to_tmem<true>(tD + 32, tA);
        publish();
        // ---- P4: sigma layer 1 -> [32, 96) --------------------------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma_ts(mD, mA, sW_a + HW_D, 4, 64, false); mma_commit(bar); }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 2); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[6], cn); }
        mma_done();
        hidden_to_tmem<true>(tD, tA);
        publish();
        // ---- P5: sigma layer 2 -> [32, 112): cols 0..63 geo_feat, col 64 density logit ------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma_ts(mD, mA, sW_a + HW_E, 4, 80, false); mma_commit(bar); }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 3); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[8], cn); }
        mma_done();
        float sigma;
        {
            hidden_to_tmem<false>(tD, tA);
            uint32_t s16[16];
            ld16(tD + 64, s16); wait_ld();
            sigma = expf(round_h(__uint_as_float(s16[0]))) * a.density_scale;
            float shv[16];
            sh4(dxv, dyv, dzv, shv);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = pack2(shv[2 * j], shv[2 * j + 1]);
            // the feature tile is dead (P2 read it last): [SH 16 | ind code 4 | 0 x12] in its chunks 0..3
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 0)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 1)) = make_uint4(w[4], w[5], w[6], w[7]);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 2)) = make_uint4(S.ind_p[0], S.ind_p[1], 0u, 0u);
            *reinterpret_cast<uint4 *>(sX + sw128_offset(t, 3)) = make_uint4(0u, 0u, 0u, 0u);
        }
        publish();
        // ---- P6: color layer 0 = A(geo) * WF0 + [SH | ind] * WF1 -> [32, 96) ------------------------------------------------------------
        if (t == 0) {
            fence_after_sync();
            issue_mma_ts(mD, mA, sW_a + HW_F0, 4, 64, false);
            issue_mma(mD, sX_a, sW_a + HW_F1, 2, 64, true);
            mma_commit(bar);
        }
        if (has_next) { gather_finish(G, cn.ok, rown, r7, 4); gather_issue<QUAD>(G, t_xy, t_yz, t_xz, &S.lvl[10], cn); }
        mma_done();
        hidden_to_tmem<true>(tD, tA);
        publish();
        // ---- P7: color layer 1 (N padded 3 -> 16) -> [32, 48) --------------------------------------------------------------------------------
        if (t == 0) { fence_after_sync(); issue_mma_ts(mD, mA, sW_a + HW_G, 4, 16, false); mma_commit(bar); }
        if (has_next) gather_finish(G, cn.ok, rown, r7, 5);
        mma_done();
        {
            uint32_t c16[16];
            ld16(tD, c16); wait_ld();
            if (live) {
                float rgb[3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float s = round_h(1.0f / (1.0f + expf(-round_h(__uint_as_float(c16[j])))));
                    rgb[j] = round_h(round_h(s * 1.002f) - 0.001f);
                }
                if (a.sigmas) __stcs(a.sigmas + m, sigma);
                if (a.rgbs) { __stcs(a.rgbs + 3 * (size_t)m, rgb[0]); __stcs(a.rgbs + 3 * (size_t)m + 1, rgb[1]); __stcs(a.rgbs + 3 * (size_t)m + 2, rgb[2]); }
                if (a.amb_aud) __stcs(a.amb_aud + m, amb_aud);
                if (a.amb_eye) __stcs(a.amb_eye + m, eye_att);
                if (a.unc) __stcs(a.unc + m, 0.6931471805599453f);       // testing: log(1 + e^0) (network.py:245,278)
            }
        }
        buf ^= 1u;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(S.tmem_base, 512);
}

size_t head4_smem_bytes() { return 1024 + HW_BYTES + (size_t)H4_WGS * 2 * HG_TILE_BYTES + sizeof(Head4Smem); }

size_t head_smem_bytes() { return 1024 + HW_BYTES + (size_t)HG_WGS * 3 * HG_TILE_BYTES + sizeof(HeadSmem); }

int launch_head_forward(const HeadArgs &a, cudaStream_t st, bool save, bool quad) {
    const size_t smem = head_smem_bytes();
    B2N_SMEM((k_head_forward<false, false>), smem);
    B2N_SMEM((k_head_forward<true, false>), smem);
    B2N_SMEM((k_head_forward<false, true>), smem);
    B2N_SMEM((k_head_forward<true, true>), smem);
    const uint32_t tiles = ceil_div<uint32_t>(a.M, HG_TILE);
    uint32_t sms = (uint32_t)sm_count();
    if (a.max_ctas > 0 && a.max_ctas < sms) sms = a.max_ctas;
    // inference without unc_net: the four-warpgroup kernel with the activations in tensor memory (B2N_HEAD_WG4=0 keeps the three-warpgroup kernel, for A/B)
    const char *wg4_env = getenv("B2N_HEAD_WG4");
    const bool wg4 = !(wg4_env && wg4_env[0] == '0');
    if (!save && !a.has_unc && wg4) {
        const size_t smem4 = head4_smem_bytes();
        B2N_SMEM(k_head_infer4<false>, smem4);
        B2N_SMEM(k_head_infer4<true>, smem4);
        uint32_t c4 = ceil_div<uint32_t>(tiles, H4_WGS);
        if (c4 > sms) c4 = sms;
        if (c4 == 0) return 0;
        if (quad) k_head_infer4<true><<<c4, H4_THREADS, smem4, st>>>(a); else k_head_infer4<false><<<c4, H4_THREADS, smem4, st>>>(a);
        return check_launch("head_forward(4 warpgroups)");
    }
    uint32_t ctas = ceil_div<uint32_t>(tiles, HG_WGS);
    if (ctas > sms) ctas = sms;
    if (ctas == 0) return 0;
    if (save) { if (quad) k_head_forward<true, true><<<ctas, HG_THREADS, smem, st>>>(a); else k_head_forward<true, false><<<ctas, HG_THREADS, smem, st>>>(a); }
    else { if (quad) k_head_forward<false, true><<<ctas, HG_THREADS, smem, st>>>(a); else k_head_forward<false, false><<<ctas, HG_THREADS, smem, st>>>(a); }
    return check_launch("head_forward");
}

// ---------------------------------------------------------------------------------------------------
// corner-quad image of the three tables (QUAD gather): quads[plane][qoff(level) + i + j * res] = the four corner values of cell (i, j),
// read through the reference's index function (gridencoder.cu:54-72, D = 2) — built by b2n_model_update, i.e. after every weight update.
// ---------------------------------------------------------------------------------------------------
struct QuadArgs { const float *tab[3]; HeadLvl lvl[12], qlvl[12]; uint32_t cells; float4 *quads; };      // cells = entries per plane

__global__ void __launch_bounds__(256) k_pack_quads(const __grid_constant__ QuadArgs qa) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= qa.cells) return;
    uint32_t l = 0;
#pragma unroll
    for (uint32_t k = 1; k < 12; k++) l += (c >= qa.qlvl[k].off) ? 1u : 0u;
    const HeadLvl g = qa.lvl[l], q = qa.qlvl[l];
    const uint32_t r = c - q.off, j = r / q.mul, i = r - j * q.mul;
    const uint32_t m0 = j * g.mul, m1 = m0 + g.mul;
    uint32_t e0, e1, e2, e3;
    if (g.mask == 0xffffffffu) { e0 = i + m0; e1 = i + 1u + m0; e2 = i + m1; e3 = i + 1u + m1; }
    else { e0 = (i ^ m0) & g.mask; e1 = ((i + 1u) ^ m0) & g.mask; e2 = (i ^ m1) & g.mask; e3 = ((i + 1u) ^ m1) & g.mask; }
    const float *b = qa.tab[blockIdx.y] + g.off;
    qa.quads[(size_t)blockIdx.y * qa.cells + c] = make_float4(__ldg(b + e0), __ldg(b + e1), __ldg(b + e2), __ldg(b + e3));
}

}  // namespace b2n

// ---------------------------------------------------------------------------------------------------
// C ABI: model object (packed weights) + head forward
// ---------------------------------------------------------------------------------------------------
using namespace b2n;

namespace b2n { __global__ void k_level_scales(float S, uint32_t H, uint32_t L, float *__restrict__ out); }

// Derive the 12 per-level constants.  Reads offsets[13] and the device-computed scales back to the host (synchronises `st`);
// runs only when the table geometry changes (model creation), never per frame / per training step.
static int derive_levels(b2n_model *m, const b2n_head_weights *w, cudaStream_t st) {
    float *d_scales = m->wsmall + HS_FLOATS;
    k_level_scales<<<1, 64, 0, st>>>(w->S, w->H, 12, d_scales);
    if (check_launch("model_update(scales)")) return 1;
    int32_t offs[13];
    float scales[12];
    uint32_t qcells = 0;
    B2N_CUDA(cudaMemcpyAsync(offs, w->offsets, sizeof(offs), cudaMemcpyDeviceToHost, st));
    B2N_CUDA(cudaMemcpyAsync(scales, d_scales, sizeof(scales), cudaMemcpyDeviceToHost, st));
    B2N_CUDA(cudaStreamSynchronize(st));
    for (int l = 0; l < 12; l++) {
        const uint32_t size = (uint32_t)(offs[l + 1] - offs[l]);
        const uint32_t res = (uint32_t)ceilf(scales[l]) + 1u, stride = res + 1u;
        B2N_REQUIRE(offs[l] >= 0 && offs[l + 1] > offs[l], "model_update: offsets must be increasing");
        const bool dense = (uint64_t)stride * stride <= size;              // gridencoder.cu:54-72 with D = 2, gridtype = hash
        HeadLvl g;
        g.scale = scales[l];
        g.off = (uint32_t)offs[l];
        if (dense) { g.mul = stride; g.mask = 0xffffffffu; }
        else {
            B2N_REQUIRE((size & (size - 1)) == 0, "model_update: hashed level %d has %u entries; the fused kernel needs a power of two (use the per-op path)", l, size);
            g.mul = 2654435761u; g.mask = size - 1u;
        }
        m->lvl[l] = g;
        // quad image: cells (i, j), i, j in [0, res): floor(u * scale + 0.5) <= res - 1 for u in [0, 1]
        m->qlvl[l] = HeadLvl{scales[l], res, 0u, qcells};
        qcells += res * res;
    }
    if (m->use_quads && qcells != m->quad_cells) {
        if (m->quads) cudaFree(m->quads);
        m->quads = nullptr;
        B2N_CUDA(cudaMalloc(&m->quads, (size_t)qcells * 3 * sizeof(float4)));
    }
    m->quad_cells = qcells;
    m->geo_offsets = w->offsets; m->geo_S = w->S; m->geo_H = w->H;
    return 0;
}

extern "C" {

int b2n_model_create(b2n_model **out, void *stream) {
    (void)stream;
    B2N_REQUIRE(out, "model_create: null pointer");
    b2n_model *m = new b2n_model();
    {   // B2N_HEAD_QUADS=0 keeps the four-reads-per-cell gather on the reference-format tables (A/B measurements)
        const char *e = getenv("B2N_HEAD_QUADS");
        m->use_quads = !(e && e[0] == '0');
    }
    if (cudaMalloc(&m->wimg, HW_BYTES) != cudaSuccess || cudaMalloc(&m->wimg_t, HT_BYTES) != cudaSuccess ||
        cudaMalloc(&m->wsmall, sizeof(float) * (HS_FLOATS + 16)) != cudaSuccess) {
        (void)cudaGetLastError();
        if (m->wimg) cudaFree(m->wimg);
        if (m->wimg_t) cudaFree(m->wimg_t);
        delete m;
        set_error("model_create: device allocation failed");
        return 3;
    }
    *out = m;
    return 0;
}

void b2n_model_destroy(b2n_model *m) {
    if (!m) return;
    cudaFree(m->wimg);
    cudaFree(m->wimg_t);
    cudaFree(m->wsmall);
    if (m->quads) cudaFree(m->quads);
    delete m;
}

int b2n_model_update(b2n_model *m, const b2n_head_weights *w, void *stream) {
    B2N_REQUIRE(m && w, "model_update: null pointer");
    B2N_REQUIRE(w->table_xy && w->table_yz && w->table_xz && w->offsets, "model_update: null table pointer");
    B2N_REQUIRE(w->aud_att_w0 && w->aud_att_w1 && w->eye_att_w0 && w->eye_att_w1 && w->sigma_w0 && w->sigma_w1 && w->sigma_w2 && w->color_w0 && w->color_w1,
                "model_update: null weight pointer");
    B2N_REQUIRE((w->unc_w0 == nullptr) == (w->unc_w1 == nullptr), "model_update: unc_w0 / unc_w1 must both be given or both be NULL");
    PackArgs pa = {};
    uint32_t n = 0;
    auto add = [&](const float *src, uint32_t off, uint32_t rows, uint32_t src_rows, uint32_t ld, uint32_t col0, uint32_t kvalid, uint32_t shift) {
        pa.r[n++] = PackRegion{src, off, rows, src_rows, ld, col0, kvalid, shift, 0, 0, 0, 0, 0};
    };
    add(w->aud_att_w0, HW_A, 64, 64, 36, 0, 36, 0);
    add(w->eye_att_w0, HW_A + 64 * 128, 16, 16, 36, 0, 36, 0);
    add(w->sigma_w0, HW_A + 80 * 128, 64, 64, 69, 0, 36, 0);
    add(w->unc_w0, HW_U, 32, 32, 36, 0, 36, 0);                            // NULL src -> zeros
    add(w->aud_att_w1, HW_B, 32, 32, 64, 0, 64, 0);
    add(w->sigma_w0, HW_C, 64, 64, 69, 36, 33, 0);
    add(w->sigma_w1, HW_D, 64, 64, 64, 0, 64, 0);
    add(w->sigma_w2, HW_E, 80, 65, 64, 0, 64, 1);                          // rotate: rows 0..63 = geo (src 1..64), row 64 = logit (src 0)
    add(w->color_w0, HW_F0, 64, 64, 84, 16, 64, 0);                        // geo_feat columns 16..79
    add(w->color_w0, HW_F1, 64, 64, 84, 0, 16, 0);                         // SH columns 0..15 at K 0..15 ...
    pa.r[n - 1].col1 = 80; pa.r[n - 1].k1_off = 16; pa.r[n - 1].k1_valid = 4;   // ... and the individual-code columns 80..83 at K 16..19 (second K step of P6)
    add(w->color_w1, HW_G, 16, 3, 64, 0, 64, 0);
    pa.n = n;
    cudaStream_t st = as_stream(stream);
    if (m->geo_offsets != w->offsets || m->geo_S != w->S || m->geo_H != w->H)
        if (int rc = derive_levels(m, w, st)) return rc;
    k_pack_head<<<ceil_div<uint32_t>(HW_BYTES / 16, 256), 256, 0, st>>>(pa, m->wimg);
    if (check_launch("model_update(pack)")) return 1;
    k_pack_small<<<1, 256, 0, st>>>(w->eye_att_w1, w->unc_w1, w->color_w0, w->color_w1, m->wsmall);
    if (check_launch("model_update(small)")) return 1;
    {   // transposed image for the backward-data kernel
        PackArgs pt = {};
        uint32_t k = 0;
        auto addt = [&](const float *src, uint32_t off, uint32_t rows, uint32_t nvalid, uint32_t ld, uint32_t col0, uint32_t kvalid, uint32_t k0) {
            pt.r[k++] = PackRegion{src, off, rows, nvalid, ld, col0, kvalid, 0, 1, k0, 0, 0, 0};
        };
        addt(w->color_w0, HT_C0G, 64, 64, 84, 16, 64, 0);
        addt(w->color_w0, HT_C0I, 16, 4, 84, 80, 64, 0);
        addt(w->sigma_w2, HT_S2A, 64, 64, 64, 0, 64, 1);
        addt(w->sigma_w2, HT_S2B, 64, 64, 64, 0, 1, 0);
        addt(w->sigma_w1, HT_S1, 64, 64, 64, 0, 64, 0);
        addt(w->sigma_w0, HT_S0X, 48, 36, 69, 0, 64, 0);
        addt(w->sigma_w0, HT_S0W, 48, 33, 69, 36, 64, 0);
        addt(w->aud_att_w1, HT_A1, 64, 64, 64, 0, 32, 0);
        addt(w->aud_att_w0, HT_A0, 48, 36, 36, 0, 64, 0);
        addt(w->eye_att_w0, HT_E0, 48, 36, 36, 0, 16, 0);
        pt.n = k;
        k_pack_head<<<ceil_div<uint32_t>(HT_BYTES / 16, 256), 256, 0, st>>>(pt, m->wimg_t);
        if (check_launch("model_update(pack transposed)")) return 1;
    }
    if (m->use_quads) {
        QuadArgs qa = {};
        qa.tab[0] = w->table_xy; qa.tab[1] = w->table_yz; qa.tab[2] = w->table_xz;
        for (int l = 0; l < 12; l++) { qa.lvl[l] = m->lvl[l]; qa.qlvl[l] = m->qlvl[l]; }
        qa.cells = m->quad_cells; qa.quads = m->quads;
        k_pack_quads<<<dim3(ceil_div<uint32_t>(m->quad_cells, 256), 3), 256, 0, st>>>(qa);
        if (check_launch("model_update(quads)")) return 1;
    }
    m->w = *w;
    m->ready = true;
    return 0;
}

}  // extern "C"

namespace b2n {
int head_forward_on_model(const b2n_model *m, const float *xyzs, const float *dirs, uint32_t M, const float *enc_a, const float *ind_code, const float *eye,
                          const int32_t *n_valid, float density_scale, float *sigmas, float *rgbs, float *amb_aud, float *amb_eye, float *unc, cudaStream_t st,
                          const float *live_deltas, const b2n_head_saved *saved, uint32_t head_ctas) {
    B2N_REQUIRE(m && m->ready, "head_forward: model has no weights (call b2n_model_update)");
    B2N_REQUIRE(xyzs && dirs && enc_a, "head_forward: null pointer");
    if (M == 0) return 0;
    HeadArgs a = {};
    a.xyzs = xyzs; a.dirs = dirs; a.M = M; a.n_valid = n_valid;
    if (m->use_quads) {
        for (int p = 0; p < 3; p++) a.tab[p] = reinterpret_cast<const float *>(m->quads + (size_t)p * m->quad_cells);
        for (int l = 0; l < 12; l++) a.lvl[l] = m->qlvl[l];
    } else {
        a.tab[0] = m->w.table_xy; a.tab[1] = m->w.table_yz; a.tab[2] = m->w.table_xz;
        for (int l = 0; l < 12; l++) a.lvl[l] = m->lvl[l];
    }
    a.bound = m->w.bound;
    {   // (x + b) / (2b) == (x + b) * (1 / 2b) exactly when 2b is a power of two
        int ex = 0;
        const float two_b = 2.0f * a.bound;
        a.inv_two_bound = (two_b > 0.0f && frexpf(two_b, &ex) == 0.5f) ? 1.0f / two_b : 0.0f;
    }
    a.wimg = m->wimg; a.wsmall = m->wsmall;
    a.enc_a = enc_a; a.ind_code = ind_code; a.eye = eye;
    a.sigmas = sigmas; a.rgbs = rgbs; a.amb_aud = amb_aud; a.amb_eye = amb_eye; a.unc = unc;
    a.has_unc = m->w.unc_w0 != nullptr;
    a.density_scale = density_scale;
    // a frame that shares the GPU with other frames in flight (head_ctas > 0): at most head_ctas CTAs, grid-strided tiles; a frame alone: all SMs, balanced shares
    a.sched = (head_ctas || saved) ? 2 : 0;      // training: the grid-strided walk measured 2 % faster per step (neighbouring tiles run at the same time on different SMs and share the L2)
    a.max_ctas = head_ctas;
    if (const char *e = getenv("B2N_HEAD_SCHED")) a.sched = atoi(e);            // A/B overrides
    if (const char *e = getenv("B2N_HEAD_CTAS")) a.max_ctas = (uint32_t)atoi(e);
    if (saved) {
        B2N_REQUIRE(!a.has_unc || saved->hu, "head_forward_train: unc_net is packed but saved->hu is NULL");
        a.sv = *saved;
    }
    return launch_head_forward(a, st, saved != nullptr, m->use_quads);
}
}  // namespace b2n

extern "C" {

int b2n_head_forward_train(const b2n_model *m, const float *xyzs, const float *dirs, uint32_t M, const float *enc_a, const float *ind_code, const float *eye,
                           float *sigmas, float *rgbs, float *amb_aud, float *amb_eye, float *unc, const b2n_head_saved *saved, void *stream) {
    B2N_REQUIRE(saved, "head_forward_train: null pointer");
    B2N_REQUIRE(saved->x36 && saved->ha && saved->he && saved->att && saved->s_in && saved->h1 && saved->h2 && saved->c_in && saved->hc && saved->misc,
                "head_forward_train: null activation buffer");
    return head_forward_on_model(m, xyzs, dirs, M, enc_a, ind_code, eye, nullptr, 1.0f, sigmas, rgbs, amb_aud, amb_eye, unc, as_stream(stream), nullptr, saved, 0);
}

int b2n_head_forward(const b2n_model *m, const float *xyzs, const float *dirs, uint32_t M, const float *enc_a, const float *ind_code, const float *eye,
                     const int32_t *n_valid, float *sigmas, float *rgbs, float *amb_aud, float *amb_eye, float *unc, void *stream) {
    return head_forward_on_model(m, xyzs, dirs, M, enc_a, ind_code, eye, n_valid, 1.0f, sigmas, rgbs, amb_aud, amb_eye, unc, as_stream(stream), nullptr, nullptr, 0);
}

}  // extern "C"
