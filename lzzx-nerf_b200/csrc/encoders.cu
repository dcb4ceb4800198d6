// encoders.cu — spherical-harmonics and frequency encoders for sm_100a.
//
// SH: the reference (shencoder/src/shencoder.cu:28-355) spells out 64 polynomials and their 192 partials.
// They are exactly  Y[l*l+l+m] = (-1)^m K(l,|m|) sqrt2[m!=0] * Pi_l^|m|(z) * (m>=0 ? Re : Im)((x+iy)^|m|)
// with Pi_l^m(z) = d^m/dz^m P_l(z) evaluated as a polynomial in z (inputs are NOT re-normalised), so this file
// evaluates them by recurrence (templated on the degree, fully unrolled) and derives the partials analytically:
//   d/dx Re_m = m Re_{m-1},  d/dy Re_m = -m Im_{m-1},  d/dx Im_m = m Im_{m-1},  d/dy Im_m = m Re_{m-1},
//   d/dz Pi_l^m = Pi_l^{m+1}.
// A CTA computes a [128 x degree^2] tile in shared memory and writes it with coalesced 128 B transactions
// (the reference's per-thread stride-64 B stores touch 32 sectors per instruction).
//
// Freq: freqencoder/src/freqencoder.cu:30-94, same __sinf(scalbnf(x,f)+phase) evaluation, one thread per output.
#include "common.cuh"
#include <math.h>

namespace b2n {

// K[l][m] = (-1)^m sqrt2[m!=0] sqrt((2l+1)/(4pi) (l-m)!/(l+m)!);  dfact[m] = (2m-1)!! = Pi_m^m
// passed by value as a kernel parameter (constant bank; no device-side state, graph-capture safe)
struct ShConsts { float K[8][8]; float dfact[9]; };

static const ShConsts &sh_consts() {
    static ShConsts c = [] {
        ShConsts t = {};
        for (int m = 0; m <= 8; m++) { double d = 1; for (int k = 1; k <= m; k++) d *= 2.0 * k - 1.0; t.dfact[m] = (float)d; }
        for (int l = 0; l < 8; l++)
            for (int m = 0; m <= l; m++) {
                double k = (2.0 * l + 1.0) / (4.0 * M_PI);
                for (int j = l - m + 1; j <= l + m; j++) k /= (double)j;
                t.K[l][m] = (float)(sqrt(k) * (m ? M_SQRT2 : 1.0) * ((m & 1) ? -1.0 : 1.0));
            }
        return t;
    }();
    return c;
}

constexpr int SH_ROWS = 128;

template <int DEG, bool GRAD>
__global__ void __launch_bounds__(SH_ROWS) k_sh_fwd(const __grid_constant__ ShConsts sc, const float *__restrict__ inputs, float *__restrict__ outputs,
                                                     uint32_t B, uint32_t D, float *__restrict__ dy_dx) {
    constexpr int C2 = DEG * DEG;
    constexpr int LD = C2 + 1;                       // odd row stride: conflict-free column writes
    extern __shared__ float sm[];                    // [SH_ROWS][LD] values (+ 3 x the same for partials)
    float *sY = sm, *sDx = sm + SH_ROWS * LD, *sDy = sDx + SH_ROWS * LD, *sDz = sDy + SH_ROWS * LD;
    const uint32_t row0 = blockIdx.x * SH_ROWS, b = row0 + threadIdx.x;
    if (b < B) {
        const float x = inputs[(size_t)b * D], y = inputs[(size_t)b * D + 1], z = inputs[(size_t)b * D + 2];
        float A[DEG], Bm[DEG];                       // (x+iy)^m
        A[0] = 1.0f; Bm[0] = 0.0f;
#pragma unroll
        for (int m = 1; m < DEG; m++) { A[m] = x * A[m - 1] - y * Bm[m - 1]; Bm[m] = x * Bm[m - 1] + y * A[m - 1]; }
        float Pi[DEG + 1][DEG];                      // Pi[m][l], zero for l < m; row DEG is the all-zero derivative seed
#pragma unroll
        for (int m = 0; m <= DEG; m++)
#pragma unroll
            for (int l = 0; l < DEG; l++) Pi[m][l] = 0.0f;
#pragma unroll
        for (int m = 0; m < DEG; m++) {
            Pi[m][m] = sc.dfact[m];
            if (m + 1 < DEG) Pi[m][m + 1] = (2.0f * m + 1.0f) * z * sc.dfact[m];
#pragma unroll
            for (int l = m + 2; l < DEG; l++) Pi[m][l] = ((2.0f * l - 1.0f) * z * Pi[m][l - 1] - (float)(l + m - 1) * Pi[m][l - 2]) * (1.0f / (float)(l - m));
        }
        float *rY = sY + threadIdx.x * LD, *rX = sDx + threadIdx.x * LD, *rYd = sDy + threadIdx.x * LD, *rZ = sDz + threadIdx.x * LD;
#pragma unroll
        for (int l = 0; l < DEG; l++) {
#pragma unroll
            for (int mm = -l; mm <= l; mm++) {
                const int m = mm < 0 ? -mm : mm;
                const float K = sc.K[l][m];
                const float P = Pi[m][l];
                const float Q = (mm >= 0) ? A[m] : Bm[m];
                const int i = l * l + l + mm;
                rY[i] = K * P * Q;
                if (GRAD) {
                    float dQx = 0.0f, dQy = 0.0f;
                    if (m > 0) {
                        if (mm >= 0) { dQx = (float)m * A[m - 1]; dQy = -(float)m * Bm[m - 1]; }
                        else         { dQx = (float)m * Bm[m - 1]; dQy = (float)m * A[m - 1]; }
                    }
                    rX[i] = K * P * dQx; rYd[i] = K * P * dQy; rZ[i] = K * Pi[m + 1][l] * Q;
                }
            }
        }
    }
    __syncthreads();
    const uint32_t rows = min((uint32_t)SH_ROWS, B - row0);
    for (uint32_t i = threadIdx.x; i < rows * C2; i += SH_ROWS) {
        const uint32_t r = i / C2, c = i - r * C2;
        __stcs(outputs + (size_t)row0 * C2 + i, sY[r * LD + c]);
    }
    if (GRAD) {                                      // dy_dx [B, 3, C2]  (shencoder.cu:126-129)
        for (uint32_t i = threadIdx.x; i < rows * 3 * C2; i += SH_ROWS) {
            const uint32_t r = i / (3 * C2), k = i - r * 3 * C2, d = k / C2, c = k - d * C2;
            const float *src = d == 0 ? sDx : (d == 1 ? sDy : sDz);
            __stcs(dy_dx + (size_t)row0 * 3 * C2 + i, src[r * LD + c]);
        }
    }
}

// shencoder.cu:359-382 — grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]
__global__ void __launch_bounds__(256) k_sh_bwd(const float *__restrict__ grad, uint32_t B, uint32_t D, uint32_t C2,
                                                 const float *__restrict__ dy_dx, float *__restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    float r = grad_inputs[t];
    for (uint32_t ch = 0; ch < C2; ch++) r = __fmaf_rn(grad[(size_t)b * C2 + ch], dy_dx[((size_t)b * D + d) * C2 + ch], r);
    grad_inputs[t] = r;
}

template <int DEG>
static int sh_launch(const float *inputs, float *outputs, uint32_t B, uint32_t D, float *dy_dx, cudaStream_t st) {
    constexpr int LD = DEG * DEG + 1;
    const uint32_t ctas = ceil_div<uint32_t>(B, SH_ROWS);
    if (dy_dx) {
        const size_t smem = sizeof(float) * 4 * SH_ROWS * LD;
        B2N_SMEM((k_sh_fwd<DEG, true>), smem);
        k_sh_fwd<DEG, true><<<ctas, SH_ROWS, smem, st>>>(sh_consts(), inputs, outputs, B, D, dy_dx);
    } else {
        k_sh_fwd<DEG, false><<<ctas, SH_ROWS, sizeof(float) * SH_ROWS * LD, st>>>(sh_consts(), inputs, outputs, B, D, nullptr);
    }
    return check_launch("sh_encode_forward");
}

// freqencoder.cu:30-58
__global__ void __launch_bounds__(256) k_freq_fwd(const float *__restrict__ inputs, uint32_t B, uint32_t D, uint32_t C, float *__restrict__ outputs) {
    const float HALF_PI = 3.141592653589793f / 2;
    const size_t total = (size_t)B * C;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(t / C), c = (uint32_t)(t - (size_t)b * C);
        float v;
        if (c < D) v = inputs[(size_t)b * D + c];
        else {
            const uint32_t col = c / D - 1, d = c % D, f = col / 2;
            v = __sinf(__fadd_rn(scalbnf(inputs[(size_t)b * D + d], (int)f), (float)(col % 2) * HALF_PI));
        }
        __stcs(outputs + t, v);
    }
}
// freqencoder.cu:63-94
__global__ void __launch_bounds__(256) k_freq_bwd(const float *__restrict__ grad, const float *__restrict__ outputs, uint32_t B, uint32_t D, uint32_t deg,
                                                   uint32_t C, float *__restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float *g = grad + (size_t)b * C, *o = outputs + (size_t)b * C;
    float r = g[d];
    g += D; o += D;
    for (uint32_t f = 0; f < deg; f++) {
        r = __fmaf_rn(scalbnf(1.0f, (int)f), __fmaf_rn(g[d], o[D + d], -__fmul_rn(g[D + d], o[d])), r);
        g += 2 * D; o += 2 * D;
    }
    grad_inputs[t] = r;
}

}  // namespace b2n

using namespace b2n;

extern "C" {

int b2n_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t degree, float *dy_dx, void *stream) {
    B2N_REQUIRE(inputs && outputs, "sh_encode_forward: null pointer");
    B2N_REQUIRE(D == 3, "SH encoder only support input dim == 3");
    B2N_REQUIRE(degree >= 1 && degree <= 8, "SH encoder only supports degree in [1, 8]");
    if (B == 0) return 0;
    cudaStream_t st = as_stream(stream);
    switch (degree) {
        case 1: return sh_launch<1>(inputs, outputs, B, D, dy_dx, st);
        case 2: return sh_launch<2>(inputs, outputs, B, D, dy_dx, st);
        case 3: return sh_launch<3>(inputs, outputs, B, D, dy_dx, st);
        case 4: return sh_launch<4>(inputs, outputs, B, D, dy_dx, st);
        case 5: return sh_launch<5>(inputs, outputs, B, D, dy_dx, st);
        case 6: return sh_launch<6>(inputs, outputs, B, D, dy_dx, st);
        case 7: return sh_launch<7>(inputs, outputs, B, D, dy_dx, st);
        default: return sh_launch<8>(inputs, outputs, B, D, dy_dx, st);
    }
}

int b2n_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t degree, const float *dy_dx, float *grad_inputs, void *stream) {
    (void)inputs;
    B2N_REQUIRE(grad && dy_dx && grad_inputs, "sh_encode_backward: null pointer");
    B2N_REQUIRE(degree >= 1 && degree <= 8, "SH encoder only supports degree in [1, 8]");
    if (B == 0) return 0;
    k_sh_bwd<<<ceil_div<uint32_t>(B * D, 256), 256, 0, as_stream(stream)>>>(grad, B, D, degree * degree, dy_dx, grad_inputs);
    return check_launch("sh_encode_backward");
}

int b2n_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *outputs, void *stream) {
    B2N_REQUIRE(inputs && outputs, "freq_encode_forward: null pointer");
    B2N_REQUIRE(D >= 1 && C == D + 2 * D * deg, "freq_encode_forward: output_dim %u != input_dim %u * (1 + 2*degree %u)", C, D, deg);
    if (B == 0) return 0;
    const size_t total = (size_t)B * C;
    const size_t want = ceil_div<size_t>(total, 256), cap = (size_t)sm_count() * 16;
    k_freq_fwd<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(inputs, B, D, C, outputs);
    return check_launch("freq_encode_forward");
}

int b2n_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *grad_inputs, void *stream) {
    B2N_REQUIRE(grad && outputs && grad_inputs, "freq_encode_backward: null pointer");
    B2N_REQUIRE(D >= 1 && C == D + 2 * D * deg, "freq_encode_backward: output_dim %u != input_dim %u * (1 + 2*degree %u)", C, D, deg);
    if (B == 0) return 0;
    k_freq_bwd<<<ceil_div<uint32_t>(B * D, 256), 256, 0, as_stream(stream)>>>(grad, outputs, B, D, deg, C, grad_inputs);
    return check_launch("freq_encode_backward");
}

}  // extern "C"
