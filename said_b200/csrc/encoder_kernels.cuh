// Wav2Vec2 front-end kernels that are not GEMM-shaped (TF modeling_wav2vec2.py:302-323, 326-368).
#pragma once
#include "common.cuh"

namespace said {

// conv0 = Conv1d(1 -> 512, k=10, s=5, no bias) followed by GroupNorm(512 groups) == per-channel
// normalisation over all frames of the clip, then GELU.  The (L0, 512) output is the largest activation
// of the whole path (32.8 MB per 5 s clip), so the raw conv output is never stored: pass 1 recomputes
// the 10-tap conv to accumulate per-channel sum / sum-of-squares (fp64), pass 2 recomputes it again and
// writes gelu(norm(conv)) once, channel-last and coalesced.  Thread == channel.
constexpr int C0_CH = 512;
constexpr int C0_K = 10;
constexpr int C0_S = 5;
constexpr int C0_TILE = 128;   // frames per smem tile

// grid (nchunk, B); partial: (B, nchunk, 2, 512) doubles
__global__ void __launch_bounds__(C0_CH)
conv0_stats_kernel(const float* __restrict__ wave, int T_a, int L0, const float* __restrict__ w /*(10,512)*/,
                   int frames_per_chunk, double* __restrict__ partial) {
    __shared__ float xs[C0_TILE * C0_S + C0_K];
    const int c = threadIdx.x, chunk = blockIdx.x, b = blockIdx.y;
    float wk[C0_K];
#pragma unroll
    for (int k = 0; k < C0_K; ++k) wk[k] = __ldg(w + k * C0_CH + c);
    const float* x = wave + (long long)b * T_a;
    const int f_begin = chunk * frames_per_chunk;
    const int f_end = min(L0, f_begin + frames_per_chunk);
    double s = 0.0, q = 0.0;
    for (int f0 = f_begin; f0 < f_end; f0 += C0_TILE) {
        const int nf = min(C0_TILE, f_end - f0);
        const int ns = (nf - 1) * C0_S + C0_K;
        __syncthreads();
        for (int i = threadIdx.x; i < ns; i += C0_CH) xs[i] = __ldg(x + (long long)f0 * C0_S + i);
        __syncthreads();
        float ts = 0.f, tq = 0.f;      // fp32 inside a tile of <= 128 frames, fp64 across tiles
        for (int f = 0; f < nf; ++f) {
            float y = 0.f;
#pragma unroll
            for (int k = 0; k < C0_K; ++k) y = fmaf(wk[k], xs[f * C0_S + k], y);
            ts += y;
            tq = fmaf(y, y, tq);
        }
        s += ts;
        q += tq;
    }
    double* pp = partial + ((long long)b * gridDim.x + chunk) * 2 * C0_CH;
    pp[c] = s;
    pp[C0_CH + c] = q;
}

// grid (ceil(L0 / C0_TILE), B); out: (B, out_stride frames, 512), frames [0, L0) of each clip written
__global__ void __launch_bounds__(C0_CH)
conv0_apply_kernel(const float* __restrict__ wave, int T_a, int L0, const float* __restrict__ w,
                   const double* __restrict__ partial, int nchunk, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, float* __restrict__ out, int out_stride) {
    __shared__ float xs[C0_TILE * C0_S + C0_K];
    const int c = threadIdx.x, b = blockIdx.y;
    const int f0 = blockIdx.x * C0_TILE;
    const int nf = min(C0_TILE, L0 - f0);
    const int ns = (nf - 1) * C0_S + C0_K;
    const float* x = wave + (long long)b * T_a;
    for (int i = threadIdx.x; i < ns; i += C0_CH) xs[i] = __ldg(x + (long long)f0 * C0_S + i);
    double s = 0.0, q = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) {
        const double* pp = partial + ((long long)b * nchunk + ch) * 2 * C0_CH;
        s += pp[c];
        q += pp[C0_CH + c];
    }
    const double mean = s / L0;
    double var = q / L0 - mean * mean;
    if (var < 0.0) var = 0.0;
    const float sc = (float)(1.0 / sqrt(var + (double)eps)) * __ldg(gamma + c);
    const float sh = __ldg(beta + c) - (float)mean * sc;
    float wk[C0_K];
#pragma unroll
    for (int k = 0; k < C0_K; ++k) wk[k] = __ldg(w + k * C0_CH + c);
    __syncthreads();
    float* o = out + ((long long)b * out_stride + f0) * C0_CH + c;
    for (int f = 0; f < nf; ++f) {
        float y = 0.f;
#pragma unroll
        for (int k = 0; k < C0_K; ++k) y = fmaf(wk[k], xs[f * C0_S + k], y);
        o[(long long)f * C0_CH] = gelu_erf(y * sc + sh);
    }
}

// Positional conv embedding input: (B, T, H) -> zero-padded, group-major (B, G, T + KP, H/G) so that
// the grouped Conv1d(H, H, k=KP, pad=KP/2, groups=G) becomes, per (clip, group), a GEMM whose row t is
// the contiguous run of KP*(H/G) floats starting at padded frame t (row stride = H/G floats).
__global__ void posconv_regroup_kernel(const float* __restrict__ x, int B, int T, int H, int G, int KP,
                                       float* __restrict__ xp) {
    const int cg = H / G, Tp = T + KP;
    const long long total = (long long)B * G * Tp * cg;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cg);
    long long r = i / cg;
    const int tp = (int)(r % Tp); r /= Tp;
    const int g = (int)(r % G);
    const int b = (int)(r / G);
    const int t = tp - KP / 2;
    xp[i] = (t >= 0 && t < T) ? __ldg(x + ((long long)b * T + t) * H + g * cg + c) : 0.f;
}

}  // namespace said
