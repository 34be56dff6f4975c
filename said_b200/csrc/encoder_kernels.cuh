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

// Per-utterance zero-mean / unit-variance normalisation on the device: y = (x - mean) / sqrt(var + 1e-7), population
// variance (HF Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm, reached from said/model/diffusion.py:188-207; the
// reference does this in numpy on the host and returns a CPU tensor).  One CTA per clip, two passes in fp64 (the second
// pass and the write re-read the clip from L2).
__global__ void __launch_bounds__(1024)
normalize_audio_kernel(const float* __restrict__ x, int T_a, float* __restrict__ y) {
    __shared__ double red[32];
    __shared__ double s_mean, s_rstd;
    const float* xb = x + (long long)blockIdx.x * T_a;
    float* yb = y + (long long)blockIdx.x * T_a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s = 0.0;
    for (int i = tid; i < T_a; i += blockDim.x) s += (double)__ldg(xb + i);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        s_mean = t / T_a;
    }
    __syncthreads();
    const double mean = s_mean;
    double q = 0.0;
    for (int i = tid; i < T_a; i += blockDim.x) { const double d = (double)__ldg(xb + i) - mean; q += d * d; }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        s_rstd = 1.0 / sqrt(t / T_a + 1e-7);
    }
    __syncthreads();
    const float m32 = (float)mean, r32 = (float)s_rstd;
    for (int i = tid; i < T_a; i += blockDim.x) yb[i] = (__ldg(xb + i) - m32) * r32;
}

// feat_extract_norm == "layer" (wav2vec2-large family, TF modeling_wav2vec2.py:275-299): conv0 = Conv1d(1 -> 512, k=10,
// s=5, bias) -> LayerNorm over the 512 channels of each frame -> GELU.  One warp per frame (16 channels per lane, the
// frame's 10 input samples broadcast from one load), output channel-last at row (b * out_stride + f).
__global__ void __launch_bounds__(256)
conv0_layernorm_gelu_kernel(const float* __restrict__ wave, int T_a, int L0, int B, const float* __restrict__ w /*(10,512)*/,
                            const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                            float eps, float* __restrict__ out, int out_stride) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * L0) return;
    const int b = (int)(gw / L0), f = (int)(gw - (long long)b * L0);
    const float* x = wave + (long long)b * T_a + (long long)f * C0_S;
    float xs[C0_K];
#pragma unroll
    for (int k = 0; k < C0_K; ++k) xs[k] = __ldg(x + k);
    float y[16];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int c = lane + 32 * j;
        float a = __ldg(bias + c);
#pragma unroll
        for (int k = 0; k < C0_K; ++k) a = fmaf(__ldg(w + k * C0_CH + c), xs[k], a);
        y[j] = a;
        s += a;
    }
    const float mean = warp_sum(s) * (1.0f / C0_CH);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float d = y[j] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C0_CH) + eps);
    float* o = out + ((long long)b * out_stride + f) * C0_CH;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int c = lane + 32 * j;
        o[c] = gelu_erf((y[j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
    }
}

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

// Device-side load_audio tail (reference said/util/audio.py:35-38): torchaudio.functional.resample (polyphase windowed-sinc FIR:
// output j = i * new + p uses filter bank row p on the input window starting at i * orig - width) applied per channel, then the
// mean over channels.  wave: (channels, n_in); bank: (new, K = 2 width + orig), built by the host exactly as torchaudio builds it;
// out: (n_out).  One thread per output sample.
__global__ void __launch_bounds__(256)
resample_mono_kernel(const float* __restrict__ wave, int channels, int n_in, int orig, int nw, int width, const float* __restrict__ bank,
                     float* __restrict__ out, int n_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const int i = j / nw, p = j - i * nw;
    const int K = 2 * width + orig;
    const float* f = bank + (long long)p * K;
    const int x0 = i * orig - width;
    float mix = 0.f;
    for (int c = 0; c < channels; ++c) {
        const float* x = wave + (long long)c * n_in;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            const int xi = x0 + k;
            if (xi >= 0 && xi < n_in) acc = fmaf(__ldg(f + k), __ldg(x + xi), acc);
        }
        mix += acc;
    }
    out[j] = mix / (float)channels;
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
