// Normalisation statistics and row-wise normalisation kernels.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "pair.cuh"

namespace said {

// GroupNorm statistics -> per-(sample, channel) scale/shift so that  gn(x)[c] = x[c]*scale + shift.
// (GroupNorm32, ldm/util.py:111-122: 32 groups, biased variance over channels-in-group x T, fp32;
//  attention.py:63-66 Normalize: same with eps 1e-6.)  x is (B', T, C) channel-last, C == 192; `cpg`
// channels per group (6 for a 192-channel tensor, 12 for one half of a 384-channel concat).
// Two launches, GN_SPLIT CTAs per sample each (a sample is only 230 KB, but one CTA per sample leaves the
// chip latency-bound and, for a single clip, idle):
//   gn_partial_kernel  per-channel sum / sum-of-squares of a quarter of the frames, fp64 (no cancellation
//                      problem in E[x^2]-mean^2), deterministic (no atomics)
//   gn_finish_kernel   reduces the partials, writes scale/shift at [b*out_ld + out_off + c] and, on the
//                      tensor-core path, materialises silu(gn(x)) for its quarter of the frames so the conv
//                      GEMM's operand loader is a plain shifted copy (the data was just read: L2 hits)
constexpr int GN_THREADS = 768;   // 4 row phases x 192 channels
constexpr int GN_SPLIT = 4;        // CTAs per sample at batch scale
constexpr int GN_SPLIT_MAX = 16;   // ... for a handful of samples (single-clip latency)
__global__ void __launch_bounds__(GN_THREADS)
gn_partial_kernel(const float* __restrict__ x, int src_samples, int T, double* __restrict__ partial /*(B', GN_SPLIT, 2, 192)*/) {
    constexpr int C = 192;
    __shared__ double s_sum[4][C];
    __shared__ double s_sq[4][C];
    pdl_wait();
    pdl_trigger();
    const int sp = blockIdx.x, b = blockIdx.y;
    const int c = threadIdx.x % C, ph = threadIdx.x / C;
    const int nsp = gridDim.x;
    const int rows = (T + nsp - 1) / nsp;
    const int t0 = sp * rows, t1 = min(T, t0 + rows);
    const float* xb = x + (long long)(b % src_samples) * T * C;   // sample b reads source sample b % src_samples (shared CFG prefix)
    double s = 0.0, q = 0.0;
    int t = t0 + ph;
    for (; t + 12 < t1; t += 16) {   // 4 independent loads in flight
        const float v0 = __ldg(xb + (long long)t * C + c), v1 = __ldg(xb + (long long)(t + 4) * C + c);
        const float v2 = __ldg(xb + (long long)(t + 8) * C + c), v3 = __ldg(xb + (long long)(t + 12) * C + c);
        s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
        q += ((double)v0 * v0 + (double)v1 * v1) + ((double)v2 * v2 + (double)v3 * v3);
    }
    for (; t < t1; t += 4) {
        const float v = __ldg(xb + (long long)t * C + c);
        s += v;
        q += (double)v * v;
    }
    s_sum[ph][c] = s;
    s_sq[ph][c] = q;
    __syncthreads();
    if (threadIdx.x < C) {
        double* pp = partial + ((long long)b * nsp + sp) * 2 * C;
        pp[c] = (s_sum[0][c] + s_sum[1][c]) + (s_sum[2][c] + s_sum[3][c]);
        pp[C + c] = (s_sq[0][c] + s_sq[1][c]) + (s_sq[2][c] + s_sq[3][c]);
    }
}

__global__ void __launch_bounds__(GN_THREADS)
gn_finish_kernel(const float* __restrict__ x, int src_samples, int T, int cpg, float eps, const double* __restrict__ partial, int nsp,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scale,
                 float* __restrict__ shift, int out_ld, int out_off, float* __restrict__ act_out, int act_ld, int act_off) {
    constexpr int C = 192;
    __shared__ double s_sum[C], s_sq[C];
    __shared__ float s_mean[32], s_rstd[32];
    pdl_wait();
    pdl_trigger();
    const int sp = blockIdx.x, b = blockIdx.y;
    const int c = threadIdx.x % C, ph = threadIdx.x / C;
    if (threadIdx.x < C) {
        double s = 0.0, q = 0.0;
        for (int k = 0; k < nsp; ++k) {
            const double* pp = partial + ((long long)b * nsp + k) * 2 * C;
            s += pp[c];
            q += pp[C + c];
        }
        s_sum[c] = s;
        s_sq[c] = q;
    }
    __syncthreads();
    const int ng = C / cpg;
    if ((int)threadIdx.x < ng) {
        double gs = 0.0, gq = 0.0;
        for (int j = 0; j < cpg; ++j) {
            gs += s_sum[threadIdx.x * cpg + j];
            gq += s_sq[threadIdx.x * cpg + j];
        }
        const double n = (double)cpg * T;
        const double mean = gs / n;
        double var = gq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int g = c / cpg;
    const float sc = s_rstd[g] * __ldg(gamma + c);
    const float sh = __ldg(beta + c) - s_mean[g] * sc;
    if (sp == 0 && threadIdx.x < C) {
        scale[(long long)b * out_ld + out_off + c] = sc;
        shift[(long long)b * out_ld + out_off + c] = sh;
    }
    if (act_out != nullptr) {
        const int rows = (T + (int)gridDim.x - 1) / (int)gridDim.x;
        const int t0 = sp * rows, t1 = min(T, t0 + rows);
        const float* xb = x + (long long)(b % src_samples) * T * C + c;
        float* ob = act_out + (long long)b * T * act_ld + act_off + c;
        for (int t2 = t0 + ph; t2 < t1; t2 += 4) ob[(long long)t2 * act_ld] = silu(__ldg(xb + (long long)t2 * C) * sc + sh);
    }
}

// Single-launch GroupNorm: one thread-block CLUSTER per sample.  Each of the cluster's CTAs holds its share of the
// sample's frames in registers (float4 per thread, all loads issued before the first use: ~57 KB in flight per CTA
// instead of a scalar load chain), reduces per-channel sum / sum-of-squares in fp64, publishes its per-group partials in
// shared memory, and after one cluster barrier every CTA reads all partials through distributed shared memory in rank
// order (deterministic), derives scale / shift and writes silu(gn(x)) for its frames from the registers.
// x is read once and the activation written once (the two-kernel version reads x twice), one launch instead of two.
// Requires ceil(ceil(T / cluster) / 8) <= GNF_MAXR.
constexpr int GNF_THREADS = 384;   // 8 row phases x 48 channel quads
constexpr int GNF_MAXR = 10;       // frames per thread
__global__ void __launch_bounds__(GNF_THREADS)
gn_fused_kernel(const float* __restrict__ x, int src_samples, int T, int cpg, float eps, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift, int out_ld, int out_off,
                float* __restrict__ act_out, int act_ld, int act_off) {
    constexpr int C = 192, Q = C / 4, PH = GNF_THREADS / Q;
    __shared__ double s_red[PH][2][C];
    __shared__ double s_part[2][32];     // this CTA's per-group sum / sum of squares (read by the whole cluster)
    __shared__ float s_mean[32], s_rstd[32];
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    pdl_wait();
    pdl_trigger();
    const int nsp = (int)cluster.num_blocks(), sp = (int)cluster.block_rank(), b = blockIdx.y;
    const int q = threadIdx.x % Q, ph = threadIdx.x / Q;
    const int rows = (T + nsp - 1) / nsp;
    const int t0 = sp * rows, t1 = min(T, t0 + rows);
    const float* xb = x + (long long)(b % src_samples) * T * C + q * 4;
    float4 v[GNF_MAXR];
#pragma unroll
    for (int i = 0; i < GNF_MAXR; ++i) {
        const int t = t0 + ph + PH * i;
        v[i] = t < t1 ? ldg4(xb + (long long)t * C) : zero4();
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
    for (int i = 0; i < GNF_MAXR; ++i) {
        s0 += (double)v[i].x; q0 += (double)v[i].x * (double)v[i].x;
        s1 += (double)v[i].y; q1 += (double)v[i].y * (double)v[i].y;
        s2 += (double)v[i].z; q2 += (double)v[i].z * (double)v[i].z;
        s3 += (double)v[i].w; q3 += (double)v[i].w * (double)v[i].w;
    }
    s_red[ph][0][q * 4 + 0] = s0; s_red[ph][0][q * 4 + 1] = s1; s_red[ph][0][q * 4 + 2] = s2; s_red[ph][0][q * 4 + 3] = s3;
    s_red[ph][1][q * 4 + 0] = q0; s_red[ph][1][q * 4 + 1] = q1; s_red[ph][1][q * 4 + 2] = q2; s_red[ph][1][q * 4 + 3] = q3;
    __syncthreads();
    const int ng = C / cpg;
    if ((int)threadIdx.x < 2 * ng) {     // thread (g, which): group g, sum (0) or sum of squares (1)
        const int g = threadIdx.x % ng, w = threadIdx.x / ng;
        double a = 0.0;
        for (int j = 0; j < cpg; ++j) {
            const int c = g * cpg + j;
            double cs = 0.0;
#pragma unroll
            for (int p2 = 0; p2 < PH; ++p2) cs += s_red[p2][w][c];
            a += cs;
        }
        s_part[w][g] = a;
    }
    cluster.sync();
    if ((int)threadIdx.x < ng) {
        double gs = 0.0, gq = 0.0;
        for (int r = 0; r < nsp; ++r) {
            const double* rp = cluster.map_shared_rank(&s_part[0][0], r);
            gs += rp[threadIdx.x];
            gq += rp[32 + threadIdx.x];
        }
        const double n = (double)cpg * T;
        const double mean = gs / n;
        double var = gq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    cluster.sync();                      // also: no CTA leaves while its partials may still be read remotely
    const float4 gm = ldg4(gamma + q * 4), bt = ldg4(beta + q * 4);
    float4 sc, sh;
    {
        const int g0 = (q * 4) / cpg, g1 = (q * 4 + 1) / cpg, g2 = (q * 4 + 2) / cpg, g3 = (q * 4 + 3) / cpg;
        sc.x = s_rstd[g0] * gm.x; sh.x = bt.x - s_mean[g0] * sc.x;
        sc.y = s_rstd[g1] * gm.y; sh.y = bt.y - s_mean[g1] * sc.y;
        sc.z = s_rstd[g2] * gm.z; sh.z = bt.z - s_mean[g2] * sc.z;
        sc.w = s_rstd[g3] * gm.w; sh.w = bt.w - s_mean[g3] * sc.w;
    }
    if (sp == 0 && ph == 0) {
        st4(scale + (long long)b * out_ld + out_off + q * 4, sc);
        st4(shift + (long long)b * out_ld + out_off + q * 4, sh);
    }
    if (act_out != nullptr) {
        float* ob = act_out + (long long)b * T * act_ld + act_off + q * 4;
#pragma unroll
        for (int i = 0; i < GNF_MAXR; ++i) {
            const int t = t0 + ph + PH * i;
            if (t < t1)
                st4(ob + (long long)t * act_ld, make_float4(silu(v[i].x * sc.x + sh.x), silu(v[i].y * sc.y + sh.y),
                                                            silu(v[i].z * sc.z + sh.z), silu(v[i].w * sc.w + sh.w)));
        }
    }
}

// LayerNorm(192) of every row, optionally preceded by a per-(sample, channel) affine (the SpatialTransformer's
// GroupNorm folded in), materialised:  y = LN(x * ps + pb) * gamma + beta   (attention.py:170-191 norm1 / norm3).
// The tcgen05 GEMMs whose output is wider than one tile (q/k/v: 3 tiles, GEGLU: 8 tiles) would otherwise redo this
// transform in their operand producers once per output tile -- the producers, not the tensor pipe, bounded them.
// 16 lanes per row (3 float4 each), two rows per warp; two-pass mean / variance in registers like ALoadLNT.
__global__ void __launch_bounds__(256)
ln192_rows_kernel(const float* __restrict__ x, int M, int T, const float* __restrict__ pre_scale, const float* __restrict__ pre_shift,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ y) {
    constexpr int C = 192;
    pdl_wait();
    pdl_trigger();
    const int row = (blockIdx.x * 256 + threadIdx.x) >> 4, l = threadIdx.x & 15;
    const bool ok = row < M;
    const int r = ok ? row : M - 1;
    const float* xr = x + (long long)r * C;
    float4 v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = ldg4(xr + (l + 16 * j) * 4);
    if (pre_scale != nullptr) {
        const int b = r / T;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 a = ldg4(pre_scale + (long long)b * C + (l + 16 * j) * 4), d = ldg4(pre_shift + (long long)b * C + (l + 16 * j) * 4);
            v[j].x = v[j].x * a.x + d.x; v[j].y = v[j].y * a.y + d.y; v[j].z = v[j].z * a.z + d.z; v[j].w = v[j].w * a.w + d.w;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float a = v[j].x - mean, b2 = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
        q += (a * a + b2 * b2) + (d * d + e * e);
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / C) + eps);
    if (!ok) return;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = (l + 16 * j) * 4;
        const float4 g = ldg4(gamma + k), bb = ldg4(beta + k);
        st4(y + (long long)row * C + k, make_float4((v[j].x - mean) * rstd * g.x + bb.x, (v[j].y - mean) * rstd * g.y + bb.y,
                                                    (v[j].z - mean) * rstd * g.z + bb.z, (v[j].w - mean) * rstd * g.w + bb.w));
    }
}

// Row-wise LayerNorm (+ optional residual add before it):  y = LN(x [+ r]) * gamma + beta.
// One warp per row; C % 128 == 0 not required, C % 4 == 0 and C <= 1024.
// Used by the Wav2Vec2 post-LN encoder layers (TF modeling_wav2vec2.py:592-609).
template <int MAXV>   // float4 per lane, C <= 128*MAXV
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ r, int M, int C, float eps,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, int act = 0 /*1: GELU after*/,
                      __half* __restrict__ y_pair = nullptr /*also (or only, y == null) as a pair tensor of C columns*/,
                      int* __restrict__ flag = nullptr) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= M) return;
    const float* xr = x + (long long)warp * C;
    const float* rr = r ? r + (long long)warp * C : nullptr;
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = (j * 32 + lane) * 4;
        float4 a = zero4();
        if (k < C) {
            a = ldg4(xr + k);
            if (rr) { const float4 d = ldg4(rr + k); a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w; }
        }
        v[j] = a;
        s += (a.x + a.y) + (a.z + a.w);
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = (j * 32 + lane) * 4;
        if (k < C) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int k = (j * 32 + lane) * 4;
        if (k < C) {
            const float4 g = ldg4(gamma + k), bb = ldg4(beta + k);
            float4 o;
            o.x = (v[j].x - mean) * rstd * g.x + bb.x;
            o.y = (v[j].y - mean) * rstd * g.y + bb.y;
            o.z = (v[j].z - mean) * rstd * g.z + bb.z;
            o.w = (v[j].w - mean) * rstd * g.w + bb.w;
            if (act == 1) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
            if (y != nullptr) st4(y + (long long)warp * C + k, o);
            if (y_pair != nullptr) {
                store_pair4(y_pair, warp, C, k, o);
                if (amax4(0.f, o) > P16_LIMIT) atomicOr(flag, 1);
            }
        }
    }
}

// Linear interpolation over frames (align_corners=True; said/model/wav2vec2.py:41-44, ATen
// upsample_linear1d: src = j * (L-1)/(T-1), lambda1 = src - floor(src)) fused with the feature
// projection's LayerNorm(512) (TF modeling_wav2vec2.py:429-434).  x: (B, Lstride >= L frames, C) -> y: (B, T, C).
// One warp per output row; T < 0 disables interpolation semantics (never used: T is always given).
template <int MAXV>
__global__ void __launch_bounds__(256)
interp_layernorm_kernel(const float* __restrict__ x, int B, int L, int Lstride, int T, int C, float eps,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                        __half* __restrict__ y_pair = nullptr, int* __restrict__ flag = nullptr) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * T) return;
    const int b = warp / T, j = warp - b * T;
    const float rscale = (T > 1) ? (float)(L - 1) / (float)(T - 1) : 0.f;
    const float src = rscale * (float)j;
    const int i0 = (int)src;
    const int i1 = i0 + ((i0 < L - 1) ? 1 : 0);
    const float l1 = src - (float)i0, l0 = 1.0f - l1;
    const float* x0 = x + ((long long)b * Lstride + i0) * C;
    const float* x1 = x + ((long long)b * Lstride + i1) * C;
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXV; ++jj) {
        const int k = (jj * 32 + lane) * 4;
        float4 a = zero4();
        if (k < C) {
            const float4 p = ldg4(x0 + k), q = ldg4(x1 + k);
            a.x = l0 * p.x + l1 * q.x; a.y = l0 * p.y + l1 * q.y; a.z = l0 * p.z + l1 * q.z; a.w = l0 * p.w + l1 * q.w;
        }
        v[jj] = a;
        s += (a.x + a.y) + (a.z + a.w);
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXV; ++jj) {
        const int k = (jj * 32 + lane) * 4;
        if (k < C) {
            const float a = v[jj].x - mean, bb = v[jj].y - mean, c = v[jj].z - mean, d = v[jj].w - mean;
            q += (a * a + bb * bb) + (c * c + d * d);
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int jj = 0; jj < MAXV; ++jj) {
        const int k = (jj * 32 + lane) * 4;
        if (k < C) {
            const float4 g = ldg4(gamma + k), bb = ldg4(beta + k);
            float4 o;
            o.x = (v[jj].x - mean) * rstd * g.x + bb.x;
            o.y = (v[jj].y - mean) * rstd * g.y + bb.y;
            o.z = (v[jj].z - mean) * rstd * g.z + bb.z;
            o.w = (v[jj].w - mean) * rstd * g.w + bb.w;
            if (y != nullptr) st4(y + (long long)warp * C + k, o);
            if (y_pair != nullptr) {
                store_pair4(y_pair, warp, C, k, o);
                if (amax4(0.f, o) > P16_LIMIT) atomicOr(flag, 1);
            }
        }
    }
}

}  // namespace said
