// Distribution-level evaluation on the device: the BCVAE encoder over sliding 120-frame windows of
// blendshape-coefficient sequences (said/model/vae.py:26-89 in eval mode; windows as script/test_evaluate.py:53-106) and the
// Frechet distance between two sets of 64-d latents (said/metric/frechet_distance.py:17-64 -> pytorch_fid
// calculate_frechet_distance).  It is the instrument that compares OUTPUT DISTRIBUTIONS of two runs of the path (fp32 vs
// tensor-core modes, epsilon-prediction long chains) where per-sample parity cannot apply.
#pragma once
#include "common.cuh"

namespace said {

constexpr int BCV_SEQ = 120, BCV_IN = 32, BCV_Z = 64;
constexpr int BCV_L1 = 118, BCV_L2 = 116, BCV_L3 = 57, BCV_L4 = 55, BCV_FLAT = 32 * BCV_L4;   // 1760

// BatchNorm (eval) is folded on the host into a per-channel scale / shift applied after the convolution / linear layer.
struct BcvaeWeights {
    const float *c1w, *c1b, *bn1s, *bn1h;     // Conv1d(32,32,3)   + BN + LeakyReLU(0.2)
    const float *c2w, *c2b, *bn2s, *bn2h;     // Conv1d(32,64,3)   + BN + LeakyReLU(0.2)
    const float *c3w, *c3b, *bn3s, *bn3h;     // Conv1d(64,64,4,s2)+ BN + LeakyReLU(0.2)
    const float *c4w, *c4b;                   // Conv1d(64,32,3), Flatten (channel-major)
    const float *f1w, *f1b, *bn4s, *bn4h;     // Linear(1760,256)  + BN + LeakyReLU(0.01)
    const float *f2w, *f2b, *bn5s, *bn5h;     // Linear(256,128)   + BN + LeakyReLU(0.01)
    const float *f3w, *f3b;                   // Linear(128,64)
    const float *muw, *mub;                   // fc_mu Linear(64,64)
};

constexpr size_t bcvae_smem_bytes() {
    return sizeof(float) * (size_t)(BCV_SEQ * BCV_IN + BCV_L1 * 32 + BCV_L2 * 64 + BCV_L3 * 64 + BCV_FLAT + 256 + 128 + 64);
}

SAID_DEVINL float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// One CTA per window.  coeffs: (B, T, 32); window w of clip b = frames [w * step, w * step + 120).  out: (B * nw, 64) latent means.
__global__ void __launch_bounds__(256)
bcvae_encode_kernel(const float* __restrict__ coeffs, int T, int nw, int step, BcvaeWeights W, float* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float* x0 = sm;                         // [120][32]  (frame, channel)
    float* x1 = x0 + BCV_SEQ * BCV_IN;      // [118][32]
    float* x2 = x1 + BCV_L1 * 32;           // [116][64]
    float* x3 = x2 + BCV_L2 * 64;           // [57][64]
    float* x4 = x3 + BCV_L3 * 64;           // [32][55]   channel-major = nn.Flatten order
    float* h1 = x4 + BCV_FLAT;              // [256]
    float* h2 = h1 + 256;                   // [128]
    float* h3 = h2 + 128;                   // [64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x / nw, w = blockIdx.x - b * nw;
    const float* src = coeffs + ((long long)b * T + (long long)w * step) * BCV_IN;
    for (int i = tid; i < BCV_SEQ * BCV_IN; i += 256) x0[i] = __ldg(src + i);
    __syncthreads();
    // conv(k, stride) over channel-last input: out[t][co] = bias[co] + sum_{ci,k} w[co][ci][k] * in[t*stride + k][ci]
    auto conv = [&](const float* in, int Ci, float* o, int Lo, int Co, int K, int stride, const float* wt, const float* bias,
                    const float* bs, const float* bh, bool chan_major) {
        for (int i = tid; i < Lo * Co; i += 256) {
            const int t = i / Co, co = i - t * Co;
            float acc = __ldg(bias + co);
            const float* wr = wt + (long long)co * Ci * K;
            for (int ci = 0; ci < Ci; ++ci)
                for (int k = 0; k < K; ++k) acc = fmaf(__ldg(wr + ci * K + k), in[(t * stride + k) * Ci + ci], acc);
            if (bs != nullptr) acc = leaky(acc * __ldg(bs + co) + __ldg(bh + co), 0.2f);
            if (chan_major) o[co * Lo + t] = acc;
            else o[i] = acc;
        }
        __syncthreads();
    };
    conv(x0, 32, x1, BCV_L1, 32, 3, 1, W.c1w, W.c1b, W.bn1s, W.bn1h, false);
    conv(x1, 32, x2, BCV_L2, 64, 3, 1, W.c2w, W.c2b, W.bn2s, W.bn2h, false);
    conv(x2, 64, x3, BCV_L3, 64, 4, 2, W.c3w, W.c3b, W.bn3s, W.bn3h, false);
    conv(x3, 64, x4, BCV_L4, 32, 3, 1, W.c4w, W.c4b, nullptr, nullptr, true);
    // linear layers: one warp per output, lanes over the contraction
    auto linear = [&](const float* in, int K, float* o, int N, const float* wt, const float* bias, const float* bs, const float* bh) {
        for (int j = warp; j < N; j += 8) {
            const float* wr = wt + (long long)j * K;
            float acc = 0.f;
            for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(wr + k), in[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                acc += __ldg(bias + j);
                if (bs != nullptr) acc = leaky(acc * __ldg(bs + j) + __ldg(bh + j), 0.01f);
                o[j] = acc;
            }
        }
        __syncthreads();
    };
    linear(x4, BCV_FLAT, h1, 256, W.f1w, W.f1b, W.bn4s, W.bn4h);
    linear(h1, 256, h2, 128, W.f2w, W.f2b, W.bn5s, W.bn5h);
    linear(h2, 128, h3, 64, W.f3w, W.f3b, nullptr, nullptr);
    linear(h3, 64, out + (long long)blockIdx.x * BCV_Z, 64, W.muw, W.mub, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
// Frechet distance between N(mu1, S1) and N(mu2, S2) fitted to two sets of 64-d latents, one CTA, fp64:
//   d = |mu1 - mu2|^2 + tr S1 + tr S2 - 2 tr (S1 S2)^(1/2),   tr (S1 S2)^(1/2) = sum_i sqrt(lambda_i(S1^(1/2) S2 S1^(1/2)))
// (covariances unbiased, like np.cov).  Symmetric eigenproblems by cyclic Jacobi rotations.
// ------------------------------------------------------------------------------------------------
constexpr int FD_D = 64;
constexpr size_t frechet_smem_bytes() { return sizeof(double) * (size_t)(4 * FD_D * FD_D + 4 * FD_D); }

// cyclic Jacobi on the symmetric matrix A (FD_D x FD_D, shared memory); on return diag(A) holds the eigenvalues and, if V != null,
// the columns of V the eigenvectors.  All 256 threads must call it.
__device__ void jacobi_eig(double* A, double* V, double* cs /*2 doubles of shared scratch*/) {
    const int tid = threadIdx.x;
    if (V != nullptr)
        for (int i = tid; i < FD_D * FD_D; i += blockDim.x) V[i] = (i / FD_D == i % FD_D) ? 1.0 : 0.0;
    __syncthreads();
    for (int sweep = 0; sweep < 30; ++sweep) {
        // convergence: off-diagonal Frobenius norm (computed redundantly by every thread from shared memory would be slow; use one warp)
        __shared__ double s_off;
        if (tid < 32) {
            double o = 0.0;
            for (int i = tid; i < FD_D * FD_D; i += 32) {
                const int r = i / FD_D, c = i % FD_D;
                if (r != c) o += A[i] * A[i];
            }
            o = warp_sum(o);
            if (tid == 0) s_off = o;
        }
        __syncthreads();
        double tr = 0.0;
        for (int i = 0; i < FD_D; ++i) tr += fabs(A[i * FD_D + i]);
        if (s_off <= 1e-30 * tr * tr + 1e-300) break;
        for (int p = 0; p < FD_D - 1; ++p)
            for (int q = p + 1; q < FD_D; ++q) {
                if (tid == 0) {
                    const double apq = A[p * FD_D + q];
                    double c = 1.0, s = 0.0;
                    if (fabs(apq) > 1e-300) {
                        const double theta = (A[q * FD_D + q] - A[p * FD_D + p]) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                    cs[0] = c;
                    cs[1] = s;
                }
                __syncthreads();
                const double c = cs[0], s = cs[1];
                if (s != 0.0) {
                    // A <- J^T A J with J the rotation in the (p, q) plane: first the columns, then the rows
                    if (tid < FD_D) {
                        const int k = tid;
                        const double akp = A[k * FD_D + p], akq = A[k * FD_D + q];
                        A[k * FD_D + p] = c * akp - s * akq;
                        A[k * FD_D + q] = s * akp + c * akq;
                        if (V != nullptr) {
                            const double vkp = V[k * FD_D + p], vkq = V[k * FD_D + q];
                            V[k * FD_D + p] = c * vkp - s * vkq;
                            V[k * FD_D + q] = s * vkp + c * vkq;
                        }
                    }
                    __syncthreads();
                    if (tid < FD_D) {
                        const int k = tid;
                        const double apk = A[p * FD_D + k], aqk = A[q * FD_D + k];
                        A[p * FD_D + k] = c * apk - s * aqk;
                        A[q * FD_D + k] = s * apk + c * aqk;
                    }
                }
                __syncthreads();
            }
    }
    __syncthreads();
}

// mean (FD_D) and unbiased covariance (FD_D x FD_D) of n rows of x, into shared memory
__device__ void mean_cov(const float* __restrict__ x, int n, double* mu, double* S, double* row /*FD_D doubles scratch*/) {
    const int tid = threadIdx.x;
    for (int i = tid; i < FD_D * FD_D; i += blockDim.x) S[i] = 0.0;
    if (tid < FD_D) {
        double s = 0.0;
        for (int r = 0; r < n; ++r) s += (double)__ldg(x + (long long)r * FD_D + tid);
        mu[tid] = s / n;
    }
    __syncthreads();
    // each thread owns 16 entries (i0..i0+3, j) ... simple mapping: entry e = tid + 256 k
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int r = 0; r < n; ++r) {
        if (tid < FD_D) row[tid] = (double)__ldg(x + (long long)r * FD_D + tid) - mu[tid];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int e = tid + 256 * k;
            acc[k] += row[e / FD_D] * row[e % FD_D];
        }
        __syncthreads();
    }
    const double inv = 1.0 / (double)(n > 1 ? n - 1 : 1);
#pragma unroll
    for (int k = 0; k < 16; ++k) S[tid + 256 * k] = acc[k] * inv;
    __syncthreads();
}

__global__ void __launch_bounds__(256)
frechet_kernel(const float* __restrict__ lat1, int n1, const float* __restrict__ lat2, int n2, double* __restrict__ out /*[4]: fd, |dmu|^2, tr1 + tr2, tr sqrt*/) {
    extern __shared__ __align__(16) double dsm[];
    double* S1 = dsm;                       // covariance 1, then its eigen-decomposition in place
    double* S2 = S1 + FD_D * FD_D;
    double* V = S2 + FD_D * FD_D;
    double* Tm = V + FD_D * FD_D;
    double* mu1 = Tm + FD_D * FD_D;
    double* mu2 = mu1 + FD_D;
    double* row = mu2 + FD_D;
    double* cs = row + FD_D;
    const int tid = threadIdx.x;
    mean_cov(lat1, n1, mu1, S1, row);
    mean_cov(lat2, n2, mu2, S2, row);
    __shared__ double s_dmu, s_tr;
    if (tid == 0) {
        double d = 0.0, t = 0.0;
        for (int i = 0; i < FD_D; ++i) {
            const double e = mu1[i] - mu2[i];
            d += e * e;
            t += S1[i * FD_D + i] + S2[i * FD_D + i];
        }
        s_dmu = d;
        s_tr = t;
    }
    __syncthreads();
    jacobi_eig(S1, V, cs);                  // S1 = V diag(l) V^T
    // R = S1^(1/2) = V diag(sqrt(max(l, 0))) V^T  -> Tm
    for (int e = tid; e < FD_D * FD_D; e += blockDim.x) {
        const int i = e / FD_D, j = e % FD_D;
        double a = 0.0;
        for (int k = 0; k < FD_D; ++k) {
            const double l = S1[k * FD_D + k];
            a += V[i * FD_D + k] * sqrt(l > 0.0 ? l : 0.0) * V[j * FD_D + k];
        }
        Tm[e] = a;
    }
    __syncthreads();
    // V <- R S2 ;  S1 <- (R S2) R, symmetrised
    for (int e = tid; e < FD_D * FD_D; e += blockDim.x) {
        const int i = e / FD_D, j = e % FD_D;
        double a = 0.0;
        for (int k = 0; k < FD_D; ++k) a += Tm[i * FD_D + k] * S2[k * FD_D + j];
        V[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < FD_D * FD_D; e += blockDim.x) {
        const int i = e / FD_D, j = e % FD_D;
        double a = 0.0;
        for (int k = 0; k < FD_D; ++k) a += V[i * FD_D + k] * Tm[k * FD_D + j];
        S1[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < FD_D * FD_D; e += blockDim.x) {
        const int i = e / FD_D, j = e % FD_D;
        if (i < j) {
            const double m = 0.5 * (S1[i * FD_D + j] + S1[j * FD_D + i]);
            S1[i * FD_D + j] = m;
            S1[j * FD_D + i] = m;
        }
    }
    __syncthreads();
    jacobi_eig(S1, nullptr, cs);
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < FD_D; ++i) {
            const double l = S1[i * FD_D + i];
            t += sqrt(l > 0.0 ? l : 0.0);
        }
        out[0] = s_dmu + s_tr - 2.0 * t;
        out[1] = s_dmu;
        out[2] = s_tr;
        out[3] = t;
    }
}

}  // namespace said
