// Attention kernels of the path (fp32, register-tiled, online softmax).
#pragma once
#include "common.cuh"
#include "pair.cuh"

namespace said {

// ------------------------------------------------------------------------------------------------
// Dense multi-head self-attention  softmax(q k^T * scale) v   (CrossAttention.forward with
// context=None, ldm/attention.py:86-128; Wav2Vec2 encoder attention, TF modeling_wav2vec2.py:466-573).
//
// q, k, v are column slices of one fused projection buffer: row (b*T + t), head h at columns
// q_off + h*HD, k_off + h*HD, v_off + h*HD, row stride ld.  Output: out[(b*T+t)*ldo + h*HD + d].
//
// One CTA = (64 queries, one head, one sample); 8 warps x 8 queries.  Keys/values are streamed through
// shared memory in blocks of 128 (K transposed [d][j], V [j][d]), so T is unbounded; per block each warp does:
//   pass 1: S(8 x 128) as a register tile (lane <-> 4 keys x 8 queries), 32 FMA per 6 smem wavefronts
//   online softmax across lanes (xor shuffles), P -> warp-private smem (row stride 12 floats:
//   conflict-free 128-bit stores)
//   pass 2: O(8 x HD) as a register tile (lane <-> 2 queries x HD/8 dims).
// ------------------------------------------------------------------------------------------------
constexpr int ATT_THREADS = 256;
constexpr int ATT_QTILE = 64;
constexpr int ATT_KB = 128;      // keys per block
constexpr int ATT_KSTR = ATT_KB + 4;
constexpr int ATT_PSTR = 12;

template <int HD>
constexpr size_t attention_smem_bytes() {
    return sizeof(float) * ((size_t)HD * ATT_KSTR /*Kt*/ + (size_t)ATT_KB * HD /*V*/ + (size_t)HD * ATT_QTILE /*Qt*/ +
                            (size_t)8 * ATT_KB * ATT_PSTR /*P*/);
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
self_attention_kernel(const float* __restrict__ qkv, int ld, int q_off, int k_off, int v_off, int T, float scale,
                      float* __restrict__ out, int ldo, int Tstr = 0 /*rows per sample in qkv / out (0: T)*/,
                      __half* __restrict__ out_pair = nullptr /*write the pair tensor (ldo columns) instead of fp32*/,
                      int* __restrict__ flag = nullptr) {
    extern __shared__ __align__(16) float smem[];
    float* Kt = smem;                               // [HD][KSTR]   keys of the current block, transposed
    float* Vs = Kt + (size_t)HD * ATT_KSTR;         // [KB][HD]
    float* Qt = Vs + (size_t)ATT_KB * HD;           // [HD][64]
    float* Ps = Qt + (size_t)HD * ATT_QTILE;        // [8 warps][KB][PSTR]
    pdl_wait();
    pdl_trigger();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * ATT_QTILE, h = blockIdx.y, b = blockIdx.z;
    if (Tstr == 0) Tstr = T;
    const float* base = qkv + (long long)b * Tstr * ld;
    constexpr int V4 = HD / 4;

    for (int i = tid; i < ATT_QTILE * V4; i += ATT_THREADS) {
        const int qi = i / V4, d4 = (i % V4) * 4;
        float4 qv = zero4();
        if (q0 + qi < T) qv = ldg4(base + (long long)(q0 + qi) * ld + q_off + h * HD + d4);
        Qt[(d4 + 0) * ATT_QTILE + qi] = qv.x; Qt[(d4 + 1) * ATT_QTILE + qi] = qv.y;
        Qt[(d4 + 2) * ATT_QTILE + qi] = qv.z; Qt[(d4 + 3) * ATT_QTILE + qi] = qv.w;
    }
    const bool active = q0 + warp * 8 < T;          // warp has at least one valid query

    float* Pw = Ps + (size_t)warp * ATT_KB * ATT_PSTR;
    const float* Qw = Qt + warp * 8;
    float mrun[8], lrun[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mrun[i] = -INFINITY; lrun[i] = 0.f; }
    // pass-2 mapping: lane -> queries {2*qp, 2*qp+1}, dims {dc*4 + 32*u .. +3 : u < HD/32}
    constexpr int DU = HD / 32;
    const int qp = lane >> 3, dc = lane & 7;
    float o[2][DU * 4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int d = 0; d < DU * 4; ++d) o[i][d] = 0.f;

    for (int j0 = 0; j0 < T; j0 += ATT_KB) {
        __syncthreads();                            // previous block fully consumed (and Q staged)
        for (int i = tid; i < ATT_KB * V4; i += ATT_THREADS) {
            const int j = i / V4, d4 = (i % V4) * 4;
            float4 kv = zero4(), vv = zero4();
            if (j0 + j < T) {
                kv = ldg4(base + (long long)(j0 + j) * ld + k_off + h * HD + d4);
                vv = ldg4(base + (long long)(j0 + j) * ld + v_off + h * HD + d4);
            }
            Kt[(d4 + 0) * ATT_KSTR + j] = kv.x; Kt[(d4 + 1) * ATT_KSTR + j] = kv.y;
            Kt[(d4 + 2) * ATT_KSTR + j] = kv.z; Kt[(d4 + 3) * ATT_KSTR + j] = kv.w;
            st4(Vs + (size_t)j * HD + d4, vv);
        }
        __syncthreads();
        if (!active) continue;
        // ---- pass 1: scores for keys j0 + c*32 + lane ----
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[i][c] = 0.f;
#pragma unroll 4
        for (int d = 0; d < HD; ++d) {
            const float4 qa = ld4(Qw + d * ATT_QTILE), qb = ld4(Qw + d * ATT_QTILE + 4);
            const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            float kv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) kv[c] = Kt[d * ATT_KSTR + c * 32 + lane];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[i][c] = fmaf(qv[i], kv[c], s[i][c]);
        }
        // ---- online softmax (per query, across the keys of this block) ----
        float cr0 = 0.f, cr1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float bm = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s[i][c] = (j0 + c * 32 + lane < T) ? s[i][c] * scale : -INFINITY;
                bm = fmaxf(bm, s[i][c]);
            }
            bm = warp_max(bm);
            const float mnew = fmaxf(mrun[i], bm);     // finite: every block holds >= 1 valid key
            float ps = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s[i][c] = expf(s[i][c] - mnew);
                ps += s[i][c];
            }
            ps = warp_sum(ps);
            const float corr = expf(mrun[i] - mnew);   // exp(-inf) = 0 on the first block
            lrun[i] = lrun[i] * corr + ps;
            mrun[i] = mnew;
            if (qp == (i >> 1)) { if (i & 1) cr1 = corr; else cr0 = corr; }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float* pr = Pw + (size_t)(c * 32 + lane) * ATT_PSTR;
            st4(pr, make_float4(s[0][c], s[1][c], s[2][c], s[3][c]));
            st4(pr + 4, make_float4(s[4][c], s[5][c], s[6][c], s[7][c]));
        }
        __syncwarp();
        // ---- pass 2: O = O*corr + P V ----
#pragma unroll
        for (int d = 0; d < DU * 4; ++d) { o[0][d] *= cr0; o[1][d] *= cr1; }
        const int jn = min(ATT_KB, T - j0);
        for (int j = 0; j < jn; ++j) {
            const float2 p = *reinterpret_cast<const float2*>(Pw + (size_t)j * ATT_PSTR + 2 * qp);
            const float* vr = Vs + (size_t)j * HD + dc * 4;
#pragma unroll
            for (int u = 0; u < DU; ++u) {
                const float4 v = ld4(vr + 32 * u);
                o[0][4 * u + 0] = fmaf(p.x, v.x, o[0][4 * u + 0]); o[0][4 * u + 1] = fmaf(p.x, v.y, o[0][4 * u + 1]);
                o[0][4 * u + 2] = fmaf(p.x, v.z, o[0][4 * u + 2]); o[0][4 * u + 3] = fmaf(p.x, v.w, o[0][4 * u + 3]);
                o[1][4 * u + 0] = fmaf(p.y, v.x, o[1][4 * u + 0]); o[1][4 * u + 1] = fmaf(p.y, v.y, o[1][4 * u + 1]);
                o[1][4 * u + 2] = fmaf(p.y, v.z, o[1][4 * u + 2]); o[1][4 * u + 3] = fmaf(p.y, v.w, o[1][4 * u + 3]);
            }
        }
        __syncwarp();
    }
    if (!active) return;
    // ---- normalise and store ----
    float l0 = 1.f, l1 = 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (qp == i) { l0 = lrun[2 * i]; l1 = lrun[2 * i + 1]; }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int q = q0 + warp * 8 + 2 * qp + i;
        if (q >= T) continue;
        const float inv = 1.0f / (i == 0 ? l0 : l1);
        const long long orow_i = (long long)b * Tstr + q;
#pragma unroll
        for (int u = 0; u < DU; ++u) {
            const float4 ov = make_float4(o[i][4 * u] * inv, o[i][4 * u + 1] * inv, o[i][4 * u + 2] * inv, o[i][4 * u + 3] * inv);
            if (out_pair != nullptr) {
                store_pair4(out_pair, orow_i, ldo, h * HD + dc * 4 + 32 * u, ov);
                if (amax4(0.f, ov) > P16_LIMIT) atomicOr(flag, 1);
            } else {
                st4(out + orow_i * ldo + h * HD + dc * 4 + 32 * u, ov);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Aligned cross-attention (BasicTransformerBlock attn2 with the alignment bias, attention.py:170-191).
// With the audio features interpolated to one per frame (diffusion.py:387) the mask leaves query i
// exactly the keys {i-1, i, i+1}, and K/V depend only on the audio, so they are
// projected once per clip ("K/V hoist") into kv: row (clip*T + t), K at column kv_off, V at
// kv_off + 192, row stride kv_ld.  Unconditional samples (context = null embedding on every frame,
// diffusion.py:397-400) see identical keys and values, so their output is the constant v_null.
// q rows exist for the conditional samples only: q[(b - n_uncond)*T + t].  For the unconditional rows the kernel
// directly writes the block's next residual state (see below).
// ------------------------------------------------------------------------------------------------
// Eight lanes per (row, head): lane j owns dims 4j..4j+3, so every global access of a warp is one contiguous 512-byte
// run (4 adjacent heads of a row) and all seven loads of a lane are independent; the three dot products are reduced
// over the 8 lanes with shuffles.
__global__ void __launch_bounds__(256)
cross_attention3_kernel(const float* __restrict__ q, const float* __restrict__ kv, int kv_ld, int kv_off,
                        const float* __restrict__ c_null, const float* __restrict__ x_res, float* __restrict__ x_out,
                        int res_rows, int n_uncond, int Bp, int T, float scale, float* __restrict__ out) {
    constexpr int C = 192, HD = 32, H = 6;
    pdl_wait();
    pdl_trigger();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long idx = gid >> 3;                       // (row, head)
    const int j = (int)(gid & 7);
    const bool live = idx < (long long)Bp * T * H;        // dead lanes still take part in the shuffles
    const long long ci = live ? idx : 0;
    const int h = (int)(ci % H);
    const long long row = ci / H;
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    const int col = h * HD + j * 4;
    // null-condition branch: attention output is the constant v_null for every query, so the whole
    // "to_out(attn2) + x" step collapses to  x_out = x_res + c_null  with c_null = W_o v_null + b_o.
    // (No early return before the shuffles: they use the full-warp mask.)
    const bool uncond = b < n_uncond;
    const long long crow = uncond ? 0 : (long long)(b - n_uncond) * T + t;
    float4 qv = zero4(), k0 = zero4(), k1 = zero4(), k2 = zero4(), v0 = zero4(), v1 = zero4(), v2 = zero4();
    const bool has0 = t > 0, has2 = t < T - 1;
    if (live && uncond) {
        const float4 a = ldg4(x_res + (row % res_rows) * C + col), c4 = ldg4(c_null + col);   // x_res may hold only the shared samples
        st4(x_out + row * C + col, make_float4(a.x + c4.x, a.y + c4.y, a.z + c4.z, a.w + c4.w));
    } else if (live) {
        const float* kr = kv + crow * kv_ld + kv_off + col;
        qv = ldg4(q + crow * C + col);
        k1 = ldg4(kr);
        v1 = ldg4(kr + C);
        if (has0) { k0 = ldg4(kr - kv_ld); v0 = ldg4(kr - kv_ld + C); }
        if (has2) { k2 = ldg4(kr + kv_ld); v2 = ldg4(kr + kv_ld + C); }
    }
    float s0 = (qv.x * k0.x + qv.y * k0.y) + (qv.z * k0.z + qv.w * k0.w);
    float s1 = (qv.x * k1.x + qv.y * k1.y) + (qv.z * k1.z + qv.w * k1.w);
    float s2 = (qv.x * k2.x + qv.y * k2.y) + (qv.z * k2.z + qv.w * k2.w);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (!live || uncond) return;
    s0 = has0 ? s0 * scale : -INFINITY;
    s1 = s1 * scale;
    s2 = has2 ? s2 * scale : -INFINITY;
    const float m = fmaxf(s0, fmaxf(s1, s2));
    const float p0 = expf(s0 - m), p1 = expf(s1 - m), p2 = expf(s2 - m);
    const float inv = 1.0f / ((p0 + p1) + p2);
    const float w0 = p0 * inv, w1 = p1 * inv, w2 = p2 * inv;
    float4 acc;
    acc.x = fmaf(w2, v2.x, fmaf(w1, v1.x, w0 * v0.x));
    acc.y = fmaf(w2, v2.y, fmaf(w1, v1.y, w0 * v0.y));
    acc.z = fmaf(w2, v2.z, fmaf(w1, v1.z, w0 * v0.z));
    acc.w = fmaf(w2, v2.w, fmaf(w1, v1.w, w0 * v0.w));
    st4(out + row * C + col, acc);
}

}  // namespace said
