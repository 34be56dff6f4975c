// Self-attention on the 5th-gen tensor cores (tcgen05), head dim 32, up to 304 keys, fp32-level accuracy.
//
//   out = softmax(q k^T * scale) v          (CrossAttention.forward with context=None, ldm/attention.py:86-128)
//
// One CTA per (head, sample).  The head's K (rows = keys) and [V_hi | V_lo]^T (rows = head dims, hi then lo) are staged once
// in shared memory as TF32 hi/lo in the UMMA K-major SWIZZLE_128B layout (a key's 32 dims, or 32 keys of one dim, are
// exactly one 128-byte swizzle row); all global loads of a staging batch are issued before the first shared store.
// Then, per tile of 128 queries (the next tile's Q rows are prefetched into registers):
//   1. Q hi/lo -> smem;  S(128 x Tk) = Q K^T with 3xTF32 (hi*hi + lo*hi + hi*lo) into TMEM (N = 256 + rest)
//   2. four threads per query row (= TMEM lane; 16 softmax warps): row max over the valid keys, then per chunk of
//      32 keys p = exp(s*scale - max), row sum, TF32 hi/lo of p -> smem (A operand); only the last chunk masks keys
//   3. O[:, 0:32] += P_hi V_hi + P_lo V_hi and O[:, 32:64] += P_hi V_lo as TWO MMAs per k-step (P_hi x [V_hi | V_lo] is one
//      N = 64 MMA): these small MMAs are bound by the single issuing thread, not by the tensor pipe
//   4. (O[:, 0:32] + O[:, 32:64]) / rowsum -> global
// The softmax threads and the single MMA-issuing thread hand work back and forth through mbarriers (operands ready /
// MMAs complete, ping-pong P buffers); one CTA per SM fills the shared memory, the chip-level parallelism comes from the
// 6 x samples CTAs.
#pragma once
#include "gemm_tc.cuh"
#include "pair.cuh"

namespace said {
namespace tc {

static_assert(BK == 32 && ROW_BYTES == 128, "attention tiles assume 128-byte SWIZZLE_128B rows");
constexpr int ATC_HD = 32;
constexpr int ATC_MAXKEYS = 304;          // K and V^T hi/lo for the whole head must fit one SM's shared memory
constexpr int ATC_PARTS = 4;               // threads per query row: each owns 32 / ATC_PARTS columns of every 32-key chunk
constexpr int ATC_CP = 32 / ATC_PARTS;     // (the softmax is instruction-latency bound: 16 warps hide what 8 could not)
constexpr int ATC_SM_THREADS = 128 * ATC_PARTS;    // warps 0-15: staging / softmax / epilogue
constexpr int ATC_MMA_WARP = ATC_SM_THREADS / 32;  // warp 16: MMA issuer
constexpr int ATC_THREADS = ATC_SM_THREADS + 32;
static_assert(ATC_CP == 8, "the softmax uses 8-column TMEM loads");
constexpr int ATC_S_COL = 0;               // TMEM columns [0, Tk): scores
constexpr int ATC_O_COL = 320;             // TMEM columns [320, 384): output accumulators (P V_hi terms | P_hi V_lo term)

inline size_t attention_tc_smem_bytes(int T) {
    const int Tk = (T + 15) / 16 * 16;
    const int nch = (T + 31) / 32;
    return (size_t)2 * Tk * 128            // K hi/lo
           + (size_t)2 * nch * 32 * 128    // V^T [hi | lo], one 64(d hi, d lo) x 32(keys) tile per key chunk
           + (size_t)2 * 2 * 128 * 128     // two operand buffers (hi/lo): P chunk ping-pong; buffer 1 doubles as the Q tile
           + 2 * ATC_SM_THREADS * 4        // row max / row sum exchange
           + 1024 + 64;
}

SAID_DEVINL float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
SAID_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(ATC_THREADS, 1)
self_attention_tc_kernel(const float* __restrict__ qkv, int ld, int q_off, int k_off, int v_off, int T, float scale,
                         float* __restrict__ out, int ldo, int Tstr /*rows per sample in qkv / out*/,
                         __half* __restrict__ out_pair /*non-null: write the pair tensor (ldo columns) instead of fp32*/,
                         int* __restrict__ flag) {
    extern __shared__ uint8_t smem_raw[];
    const int Tk = (T + 15) / 16 * 16;
    const int nch = (T + 31) / 32;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t k_hi = base, k_lo = k_hi + (uint32_t)Tk * 128u;
    const uint32_t kv_end = k_lo + (uint32_t)Tk * 128u;
    // V^T: per 32-key chunk one 64-row operand [hi dims 0..31 | lo dims 0..31] (8 KB), so P_hi * [V_hi | V_lo] is ONE N = 64 MMA
    const uint32_t v_hi = (kv_end + 1023u) & ~1023u, v_lo = v_hi + 4096u;
    const uint32_t buf0 = v_hi + (uint32_t)nch * 8192u;          // operand buffer b: hi at bufb, lo at bufb + 16 KB
    const uint32_t buf1 = buf0 + 32768u;                         // buffer 1 also holds the Q tile for the S job
    const uint32_t xch = buf1 + 32768u;                          // float[2][ATC_SM_THREADS]
    const uint32_t bars = xch + 2u * ATC_SM_THREADS * 4u;
    const uint32_t bar_s_ready = bars, bar_s_done = bars + 8u;
    auto bar_p_ready = [&](int b) { return bars + 16u + 8u * b; };
    auto bar_p_done = [&](int b) { return bars + 32u + 8u * b; };
    const uint32_t tmem_slot = bars + 48u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const float* gbase = qkv + (long long)b * Tstr * ld + h * ATC_HD;

    if (tid == ATC_SM_THREADS) {
        mbar_init(bar_s_ready, ATC_SM_THREADS);
        mbar_init(bar_s_done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_p_ready(i), ATC_SM_THREADS);
            mbar_init(bar_p_done(i), 1);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == ATC_MMA_WARP) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();      // barrier init and the TMEM allocation above overlap the previous kernel's tail
    pdl_trigger();
    const int n_qtiles = (T + 127) / 128;

    if (warp < ATC_MMA_WARP) {
        // ---------------- stage K (rows = keys) and V^T (rows = dims), TF32 hi/lo ----------------
        // All global loads of a batch are issued before the first shared-memory store (the volatile stores would
        // otherwise serialise one load round trip per iteration: staging was ~1/3 of the CTA's lifetime).
        constexpr int QR = 128 * 8 / ATC_SM_THREADS;       // float4 of a Q tile per thread
        float4 qreg[QR];                                   // raw Q rows of the next query tile (prefetched)
        auto load_q = [&](int qt) {
#pragma unroll
            for (int u = 0; u < QR; ++u) {
                const int i = tid + u * ATC_SM_THREADS, r = i >> 3, c = i & 7;
                qreg[u] = (qt * 128 + r < T) ? ldg4(gbase + (long long)(qt * 128 + r) * ld + q_off + c * 4) : zero4();
            }
        };
        load_q(0);
        constexpr int KB = 5;
        for (int i0 = tid; i0 < Tk * 8; i0 += ATC_SM_THREADS * KB) {
            float4 x[KB];
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const int i = i0 + u * ATC_SM_THREADS, key = i >> 3, c = i & 7;
                x[u] = (key < T) ? ldg4(gbase + (long long)key * ld + k_off + c * 4) : zero4();
            }
#pragma unroll
            for (int u = 0; u < KB; ++u) {
                const int i = i0 + u * ATC_SM_THREADS, key = i >> 3, c = i & 7;
                if (i < Tk * 8) {
                    const uint32_t off = (uint32_t)key * 128u + (uint32_t)((c ^ (key & 7)) << 4);
                    float4 hh, ll;
                    tf32_split4(x[u], hh, ll);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(k_hi + off), "f"(hh.x), "f"(hh.y), "f"(hh.z), "f"(hh.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(k_lo + off), "f"(ll.x), "f"(ll.y), "f"(ll.z), "f"(ll.w) : "memory");
                }
            }
        }
        // V^T: one item = 4 consecutive keys x 4 dims, transposed in registers -> one 16-byte store per dim
        for (int i0 = tid; i0 < nch * 64; i0 += ATC_SM_THREADS) {
            const int kg = i0 >> 3, c4 = i0 & 7;             // key group (4 keys), dims c4*4 .. +3
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int key = kg * 4 + u;
                x[u] = (key < T) ? ldg4(gbase + (long long)key * ld + v_off + c4 * 4) : zero4();
            }
            const int chunk = kg >> 3, kq = kg & 7;
            const float xv[4][4] = {{x[0].x, x[1].x, x[2].x, x[3].x}, {x[0].y, x[1].y, x[2].y, x[3].y},
                                    {x[0].z, x[1].z, x[2].z, x[3].z}, {x[0].w, x[1].w, x[2].w, x[3].w}};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int dd = c4 * 4 + e;
                const uint32_t off = (uint32_t)chunk * 8192u + (uint32_t)dd * 128u + (uint32_t)((kq ^ (dd & 7)) << 4);
                float4 hh, ll;
                tf32_split4(make_float4(xv[e][0], xv[e][1], xv[e][2], xv[e][3]), hh, ll);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(v_hi + off), "f"(hh.x), "f"(hh.y), "f"(hh.z), "f"(hh.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(v_lo + off), "f"(ll.x), "f"(ll.y), "f"(ll.z), "f"(ll.w) : "memory");
            }
        }
        const int row = tid & 127, part = tid >> 7;       // ATC_PARTS threads per query row: column slices of every 32-key chunk
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float sl2 = scale * 1.4426950408889634f;    // softmax in base 2: exp(x) = 2^(x log2 e)
        uint32_t s_jobs = 0, p_use[2] = {0u, 0u};
        for (int qt = 0; qt < n_qtiles; ++qt) {
            // ---------------- Q tile (into operand buffer 1) from the prefetched registers ----------------
#pragma unroll
            for (int u = 0; u < QR; ++u) {
                const int i = tid + u * ATC_SM_THREADS, r = i >> 3, c = i & 7;
                const float4 x = qreg[u];
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
                float4 hh, ll;
                tf32_split4(x, hh, ll);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf1 + off), "f"(hh.x), "f"(hh.y), "f"(hh.z), "f"(hh.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf1 + 16384u + off), "f"(ll.x), "f"(ll.y), "f"(ll.z), "f"(ll.w) : "memory");
            }
            fence_proxy_async();
            mbar_arrive(bar_s_ready);                     // job: S = Q K^T
            if (qt + 1 < n_qtiles) load_q(qt + 1);        // in flight during the MMAs and the softmax of this tile
            mbar_wait(bar_s_done, s_jobs & 1u);
            ++s_jobs;
            tc_fence_after();
            // ---------------- row max over the valid keys (each thread: its 8-column slice of every chunk) ----------------
            float mx = -INFINITY;
            for (int c0 = 0; c0 < nch; c0 += 4) {          // up to four TMEM loads in flight
                float v[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u < nch) tmem_ld8_issue(trow + ATC_S_COL + (c0 + u) * 32 + part * ATC_CP, v[u]);   // warp-uniform
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u < nch) tmem_ld_wait8(v[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c0 + u < nch - 1) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[u][e]);
                    } else if (c0 + u == nch - 1) {        // only the last chunk can hold padding keys
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if ((c0 + u) * 32 + part * ATC_CP + e < T) mx = fmaxf(mx, v[u][e]);
                    }
                }
            }
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + (uint32_t)tid * 4u), "f"(mx) : "memory");
            named_bar_sync(1, ATC_SM_THREADS);
#pragma unroll
            for (int o = 1; o < ATC_PARTS; ++o) {
                float other;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch + (uint32_t)((tid + 128 * o) & (ATC_SM_THREADS - 1)) * 4u));
                mx = fmaxf(mx, other);
            }
            const float msl2 = mx * sl2;
            // ---------------- P chunks (ping-pong buffers) and O += P V ----------------
            float ls0 = 0.f, ls1 = 0.f;
            for (int ch = 0; ch < nch; ++ch) {
                const int pb = ch & 1;
                float p[8];
                {
                    float v[8];
                    tmem_ld8_issue(trow + ATC_S_COL + ch * 32 + part * ATC_CP, v);
                    tmem_ld_wait8(v);
                    if (ch < nch - 1) {                    // warp-uniform: full chunks need no key mask
#pragma unroll
                        for (int e = 0; e < 8; ++e) p[e] = ex2_approx(v[e] * sl2 - msl2);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) p[e] = (ch * 32 + part * ATC_CP + e < T) ? ex2_approx(v[e] * sl2 - msl2) : 0.f;
                    }
                    ls0 += (p[0] + p[1]) + (p[2] + p[3]);
                    ls1 += (p[4] + p[5]) + (p[6] + p[7]);
                }
                // the job that last read this buffer (chunk ch - 2, or the S job for buffer 1) must have completed
                if (ch >= 2) mbar_wait(bar_p_done(pb), (p_use[pb] - 1u) & 1u);
                const uint32_t pbase = (pb ? buf1 : buf0) + (uint32_t)row * 128u;
#pragma unroll
                for (int c = 0; c < ATC_CP / 4; ++c) {
                    const uint32_t off = pbase + (uint32_t)(((part * (ATC_CP / 4) + c) ^ (row & 7)) << 4);
                    float4 hh, ll;
                    tf32_split4(make_float4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]), hh, ll);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(off), "f"(hh.x), "f"(hh.y), "f"(hh.z), "f"(hh.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(off + 16384u), "f"(ll.x), "f"(ll.y), "f"(ll.z), "f"(ll.w) : "memory");
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(bar_p_ready(pb));             // job: O += P_ch V_ch
                ++p_use[pb];
            }
            float lsum = ls0 + ls1;
            // ---------------- wait for the last job on each buffer, then O / rowsum -> global ----------------
            mbar_wait(bar_p_done(0), (p_use[0] - 1u) & 1u);
            if (nch >= 2) mbar_wait(bar_p_done(1), (p_use[1] - 1u) & 1u);
            tc_fence_after();
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + (uint32_t)(ATC_SM_THREADS + tid) * 4u), "f"(lsum) : "memory");
            named_bar_sync(1, ATC_SM_THREADS);
#pragma unroll
            for (int o = 1; o < ATC_PARTS; ++o) {
                float other;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch + (uint32_t)(ATC_SM_THREADS + ((tid + 128 * o) & (ATC_SM_THREADS - 1))) * 4u));
                lsum += other;
            }
            {
                float v[8], w[8];
                tmem_ld8_issue(trow + ATC_O_COL + part * ATC_CP, v);
                tmem_ld8_issue(trow + ATC_O_COL + ATC_HD + part * ATC_CP, w);
                tmem_ld_wait8(v);
                tmem_ld_wait8(w);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += w[e];
                const int q = qt * 128 + row;
                if (q < T) {
                    const float inv = 1.0f / lsum;
                    const float4 o0 = make_float4(v[0] * inv, v[1] * inv, v[2] * inv, v[3] * inv);
                    const float4 o1 = make_float4(v[4] * inv, v[5] * inv, v[6] * inv, v[7] * inv);
                    const long long orow_i = (long long)b * Tstr + q;
                    if (out_pair != nullptr) {
                        store_pair4(out_pair, orow_i, ldo, h * ATC_HD + part * ATC_CP, o0);
                        store_pair4(out_pair, orow_i, ldo, h * ATC_HD + part * ATC_CP + 4, o1);
                        if (amax4(amax4(0.f, o0), o1) > P16_LIMIT) atomicOr(flag, 1);
                    } else {
                        float* orow = out + orow_i * ldo + h * ATC_HD + part * ATC_CP;
                        st4(orow, o0);
                        st4(orow + 4, o1);
                    }
                }
            }
            tc_fence_before();                            // TMEM reads done before the next tile's MMAs overwrite S / O
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer ----------------
        uint32_t s_jobs = 0, p_cnt[2] = {0u, 0u};
        const int n1 = Tk > 256 ? 256 : Tk, n2 = Tk - n1;
        const uint32_t id1 = make_idesc_tf32(128, n1), id2 = make_idesc_tf32(128, n2 > 0 ? n2 : 16), idv = make_idesc_tf32(128, ATC_HD), idv2 = make_idesc_tf32(128, 2 * ATC_HD);
        const uint64_t dqh = make_desc(buf1), dql = make_desc(buf1 + 16384u);
        const uint64_t dkh = make_desc(k_hi), dkl = make_desc(k_lo);
        const uint64_t dkh2 = make_desc(k_hi + 256u * 128u), dkl2 = make_desc(k_lo + 256u * 128u);
        for (int qt = 0; qt < n_qtiles; ++qt) {
            mbar_wait(bar_s_ready, s_jobs & 1u);
            ++s_jobs;
            tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t adv = (uint64_t)(k4 * 2);
                const uint32_t acc = k4 != 0 ? 1u : 0u;
                mma_tf32(tmem_base + ATC_S_COL, dqh + adv, dkh + adv, id1, acc);
                mma_tf32(tmem_base + ATC_S_COL, dql + adv, dkh + adv, id1, 1u);
                mma_tf32(tmem_base + ATC_S_COL, dqh + adv, dkl + adv, id1, 1u);
                if (n2 > 0) {
                    mma_tf32(tmem_base + ATC_S_COL + 256, dqh + adv, dkh2 + adv, id2, acc);
                    mma_tf32(tmem_base + ATC_S_COL + 256, dql + adv, dkh2 + adv, id2, 1u);
                    mma_tf32(tmem_base + ATC_S_COL + 256, dqh + adv, dkl2 + adv, id2, 1u);
                }
            }
            mma_commit(bar_s_done);
            for (int ch = 0; ch < nch; ++ch) {
                const int pb = ch & 1;
                mbar_wait(bar_p_ready(pb), p_cnt[pb] & 1u);
                ++p_cnt[pb];
                tc_fence_after();
                const uint32_t pa = pb ? buf1 : buf0;
                const uint64_t dph = make_desc(pa), dpl = make_desc(pa + 16384u);
                // O[:, 0:32] += P_hi V_hi + P_lo V_hi,  O[:, 32:64] += P_hi V_lo  (summed in the epilogue): two MMAs per
                // k-step instead of three -- these N = 32 MMAs are bound by the issuing thread, not by the tensor pipe
                const uint64_t dvh = make_desc(v_hi + (uint32_t)ch * 8192u);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const uint64_t adv = (uint64_t)(k4 * 2);
                    mma_tf32(tmem_base + ATC_O_COL, dph + adv, dvh + adv, idv2, (ch | k4) != 0 ? 1u : 0u);
                    mma_tf32(tmem_base + ATC_O_COL, dpl + adv, dvh + adv, idv, 1u);
                }
                mma_commit(bar_p_done(pb));
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == ATC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
}  // namespace said
