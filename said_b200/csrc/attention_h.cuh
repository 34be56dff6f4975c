// Self-attention on tcgen05 with fp16 hi/lo operand pairs ("fp16x3"), flash-style, head dim 32, any T <= 512.
//
//   out = softmax(q k^T * scale) v          (CrossAttention.forward with context=None, ldm/attention.py:86-128)
//
// One CTA per (pair of heads, sample): the kernel is a chain of short dependent phases (MMA -> softmax -> MMA), so two independent
// groups of 8 warps -- one head each, with their own 256 tensor-memory columns, ~99 KB of shared memory (T = 300) and barriers --
// share the SM and fill each other's latency bubbles.  (Two CTAs per SM would do the same, but a kernel that uses tensor memory
// is admitted one CTA per SM: measured with the occupancy API for every block size / shared-memory size.)  Sequences whose
// operands do not fit twice (T > ~340) run one head per CTA.
//   * K and V^T of the head are staged once: fp32 -> fp16 hi/lo, each row [hi(32 dims) | lo(32 dims)] = one 128-byte
//     SWIZZLE_128B row, so the three passes of the split product (hi*hi + lo*hi + hi*lo) are just different 32-byte K-slices of
//     the same tiles; V^T per 64 keys is a 64-row tile [hi dims | lo dims] x 64 keys.
//   * per tile of 128 queries (Q pre-multiplied by scale * log2 e) and per block of 128 keys:
//       S(128 x 128) = Q K^T into TMEM; two threads per row read their 64 scores ONCE, keep the running row maximum / sum
//       (online softmax), and write P = exp2(S - m) back to TMEM IN PLACE as packed fp16 hi / lo (P never touches shared
//       memory); the running output O (in TMEM) is rescaled when the maximum moved; then
//       O[:, 0:64] += P_hi [V_hi | V_lo],  O[:, 0:32] += P_lo V_hi      (A operand from TMEM: the "TS" form of tcgen05.mma)
//   * (O[:, 0:32] + O[:, 32:64]) / rowsum -> global, as fp32 or directly in the pair format of the next GEMM's operand.
#pragma once
#include "gemm_h.cuh"

namespace said {
namespace hx {

constexpr int AH_HD = 32;
constexpr int AH_SM_THREADS = 256;                 // 8 warps: staging / softmax / epilogue, two threads per query row
constexpr int AH_THREADS = AH_SM_THREADS;          // threads per group (= per head); thread 0 of the group also issues its MMAs
constexpr int AH_KB = 128;                         // keys per block
constexpr int AH_S_COL = 0;                        // TMEM columns [0, 128): scores, then P_hi (cols 0-63) | P_lo (cols 64-127) packed 2 per column
constexpr int AH_O_COL = 128;                      // TMEM columns [128, 192): output accumulators
constexpr int AH_TMEM_COLS = 256;
constexpr int AH_MAXT = 512;

SAID_DEVINL void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
SAID_DEVINL void tmem_st32u(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
SAID_DEVINL void tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
SAID_DEVINL uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
SAID_DEVINL float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// eight fp32 values -> 16 bytes of fp16 hi and 16 bytes of fp16 lo
SAID_DEVINL void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
SAID_DEVINL void sts16(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Per-group shared memory (one group = 8 warps = one head): K, V^T, Q, exchange floats, barriers.
inline size_t attention_h_group_bytes(int T) {
    const int Tkp = (T + 15) / 16 * 16;
    const int nch = (T + 63) / 64;
    const size_t raw = (size_t)Tkp * 128 + 1024 /*V tile alignment*/ + (size_t)nch * 8192 + 16384 /*Q*/ + 2 * AH_SM_THREADS * 4 + 64;
    return (raw + 1023) / 1024 * 1024;
}
inline size_t attention_h_smem_bytes(int T, int groups) { return attention_h_group_bytes(T) * groups + 1024; }
// two heads per CTA whenever their operands fit one SM's shared memory together (T <= ~340)
inline int attention_h_groups(int T, int heads) { return (heads % 2 == 0 && attention_h_smem_bytes(T, 2) <= 227 * 1024) ? 2 : 1; }

// grid (heads / groups, samples), block 256 * groups: group g (threads [256 g, 256 g + 256)) processes head blockIdx.x * groups + g with
// its own shared memory, tensor-memory columns [256 g, 256 g + 192) and barriers; the two groups only meet at the TMEM allocation.
__global__ void __launch_bounds__(2 * AH_THREADS, 1)
self_attention_h_kernel(const float* __restrict__ qkv, int ld, int q_off, int k_off, int v_off, int T, float scale,
                        float* __restrict__ out, int ldo, int Tstr /*rows per sample in qkv / out*/,
                        __half* __restrict__ out_pair /*non-null: write the pair tensor (ldo columns) instead of fp32*/,
                        int* __restrict__ flag, uint32_t group_bytes, long long* __restrict__ trace = nullptr /*diagnostics: clock64 stamps of CTA (0,0) group 0 thread 0*/) {
    int tr_n = 0;
    const bool tr_on = trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
    auto stamp = [&]() { if (tr_on && tr_n < 62) trace[tr_n++] = clock64(); };
    stamp();
    extern __shared__ uint8_t smem_raw[];
    const int Tkp = (T + 15) / 16 * 16;
    const int nch = (T + 63) / 64;                    // 64-key chunks of V^T
    const int nkb = (T + AH_KB - 1) / AH_KB;          // 128-key blocks
    const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform (the MMA-issuing warp keeps its
    const int grp = wg >> 3, groups = blockDim.x >> 8;                      //  descriptors in uniform registers)
    const int tid = threadIdx.x & 255, warp = wg & 7;                        // (within the group)
    const uint32_t base0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t base = base0 + (uint32_t)grp * group_bytes;
    const uint32_t k_sm = base;                                            // [Tkp rows][hi 64 B | lo 64 B]
    const uint32_t v_sm = (k_sm + (uint32_t)Tkp * 128u + 1023u) & ~1023u;  // per 64-key chunk: [32 hi-dim rows | 32 lo-dim rows] x 128 B
    const uint32_t q_sm = v_sm + (uint32_t)nch * 8192u;                    // [128 rows][hi | lo]
    const uint32_t xch = q_sm + 16384u;                                    // float[2][AH_SM_THREADS]
    const uint32_t bars = xch + 2u * AH_SM_THREADS * 4u;
    const uint32_t bar_q = bars, bar_s = bars + 8u, bar_p = bars + 16u, bar_o = bars + 24u;
    const uint32_t tmem_slot = base0 + group_bytes - 8u;                   // (last bytes of group 0's region: shared by both groups)
    const int nbar = 1 + grp;                                              // named barrier of this group

    const int h = blockIdx.x * groups + grp, b = blockIdx.y;
    const float* gbase = qkv + (long long)b * Tstr * ld + h * AH_HD;

    if (tid == 0) {
        mbar_init(bar_q, AH_SM_THREADS);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, AH_SM_THREADS);
        mbar_init(bar_o, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, groups == 2 ? 512u : 256u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tmem_cta = tmem_base;
    tmem_base += (uint32_t)grp * 256u;                                     // this group's columns
    pdl_wait();
    pdl_trigger();
    const int n_qtiles = (T + 127) / 128;
    // few samples: gridDim.z = n_qtiles spreads the query tiles of a head over CTAs (each stages K / V itself)
    const int qt_begin = gridDim.z > 1 ? (int)blockIdx.z : 0, qt_end = gridDim.z > 1 ? qt_begin + 1 : n_qtiles;

    // ---------------- MMA issue (warp 0 of the group between its softmax duties: all lanes run the bookkeeping, one elected lane issues) ----------------
    const uint64_t dq = make_desc(q_sm);
    const uint32_t id64 = make_idesc_f16(128, 64), id32 = make_idesc_f16(128, 32);
    auto issue_s = [&](int kb) {     // S = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T: hi / lo are the 32-byte K-slices 0,1 / 2,3 of the same 128-byte rows
        const int nvalid = min(AH_KB, T - kb * AH_KB);
        const uint32_t ids = make_idesc_f16(128, (nvalid + 15) / 16 * 16);
        const uint64_t dk = make_desc(k_sm + (uint32_t)kb * AH_KB * 128u);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint64_t hi = (uint64_t)(k * 2), lo = (uint64_t)((2 + k) * 2);
            mma_f16(tmem_base + AH_S_COL, dq + hi, dk + hi, ids, k != 0 ? 1u : 0u);
            mma_f16(tmem_base + AH_S_COL, dq + lo, dk + hi, ids, 1u);
            mma_f16(tmem_base + AH_S_COL, dq + hi, dk + lo, ids, 1u);
        }
        mma_commit(bar_s);
    };
    auto issue_pv = [&](int kb) {    // O[:, 0:64] += P_hi [V_hi | V_lo],  O[:, 0:32] += P_lo V_hi   (A operand = P in tensor memory)
        const int nvalid = min(AH_KB, T - kb * AH_KB);
        const int ksteps = (nvalid + 15) / 16;
        for (int j = 0; j < ksteps; ++j) {
            const int key0 = kb * AH_KB + 16 * j;
            const uint64_t dv = make_desc(v_sm + (uint32_t)(key0 >> 6) * 8192u) + (uint64_t)(((key0 & 63) >> 4) * 2);
            mma_f16_ts(tmem_base + AH_O_COL, tmem_base + AH_S_COL + 8 * j, dv, id64, (kb | j) != 0 ? 1u : 0u);
            mma_f16_ts(tmem_base + AH_O_COL, tmem_base + AH_S_COL + 64 + 8 * j, dv, id32, 1u);
        }
    };
    uint32_t n_q = 0, n_p = 0;

    {
        // ---------------- stage K and V^T ----------------
        // K: row = key, [hi dims | lo dims].  V^T: per 64-key chunk, row = dim (hi rows 0-31, lo rows 32-63), 8 keys per 16-byte store.
        // The loads of a thread's first V^T item and of up to five K items (18 x 16 bytes) are issued before anything is converted:
        // staging used to take five dependent load round trips (7.7 us of a 41 us CTA at T = 300), now two.
        auto v_load = [&](int i, float4 (&xv)[8]) {
            const int kg = i >> 3, c4 = i & 7;                   // key group (8 keys), dims 4 c4 .. 4 c4 + 3
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int key = kg * 8 + u;
                xv[u] = (key < T) ? ldg4(gbase + (long long)key * ld + v_off + c4 * 4) : zero4();
            }
        };
        auto v_store = [&](int i, const float4 (&xv)[8]) {
            const int kg = i >> 3, c4 = i & 7;
            const int chunk = kg >> 3, kq = kg & 7;              // 16-byte slot of the 128-byte row
            const float* xf = reinterpret_cast<const float*>(xv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float col[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) col[u] = xf[4 * u + e];
                uint4 hi, lo;
                split8(col, hi, lo);
                const int dd = c4 * 4 + e;
                const uint32_t tile = v_sm + (uint32_t)chunk * 8192u;
                sts16(tile + (uint32_t)dd * 128u + (uint32_t)((kq ^ (dd & 7)) << 4), hi);
                sts16(tile + (uint32_t)(32 + dd) * 128u + (uint32_t)((kq ^ ((32 + dd) & 7)) << 4), lo);
            }
        };
        const int n_vitems = nch * 8 * 8;
        float4 xv0[8];
        if (tid < n_vitems) v_load(tid, xv0);
        constexpr int KIT = 5;
        for (int i0 = tid; i0 < Tkp * 4; i0 += KIT * AH_SM_THREADS) {
            float4 ld_[KIT][2];
#pragma unroll
            for (int u = 0; u < KIT; ++u) {
                const int i = i0 + u * AH_SM_THREADS, key = i >> 2, c2 = i & 3;
                const bool ok = i < Tkp * 4 && key < T;
                const float* src = gbase + (long long)(ok ? key : 0) * ld + k_off + c2 * 8;
                ld_[u][0] = ok ? ldg4(src) : zero4();
                ld_[u][1] = ok ? ldg4(src + 4) : zero4();
            }
#pragma unroll
            for (int u = 0; u < KIT; ++u) {
                const int i = i0 + u * AH_SM_THREADS, key = i >> 2, c2 = i & 3;
                if (i < Tkp * 4) {
                    const float x[8] = {ld_[u][0].x, ld_[u][0].y, ld_[u][0].z, ld_[u][0].w, ld_[u][1].x, ld_[u][1].y, ld_[u][1].z, ld_[u][1].w};
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    const uint32_t row = k_sm + (uint32_t)key * 128u;
                    sts16(row + (uint32_t)((c2 ^ (key & 7)) << 4), hi);
                    sts16(row + (uint32_t)(((4 + c2) ^ (key & 7)) << 4), lo);
                }
            }
        }
        if (tid < n_vitems) v_store(tid, xv0);
        for (int i = tid + AH_SM_THREADS; i < n_vitems; i += AH_SM_THREADS) {
            float4 xv[8];
            v_load(i, xv);
            v_store(i, xv);
        }
        stamp();                                               // K / V^T staged
        const int row = tid & 127, part = tid >> 7;            // two threads per query row: 64 of the 128 score columns each
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float qscale = scale * 1.4426950408889634f;       // softmax in base 2
        uint32_t n_s = 0, n_o = 0;
        // raw Q rows of the next query tile, prefetched into registers while the current tile is processed
        float4 qreg[2][2];
        auto load_q = [&](int qt) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = tid + u * AH_SM_THREADS, r = i >> 2, c2 = i & 3;
                const bool ok = qt * 128 + r < T;
                const float* src = gbase + (long long)(ok ? qt * 128 + r : 0) * ld + q_off + c2 * 8;
                qreg[u][0] = ok ? ldg4(src) : zero4();
                qreg[u][1] = ok ? ldg4(src + 4) : zero4();
            }
        };
        // Q tile: 128 rows x [hi | lo], pre-scaled, from the prefetched registers; then the tile's first S GEMM.  For every tile but the
        // first this runs BEFORE the previous tile's output epilogue (the Q buffer is free once the last S GEMM of a tile has
        // completed, and the tensor pipe executes S(next) after P V(last), which reads the columns S(next) overwrites): the GEMM
        // and its hand-off latency overlap the epilogue instead of following it.
        auto stage_q_and_issue = [&]() {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = tid + u * AH_SM_THREADS, r = i >> 2, c2 = i & 3;
                const float x[8] = {qreg[u][0].x * qscale, qreg[u][0].y * qscale, qreg[u][0].z * qscale, qreg[u][0].w * qscale,
                                    qreg[u][1].x * qscale, qreg[u][1].y * qscale, qreg[u][1].z * qscale, qreg[u][1].w * qscale};
                uint4 hi, lo;
                split8(x, hi, lo);
                const uint32_t ra = q_sm + (uint32_t)r * 128u;
                sts16(ra + (uint32_t)((c2 ^ (r & 7)) << 4), hi);
                sts16(ra + (uint32_t)(((4 + c2) ^ (r & 7)) << 4), lo);
            }
            tc_fence_before();
            tc::fence_proxy_async();
            mbar_arrive(bar_q);
            if (warp == 0) {
                tc::mbar_wait_tight(bar_q, n_q & 1u);
                ++n_q;
                tc_fence_after();
                if (tc::elect_one()) issue_s(0);
            }
            __syncwarp();
        };
        load_q(qt_begin);
        stage_q_and_issue();
        for (int qt = qt_begin; qt < qt_end; ++qt) {
            if (qt + 1 < qt_end) load_q(qt + 1);               // in flight during this tile's MMAs and softmax
            float m_run = -INFINITY, l_part = 0.f;
            for (int kb = 0; kb < nkb; ++kb) {
                const int nvalid = min(AH_KB, T - kb * AH_KB);         // valid keys of this block
                const int my0 = part * 64;                              // first column of this thread
                stamp();                                                // (waiting for S)
                tc::mbar_wait_tight(bar_s, n_s & 1u);
                ++n_s;
                tc_fence_after();
                stamp();                                                // S ready
                float s[4][16];
                // warp-uniform: this thread has valid score columns in this block, and its warp has at least one valid query row
                const bool have = my0 < nvalid && qt * 128 + (warp & 3) * 32 < T;
                const bool full = my0 + 64 <= nvalid;                   // no key mask needed (all blocks but the last)
                if (have) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) tc::tmem_ld16_issue(trow + AH_S_COL + my0 + 16 * u, s[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) tc::tmem_ld_wait16(s[u]);
                    if (!full) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (my0 + 16 * u + e >= nvalid) s[u][e] = -INFINITY;     // masked keys: exp2(-inf) = 0 below
                    }
                }
                float mx = -INFINITY;
                if (have) {
                    float m4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        m4[u] = fmaxf(s[u][0], s[u][1]);
#pragma unroll
                        for (int e = 2; e < 16; e += 2) m4[u] = fmaxf(m4[u], fmaxf(s[u][e], s[u][e + 1]));
                    }
                    mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                }
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + (uint32_t)tid * 4u), "f"(mx) : "memory");
                asm volatile("bar.sync %0, %1;" ::"r"(nbar), "n"(AH_SM_THREADS) : "memory");   // also: every thread holds its scores in registers now
                float other;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(xch + (uint32_t)(tid ^ 128) * 4u));
                const float m_new = fmaxf(m_run, fmaxf(mx, other));
                const float alpha = ex2f(m_run - m_new);                // 0 on the first block (m_run = -inf)
                m_run = m_new;
                // P = 2^(s - m) as packed fp16 hi / lo, written over the score columns 16 scores (8 + 8 packed words) at a time: holding
                // all 64 packed words next to the 64 scores spilled half of the scores to local memory inside this loop
                float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t ph[8], pl[8];
                    if (have) {
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            const float p0 = ex2f(s[u][2 * w] - m_new), p1 = ex2f(s[u][2 * w + 1] - m_new);
                            ls0 += p0;
                            ls1 += p1;
                            const __half2 hh = __floats2half2_rn(p0, p1);
                            const float2 hf = __half22float2(hh);
                            ph[w] = *reinterpret_cast<const uint32_t*>(&hh);
                            pl[w] = pack_h2(p0 - hf.x, p1 - hf.y);
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < 8; ++w) { ph[w] = 0u; pl[w] = 0u; }
                    }
                    tmem_st8u(trow + AH_S_COL + part * 32 + 8 * u, ph);
                    tmem_st8u(trow + AH_S_COL + 64 + part * 32 + 8 * u, pl);
                }
                l_part = l_part * alpha + (ls0 + ls1);
                // rescale the running output when the maximum moved (warp-uniform decision: tcgen05.ld/st are warp-wide)
                if (kb > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
                    float o0[16], o1[16], o[32];
                    tc::tmem_ld16_issue(trow + AH_O_COL + part * 32, o0);
                    tc::tmem_ld16_issue(trow + AH_O_COL + part * 32 + 16, o1);
                    tc::tmem_ld_wait16(o0);
                    tc::tmem_ld_wait16(o1);
#pragma unroll
                    for (int e = 0; e < 16; ++e) { o[e] = o0[e] * alpha; o[16 + e] = o1[e] * alpha; }
                    tc::tmem_st32(trow + AH_O_COL + part * 32, o);
                }
                tc::tmem_wait_st();
                tc_fence_before();
                mbar_arrive(bar_p);
                stamp();                                                // softmax of this thread done
                if (warp == 0) {
                    tc::mbar_wait_tight(bar_p, n_p & 1u);
                    ++n_p;
                    tc_fence_after();
                    if (tc::elect_one()) {
                        issue_pv(kb);
                        if (kb == nkb - 1) mma_commit(bar_o);
                        else issue_s(kb + 1);                  // executes after the P V MMAs above (tcgen05.mma runs in issue order)
                    }
                }
                __syncwarp();
            }
            if (qt + 1 < qt_end) stage_q_and_issue();          // next tile's Q and first S GEMM, under this tile's epilogue
            // ---------------- O / rowsum -> global ----------------
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + (uint32_t)(AH_SM_THREADS + tid) * 4u), "f"(l_part) : "memory");
            stamp();
            tc::mbar_wait_tight(bar_o, n_o & 1u);
            stamp();                                                    // O ready
            ++n_o;
            tc_fence_after();
            asm volatile("bar.sync %0, %1;" ::"r"(nbar), "n"(AH_SM_THREADS) : "memory");
            float l_other;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_other) : "r"(xch + (uint32_t)(AH_SM_THREADS + (tid ^ 128)) * 4u));
            const float inv = 1.0f / (l_part + l_other);
            float oa[16], ob[16];
            tc::tmem_ld16_issue(trow + AH_O_COL + part * 16, oa);            // dims 16 part .. +15 of the P V_hi terms
            tc::tmem_ld16_issue(trow + AH_O_COL + 32 + part * 16, ob);       // ... of the P_hi V_lo term
            tc::tmem_ld_wait16(oa);
            tc::tmem_ld_wait16(ob);
            const int q = qt * 128 + row;
            if (q < T) {
                const long long orow = (long long)b * Tstr + q;
                float amax = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 ov = make_float4((oa[4 * c] + ob[4 * c]) * inv, (oa[4 * c + 1] + ob[4 * c + 1]) * inv,
                                                  (oa[4 * c + 2] + ob[4 * c + 2]) * inv, (oa[4 * c + 3] + ob[4 * c + 3]) * inv);
                    const int col = h * AH_HD + part * 16 + 4 * c;
                    if (out_pair != nullptr) {
                        store_pair4(out_pair, orow, ldo, col, ov);
                        amax = amax4(amax, ov);
                    } else {
                        st4(out + orow * ldo + col, ov);
                    }
                }
                if (amax > P16_LIMIT) atomicOr(flag, 1);
            }
        }
    }
    stamp();
    if (tr_on) trace[63] = tr_n;
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        tmem_dealloc(tmem_cta, groups == 2 ? 512u : 256u);
    }
}

}  // namespace hx
}  // namespace said
