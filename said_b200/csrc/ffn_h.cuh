// Fused feed-forward of the transformer block on tcgen05 (fp16x3 operands): GEGLU projection, GEGLU, output projection folded
// with proj_out, bias and the block's residual -- ONE kernel, the 768-wide intermediate never leaves the SM.
//
//   out = [ geglu(LN3(x2) W1 + b1) | x2 ] [W2 Wp ; Wp] + b + h          (attention.py:25-51, 192, 232-234; weights folded at load)
//
// Per tile of 128 rows (persistent, one CTA per SM, 19 warps):
//   * the LayerNorm-ed rows (pair format, 96 KB) are TMA-loaded once and stay resident;
//   * for each of the 6 PAIRS of 64-column hidden chunks:  acc4 = LN3 . W1[:, pair]  (N = 256 interleaved value / gate columns,
//     K = 192) -- one N = 256 instruction reads the activation operand once for 256 columns, where two N = 128 instructions are
//     bound by shared-memory operand bandwidth (measured: 96 cycles per N = 128 MMA against 64 of arithmetic);
//   * two groups of 8 epilogue warps read one half of acc4 each, apply bias + GEGLU and split to fp16 hi / lo; the even chunk
//     goes straight into shared memory in the UMMA operand layout, the odd chunk into the last 64 columns of tensor memory (packed
//     fp16 pairs, the A operand of a TS-form MMA) -- two buffers, so neither group waits for the other; the MMA warp accumulates
//     acc5 += chunk . W2[chunk rows, :]  (N = 192, K = 64) behind the NEXT pair's first GEMM, so the tensor pipe always has
//     queued work while a chunk is in the epilogue warps;
//   * finally acc5 += x2 . Wp (K = 192, the residual stream's own pair tensor re-using the resident operand slots), and the
//     epilogue adds bias and the block input h and writes fp32 (the residual rows are prefetched while the last MMAs run).
// Weights stream through a three-slot ring of 32 KB PLANES (the hi plane of a k-chunk feeds the hi.hi and lo.hi products, the lo
// plane the hi.lo product), cp.async.bulk from pre-swizzled images: 1.9 MB per tile from L2.
// What this removes from the unfused path: the (M, 768) GEGLU tensor's HBM round trip (118 MB written + read per block and
// step at batch 64), one launch, and the second kernel's pipeline fill / drain.
#pragma once
#include "gemm_h.cuh"
#include "attention_h.cuh"   // mma_f16_ts

namespace said {
namespace hx {

constexpr int FFN_C = 192;                 // model channels (K of the first GEMM, N of the second)
constexpr int FFN_JC = 64;                 // hidden columns per chunk
constexpr int FFN_NJ = 12;                 // chunks: 768 hidden columns
constexpr int FFN_NP = FFN_NJ / 2;         // chunk pairs: one N = 256 accumulator each
constexpr int FFN_EPI_WARPS = 16;
constexpr int FFN_THREADS = (FFN_EPI_WARPS + 3) * 32;
constexpr int FFN_A_BYTES = 3 * A_STAGE;               // resident operand: 128 rows x 192 columns, hi + lo (96 KB)
constexpr int FFN_FF_BYTES = A_STAGE;                  // one 128 x 64 GEGLU chunk, hi + lo (32 KB); doubles as the final epilogue's staging
constexpr int FFN_W1_PLANE = 256 * HROW;               // 32 KB: hi or lo plane of one k-chunk of a W1 pair (256 interleaved columns)
constexpr int FFN_W2_PLANE = FFN_C * HROW;             // 24 KB: hi or lo plane of one k-chunk of W2 (192 columns)
constexpr int FFN_W_SLOT = FFN_W1_PLANE;
constexpr int FFN_W_SLOTS = 3;
constexpr size_t FFN_SMEM_BYTES = (size_t)FFN_A_BYTES + FFN_FF_BYTES + (size_t)FFN_W_SLOTS * FFN_W_SLOT + 1024 + 256;
static_assert(FFN_SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int FFN_ACC4_COL = 0;            // 256-column accumulator of the GEGLU GEMM (one chunk pair)
constexpr int FFN_ACC5_COL = 256;          // 192-column accumulator of the output GEMM
constexpr int FFN_FFT_COL = 448;           // odd chunks as a tensor-memory operand: 32 columns of packed fp16 hi, then 32 of lo

SAID_DEVINL void tmem_st4u(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct FfnParams {
    CUtensorMap map_ln;      // LN3(x2) pair tensor (Mp, [hi 192 | lo 192])
    CUtensorMap map_x2;      // x2 pair tensor
    int M;
    float scale1;            // 2^-exp of the W1 image
    float* part;             // leftover tiles split across CTAs: partial output accumulators (leftover tile, piece, 128, 192) fp32
    int* sync;               // ... and their arrival / completion counters (2 x 32 ints, zero between launches)
    int split;               // 0: whole tiles only
    long long* trace;        // diagnostics (dbg & 16): clock64 stamps of CTA 0's second tile
    int dbg;                 // diagnostics (said_op_gemm_h_bench): 1 no activation loads, 2 no weight copies, 4 no epilogue work, 8 no MMAs
};

template <class EP>
__global__ void __launch_bounds__(FFN_THREADS, 1)
ffn_h_kernel(const __grid_constant__ FfnParams p, const uint8_t* __restrict__ W1img /*bn 256: 6 n-tiles x 3 k-chunks x {hi, lo}*/,
             const uint8_t* __restrict__ W2img /*bn 192: 15 k-chunks x {hi, lo}*/, const float* __restrict__ bias1 /*(1536) interleaved*/,
             int* __restrict__ flag, EP ep) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_buf = smem_base;
    const uint32_t ff_buf = a_buf + FFN_A_BYTES;
    const uint32_t w_base = ff_buf + FFN_FF_BYTES;
    const uint32_t bar_base = w_base + FFN_W_SLOTS * FFN_W_SLOT;
    const uint32_t bar_a_full = bar_base, bar_a_mid = bar_base + 8, bar_a2_full = bar_base + 16, bar_a_empty = bar_base + 24;
    auto bar_w_full = [&](int s) { return bar_base + 32u + 8u * s; };     // up to 4 slots
    auto bar_w_empty = [&](int s) { return bar_base + 64u + 8u * s; };
    const uint32_t bar_acc4_full = bar_base + 96, bar_acc4_empty = bar_base + 104;
    const uint32_t bar_acc5_full = bar_base + 128, bar_acc5_empty = bar_base + 136;
    // chunk buffers, one per chunk parity / epilogue group ([0]: shared memory, [1]: tensor memory).  mbarrier waits are by phase
    // parity, so a waiter must never be two phases ahead of its barrier: a group reaches pair k's wait only after the GEMM of pair
    // k, which is queued behind the output GEMMs of pair k - 2.
    auto bar_ff_full = [&](int g) { return bar_base + 112u + 8u * g; };
    auto bar_ff_empty = [&](int g) { return bar_base + 144u + 8u * g; };
    const uint32_t tmem_slot = bar_base + 160;
    auto w_st = [&](int s) { return w_base + (uint32_t)s * FFN_W_SLOT; };

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);    // provably warp-uniform: role branches do not diverge
    const int total_tiles = (p.M + HBM - 1) / HBM;
    // Work items.  Whole tiles go round-robin.  When the last round would leave most SMs idle (batch 64: 301 tiles on 148 SMs),
    // each leftover tile is cut into its 6 chunk pairs, handed to the lowest-numbered CTAs as their last item: a piece computes
    // one pair's contribution to the output accumulator (the last piece also the x2 . Wp term), parks it in global memory, and once
    // all 6 have arrived reduces one 32-column slice of the tile in a fixed order (results do not depend on timing).
    const int rounds = total_tiles / (int)gridDim.x, left = total_tiles - rounds * (int)gridDim.x;
    const bool split = p.split && rounds >= 1 && left > 0 && left * FFN_NP <= (int)gridDim.x && left <= 32;
    const int my_tiles = split ? rounds + ((int)blockIdx.x < left * FFN_NP ? 1 : 0)
                               : (((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);
    struct Item { int mt, pr0, npr, x2, lt, piece; };      // tile; first chunk pair, pairs; with the x2 term; leftover index, piece (-1: whole tile)
    auto item = [&](int i) -> Item {
        if (split && i >= rounds) {
            const int lt = (int)blockIdx.x / FFN_NP, pc = (int)blockIdx.x % FFN_NP;
            return Item{rounds * (int)gridDim.x + lt, pc, 1, pc == FFN_NP - 1 ? 1 : 0, lt, pc};
        }
        return Item{(int)blockIdx.x + i * (int)gridDim.x, 0, FFN_NP, 1, 0, -1};
    };

    if (tid == FFN_EPI_WARPS * 32) {
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_mid, 1);
        mbar_init(bar_a2_full, 1);
        mbar_init(bar_a_empty, 1);
        for (int s = 0; s < FFN_W_SLOTS; ++s) {
            mbar_init(bar_w_full(s), 1);
            mbar_init(bar_w_empty(s), 1);
        }
        mbar_init(bar_acc4_full, 1);
        mbar_init(bar_acc4_empty, FFN_EPI_WARPS * 32);
        mbar_init(bar_ff_full(0), FFN_EPI_WARPS * 16);   // one group of 8 warps per chunk
        mbar_init(bar_ff_full(1), FFN_EPI_WARPS * 16);
        mbar_init(bar_ff_empty(0), 1);
        mbar_init(bar_ff_empty(1), 1);
        mbar_init(bar_acc5_full, 1);
        mbar_init(bar_acc5_empty, FFN_EPI_WARPS * 32);
        fence_mbar_init();
    }
    if (tid == (FFN_EPI_WARPS + 1) * 32) {
        tma_prefetch_desc(&p.map_ln);
        tma_prefetch_desc(&p.map_x2);
    }
    __syncwarp();
    if (warp == FFN_EPI_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != FFN_EPI_WARPS + 2) pdl_wait();
    pdl_trigger();
    const bool trace_cta = (p.dbg & 16) && blockIdx.x == 0;

    if (warp < FFN_EPI_WARPS) {
        // ===================== epilogue warps: GEGLU chunks, then the tile's output =====================
        const int q = warp & 3, part4 = warp >> 2;           // TMEM lane quarter; which chunks of the output accumulator
        const int grp = warp >> 3, part = (warp >> 2) & 1;   // GEGLU: group g takes the g-th chunk of every pair (acc4 columns 128 g ..)
        const int row = q * 32 + lane;                       // row of the tile = TMEM lane
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const int lr = lane >> 2, lq = lane & 3;             // final epilogue, after the transpose: rows lr + 8 i, float4 column lq
        const uint32_t stg = ff_buf + (uint32_t)warp * (32 * 16 * 4);
        const bool has_res = ep.tc_has_res() && !(p.dbg & 4);
        constexpr int NCH = FFN_C / 16;                      // 12 output chunks of 16 columns, 3 per warp
        constexpr int MYCH = NCH / 4;
        uint32_t kp = 0;                                     // running pair index
        for (int i = 0; i < my_tiles; ++i) {
            const Item it = item(i);
            const int mt = it.mt;
            float amax = 0.f;
            for (int pr = it.pr0; pr < it.pr0 + it.npr; ++pr, ++kp) {
                const int j = 2 * pr + grp;                  // this group's chunk of the pair
                tc::mbar_wait_tight(bar_acc4_full, kp & 1u);
                tc_fence_after();
                const bool tr = trace_cta && i == 1 && (warp & 7) == 0 && lane == 0;
                if (tr) p.trace[64 + j * 5 + 0] = clock64();
                if (p.dbg & 4) {
                    tc_fence_before();
                    mbar_arrive(bar_acc4_empty);
                    tc::mbar_wait_tight(bar_ff_empty(grp), (kp & 1u) ^ 1u);
                    mbar_arrive(bar_ff_full(grp));
                    continue;
                }
                // this thread's 4 chunks of 16 accumulator columns: 16 (part + 2 u), u = 0..3, within the group's 128 columns
                float v[4][16];
#pragma unroll
                for (int u = 0; u < 4; ++u) tc::tmem_ld16_issue(trow + FFN_ACC4_COL + grp * 128 + (part + 2 * u) * 16, v[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) tc::tmem_ld_wait16(v[u]);
                tc_fence_before();
                mbar_arrive(bar_acc4_empty);                 // the accumulator is in registers: the next pair's GEMM may overwrite it
                if (tr) p.trace[64 + j * 5 + 1] = clock64();
                uint4 hi[4], lo[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float* bp = bias1 + j * 128 + (part + 2 * u) * 16;    // value / gate interleaved
                    float o[8];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 bb = ldg4(bp + c4 * 4);
                        o[2 * c4] = fmaf(v[u][4 * c4], p.scale1, bb.x) * gelu_erf_fast(fmaf(v[u][4 * c4 + 1], p.scale1, bb.y));
                        o[2 * c4 + 1] = fmaf(v[u][4 * c4 + 2], p.scale1, bb.z) * gelu_erf_fast(fmaf(v[u][4 * c4 + 3], p.scale1, bb.w));
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(o[e]));
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 hh = __floats2half2_rn(o[2 * e], o[2 * e + 1]);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(o[2 * e] - hf.x, o[2 * e + 1] - hf.y);
                        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        l[e] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    hi[u] = make_uint4(h[0], h[1], h[2], h[3]);
                    lo[u] = make_uint4(l[0], l[1], l[2], l[3]);
                }
                // the group's chunk buffer is free once the output GEMM of its previous chunk has read it
                if (tr) p.trace[64 + j * 5 + 2] = clock64();
                tc::mbar_wait_tight(bar_ff_empty(grp), (kp & 1u) ^ 1u);
                if (tr) p.trace[64 + j * 5 + 3] = clock64();
                if (grp == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = part + 2 * u;          // 16-byte slot of the 128-byte operand row: hidden columns 8 c .. 8 c + 7
                        const uint32_t a = ff_buf + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(hi[u].x), "r"(hi[u].y), "r"(hi[u].z), "r"(hi[u].w) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + A_PLANE), "r"(lo[u].x), "r"(lo[u].y), "r"(lo[u].z), "r"(lo[u].w) : "memory");
                    }
                    tc::fence_proxy_async();
                } else {
                    tc_fence_after();
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = part + 2 * u;          // hidden columns 8 c .. 8 c + 7 = packed columns 4 c .. 4 c + 3
                        tmem_st4u(trow + FFN_FFT_COL + 4 * c, hi[u].x, hi[u].y, hi[u].z, hi[u].w);
                        tmem_st4u(trow + FFN_FFT_COL + 32 + 4 * c, lo[u].x, lo[u].y, lo[u].z, lo[u].w);
                    }
                    tc::tmem_wait_st();
                    tc_fence_before();
                }
                mbar_arrive(bar_ff_full(grp));
                if (tr) p.trace[64 + j * 5 + 4] = clock64();
            }
            if (amax > P16_LIMIT) atomicOr(flag, 1);
            // ---------------- the tile's output: acc5 (+ bias + residual) -> global, through the transposing staging ----------------
            const bool piece = it.piece >= 0;
            const int mrow0 = mt * HBM + q * 32 + lr;
            typename EP::RowCtx rc[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) rc[ii] = ep.tc_row(mrow0 + 8 * ii, p.M);
            float4 pf[MYCH][4];                              // the residual rows: in flight while the last MMAs run
#pragma unroll
            for (int jj = 0; jj < MYCH; ++jj)
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
                    pf[jj][ii] = (has_res && !piece) ? ep.tc_prefetch4(rc[ii], (4 * jj + part4) * 16 + lq * 4) : zero4();
            float* const ppart = piece ? p.part + ((size_t)(it.lt * FFN_NP + it.piece) * HBM) * FFN_C : nullptr;
            if (trace_cta && i == 1 && tid == 0) p.trace[130] = clock64();
            mbar_wait(bar_acc5_full, (uint32_t)i & 1u);
            tc_fence_after();
            if (trace_cta && i == 1 && tid == 0) p.trace[131] = clock64();
            // (all MMAs of the tile have completed: the chunk buffer is idle and serves as the staging area)
#pragma unroll
            for (int jj = 0; jj < MYCH; ++jj) {
                if (p.dbg & 4) break;
                const int jc = 4 * jj + part4;
                const int n = jc * 16 + lq * 4;
                const typename EP::ColCtx cc = ep.tc_col(n);
                float v[16];
                tmem_ld16(trow + FFN_ACC5_COL + jc * 16, v);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const uint32_t a = stg + (uint32_t)lane * 64u + (uint32_t)((c4 ^ ((lane >> 1) & 3)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[4 * c4]), "f"(v[4 * c4 + 1]),
                                 "f"(v[4 * c4 + 2]), "f"(v[4 * c4 + 3]) : "memory");
                }
                __syncwarp();
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int r = lr + 8 * ii;
                    const uint32_t a = stg + (uint32_t)r * 64u + (uint32_t)((lq ^ ((r >> 1) & 3)) << 4);
                    float4 acc;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "r"(a));
                    const int m = mrow0 + 8 * ii;
                    if (piece) st4(ppart + (size_t)(q * 32 + r) * FFN_C + n, acc);            // raw accumulator; scaled when reduced
                    else if (m < p.M) ep.store4(rc[ii], cc, m, n, acc, pf[jj][ii]);
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(bar_acc5_empty);
            if (trace_cta && i == 1 && tid == 0) p.trace[132] = clock64();
            if (piece && !(p.dbg & 4)) {
                // all six pieces of the tile are resident (they are the last items of the six lowest CTAs of the tile's group, and
                // the grid never exceeds the SM count): wait for them, then reduce this piece's 32-column slice in piece order
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(FFN_EPI_WARPS * 32) : "memory");
                if (tid == 0) {
                    atomicAdd(p.sync + it.lt, 1);
                    uint32_t spins = 0;
                    while (*reinterpret_cast<volatile int*>(p.sync + it.lt) < FFN_NP) {
                        if (++spins > (1u << 24)) __trap();
                        __nanosleep(100);
                    }
                    __threadfence();
                }
                asm volatile("bar.sync 1, %0;" ::"n"(FFN_EPI_WARPS * 32) : "memory");
                const int rr = tid >> 2, m = mt * HBM + rr;
                if (m < p.M) {
                    const typename EP::RowCtx rcx = ep.tc_row(m, p.M);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int n = 32 * it.piece + (tid & 3) * 8 + 4 * hh;
                        const float4 res = has_res ? ep.tc_prefetch4(rcx, n) : zero4();
                        float4 acc = zero4();
                        for (int k = 0; k < FFN_NP; ++k) {
                            const float4 t = __ldcg(reinterpret_cast<const float4*>(p.part + ((size_t)(it.lt * FFN_NP + k) * HBM + rr) * FFN_C + n));
                            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                        }
                        ep.store4(rcx, ep.tc_col(n), m, n, acc, res);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(FFN_EPI_WARPS * 32) : "memory");
                if (tid == 0 && atomicAdd(p.sync + 32 + it.lt, 1) == FFN_NP - 1) {   // last one out re-arms the counters
                    p.sync[it.lt] = 0;
                    p.sync[32 + it.lt] = 0;
                }
            }
            // the staging area becomes the next tile's chunk buffer: every epilogue warp must be done with it
            asm volatile("bar.sync 1, %0;" ::"n"(FFN_EPI_WARPS * 32) : "memory");
        }
    } else if (warp == FFN_EPI_WARPS) {
        // ===================== MMA issuer =====================
        // The whole warp runs this code, so that loop state, slot indices and descriptors are warp-uniform and live in uniform
        // registers; one elected lane issues.  (Issued from a single-lane branch, every tcgen05.mma was preceded by register ->
        // uniform-register moves and a divergence loop: ~150 cycles of issue per MMA against 96..128 of tensor work.)
        {
            const uint32_t id256 = make_idesc_f16(HBM, 256), id192 = make_idesc_f16(HBM, FFN_C);
            const bool do_mma = !(p.dbg & 8);
            int sw = 0;
            uint32_t pw = 0, n_ff = 0, n_acc4 = 0, n_x2 = 0;
            auto next_w = [&]() { if (++sw == FFN_W_SLOTS) { sw = 0; pw ^= 1u; } };
            // acc (+)= A(128 x 64: hi at a, lo at a + A_PLANE) . W(k-chunk: its hi plane, then its lo plane, from the ring)
            auto chunk_mma = [&](uint32_t tacc, uint32_t a, uint32_t idesc, bool first) {
                const uint64_t dah = make_desc(a), dal = make_desc(a + A_PLANE);
                mbar_wait(bar_w_full(sw), pw);
                tc_fence_after();
                const uint64_t dbh = make_desc(w_st(sw));
                if (elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int k4 = 0; k4 < HBK / 16; ++k4) {
                            const uint64_t adv = (uint64_t)(k4 * 2);    // 16 fp16 = 32 bytes = 2 x 16-byte units along K
                            mma_f16(tacc, dah + adv, dbh + adv, idesc, (first && k4 == 0) ? 0u : 1u);
                            mma_f16(tacc, dal + adv, dbh + adv, idesc, 1u);
                        }
                    }
                    mma_commit(bar_w_empty(sw));
                }
                __syncwarp();
                next_w();
                mbar_wait(bar_w_full(sw), pw);
                tc_fence_after();
                const uint64_t dbl = make_desc(w_st(sw));
                if (elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int k4 = 0; k4 < HBK / 16; ++k4) mma_f16(tacc, dah + (uint64_t)(k4 * 2), dbl + (uint64_t)(k4 * 2), idesc, 1u);
                    }
                    mma_commit(bar_w_empty(sw));
                }
                __syncwarp();
                next_w();
            };
            // the same with the activation operand in tensor memory (packed fp16: 8 columns per 16-element k-step; hi, then lo 32 columns on)
            auto chunk_mma_ts = [&](uint32_t tacc, uint32_t ta, uint32_t idesc) {
                mbar_wait(bar_w_full(sw), pw);
                tc_fence_after();
                const uint64_t dbh = make_desc(w_st(sw));
                if (elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int k4 = 0; k4 < HBK / 16; ++k4) {
                            const uint64_t adv = (uint64_t)(k4 * 2);
                            mma_f16_ts(tacc, ta + 8 * k4, dbh + adv, idesc, 1u);
                            mma_f16_ts(tacc, ta + 32 + 8 * k4, dbh + adv, idesc, 1u);
                        }
                    }
                    mma_commit(bar_w_empty(sw));
                }
                __syncwarp();
                next_w();
                mbar_wait(bar_w_full(sw), pw);
                tc_fence_after();
                const uint64_t dbl = make_desc(w_st(sw));
                if (elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int k4 = 0; k4 < HBK / 16; ++k4) mma_f16_ts(tacc, ta + 8 * k4, dbl + (uint64_t)(k4 * 2), idesc, 1u);
                    }
                    mma_commit(bar_w_empty(sw));
                }
                __syncwarp();
                next_w();
            };
            auto commit = [&](uint32_t bar) {
                if (elect_one()) mma_commit(bar);
                __syncwarp();
            };
            auto g4 = [&]() {            // GEGLU GEMM of the next chunk pair
                mbar_wait(bar_acc4_empty, (n_acc4 & 1u) ^ 1u);
                ++n_acc4;
                tc_fence_after();
                for (int kc = 0; kc < 3; ++kc) chunk_mma(tmem_base + FFN_ACC4_COL, a_buf + (uint32_t)kc * A_STAGE, id256, kc == 0);
                commit(bar_acc4_full);
            };
            for (int i = 0; i < my_tiles; ++i) {
                const Item it = item(i);
                const int pr_end = it.pr0 + it.npr;
                mbar_wait(bar_a_full, (uint32_t)i & 1u);
                tc_fence_after();
                const bool tr = trace_cta && i == 1 && lane == 0;
                if (tr) p.trace[60] = clock64();
                g4();
                if (it.npr == 1 && it.x2) commit(bar_a_mid);
                if (tr) p.trace[61] = clock64();
                for (int pr = it.pr0; pr < pr_end; ++pr) {
                    if (pr + 1 < pr_end) {
                        g4();
                        if (pr + 2 == pr_end && it.x2) commit(bar_a_mid);   // the LayerNorm-ed rows are no longer needed: x2 may be loaded over them
                    }
                    if (pr == it.pr0) {                      // the previous tile's output accumulator must have been drained
                        mbar_wait(bar_acc5_empty, ((uint32_t)i & 1u) ^ 1u);
                        tc_fence_after();
                    }
                    for (int g = 0; g < 2; ++g) {
                        const int j = 2 * pr + g;
                        if (tr) p.trace[j * 4 + 0] = clock64();
                        tc::mbar_wait_tight(bar_ff_full(g), n_ff & 1u);
                        tc_fence_after();
                        if (tr) p.trace[j * 4 + 1] = clock64();
                        if (g == 0) chunk_mma(tmem_base + FFN_ACC5_COL, ff_buf, id192, pr == it.pr0);
                        else chunk_mma_ts(tmem_base + FFN_ACC5_COL, tmem_base + FFN_FFT_COL, id192);
                        commit(bar_ff_empty(g));
                        if (tr) p.trace[j * 4 + 2] = clock64();
                    }
                    ++n_ff;
                }
                if (it.x2) {
                    mbar_wait(bar_a2_full, n_x2 & 1u);
                    ++n_x2;
                    tc_fence_after();
                    for (int kc = 0; kc < 3; ++kc) chunk_mma(tmem_base + FFN_ACC5_COL, a_buf + (uint32_t)kc * A_STAGE, id192, false);
                }
                commit(bar_a_empty);
                commit(bar_acc5_full);
                if (tr) p.trace[62] = clock64();
            }
        }
        __syncwarp();
    } else if (warp == FFN_EPI_WARPS + 1) {
        // ===================== activation tiles: TMA =====================
        if (lane == 0) {
            uint32_t n_x2 = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const Item it = item(i);
                const int m0 = it.mt * HBM;
                mbar_wait(bar_a_empty, ((uint32_t)i & 1u) ^ 1u);
                if (p.dbg & 1) {
                    mbar_arrive(bar_a_full);
                    if (it.x2) {
                        mbar_wait(bar_a_mid, n_x2 & 1u);
                        ++n_x2;
                        mbar_arrive(bar_a2_full);
                    }
                    continue;
                }
                mbar_arrive_expect_tx(bar_a_full, (uint32_t)FFN_A_BYTES);
                for (int c = 0; c < 3; ++c) {
                    tma_load_2d(a_buf + c * A_STAGE, &p.map_ln, c * HBK, m0, bar_a_full);
                    tma_load_2d(a_buf + c * A_STAGE + A_PLANE, &p.map_ln, FFN_C + c * HBK, m0, bar_a_full);
                }
                if (!it.x2) continue;
                mbar_wait(bar_a_mid, n_x2 & 1u);
                ++n_x2;
                mbar_arrive_expect_tx(bar_a2_full, (uint32_t)FFN_A_BYTES);
                for (int c = 0; c < 3; ++c) {
                    tma_load_2d(a_buf + c * A_STAGE, &p.map_x2, c * HBK, m0, bar_a2_full);
                    tma_load_2d(a_buf + c * A_STAGE + A_PLANE, &p.map_x2, FFN_C + c * HBK, m0, bar_a2_full);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== weight planes: bulk copies, in the order the MMA warp consumes them =====================
        if (lane == 0) {
            int sw = 0;
            uint32_t pw = 0;
            auto put = [&](const uint8_t* src, uint32_t bytes) {
                mbar_wait(bar_w_empty(sw), pw ^ 1u);
                if (p.dbg & 2) {
                    mbar_arrive(bar_w_full(sw));
                } else {
                    mbar_arrive_expect_tx(bar_w_full(sw), bytes);
                    tc::bulk_g2s(w_st(sw), src, bytes, bar_w_full(sw));
                }
                if (++sw == FFN_W_SLOTS) { sw = 0; pw ^= 1u; }
            };
            auto w1 = [&](int pr) {      // 3 k-chunks x {hi, lo} of chunk pair pr
                for (int pl = 0; pl < 6; ++pl) put(W1img + ((size_t)pr * 6 + pl) * FFN_W1_PLANE, FFN_W1_PLANE);
            };
            auto w2 = [&](int kc) {      // {hi, lo} of k-chunk kc
                for (int pl = 0; pl < 2; ++pl) put(W2img + ((size_t)kc * 2 + pl) * FFN_W2_PLANE, FFN_W2_PLANE);
            };
            for (int i = 0; i < my_tiles; ++i) {
                const Item it = item(i);
                const int pr_end = it.pr0 + it.npr;
                w1(it.pr0);
                for (int pr = it.pr0; pr < pr_end; ++pr) {
                    if (pr + 1 < pr_end) w1(pr + 1);
                    w2(2 * pr);
                    w2(2 * pr + 1);
                }
                if (it.x2)
                    for (int kc = 0; kc < 3; ++kc) w2(FFN_NJ + kc);
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FFN_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <class EP>
inline cudaError_t launch_ffn_h(cudaStream_t st, int num_sms, const FfnParams& p, const uint8_t* W1img, const uint8_t* W2img,
                                const float* bias1, int* flag, const EP& ep, bool pdl = false) {
    auto kern = ffn_h_kernel<EP>;
    cudaError_t e = tc::configure_once((const void*)kern, (int)FFN_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const int total_tiles = (p.M + HBM - 1) / HBM;
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    return launch_ex(kern, dim3(grid), dim3(FFN_THREADS), FFN_SMEM_BYTES, st, pdl, 1, p, W1img, W2img, bias1, flag, ep);
}

}  // namespace hx
}  // namespace said
