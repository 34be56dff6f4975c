// tcgen05 GEMM over PRE-SPLIT fp16 operand pairs, fed by TMA -- the contraction engine of the step loop ("fp16x3" mode).
//
//   out[m, n] = epilogue( sum_k  A[m + shift(k), k] * W[n, k] )
//
// Operand format ("P16", pair tensor): every activation that feeds a contraction is stored ONCE, by the kernel that produces
// it, as two fp16 planes per row -- row r = [hi(C) | lo(C)], hi = fp16(x), lo = fp16(x - hi) -- the same 4 bytes per element as
// fp32, ~22 significant bits.  The product is formed as hi*hi + lo*hi + hi*lo on the tensor cores (kind::f16, fp32 accumulate in
// TMEM): the precision of the 3xTF32 scheme (gemm_tc.cuh) at twice its MMA rate and half its shared-memory traffic, and the
// hi/lo split is paid once per element instead of once per consuming tile (x3 conv taps, x3..8 output tiles).
// Weights are pre-split on the host the same way, pre-scaled by a power of two per matrix so that the lo plane stays in fp16's
// normal range (the epilogue multiplies the accumulator by the inverse power of two: exact).
//
// Because the operands already have their final shared-memory form, there are no producer warps:
//   * warp 9 (one thread) issues cp.async.bulk.tensor (TMA, SASS UTMALDG) loads of 128-row x 64-column boxes of the hi and lo
//     planes straight into a ring of SWIZZLE_128B stages; a Conv1d(k=3) is three row-shifted boxes of the same tensor (the
//     activations keep one zero row between clips, so a tile may straddle clips), a channel concat is boxes from two tensor maps;
//     out-of-range rows are zero-filled by the TMA unit;
//   * warp 10 (one thread) streams the pre-swizzled weight tile images with cp.async.bulk (UBLKCP);
//   * warp 8 (one thread) issues tcgen05.mma.kind::f16 (M = 128, N <= 192, K = 16) into one of two TMEM accumulators;
//   * warps 0-7 drain the other accumulator (tcgen05.ld), transpose through shared memory and run the fused epilogue.
// Persistent, one CTA per SM; leftover tiles are cut into N-slivers for idle CTAs (as in gemm_tc.cuh).
#pragma once
#include <cuda.h>        // CUtensorMap + enums (types only; cuTensorMapEncodeTiled is fetched through cudart at run time)
#include <cuda_fp16.h>

#include "gemm_tc.cuh"
#include "pair.cuh"

namespace said {
namespace hx {

using tc::elect_one;
using tc::fence_mbar_init;
using tc::make_desc;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mma_commit;
using tc::smem_u32;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_alloc;
using tc::tmem_dealloc;
using tc::tmem_ld16;

constexpr int HBM = 128;                    // rows per tile (UMMA M)
constexpr int HBK = 64;                     // fp16 elements per stage row: 128 bytes = one SWIZZLE_128B row
constexpr int HROW = 128;                   // bytes per stage row
constexpr int A_PLANE = HBM * HROW;         // 16 KB: one plane (hi or lo) of an activation stage
constexpr int A_STAGE = 2 * A_PLANE;
constexpr int H_EPI_WARPS = 16;                     // 4 per TMEM lane quarter: the epilogue is latency-bound (TMEM -> smem -> global)
constexpr int H_EPI_PARTS = H_EPI_WARPS / 4;         // warps sharing a lane quarter split the 16-column chunks round-robin
constexpr int H_THREADS = (H_EPI_WARPS + 3) * 32;   // + MMA warp, activation-TMA warp, weight-copy warp
constexpr int H_MAX_SEG = 4;
static_assert(tc::ROW_BYTES == HROW, "make_desc() assumes 128-byte rows");

// ------------------------------------------------------------------------------------------------ PTX wrappers
SAID_DEVINL uint32_t make_idesc_f16(int M, int N) {   // F16 x F16 -> F32, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
SAID_DEVINL void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
SAID_DEVINL void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
SAID_DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------------------------------------------------------------------ kernel parameters
// The contraction dimension is a sequence of 64-column chunks described by up to four segments; chunk j of a segment loads
// columns [col0 + 64 j, +64) of rows [m0 + row_shift, +128) of tensor `map` (hi plane) and the same box `lo_off` columns to the
// right (lo plane).  Weight images follow the same chunk order.
struct HSeg { int map, nchunks, col0, lo_off, row_shift; };
struct HParams {
    CUtensorMap maps[3];
    int M, N, nk, nseg, nmaps;
    HSeg seg[H_MAX_SEG];
    int w_block_bytes;       // bytes between consecutive (n-tile, k-chunk) blocks of the weight image (2 * BN * 128)
    int w_nk_total, w_kc0;   // the image holds w_nk_total chunks per n-tile; this launch contracts chunks [w_kc0, w_kc0 + nk) of them
                             // (split-K: long contractions run as several launches that accumulate through the fp32 residual input,
                             //  which rounds to nearest, instead of one long chain of truncating tensor-core accumulations)
    int sliver;              // 0: leftover tiles are not cut into slivers
    int dbg;                 // diagnostics (said_op_gemm_h_bench): 1 no activation loads, 2 no weight copies, 4 no epilogue I/O, 8 no MMAs
};

template <int BN>
struct HCfg {
    static constexpr int B_PLANE = BN * HROW;
    static constexpr int B_STAGE = 2 * B_PLANE;
    static constexpr int A_STAGES = 3;
    static constexpr int B_STAGES = BN >= 192 ? 2 : (BN >= 128 ? 3 : 4);
    static constexpr int ACC_STRIDE = 256;
    static constexpr int TMEM_COLS = 512;
    static constexpr int EPI_STAGE_BYTES = H_EPI_WARPS * 32 * 16 * 4;
    static constexpr size_t SMEM_BYTES = (size_t)A_STAGES * A_STAGE + (size_t)B_STAGES * B_STAGE + EPI_STAGE_BYTES + 1024 + 256;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN, class EP>
__global__ void __launch_bounds__(H_THREADS, 1)
gemm_h_kernel(const __grid_constant__ HParams p, const uint8_t* __restrict__ Wimg, EP ep) {
    using Cfg = HCfg<BN>;
    constexpr int AS = Cfg::A_STAGES, BS = Cfg::B_STAGES;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + AS * A_STAGE;
    const uint32_t epi_base = b_base + BS * Cfg::B_STAGE;
    const uint32_t bar_base = epi_base + Cfg::EPI_STAGE_BYTES;
    auto fulla_bar = [&](int s) { return bar_base + 8u * s; };
    auto emptya_bar = [&](int s) { return bar_base + 8u * (AS + s); };
    auto fullb_bar = [&](int s) { return bar_base + 8u * (2 * AS + s); };
    auto emptyb_bar = [&](int s) { return bar_base + 8u * (2 * AS + BS + s); };
    auto accf_bar = [&](int b) { return bar_base + 8u * (2 * AS + 2 * BS + b); };
    auto acce_bar = [&](int b) { return bar_base + 8u * (2 * AS + 2 * BS + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * AS + 2 * BS + 4);
    auto a_st = [&](int s) { return smem_base + s * A_STAGE; };
    auto b_st = [&](int s) { return b_base + s * Cfg::B_STAGE; };

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);    // provably warp-uniform: role branches do not diverge
    const int nk = p.nk;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int total_tiles = ((p.M + HBM - 1) / HBM) * n_tiles;
    const int tq = total_tiles / (int)gridDim.x, tr = total_tiles % (int)gridDim.x;
    // tail balancing (see gemm_tc.cuh): each leftover tile is cut along N into S slivers handed to different CTAs as their last item
    int sl_S = 1;
    if (BN == 192 && tq >= 1 && tr > 0 && p.sliver) {
        const int cand[5] = {12, 6, 4, 3, 2};
        for (int k = 0; k < 5; ++k)
            if (cand[k] * tr <= (int)gridDim.x) { sl_S = cand[k]; break; }
    }
    const bool sliver_mode = sl_S > 1;
    const int my_tiles = sliver_mode ? tq + ((int)blockIdx.x < tr * sl_S ? 1 : 0) : tq + ((int)blockIdx.x < tr ? 1 : 0);
    const int tile0 = sliver_mode ? (int)blockIdx.x * tq : (int)blockIdx.x * tq + min((int)blockIdx.x, tr);
    struct Item { int mt, nt, n0, nw; };
    auto item = [&](int i) -> Item {
        if (sliver_mode && i >= tq) {
            const int s = (int)blockIdx.x;
            const int tile = (int)gridDim.x * tq + s / sl_S, mt = tile / n_tiles;
            return Item{mt, tile - mt * n_tiles, (s % sl_S) * (BN / sl_S), BN / sl_S};
        }
        const int tile = tile0 + i;
        const int mt = tile / n_tiles;
        return Item{mt, tile - mt * n_tiles, 0, BN};
    };

    if (tid == H_EPI_WARPS * 32) {
        for (int s = 0; s < AS; ++s) {
            mbar_init(fulla_bar(s), 1);                   // the TMA thread's arrive.expect_tx (+ tx bytes of both planes)
            mbar_init(emptya_bar(s), 1);                  // one tcgen05.commit
        }
        for (int s = 0; s < BS; ++s) {
            mbar_init(fullb_bar(s), 1);
            mbar_init(emptyb_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);                    // tcgen05.commit after the tile's last MMA
            mbar_init(acce_bar(b), H_EPI_WARPS * 32);     // every epilogue thread after draining
        }
        fence_mbar_init();
    }
    if (tid == (H_EPI_WARPS + 1) * 32) {
        for (int i = 0; i < p.nmaps; ++i) tma_prefetch_desc(&p.maps[i]);
    }
    __syncwarp();
    if (warp == H_EPI_WARPS) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (warp != H_EPI_WARPS + 2) pdl_wait();     // weights are constants: their copies may start under the previous kernel's tail
    pdl_trigger();

    if (warp < H_EPI_WARPS) {
        // ===================== epilogue (same structure as gemm_tc_kernel) =====================
        constexpr int NCH = BN / 16;
        constexpr int MYCH = (NCH + H_EPI_PARTS - 1) / H_EPI_PARTS;
        constexpr int EPI_PF = MYCH < 2 ? MYCH : 2;
        const int q = warp & 3, half = warp >> 2;          // lane quarter; part of the chunk round-robin
        const int lr = lane >> 2, lq = lane & 3;
        const uint32_t stg = epi_base + (uint32_t)warp * (32 * 16 * 4);
        const bool has_res = ep.tc_has_res() && !(p.dbg & 4);
        float4 pf[EPI_PF][4];
#pragma unroll
        for (int jj = 0; jj < EPI_PF; ++jj)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) pf[jj][ii] = zero4();
        for (int i = 0; i < my_tiles; ++i) {
            const Item it_ = item(i);
            const int mt = it_.mt, nt = it_.nt;
            const int nch_i = it_.nw / 16;
            const int ncol0 = nt * BN + it_.n0;
            const int buf = i & 1;
            const int mrow0 = mt * HBM + q * 32 + lr;
            typename EP::RowCtx rc[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) rc[ii] = ep.tc_row(mrow0 + 8 * ii, p.M);
            if (has_res) {
#pragma unroll
                for (int jj = 0; jj < EPI_PF; ++jj) {
                    const int j = H_EPI_PARTS * jj + half;
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) pf[jj][ii] = ep.tc_prefetch4(rc[ii], ncol0 + (j < nch_i ? j : 0) * 16 + lq * 4);
                }
            }
            mbar_wait(accf_bar(buf), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Cfg::ACC_STRIDE);
#pragma unroll
            for (int jj = 0; jj < MYCH; ++jj) {
                const int j = H_EPI_PARTS * jj + half;
                if (j < nch_i) {                           // warp-uniform
                    const typename EP::ColCtx cc = ep.tc_col(ncol0 + j * 16 + lq * 4);   // bias: in flight while the accumulator is read
                    float v[16];
                    tmem_ld16(taddr + j * 16, v);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const uint32_t a = stg + (uint32_t)lane * 64u + (uint32_t)((c4 ^ ((lane >> 1) & 3)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[4 * c4]), "f"(v[4 * c4 + 1]),
                                     "f"(v[4 * c4 + 2]), "f"(v[4 * c4 + 3]) : "memory");
                    }
                    __syncwarp();
                    const int n = ncol0 + j * 16 + lq * 4;
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const int r = lr + 8 * ii;
                        const uint32_t a = stg + (uint32_t)r * 64u + (uint32_t)((lq ^ ((r >> 1) & 3)) << 4);
                        float4 acc;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "r"(a));
                        const int m = mrow0 + 8 * ii;
                        if (m < p.M && !(p.dbg & 4)) ep.store4(rc[ii], cc, m, n, acc, pf[jj % EPI_PF][ii]);
                    }
                    __syncwarp();
                    const int jn = j + H_EPI_PARTS * EPI_PF;
                    if (has_res && jn < nch_i) {
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) pf[jj % EPI_PF][ii] = ep.tc_prefetch4(rc[ii], ncol0 + jn * 16 + lq * 4);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acce_bar(buf));
        }
    } else if (warp == H_EPI_WARPS) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (state and descriptors stay in uniform registers); one elected lane issues.  From a
        // single-lane branch every tcgen05.mma cost a divergence loop plus register -> uniform-register moves, which at N <= 192
        // (96 tensor cycles per MMA) made the issue stream the limiter (profiles/r2_mma_rate.md).
        {
            const bool do_mma = !(p.dbg & 8);
            int sa = 0, sb = 0, sa0 = 0;
            uint32_t pa = 0, pb = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                const Item it_ = item(i);
                // A-stationary: when the whole K extent fits the activation ring, consecutive n-tiles of one row block reuse the
                // stages the first of them loaded (GEGLU: 8 n-tiles, q/k/v: 3) instead of reloading them from L2
                const bool a_first = !(nk <= AS && i > 0 && item(i - 1).mt == it_.mt);
                const bool a_last = !(nk <= AS && i + 1 < my_tiles && item(i + 1).mt == it_.mt);
                if (a_first) sa0 = sa;
                const uint32_t idesc = make_idesc_f16(HBM, it_.nw);
                const uint32_t boff = (uint32_t)it_.n0 * HROW;               // n0 is a multiple of 16 rows: whole swizzle atoms
                mbar_wait(acce_bar(buf), ((uint32_t)(i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::ACC_STRIDE);
                for (int kc = 0; kc < nk; ++kc) {
                    int sa_use = sa;
                    if (a_first) {
                        mbar_wait(fulla_bar(sa), pa);
                    } else {
                        sa_use = sa0 + kc;
                        if (sa_use >= AS) sa_use -= AS;
                    }
                    mbar_wait(fullb_bar(sb), pb);
                    tc_fence_after();
                    const uint64_t dah = make_desc(a_st(sa_use)), dal = make_desc(a_st(sa_use) + A_PLANE);
                    const uint64_t dbh = make_desc(b_st(sb) + boff), dbl = make_desc(b_st(sb) + Cfg::B_PLANE + boff);
                    if (elect_one()) {
                        if (do_mma) {
#pragma unroll
                            for (int k4 = 0; k4 < HBK / 16; ++k4) {
                                const uint64_t adv = (uint64_t)(k4 * 2);   // 16 fp16 = 32 bytes = 2 x 16-byte units along K
                                mma_f16(tacc, dah + adv, dbh + adv, idesc, (kc | k4) != 0 ? 1u : 0u);
                                mma_f16(tacc, dal + adv, dbh + adv, idesc, 1u);
                                mma_f16(tacc, dah + adv, dbl + adv, idesc, 1u);
                            }
                        }
                        if (a_last) mma_commit(emptya_bar(sa_use));      // the stage is free once the LAST n-tile's MMAs have read it
                        mma_commit(emptyb_bar(sb));
                        if (kc == nk - 1) mma_commit(accf_bar(buf));
                    }
                    __syncwarp();
                    if (a_first && ++sa == AS) { sa = 0; pa ^= 1u; }
                    if (++sb == BS) { sb = 0; pb ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == H_EPI_WARPS + 1) {
        // ===================== activation tiles: TMA =====================
        if (lane == 0) {
            int sa = 0;
            uint32_t pa = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int m0 = item(i).mt * HBM;
                if (nk <= AS && i > 0 && item(i - 1).mt == item(i).mt) continue;   // A-stationary (see the MMA issuer)
                for (int s = 0; s < p.nseg; ++s) {
                    const HSeg sg = p.seg[s];
                    const CUtensorMap* map = &p.maps[sg.map];
                    for (int j = 0; j < sg.nchunks; ++j) {
                        mbar_wait(emptya_bar(sa), pa ^ 1u);
                        if (p.dbg & 1) {
                            mbar_arrive(fulla_bar(sa));
                        } else {
                            mbar_arrive_expect_tx(fulla_bar(sa), (uint32_t)A_STAGE);
                            tma_load_2d(a_st(sa), map, sg.col0 + j * HBK, m0 + sg.row_shift, fulla_bar(sa));
                            tma_load_2d(a_st(sa) + A_PLANE, map, sg.lo_off + sg.col0 + j * HBK, m0 + sg.row_shift, fulla_bar(sa));
                        }
                        if (++sa == AS) { sa = 0; pa ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== weight tile images: bulk copies =====================
        if (lane == 0) {
            constexpr uint32_t bytes = (uint32_t)Cfg::B_STAGE;
            int sb = 0;
            uint32_t pb = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const Item it_ = item(i);
                const uint8_t* wsrc = Wimg + ((size_t)it_.nt * p.w_nk_total + p.w_kc0) * p.w_block_bytes;
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(emptyb_bar(sb), pb ^ 1u);
                    const uint8_t* blk = wsrc + (size_t)kc * p.w_block_bytes;
                    if (p.dbg & 2) {
                        mbar_arrive(fullb_bar(sb));
                    } else if (it_.nw == BN) {
                        mbar_arrive_expect_tx(fullb_bar(sb), bytes);
                        tc::bulk_g2s(b_st(sb), blk, bytes, fullb_bar(sb));
                    } else {   // sliver: only its rows of the hi and lo planes
                        const uint32_t part = (uint32_t)it_.nw * HROW, roff = (uint32_t)it_.n0 * HROW;
                        mbar_arrive_expect_tx(fullb_bar(sb), 2u * part);
                        tc::bulk_g2s(b_st(sb) + roff, blk + roff, part, fullb_bar(sb));
                        tc::bulk_g2s(b_st(sb) + Cfg::B_PLANE + roff, blk + Cfg::B_PLANE + roff, part, fullb_bar(sb));
                    }
                    if (++sb == BS) { sb = 0; pb ^= 1u; }
                }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == H_EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
// Pack a K-major x N weight matrix Wt (K rows, ldw floats per row) into per-(n-tile, k-chunk) shared-memory images:
//   block(nt, kc) = [hi plane: BN rows x 128 B, SWIZZLE_128B] [lo plane], fp16 of w * 2^scale_exp.
// Returns the power-of-two exponent applied (the epilogue multiplies by 2^-scale_exp).
inline int pack_weights_h(const float* Wt, int K, int N, int ldw, int BN, std::vector<uint16_t>& out) {
    float amax = 0.f;
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) amax = std::max(amax, std::fabs(Wt[(size_t)k * ldw + n]));
    int e = 0;
    if (amax > 0.f) {
        int ex;
        std::frexp(amax, &ex);          // amax = f * 2^ex, f in [0.5, 1)
        e = 13 - ex;                    // scaled amax in [2^12, 2^13): two binades of head-room below fp16's 65504
        if (e > 24) e = 24;
        if (e < -24) e = -24;
    }
    const float sc = std::ldexp(1.0f, e);
    const int ntiles = (N + BN - 1) / BN, nk = K / HBK;
    const size_t plane = (size_t)BN * HBK, block = 2 * plane;
    out.assign((size_t)ntiles * nk * block, 0);
    for (int nt = 0; nt < ntiles; ++nt)
        for (int kc = 0; kc < nk; ++kc) {
            uint16_t* hi = out.data() + ((size_t)nt * nk + kc) * block;
            uint16_t* lo = hi + plane;
            for (int r = 0; r < BN; ++r) {
                const int n = nt * BN + r;
                if (n >= N) continue;
                for (int kk = 0; kk < HBK; ++kk) {
                    const float w = Wt[(size_t)(kc * HBK + kk) * ldw + n] * sc;
                    const __half h = __float2half_rn(w);
                    const __half l = __float2half_rn(w - __half2float(h));
                    const int chunk = kk >> 3;                       // 16-byte chunk (8 halves) of the 128-byte row
                    const int sw = (chunk ^ (r & 7)) & 7;
                    const size_t idx = (size_t)r * HBK + (size_t)sw * 8 + (kk & 7);
                    hi[idx] = __half_as_ushort(h);
                    lo[idx] = __half_as_ushort(l);
                }
            }
        }
    return e;
}

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tensorMapEncodeTiled tensor_map_encoder() {
    static PFN_tensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_tensorMapEncodeTiled)p;
    }
    return fn;
}
// Tensor map over a pair tensor: rows x [hi(C) | lo(C)] fp16, row pitch `pitch_halfs` (>= 2C; default 2C), boxes of 64 columns x 128 rows.
inline bool make_pair_map(CUtensorMap* m, const void* base, int C, long long rows, long long pitch_halfs = 0) {
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    if (pitch_halfs == 0) pitch_halfs = 2LL * C;
    cuuint64_t dims[2] = {(cuuint64_t)(2 * C), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_halfs * 2};
    cuuint32_t box[2] = {(cuuint32_t)HBK, (cuuint32_t)HBM};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, class EP>
inline cudaError_t launch_gemm_h(cudaStream_t st, int num_sms, const HParams& p, const uint8_t* Wimg, const EP& ep, bool pdl = false) {
    using Cfg = HCfg<BN>;
    auto kern = gemm_h_kernel<BN, EP>;
    cudaError_t e = tc::configure_once((const void*)kern, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const int total_tiles = ((p.M + HBM - 1) / HBM) * ((p.N + BN - 1) / BN);
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    return launch_ex(kern, dim3(grid), dim3(H_THREADS), Cfg::SMEM_BYTES, st, pdl, 1, p, Wimg, ep);
}

}  // namespace hx
}  // namespace said
