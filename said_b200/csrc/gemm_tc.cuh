// tcgen05 (5th-gen tensor core) GEMM with fused operand transforms and epilogues, sm_100a only.
//
//   out[m, n] = epilogue( sum_k  load(m, k) * W[n, k] )
//
// Same loaders / epilogues as gemm_simt.cuh (normalisation + activation folded into the A operand,
// bias / residual / GEGLU folded into the epilogue), but the contraction runs on the tensor cores.
// Persistent, one CTA per SM, 576 threads, warp-specialised (details at gemm_tc_kernel below):
//
//   * A (activations): 8 producer warps cp.async 16-byte chunks of fp32 straight into a 5-deep ring of "hi" stages in
//     the UMMA canonical K-major SWIZZLE_128B layout (32 fp32 = 128 B per row, 16-byte chunk index XOR row%8); the raw
//     fp32 IS the TF32 hi operand (the tensor core ignores the low 13 mantissa bits).  In place they apply the loader
//     transform (if any) and write the TF32 lo part (x - hi, 3 ALU instructions per value: tf32_split4) into a separate
//     2-deep lo ring, then fence.proxy.async + mbarrier arrive.  Four chunks of global loads stay in flight.
//   * B (weights): pre-split and pre-swizzled on the host into ready-to-use tile images, one contiguous
//     block per (n-tile, k-chunk); a single cp.async.bulk (TMA engine, UBLKCP) per stage lands it and
//     completes the stage's mbarrier transaction count.
//   * MMA: one elected thread issues tcgen05.mma.kind::tf32 (M=128, N<=BN, K=8) into one of two TMEM accumulators.
//     NSPLIT == 3 issues hi*hi + lo*hi + hi*lo ("3xTF32": fp32-level accuracy, which is what keeps
//     the path inside the reference's fp32 tolerance); NSPLIT == 1 issues hi*hi only.
//     tcgen05.commit releases the stages back to the producers and finally signals the epilogue.
//   * Epilogue: 8 warps drain the other accumulator (tcgen05.ld 32x32b.x16), transpose through shared memory for
//     sector-coalesced global I/O, apply the epilogue (residual prefetched two chunks ahead) and store.
//   * Scheduling: contiguous runs of tiles per CTA; the T % G leftover tiles are cut along N into slivers for idle CTAs.
//
// Measured history of the pipeline (what starved what, and the fixes): profiles/r1_gemm_pipeline.md.
#pragma once
#include <cstring>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "common.cuh"

namespace said {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                    // fp32 elements per stage row: 128 bytes = one SWIZZLE_128B row (16 -> SWIZZLE_64B)
constexpr int ROW_BYTES = BK * 4;
constexpr int ROW_CHUNKS = ROW_BYTES / 16;            // 16-byte chunks per row (4)
constexpr int A_TILE_BYTES = BM * ROW_BYTES;
// byte offset of 16-byte chunk `c` of row `r` inside a K-major swizzled tile whose base is 1024-byte aligned:
// the swizzle XORs the chunk index with address bits [7, 7 + log2(ROW_CHUNKS))
SAID_DEVINL uint32_t swz_off(int r, int c) {
    return (uint32_t)r * ROW_BYTES + (uint32_t)((c ^ ((r * ROW_BYTES) >> 7)) & (ROW_CHUNKS - 1)) * 16u;
}

SAID_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SAID_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
SAID_DEVINL void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SAID_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SAID_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking poll (try_wait may suspend the thread for a hardware-defined time)
SAID_DEVINL bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A wrong barrier protocol must surface as an error, not as a hung GPU: trap after ~seconds of spinning.
SAID_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
        if (spins > 8) __nanosleep(40);          // long waits (epilogue, idle roles) must not burn issue slots
    }
}
// Latency-critical hand-offs (attention: MMA <-> softmax every few hundred cycles): mbarrier.try_wait already suspends the thread in
// hardware until the phase completes or a time limit expires, so no __nanosleep (whose granularity is of the order of the waits).
SAID_DEVINL void mbar_wait_tight(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
SAID_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SAID_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SAID_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SAID_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

SAID_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
// Column count as an IMMEDIATE: with a register operand the tool chain cannot know how much tensor memory a CTA takes and admits
// one CTA per SM (measured: occupancy 1 for a 256-column kernel); with an immediate two 256-column CTAs can share an SM.
template <uint32_t NCOLS>
SAID_DEVINL void tmem_alloc_imm(uint32_t dst_smem) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
SAID_DEVINL void tmem_dealloc_imm(uint32_t taddr) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
SAID_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SAID_DEVINL void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, TF32 inputs, fp32 accumulate
SAID_DEVINL void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one lane of a converged warp (the issuing lane of tcgen05.mma / commit in warp-uniform code)
SAID_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
SAID_DEVINL void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
SAID_DEVINL void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form of the above: issue the load, do other work (or issue more loads), then wait.  The wait names the
// destination registers as read-write operands so the compiler cannot schedule their uses above it.
SAID_DEVINL void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
SAID_DEVINL void tmem_ld_wait16(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}

SAID_DEVINL void tmem_ld8_issue(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
}
SAID_DEVINL void tmem_ld_wait8(float (&v)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])
                 :
                 : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): K-major, rows of ROW_BYTES, hardware swizzle
// matching the row width (SWIZZLE_64B for 64-byte rows, SWIZZLE_128B for 128-byte rows), 8-row groups
// 8 * ROW_BYTES apart.
SAID_DEVINL uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);         // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;           // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (sm_100)
    d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;       // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): TF32 x TF32 -> F32, both K-major, M x N
SAID_DEVINL uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

SAID_DEVINL float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// TF32 hi / lo split of a finite fp32 value, 3 ALU instructions per value (cvt.rna.tf32.f32 alone is 4: add, inf/nan
// test, select, mask; the split with two of them is 9):
//   hi = round-to-nearest (ties away) to 10 mantissa bits, by integer add + mask on the bit pattern
//   lo = x - hi exactly (fp32 subtraction of nearby values is exact); the tensor core ignores the low 13 mantissa bits
//        of its TF32 inputs, so lo is effectively truncated to 11 significant bits: |error| <= 2^-10 |lo| <= 2^-21 |x|,
//        the same order as the dropped lo*lo term of the 3xTF32 product.
SAID_DEVINL float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
SAID_DEVINL void tf32_split4(const float4& x, float4& h, float4& l) {
    h.x = tf32_hi(x.x); h.y = tf32_hi(x.y); h.z = tf32_hi(x.z); h.w = tf32_hi(x.w);
    l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
}

SAID_DEVINL void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {   // src_bytes 0 -> zero fill
    // L2::256B: a miss brings the whole 256-byte segment of the row into L2, so HBM sees long bursts instead of
    // the 64/128-byte slices a K-chunked tile load would otherwise request
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
SAID_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SAID_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct TcDims {
    int M, N, K;
    int w_block_floats;   // floats between consecutive (n-tile, k-chunk) blocks of the weight image
    int dbg;              // diagnostics (said_op_gemm_tc_bench): 2 skip weight copies, 4 skip epilogue I/O, 8 skip MMAs
};

constexpr int EPI_WARPS = 8;              // warps 0-7: TMEM lane quarter = warp % 4, column half = warp / 4
constexpr int LOADER_WARPS = 8;
constexpr int LOADER_THREADS = LOADER_WARPS * 32;
constexpr int LOADER_TID0 = (EPI_WARPS + 2) * 32;   // warps 0-7 epilogue, 8 MMA, 9 weights, 10.. loaders
constexpr int THREADS2 = LOADER_TID0 + LOADER_THREADS;   // 576
constexpr int LROWS = BM * ROW_CHUNKS / LOADER_THREADS;   // 16-byte slots per loader thread per stage (4)
constexpr int LROW_STEP = LOADER_THREADS / ROW_CHUNKS;    // row distance between a thread's slots (32)

template <int BN, int NSPLIT>
struct TcCfg {
    static constexpr int NPARTS = NSPLIT == 3 ? 2 : 1;                 // hi (+ lo) copies of each operand
    static constexpr int B_TILE_BYTES = BN * ROW_BYTES;
    static constexpr int B_STAGE_BYTES = NPARTS * B_TILE_BYTES;
    // Activations: cp.async lands directly in a "hi" stage (raw fp32 = the TF32 hi operand); the lo tiles live in their
    // own, shorter ring because a lo tile only exists between the transform and the completion of its MMAs.  The hi
    // ring depth is what hides the global-load latency: item i + A_STAGES - 1 is issued when MMA(i - 1) completes, so
    // the latency budget is (A_STAGES - 2) chunk times of MMA (with 3 coupled hi+lo stages it was ONE, and the MMA
    // thread spent ~half of its time waiting for the producers: profiles/r1_gemm_pipeline.md).
    static constexpr int A_STAGES = NSPLIT == 3 ? 5 : 6;
    static constexpr int A_LO_STAGES = NSPLIT == 3 ? 2 : 0;
    static constexpr int B_STAGES = NSPLIT == 3 ? 2 : 4;               // weights: bulk copies
    static constexpr int ACC_STRIDE = 256;                             // TMEM columns between the two accumulators
    static constexpr int TMEM_COLS = 512;
    static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 16 * 4;    // per warp: 32 rows x 16 columns fp32 (transpose staging)
    static constexpr size_t SMEM_BYTES = (size_t)(A_STAGES + A_LO_STAGES) * A_TILE_BYTES + (size_t)B_STAGES * B_STAGE_BYTES + EPI_STAGE_BYTES +
                                         1024 /*alignment slack*/ + 256 /*barriers*/;
    static constexpr int W_BLOCK_FLOATS = NPARTS * B_TILE_BYTES / 4;   // floats copied per (n-tile, k-chunk)
};

// Persistent, warp-specialised:
//   warps 0-7   epilogue: drain one of two TMEM accumulators while the next tile's MMAs fill the other
//   warp  8     MMA issuer (one elected thread)
//   warp  9     weight-tile copies (one elected thread, cp.async.bulk) into their own ring
//   warps 10-17 A producers: cp.async 16-byte chunks straight into the swizzled hi tile of a 3-deep ring (two
//               items in flight across tile boundaries; each thread later touches only the chunks it copied, so
//               cp.async.wait_group is the only synchronisation), then in place: transform (if any), TF32 lo part
template <int BN, int NSPLIT, class AL, class EP>
__global__ void __launch_bounds__(THREADS2, 1)
gemm_tc_kernel(TcDims d, AL al, const float* __restrict__ Wp, EP ep) {
    using Cfg = TcCfg<BN, NSPLIT>;
    constexpr int AS = Cfg::A_STAGES, ALS = Cfg::A_LO_STAGES, BS = Cfg::B_STAGES;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
    static_assert(NSPLIT != 3 || ALS >= 2, "3xTF32 needs a lo ring");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t alo_base = smem_base + AS * A_TILE_BYTES;
    const uint32_t b_base = alo_base + ALS * A_TILE_BYTES;
    const uint32_t epi_base = b_base + BS * Cfg::B_STAGE_BYTES;
    const uint32_t bar_base = epi_base + Cfg::EPI_STAGE_BYTES;
    auto fulla_bar = [&](int s) { return bar_base + 8u * s; };
    auto emptya_bar = [&](int s) { return bar_base + 8u * (AS + s); };
    auto fullb_bar = [&](int s) { return bar_base + 8u * (2 * AS + s); };
    auto emptyb_bar = [&](int s) { return bar_base + 8u * (2 * AS + BS + s); };
    auto accf_bar = [&](int b) { return bar_base + 8u * (2 * AS + 2 * BS + b); };
    auto acce_bar = [&](int b) { return bar_base + 8u * (2 * AS + 2 * BS + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * AS + 2 * BS + 4);
    auto a_hi = [&](int s) { return smem_base + s * A_TILE_BYTES; };
    auto a_lo = [&](int s) { return alo_base + s * A_TILE_BYTES; };
    auto b_hi = [&](int s) { return b_base + s * Cfg::B_STAGE_BYTES; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = d.K / BK;
    const int n_tiles = (d.N + BN - 1) / BN;
    const int total_tiles = ((d.M + BM - 1) / BM) * n_tiles;
    // blocked assignment: CTA c owns a contiguous run of tiles (n fastest), so the n-tiles of one row block are
    // processed back to back by the same CTA (row statistics computed once, activation rows hot in L1/L2)
    const int tq = total_tiles / (int)gridDim.x, tr = total_tiles % (int)gridDim.x;
    // Tail balancing: with T tiles on G CTAs the tr = T % G leftover tiles used to cost
    // a whole extra round (300 tiles on 148 SMs: 3 rounds for 2.03; 150 tiles: 2 for 1.01).  Each leftover tile is instead cut
    // along N into S slivers (S * tr <= G, sliver width a multiple of 16) handed to different CTAs as their LAST item: a sliver
    // repeats the tile's operand production but only 1/S of its MMAs, weight traffic and epilogue.
    int sl_S = 1;
    if (BN == 192 && tq >= 1 && tr > 0 && !(d.dbg & 32)) {
        const int cand[5] = {12, 6, 4, 3, 2};
        for (int k = 0; k < 5; ++k)
            if (cand[k] * tr <= (int)gridDim.x) { sl_S = cand[k]; break; }
    }
    const bool sliver_mode = sl_S > 1;
    const int my_tiles = sliver_mode ? tq + ((int)blockIdx.x < tr * sl_S ? 1 : 0) : tq + ((int)blockIdx.x < tr ? 1 : 0);
    const int tile0 = sliver_mode ? (int)blockIdx.x * tq : (int)blockIdx.x * tq + min((int)blockIdx.x, tr);
    struct Item { int mt, nt, n0, nw; };
    auto item = [&](int i) -> Item {       // i-th work item of this CTA (the same for every role)
        if (sliver_mode && i >= tq) {
            const int s = (int)blockIdx.x;
            const int tile = (int)gridDim.x * tq + s / sl_S, mt = tile / n_tiles;
            return Item{mt, tile - mt * n_tiles, (s % sl_S) * (BN / sl_S), BN / sl_S};
        }
        const int tile = tile0 + i;
        const int mt = tile / n_tiles;
        return Item{mt, tile - mt * n_tiles, 0, BN};
    };

    if (tid == EPI_WARPS * 32) {
        for (int s = 0; s < AS; ++s) {
            mbar_init(fulla_bar(s), LOADER_THREADS);      // every loader thread, after its chunks are final
            mbar_init(emptya_bar(s), 1);                  // one tcgen05.commit
        }
        for (int s = 0; s < BS; ++s) {
            mbar_init(fullb_bar(s), 1);                   // the weight-copy thread's arrive.expect_tx (+ tx bytes)
            mbar_init(emptyb_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);                    // tcgen05.commit after the tile's last MMA
            mbar_init(acce_bar(b), EPI_WARPS * 32);       // every epilogue thread after draining
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == EPI_WARPS) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // PDL: everything above overlapped the previous kernel's tail.  The weight-copy warp does not wait at all (weights
    // are constants), so the B ring is already full when the activations become available.
    if (warp != EPI_WARPS + 1) pdl_wait();
    pdl_trigger();

    if (warp < EPI_WARPS) {
        // ===================== epilogue =====================
        // 8 warps: warp w drains TMEM lanes [32 (w%4), +32) (= rows of the tile) and the 16-column chunks j with
        // j % 2 == w / 4.  tcgen05.ld hands each lane one row; a warp-private swizzled smem transpose turns that
        // into 4 lanes per row (64 contiguous bytes per row, 8 rows per instruction) so residual loads and output
        // stores are sector-coalesced instead of 32 scattered 16-byte accesses per instruction.  Residual float4s
        // are loaded EPI_PF chunks ahead (before the accumulator is even complete).
        constexpr int NCH = BN / 16;
        constexpr int MYCH = (NCH + 1) / 2;                // chunks per warp (column half)
        constexpr int EPI_PF = MYCH < 2 ? MYCH : 2;
        const int q = warp & 3, half = warp >> 2;
        const int lr = lane >> 2, lq = lane & 3;           // after the transpose: rows lr + 8 i, float4 column lq
        const uint32_t stg = epi_base + (uint32_t)warp * (32 * 16 * 4);
        const bool has_res = ep.tc_has_res() && !(d.dbg & 4);
        float4 pf[EPI_PF][4];
#pragma unroll
        for (int jj = 0; jj < EPI_PF; ++jj)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) pf[jj][ii] = zero4();
        for (int i = 0; i < my_tiles; ++i) {
            const Item it_ = item(i);
            const int mt = it_.mt, nt = it_.nt;
            const int nch_i = it_.nw / 16;                 // 16-column chunks of this item (NCH, or fewer for a sliver)
            const int ncol0 = nt * BN + it_.n0;            // first output column of this item
            const int buf = i & 1;
            const int mrow0 = mt * BM + q * 32 + lr;       // + 8 ii
            typename EP::RowCtx rc[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) rc[ii] = ep.tc_row(mrow0 + 8 * ii, d.M);
            if (has_res) {                                 // CTA-uniform; the loads themselves are unconditional
#pragma unroll
                for (int jj = 0; jj < EPI_PF; ++jj) {
                    const int j = 2 * jj + half;
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) pf[jj][ii] = ep.tc_prefetch4(rc[ii], ncol0 + (j < nch_i ? j : 0) * 16 + lq * 4);
                }
            }
            mbar_wait(accf_bar(buf), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Cfg::ACC_STRIDE);
#pragma unroll
            for (int jj = 0; jj < MYCH; ++jj) {
                const int j = 2 * jj + half;
                if (j < nch_i) {                           // warp-uniform
                    float v[16];
                    tmem_ld16(taddr + j * 16, v);
                    // lane = row `lane`: write 4 float4 with the float4-column XOR-swizzled by (row >> 1) & 3
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
                        if (m < d.M && !(d.dbg & 4)) ep.store4(rc[ii], m, n, acc, pf[jj % EPI_PF][ii]);
                    }
                    __syncwarp();
                    const int jn = j + 2 * EPI_PF;
                    if (has_res && jn < nch_i) {           // warp-uniform
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) pf[jj % EPI_PF][ii] = ep.tc_prefetch4(rc[ii], ncol0 + jn * 16 + lq * 4);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acce_bar(buf));
        }
    } else if (warp == EPI_WARPS) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int sa = 0, sb = 0, sl = 0;
            uint32_t pa = 0, pb = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                const Item it_ = item(i);
                const uint32_t idesc = make_idesc_tf32(BM, it_.nw);          // a sliver multiplies by rows [n0, n0 + nw) of the weight tile
                const uint32_t boff = (uint32_t)it_.n0 * ROW_BYTES;          // (n0 is a multiple of 16 rows: whole 1024-byte swizzle atoms)
                mbar_wait(acce_bar(buf), ((uint32_t)(i >> 1) & 1u) ^ 1u);   // accumulator drained (first use: free)
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::ACC_STRIDE);
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(fulla_bar(sa), pa);
                    mbar_wait(fullb_bar(sb), pb);
                    tc_fence_after();
                    const uint64_t da = make_desc(a_hi(sa));
                    const uint64_t db = make_desc(b_hi(sb) + boff);
                    const uint64_t dal = make_desc(a_lo(sl));
                    const uint64_t dbl = make_desc(b_hi(sb) + Cfg::B_TILE_BYTES + boff);
#pragma unroll
                    for (int k4 = 0; k4 < ((d.dbg & 8) ? 0 : BK / 8); ++k4) {
                        const uint64_t adv = (uint64_t)(k4 * 2);   // 8 tf32 = 32 bytes = 2 x 16-byte units along K (inside one swizzle row)
                        mma_tf32(tacc, da + adv, db + adv, idesc, (kc | k4) != 0 ? 1u : 0u);
                        if constexpr (NSPLIT == 3) {
                            mma_tf32(tacc, dal + adv, db + adv, idesc, 1u);
                            mma_tf32(tacc, da + adv, dbl + adv, idesc, 1u);
                        }
                    }
                    mma_commit(emptya_bar(sa));   // stages free once these MMAs have read them
                    mma_commit(emptyb_bar(sb));
                    if (++sa == AS) { sa = 0; pa ^= 1u; }
                    if (++sb == BS) { sb = 0; pb ^= 1u; }
                    if (ALS > 0 && ++sl == ALS) sl = 0;
                }
                mma_commit(accf_bar(buf));      // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp == EPI_WARPS + 1) {
        // ===================== weight copies =====================
        if (lane == 0) {
            constexpr uint32_t bytes = (uint32_t)Cfg::W_BLOCK_FLOATS * 4u;
            int sb = 0;
            uint32_t pb = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const Item it_ = item(i);
                const float* wsrc = Wp + (size_t)it_.nt * nk * d.w_block_floats;
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(emptyb_bar(sb), pb ^ 1u);
                    if (d.dbg & 2) {
                        mbar_arrive(fullb_bar(sb));
                    } else if (it_.nw == BN) {
                        mbar_arrive_expect_tx(fullb_bar(sb), bytes);
                        bulk_g2s(b_hi(sb), wsrc + (size_t)kc * d.w_block_floats, bytes, fullb_bar(sb));
                    } else {   // sliver: only its rows of the hi and (3xTF32) lo tiles
                        const uint32_t part = (uint32_t)it_.nw * ROW_BYTES, roff = (uint32_t)it_.n0 * ROW_BYTES;
                        const float* blk = wsrc + (size_t)kc * d.w_block_floats;
                        mbar_arrive_expect_tx(fullb_bar(sb), part * Cfg::NPARTS);
                        bulk_g2s(b_hi(sb) + roff, blk + roff / 4, part, fullb_bar(sb));
                        if (Cfg::NPARTS == 2)
                            bulk_g2s(b_hi(sb) + Cfg::B_TILE_BYTES + roff, blk + (Cfg::B_TILE_BYTES + roff) / 4, part, fullb_bar(sb));
                    }
                    if (++sb == BS) { sb = 0; pb ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== A producers =====================
        const int lt = tid - LOADER_TID0;
        const int c = lt & (ROW_CHUNKS - 1);   // 16-byte chunk of the row this thread owns
        const int rb = lt / ROW_CHUNKS;        // rows rb + LROW_STEP * i
        uint32_t off[LROWS];
#pragma unroll
        for (int i = 0; i < LROWS; ++i) off[i] = swz_off(rb + LROW_STEP * i, c);
        const bool ident = al.identity();
        const int total_items = my_tiles * nk;
        // ---- issue pointer: item `ii` (tile ti_i, chunk kc_i) goes to hi stage sa_i; runs up to AS - 1 items ahead of
        //      the transform pointer, across tile boundaries
        int ii = 0, ti_i = 0, kc_i = 0, sa_i = 0;
        uint32_t pa_i = 0;
        typename AL::ICtx ic[LROWS];
        auto issue_one = [&]() {                 // the stage must be free (MMAs of item ii - AS complete)
            if (kc_i == 0) {
                const int m0 = item(ti_i).mt * BM;
#pragma unroll
                for (int i = 0; i < LROWS; ++i) ic[i] = al.iprep(m0 + rb + LROW_STEP * i);
            }
            const uint32_t dst = a_hi(sa_i);
#pragma unroll
            for (int i = 0; i < LROWS; ++i) {
                bool valid;
                const float* src = al.isrc(ic[i], kc_i * BK + c * 4, valid);
                cp_async16(dst + off[i], src, valid ? 16u : 0u);
            }
            cp_async_commit();
            if (++kc_i == nk) { kc_i = 0; ++ti_i; }
            if (++sa_i == AS) { sa_i = 0; pa_i ^= 1u; }
            ++ii;
        };
        for (int j = 0; j < AS - 1 && ii < total_items; ++j) issue_one();   // first use of every stage: free
        // ---- transform pointer
        typename AL::Ctx ctx[LROWS];
        int it = 0, sa_x = 0, sl_x = 0;
        int sj = 0;                              // hi stage (and its phase) of item it - ALS, whose MMAs free lo stage sl_x
        uint32_t pj = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int m0 = item(ti).mt * BM;
            const bool new_rows = ti == 0 || item(ti).mt != item(ti - 1).mt;
            for (int kc = 0; kc < nk; ++kc, ++it) {
                // keep the ring full without blocking: a stage is free once the MMAs that read it have completed
                while (ii < total_items && ii - it < AS && mbar_test_wait(emptya_bar(sa_i), pa_i ^ 1u)) issue_one();
                if (ii == it) {                  // nothing in flight for the current item: must block
                    mbar_wait(emptya_bar(sa_i), pa_i ^ 1u);
                    issue_one();
                }
                // this thread's chunks of item `it` have landed when at most ii - it - 1 newer groups are pending
                switch (ii - it - 1) {
                    case 0: cp_async_wait<0>(); break;
                    case 1: cp_async_wait<1>(); break;
                    case 2: cp_async_wait<2>(); break;
                    case 3: cp_async_wait<3>(); break;
                    case 4: cp_async_wait<4>(); break;
                    default: cp_async_wait<5>(); break;
                }
                if (kc == 0 && !ident && new_rows) {
#pragma unroll
                    for (int i = 0; i < LROWS; ++i) ctx[i] = al.prep(m0 + rb + LROW_STEP * i, c);
                }
                const uint32_t abase = a_hi(sa_x);
                if (NSPLIT == 3 || !ident) {
                    if constexpr (NSPLIT == 3) {
                        if (it >= ALS) {         // lo stage sl_x was last read by the MMAs of item it - ALS
                            mbar_wait(emptya_bar(sj), pj);
                            if (++sj == AS) { sj = 0; pj ^= 1u; }
                        }
                    }
                    const uint32_t lbase = a_lo(sl_x);
                    float4 xs[LROWS];          // all shared loads first: the volatile asm statements keep program order
#pragma unroll
                    for (int i = 0; i < LROWS; ++i)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(xs[i].x), "=f"(xs[i].y), "=f"(xs[i].z), "=f"(xs[i].w) : "r"(abase + off[i]));
#pragma unroll
                    for (int i = 0; i < LROWS; ++i) {
                        float4 x = xs[i];
                        float4 h;
                        if (!ident) {
                            x = al.xform(ctx[i], kc * BK + c * 4, x);
                            h.x = tf32_hi(x.x); h.y = tf32_hi(x.y); h.z = tf32_hi(x.z); h.w = tf32_hi(x.w);
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(abase + off[i]), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                        } else {
                            // untouched fp32 in the hi tile: the tensor core reads its upper 19 bits (TF32 truncation)
                            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                        }
                        if constexpr (NSPLIT == 3) {
                            float4 l;      // unrounded: the tensor core truncates it (see tf32_split4)
                            l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lbase + off[i]), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                        }
                    }
                }
                fence_proxy_async();
                mbar_arrive(fulla_bar(sa_x));
                if (++sa_x == AS) sa_x = 0;
                if (ALS > 0 && ++sl_x == ALS) sl_x = 0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// =====================================================================================================
// A-operand-in-TMEM variant ("TS" MMA).  With both operands in shared memory the 3xTF32 kernel above is
// bound by shared-memory bandwidth, not by the tensor pipe: per 32-wide K chunk the three MMAs read 48 KB of A
// and 72 KB of B, the producers move another 48 KB, the weight copy writes 48 KB -- 216 KB against 128 B/cycle
// is ~1700 cycles for 1152 cycles of MMA.  Here the activation tile never touches shared memory: each
// producer thread owns one row of the tile (= one TMEM lane), loads its 32 K-values straight from global into
// registers (L2::256B so HBM sees long bursts; the 128-byte row segment is one cache line), transforms and
// splits them, and writes hi and lo with tcgen05.st into a 2-stage A ring in TMEM (columns 384..511, next to
// the two 192-column accumulators).  tcgen05.mma then takes A from TMEM and only B from shared memory.
// MEASURED (profiles/r1_gemm_variants.md): correct, but 20-30 % slower than the shared-memory-A kernel at the
// 96-register budget of a 576-thread CTA (the per-thread 32-value hi/lo staging spills, and row-per-thread global
// loads cost 8x the L1 wavefronts of the coalesced cp.async path).  Not the default; selected with SAID_TC_TMEM_A=1.
// =====================================================================================================
SAID_DEVINL void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
SAID_DEVINL void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
          "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
          "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
          "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
SAID_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int BN, int NSPLIT>
struct TcaCfg {
    static constexpr int NPARTS = NSPLIT == 3 ? 2 : 1;
    static constexpr int B_TILE_BYTES = BN * ROW_BYTES;
    static constexpr int B_STAGE_BYTES = NPARTS * B_TILE_BYTES;
    static constexpr int B_STAGES = NSPLIT == 3 ? 3 : 6;
    static constexpr int ACC_STRIDE = 192;                             // TMEM columns between the two accumulators
    static constexpr int A_COL = 384;                                  // TMEM A ring: stage s at A_COL + 64 s: hi [0,32), lo [32,64)
    static constexpr int TMEM_COLS = 512;
    static constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * 16 * 4;
    static constexpr size_t SMEM_BYTES = (size_t)B_STAGES * B_STAGE_BYTES + EPI_STAGE_BYTES + 1024 + 256;
    static constexpr int W_BLOCK_FLOATS = NPARTS * B_TILE_BYTES / 4;
};

template <int BN, int NSPLIT, class AL, class EP>
__global__ void __launch_bounds__(THREADS2, 1)
gemm_tca_kernel(TcDims d, AL al, const float* __restrict__ Wp, EP ep) {
    using Cfg = TcaCfg<BN, NSPLIT>;
    constexpr int BS = Cfg::B_STAGES;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 192, "UMMA N (two accumulators + the A ring must fit 512 TMEM columns)");
    static_assert(BK == 32, "one A stage = 32 K-values = 32 TMEM columns");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + BS * Cfg::B_STAGE_BYTES;
    const uint32_t bar_base = epi_base + Cfg::EPI_STAGE_BYTES;
    auto fulla_bar = [&](int s) { return bar_base + 8u * s; };
    auto emptya_bar = [&](int s) { return bar_base + 8u * (2 + s); };
    auto fullb_bar = [&](int s) { return bar_base + 8u * (4 + s); };
    auto emptyb_bar = [&](int s) { return bar_base + 8u * (4 + BS + s); };
    auto accf_bar = [&](int b) { return bar_base + 8u * (4 + 2 * BS + b); };
    auto acce_bar = [&](int b) { return bar_base + 8u * (4 + 2 * BS + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (4 + 2 * BS + 4);
    auto b_hi = [&](int s) { return smem_base + s * Cfg::B_STAGE_BYTES; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nk = d.K / BK;
    const int n_tiles = (d.N + BN - 1) / BN;
    const int total_tiles = ((d.M + BM - 1) / BM) * n_tiles;
    const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == EPI_WARPS * 32) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(fulla_bar(s), 128);                 // the four producer warps (one per TMEM lane quarter) of that stage
            mbar_init(emptya_bar(s), 1);
        }
        for (int s = 0; s < BS; ++s) {
            mbar_init(fullb_bar(s), 1);
            mbar_init(emptyb_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accf_bar(b), 1);
            mbar_init(acce_bar(b), EPI_WARPS * 32);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == EPI_WARPS) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < EPI_WARPS) {
        // ===================== epilogue =====================
        // 8 warps: warp w drains TMEM lanes [32 (w%4), +32) (= rows of the tile) and the 16-column chunks j with
        // j % 2 == w / 4.  tcgen05.ld hands each lane one row; a warp-private swizzled smem transpose turns that
        // into 4 lanes per row (64 contiguous bytes per row, 8 rows per instruction) so residual loads and output
        // stores are sector-coalesced instead of 32 scattered 16-byte accesses per instruction.  Residual float4s
        // are loaded EPI_PF chunks ahead (before the accumulator is even complete).
        constexpr int NCH = BN / 16;
        constexpr int MYCH = (NCH + 1) / 2;                // chunks per warp (column half)
        constexpr int EPI_PF = MYCH < 2 ? MYCH : 2;
        const int q = warp & 3, half = warp >> 2;
        const int lr = lane >> 2, lq = lane & 3;           // after the transpose: rows lr + 8 i, float4 column lq
        const uint32_t stg = epi_base + (uint32_t)warp * (32 * 16 * 4);
        const bool has_res = ep.tc_has_res() && !(d.dbg & 4);
        float4 pf[EPI_PF][4];
#pragma unroll
        for (int jj = 0; jj < EPI_PF; ++jj)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) pf[jj][ii] = zero4();
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
            const int buf = i & 1;
            const int mrow0 = mt * BM + q * 32 + lr;       // + 8 ii
            typename EP::RowCtx rc[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) rc[ii] = ep.tc_row(mrow0 + 8 * ii, d.M);
            if (has_res) {                                 // CTA-uniform; the loads themselves are unconditional
#pragma unroll
                for (int jj = 0; jj < EPI_PF; ++jj) {
                    const int j = 2 * jj + half;
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) pf[jj][ii] = ep.tc_prefetch4(rc[ii], nt * BN + (j < NCH ? j : 0) * 16 + lq * 4);
                }
            }
            mbar_wait(accf_bar(buf), (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Cfg::ACC_STRIDE);
#pragma unroll
            for (int jj = 0; jj < MYCH; ++jj) {
                const int j = 2 * jj + half;
                if (j < NCH) {                             // warp-uniform
                    float v[16];
                    tmem_ld16(taddr + j * 16, v);
                    // lane = row `lane`: write 4 float4 with the float4-column XOR-swizzled by (row >> 1) & 3
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const uint32_t a = stg + (uint32_t)lane * 64u + (uint32_t)((c4 ^ ((lane >> 1) & 3)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[4 * c4]), "f"(v[4 * c4 + 1]),
                                     "f"(v[4 * c4 + 2]), "f"(v[4 * c4 + 3]) : "memory");
                    }
                    __syncwarp();
                    const int n = nt * BN + j * 16 + lq * 4;
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const int r = lr + 8 * ii;
                        const uint32_t a = stg + (uint32_t)r * 64u + (uint32_t)((lq ^ ((r >> 1) & 3)) << 4);
                        float4 acc;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "r"(a));
                        const int m = mrow0 + 8 * ii;
                        if (m < d.M && !(d.dbg & 4)) ep.store4(rc[ii], m, n, acc, pf[jj % EPI_PF][ii]);
                    }
                    __syncwarp();
                    const int jn = j + 2 * EPI_PF;
                    if (has_res && jn < NCH) {             // warp-uniform
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) pf[jj % EPI_PF][ii] = ep.tc_prefetch4(rc[ii], nt * BN + jn * 16 + lq * 4);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acce_bar(buf));
        }
    } else if (warp == EPI_WARPS) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            int sb = 0, it = 0;
            uint32_t pb = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                mbar_wait(acce_bar(buf), ((uint32_t)(i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::ACC_STRIDE);
                for (int kc = 0; kc < nk; ++kc, ++it) {
                    const int sa = it & 1;
                    mbar_wait(fulla_bar(sa), (uint32_t)(it >> 1) & 1u);
                    mbar_wait(fullb_bar(sb), pb);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + (uint32_t)(Cfg::A_COL + sa * 64);
                    const uint64_t db = make_desc(b_hi(sb));
                    const uint64_t dbl = make_desc(b_hi(sb) + Cfg::B_TILE_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < BK / 8; ++k4) {
                        const uint64_t adv = (uint64_t)(k4 * 2);
                        mma_tf32_ts(tacc, ta + k4 * 8, db + adv, idesc, (kc | k4) != 0 ? 1u : 0u);
                        if constexpr (NSPLIT == 3) {
                            mma_tf32_ts(tacc, ta + 32 + k4 * 8, db + adv, idesc, 1u);
                            mma_tf32_ts(tacc, ta + k4 * 8, dbl + adv, idesc, 1u);
                        }
                    }
                    mma_commit(emptya_bar(sa));
                    mma_commit(emptyb_bar(sb));
                    if (++sb == BS) { sb = 0; pb ^= 1u; }
                }
                mma_commit(accf_bar(buf));
            }
        }
        __syncwarp();
    } else if (warp == EPI_WARPS + 1) {
        // ===================== weight copies =====================
        if (lane == 0) {
            constexpr uint32_t bytes = (uint32_t)Cfg::W_BLOCK_FLOATS * 4u;
            int sb = 0;
            uint32_t pb = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int nt = tile % n_tiles;
                const float* wsrc = Wp + (size_t)nt * nk * d.w_block_floats;
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(emptyb_bar(sb), pb ^ 1u);
                    mbar_arrive_expect_tx(fullb_bar(sb), bytes);
                    bulk_g2s(b_hi(sb), wsrc + (size_t)kc * d.w_block_floats, bytes, fullb_bar(sb));
                    if (++sb == BS) { sb = 0; pb ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== A producers: one thread = one row of the tile = one TMEM lane =====================
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int pair = (warp - (EPI_WARPS + 2)) >> 2;    // 0: even items (A stage 0), 1: odd items (A stage 1)
        const int row = q * 32 + lane;
        const uint32_t tst = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::A_COL + pair * 64);
        const int total_items = my_tiles * nk;
        const bool ident = al.identity();
        typename AL::ICtx ic;
        typename AL::Ctx cx;
        int ic_tile = -1, cx_tile = -1;
        auto load_item = [&](int j, float4 (&x)[8]) {
            const int ti = j / nk, kc = j - ti * nk;
            if (ti != ic_tile) {
                ic = al.iprep(((blockIdx.x + ti * gridDim.x) / n_tiles) * BM + row);
                ic_tile = ti;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                bool valid;
                const float* src = al.isrc(ic, kc * BK + c * 4, valid);
                x[c] = valid ? ldg4_l2pf(src) : zero4();
            }
        };
        float4 x[8];
        if (pair < total_items) load_item(pair, x);
        uint32_t use = 0;
        for (int j = pair; j < total_items; j += 2, ++use) {
            const int ti = j / nk, kc = j - ti * nk;
            if (!ident && ti != cx_tile) {
                const int m = ((blockIdx.x + ti * gridDim.x) / n_tiles) * BM + row;
                if constexpr (AL::kTag == 1) cx = al.prep_row(m);
                else cx = al.prep(m, 0);
                cx_tile = ti;
            }
            float hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = x[c];
                if (!ident) v = al.xform(cx, kc * BK + c * 4, v);
                hi[4 * c] = rna_tf32(v.x); hi[4 * c + 1] = rna_tf32(v.y); hi[4 * c + 2] = rna_tf32(v.z); hi[4 * c + 3] = rna_tf32(v.w);
                if constexpr (NSPLIT == 3) {
                    lo[4 * c] = rna_tf32(v.x - hi[4 * c]); lo[4 * c + 1] = rna_tf32(v.y - hi[4 * c + 1]);
                    lo[4 * c + 2] = rna_tf32(v.z - hi[4 * c + 2]); lo[4 * c + 3] = rna_tf32(v.w - hi[4 * c + 3]);
                }
            }
            if (j + 2 < total_items) load_item(j + 2, x);         // next item's loads fly while we wait for the stage
            mbar_wait(emptya_bar(pair), (use & 1u) ^ 1u);         // the MMAs that read this A stage have completed
            tc_fence_after();
            tmem_st32(tst, hi);
            if constexpr (NSPLIT == 3) tmem_st32(tst + 32, lo);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(fulla_bar(pair));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: one process may drive several GPUs (one engine
// each), so "already configured" is remembered per (kernel, device), not per kernel.
inline cudaError_t configure_once(const void* kern, int smem_bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kern, dev})) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) done.insert({kern, dev});
    return e;
}

template <int BN, int NSPLIT, class AL, class EP>
inline cudaError_t launch_gemm_tca(cudaStream_t st, int num_sms, int M, int N, int K, const AL& al, const float* Wp,
                                   int w_block_floats, const EP& ep) {
    using Cfg = TcaCfg<BN, NSPLIT>;
    auto kern = gemm_tca_kernel<BN, NSPLIT, AL, EP>;
    {
        cudaError_t e = configure_once((const void*)kern, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    TcDims d{M, N, K, w_block_floats, 0};
    const int total_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    kern<<<grid, THREADS2, Cfg::SMEM_BYTES, st>>>(d, al, Wp, ep);
    return cudaGetLastError();
}

// Host: pack a K-major x N weight matrix Wt (K rows, ldw floats per row, columns [0, N)) into the
// per-(n-tile, k-chunk) shared-memory images the kernel copies verbatim:
//   block(nt, kc) = [hi tile: BN rows x 128 B, swizzled] [lo tile]      (lo only when nsplit == 3)
// Columns beyond N (last tile) are zero.
inline float host_rna_tf32(float x) {
    uint32_t b;
    memcpy(&b, &x, 4);
    if ((b & 0x7F800000u) == 0x7F800000u) return x;   // inf / nan
    b = (b + 0x1000u) & 0xFFFFE000u;                 // round to nearest, ties away, on the magnitude
    float r;
    memcpy(&r, &b, 4);
    return r;
}
inline void pack_weights_tc(const float* Wt, int K, int N, int ldw, int BN, int nsplit, std::vector<float>& out) {
    const int nparts = nsplit == 3 ? 2 : 1;
    const int ntiles = (N + BN - 1) / BN, nk = K / BK;
    const size_t tile_floats = (size_t)BN * BK;
    const size_t block = (size_t)nparts * tile_floats;
    out.assign((size_t)ntiles * nk * block, 0.f);
    for (int nt = 0; nt < ntiles; ++nt)
        for (int kc = 0; kc < nk; ++kc) {
            float* hi = out.data() + ((size_t)nt * nk + kc) * block;
            float* lo = hi + tile_floats;
            for (int r = 0; r < BN; ++r) {
                const int n = nt * BN + r;
                if (n >= N) continue;
                for (int kk = 0; kk < BK; ++kk) {
                    const float w = Wt[(size_t)(kc * BK + kk) * ldw + n];
                    const int chunk = kk >> 2;
                    const int sw = (chunk ^ ((r * ROW_BYTES) >> 7)) & (ROW_CHUNKS - 1);
                    const size_t idx = (size_t)r * BK + (size_t)sw * 4 + (kk & 3);
                    const float h = host_rna_tf32(w);
                    hi[idx] = h;
                    if (nparts == 2) lo[idx] = host_rna_tf32(w - h);
                }
            }
        }
}

template <int BN, int NSPLIT, class AL, class EP>
inline cudaError_t launch_gemm_tc(cudaStream_t st, int num_sms, int M, int N, int K, const AL& al, const float* Wp,
                                  int w_block_floats, const EP& ep, int dbg = 0, bool pdl = false) {
    using Cfg = TcCfg<BN, NSPLIT>;
    auto kern = gemm_tc_kernel<BN, NSPLIT, AL, EP>;
    {
        cudaError_t e = configure_once((const void*)kern, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    TcDims d{M, N, K, w_block_floats, dbg};
    const int total_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    return launch_ex(kern, dim3(grid), dim3(THREADS2), Cfg::SMEM_BYTES, st, pdl, 1, d, al, Wp, ep);
}

}  // namespace tc
}  // namespace said
