// tcgen05 (5th-gen tensor core) GEMM with fused operand transforms and epilogues, sm_100a only.
//
//   out[m, n] = epilogue( sum_k  load(m, k) * W[n, k] )
//
// Same loaders / epilogues as gemm_simt.cuh (normalisation + activation folded into the A operand,
// bias / residual / GEGLU folded into the epilogue), but the contraction runs on the tensor cores:
//
//   * A (activations): 4 loader warps read fp32 from global (coalesced 16-byte chunks), apply the
//     loader transform, split every value into TF32 hi + TF32 lo  (x = hi + lo + O(2^-22 x)) and write
//     both into shared memory in the UMMA canonical K-major SWIZZLE_128B layout (32 fp32 = 128 B per row,
//     16-byte chunk index XOR row%8), followed by fence.proxy.async + mbarrier arrive.
//   * B (weights): pre-split and pre-swizzled on the host into ready-to-use tile images, one contiguous
//     block per (n-tile, k-chunk); a single cp.async.bulk (TMA engine, UBLKCP) per stage lands it and
//     completes the stage's mbarrier transaction count.
//   * MMA: one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) with the accumulator in
//     TMEM.  NSPLIT == 3 issues hi*hi + lo*hi + hi*lo ("3xTF32": fp32-level accuracy, which is what keeps
//     the path inside the reference's fp32 tolerance); NSPLIT == 1 issues hi*hi only.
//     tcgen05.commit releases the stage back to the producers and finally signals the epilogue.
//   * Epilogue: the 4 loader warps read their 32 TMEM lanes (tcgen05.ld 32x32b.x16), apply the epilogue and
//     store.
//
// One CTA = one 128 x BN output tile; STAGES-deep mbarrier pipeline over K in chunks of 32.
#pragma once
#include <cstring>
#include <vector>

#include "common.cuh"

namespace said {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                    // fp32 elements per stage row = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BM * 128;
constexpr int THREADS = 192;              // warps 0-3 loaders + epilogue, warp 4 MMA, warp 5 weight copies

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
// A wrong barrier protocol must surface as an error, not as a hung GPU: trap after ~seconds of spinning.
SAID_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
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

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
SAID_DEVINL uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
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

struct TcDims {
    int M, N, K;
    int w_block_floats;   // floats between consecutive (n-tile, k-chunk) blocks of the weight image
};

template <int BN, int NSPLIT>
struct TcCfg {
    static constexpr int NPARTS = NSPLIT == 3 ? 2 : 1;                 // hi (+ lo) copies of each operand
    static constexpr int B_TILE_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = NPARTS * (A_TILE_BYTES + B_TILE_BYTES);
    static constexpr int STAGES = (NSPLIT == 3) ? 2 : 4;
    static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    // floats per (n-tile, k-chunk) block of the packed weight image
    static constexpr int W_BLOCK_FLOATS = NPARTS * B_TILE_BYTES / 4;
};

template <int BN, int NSPLIT, class AL, class EP>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(TcDims d, AL al, const float* __restrict__ Wp, EP ep) {
    using Cfg = TcCfg<BN, NSPLIT>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t acc_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
    auto a_hi = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto b_hi = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::NPARTS * A_TILE_BYTES; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM;
    const int nt = blockIdx.y;
    const int nk = d.K / BK;

    if (tid == 128) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 128 + 1);   // 128 loader threads + the weight-copy thread (with tx bytes)
            mbar_init(empty_bar(s), 1);        // one tcgen05.commit
        }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 4) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < 4) {
        // ===================== A producers =====================
        const int c = tid & 7;                 // 16-byte chunk of the 128-byte row this thread owns
        const int rbase = tid >> 3;            // rows rbase + 16 i
        const uint32_t swz = (uint32_t)((c ^ (rbase & 7)) << 4);
        typename AL::Ctx ctx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ctx[i] = al.prep(m0 + rbase + 16 * i, c);
        for (int kc = 0; kc < nk; ++kc) {
            const int s = kc % STAGES;
            const uint32_t u = (uint32_t)(kc / STAGES);
            float4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = al.load4(ctx[i], kc * BK + c * 4);
            mbar_wait(empty_bar(s), (u & 1u) ^ 1u);
            const uint32_t abase = a_hi(s);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t off = (uint32_t)(rbase + 16 * i) * 128u + swz;
                float4 h;
                h.x = rna_tf32(x[i].x); h.y = rna_tf32(x[i].y); h.z = rna_tf32(x[i].z); h.w = rna_tf32(x[i].w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(abase + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                if constexpr (NSPLIT == 3) {
                    float4 l;
                    l.x = rna_tf32(x[i].x - h.x); l.y = rna_tf32(x[i].y - h.y);
                    l.z = rna_tf32(x[i].z - h.z); l.w = rna_tf32(x[i].w - h.w);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(abase + A_TILE_BYTES + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                }
            }
            fence_proxy_async();
            mbar_arrive(full_bar(s));
        }
        // ===================== epilogue =====================
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int m = m0 + tid;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int j = 0; j < BN / 16; ++j) {
            float v[16];
            tmem_ld16(taddr + j * 16, v);
            if (m < d.M) ep.store16(m, nt * BN + j * 16, v);
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % STAGES;
                const uint32_t u = (uint32_t)(kc / STAGES);
                mbar_wait(full_bar(s), u & 1u);
                tc_fence_after();
                const uint64_t da = make_desc_sw128(a_hi(s));
                const uint64_t db = make_desc_sw128(b_hi(s));
                const uint64_t dal = make_desc_sw128(a_hi(s) + A_TILE_BYTES);
                const uint64_t dbl = make_desc_sw128(b_hi(s) + Cfg::B_TILE_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < BK / 8; ++k4) {
                    const uint64_t adv = (uint64_t)(k4 * 2);   // 8 tf32 = 32 bytes = 2 x 16-byte units along K
                    mma_tf32(tmem_base, da + adv, db + adv, idesc, (kc | k4) != 0 ? 1u : 0u);
                    if constexpr (NSPLIT == 3) {
                        mma_tf32(tmem_base, dal + adv, db + adv, idesc, 1u);
                        mma_tf32(tmem_base, da + adv, dbl + adv, idesc, 1u);
                    }
                }
                mma_commit(empty_bar(s));   // stage free once these MMAs have read it
            }
            mma_commit(acc_bar);            // accumulator complete
        }
        __syncwarp();
    } else {
        // ===================== weight copies =====================
        if (lane == 0) {
            const float* wsrc = Wp + (size_t)nt * nk * d.w_block_floats;
            constexpr uint32_t bytes = (uint32_t)Cfg::W_BLOCK_FLOATS * 4u;
            for (int kc = 0; kc < nk; ++kc) {
                const int s = kc % STAGES;
                const uint32_t u = (uint32_t)(kc / STAGES);
                mbar_wait(empty_bar(s), (u & 1u) ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), bytes);
                bulk_g2s(b_hi(s), wsrc + (size_t)kc * d.w_block_floats, bytes, full_bar(s));
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
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
    const size_t block = (size_t)nparts * BN * 32;
    out.assign((size_t)ntiles * nk * block, 0.f);
    for (int nt = 0; nt < ntiles; ++nt)
        for (int kc = 0; kc < nk; ++kc) {
            float* hi = out.data() + ((size_t)nt * nk + kc) * block;
            float* lo = hi + (size_t)BN * 32;
            for (int r = 0; r < BN; ++r) {
                const int n = nt * BN + r;
                if (n >= N) continue;
                for (int kk = 0; kk < 32; ++kk) {
                    const float w = Wt[(size_t)(kc * BK + kk) * ldw + n];
                    const int chunk = kk >> 2;
                    const size_t idx = (size_t)r * 32 + (size_t)((chunk ^ (r & 7)) << 2) + (kk & 3);
                    const float h = host_rna_tf32(w);
                    hi[idx] = h;
                    if (nparts == 2) lo[idx] = host_rna_tf32(w - h);
                }
            }
        }
}

template <int BN, int NSPLIT, class AL, class EP>
inline cudaError_t launch_gemm_tc(cudaStream_t st, int M, int N, int K, const AL& al, const float* Wp, int w_block_floats,
                                  const EP& ep) {
    using Cfg = TcCfg<BN, NSPLIT>;
    static bool configured = false;
    auto kern = gemm_tc_kernel<BN, NSPLIT, AL, EP>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    TcDims d{M, N, K, w_block_floats};
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    kern<<<grid, THREADS, Cfg::SMEM_BYTES, st>>>(d, al, Wp, ep);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace said
