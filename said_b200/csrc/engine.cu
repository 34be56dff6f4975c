// libsaid_sm100.so -- engine and C ABI of the B200-native SAiD inference hot path.
//
// Host-side "program" for the three device phases of SAID.inference() (reference
// said/model/diffusion.py:308-472):
//   encode_audio     Wav2Vec2 feature encoder + transformer        (said/model/wav2vec2.py:14-82)
//   prepare_context  cross-attention K/V of all 4 blocks, hoisted   (said/model/ldm/attention.py:90-91)
//   denoise          N x [UNet1D forward (both CFG branches) + CFG combine + scheduler step + blend]
// Every arithmetic op is a hand-written kernel from the .cuh files next to this one; this file only
// sequences launches, owns workspaces and packs weights.  No cuBLAS / cuDNN / torch.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/said_b200.h"
#include "attention.cuh"
#include "attention_h.cuh"
#include "attention_tc.cuh"
#include "diffusion_kernels.cuh"
#include "encoder_kernels.cuh"
#include "eval_kernels.cuh"
#include "gemm_simt.cuh"
#include "gemm_h.cuh"
#include "ffn_h.cuh"
#include "gemm_tc.cuh"
#include "norm_kernels.cuh"
#include "pair_kernels.cuh"

using namespace said;

namespace {

thread_local std::string g_err;
int fail(const std::string& m) {
    g_err = m;
    return 1;
}
#define CK(x)                                                                                              \
    do {                                                                                                   \
        cudaError_t _e = (x);                                                                              \
        if (_e != cudaSuccess) return fail(std::string(__FILE__) + ":" + std::to_string(__LINE__) + " " +  \
                                           #x + ": " + cudaGetErrorString(_e));                            \
    } while (0)
#define CKI(x)                      \
    do {                            \
        int _r = (x);               \
        if (_r != 0) return _r;     \
    } while (0)

struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
    int64_t numel() const { return (int64_t)data.size(); }
};

unsigned long long g_alloc_gen = 0;   // bumped whenever a workspace moves: cached step graphs bake workspace pointers
struct DevBuf {
    float* p = nullptr;
    size_t cap = 0;   // floats
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        ++g_alloc_gen;
        cudaError_t e = cudaMalloc((void**)&p, n * sizeof(float));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    // workspaces of the padded-row (fp16x3) path are zero-filled when (re)allocated: rows nobody writes stay finite
    cudaError_t ensure_zero(size_t n) {
        if (n <= cap) return cudaSuccess;
        cudaError_t e = ensure(n);
        if (e == cudaSuccess) e = cudaMemset(p, 0, cap * sizeof(float));
        if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the memset runs on the legacy stream; the engine's streams do not wait for it
        return e;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

constexpr int C = 192;        // model_channels (unet_1d_condition.py:40)
constexpr int HEADS = 6;
constexpr int HD = 32;
constexpr int FF = 768;       // GEGLU inner width (attention.py:36-38)
constexpr int TE = 768;       // time_embed_dim

struct ResBlockW {
    int cin = C;
    bool skip = false;
    int k2 = 3 * C;            // contraction of the second conv (+ cin when the 1x1 skip is fused in)
    float *gn1_g, *gn1_b, *w1, *b1, *gn2_g, *gn2_b, *w2, *b2;
};
struct TransformerW {
    float *gn_g, *gn_b, *ln1_g, *ln1_b, *wqkv, *wo1, *bo1, *ln2_g, *ln2_b, *wq2, *wo2, *bo2;
    float *ln3_g, *ln3_b, *wff1, *bff1, *wff2, *bff2, *wproj, *bproj;
    float *wffp, *bffp;        // ff.net.2 and proj_out folded into one (768 + 192) x 192 GEMM: [ff | x] [W2 Wp ; Wp]
};
struct EncLayerW {
    float *wqkv, *bqkv, *wo, *bo, *ln1_g, *ln1_b, *wff1, *bff1, *wff2, *bff2, *ln2_g, *ln2_b;
};

}  // namespace

struct said_engine {
    int device = 0;
    int num_sms = 0;
    bool pdl = getenv("SAID_PDL") != nullptr;   // programmatic dependent launch for the step-loop kernels: measured slower at batch scale, opt-in (common.cuh)
    // ... but ON for the small-row tensor-core path (single clips): its launches fill a fraction of the SMs, so the next kernel's
    // CTAs really are resident early and their prologues (barrier init, TMEM allocation, descriptor prefetch) overlap the
    // predecessor's tail: 0.54 -> 0.48 ms per step for one 5 s clip.  SAID_NO_PDL_SMALL switches it off.
    bool pdl_small = getenv("SAID_NO_PDL_SMALL") == nullptr;
    bool gn_fused = getenv("SAID_GN_TWO_PASS") == nullptr;   // cluster GroupNorm (one launch); the env switch keeps the two-kernel version reachable for A/B timing
    cudaStream_t own_stream = nullptr;
    // Two half-batches on two streams (fp16x3 path, mid-size batches): layer L+1 of one half starts on the SMs that layer L of the
    // other half has already left, which hides part of the row-tile quantisation.  Measured (ms per step, split / whole): 32 clips
    // (151 row tiles) 1.36 / 1.43, 48 clips (226) 1.72 / 1.77, 64 clips (301) 2.41 / 2.26 -- so it is used between
    // split_min_tiles and split_max_tiles only.  Workspaces are the same buffers, one row range per half.
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    struct HalfCtx { int row_base, samp_base, clip_base, half; } hctx{0, 0, 0, 0};
    int split_min_tiles = getenv("SAID_SPLIT_MIN_TILES") ? atoi(getenv("SAID_SPLIT_MIN_TILES")) : 150;
    int split_max_tiles = getenv("SAID_SPLIT_MAX_TILES") ? atoi(getenv("SAID_SPLIT_MAX_TILES")) : 260;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    long long launches = 0;

    // ---- optional per-kernel-family timing (bench.py roofline): one event after every launch
    enum Tag { TAG_GEMM_CONV = 0, TAG_GEMM_LN, TAG_GEMM_PLAIN, TAG_ATTN, TAG_XATTN, TAG_GN, TAG_STEP, TAG_OTHER, TAG_COUNT };
    bool prof_on = false;
    int cur_tag = TAG_OTHER;
    std::vector<std::pair<int, cudaEvent_t>> prof_ev;
    int after_launch(cudaStream_t st) {
        ++launches;
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) return fail(std::string("kernel launch failed: ") + cudaGetErrorString(le));
        if (prof_on) {
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) != cudaSuccess || cudaEventRecord(ev, st) != cudaSuccess) return fail("profiling event failed");
            prof_ev.emplace_back(cur_tag, ev);
        }
        cur_tag = TAG_OTHER;
        return 0;
    }
    // ---- tensor-core path: per GEMM weight, a pre-split / pre-swizzled tile image (gemm_tc.cuh)
    struct TcW { float* img; int K, N, bn; };
    std::map<const float*, TcW> tcmap;   // keyed by the SIMT "Wt" device pointer of the same weight
    int precision = 1;                   // 0: fp32 FFMA, 1: 3xTF32 tcgen05 (fp32-level), 2: 1xTF32 tcgen05, 3: fp16 hi/lo x3 tcgen05 via TMA
    // ---- fp16x3 path (gemm_h.cuh): per GEMM weight, an fp16 hi/lo tile image pre-scaled by 2^exp
    struct HW { uint8_t* img; int K, N, bn, exp; };
    std::map<const float*, HW> hmap;     // keyed like tcmap
    std::map<const float*, HW> hmap32;   // the same weights as 32-column tile images: small row counts (single clips) get 6x the CTAs
    std::map<const float*, HW> hmap64;   // ... and as 64-column images: 3x the CTAs for a handful of clips
    std::map<const float*, HW> hmap96;   // ... and as 96-column images: 2x the CTAs up to half a wave of row tiles
    bool reg_h = false;                  // register_tc also builds the fp16 image
    bool reg_h32 = false;                // ... and the 32-column image (denoiser weights only)
    int register_tc(const float* key, const float* host_wt, int K, int N, int ldw) {
        const int bn = (N % 192 == 0) ? 192 : (N % 128 == 0 ? 128 : (N == 32 ? 32 : 0));
        if (bn == 0 || K % tc::BK != 0) return 0;
        std::vector<float> img;
        tc::pack_weights_tc(host_wt, K, N, ldw, bn, 3, img);
        float* d = nullptr;
        CKI(upload(img, &d));
        tcmap[key] = TcW{d, K, N, bn};
        if (reg_h && K % hx::HBK == 0) {
            CKI(register_h(key, host_wt, K, N, ldw, bn));
            if (reg_h32 && bn != 32 && N % 32 == 0) CKI(register_h(key, host_wt, K, N, ldw, 32, &hmap32));
            if (reg_h32 && bn != 32 && N % 64 == 0) CKI(register_h(key, host_wt, K, N, ldw, 64, &hmap64));
            if (reg_h32 && bn != 32 && N % 96 == 0) CKI(register_h(key, host_wt, K, N, ldw, 96, &hmap96));
        }
        return 0;
    }
    int register_h(const float* key, const float* host_wt, int K, int N, int ldw, int bn, std::map<const float*, HW>* into = nullptr) {
        std::vector<uint16_t> himg;
        const int e = hx::pack_weights_h(host_wt, K, N, ldw, bn, himg);
        uint8_t* hd = nullptr;
        CK(cudaMalloc((void**)&hd, himg.size() * sizeof(uint16_t)));
        arena.push_back(reinterpret_cast<float*>(hd));
        CK(cudaMemcpy(hd, himg.data(), himg.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        (into ? *into : hmap)[key] = HW{hd, K, N, bn, e};
        return 0;
    }
    // few row tiles (single clips: 5 tiles of 128 rows): tile N by 32 instead of 192 so that the launch has 6x the CTAs
    bool small_rows(int M) const { return ((M + hx::HBM - 1) / hx::HBM) * 6 <= num_sms; }   // (N = 32 MMAs cost 68 cycles against 96 for N = 192: only worth it while the 6x CTAs still fit one wave)
    // The fused feed-forward (ffn_h.cuh): out = [geglu(ln W1 + b1) | x2] Wffp + bias + residual of `ep`.  W1's image for this
    // kernel has 256-column n-tiles and is registered under the GEGLU bias pointer.
    bool mid_rows(int M) const { return !small_rows(M) && ((M + hx::HBM - 1) / hx::HBM) * 3 <= num_sms; }   // 25..49 row tiles: 64-column tiles
    bool half_rows(int M) const { return !small_rows(M) && !mid_rows(M) && ((M + hx::HBM - 1) / hx::HBM) * 2 <= num_sms; }   // 50..74 row tiles: 96-column tiles
    // workspace of the fused feed-forward's split leftover tiles (ffn_h.cuh): partial accumulators + counters
    DevBuf ffn_part, ffn_sync;
    static constexpr int FFN_SPLIT_TILES = 16;
    bool ffn_split = getenv("SAID_FFN_NOSPLIT") == nullptr;
    bool lean_epi = getenv("SAID_NO_LEAN_EPI") == nullptr;
    int ffn_min_tiles = getenv("SAID_FFN_MIN_TILES") ? atoi(getenv("SAID_FFN_MIN_TILES")) : 75;   // measured at 16 / 24 / 32 / 48 clips: 75 -> 0.95 / 1.06 / 1.43 / 1.78 ms per step, 148 -> 0.95 / 1.12 / 1.44 / 1.77, never -> 0.98 / 1.13 / 1.45 / 1.83
    int ensure_ffn_split(size_t M) {
        (void)M;   // fixed size: room for FFN_SPLIT_TILES leftover tiles per half-batch (two halves may run at once)
        CK(ffn_sync.ensure_zero(2 * 64));
        CK(ffn_part.ensure((size_t)2 * FFN_SPLIT_TILES * hx::FFN_NP * hx::HBM * hx::FFN_C));
        return 0;
    }
    template <class EP>
    int ffn_h(cudaStream_t st, int M, const __half* pln, const __half* px2, const float* w1key, const float* bias1, const float* wffp,
              EP ep, int tag, int dbg = 0, long long* trace = nullptr) {
        auto i1 = hmap.find(w1key), i2 = hmap.find(wffp);
        if (i1 == hmap.end() || i2 == hmap.end()) return fail("fused feed-forward: weight images not registered");
        const HW &w1 = i1->second, &w2 = i2->second;
        if (w1.K != hx::FFN_C || w1.N != 2 * hx::FFN_NJ * hx::FFN_JC || w1.bn != 256 || w2.K != hx::FFN_NJ * hx::FFN_JC + hx::FFN_C ||
            w2.N != hx::FFN_C || w2.bn != 192)
            return fail("fused feed-forward: unexpected weight geometry");
        hx::FfnParams p;
        memset(&p, 0, sizeof(p));
        if (!hx::make_pair_map(&p.map_ln, pln, hx::FFN_C, M) || !hx::make_pair_map(&p.map_x2, px2, hx::FFN_C, M))
            return fail("cuTensorMapEncodeTiled failed (driver entry point unavailable or bad tensor geometry)");
        p.M = M;
        p.scale1 = std::ldexp(1.0f, -w1.exp);
        p.dbg = trace ? dbg : (dbg & ~16);
        p.trace = trace;
        {
            const int tiles = (M + hx::HBM - 1) / hx::HBM, grid = tiles < num_sms ? tiles : num_sms;
            const int left = tiles - (tiles / grid) * grid;
            p.split = (ffn_split && ffn_sync.p != nullptr && ffn_part.p != nullptr && left > 0 && left <= FFN_SPLIT_TILES) ? 1 : 0;
            p.part = ffn_part.p + (size_t)hctx.half * FFN_SPLIT_TILES * hx::FFN_NP * hx::HBM * hx::FFN_C;
            p.sync = reinterpret_cast<int*>(ffn_sync.p) + hctx.half * 64;
        }
        ep.acc_scale = std::ldexp(1.0f, -w2.exp);
        cur_tag = tag;
        cudaError_t e;
        if constexpr (std::is_same<EP, EpiStd>::value) {
            const int F = lean_epi ? lean_flags(ep, hx::FFN_C) : -1;
            if (F == 9) e = hx::launch_ffn_h(st, num_sms, p, w1.img, w2.img, bias1, status_flag, make_lean<9>(ep, hx::FFN_C), pdl);
            else if (F == 1) e = hx::launch_ffn_h(st, num_sms, p, w1.img, w2.img, bias1, status_flag, make_lean<1>(ep, hx::FFN_C), pdl);
            else e = hx::launch_ffn_h(st, num_sms, p, w1.img, w2.img, bias1, status_flag, ep, pdl);
        } else {
            e = hx::launch_ffn_h(st, num_sms, p, w1.img, w2.img, bias1, status_flag, ep, pdl);
        }
        if (e != cudaSuccess) return fail(std::string("fused feed-forward launch failed: ") + cudaGetErrorString(e));
        return after_launch(st);
    }
    // EpiLean feature set of an EpiStd configuration, or -1 if it needs the general epilogue
    static int lean_flags(const EpiStd& ep, int N) {
        if (!ep.out || ep.acc_in || ep.out_pair || ep.out_period != 0 || ep.act != 0 || N % 192 != 0) return -1;
        if (ep.emb && !ep.step_ptr) return -1;            // per-sample embedding rows (said_denoiser_forward): general epilogue
        if (ep.res_scale && !ep.res) return -1;
        return (ep.res ? 1 : 0) | (ep.res_scale ? 2 : 0) | (ep.emb ? 4 : 0) | (ep.res && ep.res_mod > 0 ? 8 : 0);
    }
    template <int F>
    static EpiLean<F> make_lean(const EpiStd& ep, int N) {
        return EpiLean<F>{ep.out, ep.ldo, N, ep.bias, ep.res, ep.ldr, ep.res_scale, ep.res_shift, ep.T > 0 ? ep.T : 1, ep.res_aff_ld,
                          ep.acc_scale, ep.emb, ep.emb_ld, ep.step_ptr, ep.res_mod > 0 ? ep.res_mod : 1};
    }
    template <int BN>
    bool launch_lean(int F, cudaStream_t st, const hx::HParams& p, const uint8_t* img, const EpiStd& ep, int N, cudaError_t* e) {
        switch (F) {
            case 0: *e = hx::launch_gemm_h<BN>(st, num_sms, p, img, make_lean<0>(ep, N), pdl); return true;
            case 1: *e = hx::launch_gemm_h<BN>(st, num_sms, p, img, make_lean<1>(ep, N), pdl); return true;
            case 3: *e = hx::launch_gemm_h<BN>(st, num_sms, p, img, make_lean<3>(ep, N), pdl); return true;
            case 4: *e = hx::launch_gemm_h<BN>(st, num_sms, p, img, make_lean<4>(ep, N), pdl); return true;
            case 9: *e = hx::launch_gemm_h<BN>(st, num_sms, p, img, make_lean<9>(ep, N), pdl); return true;
            default: return false;
        }
    }
    // One contraction on the fp16x3 path.  The K dimension is the concatenation of `segs`: columns [col0, col0 + ncols) of
    // the pair tensor `src` (C columns, `rows` rows), rows shifted by row_shift (Conv1d taps).  Weight = the image of `wkey`.
    struct HSrc { const __half* base; int C; long long rows; long long pitch_halfs = 0; /* 0: 2 * C */ };
    struct HSegSpec { HSrc src; int col0, ncols, row_shift; };
    template <class EP>
    int gemm_h(cudaStream_t st, int M, int N, std::initializer_list<HSegSpec> segs, const float* wkey, EP ep, int tag, int dbg = 0,
               int k0 = 0 /*first contraction index of this launch within the weight (split-K), multiple of 64*/) {
        auto it = hmap.find(wkey);
        if (it == hmap.end()) return fail("fp16x3 gemm: weight image not registered");
        if (small_rows(M) && it->second.bn != 32) {
            auto it32 = hmap32.find(wkey);
            if (it32 != hmap32.end()) it = it32;
        } else if (mid_rows(M) && it->second.bn == 192) {
            auto it64 = hmap64.find(wkey);
            if (it64 != hmap64.end()) it = it64;
        } else if (half_rows(M) && it->second.bn == 192) {
            auto it96 = hmap96.find(wkey);
            if (it96 != hmap96.end()) it = it96;
        }
        const HW& w = it->second;
        hx::HParams p;
        memset(&p, 0, sizeof(p));
        const __half* bases[3] = {nullptr, nullptr, nullptr};
        int nmaps = 0, nk = 0, ns = 0;
        for (const HSegSpec& sg : segs) {
            if (ns >= hx::H_MAX_SEG) return fail("fp16x3 gemm: too many segments");
            if (sg.ncols % hx::HBK != 0 || sg.col0 % 8 != 0) return fail("fp16x3 gemm: segment not a multiple of 64 columns");
            int mi = -1;
            for (int j = 0; j < nmaps; ++j)
                if (bases[j] == sg.src.base) mi = j;
            if (mi < 0) {
                if (nmaps >= 3) return fail("fp16x3 gemm: too many source tensors");
                if (!hx::make_pair_map(&p.maps[nmaps], sg.src.base, sg.src.C, sg.src.rows, sg.src.pitch_halfs))
                    return fail("cuTensorMapEncodeTiled failed (driver entry point unavailable or bad tensor geometry)");
                bases[nmaps] = sg.src.base;
                mi = nmaps++;
            }
            p.seg[ns++] = hx::HSeg{mi, sg.ncols / hx::HBK, sg.col0, sg.src.C, sg.row_shift};
            nk += sg.ncols / hx::HBK;
        }
        if (k0 % hx::HBK != 0 || k0 + nk * hx::HBK > w.K || (k0 == 0 && nk * hx::HBK != w.K && dbg != -1) || N != w.N)
            return fail("fp16x3 gemm: shape does not match the registered weight");
        if (dbg == -1) dbg = 0;    // -1: a deliberate partial contraction starting at k0 = 0
        p.M = M;
        p.N = N;
        p.nk = nk;
        p.nseg = ns;
        p.nmaps = nmaps;
        p.w_block_bytes = 2 * w.bn * hx::HROW;
        p.w_nk_total = w.K / hx::HBK;
        p.w_kc0 = k0 / hx::HBK;
        p.sliver = 1;
        p.dbg = dbg;
        ep.acc_scale = std::ldexp(1.0f, -w.exp);
        cur_tag = tag;
        cudaError_t e = cudaErrorInvalidValue;
        if constexpr (std::is_same<EP, EpiStd>::value) {
            // the epilogue with its feature set fixed at compile time (EpiLean): EpiStd::store4 tests nine run-time features per
            // call, and on this kernel the epilogue warps' instruction count is on the critical path of every thin-K layer
            int F = lean_epi ? lean_flags(ep, N) : -1;
            if (F >= 0) {
                bool done = true;
                if (w.bn == 192) done = launch_lean<192>(F, st, p, w.img, ep, N, &e);
                else if (w.bn == 96) done = launch_lean<96>(F, st, p, w.img, ep, N, &e);
                else if (w.bn == 64) done = launch_lean<64>(F, st, p, w.img, ep, N, &e);
                else if (w.bn == 32) done = launch_lean<32>(F, st, p, w.img, ep, N, &e);
                else done = false;
                if (!done) F = -1;
                if (F >= 0) {
                    if (e != cudaSuccess) return fail(std::string("fp16x3 gemm launch failed: ") + cudaGetErrorString(e));
                    return after_launch(st);
                }
            }
        }
        if (w.bn == 192) e = hx::launch_gemm_h<192>(st, num_sms, p, w.img, ep, pdl);
        else if (w.bn == 128) e = hx::launch_gemm_h<128>(st, num_sms, p, w.img, ep, pdl);
        else if (w.bn == 96) e = hx::launch_gemm_h<96>(st, num_sms, p, w.img, ep, pdl);
        else if (w.bn == 64) e = hx::launch_gemm_h<64>(st, num_sms, p, w.img, ep, pdl);
        else if (w.bn == 32) e = hx::launch_gemm_h<32>(st, num_sms, p, w.img, ep, pdl);
        if (e != cudaSuccess) return fail(std::string("fp16x3 gemm launch failed: ") + cudaGetErrorString(e));
        return after_launch(st);
    }
    // device-side status word: bit 0 = a pair-format producer saw |x| >= fp16 max (the fp16x3 path cannot represent it)
    int* status_flag = nullptr;
    // alignment band of the cross-attention (attention.py:177-189) for the current (frames, context frames)
    int2* band_dev = nullptr;
    int band_T = -1, band_Tc = -1, band_cap = 0;
    int ensure_band(int T, int Tc, cudaStream_t st) {
        if (T == band_T && Tc == band_Tc) return 0;
        std::vector<int2> hb((size_t)T);
        const double ratio = (double)Tc / (double)T;           // c_x_ratio
        const double half = ratio / 2 + 1;                      // c_kh_size with pad = 1
        for (int i = 0; i < T; ++i) {
            const double mid = (i + 0.5) * ratio;
            long lo = (long)std::nearbyint(mid - half), hi = (long)std::nearbyint(mid + half);   // Python round(): ties to even
            if (lo < 0) lo = 0;
            if (hi > Tc) hi = Tc;
            if (hi - lo < 1 || hi - lo > XATT_MAXW)
                return fail("cross-attention alignment window of " + std::to_string(hi - lo) + " context frames per query (context " +
                            std::to_string(Tc) + " frames, sample " + std::to_string(T) + "): supported 1.." + std::to_string(XATT_MAXW));
            hb[i] = make_int2((int)lo, (int)(hi - lo));
        }
        if (T > band_cap) {
            if (band_dev) cudaFree(band_dev);
            band_dev = nullptr;
            band_cap = 0;
            ++g_alloc_gen;
            CK(cudaMalloc((void**)&band_dev, (size_t)T * sizeof(int2)));
            band_cap = T;
        }
        CK(cudaMemcpyAsync(band_dev, hb.data(), (size_t)T * sizeof(int2), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                          // hb is a temporary
        band_T = T;
        band_Tc = Tc;
        ++g_alloc_gen;                                          // the table content is baked into nothing, but keep graphs honest
        return 0;
    }
    // the tcgen05 kernel's loader threads share a row among ROW_CHUNKS lanes
    static ALoadLNT<tc::ROW_CHUNKS> to_tc_loader(const ALoadLN& a) {
        return ALoadLNT<tc::ROW_CHUNKS>{a.X, a.M, a.T, a.pre_scale, a.pre_shift, a.gamma, a.beta, a.eps};
    }
    static const ALoadPlain& to_tc_loader(const ALoadPlain& a) { return a; }
    static const ALoadConv3& to_tc_loader(const ALoadConv3& a) { return a; }
    template <class AL, class EP>
    int gemm_tc_dispatch(cudaStream_t st, int M, int N, int K, const AL& al, const TcW& w, const EP& ep) {
        const int stride = 2 * w.bn * tc::BK;   // image holds hi + lo tiles
        const bool x3 = precision == 1 || precision == 3;   // fp16x3 mode: the GEMMs that stay on these kernels run 3xTF32
        cudaError_t e = cudaErrorInvalidValue;
        if (a_in_tmem && w.bn != 128) {   // activations through TMEM (TS MMA): experimental, see gemm_tc.cuh
            if (w.bn == 192) {
                e = x3 ? tc::launch_gemm_tca<192, 3>(st, num_sms, M, N, K, al, w.img, stride, ep)
                       : tc::launch_gemm_tca<192, 1>(st, num_sms, M, N, K, al, w.img, stride, ep);
            } else if (w.bn == 32) {
                e = x3 ? tc::launch_gemm_tca<32, 3>(st, num_sms, M, N, K, al, w.img, stride, ep)
                       : tc::launch_gemm_tca<32, 1>(st, num_sms, M, N, K, al, w.img, stride, ep);
            }
        } else if (w.bn == 128) {
            if constexpr (std::is_same<AL, ALoadPlain>::value && std::is_same<EP, EpiStd>::value)   // encoder conv stack (N = 512)
                e = x3 ? tc::launch_gemm_tc<128, 3>(st, num_sms, M, N, K, al, w.img, stride, ep, 0, pdl)
                       : tc::launch_gemm_tc<128, 1>(st, num_sms, M, N, K, al, w.img, stride, ep, 0, pdl);
        } else if (w.bn == 192) {
            e = x3 ? tc::launch_gemm_tc<192, 3>(st, num_sms, M, N, K, al, w.img, stride, ep, 0, pdl)
                   : tc::launch_gemm_tc<192, 1>(st, num_sms, M, N, K, al, w.img, stride, ep, 0, pdl);
        } else if (w.bn == 32) {
            e = x3 ? tc::launch_gemm_tc<32, 3>(st, num_sms, M, N, K, al, w.img, stride, ep, 0, pdl)
                   : tc::launch_gemm_tc<32, 1>(st, num_sms, M, N, K, al, w.img, stride, ep, 0, pdl);
        }
        if (e != cudaSuccess) return fail(std::string("tcgen05 gemm launch failed: ") + cudaGetErrorString(e));
        return after_launch(st);
    }
    int enc_precision = 0;               // audio encoder GEMMs: IEEE fp32 FFMA by default (exact parity with the goldens)
    int a_in_tmem = 0;                   // 1: activations through TMEM (TS MMA, gemm_tca_kernel) -- measured slower, kept for study
    static constexpr int TC_MIN_ROWS_DEFAULT = 2048, H_MIN_ROWS_DEFAULT = 512;
    int tc_min_rows = TC_MIN_ROWS_DEFAULT;   // 3xTF32 / TF32 denoiser and every tensor-core encoder: below this many rows the small-tile FFMA kernel spreads better over the SMs
    int h_min_rows = H_MIN_ROWS_DEFAULT;     // fp16x3 denoiser: its 32-column tile images keep the tensor-core path ahead down to a single 5 s clip
                                             // (602 rows: 0.54 vs 0.82 ms per step; two clips: 0.55 vs 1.06); shorter inputs stay on the IEEE fp32 kernels
    template <class AL, class EP>
    int gemm(cudaStream_t st, int M, int N, int K, const AL& al, const float* Wt, int ldw, const EP& ep, int batch = 1,
             int wz_mod = 1, long long w_zstride = 0) {
        cur_tag = AL::kTag;
        if (precision != 0 && batch == 1 && M >= tc_min_rows) {
            auto it = tcmap.find(Wt);
            if (it != tcmap.end() && it->second.K == K && it->second.N == N)
                return gemm_tc_dispatch(st, M, N, K, to_tc_loader(al), it->second, ep);
        }
        cudaError_t e = launch_gemm(st, M, N, K, al, Wt, ldw, ep, batch, wz_mod, w_zstride, pdl);
        if (e != cudaSuccess) return fail(std::string("gemm launch failed: ") + cudaGetErrorString(e));
        return after_launch(st);
    }

    std::map<std::string, HostTensor> raw;
    bool ready = false;
    std::vector<float*> arena;   // device allocations holding packed weights

    // ---- configuration (inferred at commit) ----
    int in_ch = 32, ctx_dim = 768;
    int enc_hidden = 768, enc_layers = 12, enc_heads = 12, enc_ffn = 3072, enc_conv_dim = 512;
    int n_conv = 7, conv_k[8] = {10, 3, 3, 3, 3, 2, 2, 0}, conv_s[8] = {5, 2, 2, 2, 2, 2, 2, 0};
    int pos_k = 128, pos_g = 16;
    int proj_dim = 0;            // audio_proj_layer output width (0: absent)

    // ---- packed denoiser weights ----
    TimeEmbedWeights te{};
    float *w_in = nullptr, *b_in = nullptr;
    ResBlockW rb[5];
    TransformerW tr[4];
    float *out_gn_g = nullptr, *out_gn_b = nullptr, *w_out = nullptr, *b_out = nullptr;
    float *w_kv = nullptr;       // (ctx_dim, 4*384): per block [K(192) | V(192)]
    float *null_emb = nullptr;   // (ctx_dim)
    float *w_aproj = nullptr, *b_aproj = nullptr;

    // ---- packed encoder weights ----
    float *c0_w = nullptr, *c0_g = nullptr, *c0_b = nullptr;
    // wav2vec2-large family (feat_extract_norm="layer", conv_bias, do_stable_layer_norm): per-conv bias + LayerNorm, pre-LN layers
    bool enc_fe_layer_norm = false, enc_stable_ln = false;
    float *c0_bias = nullptr, *conv_bias_p[8] = {nullptr}, *conv_ln_g[8] = {nullptr}, *conv_ln_b[8] = {nullptr};
    float *conv_w[8] = {nullptr};
    float *fp_ln_g = nullptr, *fp_ln_b = nullptr, *fp_w = nullptr, *fp_b = nullptr;
    float *pos_w = nullptr, *pos_b = nullptr, *enc_ln_g = nullptr, *enc_ln_b = nullptr;
    std::vector<EncLayerW> enc;

    // ---- workspaces ----
    DevBuf cnull;
    DevBuf act[7], gnbuf, qkv, ao, q2, ffb, eps, ss, ss_st, emb_tab, tvals, step_tab, kv, vnull, lat, init_lat, vnull_tmp;
    DevBuf e_a, e_b, e_c, e_d, e_qkv, e_ff, e_xp, e_emb;
    // pair-format (fp16 hi/lo) operands of the fp16x3 path, sized in floats (a pair element takes 4 bytes like an fp32 one)
    DevBuf p_gn, p_raw, p_ln, p_ao, p_x2, p_ff;
    DevBuf pe_a, pe_b, pe_x, pe_att, pe_ff;     // encoder, fp16x3 mode: conv ping-pong, layer input, attention output, FFN intermediate
    double* c0_partial = nullptr;
    double* gn_partial = nullptr;
    size_t gn_partial_cap = 0;
    size_t c0_partial_cap = 0;
    int* step_ctr = nullptr;
    int ctx_B = 0, ctx_T = 0, ctx_uncond = 0;
    // The instantiated step graph is kept across denoise() calls and replayed as long as everything baked into its
    // kernel nodes is unchanged (shapes, scalars, user tensors read per step, workspace addresses, precision).
    struct GraphKey {
        int B, T, Tc, do_cfg, n_steps, pred_type, scheduler, precision, tc_min_rows;
        float gscale, grescale, latent_scale;
        const void *eta_noise, *edit_noise, *mask, *intermediates;
        unsigned long long alloc_gen;
    };
    GraphKey graph_key{};
    long long graph_launches = 0;     // kernel nodes in the cached graph
    long long graph_captures = 0;     // how many times a step graph was captured (tests: the cache works)
    cudaGraphExec_t graph_exec = nullptr;
    void drop_graph() {
        if (graph_exec) {
            cudaGraphExecDestroy(graph_exec);
            graph_exec = nullptr;
        }
    }
    DevBuf result_buf;

    ~said_engine() {
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        for (float* p : arena) cudaFree(p);
        for (float* p : bcv_arena) cudaFree(p);
        if (c0_partial) cudaFree(c0_partial);
        if (gn_partial) cudaFree(gn_partial);
        if (step_ctr) cudaFree(step_ctr);
        if (status_flag) cudaFree(status_flag);
        if (band_dev) cudaFree(band_dev);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
        if (own_stream) cudaStreamDestroy(own_stream);
        if (stream2) cudaStreamDestroy(stream2);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }

    // ------------------------------------------------------------------ weights
    const HostTensor* find(const std::string& name) const {
        auto it = raw.find(name);
        return it == raw.end() ? nullptr : &it->second;
    }
    int need(const std::string& name, std::initializer_list<int64_t> shape, const HostTensor** out) {
        const HostTensor* t = find(name);
        if (!t) return fail("missing tensor '" + name + "'");
        std::vector<int64_t> want(shape);
        if (t->shape != want) {
            std::string s = "tensor '" + name + "' has shape (";
            for (auto d : t->shape) s += std::to_string(d) + ",";
            s += ") expected (";
            for (auto d : want) s += std::to_string(d) + ",";
            return fail(s + ")");
        }
        *out = t;
        return 0;
    }
    int upload(const std::vector<float>& h, float** dptr) {
        float* d = nullptr;
        CK(cudaMalloc((void**)&d, std::max<size_t>(h.size(), 4) * sizeof(float)));
        arena.push_back(d);
        CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        *dptr = d;
        return 0;
    }
    int upload_raw(const std::string& name, std::initializer_list<int64_t> shape, float** dptr) {
        const HostTensor* t;
        CKI(need(name, shape, &t));
        return upload(t->data, dptr);
    }
    int commit();
    int commit_denoiser();
    int commit_encoder();
    // ---- evaluation embedder (BCVAE encoder, said/model/vae.py) : separate weight set, separate commit
    BcvaeWeights bcv{};
    bool bcv_ready = false;
    std::vector<float*> bcv_arena;
    int commit_bcvae();

    // ------------------------------------------------------------------ programs
    int encode_audio(const float* wave, int B, int T_a, int T, float* emb_out, cudaStream_t st);
    int encode_audio_h(const float* wave, int B, int T_a, int T, float* emb_out, cudaStream_t st);
    int prepare_context(const float* emb, int B, int T, int with_uncond, cudaStream_t st);
    int ensure_denoiser_ws(int Bp, int T);
    int forward(cudaStream_t st, const float* x, int src_batch, int Bp, int n_uncond, int T, const float* emb_table,
                const int* step_ptr, float* eps_out, float* taps);
    int forward_h(cudaStream_t st, const float* x, int src_batch, int Bp, int n_uncond, int T, const float* emb_table,
                  const int* step_ptr, float* eps_out, float* taps);
    bool enc_split_k = getenv("SAID_ENC_NO_SPLITK") == nullptr;   // fp16x3 encoder: contractions longer than 768 run as split-K launches (accuracy)
    bool fused_ffn = getenv("SAID_NO_FUSED_FFN") == nullptr;   // fp16x3 path: GEGLU + ff2 + proj_out as one kernel (ffn_h.cuh); the env switch keeps the two-GEMM form for A/B runs
    bool attn_h = getenv("SAID_ATTN_TF32") == nullptr;   // fp16x3 path: flash-style fp16 hi/lo attention (attention_h.cuh); the env switch keeps the 3xTF32 kernel reachable for A/B runs
    bool use_h(int M) const { return precision == 3 && M >= h_min_rows && in_ch == 32; }
    int denoise(const said_denoise_args& a, cudaStream_t user);
};

// =====================================================================================================
// Weight packing (host side, one-off)
// =====================================================================================================
namespace {

// Linear / Conv1d weight (Co, Ci, KT) -> K-major "Wt": dst[(row_off + tap*Ci + ci) * ldd + col_off + co]
void pack_w(const HostTensor& w, int Co, int Ci, int KT, std::vector<float>& dst, int ldd, int row_off, int col_off) {
    const float* s = w.data.data();
    for (int co = 0; co < Co; ++co)
        for (int ci = 0; ci < Ci; ++ci)
            for (int tap = 0; tap < KT; ++tap)
                dst[(size_t)(row_off + tap * Ci + ci) * ldd + col_off + co] = s[((size_t)co * Ci + ci) * KT + tap];
}

}  // namespace

int said_engine::commit_denoiser() {
    const std::string P = "denoiser.model.";
    const HostTensor* t;
    // ---- shapes -> configuration
    const HostTensor* win = find(P + "input_blocks.0.0.weight");
    if (!win || win->shape.size() != 3 || win->shape[0] != C || win->shape[2] != 3)
        return fail("denoiser.model.input_blocks.0.0.weight missing or not (192, in_channels, 3)");
    in_ch = (int)win->shape[1];
    if (in_ch % 4 != 0 || (3 * in_ch) % GEMM_BK != 0) return fail("in_channels must make 3*in_channels a multiple of 16");
    const HostTensor* wk = find(P + "input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight");
    if (!wk || wk->shape.size() != 2 || wk->shape[0] != C) return fail("attn2.to_k.weight missing or malformed");
    ctx_dim = (int)wk->shape[1];
    if (ctx_dim % GEMM_BK != 0) return fail("context dim must be a multiple of 16");

    // ---- time embedding
    CKI(upload_raw("time_freqs", {C / 2}, (float**)&te.freqs));
    CKI(upload_raw(P + "time_embed.0.weight", {TE, C}, (float**)&te.w1));
    CKI(upload_raw(P + "time_embed.0.bias", {TE}, (float**)&te.b1));
    CKI(upload_raw(P + "time_embed.2.weight", {TE, TE}, (float**)&te.w2));
    CKI(upload_raw(P + "time_embed.2.bias", {TE}, (float**)&te.b2));

    // ---- input conv
    {
        std::vector<float> w((size_t)3 * in_ch * C);
        pack_w(*win, C, in_ch, 3, w, C, 0, 0);
        CKI(upload(w, &w_in));
        CKI(register_tc(w_in, w.data(), 3 * in_ch, C, C));
        CKI(upload_raw(P + "input_blocks.0.0.bias", {C}, &b_in));
    }
    // ---- ResBlocks in execution order (openaimodel.py:697-704)
    const char* rb_paths[5] = {"input_blocks.1.0", "middle_block.0", "middle_block.2", "output_blocks.0.0", "output_blocks.1.0"};
    const int rb_cin[5] = {C, C, C, 2 * C, 2 * C};
    for (int i = 0; i < 5; ++i) {
        ResBlockW& r = rb[i];
        const std::string p = P + rb_paths[i] + ".";
        r.cin = rb_cin[i];
        r.skip = r.cin != C;
        CKI(upload_raw(p + "in_layers.0.weight", {r.cin}, &r.gn1_g));
        CKI(upload_raw(p + "in_layers.0.bias", {r.cin}, &r.gn1_b));
        CKI(need(p + "in_layers.2.weight", {C, r.cin, 3}, &t));
        std::vector<float> w1((size_t)3 * r.cin * C);
        pack_w(*t, C, r.cin, 3, w1, C, 0, 0);
        CKI(upload(w1, &r.w1));
        CKI(register_tc(r.w1, w1.data(), 3 * r.cin, C, C));
        CKI(upload_raw(p + "in_layers.2.bias", {C}, &r.b1));
        CKI(upload_raw(p + "emb_layers.1.weight", {C, TE}, (float**)&te.wr[i]));
        CKI(upload_raw(p + "emb_layers.1.bias", {C}, (float**)&te.br[i]));
        CKI(upload_raw(p + "out_layers.0.weight", {C}, &r.gn2_g));
        CKI(upload_raw(p + "out_layers.0.bias", {C}, &r.gn2_b));
        CKI(need(p + "out_layers.3.weight", {C, C, 3}, &t));
        r.k2 = 3 * C + (r.skip ? r.cin : 0);
        std::vector<float> w2((size_t)r.k2 * C);
        pack_w(*t, C, C, 3, w2, C, 0, 0);
        CKI(need(p + "out_layers.3.bias", {C}, &t));
        std::vector<float> b2 = t->data;
        if (r.skip) {   // 1x1 skip_connection fused as extra contraction rows; biases add
            CKI(need(p + "skip_connection.weight", {C, r.cin, 1}, &t));
            pack_w(*t, C, r.cin, 1, w2, C, 3 * C, 0);
            CKI(need(p + "skip_connection.bias", {C}, &t));
            for (int j = 0; j < C; ++j) b2[j] += t->data[j];
        }
        CKI(upload(w2, &r.w2));
        CKI(register_tc(r.w2, w2.data(), 3 * C, C, C));
        if (r.skip) CKI(register_tc(r.w2 + (size_t)3 * C * C, w2.data() + (size_t)3 * C * C, r.cin, C, C));
        CKI(upload(b2, &r.b2));
        if (r.skip) {   // fp16x3 path: second conv + 1x1 skip as one K = 576 + 384 contraction; image keyed by the (unique) bias pointer
            CKI(register_h(r.b2, w2.data(), r.k2, C, C, 192));
            CKI(register_h(r.b2, w2.data(), r.k2, C, C, 32, &hmap32));
            CKI(register_h(r.b2, w2.data(), r.k2, C, C, 64, &hmap64));
            CKI(register_h(r.b2, w2.data(), r.k2, C, C, 96, &hmap96));
        }
    }
    // ---- SpatialTransformers in execution order
    const char* tr_paths[4] = {"input_blocks.1.1", "middle_block.1", "output_blocks.0.1", "output_blocks.1.1"};
    std::vector<float> wkv((size_t)ctx_dim * 4 * 2 * C);
    for (int i = 0; i < 4; ++i) {
        TransformerW& s = tr[i];
        const std::string p = P + tr_paths[i] + ".";
        const std::string b = p + "transformer_blocks.0.";
        CKI(upload_raw(p + "norm.weight", {C}, &s.gn_g));
        CKI(upload_raw(p + "norm.bias", {C}, &s.gn_b));
        CKI(upload_raw(b + "norm1.weight", {C}, &s.ln1_g));
        CKI(upload_raw(b + "norm1.bias", {C}, &s.ln1_b));
        CKI(upload_raw(b + "norm2.weight", {C}, &s.ln2_g));
        CKI(upload_raw(b + "norm2.bias", {C}, &s.ln2_b));
        CKI(upload_raw(b + "norm3.weight", {C}, &s.ln3_g));
        CKI(upload_raw(b + "norm3.bias", {C}, &s.ln3_b));
        std::vector<float> qkvw((size_t)C * 3 * C);
        const char* nm[3] = {"attn1.to_q.weight", "attn1.to_k.weight", "attn1.to_v.weight"};
        for (int j = 0; j < 3; ++j) {
            CKI(need(b + nm[j], {C, C}, &t));
            pack_w(*t, C, C, 1, qkvw, 3 * C, 0, j * C);
        }
        CKI(upload(qkvw, &s.wqkv));
        CKI(register_tc(s.wqkv, qkvw.data(), C, 3 * C, 3 * C));
        std::vector<float> w((size_t)C * C);
        CKI(need(b + "attn1.to_out.0.weight", {C, C}, &t));
        pack_w(*t, C, C, 1, w, C, 0, 0);
        CKI(upload(w, &s.wo1));
        CKI(register_tc(s.wo1, w.data(), C, C, C));
        CKI(upload_raw(b + "attn1.to_out.0.bias", {C}, &s.bo1));
        CKI(need(b + "attn2.to_q.weight", {C, C}, &t));
        pack_w(*t, C, C, 1, w, C, 0, 0);
        CKI(upload(w, &s.wq2));
        CKI(register_tc(s.wq2, w.data(), C, C, C));
        CKI(need(b + "attn2.to_out.0.weight", {C, C}, &t));
        pack_w(*t, C, C, 1, w, C, 0, 0);
        CKI(upload(w, &s.wo2));
        CKI(register_tc(s.wo2, w.data(), C, C, C));
        CKI(upload_raw(b + "attn2.to_out.0.bias", {C}, &s.bo2));
        CKI(need(b + "attn2.to_k.weight", {C, ctx_dim}, &t));
        pack_w(*t, C, ctx_dim, 1, wkv, 8 * C, 0, i * 2 * C);
        CKI(need(b + "attn2.to_v.weight", {C, ctx_dim}, &t));
        pack_w(*t, C, ctx_dim, 1, wkv, 8 * C, 0, i * 2 * C + C);
        // GEGLU projection: interleave (value_j, gate_j) columns (attention.py:31-32: value = first half)
        CKI(need(b + "ff.net.0.proj.weight", {2 * FF, C}, &t));
        std::vector<float> wff1((size_t)C * 2 * FF), bff1((size_t)2 * FF);
        for (int o = 0; o < 2 * FF; ++o) {
            const int col = o < FF ? 2 * o : 2 * (o - FF) + 1;
            for (int k = 0; k < C; ++k) wff1[(size_t)k * 2 * FF + col] = t->data[(size_t)o * C + k];
        }
        CKI(upload(wff1, &s.wff1));
        CKI(register_tc(s.wff1, wff1.data(), C, 2 * FF, 2 * FF));
        CKI(need(b + "ff.net.0.proj.bias", {2 * FF}, &t));
        for (int o = 0; o < 2 * FF; ++o) bff1[o < FF ? 2 * o : 2 * (o - FF) + 1] = t->data[o];
        CKI(upload(bff1, &s.bff1));
        if (reg_h && C == hx::FFN_C && FF == hx::FFN_NJ * hx::FFN_JC) CKI(register_h(s.bff1, wff1.data(), C, 2 * FF, 2 * FF, 256));
        CKI(need(b + "ff.net.2.weight", {C, FF}, &t));
        std::vector<float> wff2((size_t)FF * C);
        pack_w(*t, C, FF, 1, wff2, C, 0, 0);
        CKI(upload(wff2, &s.wff2));
        CKI(register_tc(s.wff2, wff2.data(), FF, C, C));
        CKI(upload_raw(b + "ff.net.2.bias", {C}, &s.bff2));
        CKI(need(p + "proj_out.weight", {C, C, 1}, &t));
        pack_w(*t, C, C, 1, w, C, 0, 0);
        CKI(upload(w, &s.wproj));
        CKI(register_tc(s.wproj, w.data(), C, C, C));
        CKI(upload_raw(p + "proj_out.bias", {C}, &s.bproj));
        {   // out = (ff W2 + b2 + x) Wp + bp + h  ==  [ff | x] [W2 Wp ; Wp] + (b2 Wp + bp) + h   (products in fp64)
            std::vector<float> wc((size_t)(FF + C) * C), bc((size_t)C);
            const std::vector<float>& wp = w;   // (192 k, 192 n) K-major
            for (int k = 0; k < FF; ++k)
                for (int n = 0; n < C; ++n) {
                    double acc = 0.0;
                    for (int j = 0; j < C; ++j) acc += (double)wff2[(size_t)k * C + j] * (double)wp[(size_t)j * C + n];
                    wc[(size_t)k * C + n] = (float)acc;
                }
            std::copy(wp.begin(), wp.end(), wc.begin() + (size_t)FF * C);
            const HostTensor *b2t, *bpt;
            CKI(need(b + "ff.net.2.bias", {C}, &b2t));
            CKI(need(p + "proj_out.bias", {C}, &bpt));
            for (int n = 0; n < C; ++n) {
                double acc = bpt->data[n];
                for (int j = 0; j < C; ++j) acc += (double)b2t->data[j] * (double)wp[(size_t)j * C + n];
                bc[n] = (float)acc;
            }
            CKI(upload(wc, &s.wffp));
            CKI(register_tc(s.wffp, wc.data(), FF + C, C, C));
            CKI(upload(bc, &s.bffp));
        }
    }
    CKI(upload(wkv, &w_kv));
    CKI(register_tc(w_kv, wkv.data(), ctx_dim, 8 * C, 8 * C));
    // ---- output head
    CKI(upload_raw(P + "out.0.weight", {C}, &out_gn_g));
    CKI(upload_raw(P + "out.0.bias", {C}, &out_gn_b));
    CKI(need(P + "out.2.weight", {in_ch, C, 3}, &t));
    std::vector<float> wo((size_t)3 * C * in_ch);
    pack_w(*t, in_ch, C, 3, wo, in_ch, 0, 0);
    CKI(upload(wo, &w_out));
    CKI(register_tc(w_out, wo.data(), 3 * C, in_ch, in_ch));
    CKI(upload_raw(P + "out.2.bias", {in_ch}, &b_out));
    // ---- null condition, optional audio projection
    CKI(need("null_cond_emb", {1, 1, ctx_dim}, &t));
    CKI(upload(t->data, &null_emb));
    proj_dim = 0;
    if (const HostTensor* pw = find("audio_proj_layer.weight")) {
        if (pw->shape.size() != 2) return fail("audio_proj_layer.weight malformed");
        proj_dim = (int)pw->shape[0];
        const int kin = (int)pw->shape[1];
        if (proj_dim != ctx_dim) return fail("audio_proj_layer output width != denoiser context dim");
        std::vector<float> w((size_t)kin * proj_dim);
        pack_w(*pw, proj_dim, kin, 1, w, proj_dim, 0, 0);
        CKI(upload(w, &w_aproj));
        CKI(upload_raw("audio_proj_layer.bias", {proj_dim}, &b_aproj));
    }
    return 0;
}

int said_engine::commit_encoder() {
    const std::string P = "audio_encoder.";
    const HostTensor* t;
    // ---- feature encoder (TF modeling_wav2vec2.py:254-323, 382-419)
    n_conv = 0;
    while (find(P + "feature_extractor.conv_layers." + std::to_string(n_conv) + ".conv.weight")) ++n_conv;
    if (n_conv != 7) return fail("audio encoder: expected 7 conv feature layers (wav2vec2-base family), found " + std::to_string(n_conv));
    const int ks[7] = {10, 3, 3, 3, 3, 2, 2};
    CKI(need(P + "feature_extractor.conv_layers.0.conv.weight", {C0_CH, 1, C0_K}, &t));
    enc_conv_dim = C0_CH;
    {
        std::vector<float> w((size_t)C0_K * C0_CH);
        for (int c = 0; c < C0_CH; ++c)
            for (int k = 0; k < C0_K; ++k) w[(size_t)k * C0_CH + c] = t->data[(size_t)c * C0_K + k];
        CKI(upload(w, &c0_w));
    }
    // feat_extract_norm: "group" (base: GroupNorm on layer 0 only, no conv bias) or "layer" (large: bias + LayerNorm over
    // channels after every conv).  The state dict tells them apart; do_stable_layer_norm does not show in the names and
    // arrives as the pseudo tensor "audio_encoder.config.do_stable_layer_norm" (set by the Python layer from the config).
    enc_fe_layer_norm = find(P + "feature_extractor.conv_layers.1.layer_norm.weight") != nullptr;
    const bool has_conv_bias = find(P + "feature_extractor.conv_layers.0.conv.bias") != nullptr;
    if (enc_fe_layer_norm != has_conv_bias)
        return fail("audio encoder: supported feature extractors are (group norm, no conv bias) and (layer norm, conv bias)");
    {
        const HostTensor* sl = find(P + "config.do_stable_layer_norm");
        enc_stable_ln = sl && sl->numel() == 1 && sl->data[0] != 0.f;
    }
    CKI(upload_raw(P + "feature_extractor.conv_layers.0.layer_norm.weight", {C0_CH}, &c0_g));
    CKI(upload_raw(P + "feature_extractor.conv_layers.0.layer_norm.bias", {C0_CH}, &c0_b));
    if (enc_fe_layer_norm) {
        CKI(upload_raw(P + "feature_extractor.conv_layers.0.conv.bias", {C0_CH}, &c0_bias));
        for (int i = 1; i < 7; ++i) {
            const std::string q = P + "feature_extractor.conv_layers." + std::to_string(i) + ".";
            CKI(upload_raw(q + "conv.bias", {C0_CH}, &conv_bias_p[i]));
            CKI(upload_raw(q + "layer_norm.weight", {C0_CH}, &conv_ln_g[i]));
            CKI(upload_raw(q + "layer_norm.bias", {C0_CH}, &conv_ln_b[i]));
        }
    }
    for (int i = 1; i < 7; ++i) {
        conv_k[i] = ks[i];
        CKI(need(P + "feature_extractor.conv_layers." + std::to_string(i) + ".conv.weight", {C0_CH, C0_CH, ks[i]}, &t));
        std::vector<float> w((size_t)ks[i] * C0_CH * C0_CH);
        pack_w(*t, C0_CH, C0_CH, ks[i], w, C0_CH, 0, 0);
        CKI(upload(w, &conv_w[i]));
        CKI(register_tc(conv_w[i], w.data(), ks[i] * C0_CH, C0_CH, C0_CH));
    }
    // ---- feature projection
    const HostTensor* pw = find(P + "feature_projection.projection.weight");
    if (!pw || pw->shape.size() != 2 || pw->shape[1] != C0_CH) return fail("feature_projection.projection.weight missing or malformed");
    enc_hidden = (int)pw->shape[0];
    const int H = enc_hidden;
    if (H % 64 != 0 || H > 1024) return fail("audio encoder: hidden size must be a multiple of 64 and <= 1024");
    enc_heads = H / 64;
    CKI(upload_raw(P + "feature_projection.layer_norm.weight", {C0_CH}, &fp_ln_g));
    CKI(upload_raw(P + "feature_projection.layer_norm.bias", {C0_CH}, &fp_ln_b));
    {
        std::vector<float> w((size_t)C0_CH * H);
        pack_w(*pw, H, C0_CH, 1, w, H, 0, 0);
        CKI(upload(w, &fp_w));
        CKI(register_tc(fp_w, w.data(), C0_CH, H, H));
        CKI(upload_raw(P + "feature_projection.projection.bias", {H}, &fp_b));
    }
    // ---- positional conv: fold weight norm (w = v * g / ||v||, norm over (out, in) per tap), regroup
    {
        const std::string pc = P + "encoder.pos_conv_embed.conv.";
        const HostTensor* g = find(pc + "weight_g");
        const HostTensor* v = find(pc + "weight_v");
        if (!g) g = find(pc + "parametrizations.weight.original0");
        if (!v) v = find(pc + "parametrizations.weight.original1");
        if (!g || !v || v->shape.size() != 3 || v->shape[0] != H) return fail("pos_conv_embed weight_g / weight_v missing or malformed");
        const int cg = (int)v->shape[1];
        pos_k = (int)v->shape[2];
        if (H % cg != 0) return fail("pos_conv_embed: hidden not divisible by group width");
        pos_g = H / cg;
        if (g->numel() != pos_k) return fail("pos_conv_embed weight_g must have one entry per kernel tap");
        if (pos_k % 2 != 0) return fail("pos_conv_embed: odd kernel sizes are not supported");
        if (cg % 4 != 0 || (pos_k * cg) % GEMM_BK != 0) return fail("pos_conv_embed: unsupported group width");
        std::vector<double> nrm(pos_k, 0.0);
        for (int co = 0; co < H; ++co)
            for (int ci = 0; ci < cg; ++ci)
                for (int k = 0; k < pos_k; ++k) {
                    const double x = v->data[((size_t)co * cg + ci) * pos_k + k];
                    nrm[k] += x * x;
                }
        std::vector<float> w((size_t)pos_g * pos_k * cg * cg);
        for (int gi = 0; gi < pos_g; ++gi)
            for (int co = 0; co < cg; ++co)
                for (int ci = 0; ci < cg; ++ci)
                    for (int k = 0; k < pos_k; ++k) {
                        const float fac = g->data[k] / (float)std::sqrt(nrm[k]);
                        w[((size_t)gi * pos_k * cg + (size_t)k * cg + ci) * cg + co] =
                            v->data[((size_t)(gi * cg + co) * cg + ci) * pos_k + k] * fac;
                    }
        CKI(upload(w, &pos_w));
        CKI(upload_raw(pc + "bias", {H}, &pos_b));
    }
    CKI(upload_raw(P + "encoder.layer_norm.weight", {H}, &enc_ln_g));
    CKI(upload_raw(P + "encoder.layer_norm.bias", {H}, &enc_ln_b));
    // ---- transformer layers (post-LN; TF modeling_wav2vec2.py:576-609)
    enc_layers = 0;
    while (find(P + "encoder.layers." + std::to_string(enc_layers) + ".final_layer_norm.weight")) ++enc_layers;
    if (enc_layers == 0) return fail("audio encoder: no transformer layers found");
    const HostTensor* f1 = find(P + "encoder.layers.0.feed_forward.intermediate_dense.weight");
    if (!f1 || f1->shape.size() != 2) return fail("intermediate_dense.weight missing");
    enc_ffn = (int)f1->shape[0];
    if (enc_ffn % GEMM_BK != 0) return fail("audio encoder: ffn width must be a multiple of 16");
    enc.assign(enc_layers, EncLayerW{});
    for (int l = 0; l < enc_layers; ++l) {
        EncLayerW& L = enc[l];
        const std::string p = P + "encoder.layers." + std::to_string(l) + ".";
        std::vector<float> wqkv((size_t)H * 3 * H), bqkv((size_t)3 * H);
        const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
        for (int j = 0; j < 3; ++j) {
            CKI(need(p + "attention." + nm[j] + ".weight", {H, H}, &t));
            pack_w(*t, H, H, 1, wqkv, 3 * H, 0, j * H);
            CKI(need(p + "attention." + nm[j] + ".bias", {H}, &t));
            std::copy(t->data.begin(), t->data.end(), bqkv.begin() + (size_t)j * H);
        }
        CKI(upload(wqkv, &L.wqkv));
        CKI(register_tc(L.wqkv, wqkv.data(), H, 3 * H, 3 * H));
        CKI(upload(bqkv, &L.bqkv));
        std::vector<float> w((size_t)H * H);
        CKI(need(p + "attention.out_proj.weight", {H, H}, &t));
        pack_w(*t, H, H, 1, w, H, 0, 0);
        CKI(upload(w, &L.wo));
        CKI(register_tc(L.wo, w.data(), H, H, H));
        CKI(upload_raw(p + "attention.out_proj.bias", {H}, &L.bo));
        CKI(upload_raw(p + "layer_norm.weight", {H}, &L.ln1_g));
        CKI(upload_raw(p + "layer_norm.bias", {H}, &L.ln1_b));
        CKI(need(p + "feed_forward.intermediate_dense.weight", {enc_ffn, H}, &t));
        std::vector<float> w1((size_t)H * enc_ffn);
        pack_w(*t, enc_ffn, H, 1, w1, enc_ffn, 0, 0);
        CKI(upload(w1, &L.wff1));
        CKI(register_tc(L.wff1, w1.data(), H, enc_ffn, enc_ffn));
        CKI(upload_raw(p + "feed_forward.intermediate_dense.bias", {enc_ffn}, &L.bff1));
        CKI(need(p + "feed_forward.output_dense.weight", {H, enc_ffn}, &t));
        std::vector<float> w2((size_t)enc_ffn * H);
        pack_w(*t, H, enc_ffn, 1, w2, H, 0, 0);
        CKI(upload(w2, &L.wff2));
        CKI(register_tc(L.wff2, w2.data(), enc_ffn, H, H));
        CKI(upload_raw(p + "feed_forward.output_dense.bias", {H}, &L.bff2));
        CKI(upload_raw(p + "final_layer_norm.weight", {H}, &L.ln2_g));
        CKI(upload_raw(p + "final_layer_norm.bias", {H}, &L.ln2_b));
    }
    return 0;
}

int said_engine::commit_bcvae() {
    CK(cudaSetDevice(device));
    CK(cudaDeviceSynchronize());
    for (float* p : bcv_arena) cudaFree(p);
    bcv_arena.clear();
    bcv_ready = false;
    const std::string P = "bcvae.encoder.";
    auto up = [&](const std::vector<float>& h, const float** d) -> int {
        float* p = nullptr;
        CK(cudaMalloc((void**)&p, std::max<size_t>(h.size(), 4) * sizeof(float)));
        bcv_arena.push_back(p);
        CK(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
        *d = p;
        return 0;
    };
    auto raw_t = [&](const std::string& name, std::initializer_list<int64_t> shape, const float** d) -> int {
        const HostTensor* t;
        CKI(need(P + name, shape, &t));
        return up(t->data, d);
    };
    // BatchNorm1d in eval mode -> per-channel scale / shift (vae.py:43-64; eps 1e-5)
    auto bn = [&](const std::string& name, int64_t c, const float** sc, const float** sh) -> int {
        const HostTensor *g, *b, *rm, *rv;
        CKI(need(P + name + ".weight", {c}, &g));
        CKI(need(P + name + ".bias", {c}, &b));
        CKI(need(P + name + ".running_mean", {c}, &rm));
        CKI(need(P + name + ".running_var", {c}, &rv));
        std::vector<float> s(c), h(c);
        for (int64_t i = 0; i < c; ++i) {
            const double k = (double)g->data[i] / std::sqrt((double)rv->data[i] + 1e-5);
            s[i] = (float)k;
            h[i] = (float)((double)b->data[i] - (double)rm->data[i] * k);
        }
        CKI(up(s, sc));
        return up(h, sh);
    };
    CKI(raw_t("conv_layers.0.weight", {32, BCV_IN, 3}, &bcv.c1w)); CKI(raw_t("conv_layers.0.bias", {32}, &bcv.c1b)); CKI(bn("conv_layers.1", 32, &bcv.bn1s, &bcv.bn1h));
    CKI(raw_t("conv_layers.3.weight", {64, 32, 3}, &bcv.c2w));     CKI(raw_t("conv_layers.3.bias", {64}, &bcv.c2b)); CKI(bn("conv_layers.4", 64, &bcv.bn2s, &bcv.bn2h));
    CKI(raw_t("conv_layers.6.weight", {64, 64, 4}, &bcv.c3w));     CKI(raw_t("conv_layers.6.bias", {64}, &bcv.c3b)); CKI(bn("conv_layers.7", 64, &bcv.bn3s, &bcv.bn3h));
    CKI(raw_t("conv_layers.9.weight", {32, 64, 3}, &bcv.c4w));     CKI(raw_t("conv_layers.9.bias", {32}, &bcv.c4b));
    CKI(raw_t("fc_layers.0.weight", {256, BCV_FLAT}, &bcv.f1w));   CKI(raw_t("fc_layers.0.bias", {256}, &bcv.f1b)); CKI(bn("fc_layers.1", 256, &bcv.bn4s, &bcv.bn4h));
    CKI(raw_t("fc_layers.3.weight", {128, 256}, &bcv.f2w));        CKI(raw_t("fc_layers.3.bias", {128}, &bcv.f2b)); CKI(bn("fc_layers.4", 128, &bcv.bn5s, &bcv.bn5h));
    CKI(raw_t("fc_layers.6.weight", {BCV_Z, 128}, &bcv.f3w));      CKI(raw_t("fc_layers.6.bias", {BCV_Z}, &bcv.f3b));
    CKI(raw_t("fc_mu.weight", {BCV_Z, BCV_Z}, &bcv.muw));          CKI(raw_t("fc_mu.bias", {BCV_Z}, &bcv.mub));
    bcv_ready = true;
    for (auto it = raw.begin(); it != raw.end();) it = it->first.compare(0, 6, "bcvae.") == 0 ? raw.erase(it) : std::next(it);
    return 0;
}

int said_engine::commit() {
    CK(cudaSetDevice(device));
    CK(cudaDeviceSynchronize());
    drop_graph();                     // its nodes hold weight pointers
    for (float* p : arena) cudaFree(p);
    arena.clear();
    tcmap.clear();
    hmap.clear();
    hmap32.clear();
    hmap64.clear();
    hmap96.clear();
    ready = false;
    ctx_B = ctx_T = 0;
    reg_h = true;
    reg_h32 = true;
    const int rc_d = commit_denoiser();
    reg_h32 = false;
    const int rc_e = rc_d == 0 ? commit_encoder() : 0;
    reg_h = false;
    CKI(rc_d);
    CKI(rc_e);
    const int enc_out = proj_dim > 0 ? proj_dim : enc_hidden;
    if (enc_out != ctx_dim)
        return fail("audio feature width " + std::to_string(enc_out) + " != denoiser context dim " + std::to_string(ctx_dim));
    ready = true;
    raw.clear();    // the host staging copies are consumed: a later load of a shallower model must not see this one's layers
    return 0;
}

// =====================================================================================================
// Launch helpers
// =====================================================================================================
namespace {

EpiStd mk_epi(float* out, long long ldo, int N) {
    EpiStd e;
    memset(&e, 0, sizeof(e));
    e.out = out;
    e.ldo = ldo;
    e.N = N;
    e.T = 1;
    e.zdiv = 1;
    e.acc_scale = 1.0f;
    return e;
}
ALoadPlain mk_plain(const float* A, long long lda, int M) {
    ALoadPlain a;
    a.A = A;
    a.lda = lda;
    a.zstride = 0;
    a.zdiv = 1;
    a.zstride2 = 0;
    a.M = M;
    a.A2 = nullptr;
    a.lda2 = 0;
    a.K0 = 0x7fffffff;
    return a;
}

}  // namespace

#define LAUNCH_CHECK() CKI(after_launch(st))

// =====================================================================================================
// Audio encoder
// =====================================================================================================
int said_engine::encode_audio(const float* wave, int B, int T_a, int T, float* emb_out, cudaStream_t st) {
    struct PrecisionScope {   // the encoder has its own precision knob
        int& p; int saved;
        PrecisionScope(int& p_, int v) : p(p_), saved(p_) { p = v; }
        ~PrecisionScope() { p = saved; }
    } scope(precision, enc_precision == 3 ? 0 : enc_precision);   // fp16x3 encoder below its row threshold: the IEEE fp32 kernels throughout
    if (!ready) return fail("weights not committed");
    if (B <= 0 || T <= 0) return fail("encode_audio: empty batch");
    if (enc_precision == 3 && (long long)B * T >= tc_min_rows) {
        scope.p = 3;
        return encode_audio_h(wave, B, T_a, T, emb_out, st);
    }
    int L[8];
    L[0] = (T_a - conv_k[0]) / conv_s[0] + 1;
    if (T_a < conv_k[0]) return fail("encode_audio: waveform shorter than the first conv kernel");
    for (int i = 1; i < n_conv; ++i) {
        L[i] = (L[i - 1] - conv_k[i]) / conv_s[i] + 1;
        if (L[i - 1] < conv_k[i] || L[i] < 1) return fail("encode_audio: waveform too short for the conv stack");
    }
    const int H = enc_hidden, CD = enc_conv_dim;
    const int Lf = L[n_conv - 1];
    // ---- conv0 + per-channel norm + GELU
    const int nchunk = 32;
    const int fpc = (L[0] + nchunk - 1) / nchunk;
    const size_t need_partial = (size_t)B * nchunk * 2 * CD;
    if (need_partial > c0_partial_cap) {
        if (c0_partial) cudaFree(c0_partial);
        c0_partial = nullptr;
        c0_partial_cap = 0;
        CK(cudaMalloc((void**)&c0_partial, need_partial * sizeof(double)));
        c0_partial_cap = need_partial;
    }
    // Per-clip frame strides S[i] with S[i-1] = 2 S[i] (all strides are 2): output frame j of clip b is then row
    // m = b*S[i] + j of ONE flat overlapping-row matrix over the previous layer (row m starts at frame 2m), so the
    // whole batch is a single GEMM (no per-clip launches, tcgen05-eligible).  Rows j >= L[i] are padding: they
    // read only padding / neighbouring frames and are never read by valid rows.
    int S[8];
    {
        int s_last = 1;
        for (int i = 0; i < n_conv; ++i) {
            const int sh = n_conv - 1 - i;
            s_last = std::max(s_last, (L[i] + (1 << sh) - 1) >> sh);
        }
        for (int i = 0; i < n_conv; ++i) S[i] = s_last << (n_conv - 1 - i);
    }
    for (int i = 1; i < n_conv; ++i)
        if (conv_s[i] != 2) return fail("audio encoder: conv strides other than (5,2,2,2,2,2,2) are not supported");
    CK(e_a.ensure((size_t)(B * S[0] + 4) * CD));
    CK(e_b.ensure((size_t)(B * S[1] + 4) * CD));
    if (enc_fe_layer_norm) {
        const long long warps = (long long)B * L[0];
        conv0_layernorm_gelu_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(wave, T_a, L[0], B, c0_w, c0_bias, c0_g, c0_b, 1e-5f,
                                                                                          e_a.p, S[0]);
        LAUNCH_CHECK();
    } else {
        conv0_stats_kernel<<<dim3(nchunk, B), C0_CH, 0, st>>>(wave, T_a, L[0], c0_w, fpc, c0_partial);
        LAUNCH_CHECK();
        conv0_apply_kernel<<<dim3((L[0] + C0_TILE - 1) / C0_TILE, B), C0_CH, 0, st>>>(wave, T_a, L[0], c0_w, c0_partial, nchunk,
                                                                                       c0_g, c0_b, 1e-5f, e_a.p, S[0]);
        LAUNCH_CHECK();
    }
    // ---- conv1..6 (stride 2, GELU): one flat overlapping-row GEMM per layer
    float* src = e_a.p;
    float* dst = e_b.p;
    for (int i = 1; i < n_conv; ++i) {
        const int rows = B * S[i];
        ALoadPlain al = mk_plain(src, (long long)conv_s[i] * CD, rows);
        EpiStd ep = mk_epi(dst, CD, CD);
        if (enc_fe_layer_norm) {   // conv + bias, then LayerNorm(512) + GELU over every (also padding) row, in place
            ep.bias = conv_bias_p[i];
            CKI(gemm(st, rows, CD, conv_k[i] * CD, al, conv_w[i], CD, ep));
            layernorm_rows_kernel<8><<<(unsigned)(((long long)rows * 32 + 255) / 256), 256, 0, st>>>(dst, nullptr, rows, CD, 1e-5f, conv_ln_g[i],
                                                                                                  conv_ln_b[i], dst, 1);
            LAUNCH_CHECK();
        } else {
            ep.act = 1;
            CKI(gemm(st, rows, CD, conv_k[i] * CD, al, conv_w[i], CD, ep));
        }
        std::swap(src, dst);
    }
    // src now holds (B, S_last, CD), valid frames [0, Lf) of each clip
    const int M = B * T;
    CK(e_c.ensure((size_t)M * std::max(CD, H)));
    CK(e_d.ensure((size_t)M * H));
    CK(e_qkv.ensure((size_t)M * 3 * H));
    CK(e_ff.ensure((size_t)M * enc_ffn));
    CK(e_xp.ensure((size_t)B * pos_g * (T + pos_k) * (H / pos_g)));
    // ---- interpolate to T frames + LayerNorm(512)  -> e_c (M, CD)
    interp_layernorm_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(src, B, Lf, S[n_conv - 1], T, CD, 1e-5f, fp_ln_g, fp_ln_b, e_c.p);
    LAUNCH_CHECK();
    // ---- projection 512 -> H  -> e_d
    {
        EpiStd ep = mk_epi(e_d.p, H, H);
        ep.bias = fp_b;
        CKI(gemm(st, M, H, CD, mk_plain(e_c.p, CD, M), fp_w, H, ep));
    }
    // ---- positional conv embedding: x + gelu(conv(x)) then LayerNorm   (TF :690-693)
    {
        const int cg = H / pos_g, Tp = T + pos_k;
        const long long tot = (long long)B * pos_g * Tp * cg;
        posconv_regroup_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(e_d.p, B, T, H, pos_g, pos_k, e_xp.p);
        LAUNCH_CHECK();
        ALoadPlain al = mk_plain(e_xp.p, cg, T);
        al.zdiv = pos_g;
        al.zstride = (long long)pos_g * Tp * cg;
        al.zstride2 = (long long)Tp * cg;
        EpiStd ep = mk_epi(e_c.p, H, cg);
        ep.bias = pos_b;
        ep.act = 1;
        ep.res = e_d.p;
        ep.ldr = H;
        ep.zdiv = pos_g;
        ep.zs0 = (long long)T * H;
        ep.zs1 = cg;
        ep.bias_zs = cg;
        CKI(gemm(st, T, cg, pos_k * cg, al, pos_w, cg, ep, B * pos_g, pos_g, (long long)pos_k * cg * cg));
        if (!enc_stable_ln) {   // post-LN encoder: LayerNorm right after the positional embedding (TF :690-693)
            layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(e_c.p, nullptr, M, H, 1e-5f, enc_ln_g, enc_ln_b, e_d.p);
            LAUNCH_CHECK();
        }
    }
    CK(cudaFuncSetAttribute(self_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)attention_smem_bytes<64>()));
    if (enc_stable_ln) {
        // ---- Wav2Vec2EncoderStableLayerNorm (TF modeling_wav2vec2.py:730-803, 612-655): pre-LN layers over h = e_c,
        //      h += attn(LN1(h)); h += ffn(LN2(h)); one LayerNorm after the last layer.  The residual adds run in place
        //      (epilogue reads and writes the same element).
        float* h = e_c.p;
        for (int l = 0; l < enc_layers; ++l) {
            const EncLayerW& W = enc[l];
            layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(h, nullptr, M, H, 1e-5f, W.ln1_g, W.ln1_b, e_d.p);
            LAUNCH_CHECK();
            {
                EpiStd ep = mk_epi(e_qkv.p, 3 * H, 3 * H);
                ep.bias = W.bqkv;
                CKI(gemm(st, M, 3 * H, H, mk_plain(e_d.p, H, M), W.wqkv, 3 * H, ep));
            }
            self_attention_kernel<64><<<dim3((T + ATT_QTILE - 1) / ATT_QTILE, enc_heads, B), ATT_THREADS,
                                        attention_smem_bytes<64>(), st>>>(e_qkv.p, 3 * H, 0, H, 2 * H, T, 0.125f, e_d.p, H);
            LAUNCH_CHECK();
            {
                EpiStd ep = mk_epi(h, H, H);
                ep.bias = W.bo;
                ep.res = h;
                ep.ldr = H;
                CKI(gemm(st, M, H, H, mk_plain(e_d.p, H, M), W.wo, H, ep));
            }
            layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(h, nullptr, M, H, 1e-5f, W.ln2_g, W.ln2_b, e_d.p);
            LAUNCH_CHECK();
            {
                EpiStd ep = mk_epi(e_ff.p, enc_ffn, enc_ffn);
                ep.bias = W.bff1;
                ep.act = 1;
                CKI(gemm(st, M, enc_ffn, H, mk_plain(e_d.p, H, M), W.wff1, enc_ffn, ep));
            }
            {
                EpiStd ep = mk_epi(h, H, H);
                ep.bias = W.bff2;
                ep.res = h;
                ep.ldr = H;
                CKI(gemm(st, M, H, enc_ffn, mk_plain(e_ff.p, enc_ffn, M), W.wff2, H, ep));
            }
        }
        float* dst_ln = proj_dim == 0 ? emb_out : e_d.p;
        layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(h, nullptr, M, H, 1e-5f, enc_ln_g, enc_ln_b, dst_ln);
        LAUNCH_CHECK();
        if (proj_dim != 0) {   // audio_proj_layer (diffusion.py:228-229)
            EpiStd ep = mk_epi(emb_out, proj_dim, proj_dim);
            ep.bias = b_aproj;
            CKI(gemm(st, M, proj_dim, H, mk_plain(e_d.p, H, M), w_aproj, proj_dim, ep));
        }
        return 0;
    }
    // ---- transformer layers; x lives in e_d
    const bool last_direct = proj_dim == 0;
    for (int l = 0; l < enc_layers; ++l) {
        const EncLayerW& W = enc[l];
        {
            EpiStd ep = mk_epi(e_qkv.p, 3 * H, 3 * H);
            ep.bias = W.bqkv;
            CKI(gemm(st, M, 3 * H, H, mk_plain(e_d.p, H, M), W.wqkv, 3 * H, ep));
        }
        self_attention_kernel<64><<<dim3((T + ATT_QTILE - 1) / ATT_QTILE, enc_heads, B), ATT_THREADS,
                                    attention_smem_bytes<64>(), st>>>(e_qkv.p, 3 * H, 0, H, 2 * H, T, 0.125f, e_c.p, H);
        LAUNCH_CHECK();
        {
            EpiStd ep = mk_epi(e_qkv.p, H, H);   // attention output projection + residual -> e_qkv (as M x H)
            ep.bias = W.bo;
            ep.res = e_d.p;
            ep.ldr = H;
            CKI(gemm(st, M, H, H, mk_plain(e_c.p, H, M), W.wo, H, ep));
        }
        layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(e_qkv.p, nullptr, M, H, 1e-5f, W.ln1_g, W.ln1_b, e_d.p);
        LAUNCH_CHECK();
        {
            EpiStd ep = mk_epi(e_ff.p, enc_ffn, enc_ffn);
            ep.bias = W.bff1;
            ep.act = 1;
            CKI(gemm(st, M, enc_ffn, H, mk_plain(e_d.p, H, M), W.wff1, enc_ffn, ep));
        }
        {
            EpiStd ep = mk_epi(e_c.p, H, H);
            ep.bias = W.bff2;
            ep.res = e_d.p;
            ep.ldr = H;
            CKI(gemm(st, M, H, enc_ffn, mk_plain(e_ff.p, enc_ffn, M), W.wff2, H, ep));
        }
        float* dst_ln = (l == enc_layers - 1 && last_direct) ? emb_out : e_d.p;
        layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(e_c.p, nullptr, M, H, 1e-5f, W.ln2_g, W.ln2_b, dst_ln);
        LAUNCH_CHECK();
    }
    if (!last_direct) {   // audio_proj_layer (diffusion.py:228-229)
        EpiStd ep = mk_epi(emb_out, proj_dim, proj_dim);
        ep.bias = b_aproj;
        CKI(gemm(st, M, proj_dim, H, mk_plain(e_d.p, H, M), w_aproj, proj_dim, ep));
    }
    return 0;
}

// =====================================================================================================
// Audio encoder, fp16x3 mode: every dense contraction (conv1..6, feature projection, q/k/v, attention output, feed-forward)
// is a TMA-fed tcgen05 kind::f16 GEMM over pair-format operands (gemm_h.cuh); the kernels in between write pairs.
// A strided Conv1d(k, stride 2) over the channel-last pair tensor is k segments of one GEMM: tap t is a tensor map whose
// row pitch is two frames and whose base is shifted by t frames.  conv0, the positional conv (grouped, K = 6144 per group)
// and the attention core stay on their fp32 kernels.
// =====================================================================================================
int said_engine::encode_audio_h(const float* wave, int B, int T_a, int T, float* emb_out, cudaStream_t st) {
    int L[8];
    L[0] = (T_a - conv_k[0]) / conv_s[0] + 1;
    if (T_a < conv_k[0]) return fail("encode_audio: waveform shorter than the first conv kernel");
    for (int i = 1; i < n_conv; ++i) {
        L[i] = (L[i - 1] - conv_k[i]) / conv_s[i] + 1;
        if (L[i - 1] < conv_k[i] || L[i] < 1) return fail("encode_audio: waveform too short for the conv stack");
        if (conv_s[i] != 2 || conv_k[i] > 3) return fail("audio encoder: conv layers other than (k <= 3, stride 2) are not supported");
    }
    const int H = enc_hidden, CD = enc_conv_dim;
    const int Lf = L[n_conv - 1];
    int S[8];
    {
        int s_last = 1;
        for (int i = 0; i < n_conv; ++i) {
            const int sh = n_conv - 1 - i;
            s_last = std::max(s_last, (L[i] + (1 << sh) - 1) >> sh);
        }
        for (int i = 0; i < n_conv; ++i) S[i] = s_last << (n_conv - 1 - i);
    }
    if (!status_flag) {
        CK(cudaMalloc((void**)&status_flag, sizeof(int)));
        CK(cudaMemset(status_flag, 0, sizeof(int)));
    }
    CK(e_a.ensure((size_t)(B * S[0] + 4) * CD));
    CK(e_b.ensure((size_t)(B * S[1] + 4) * CD));
    CK(pe_a.ensure_zero((size_t)(B * S[0] + 4) * CD));
    CK(pe_b.ensure_zero((size_t)(B * S[1] + 4) * CD));
    // ---- conv0 + norm + GELU (fp32 kernels), then the pair format
    if (enc_fe_layer_norm) {
        const long long warps = (long long)B * L[0];
        conv0_layernorm_gelu_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(wave, T_a, L[0], B, c0_w, c0_bias, c0_g, c0_b, 1e-5f,
                                                                                          e_a.p, S[0]);
        LAUNCH_CHECK();
    } else {
        const int nchunk = 32;
        const int fpc = (L[0] + nchunk - 1) / nchunk;
        const size_t need_partial = (size_t)B * nchunk * 2 * CD;
        if (need_partial > c0_partial_cap) {
            if (c0_partial) cudaFree(c0_partial);
            c0_partial = nullptr;
            c0_partial_cap = 0;
            CK(cudaMalloc((void**)&c0_partial, need_partial * sizeof(double)));
            c0_partial_cap = need_partial;
        }
        conv0_stats_kernel<<<dim3(nchunk, B), C0_CH, 0, st>>>(wave, T_a, L[0], c0_w, fpc, c0_partial);
        LAUNCH_CHECK();
        conv0_apply_kernel<<<dim3((L[0] + C0_TILE - 1) / C0_TILE, B), C0_CH, 0, st>>>(wave, T_a, L[0], c0_w, c0_partial, nchunk,
                                                                                       c0_g, c0_b, 1e-5f, e_a.p, S[0]);
        LAUNCH_CHECK();
    }
    auto to_pair = [&](const float* src, long long rows, int Cc, __half* dst) -> int {
        const long long nq = rows * (Cc / 4);
        f32_to_pair_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(src, rows, Cc, dst, status_flag);
        LAUNCH_CHECK();
        return 0;
    };
    __half* psrc = reinterpret_cast<__half*>(pe_a.p);
    __half* pdst = reinterpret_cast<__half*>(pe_b.p);
    CKI(to_pair(e_a.p, (long long)B * S[0], CD, psrc));
    // ---- conv1..6 (stride 2, GELU): row m of layer i reads frames 2m + tap of layer i - 1
    float* last_f32 = e_b.p;       // the last conv layer's output in fp32 (input of the interpolation)
    for (int i = 1; i < n_conv; ++i) {
        const int rows = B * S[i];
        const bool last = i == n_conv - 1;
        HSrc taps[3];
        for (int t = 0; t < conv_k[i]; ++t) taps[t] = HSrc{psrc + (size_t)t * 2 * CD, CD, rows, 4LL * CD};
        EpiStd ep = mk_epi(nullptr, CD, CD);
        ep.flag = status_flag;
        ep.pair_C = CD;
        if (enc_fe_layer_norm) {   // conv + bias -> fp32, then LayerNorm(512) + GELU -> pair (fp32 for the last layer)
            ep.out = e_b.p;
            ep.bias = conv_bias_p[i];
        } else {
            ep.act = 1;
            if (last) ep.out = last_f32;
            else ep.out_pair = pdst;
        }
        if (enc_split_k) {
            // one launch per tap (K = 512 each); the partial sums travel in fp32 through e_b (the epilogue adds them with round to
            // nearest before bias / GELU), so no tensor-core accumulation chain is longer than 512 / 16 x 3 MMAs
            for (int t = 0; t < conv_k[i]; ++t) {
                const bool fin = t == conv_k[i] - 1;
                EpiStd pe = fin ? ep : mk_epi(e_b.p, CD, CD);
                if (t > 0) {
                    pe.acc_in = e_b.p;
                    pe.ld_acc = CD;
                }
                CKI(gemm_h(st, rows, CD, {{taps[t], 0, CD, 0}}, conv_w[i], pe, TAG_OTHER, t == 0 ? -1 : 0, t * CD));
            }
        } else if (conv_k[i] == 3) {
            CKI(gemm_h(st, rows, CD, {{taps[0], 0, CD, 0}, {taps[1], 0, CD, 0}, {taps[2], 0, CD, 0}}, conv_w[i], ep, TAG_OTHER));
        } else {
            CKI(gemm_h(st, rows, CD, {{taps[0], 0, CD, 0}, {taps[1], 0, CD, 0}}, conv_w[i], ep, TAG_OTHER));
        }
        if (enc_fe_layer_norm) {
            layernorm_rows_kernel<8><<<(unsigned)(((long long)rows * 32 + 255) / 256), 256, 0, st>>>(
                e_b.p, nullptr, rows, CD, 1e-5f, conv_ln_g[i], conv_ln_b[i], last ? last_f32 : nullptr, 1, last ? nullptr : pdst, status_flag);
            LAUNCH_CHECK();
        }
        std::swap(psrc, pdst);
    }
    const int M = B * T;
    CK(e_c.ensure((size_t)M * std::max(CD, H)));
    CK(e_d.ensure((size_t)M * H));
    CK(e_qkv.ensure((size_t)M * 3 * H));
    CK(e_xp.ensure((size_t)B * pos_g * (T + pos_k) * (H / pos_g)));
    CK(pe_x.ensure_zero((size_t)M * std::max(CD, H)));
    CK(pe_att.ensure_zero((size_t)M * H));
    CK(pe_ff.ensure_zero((size_t)M * enc_ffn));
    __half* px = reinterpret_cast<__half*>(pe_x.p);
    __half* patt = reinterpret_cast<__half*>(pe_att.p);
    __half* pffn = reinterpret_cast<__half*>(pe_ff.p);
    // ---- interpolate to T frames + LayerNorm(512) -> pair
    interp_layernorm_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(last_f32, B, Lf, S[n_conv - 1], T, CD, 1e-5f, fp_ln_g, fp_ln_b,
                                                                      (float*)nullptr, px, status_flag);
    LAUNCH_CHECK();
    {   // projection 512 -> H  -> e_d (fp32)
        EpiStd ep = mk_epi(e_d.p, H, H);
        ep.bias = fp_b;
        CKI(gemm_h(st, M, H, {{HSrc{px, CD, M}, 0, CD, 0}}, fp_w, ep, TAG_OTHER));
    }
    // ---- positional conv embedding: x + gelu(conv(x)) then LayerNorm (fp32 batched grouped GEMM, as in encode_audio)
    {
        const int cg = H / pos_g, Tpd = T + pos_k;
        const long long tot = (long long)B * pos_g * Tpd * cg;
        posconv_regroup_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(e_d.p, B, T, H, pos_g, pos_k, e_xp.p);
        LAUNCH_CHECK();
        ALoadPlain al = mk_plain(e_xp.p, cg, T);
        al.zdiv = pos_g;
        al.zstride = (long long)pos_g * Tpd * cg;
        al.zstride2 = (long long)Tpd * cg;
        EpiStd ep = mk_epi(e_c.p, H, cg);
        ep.bias = pos_b;
        ep.act = 1;
        ep.res = e_d.p;
        ep.ldr = H;
        ep.zdiv = pos_g;
        ep.zs0 = (long long)T * H;
        ep.zs1 = cg;
        ep.bias_zs = cg;
        CKI(gemm(st, T, cg, pos_k * cg, al, pos_w, cg, ep, B * pos_g, pos_g, (long long)pos_k * cg * cg));
    }
    CK(cudaFuncSetAttribute(self_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attention_smem_bytes<64>()));
    auto ln = [&](const float* src, const float* g, const float* b, float* dst, __half* dst_pair) -> int {
        layernorm_rows_kernel<8><<<(M * 32 + 255) / 256, 256, 0, st>>>(src, nullptr, M, H, 1e-5f, g, b, dst, 0, dst_pair, status_flag);
        LAUNCH_CHECK();
        return 0;
    };
    auto attention = [&]() -> int {
        self_attention_kernel<64><<<dim3((T + ATT_QTILE - 1) / ATT_QTILE, enc_heads, B), ATT_THREADS, attention_smem_bytes<64>(), st>>>(
            e_qkv.p, 3 * H, 0, H, 2 * H, T, 0.125f, nullptr, H, T, patt, status_flag);
        LAUNCH_CHECK();
        return 0;
    };
    const HSrc sx{px, H, M}, satt{patt, H, M}, sff{pffn, enc_ffn, M};
    CK(e_ff.ensure((size_t)M * H));
    // feed-forward output projection (K = ffn = 3072 / 4096): split-K in 768-wide parts whose partial sums travel in fp32 through e_ff
    auto ffn2 = [&](const float* wkey, const EpiStd& fin) -> int {
        const int part = 768;
        if (!enc_split_k || enc_ffn % part != 0 || enc_ffn <= part) return gemm_h(st, M, H, {{sff, 0, enc_ffn, 0}}, wkey, fin, TAG_OTHER);
        const int np = enc_ffn / part;
        for (int j = 0; j < np; ++j) {
            EpiStd pe = j == np - 1 ? fin : mk_epi(e_ff.p, H, H);
            if (j > 0) {
                pe.acc_in = e_ff.p;
                pe.ld_acc = H;
            }
            CKI(gemm_h(st, M, H, {{sff, j * part, part, 0}}, wkey, pe, TAG_OTHER, j == 0 ? -1 : 0, j * part));
        }
        return 0;
    };
    if (enc_stable_ln) {
        // pre-LN layers (TF modeling_wav2vec2.py:612-655, 730-803) over h = e_c: h += attn(LN1(h)); h += ffn(LN2(h)); final LayerNorm
        float* h = e_c.p;
        for (int l = 0; l < enc_layers; ++l) {
            const EncLayerW& W = enc[l];
            CKI(ln(h, W.ln1_g, W.ln1_b, nullptr, px));
            {
                EpiStd ep = mk_epi(e_qkv.p, 3 * H, 3 * H);
                ep.bias = W.bqkv;
                CKI(gemm_h(st, M, 3 * H, {{sx, 0, H, 0}}, W.wqkv, ep, TAG_OTHER));
            }
            CKI(attention());
            {
                EpiStd ep = mk_epi(h, H, H);
                ep.bias = W.bo;
                ep.res = h;
                ep.ldr = H;
                CKI(gemm_h(st, M, H, {{satt, 0, H, 0}}, W.wo, ep, TAG_OTHER));
            }
            CKI(ln(h, W.ln2_g, W.ln2_b, nullptr, px));
            {
                EpiStd ep = mk_epi(nullptr, enc_ffn, enc_ffn);
                ep.bias = W.bff1;
                ep.act = 1;
                ep.out_pair = pffn;
                ep.pair_C = enc_ffn;
                ep.flag = status_flag;
                CKI(gemm_h(st, M, enc_ffn, {{sx, 0, H, 0}}, W.wff1, ep, TAG_OTHER));
            }
            {
                EpiStd ep = mk_epi(h, H, H);
                ep.bias = W.bff2;
                ep.res = h;
                ep.ldr = H;
                CKI(ffn2(W.wff2, ep));
            }
        }
        float* dst_ln = proj_dim == 0 ? emb_out : e_d.p;
        CKI(ln(h, enc_ln_g, enc_ln_b, dst_ln, nullptr));
    } else {
        // post-LN layers (TF :576-609): x lives in e_d (fp32) and px (pair)
        CKI(ln(e_c.p, enc_ln_g, enc_ln_b, e_d.p, px));
        for (int l = 0; l < enc_layers; ++l) {
            const EncLayerW& W = enc[l];
            {
                EpiStd ep = mk_epi(e_qkv.p, 3 * H, 3 * H);
                ep.bias = W.bqkv;
                CKI(gemm_h(st, M, 3 * H, {{sx, 0, H, 0}}, W.wqkv, ep, TAG_OTHER));
            }
            CKI(attention());
            {   // attention output projection + residual -> e_c
                EpiStd ep = mk_epi(e_c.p, H, H);
                ep.bias = W.bo;
                ep.res = e_d.p;
                ep.ldr = H;
                CKI(gemm_h(st, M, H, {{satt, 0, H, 0}}, W.wo, ep, TAG_OTHER));
            }
            CKI(ln(e_c.p, W.ln1_g, W.ln1_b, e_d.p, px));
            {
                EpiStd ep = mk_epi(nullptr, enc_ffn, enc_ffn);
                ep.bias = W.bff1;
                ep.act = 1;
                ep.out_pair = pffn;
                ep.pair_C = enc_ffn;
                ep.flag = status_flag;
                CKI(gemm_h(st, M, enc_ffn, {{sx, 0, H, 0}}, W.wff1, ep, TAG_OTHER));
            }
            {
                EpiStd ep = mk_epi(e_c.p, H, H);
                ep.bias = W.bff2;
                ep.res = e_d.p;
                ep.ldr = H;
                CKI(ffn2(W.wff2, ep));
            }
            const bool final_direct = l == enc_layers - 1 && proj_dim == 0;
            CKI(ln(e_c.p, W.ln2_g, W.ln2_b, final_direct ? emb_out : e_d.p, final_direct ? nullptr : px));
        }
    }
    if (proj_dim != 0) {   // audio_proj_layer (diffusion.py:228-229): fp32 GEMM of the generic path
        EpiStd ep = mk_epi(emb_out, proj_dim, proj_dim);
        ep.bias = b_aproj;
        CKI(gemm(st, M, proj_dim, H, mk_plain(e_d.p, H, M), w_aproj, proj_dim, ep));
    }
    return 0;
}

// =====================================================================================================
// Context hoist
// =====================================================================================================
int said_engine::prepare_context(const float* emb, int B, int T, int with_uncond, cudaStream_t st) {
    if (!ready) return fail("weights not committed");
    if (B <= 0 || T <= 0) return fail("prepare_context: empty batch");
    const int M = B * T, N = 8 * C;
    CK(kv.ensure((size_t)M * N));
    CK(vnull.ensure((size_t)4 * C));
    CK(cnull.ensure((size_t)4 * C));
    CK(vnull_tmp.ensure((size_t)N));
    {
        EpiStd ep = mk_epi(kv.p, N, N);
        CKI(gemm(st, M, N, ctx_dim, mk_plain(emb, ctx_dim, M), w_kv, N, ep));
    }
    if (with_uncond) {
        EpiStd ep = mk_epi(vnull_tmp.p, N, N);
        CKI(gemm(st, 1, N, ctx_dim, mk_plain(null_emb, ctx_dim, 1), w_kv, N, ep));
        for (int l = 0; l < 4; ++l)
            CK(cudaMemcpyAsync(vnull.p + l * C, vnull_tmp.p + l * 2 * C + C, C * sizeof(float), cudaMemcpyDeviceToDevice, st));
        // c_null[l] = to_out(v_null[l]) + bias: what the null-condition branch adds to its residual stream in block l
        for (int l = 0; l < 4; ++l) {
            EpiStd ep2 = mk_epi(cnull.p + l * C, C, C);
            ep2.bias = tr[l].bo2;
            CKI(gemm(st, 1, C, C, mk_plain(vnull.p + l * C, C, 1), tr[l].wo2, C, ep2));
        }
    }
    ctx_B = B;
    ctx_T = T;
    ctx_uncond = with_uncond ? 1 : 0;
    return 0;
}

// =====================================================================================================
// Denoiser forward
// =====================================================================================================
int said_engine::ensure_denoiser_ws(int Bp, int T) {
    // sized for the padded row space of the fp16x3 path (one zero row between clips); the other paths use the first Bp*T rows
    const size_t M = (size_t)Bp * (T + 1);
    for (auto& a : act) CK(a.ensure_zero(M * C));
    CK(gnbuf.ensure_zero(M * 2 * C));
    CK(qkv.ensure_zero(M * 3 * C));
    CK(ao.ensure_zero(M * C));
    CK(q2.ensure_zero(M * C));
    CK(ffb.ensure_zero(M * FF));
    CK(eps.ensure(M * in_ch));
    CK(ss.ensure((size_t)Bp * 2 * C * 2));
    CK(ss_st.ensure((size_t)Bp * C * 2));
    if (precision == 3) {
        CK(p_raw.ensure_zero(M * 2 * C));
        CK(p_ln.ensure_zero(M * C));
        CK(p_x2.ensure_zero(M * C));
        CKI(ensure_ffn_split(M));
    }
    if (!status_flag) {
        CK(cudaMalloc((void**)&status_flag, sizeof(int)));
        CK(cudaMemset(status_flag, 0, sizeof(int)));
    }
    if (!step_ctr) CK(cudaMalloc((void**)&step_ctr, sizeof(int)));
    if ((size_t)Bp * GN_SPLIT_MAX * 2 * C > gn_partial_cap) {
        if (gn_partial) cudaFree(gn_partial);
        gn_partial = nullptr;
        gn_partial_cap = 0;
        CK(cudaMalloc((void**)&gn_partial, (size_t)Bp * GN_SPLIT_MAX * 2 * C * sizeof(double)));
        gn_partial_cap = (size_t)Bp * GN_SPLIT_MAX * 2 * C;
    }
    CK(cudaFuncSetAttribute(self_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)attention_smem_bytes<32>()));
    if (T <= tc::ATC_MAXKEYS)
        CK(cudaFuncSetAttribute(tc::self_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)tc::attention_tc_smem_bytes(T)));
    if (T <= hx::AH_MAXT)
        CK(cudaFuncSetAttribute(hx::self_attention_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)hx::attention_h_smem_bytes(T, hx::attention_h_groups(T, HEADS))));
    return 0;
}

// x: (src_batch, T, in_ch) latents; sample b of the Bp denoiser samples reads clip b % src_batch.
// Samples [0, n_uncond) are the null-condition branch.  emb_table: (rows, 5, 192), row = *step_ptr
// or the sample index when step_ptr is null.
int said_engine::forward(cudaStream_t st, const float* x, int src_batch, int Bp, int n_uncond, int T,
                         const float* emb_table, const int* step_ptr, float* eps_out, float* taps) {
    const int M = Bp * T;
    const int Mc = (Bp - n_uncond) * T;
    // Under classifier-free guidance both branches are fed the same latents (diffusion.py:421-423) and differ only
    // from the first cross-attention on, so the input conv, the first ResBlock and the first block's
    // GroupNorm / self-attention / out-projection are computed ONCE for the Bs = B conditional samples and read by
    // both branches (sample b of a shared tensor = row block b % Bs).
    const bool share = n_uncond > 0 && 2 * n_uncond == Bp && taps == nullptr;
    const int Bs = share ? Bp - n_uncond : Bp;
    const int Ms = Bs * T;
    float* h0 = act[0].p; float* h1 = act[1].p; float* A = act[2].p; float* Bb = act[3].p;
    float* t1 = act[4].p; float* x1 = act[5].p; float* x2 = act[6].p;
    float* sc = ss.p; float* sh = ss.p + (size_t)Bp * 2 * C;
    float* sc_st = ss_st.p; float* sh_st = ss_st.p + (size_t)Bp * C;
    const float att_scale = 1.0f / sqrtf((float)HD);
    int tap_idx = 0;
    auto tap = [&](const float* p) -> int {
        if (taps) {
            CK(cudaMemcpyAsync(taps + (size_t)tap_idx * M * C, p, (size_t)M * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
            ++tap_idx;
        }
        return 0;
    };
    // GroupNorm of `nb` samples whose data is sample (b % src_nb) of src
    auto gn = [&](const float* src, int src_nb, int nb, int cpg, float eps_, const float* g, const float* b, float* osc, float* osh,
                  int ld, int off, float* act_out = nullptr, int act_ld = 0, int act_off = 0) -> int {
        {   // one launch, one cluster per sample (4 CTAs at batch scale, 8 for a handful of samples)
            const int cl = nb * GN_SPLIT >= num_sms ? GN_SPLIT : 8;
            const int rows = (T + cl - 1) / cl;
            if (gn_fused && (rows + 7) / 8 <= GNF_MAXR) {
                cur_tag = TAG_GN;
                CK(launch_ex(gn_fused_kernel, dim3(cl, nb), dim3(GNF_THREADS), 0, st, pdl, cl, src, src_nb, T, cpg, eps_, g, b, osc, osh, ld, off,
                             act_out, act_ld, act_off));
                LAUNCH_CHECK();
                return 0;
            }
        }
        const int nsp = nb * GN_SPLIT >= num_sms ? GN_SPLIT : GN_SPLIT_MAX;   // few samples: more CTAs each
        cur_tag = TAG_GN;
        CK(launch_ex(gn_partial_kernel, dim3(nsp, nb), dim3(GN_THREADS), 0, st, pdl, 1, src, src_nb, T, gn_partial));
        LAUNCH_CHECK();
        cur_tag = TAG_GN;
        CK(launch_ex(gn_finish_kernel, dim3(act_out ? nsp : 1, nb), dim3(GN_THREADS), 0, st, pdl, 1, src, src_nb, T, cpg, eps_,
                     (const double*)gn_partial, nsp, g, b, osc, osh, ld, off, act_out, act_ld, act_off));
        LAUNCH_CHECK();
        return 0;
    };
    // Tensor-core path: silu(gn(x)) is materialised once by the statistics kernel (gnbuf) instead of being
    // recomputed by the conv loader for each of the 3 taps -- the loader's instruction cost, not HBM, is what
    // bounds the tcgen05 conv GEMMs.
    const bool mat = precision != 0 && M >= tc_min_rows;
    float* gnb = gnbuf.p;
    // ResBlock (openaimodel.py:207-227) over nb samples: in (a [, skip]) -> out; a holds a_nb samples, skip skip_nb
    auto resblock = [&](int i, const float* a, int a_nb, const float* skip, int skip_nb, int nb, float* out) -> int {
        const ResBlockW& W = rb[i];
        const int cin = W.cin;
        const int m = nb * T;
        if (skip) {
            CKI(gn(a, a_nb, nb, 12, 1e-5f, W.gn1_g, W.gn1_b, sc, sh, cin, 0, mat ? gnb : nullptr, cin, 0));
            CKI(gn(skip, skip_nb, nb, 12, 1e-5f, W.gn1_g + C, W.gn1_b + C, sc, sh, cin, C, mat ? gnb : nullptr, cin, C));
        } else {
            CKI(gn(a, a_nb, nb, 6, 1e-5f, W.gn1_g, W.gn1_b, sc, sh, cin, 0, mat ? gnb : nullptr, cin, 0));
        }
        {
            ALoadConv3 al{a, skip, C, skip ? C : 0, cin, T, m, a_nb, sc, sh, 3 * cin, skip_nb};
            if (mat) al = ALoadConv3{gnb, nullptr, cin, 0, cin, T, m, nb, nullptr, nullptr, 3 * cin, 0};
            EpiStd ep = mk_epi(t1, C, C);
            ep.bias = W.b1;
            ep.emb = emb_table + (size_t)i * C;
            ep.emb_ld = 5 * C;
            ep.step_ptr = step_ptr;
            ep.T = T;
            CKI(gemm(st, m, C, 3 * cin, al, W.w1, C, ep));
        }
        CKI(gn(t1, nb, nb, 6, 1e-5f, W.gn2_g, W.gn2_b, sc, sh, C, 0, mat ? gnb : nullptr, C, 0));
        ALoadConv3 al2c{t1, nullptr, C, 0, C, T, m, nb, sc, sh, 3 * C, 0};
        if (mat) al2c = ALoadConv3{gnb, nullptr, C, 0, C, T, m, nb, nullptr, nullptr, 3 * C, 0};
        if (skip) {
            // second conv over t1, then the 1x1 skip_connection over the raw concat as a second GEMM that
            // accumulates through the residual input (one loader cannot address three tensors)
            EpiStd ep = mk_epi(x1, C, C);
            ep.bias = W.b2;
            CKI(gemm(st, m, C, 3 * C, al2c, W.w2, C, ep));
            ALoadConv3 al2{a, skip, C, C, cin, T, m, a_nb, nullptr, nullptr, 0, skip_nb};   // K3 = 0: raw centre tap only
            EpiStd ep2 = mk_epi(out, C, C);
            ep2.res = x1;
            ep2.ldr = C;
            CKI(gemm(st, m, C, cin, al2, W.w2 + (size_t)3 * C * C, C, ep2));
        } else {
            EpiStd ep = mk_epi(out, C, C);
            ep.bias = W.b2;
            ep.res = a;
            ep.ldr = C;
            ep.res_mod = a_nb * T;
            CKI(gemm(st, m, C, 3 * C, al2c, W.w2, C, ep));
        }
        return 0;
    };
    // SpatialTransformer + BasicTransformerBlock (attention.py:223-234, 167-193): h (h_nb samples) -> out (Bp samples).
    // When h_nb < Bp (shared CFG prefix) everything up to the first cross-attention runs on the h_nb samples.
    auto transformer = [&](int i, const float* h, int h_nb, float* out) -> int {
        const TransformerW& W = tr[i];
        const int mh = h_nb * T;
        const bool shared_front = h_nb < Bp;
        CKI(gn(h, h_nb, h_nb, 6, 1e-6f, W.gn_g, W.gn_b, sc_st, sh_st, C, 0));
        // Tensor-core path: LN outputs that feed a GEMM wider than one output tile are materialised once (gnbuf is
        // free inside the transformer block) so the GEMM's producers run the plain copy path
        auto ln_rows = [&](const float* src, int m, const float* ps, const float* pb, const float* g, const float* b, float* dst) -> int {
            cur_tag = TAG_GN;
            CK(launch_ex(ln192_rows_kernel, dim3((unsigned)(((long long)m * 16 + 255) / 256)), dim3(256), 0, st, pdl, 1, src, m, T, ps, pb, g, b, 1e-5f, dst));
            LAUNCH_CHECK();
            return 0;
        };
        if (mat) {   // q,k,v = LN1(GN(h)) W   (no bias)
            CKI(ln_rows(h, mh, sc_st, sh_st, W.ln1_g, W.ln1_b, gnb));
            EpiStd ep = mk_epi(qkv.p, 3 * C, 3 * C);
            CKI(gemm(st, mh, 3 * C, C, mk_plain(gnb, C, mh), W.wqkv, 3 * C, ep));
        } else {
            ALoadLN al{h, mh, T, sc_st, sh_st, W.ln1_g, W.ln1_b, 1e-5f};
            EpiStd ep = mk_epi(qkv.p, 3 * C, 3 * C);
            CKI(gemm(st, mh, 3 * C, C, al, W.wqkv, 3 * C, ep));
        }
        cur_tag = TAG_ATTN;
        if (mat && T <= tc::ATC_MAXKEYS) {
            CK(launch_ex(tc::self_attention_tc_kernel, dim3(HEADS, h_nb), dim3(tc::ATC_THREADS), tc::attention_tc_smem_bytes(T), st, pdl, 1,
                         (const float*)qkv.p, 3 * C, 0, C, 2 * C, T, att_scale, ao.p, C, T, (__half*)nullptr, (int*)nullptr));
        } else {
            CK(launch_ex(self_attention_kernel<32>, dim3((T + ATT_QTILE - 1) / ATT_QTILE, HEADS, h_nb), dim3(ATT_THREADS),
                         attention_smem_bytes<32>(), st, pdl, 1, (const float*)qkv.p, 3 * C, 0, C, 2 * C, T, att_scale, ao.p, C, T,
                         (__half*)nullptr, (int*)nullptr));
        }
        LAUNCH_CHECK();
        {   // x1 = to_out(attn) + GN(h)
            EpiStd ep = mk_epi(x1, C, C);
            ep.bias = W.bo1;
            ep.res = h;
            ep.ldr = C;
            ep.res_scale = sc_st;
            ep.res_shift = sh_st;
            ep.res_aff_ld = C;
            ep.T = T;
            CKI(gemm(st, mh, C, C, mk_plain(ao.p, C, mh), W.wo1, C, ep));
        }
        // rows of x1 that belong to the conditional samples: all of them when the front is shared
        const size_t x1c = shared_front ? 0 : (size_t)n_uncond * T * C;
        if (Mc > 0) {   // cross-attention queries, conditional samples only
            ALoadLN al{x1 + x1c, Mc, T, nullptr, nullptr, W.ln2_g, W.ln2_b, 1e-5f};
            EpiStd ep = mk_epi(q2.p, C, C);
            CKI(gemm(st, Mc, C, C, al, W.wq2, C, ep));
        }
        {
            const long long tot = (long long)M * HEADS * 8;   // 8 lanes per (row, head)
            if (tot >= (1LL << 31)) return fail("cross-attention: batch x frames exceeds the kernel's 32-bit index range");
            cur_tag = TAG_XATTN;
            CK(launch_ex(cross_attention_band_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st, pdl, 1, (const float*)q2.p,
                         (const float*)kv.p, 8 * C, i * 2 * C, (const int2*)band_dev, (const float*)(cnull.p + i * C), (const float*)x1, x2,
                         mh, n_uncond, Bp, T, T, ctx_T, att_scale, ao.p, (__half*)nullptr, (int*)nullptr));
            LAUNCH_CHECK();
        }
        if (Mc > 0) {   // x2 = to_out(attn2) + x1 on the conditional rows (the kernel above wrote the unconditional ones)
            const size_t r0 = (size_t)n_uncond * T * C;
            EpiStd ep = mk_epi(x2 + r0, C, C);
            ep.bias = W.bo2;
            ep.res = x1 + x1c;
            ep.ldr = C;
            CKI(gemm(st, Mc, C, C, mk_plain(ao.p + r0, C, Mc), W.wo2, C, ep));
        }
        if (mat) {   // GEGLU over the materialised LN3(x2)
            CKI(ln_rows(x2, M, nullptr, nullptr, W.ln3_g, W.ln3_b, gnb));
            EpiGeglu ep{ffb.p, FF, 2 * FF, W.bff1};
            CKI(gemm(st, M, 2 * FF, C, mk_plain(gnb, C, M), W.wff1, 2 * FF, ep));
        } else {
            ALoadLN al{x2, M, T, nullptr, nullptr, W.ln3_g, W.ln3_b, 1e-5f};
            EpiGeglu ep{ffb.p, FF, 2 * FF, W.bff1};
            CKI(gemm(st, M, 2 * FF, C, al, W.wff1, 2 * FF, ep));
        }
        if (mat) {   // out = proj_out(ff2(ff) + x2) + h as ONE GEMM over [ff | x2] with the folded weights
            ALoadPlain al = mk_plain(ffb.p, FF, M);
            al.A2 = x2;
            al.lda2 = C;
            al.K0 = FF;
            EpiStd ep = mk_epi(out, C, C);
            ep.bias = W.bffp;
            ep.res = h;
            ep.ldr = C;
            ep.res_mod = mh;
            CKI(gemm(st, M, C, FF + C, al, W.wffp, C, ep));
            return 0;
        }
        float* x3 = shared_front ? t1 : x1;   // x1 may hold only the shared rows
        {   // x3 = ff2 + x2
            EpiStd ep = mk_epi(x3, C, C);
            ep.bias = W.bff2;
            ep.res = x2;
            ep.ldr = C;
            CKI(gemm(st, M, C, FF, mk_plain(ffb.p, FF, M), W.wff2, C, ep));
        }
        {   // out = proj_out(x3) + h
            EpiStd ep = mk_epi(out, C, C);
            ep.bias = W.bproj;
            ep.res = h;
            ep.ldr = C;
            ep.res_mod = mh;
            CKI(gemm(st, M, C, C, mk_plain(x3, C, M), W.wproj, C, ep));
        }
        return 0;
    };

    {   // input conv (openaimodel.py:473-479)
        ALoadConv3 al{x, nullptr, in_ch, 0, in_ch, T, Ms, src_batch, nullptr, nullptr, 3 * in_ch, 0};
        EpiStd ep = mk_epi(h0, C, C);
        ep.bias = b_in;
        CKI(gemm(st, Ms, C, 3 * in_ch, al, w_in, C, ep));
    }
    CKI(tap(h0));
    CKI(resblock(0, h0, Bs, nullptr, 0, Bs, A));       CKI(tap(A));
    CKI(transformer(0, A, Bs, h1));                    CKI(tap(h1));
    CKI(resblock(1, h1, Bp, nullptr, 0, Bp, Bb));      CKI(tap(Bb));
    CKI(transformer(1, Bb, Bp, A));                    CKI(tap(A));
    CKI(resblock(2, A, Bp, nullptr, 0, Bp, Bb));       CKI(tap(Bb));
    CKI(resblock(3, Bb, Bp, h1, Bp, Bp, A));           CKI(tap(A));
    CKI(transformer(2, A, Bp, Bb));                    CKI(tap(Bb));
    CKI(resblock(4, Bb, Bp, h0, Bs, Bp, A));           CKI(tap(A));
    CKI(transformer(3, A, Bp, Bb));                    CKI(tap(Bb));
    {   // out: GN + SiLU + conv3 -> in_ch   (openaimodel.py:665-669)
        CKI(gn(Bb, Bp, Bp, 6, 1e-5f, out_gn_g, out_gn_b, sc, sh, C, 0, mat ? gnb : nullptr, C, 0));
        ALoadConv3 al{Bb, nullptr, C, 0, C, T, M, Bp, sc, sh, 3 * C, 0};
        if (mat) al = ALoadConv3{gnb, nullptr, C, 0, C, T, M, Bp, nullptr, nullptr, 3 * C, 0};
        EpiStd ep = mk_epi(eps_out, in_ch, in_ch);
        ep.bias = b_out;
        CKI(gemm(st, M, in_ch, 3 * C, al, w_out, in_ch, ep));
    }
    return 0;
}

// =====================================================================================================
// Denoiser forward, fp16x3 path: every contraction is a TMA-fed tcgen05 kind::f16 GEMM over pair-format operands
// (gemm_h.cuh); GroupNorm / LayerNorm / attention kernels write their outputs directly in that format.
// Row space: sample b, frame t -> row b * (T + 1) + t; row T of every sample is a zero row in the conv operands, so
// Conv1d(k=3, pad=1) is three row-shifted TMA boxes and tiles may straddle clips.  Same arguments as forward().
// =====================================================================================================
int said_engine::forward_h(cudaStream_t st, const float* x, int src_batch, int Bp, int n_uncond, int T,
                           const float* emb_table, const int* step_ptr, float* eps_out, float* taps) {
    const int Tp = T + 1;
    const int Mp = Bp * Tp;
    const int Mcp = (Bp - n_uncond) * Tp;
    const bool share = n_uncond > 0 && 2 * n_uncond == Bp && taps == nullptr;   // see forward(): shared guidance prefix
    const int Bs = share ? Bp - n_uncond : Bp;
    const int Msp = Bs * Tp;
    struct PdlScope {
        bool& p; bool saved;
        PdlScope(bool& p_, bool v) : p(p_), saved(p_) { p = v; }
        ~PdlScope() { p = saved; }
    } pdl_scope(pdl, pdl || (pdl_small && (small_rows(Mp) || mid_rows(Mp) || half_rows(Mp))));
    // (a half-batch works in its own row range of every workspace: hctx, set by denoise())
    const size_t rowb = (size_t)hctx.row_base;
    float* h0 = act[0].p + rowb * C; float* h1 = act[1].p + rowb * C; float* A = act[2].p + rowb * C; float* Bb = act[3].p + rowb * C;
    float* t1 = act[4].p + rowb * C; float* x1 = act[5].p + rowb * C; float* x2 = act[6].p + rowb * C;
    float* sc_st = ss_st.p + (size_t)hctx.samp_base * 2 * C; float* sh_st = sc_st + (size_t)Bp * C;
    float* const qkv_p = qkv.p + rowb * 3 * C;
    float* const q2_p = q2.p + rowb * C;
    const float* const kv_p = kv.p + (size_t)hctx.clip_base * ctx_T * 8 * C;
    __half* pgn = reinterpret_cast<__half*>(gnbuf.p) + rowb * 4 * C;     // conv operand: silu(gn(x)), 192 or 384 columns
    __half* praw = reinterpret_cast<__half*>(p_raw.p) + rowb * 4 * C;    // raw concat [h | skip] (1x1 skip_connection operand), 384 columns
    __half* pln = reinterpret_cast<__half*>(p_ln.p) + rowb * 2 * C;      // LayerNorm outputs
    __half* pao = reinterpret_cast<__half*>(ao.p) + rowb * 2 * C;        // attention outputs
    __half* px2 = reinterpret_cast<__half*>(p_x2.p) + rowb * 2 * C;      // residual stream before the feed-forward, as an operand
    __half* pff = reinterpret_cast<__half*>(ffb.p) + rowb * 2 * FF;      // GEGLU output, 768 columns
    const float att_scale = 1.0f / sqrtf((float)HD);
    int tap_idx = 0;
    auto tap = [&](const float* p) -> int {
        if (taps) {
            CK(cudaMemcpy2DAsync(taps + (size_t)tap_idx * Bp * T * C, (size_t)T * C * sizeof(float), p, (size_t)Tp * C * sizeof(float),
                                 (size_t)T * C * sizeof(float), (size_t)Bp, cudaMemcpyDeviceToDevice, st));
            ++tap_idx;
        }
        return 0;
    };
    // clusters of 4 CTAs per sample, 10 frames per thread at batch scale; 8 CTAs for long clips / few samples.  (8 CTAs x 5 frames
    // per thread at batch scale -- more resident warps -- measured slower: 0.58 vs 0.49 ms per step; SAID_GN_CL8 selects it.)
    static const bool gn_cl8 = getenv("SAID_GN_CL8") != nullptr;
    static const int gn_two_env = getenv("SAID_GN_TWO") ? atoi(getenv("SAID_GN_TWO")) : 0;   // A/B: stats + apply as two plain launches, this many slabs per sample (<= 16)
    const int gn_two = gn_two_env > GN_SPLIT_MAX ? GN_SPLIT_MAX : gn_two_env;
    const int gn_cl = (!gn_cl8 && T <= 320 && Bp * GN_SPLIT >= num_sms) ? GN_SPLIT : 8;
    const bool gn_r5 = gn_cl == 8 && ((T + 7) / 8 + 7) / 8 <= 5;
    if (((T + gn_cl - 1) / gn_cl + 7) / 8 > GNF_MAXR) return fail("fp16x3 path: at most 640 frames per clip");
    // GroupNorm of nb samples (data = sample b % src_nb of src): scale/shift and / or pair-format outputs
    auto gnp = [&](const float* src, int src_nb, int nb, int cpg, float eps_, const float* g, const float* b, float* osc, float* osh,
                   __half* act_pair, __half* raw_pair, int act_C, int act_off) -> int {
        cur_tag = TAG_GN;
        if (gn_two) {
            const int slabs = gn_two;
            CK(launch_ex(gn_stats_kernel, dim3(slabs, nb), dim3(GN2_THREADS), 0, st, pdl, 1, src, src_nb, T, Tp, gn_partial));
            LAUNCH_CHECK();
            cur_tag = TAG_GN;
            CK(launch_ex(gn_apply_kernel, dim3(slabs, nb), dim3(GN2_THREADS), 0, st, pdl, 1, src, src_nb, T, Tp, cpg, eps_,
                         (const double*)gn_partial, g, b, osc, osh, C, 0, act_pair, raw_pair, act_C, act_off, status_flag));
            LAUNCH_CHECK();
            return 0;
        }
        if (gn_r5)
            CK(launch_ex(gn_pair_kernel<5>, dim3(gn_cl, nb), dim3(GNF_THREADS), 0, st, pdl, gn_cl, src, src_nb, T, Tp, cpg, eps_, g, b, osc, osh,
                         C, 0, act_pair, raw_pair, act_C, act_off, status_flag));
        else
            CK(launch_ex(gn_pair_kernel<GNF_MAXR>, dim3(gn_cl, nb), dim3(GNF_THREADS), 0, st, pdl, gn_cl, src, src_nb, T, Tp, cpg, eps_, g, b, osc,
                         osh, C, 0, act_pair, raw_pair, act_C, act_off, status_flag));
        LAUNCH_CHECK();
        return 0;
    };
    auto ln_pair = [&](const float* src, int m, const float* ps, const float* pb, const float* g, const float* b, __half* dst, __half* raw) -> int {
        cur_tag = TAG_GN;
        CK(launch_ex(ln192_pair_kernel, dim3((unsigned)(((long long)m * 16 + 255) / 256)), dim3(256), 0, st, pdl, 1, src, m, Tp, ps, pb, g, b,
                     1e-5f, dst, raw, status_flag));
        LAUNCH_CHECK();
        return 0;
    };
    auto psrc = [](const __half* base, int Cc, long long rows) { return HSrc{base, Cc, rows}; };
    // ResBlock (openaimodel.py:207-227)
    auto resblock = [&](int i, const float* a, int a_nb, const float* skip, int skip_nb, int nb, float* out) -> int {
        const ResBlockW& W = rb[i];
        const int cin = W.cin;
        const int m = nb * Tp;
        if (skip) {
            CKI(gnp(a, a_nb, nb, 12, 1e-5f, W.gn1_g, W.gn1_b, nullptr, nullptr, pgn, praw, cin, 0));
            CKI(gnp(skip, skip_nb, nb, 12, 1e-5f, W.gn1_g + C, W.gn1_b + C, nullptr, nullptr, pgn, praw, cin, C));
        } else {
            CKI(gnp(a, a_nb, nb, 6, 1e-5f, W.gn1_g, W.gn1_b, nullptr, nullptr, pgn, nullptr, cin, 0));
        }
        {
            const HSrc g1 = psrc(pgn, cin, m);
            EpiStd ep = mk_epi(t1, C, C);
            ep.bias = W.b1;
            ep.emb = emb_table + (size_t)i * C;
            ep.emb_ld = 5 * C;
            ep.step_ptr = step_ptr;
            ep.T = Tp;
            CKI(gemm_h(st, m, C, {{g1, 0, cin, -1}, {g1, 0, cin, 0}, {g1, 0, cin, 1}}, W.w1, ep, TAG_GEMM_CONV));
        }
        CKI(gnp(t1, nb, nb, 6, 1e-5f, W.gn2_g, W.gn2_b, nullptr, nullptr, pgn, nullptr, C, 0));
        const HSrc g2 = psrc(pgn, C, m);
        EpiStd ep = mk_epi(out, C, C);
        ep.bias = W.b2;
        if (skip) {   // second conv and the 1x1 skip_connection over the raw concat as ONE contraction (K = 576 + 384)
            CKI(gemm_h(st, m, C, {{g2, 0, C, -1}, {g2, 0, C, 0}, {g2, 0, C, 1}, {psrc(praw, cin, m), 0, cin, 0}}, W.b2 /*key of the fused image*/,
                       ep, TAG_GEMM_CONV));
        } else {
            ep.res = a;
            ep.ldr = C;
            ep.res_mod = a_nb * Tp;
            CKI(gemm_h(st, m, C, {{g2, 0, C, -1}, {g2, 0, C, 0}, {g2, 0, C, 1}}, W.w2, ep, TAG_GEMM_CONV));
        }
        return 0;
    };
    // SpatialTransformer + BasicTransformerBlock (attention.py:223-234, 167-193)
    auto transformer = [&](int i, const float* h, int h_nb, float* out) -> int {
        const TransformerW& W = tr[i];
        const int mh = h_nb * Tp;
        const bool shared_front = h_nb < Bp;
        CKI(gnp(h, h_nb, h_nb, 6, 1e-6f, W.gn_g, W.gn_b, sc_st, sh_st, nullptr, nullptr, C, 0));
        CKI(ln_pair(h, mh, sc_st, sh_st, W.ln1_g, W.ln1_b, pln, nullptr));
        {   // q,k,v = LN1(GN(h)) W   (no bias)
            EpiStd ep = mk_epi(qkv_p, 3 * C, 3 * C);
            CKI(gemm_h(st, mh, 3 * C, {{psrc(pln, C, mh), 0, C, 0}}, W.wqkv, ep, TAG_GEMM_PLAIN));
        }
        cur_tag = TAG_ATTN;
        if (T <= hx::AH_MAXT && attn_h) {
            const int ag = (h_nb * HEADS / 2 >= num_sms) ? hx::attention_h_groups(T, HEADS) : 1;   // few samples: one head per CTA, twice the CTAs
            const int nqt = (T + 127) / 128;
            const int zs = (ag == 1 && h_nb * HEADS * nqt <= num_sms) ? nqt : 1;                    // ... and one query tile per CTA while that fits a wave
            CK(launch_ex(hx::self_attention_h_kernel, dim3(HEADS / ag, h_nb, zs), dim3(hx::AH_THREADS * ag), hx::attention_h_smem_bytes(T, ag), st, pdl, 1,
                         (const float*)qkv_p, 3 * C, 0, C, 2 * C, T, att_scale, (float*)nullptr, C, Tp, pao, status_flag,
                         (uint32_t)hx::attention_h_group_bytes(T), (long long*)nullptr));
        } else if (T <= tc::ATC_MAXKEYS) {
            CK(launch_ex(tc::self_attention_tc_kernel, dim3(HEADS, h_nb), dim3(tc::ATC_THREADS), tc::attention_tc_smem_bytes(T), st, pdl, 1,
                         (const float*)qkv_p, 3 * C, 0, C, 2 * C, T, att_scale, (float*)nullptr, C, Tp, pao, status_flag));
        } else {
            CK(launch_ex(self_attention_kernel<32>, dim3((T + ATT_QTILE - 1) / ATT_QTILE, HEADS, h_nb), dim3(ATT_THREADS),
                         attention_smem_bytes<32>(), st, pdl, 1, (const float*)qkv_p, 3 * C, 0, C, 2 * C, T, att_scale, (float*)nullptr, C, Tp,
                         pao, status_flag));
        }
        LAUNCH_CHECK();
        {   // x1 = to_out(attn) + GN(h)
            EpiStd ep = mk_epi(x1, C, C);
            ep.bias = W.bo1;
            ep.res = h;
            ep.ldr = C;
            ep.res_scale = sc_st;
            ep.res_shift = sh_st;
            ep.res_aff_ld = C;
            ep.T = Tp;
            CKI(gemm_h(st, mh, C, {{psrc(pao, C, mh), 0, C, 0}}, W.wo1, ep, TAG_GEMM_PLAIN));
        }
        const size_t x1c = shared_front ? 0 : (size_t)n_uncond * Tp * C;   // first row of the conditional samples in x1
        const size_t r0 = (size_t)n_uncond * Tp;                           // ... in the full-batch tensors
        if (Mcp > 0) {   // cross-attention queries, conditional samples only
            CKI(ln_pair(x1 + x1c, Mcp, nullptr, nullptr, W.ln2_g, W.ln2_b, pln, nullptr));
            EpiStd ep = mk_epi(q2_p, C, C);
            CKI(gemm_h(st, Mcp, C, {{psrc(pln, C, Mcp), 0, C, 0}}, W.wq2, ep, TAG_GEMM_LN));
        }
        {
            const long long tot = (long long)Bp * T * HEADS * 8;   // 8 lanes per (frame, head)
            if (tot >= (1LL << 31) || (long long)Bp * (T + 1) >= (1LL << 31)) return fail("cross-attention: batch x frames exceeds the kernel's 32-bit index range");
            cur_tag = TAG_XATTN;
            CK(launch_ex(cross_attention_band_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st, pdl, 1, (const float*)q2_p,
                         kv_p, 8 * C, i * 2 * C, (const int2*)band_dev, (const float*)(cnull.p + i * C), (const float*)x1, x2,
                         mh, n_uncond, Bp, T, Tp, ctx_T, att_scale, (float*)nullptr, pao, status_flag));
            LAUNCH_CHECK();
        }
        if (Mcp > 0) {   // x2 = to_out(attn2) + x1 on the conditional rows (the kernel above wrote the unconditional ones)
            EpiStd ep = mk_epi(x2 + r0 * C, C, C);
            ep.bias = W.bo2;
            ep.res = x1 + x1c;
            ep.ldr = C;
            CKI(gemm_h(st, Mcp, C, {{psrc(pao + r0 * 2 * C, C, Mcp), 0, C, 0}}, W.wo2, ep, TAG_GEMM_PLAIN));
        }
        CKI(ln_pair(x2, Mp, nullptr, nullptr, W.ln3_g, W.ln3_b, pln, px2));
        // (below half a wave of row tiles the fused kernel leaves most SMs idle for a whole tile time: two finer-tiled GEMMs spread better)
        if (fused_ffn && (Mp + hx::HBM - 1) / hx::HBM >= ffn_min_tiles) {   // GEGLU, ff2 and proj_out in one kernel: the 768-wide intermediate stays in shared memory
            EpiStd ep = mk_epi(out, C, C);
            ep.bias = W.bffp;
            ep.res = h;
            ep.ldr = C;
            ep.res_mod = mh;
            CKI(ffn_h(st, Mp, pln, px2, W.bff1, W.bff1, W.wffp, ep, TAG_GEMM_PLAIN));
            return 0;
        }
        {   // GEGLU
            EpiGegluPair ep{pff, FF, 2 * FF, W.bff1, 1.0f, status_flag};
            CKI(gemm_h(st, Mp, 2 * FF, {{psrc(pln, C, Mp), 0, C, 0}}, W.wff1, ep, TAG_GEMM_PLAIN));
        }
        {   // out = proj_out(ff2(ff) + x2) + h as ONE contraction over [ff | x2] with the folded weights
            EpiStd ep = mk_epi(out, C, C);
            ep.bias = W.bffp;
            ep.res = h;
            ep.ldr = C;
            ep.res_mod = mh;
            CKI(gemm_h(st, Mp, C, {{psrc(pff, FF, Mp), 0, FF, 0}, {psrc(px2, C, Mp), 0, C, 0}}, W.wffp, ep, TAG_GEMM_PLAIN));
        }
        return 0;
    };

    {   // input conv (openaimodel.py:473-479): K = 96, straight from the fp32 latents (B, T, in_ch) into the padded row space
        ALoadConv3 al{x, nullptr, in_ch, 0, in_ch, Tp, Msp, src_batch, nullptr, nullptr, 3 * in_ch, 0};
        al.Tsrc = T;
        EpiStd ep = mk_epi(h0, C, C);
        ep.bias = b_in;
        CKI(gemm(st, Msp, C, 3 * in_ch, al, w_in, C, ep));
    }
    CKI(tap(h0));
    CKI(resblock(0, h0, Bs, nullptr, 0, Bs, A));       CKI(tap(A));
    CKI(transformer(0, A, Bs, h1));                    CKI(tap(h1));
    CKI(resblock(1, h1, Bp, nullptr, 0, Bp, Bb));      CKI(tap(Bb));
    CKI(transformer(1, Bb, Bp, A));                    CKI(tap(A));
    CKI(resblock(2, A, Bp, nullptr, 0, Bp, Bb));       CKI(tap(Bb));
    CKI(resblock(3, Bb, Bp, h1, Bp, Bp, A));           CKI(tap(A));
    CKI(transformer(2, A, Bp, Bb));                    CKI(tap(Bb));
    CKI(resblock(4, Bb, Bp, h0, Bs, Bp, A));           CKI(tap(A));
    CKI(transformer(3, A, Bp, Bb));                    CKI(tap(Bb));
    {   // out: GN + SiLU + conv3 -> in_ch   (openaimodel.py:665-669), written densely as (Bp, T, in_ch)
        CKI(gnp(Bb, Bp, Bp, 6, 1e-5f, out_gn_g, out_gn_b, nullptr, nullptr, pgn, nullptr, C, 0));
        const HSrc g = psrc(pgn, C, Mp);
        EpiStd ep = mk_epi(eps_out, in_ch, in_ch);
        ep.bias = b_out;
        ep.out_period = Tp;
        ep.out_valid = T;
        CKI(gemm_h(st, Mp, in_ch, {{g, 0, C, -1}, {g, 0, C, 0}, {g, 0, C, 1}}, w_out, ep, TAG_GEMM_CONV));
    }
    return 0;
}

// =====================================================================================================
// Denoising loop
// =====================================================================================================
int said_engine::denoise(const said_denoise_args& a, cudaStream_t user) {
    if (!ready) return fail("weights not committed");
    if (a.B <= 0 || a.T <= 0 || a.n_steps < 0) return fail("denoise: bad sizes");
    if (a.scheduler != 0 && a.scheduler != 1) return fail("denoise: scheduler must be 0 (DDIM) or 1 (DDPM)");
    if (ctx_B != a.B || ctx_T <= 0 || ctx_uncond != (a.do_cfg ? 1 : 0))
        return fail("denoise: said_prepare_context was not called for this (B, cfg)");
    const int B = a.B, T = a.T, Bp = a.do_cfg ? 2 * B : B;
    const long long n = (long long)T * in_ch, tot = (long long)B * n;
    CKI(ensure_denoiser_ws(Bp, T));
    CKI(ensure_band(T, ctx_T, own_stream));
    CK(lat.ensure((size_t)tot));
    CK(init_lat.ensure((size_t)tot));
    cudaStream_t st = own_stream;
    CK(cudaEventRecord(ev_in, user));
    CK(cudaStreamWaitEvent(st, ev_in, 0));

    if (a.resume) {   // continuation of a chunked loop: the latents come back as they were returned; init_lat keeps the first call's copy
        prepare_latents_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a.init_src_dev, 1.0f, nullptr, 1.0f, 0.0f, lat.p, nullptr, tot);
    } else {
        prepare_latents_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a.init_src_dev, a.init_scale, a.edit_noise_dev, a.edit_sqrt_a,
                                                                             a.edit_sqrt_b, lat.p, init_lat.p, tot);
    }
    LAUNCH_CHECK();
    if (a.n_steps == 0) {
        finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(lat.p, a.latent_scale, a.result_dev, tot);
        LAUNCH_CHECK();
    } else {
        CK(tvals.ensure((size_t)a.n_steps));
        CK(step_tab.ensure((size_t)a.n_steps * 8));
        CK(emb_tab.ensure((size_t)a.n_steps * 5 * C));
        CK(cudaMemcpyAsync(tvals.p, a.timesteps_host, a.n_steps * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(step_tab.p, a.step_table_host, (size_t)a.n_steps * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
        time_embed_table_kernel<<<a.n_steps, 256, 0, st>>>(tvals.p, te, emb_tab.p, nullptr);
        LAUNCH_CHECK();
        set_int_kernel<<<1, 1, 0, st>>>(step_ctr, 0);
        LAUNCH_CHECK();

        StepParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.pred = eps.p;
        sp.latents = lat.p;
        sp.B = B;
        sp.n = (int)n;
        sp.do_cfg = a.do_cfg;
        sp.gscale = a.guidance_scale;
        sp.grescale = a.guidance_rescale;
        sp.one_minus_grescale = (float)(1.0 - (double)a.guidance_rescale);
        sp.pred_type = a.prediction_type;
        sp.table = step_tab.p;
        sp.step_ptr = step_ctr;
        sp.n_steps = a.more ? a.n_steps + 1 : a.n_steps;   // "last iteration" (un-noised blend, result) only in the loop's final chunk
        sp.eta_noise = a.eta_noise_dev;
        sp.scheduler = a.scheduler;
        sp.init_latents = init_lat.p;
        sp.edit_noise = a.edit_noise_dev;
        sp.mask = a.mask_dev;
        sp.intermediates = a.intermediates_dev;
        sp.latent_scale = a.latent_scale;
        CK(result_buf.ensure((size_t)tot));
        sp.result = result_buf.p;       // engine-owned so that the cached graph does not depend on the caller's output tensor
        if (sp.mask && !sp.edit_noise) return fail("denoise: mask given without edit noise");

        // two half-batches on two streams (see `stream2`): plain generation only -- per-step user tensors (variance noise,
        // intermediates, editing) are indexed by the full batch inside the step kernel
        const bool halves = use_h(Bp * T) && !prof_on && B >= 2 && split_min_tiles > 0 &&
                            (Bp * (T + 1) + hx::HBM - 1) / hx::HBM >= split_min_tiles &&
                            (Bp * (T + 1) + hx::HBM - 1) / hx::HBM <= split_max_tiles && !a.eta_noise_dev && !a.intermediates_dev &&
                            !a.mask_dev && !a.edit_noise_dev;
        if (halves && !stream2) {
            CK(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        auto one_step = [&]() -> int {
            if (halves) {
                const int mult = a.do_cfg ? 2 : 1;
                CK(cudaEventRecord(ev_fork, st));
                CK(cudaStreamWaitEvent(stream2, ev_fork, 0));
                int rc = 0;
                for (int hh = 0; hh < 2 && rc == 0; ++hh) {
                    const int b0 = hh == 0 ? 0 : B / 2, bh = hh == 0 ? B / 2 : B - B / 2;
                    cudaStream_t sh = hh == 0 ? st : stream2;
                    hctx = HalfCtx{mult * b0 * (T + 1), mult * b0, b0, hh};
                    float* eps_h = eps.p + (size_t)mult * b0 * n;
                    rc = forward_h(sh, lat.p + (size_t)b0 * n, bh, mult * bh, a.do_cfg ? bh : 0, T, emb_tab.p, step_ctr, eps_h, nullptr);
                    if (rc == 0) {
                        StepParams sq = sp;
                        sq.pred = eps_h;
                        sq.latents = lat.p + (size_t)b0 * n;
                        sq.init_latents = init_lat.p + (size_t)b0 * n;
                        sq.result = result_buf.p + (size_t)b0 * n;
                        sq.B = bh;
                        cur_tag = TAG_STEP;
                        cudaError_t le = launch_ex(ddim_step_kernel, dim3(bh, DDIM_SPLIT), dim3(256), 0, sh, pdl, 1, sq);
                        if (le != cudaSuccess) rc = fail(std::string("step kernel launch failed: ") + cudaGetErrorString(le));
                        else rc = after_launch(sh);
                    }
                }
                hctx = HalfCtx{0, 0, 0, 0};
                // (join even after an error: a capture in progress must see its forked stream rejoin)
                cudaEventRecord(ev_join, stream2);
                cudaStreamWaitEvent(st, ev_join, 0);
                if (rc != 0) return rc;
                CK(launch_ex(add_int_kernel, dim3(1), dim3(1), 0, st, pdl, 1, step_ctr, 1));
                LAUNCH_CHECK();
                return 0;
            }
            if (use_h(Bp * T)) CKI(forward_h(st, lat.p, B, Bp, a.do_cfg ? B : 0, T, emb_tab.p, step_ctr, eps.p, nullptr));
            else CKI(forward(st, lat.p, B, Bp, a.do_cfg ? B : 0, T, emb_tab.p, step_ctr, eps.p, nullptr));
            cur_tag = TAG_STEP;
            CK(launch_ex(ddim_step_kernel, dim3(B, DDIM_SPLIT), dim3(256), 0, st, pdl, 1, sp));
            LAUNCH_CHECK();
            CK(launch_ex(add_int_kernel, dim3(1), dim3(1), 0, st, pdl, 1, step_ctr, 1));
            LAUNCH_CHECK();
            return 0;
        };
        if (a.use_graph && a.n_steps > 1) {
            GraphKey key;
            memset(&key, 0, sizeof(key));
            key.B = B; key.T = T; key.Tc = ctx_T; key.do_cfg = a.do_cfg; key.n_steps = sp.n_steps; key.pred_type = a.prediction_type;
            key.scheduler = a.scheduler; key.precision = precision; key.tc_min_rows = (tc_min_rows * 100003 + h_min_rows) * 2 + (halves ? 1 : 0);
            key.gscale = a.guidance_scale; key.grescale = a.guidance_rescale; key.latent_scale = a.latent_scale;
            key.eta_noise = a.eta_noise_dev; key.edit_noise = a.edit_noise_dev; key.mask = a.mask_dev;
            key.intermediates = a.intermediates_dev;
            key.alloc_gen = g_alloc_gen;
            if (!graph_exec || memcmp(&key, &graph_key, sizeof(key)) != 0) {
                if (graph_exec) {
                    CK(cudaStreamSynchronize(st));
                    drop_graph();
                }
                const long long before = launches;
                cudaGraph_t graph = nullptr;
                CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                const int rc = one_step();
                cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (rc != 0) {
                    if (graph) cudaGraphDestroy(graph);
                    return rc;
                }
                CK(ce);
                graph_launches = launches - before;
                launches = before;
                CK(cudaGraphInstantiate(&graph_exec, graph, 0));
                cudaGraphDestroy(graph);
                graph_key = key;
                ++graph_captures;
            }
            for (int s = 0; s < a.n_steps; ++s) CK(cudaGraphLaunch(graph_exec, st));
            launches += graph_launches * a.n_steps;
        } else {
            for (int s = 0; s < a.n_steps; ++s) CKI(one_step());
        }
        if (!a.more && a.result_dev) CK(cudaMemcpyAsync(a.result_dev, result_buf.p, (size_t)tot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (a.latents_out_dev)
        CK(cudaMemcpyAsync(a.latents_out_dev, lat.p, (size_t)tot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaEventRecord(ev_out, st));
    CK(cudaStreamWaitEvent(user, ev_out, 0));
    return 0;
}

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* said_last_error(void) { return g_err.c_str(); }
int said_version(void) { return 1; }

int said_create(int device, said_engine** out) {
    if (!out) return fail("said_create: null out pointer");
    *out = nullptr;
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail("said_create: no such CUDA device " + std::to_string(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(std::string("said_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                    std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only");
    CK(cudaSetDevice(device));
    std::unique_ptr<said_engine> e(new said_engine());
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
    *out = e.release();
    return 0;
}

void said_destroy(said_engine* e) { delete e; }

int said_set_tensor(said_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    if (!e || !name || !host_data || ndim < 0 || ndim > 8) return fail("said_set_tensor: bad arguments");
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 0) return fail("said_set_tensor: negative dimension");
        t.shape.push_back(shape[i]);
        n *= shape[i];
    }
    t.data.assign(host_data, host_data + n);
    e->raw[name] = std::move(t);
    e->ready = false;
    return 0;
}

int said_commit_weights(said_engine* e) {
    if (!e) return fail("null engine");
    return e->commit();
}

int said_weights_ready(const said_engine* e) { return e && e->ready ? 1 : 0; }

int said_get_config(const said_engine* e, int* in_channels, int* ctx_dim, int* enc_hidden) {
    if (!e || !e->ready) return fail("weights not committed");
    if (in_channels) *in_channels = e->in_ch;
    if (ctx_dim) *ctx_dim = e->ctx_dim;
    if (enc_hidden) *enc_hidden = e->enc_hidden;
    return 0;
}

int said_normalize_audio(said_engine* e, const float* wave_dev, int B, int T_a, float* out_dev, void* stream) {
    if (!e) return fail("null engine");
    if (B <= 0 || T_a <= 0 || !wave_dev || !out_dev) return fail("said_normalize_audio: bad arguments");
    CK(cudaSetDevice(e->device));
    normalize_audio_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(wave_dev, T_a, out_dev);
    ++e->launches;
    CK(cudaGetLastError());
    return 0;
}

int said_resample_mono(said_engine* e, const float* wave_dev, int channels, int n_in, int orig, int nw, int width, const float* bank_dev,
                       float* out_dev, int n_out, void* stream) {
    if (!e) return fail("null engine");
    if (!wave_dev || !bank_dev || !out_dev || channels <= 0 || n_in <= 0 || orig <= 0 || nw <= 0 || width < 0 || n_out <= 0)
        return fail("said_resample_mono: bad arguments");
    CK(cudaSetDevice(e->device));
    resample_mono_kernel<<<(n_out + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wave_dev, channels, n_in, orig, nw, width, bank_dev, out_dev, n_out);
    ++e->launches;
    CK(cudaGetLastError());
    return 0;
}

int said_encode_audio(said_engine* e, const float* wave_dev, int B, int T_a, int T, float* emb_out_dev, void* stream) {
    if (!e) return fail("null engine");
    CK(cudaSetDevice(e->device));
    return e->encode_audio(wave_dev, B, T_a, T, emb_out_dev, (cudaStream_t)stream);
}

int said_prepare_context(said_engine* e, const float* emb_dev, int B, int T, int with_uncond, void* stream) {
    if (!e) return fail("null engine");
    CK(cudaSetDevice(e->device));
    return e->prepare_context(emb_dev, B, T, with_uncond, (cudaStream_t)stream);
}

int said_denoise(said_engine* e, const said_denoise_args* args, void* stream) {
    if (!e || !args) return fail("null engine / args");
    CK(cudaSetDevice(e->device));
    return e->denoise(*args, (cudaStream_t)stream);
}

int said_denoiser_forward(said_engine* e, const float* x_dev, const float* timesteps_host, const float* ctx_dev, int Bp,
                          int T, int T_ctx, float* out_dev, float* taps_dev, void* stream) {
    if (!e) return fail("null engine");
    if (!e->ready) return fail("weights not committed");
    if (Bp <= 0 || T <= 0 || T_ctx <= 0) return fail("denoiser_forward: empty batch");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    CKI(e->prepare_context(ctx_dev, Bp, T_ctx, 0, st));
    CKI(e->ensure_denoiser_ws(Bp, T));
    CKI(e->ensure_band(T, T_ctx, st));
    CK(e->tvals.ensure((size_t)Bp));
    CK(e->emb_tab.ensure((size_t)Bp * 5 * C));
    CK(cudaMemcpyAsync(e->tvals.p, timesteps_host, Bp * sizeof(float), cudaMemcpyHostToDevice, st));
    time_embed_table_kernel<<<Bp, 256, 0, st>>>(e->tvals.p, e->te, e->emb_tab.p, nullptr);
    ++e->launches;
    CK(cudaGetLastError());
    if (e->use_h(Bp * T)) CKI(e->forward_h(st, x_dev, Bp, Bp, 0, T, e->emb_tab.p, nullptr, out_dev, taps_dev));
    else CKI(e->forward(st, x_dev, Bp, Bp, 0, T, e->emb_tab.p, nullptr, out_dev, taps_dev));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int said_op_ffn_h(said_engine* e, const float* ln_dev, const float* x2_dev, const float* res_dev, int M, const float* w1_host,
                  const float* b1_dev, const float* w2_host, const float* b2_dev, float* out_dev, void* stream) {
    // unit test of the fused feed-forward (ffn_h.cuh): out = [geglu(ln W1 + b1) | x2] W2 + b2 + res, W1 (192, 1536) K-major with
    // value / gate columns interleaved (the layout commit_denoiser builds), b1 (1536) interleaved alike, W2 (960, 192), all fp32;
    // ln and x2 are converted to the pair format first
    if (!e) return fail("null engine");
    if (M <= 0 || !ln_dev || !x2_dev || !w1_host || !b1_dev || !w2_host || !out_dev) return fail("said_op_ffn_h: bad arguments");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int Cc = hx::FFN_C, FFc = hx::FFN_NJ * hx::FFN_JC;
    static DevBuf pairs;
    CK(pairs.ensure((size_t)M * Cc * 2));
    if (!e->status_flag) {
        CK(cudaMalloc((void**)&e->status_flag, sizeof(int)));
        CK(cudaMemset(e->status_flag, 0, sizeof(int)));
    }
    CKI(e->ensure_ffn_split((size_t)M));
    __half* pl = reinterpret_cast<__half*>(pairs.p);
    __half* px = pl + (size_t)M * 2 * Cc;
    const long long nq = (long long)M * (Cc / 4);
    f32_to_pair_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(ln_dev, M, Cc, pl, e->status_flag);
    f32_to_pair_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(x2_dev, M, Cc, px, e->status_flag);
    CK(cudaGetLastError());
    std::vector<uint16_t> i1, i2;
    const int e1 = hx::pack_weights_h(w1_host, Cc, 2 * FFc, 2 * FFc, 256, i1);
    const int e2 = hx::pack_weights_h(w2_host, FFc + Cc, Cc, Cc, 192, i2);
    uint8_t* wd = nullptr;
    const size_t b1s = i1.size() * sizeof(uint16_t), b2s = i2.size() * sizeof(uint16_t);
    CK(cudaMalloc((void**)&wd, b1s + b2s));
    CK(cudaMemcpyAsync(wd, i1.data(), b1s, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(wd + b1s, i2.data(), b2s, cudaMemcpyHostToDevice, st));
    const float *k1 = reinterpret_cast<const float*>(wd), *k2 = reinterpret_cast<const float*>(wd + b1s);
    e->hmap[k1] = said_engine::HW{wd, Cc, 2 * FFc, 256, e1};
    e->hmap[k2] = said_engine::HW{wd + b1s, FFc + Cc, Cc, 192, e2};
    EpiStd ep = mk_epi(out_dev, Cc, Cc);
    ep.bias = b2_dev;
    if (res_dev) { ep.res = res_dev; ep.ldr = Cc; }
    const int rc = e->ffn_h(st, M, pl, px, k1, b1_dev, k2, ep, said_engine::TAG_GEMM_PLAIN);
    cudaError_t se = cudaStreamSynchronize(st);
    e->hmap.erase(k1);
    e->hmap.erase(k2);
    cudaFree(wd);
    if (rc != 0) return rc;
    CK(se);
    return 0;
}

int said_op_gemm_h_bench(said_engine* e, int M, int Cin, int taps, int N, int with_residual, int dbg, int iters, float* ms_out) {
    // diagnostics: average milliseconds of the fp16x3 GEMM on scratch (zero) operands, with parts of it disabled by dbg
    if (!e || !ms_out) return fail("said_op_gemm_h_bench: bad arguments");
    if (!(taps == 1 || taps == 3) || Cin % hx::HBK != 0 || (N % 192 != 0 && N != 32)) return fail("said_op_gemm_h_bench: unsupported shape");
    CK(cudaSetDevice(e->device));
    static DevBuf a, o, r;
    if (with_residual == 3) {   // the fused feed-forward (M rows; Cin, taps, N ignored) on zero operands
        constexpr int C = hx::FFN_C, FF = hx::FFN_NJ * hx::FFN_JC;
        CK(a.ensure_zero((size_t)M * C * 2));     // two pair tensors (ln, x2) of 2 * C halves per row
        CK(o.ensure_zero((size_t)M * C));
        CK(r.ensure_zero((size_t)M * C));
        static DevBuf b1;
        CK(b1.ensure_zero((size_t)2 * FF));
        const size_t w1b = (size_t)hx::FFN_NP * 6 * hx::FFN_W1_PLANE, w2b = (size_t)(hx::FFN_NJ + 3) * 2 * hx::FFN_W2_PLANE;
        uint8_t* wd = nullptr;
        CK(cudaMalloc((void**)&wd, w1b + w2b));
        CK(cudaMemset(wd, 0, w1b + w2b));
        const float *k1 = reinterpret_cast<const float*>(wd), *k2 = reinterpret_cast<const float*>(wd + w1b);
        e->hmap[k1] = said_engine::HW{wd, C, 2 * FF, 256, 0};
        e->hmap[k2] = said_engine::HW{wd + w1b, FF + C, C, 192, 0};
        if (!e->status_flag) {
            CK(cudaMalloc((void**)&e->status_flag, sizeof(int)));
            CK(cudaMemset(e->status_flag, 0, sizeof(int)));
        }
        EpiStd ep = mk_epi(o.p, C, C);
        ep.res = r.p;
        ep.ldr = C;
        CKI(e->ensure_ffn_split((size_t)M));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        int rc = 0;
        const __half* pa = reinterpret_cast<const __half*>(a.p);
        long long* trace_dev = nullptr;
        if (dbg & 16) {
            CK(cudaMalloc((void**)&trace_dev, 256 * sizeof(long long)));
            CK(cudaMemset(trace_dev, 0, 256 * sizeof(long long)));
        }
        for (int it = -2; it < iters && rc == 0; ++it) {
            if (it == 0) CK(cudaEventRecord(e0, 0));
            rc = e->ffn_h((cudaStream_t)0, M, pa, pa + (size_t)M * 2 * C, k1, b1.p, k2, ep, said_engine::TAG_GEMM_PLAIN, dbg, trace_dev);
        }
        if (trace_dev) {
            long long h[256];
            cudaDeviceSynchronize();
            cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
            cudaFree(trace_dev);
            const long long t0 = h[60];
            fprintf(stderr, "[ffn trace] tile start 0, g4(0) issued %lld, tile issued %lld; epilogue: wait %lld, acc5 ready %lld, done %lld\n", h[61] - t0,
                    h[62] - t0, h[130] - t0, h[131] - t0, h[132] - t0);
            for (int j = 0; j < hx::FFN_NJ; ++j)
                fprintf(stderr, "[ffn trace] j=%2d  mma: g4(j+1) issued %6lld, ff_full %6lld, g5 issued %6lld | geglu: acc4_full %6lld, loaded %6lld, math done %6lld, ff_empty %6lld, arrived %6lld\n",
                        j, h[j * 4] - t0, h[j * 4 + 1] - t0, h[j * 4 + 2] - t0, h[64 + j * 5] - t0, h[64 + j * 5 + 1] - t0, h[64 + j * 5 + 2] - t0,
                        h[64 + j * 5 + 3] - t0, h[64 + j * 5 + 4] - t0);
        }
        cudaError_t se = cudaEventRecord(e1, 0);
        if (se == cudaSuccess) se = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (se == cudaSuccess) se = cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        e->hmap.erase(k1);
        e->hmap.erase(k2);
        cudaFree(wd);
        if (rc != 0) return rc;
        CK(se);
        *ms_out = ms / iters;
        return 0;
    }
    CK(a.ensure_zero((size_t)M * Cin));
    CK(o.ensure_zero((size_t)M * N));
    CK(r.ensure_zero((size_t)M * N));
    const int K = taps * Cin;
    const int bnb = N == 32 ? 32 : 192;
    const size_t wbytes = (size_t)(N / bnb) * (K / hx::HBK) * 2 * bnb * hx::HROW;
    uint8_t* wd = nullptr;
    CK(cudaMalloc((void**)&wd, wbytes));
    CK(cudaMemset(wd, 0, wbytes));
    const float* key = reinterpret_cast<const float*>(wd);
    e->hmap[key] = said_engine::HW{wd, K, N, bnb, 0};
    const said_engine::HSrc src{reinterpret_cast<const __half*>(a.p), Cin, M};
    EpiStd ep = mk_epi(o.p, N, N);
    if (with_residual == 1) { ep.res = r.p; ep.ldr = N; }
    static DevBuf gb;
    CK(gb.ensure_zero((size_t)N));
    if (!e->status_flag) {
        CK(cudaMalloc((void**)&e->status_flag, sizeof(int)));
        CK(cudaMemset(e->status_flag, 0, sizeof(int)));
    }
    const EpiGegluPair eg{reinterpret_cast<__half*>(o.p), N / 2, N, gb.p, 1.0f, e->status_flag};   // with_residual == 2: the GEGLU epilogue (pair output)
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int rc = 0;
    for (int it = -2; it < iters && rc == 0; ++it) {
        if (it == 0) CK(cudaEventRecord(e0, 0));
        if (with_residual == 2) rc = e->gemm_h((cudaStream_t)0, M, N, {{src, 0, Cin, 0}}, key, eg, said_engine::TAG_GEMM_PLAIN, dbg);
        else if (taps == 1) rc = e->gemm_h((cudaStream_t)0, M, N, {{src, 0, Cin, 0}}, key, ep, said_engine::TAG_GEMM_PLAIN, dbg);
        else rc = e->gemm_h((cudaStream_t)0, M, N, {{src, 0, Cin, -1}, {src, 0, Cin, 0}, {src, 0, Cin, 1}}, key, ep, said_engine::TAG_GEMM_CONV, dbg);
    }
    cudaError_t se = cudaEventRecord(e1, 0);
    if (se == cudaSuccess) se = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (se == cudaSuccess) se = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e->hmap.erase(key);
    cudaFree(wd);
    if (rc != 0) return rc;
    CK(se);
    *ms_out = ms / iters;
    return 0;
}

int said_eval_commit_bcvae(said_engine* e) {
    if (!e) return fail("null engine");
    return e->commit_bcvae();
}

int said_eval_bcvae_latents(said_engine* e, const float* coeffs_dev, int B, int T, int step, float* latents_out_dev, void* stream) {
    if (!e) return fail("null engine");
    if (!e->bcv_ready) return fail("BCVAE weights not committed (said_set_tensor \"bcvae.encoder.*\" + said_eval_commit_bcvae)");
    if (B <= 0 || step <= 0 || T < BCV_SEQ) return fail("said_eval_bcvae_latents: need B > 0, step > 0 and at least 120 frames");
    CK(cudaSetDevice(e->device));
    const int nw = (T - BCV_SEQ) / step + 1;
    CK(cudaFuncSetAttribute(bcvae_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bcvae_smem_bytes()));
    bcvae_encode_kernel<<<B * nw, 256, bcvae_smem_bytes(), (cudaStream_t)stream>>>(coeffs_dev, T, nw, step, e->bcv, latents_out_dev);
    ++e->launches;
    CK(cudaGetLastError());
    return 0;
}

int said_eval_frechet(said_engine* e, const float* lat1_dev, int n1, const float* lat2_dev, int n2, double* out4_host, void* stream) {
    if (!e || !out4_host) return fail("said_eval_frechet: bad arguments");
    if (n1 < 2 || n2 < 2) return fail("said_eval_frechet: at least two latents per set");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, 4 * sizeof(double)));
    CK(cudaFuncSetAttribute(frechet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)frechet_smem_bytes()));
    frechet_kernel<<<1, 256, frechet_smem_bytes(), st>>>(lat1_dev, n1, lat2_dev, n2, d);
    ++e->launches;
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) le = cudaMemcpyAsync(out4_host, d, 4 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (le == cudaSuccess) le = cudaStreamSynchronize(st);
    cudaFree(d);
    CK(le);
    return 0;
}

int said_check_status(said_engine* e, void* stream, int* status_out) {
    if (!e || !status_out) return fail("said_check_status: bad arguments");
    *status_out = 0;
    if (!e->status_flag) return 0;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    int v = 0;
    CK(cudaMemcpyAsync(&v, e->status_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (v != 0) CK(cudaMemsetAsync(e->status_flag, 0, sizeof(int), st));
    *status_out = v;
    return 0;
}

int said_op_gemm_h(said_engine* e, const float* a_dev, int M, int Cin, int taps, const float* wt_host, int N, const float* bias_dev,
                   float* out_dev, void* stream) {
    // unit test of the fp16x3 GEMM: out (M, N) = sum over taps of A[m + tap - (taps - 1) / 2, :] . Wt[tap * Cin : (tap + 1) * Cin, :] + bias
    // (rows outside [0, M) are zero); A is converted to the pair format first.  N must be 192 or 32, Cin a multiple of 64.
    if (!e) return fail("null engine");
    if (!(taps == 1 || taps == 3) || Cin % hx::HBK != 0 || (N != 192 && N != 32 && N % 192 != 0)) return fail("said_op_gemm_h: unsupported shape");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    static DevBuf apair;
    CK(apair.ensure((size_t)M * Cin));
    if (!e->status_flag) {
        CK(cudaMalloc((void**)&e->status_flag, sizeof(int)));
        CK(cudaMemset(e->status_flag, 0, sizeof(int)));
    }
    const long long nq = (long long)M * (Cin / 4);
    f32_to_pair_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(a_dev, M, Cin, reinterpret_cast<__half*>(apair.p), e->status_flag);
    CK(cudaGetLastError());
    const int K = taps * Cin, bn = N == 32 ? 32 : 192;
    std::vector<uint16_t> himg;
    const int ex = hx::pack_weights_h(wt_host, K, N, N, bn, himg);
    uint8_t* wd = nullptr;
    CK(cudaMalloc((void**)&wd, himg.size() * sizeof(uint16_t)));
    CK(cudaMemcpyAsync(wd, himg.data(), himg.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    const float* key = reinterpret_cast<const float*>(wd);
    e->hmap[key] = said_engine::HW{wd, K, N, bn, ex};
    const said_engine::HSrc src{reinterpret_cast<const __half*>(apair.p), Cin, M};
    EpiStd ep = mk_epi(out_dev, N, N);
    ep.bias = bias_dev;
    int rc;
    if (taps == 1) rc = e->gemm_h(st, M, N, {{src, 0, Cin, 0}}, key, ep, said_engine::TAG_GEMM_PLAIN);
    else rc = e->gemm_h(st, M, N, {{src, 0, Cin, -1}, {src, 0, Cin, 0}, {src, 0, Cin, 1}}, key, ep, said_engine::TAG_GEMM_CONV);
    cudaError_t se = cudaStreamSynchronize(st);
    e->hmap.erase(key);
    cudaFree(wd);
    if (rc != 0) return rc;
    CK(se);
    return 0;
}

int said_op_ddim_step(said_engine* e, const float* pred_dev, float* latents_dev, int B, int n, int do_cfg,
                      float guidance_scale, float guidance_rescale, int prediction_type, const float* row8_host,
                      const float* eta_noise_dev, int scheduler, void* stream) {
    if (!e) return fail("null engine");
    if (scheduler != 0 && scheduler != 1) return fail("said_op_ddim_step: scheduler must be 0 (DDIM) or 1 (DDPM)");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(e->step_tab.ensure(8));
    if (!e->step_ctr) CK(cudaMalloc((void**)&e->step_ctr, sizeof(int)));
    CK(cudaMemcpyAsync(e->step_tab.p, row8_host, 8 * sizeof(float), cudaMemcpyHostToDevice, st));
    set_int_kernel<<<1, 1, 0, st>>>(e->step_ctr, 0);
    StepParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.pred = pred_dev;
    sp.latents = latents_dev;
    sp.B = B;
    sp.n = n;
    sp.do_cfg = do_cfg;
    sp.gscale = guidance_scale;
    sp.grescale = guidance_rescale;
    sp.one_minus_grescale = (float)(1.0 - (double)guidance_rescale);
    sp.pred_type = prediction_type;
    sp.table = e->step_tab.p;
    sp.step_ptr = e->step_ctr;
    sp.n_steps = 2;   // never "last": no result write
    sp.scheduler = scheduler;
    sp.eta_noise = eta_noise_dev;
    sp.latent_scale = 1.0f;
    ddim_step_kernel<<<dim3(B, DDIM_SPLIT), 256, 0, st>>>(sp);
    e->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

int said_op_self_attention(said_engine* e, const float* qkv_dev, int B, int T, int heads, int head_dim, float* out_dev,
                           void* stream) {
    if (!e) return fail("null engine");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int Cw = heads * head_dim;
    const float scale = 1.0f / sqrtf((float)head_dim);
    const dim3 grid((T + ATT_QTILE - 1) / ATT_QTILE, heads, B);
    if (head_dim == 32) {
        CK(cudaFuncSetAttribute(self_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attention_smem_bytes<32>()));
        self_attention_kernel<32><<<grid, ATT_THREADS, attention_smem_bytes<32>(), st>>>(qkv_dev, 3 * Cw, 0, Cw, 2 * Cw, T, scale, out_dev, Cw);
    } else if (head_dim == 64) {
        CK(cudaFuncSetAttribute(self_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attention_smem_bytes<64>()));
        self_attention_kernel<64><<<grid, ATT_THREADS, attention_smem_bytes<64>(), st>>>(qkv_dev, 3 * Cw, 0, Cw, 2 * Cw, T, scale, out_dev, Cw);
    } else {
        return fail("self_attention: head_dim must be 32 or 64");
    }
    ++e->launches;
    CK(cudaGetLastError());
    return 0;
}

int said_op_self_attention_tc(said_engine* e, const float* qkv_dev, int B, int T, int heads, float* out_dev, void* stream) {
    if (!e) return fail("null engine");
    if (T > tc::ATC_MAXKEYS) return fail("self_attention_tc: at most 304 keys");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int Cw = heads * 32;
    CK(cudaFuncSetAttribute(tc::self_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::attention_tc_smem_bytes(T)));
    tc::self_attention_tc_kernel<<<dim3(heads, B), tc::ATC_THREADS, tc::attention_tc_smem_bytes(T), st>>>(
        qkv_dev, 3 * Cw, 0, Cw, 2 * Cw, T, 1.0f / sqrtf(32.0f), out_dev, Cw, T, nullptr, nullptr);
    ++e->launches;
    CK(cudaGetLastError());
    return 0;
}

int said_op_self_attention_h(said_engine* e, const float* qkv_dev, int B, int T, int heads, float* out_dev, void* stream) {
    if (!e) return fail("null engine");
    if (T > hx::AH_MAXT) return fail("self_attention_h: at most 512 keys");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int Cw = heads * 32;
    const int ag = hx::attention_h_groups(T, heads);
    const size_t smem = hx::attention_h_smem_bytes(T, ag);
    CK(cudaFuncSetAttribute(hx::self_attention_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long* tr = nullptr;
    if (getenv("SAID_ATTN_TRACE")) {
        CK(cudaMalloc((void**)&tr, 64 * sizeof(long long)));
        CK(cudaMemset(tr, 0, 64 * sizeof(long long)));
    }
    hx::self_attention_h_kernel<<<dim3(heads / ag, B), hx::AH_THREADS * ag, smem, st>>>(
        qkv_dev, 3 * Cw, 0, Cw, 2 * Cw, T, 1.0f / sqrtf(32.0f), out_dev, Cw, T, nullptr, nullptr, (uint32_t)hx::attention_h_group_bytes(T), tr);
    ++e->launches;
    CK(cudaGetLastError());
    if (tr) {   // diagnostics: phase time line of CTA (0, 0), group 0, thread 0 (cycles since kernel entry)
        long long h[64];
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(tr);
        fprintf(stderr, "[said] attention_h trace T=%d B=%d groups=%d:", T, B, ag);
        for (int i = 1; i < (int)h[63] && i < 62; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
        fprintf(stderr, "\n");
    }
    return 0;
}

long long said_launch_count(const said_engine* e) { return e ? e->launches : 0; }
long long said_graph_captures(const said_engine* e) { return e ? e->graph_captures : 0; }

int said_op_gemm_tc_bench(said_engine* e, int M, int K, int nsplit, int with_residual, int dbg, int iters, float* ms_out) {
    // diagnostics: times the tcgen05 GEMM (N = 192, plain loader) on scratch buffers with parts of it disabled
    if (!e) return fail("null engine");
    CK(cudaSetDevice(e->device));
    const int N = 192;
    static DevBuf a, w, o, r;
    CK(a.ensure((size_t)M * K));
    CK(o.ensure((size_t)M * N));
    CK(r.ensure((size_t)M * N));
    CK(w.ensure((size_t)K * N * 2));
    CK(cudaMemset(a.p, 0, (size_t)M * K * 4));
    CK(cudaMemset(r.p, 0, (size_t)M * N * 4));
    CK(cudaMemset(w.p, 0, (size_t)K * N * 2 * 4));
    EpiStd ep = mk_epi(o.p, N, N);
    if (with_residual) { ep.res = r.p; ep.ldr = N; }
    ALoadPlain al = mk_plain(a.p, K, M);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int stride = 2 * N * tc::BK;
    for (int it = -2; it < iters; ++it) {
        if (it == 0) CK(cudaEventRecord(e0, 0));
        cudaError_t le;
        if (dbg & 256)   // A-in-TMEM variant
            le = nsplit == 3 ? tc::launch_gemm_tca<192, 3>(0, e->num_sms, M, N, K, al, w.p, stride, ep)
                             : tc::launch_gemm_tca<192, 1>(0, e->num_sms, M, N, K, al, w.p, stride, ep);
        else
            le = nsplit == 3 ? tc::launch_gemm_tc<192, 3>(0, e->num_sms, M, N, K, al, w.p, stride, ep, dbg)
                             : tc::launch_gemm_tc<192, 1>(0, e->num_sms, M, N, K, al, w.p, stride, ep, dbg);
        CK(le);
    }
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

int said_set_precision(said_engine* e, int mode, int tc_min_rows, int encoder_mode) {
    if (!e) return fail("null engine");
    if (mode < 0 || mode > 3) return fail("said_set_precision: mode must be 0 (fp32 FFMA), 1 (3xTF32 tcgen05), 2 (TF32 tcgen05) or 3 (fp16 hi/lo x3 tcgen05)");
    if (encoder_mode < 0 || encoder_mode > 3) return fail("said_set_precision: encoder_mode must be 0, 1, 2 or 3");
    e->precision = mode;
    e->enc_precision = encoder_mode;
    if (tc_min_rows > 0) e->tc_min_rows = e->h_min_rows = tc_min_rows;
    else if (tc_min_rows < 0) { e->tc_min_rows = said_engine::TC_MIN_ROWS_DEFAULT; e->h_min_rows = said_engine::H_MIN_ROWS_DEFAULT; }
    e->a_in_tmem = getenv("SAID_TC_TMEM_A") ? 1 : 0;   // study aid: SAID_TC_TMEM_A=1 selects the A-through-TMEM kernel
    return 0;
}

int said_profile_begin(said_engine* e) {
    if (!e) return fail("null engine");
    for (auto& pe : e->prof_ev) cudaEventDestroy(pe.second);
    e->prof_ev.clear();
    e->prof_on = true;
    return 0;
}

int said_profile_end(said_engine* e, double* ms_out, long long* count_out, int n) {
    if (!e) return fail("null engine");
    e->prof_on = false;
    CK(cudaSetDevice(e->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < n; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
    // time attributed to a launch = interval between the previous launch's end event and its own end
    // event (kernels of one stream run back to back, so this is the kernel's duration plus launch gap)
    for (size_t i = 1; i < e->prof_ev.size(); ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e->prof_ev[i - 1].second, e->prof_ev[i].second));
        const int tag = e->prof_ev[i].first;
        if (tag >= 0 && tag < n) { ms_out[tag] += ms; count_out[tag] += 1; }
    }
    for (auto& pe : e->prof_ev) cudaEventDestroy(pe.second);
    e->prof_ev.clear();
    return 0;
}

}  // extern "C"
