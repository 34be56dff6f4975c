// Time-embedding tables and the fused CFG-combine + scheduler-step + editing-blend kernel.
#pragma once
#include "common.cuh"

namespace said {

// ------------------------------------------------------------------------------------------------
// Time embedding, hoisted out of the step loop: everything the denoiser derives from the timestep
// (ldm/util.py:66-90 timestep_embedding -> openaimodel.py:463-468 time_embed -> the five ResBlocks'
// emb_layers, openaimodel.py:171-177) depends only on t, so it is evaluated once per distinct
// timestep into a table  out[row][r][192]  (r = ResBlock in execution order).  One CTA per row.
// Weights are in their PyTorch (out, in) layout; one warp per output, lanes over the contraction.
// ------------------------------------------------------------------------------------------------
struct TimeEmbedWeights {
    const float* freqs;          // (96)  exp(-ln(1e4) k / 96), computed by the host exactly as the reference does
    const float* w1; const float* b1;   // time_embed.0  (768,192)
    const float* w2; const float* b2;   // time_embed.2  (768,768)
    const float* wr[5]; const float* br[5];   // emb_layers.1 of the 5 ResBlocks (192,768)
};

__global__ void __launch_bounds__(256)
time_embed_table_kernel(const float* __restrict__ tvals, TimeEmbedWeights w, float* __restrict__ out /*(rows,5,192)*/,
                        float* __restrict__ emb_out /*(rows,768) or null*/) {
    __shared__ float x0[192], h1[768], e2[768];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float t = tvals[row];
    if (tid < 192) {
        const float a = t * __ldg(w.freqs + (tid % 96));
        x0[tid] = tid < 96 ? cosf(a) : sinf(a);     // [cos | sin]  (util.py:83)
    }
    __syncthreads();
    for (int o = warp; o < 768; o += 8) {
        float acc = 0.f;
        for (int k = lane; k < 192; k += 32) acc = fmaf(__ldg(w.w1 + o * 192 + k), x0[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) h1[o] = silu(acc + __ldg(w.b1 + o));
    }
    __syncthreads();
    for (int o = warp; o < 768; o += 8) {
        float acc = 0.f;
        for (int k = lane; k < 768; k += 32) acc = fmaf(__ldg(w.w2 + o * 768 + k), h1[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            const float e = acc + __ldg(w.b2 + o);
            if (emb_out) emb_out[(long long)row * 768 + o] = e;
            e2[o] = silu(e);
        }
    }
    __syncthreads();
    for (int oo = warp; oo < 5 * 192; oo += 8) {
        const int r = oo / 192, o = oo - r * 192;
        const float* wr = w.wr[r];
        float acc = 0.f;
        for (int k = lane; k < 768; k += 32) acc = fmaf(__ldg(wr + o * 768 + k), e2[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[((long long)row * 5 + r) * 192 + o] = acc + __ldg(w.br[r] + o);
    }
}

// ------------------------------------------------------------------------------------------------
// One diffusion-step epilogue (diffusion.py:430-456): classifier-free-guidance combine, optional
// rescale_noise_cfg, DDIMScheduler.step (diffusers 0.19 scheduling_ddim.py, restated in
// said_b200/scheduler.py), optional eta noise, optional editing blend, optional intermediate dump.
// grid (clip, DDIM_SPLIT): each CTA updates a slice of the clip (for the rescale every CTA of a clip reduces the
// whole clip itself, redundantly but deterministically).  All per-step scalars come from a
// host-built table row indexed by the device-resident step counter so that one captured CUDA graph
// serves every step.  Arithmetic uses explicit round-to-nearest mul/add (no FMA contraction) in the
// reference's operation order, so the step is bit-identical to the fp32 CPU formulas.
// table row: [sqrt_a, sqrt_b, sqrt_a_prev, dir_coef, sigma, clip(-1: off), blend_sa, blend_sb]
// ------------------------------------------------------------------------------------------------
struct StepParams {
    const float* pred;          // (B', n): [uncond(B) ; cond(B)] under CFG, else (B, n)
    float* latents;             // (B, n) in/out
    int B, n;                   // n = T * in_channels
    int do_cfg;
    float gscale, grescale, one_minus_grescale;
    int pred_type;              // 0 epsilon, 1 sample, 2 v_prediction
    const float* table;         // (n_steps, 8)
    const int* step_ptr;
    int n_steps;
    const float* eta_noise;     // (n_steps, B, n) or null
    const float* init_latents;  // (B, n) or null   (editing blend active iff mask != null)
    const float* edit_noise;    // (B, n)
    const float* mask;          // (B, n) or null
    float* intermediates;       // (n_steps, B, n) or null: latents / latent_scale before the step
    float latent_scale;
    float* result;              // (B, n): clamp(latents / latent_scale, 0, 1), written on the last step
    int scheduler;              // 0: DDIMScheduler.step, 1: DDPMScheduler.step (ancestral; table row holds c0, c1, std)
};

constexpr int DDIM_SPLIT = 8;
__global__ void __launch_bounds__(256)
ddim_step_kernel(StepParams p) {
    __shared__ double red[4][8];
    __shared__ float s_ratio;
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x, tid = threadIdx.x;
    const int step = *p.step_ptr;
    const float* row = p.table + (long long)step * 8;
    const float sa = row[0], sb = row[1], sap = row[2], dir = row[3], sigma = row[4], clip = row[5];
    const float bsa = row[6], bsb = row[7];
    const float* cond = p.pred + (long long)(p.do_cfg ? p.B + b : b) * p.n;
    const float* unc = p.pred + (long long)b * p.n;
    float* lat = p.latents + (long long)b * p.n;

    float ratio = 1.f;
    if (p.do_cfg && p.grescale > 0.f) {
        // rescale_noise_cfg: unbiased std over all elements of the clip, of cond and of the combined prediction
        double s0 = 0, q0 = 0, s1 = 0, q1 = 0;
        for (int i = tid; i < p.n; i += blockDim.x) {
            const float c = cond[i], u = unc[i];
            const float g = __fadd_rn(c, __fmul_rn(p.gscale, __fsub_rn(c, u)));
            s0 += c; q0 += (double)c * c; s1 += g; q1 += (double)g * g;
        }
        s0 = warp_sum(s0); q0 = warp_sum(q0); s1 = warp_sum(s1); q1 = warp_sum(q1);
        if ((tid & 31) == 0) { red[0][tid >> 5] = s0; red[1][tid >> 5] = q0; red[2][tid >> 5] = s1; red[3][tid >> 5] = q1; }
        __syncthreads();
        if (tid == 0) {
            double t[4];
            for (int k = 0; k < 4; ++k) { t[k] = 0; for (int w = 0; w < 8; ++w) t[k] += red[k][w]; }
            const double n = p.n;
            const double var_c = (t[1] - t[0] * t[0] / n) / (n - 1.0);
            const double var_g = (t[3] - t[2] * t[2] / n) / (n - 1.0);
            const float std_c = (float)sqrt(var_c > 0 ? var_c : 0.0), std_g = (float)sqrt(var_g > 0 ? var_g : 0.0);
            s_ratio = __fdiv_rn(std_c, std_g);
        }
        __syncthreads();
        ratio = s_ratio;
    }

    const long long soff = ((long long)step * p.B + b) * p.n;
    const bool last = step == p.n_steps - 1;
    const int per = (p.n + gridDim.y - 1) / gridDim.y;
    const int i_end = min(p.n, (int)(blockIdx.y + 1) * per);
    for (int i = blockIdx.y * per + tid; i < i_end; i += blockDim.x) {
        const float x = lat[i];
        if (p.intermediates) p.intermediates[soff + i] = __fdiv_rn(x, p.latent_scale);
        float e = cond[i];
        if (p.do_cfg) {
            const float u = unc[i];
            e = __fadd_rn(e, __fmul_rn(p.gscale, __fsub_rn(e, u)));
            if (p.grescale > 0.f) {
                const float resc = __fmul_rn(e, ratio);
                e = __fadd_rn(__fmul_rn(p.grescale, resc), __fmul_rn(p.one_minus_grescale, e));
            }
        }
        float x0, eps;
        if (p.pred_type == 0) {
            x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sb, e)), sa);
            eps = e;
        } else if (p.pred_type == 1) {
            x0 = e;
            eps = __fdiv_rn(__fsub_rn(x, __fmul_rn(sa, x0)), sb);
        } else {
            x0 = __fsub_rn(__fmul_rn(sa, x), __fmul_rn(sb, e));
            eps = __fadd_rn(__fmul_rn(sa, e), __fmul_rn(sb, x));
        }
        if (clip >= 0.f) x0 = fminf(fmaxf(x0, -clip), clip);
        // DDIM: sqrt(a_prev) x0 + dir eps.  DDPM (scheduling_ddpm.py step): pred_original_sample_coeff x0 + current_sample_coeff x,
        // the two coefficients arriving in the same table slots; the variance noise term below is shared (sigma = std, 0 at t = 0).
        float prev = p.scheduler == 1 ? __fadd_rn(__fmul_rn(sap, x0), __fmul_rn(dir, x)) : __fadd_rn(__fmul_rn(sap, x0), __fmul_rn(dir, eps));
        if (p.eta_noise) prev = __fadd_rn(prev, __fmul_rn(sigma, p.eta_noise[soff + i]));
        if (p.mask) {
            const long long bi = (long long)b * p.n + i;
            const float m = p.mask[bi], init = p.init_latents[bi];
            // last step: the un-noised init (diffusion.py:447-448); table has blend (1, 0) there
            const float noisy = last ? init : __fadd_rn(__fmul_rn(bsa, init), __fmul_rn(bsb, p.edit_noise[bi]));
            prev = __fadd_rn(__fmul_rn(noisy, m), __fmul_rn(prev, __fsub_rn(1.0f, m)));
        }
        lat[i] = prev;
        if (last && p.result) p.result[(long long)b * p.n + i] = fminf(fmaxf(__fdiv_rn(prev, p.latent_scale), 0.f), 1.f);
    }
}

// latents = src * scale ; init_copy = latents ; [editing] latents = sa * latents + sb * noise
// (diffusion.py:363-385, scheduler add_noise)
__global__ void prepare_latents_kernel(const float* __restrict__ src, float scale, const float* __restrict__ noise,
                                       float sa, float sb, float* __restrict__ latents, float* __restrict__ init_copy,
                                       long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = __fmul_rn(src[i], scale);
    if (init_copy) init_copy[i] = v;
    if (noise) v = __fadd_rn(__fmul_rn(sa, v), __fmul_rn(sb, noise[i]));
    latents[i] = v;
}

// result = clamp(latents / latent_scale, 0, 1)   (diffusion.py:470) -- used when the loop is empty
__global__ void finalize_kernel(const float* __restrict__ latents, float latent_scale, float* __restrict__ result, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) result[i] = fminf(fmaxf(__fdiv_rn(latents[i], latent_scale), 0.f), 1.f);
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }
__global__ void add_int_kernel(int* p, int v) {
    pdl_wait();
    pdl_trigger();
    *p += v;
}

}  // namespace said
