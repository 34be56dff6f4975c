/* said_b200 -- C ABI of the B200-native SAiD inference hot path (libsaid_sm100.so).
 *
 * Drop-in boundary: these are the entry points a host binds (ctypes in said_b200/_lib.py; see
 * INTEGRATION.md for the stub a maintainer of the reference would add) to replace the PyTorch-eager
 * implementation of `SAID.inference()` (reference said/model/diffusion.py:308-472) and the modules it
 * calls.  Plain pointers and sizes only; no C++ or torch types cross the boundary.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; said_last_error() describes the failure
 *     (thread-local string, valid until the next call on that thread);
 *   - "dev" pointers are device pointers on the engine's device, borrowed for the duration of the call;
 *     "host" pointers are ordinary host memory;
 *   - all tensors are contiguous float32, row-major, channel-last: (clip, frame, channel);
 *   - work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed as void*) and is
 *     NOT synchronised before returning, except where a function says so;
 *   - one engine per (process, device); an engine is not thread-safe; engines are independent.
 */
#ifndef SAID_B200_H
#define SAID_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SAID_API __attribute__((visibility("default")))
#else
#define SAID_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct said_engine said_engine;

/* Library / device */
SAID_API const char* said_last_error(void);
SAID_API int said_version(void);                                   /* ABI version, currently 1 */
SAID_API int said_create(int device, said_engine** out);           /* fails unless the device is sm_100 */
SAID_API void said_destroy(said_engine* e);

/* Weights.  Replaces nn.Module.load_state_dict + .to(device) for the kernels' purposes
 * (reference script/inference.py:152-159).  `name` is the reference state-dict key
 * ("denoiser.model.input_blocks.0.0.weight", "audio_encoder.encoder.layers.3.attention.q_proj.bias",
 * "null_cond_emb", "audio_proj_layer.weight", ...; both weight-norm spellings of the positional conv
 * are accepted).  The data is copied.  "time_freqs" (96 floats: exp(-ln(1e4) k/96), reference
 * said/model/ldm/util.py:75-78) must also be supplied by the host so that it is bit-identical to the
 * value the reference computes with torch.  "audio_encoder.config.do_stable_layer_norm" (one float, 0 or 1; absent = 0) is
 * the one Wav2Vec2Config switch that cannot be read off the parameter names (wav2vec2-large family: pre-LN layers).
 * said_commit_weights() validates the set, infers the rest of the configuration from the names and shapes (e.g. a LayerNorm
 * per conv layer = feat_extract_norm "layer"), repacks into the kernels' layouts and uploads. */
SAID_API int said_set_tensor(said_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim);
SAID_API int said_commit_weights(said_engine* e);
/* 1 if said_commit_weights has succeeded since the last said_set_tensor */
SAID_API int said_weights_ready(const said_engine* e);
/* configuration inferred at commit: in_channels, context dim seen by the denoiser, encoder hidden size */
SAID_API int said_get_config(const said_engine* e, int* in_channels, int* ctx_dim, int* enc_hidden);

/* Audio encoder: SAID.get_audio_embedding (reference said/model/diffusion.py:209-230 ->
 * said/model/wav2vec2.py:14-82): processed waveform (B, T_a) -> features (B, T, ctx_dim). */
SAID_API int said_encode_audio(said_engine* e, const float* wave_dev, int B, int T_a, int T, float* emb_out_dev, void* stream);

/* Device-side SAID.process_audio (reference said/model/diffusion.py:188-207 -> HF Wav2Vec2FeatureExtractor):
 * per-utterance (x - mean) / sqrt(var + 1e-7) with the population variance, for B equal-length raw clips that are already
 * on the device.  wave_dev (B, T_a) -> out_dev (B, T_a); in place allowed.   */
SAID_API int said_normalize_audio(said_engine* e, const float* wave_dev, int B, int T_a, float* out_dev, void* stream);

/* Device-side tail of load_audio (reference said/util/audio.py:35-38): torchaudio.functional.resample (polyphase windowed-sinc FIR)
 * per channel followed by the mean over channels.  wave_dev (channels, n_in); the rates are given reduced by their gcd (orig, nw);
 * bank_dev (nw, 2 * width + orig) is the filter bank, built by the host with torchaudio's published formula; out_dev (n_out),
 * n_out = ceil(nw * n_in / orig).   */
SAID_API int said_resample_mono(said_engine* e, const float* wave_dev, int channels, int n_in, int orig, int nw, int width,
                                const float* bank_dev, float* out_dev, int n_out, void* stream);

/* Hoist of everything the step loop needs from the audio features: cross-attention keys/values of all
 * four transformer blocks (reference said/model/ldm/attention.py:90-91 evaluates them on every step)
 * and, when with_uncond != 0, the constant value vector of the null-condition branch
 * (reference diffusion.py:397-400).  emb_dev: (B, T, ctx_dim). */
SAID_API int said_prepare_context(said_engine* e, const float* emb_dev, int B, int T, int with_uncond, void* stream);

/* The denoising loop (reference diffusion.py:409-470). */
typedef struct said_denoise_args {
    int B, T, n_steps;              /* clips, frames, loop iterations actually run (after strength) */
    const float* timesteps_host;    /* (n_steps) timestep of each iteration, as float */
    const float* step_table_host;   /* (n_steps, 8): sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), dir coef, sigma,
                                       clip range (<0: no clipping), blend sqrt(a_next), blend sqrt(1-a_next) */
    int prediction_type;            /* 0 epsilon, 1 sample, 2 v_prediction */
    int do_cfg;                     /* guidance_scale > 1: two denoiser branches per clip */
    float guidance_scale, guidance_rescale;
    float latent_scale;
    const float* init_src_dev;      /* (B,T,C): the randn draw (generation) or init_samples (editing) */
    float init_scale;               /* latent_scale * init_noise_sigma (diffusion.py:370) */
    const float* edit_noise_dev;    /* (B,T,C) or NULL: noise added to init_samples (diffusion.py:376-385) */
    float edit_sqrt_a, edit_sqrt_b; /* add_noise coefficients at the starting timestep */
    const float* mask_dev;          /* (B,T,C) or NULL: 1 = keep init_samples (diffusion.py:446-456) */
    const float* eta_noise_dev;     /* (n_steps,B,T,C) or NULL: variance noise for eta > 0 */
    float* intermediates_dev;       /* (n_steps,B,T,C) or NULL: latents/latent_scale before each step */
    float* result_dev;              /* (B,T,C): clamp(latents/latent_scale, 0, 1) */
    float* latents_out_dev;         /* (B,T,C) or NULL: final latents before the clamp */
    int use_graph;                  /* replay one captured CUDA graph per step */
    int scheduler;                  /* 0 DDIMScheduler.step; 1 DDPMScheduler.step (ancestral sampling, reference diffusion.py:55,404:
                                       table slots 2..4 then hold pred_original_sample_coeff, current_sample_coeff, sqrt(variance)
                                       and eta_noise_dev the per-step variance noise) */
    int resume;                     /* 1: this call continues a loop started by an earlier call (chunked loops: the per-step
                                       variance noise of eta > 0 / DDPM is then drawn and held chunk by chunk instead of for all
                                       steps): init_src_dev holds the latents the previous call returned in latents_out_dev, no
                                       scaling / noising is applied and the engine keeps its copy of the un-noised init_samples */
    int more;                       /* 1: another chunk follows: the last iteration of this call is NOT the loop's last (no final
                                       un-noised blend, no result); result_dev may be NULL */
} said_denoise_args;
SAID_API int said_denoise(said_engine* e, const said_denoise_args* args, void* stream);

/* One denoiser forward: SAID.forward / UNet1DConditionModel.forward (reference
 * said/model/diffusion.py:127-155, said/model/unet_1d_condition.py:51-77).
 * x_dev (Bp,T,C), timesteps_host (Bp) as float, ctx_dev (Bp,T_ctx,ctx_dim) -> out_dev (Bp,T,C).  T_ctx may differ from T: the
 * cross-attention then uses the reference's general alignment window (said/model/ldm/attention.py:170-189).
 * taps_dev (optional, 10 x (Bp,T,192)): outputs of the ten UNet blocks in execution order (parity tests).
 * Synchronises the stream before returning. */
SAID_API int said_denoiser_forward(said_engine* e, const float* x_dev, const float* timesteps_host, const float* ctx_dev,
                          int Bp, int T, int T_ctx, float* out_dev, float* taps_dev, void* stream);

/* Synchronises `stream` and returns (and clears) the engine's device-side status word: bit 0 = an activation reached fp16's
 * range limit (|x| >= 65000) on the fp16x3 path, whose operands are fp16 hi/lo pairs -- the results of that call are invalid;
 * rerun with precision mode 1 (3xTF32).  The Python layer calls this at the end of every inference() and raises. */
SAID_API int said_check_status(said_engine* e, void* stream, int* status_out);

/* Unit entry points used by the parity tests (each is one kernel of the path). */
SAID_API int said_op_ddim_step(said_engine* e, const float* pred_dev, float* latents_dev, int B, int n, int do_cfg,
                      float guidance_scale, float guidance_rescale, int prediction_type, const float* row8_host,
                      const float* eta_noise_dev, int scheduler, void* stream);
SAID_API int said_op_self_attention(said_engine* e, const float* qkv_dev, int B, int T, int heads, int head_dim,
                           float* out_dev, void* stream);

/* Unit entry point of the fp16x3 GEMM (gemm_h.cuh): out (M,N) = sum over taps of A[m + tap - (taps-1)/2, :] . Wt[tap*Cin:(tap+1)*Cin, :]
 * + bias, rows outside [0, M) read as zero (Conv1d padding); a_dev (M,Cin) fp32 is converted to the fp16 hi/lo pair format on
 * the device, wt_host (taps*Cin, N) is the K-major weight matrix, packed and uploaded by the call.  taps 1 or 3, Cin % 64 == 0,
 * N = 32 or a multiple of 192.  Synchronises. */
SAID_API int said_op_gemm_h(said_engine* e, const float* a_dev, int M, int Cin, int taps, const float* wt_host, int N,
                            const float* bias_dev, float* out_dev, void* stream);

/* Unit entry point of the fused feed-forward kernel (ffn_h.cuh; reference said/model/ldm/attention.py:25-51 FeedForward with GEGLU,
 * :232-234 proj_out + residual, as folded at load): out (M,192) = [geglu(ln . W1 + b1) | x2] . W2 + b2 + res, where
 * geglu(v | g) = v * gelu(g).  ln_dev, x2_dev, res_dev (optional) (M,192) fp32 on the device; w1_host (192, 1536) K-major with
 * value / gate columns interleaved (column 2j = value j, 2j+1 = gate j), b1_dev (1536) interleaved alike; w2_host (960, 192)
 * K-major (rows 0..767 multiply the GEGLU output, rows 768..959 multiply x2); b2_dev (192) or NULL.  Synchronises. */
SAID_API int said_op_ffn_h(said_engine* e, const float* ln_dev, const float* x2_dev, const float* res_dev, int M,
                           const float* w1_host, const float* b1_dev, const float* w2_host, const float* b2_dev, float* out_dev,
                           void* stream);

/* Diagnostics: average milliseconds of the fp16x3 GEMM (M rows, K = taps * Cin, N a multiple of 192) over `iters` launches on
 * zero-filled scratch operands; dbg bits disable parts of the kernel (1 activation TMA loads, 2 weight copies, 4 epilogue I/O,
 * 8 MMAs) to attribute time.  Synchronises. */
SAID_API int said_op_gemm_h_bench(said_engine* e, int M, int Cin, int taps, int N, int with_residual, int dbg, int iters, float* ms_out);

/* The tcgen05 (3xTF32) self-attention kernel: head_dim 32, T <= 304 (longer sequences use the FFMA kernel). */
SAID_API int said_op_self_attention_tc(said_engine* e, const float* qkv_dev, int B, int T, int heads, float* out_dev, void* stream);

/* The fp16x3 self-attention kernel (attention_h.cuh): tcgen05 kind::f16 over fp16 hi/lo pairs, online softmax with the
 * probabilities kept in tensor memory, head_dim 32, T <= 512. */
SAID_API int said_op_self_attention_h(said_engine* e, const float* qkv_dev, int B, int T, int heads, float* out_dev, void* stream);

/* Distribution-level evaluation on the device (reference said/model/vae.py:26-89, script/test_evaluate.py:53-106,
 * said/metric/frechet_distance.py:17-64): the BCVAE encoder (eval mode) over sliding 120-frame windows of coefficient sequences,
 * and the Frechet distance between two sets of 64-d latents.  Weights: said_set_tensor with the reference's BCVAE state-dict keys
 * prefixed "bcvae." ("bcvae.encoder.conv_layers.0.weight", ..., running_mean / running_var included), then said_eval_commit_bcvae.
 *   said_eval_bcvae_latents: coeffs_dev (B,T,32) -> latents_out_dev (B * nw, 64), nw = (T - 120) / step + 1, window-major per clip.
 *   said_eval_frechet: out4_host = {distance, |mu1 - mu2|^2, tr S1 + tr S2, tr (S1 S2)^(1/2)}; synchronises. */
SAID_API int said_eval_commit_bcvae(said_engine* e);
SAID_API int said_eval_bcvae_latents(said_engine* e, const float* coeffs_dev, int B, int T, int step, float* latents_out_dev, void* stream);
SAID_API int said_eval_frechet(said_engine* e, const float* lat1_dev, int n1, const float* lat2_dev, int n2, double* out4_host, void* stream);

/* Number of kernels this engine has launched (graph replays counted node by node). */
SAID_API long long said_launch_count(const said_engine* e);
/* Number of times said_denoise had to capture + instantiate its per-step CUDA graph (the instantiated graph is cached and
 * replayed by later calls whose shapes, scalars, per-step user tensors and workspaces are unchanged). */
SAID_API long long said_graph_captures(const said_engine* e);

/* Diagnostics: average milliseconds of the tcgen05 GEMM (N = 192, plain loader, M x K activations) over
 * `iters` launches on scratch buffers; dbg bits disable parts of the kernel (2 weight copies,
 * 4 epilogue I/O, 8 MMAs) to attribute time.  Synchronises. */
SAID_API int said_op_gemm_tc_bench(said_engine* e, int M, int K, int nsplit, int with_residual, int dbg, int iters, float* ms_out);

/* Contraction precision of the denoiser's Linear / Conv1d layers:
 *   0  IEEE fp32 FFMA (CUDA cores);
 *   1  tcgen05 tensor cores, 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulate): fp32-level accuracy [default];
 *   2  tcgen05 tensor cores, single TF32 pass (what the reference gets from cuDNN for its convs on a GPU);
 *   3  tcgen05 tensor cores, fp16 hi/lo operand pairs, three passes (hi*hi + lo*hi + hi*lo, fp32 accumulate), operands
 *      pre-split by their producers and loaded by TMA: the accuracy of mode 1 at twice its tensor-core rate; activations
 *      must stay below fp16's range (65504), checked on the device (said_check_status).
 * GEMMs with fewer rows than a threshold stay on the FFMA kernel, so results are bit-reproducible across batch sizes only within
 * one regime (mode 0 is reproducible across all sizes).  Defaults: 512 rows for the mode-3 denoiser (one 5 s clip under
 * guidance is 602 rows: its 32-column weight tile images keep the tensor-core path ahead of the FFMA kernels from there), 2048
 * rows for modes 1 / 2 and for the encoder.  tc_min_rows > 0 sets every threshold to that value, 0 keeps the current ones,
 * < 0 restores the defaults.
 * encoder_mode: the same choice for the Wav2Vec2 encoder's GEMMs (default 0: the encoder runs once per clip, and
 * its long contractions (K up to 3072) lose about a decimal digit under the tensor cores' truncating accumulation). */
SAID_API int said_set_precision(said_engine* e, int mode, int tc_min_rows, int encoder_mode);

/* Per-kernel-family timing for bench.py's roofline: between begin and end every launch is followed by a
 * CUDA event on its stream (run said_denoise with use_graph = 0 in between).  said_profile_end
 * synchronises the device and returns, per family, the summed device time (ms) and the launch count.
 * Families: 0 conv-loader GEMM, 1 LayerNorm-loader GEMM, 2 plain GEMM, 3 self-attention,
 * 4 aligned cross-attention, 5 GroupNorm statistics, 6 CFG + scheduler step, 7 other. */
#define SAID_PROFILE_FAMILIES 8
SAID_API int said_profile_begin(said_engine* e);
SAID_API int said_profile_end(said_engine* e, double* ms_out, long long* count_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* SAID_B200_H */
