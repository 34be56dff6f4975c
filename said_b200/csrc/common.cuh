// Shared device helpers for the sm_100a kernels of the SAiD inference path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace said {

#define SAID_DEVINL __device__ __forceinline__

SAID_DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// read-only 16-byte load whose L2 miss fetches the surrounding 256-byte segment (long HBM bursts for row-walking readers)
SAID_DEVINL float4 ldg4_l2pf(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor grid has completed
// and its memory operations are visible, so everything a kernel does BEFORE it (barrier init, TMEM allocation, constant
// weight prefetch) overlaps the predecessor's tail.  Every kernel of the step loop waits before it touches an activation
// buffer (read or write), which makes the ordering transitive; pdl_trigger() after the wait lets the next kernel in.
// Both are no-ops for a kernel launched without the attribute.
// MEASURED (profiles/r1_pdl.md): inside the per-step CUDA graph PDL is a net loss on this path (batch 64: 4.08 vs 4.03 ms
// per step; single clip: 0.95 vs 0.82 ms, 1.29 ms with the trigger before the wait), so it is OFF unless SAID_PDL=1.
SAID_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SAID_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Launch with an optional PDL attribute and an optional cluster width (x dimension).
template <class... KA, class... A>
inline cudaError_t launch_ex(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster_x, A&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster_x;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
SAID_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
SAID_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
SAID_DEVINL float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// SiLU as torch computes it: x * sigmoid(x) = x / (1 + exp(-x))   (openaimodel.py:155, nn.SiLU)
SAID_DEVINL float silu(float x) { return x / (1.0f + expf(-x)); }
// the same on the hardware exp2 / reciprocal units (relative error ~2e-7): the GroupNorm -> SiLU -> pair kernel is issue-bound, and
// expf + an IEEE division are two thirds of its arithmetic
SAID_DEVINL float silu_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
// exact (erf) GELU: attention.py:32 F.gelu default, transformers "gelu" activation
SAID_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// exact-GELU with erf from Abramowitz & Stegun 7.1.26 on the hardware exp2 / reciprocal units: |gelu error| <= 5e-7 (a few fp32 ulps
// at 1.0) for a third of erff's instructions.  Used by the tensor-core GEGLU epilogue, where the activation math -- not the MMA --
// was the per-tile critical path.
SAID_DEVINL float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = fmaf(-pl * t, e, 1.0f);           // erf(|x| / sqrt 2)
    return 0.5f * x + 0.5f * fabsf(x) * erf_abs;             // x * Phi(x):  0.5 x (1 + sign(x) erf(|x|/sqrt 2))
}

SAID_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SAID_DEVINL double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SAID_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace said
