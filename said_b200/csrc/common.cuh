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
SAID_DEVINL float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
SAID_DEVINL void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
SAID_DEVINL float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// SiLU as torch computes it: x * sigmoid(x) = x / (1 + exp(-x))   (openaimodel.py:155, nn.SiLU)
SAID_DEVINL float silu(float x) { return x / (1.0f + expf(-x)); }
// exact (erf) GELU: attention.py:32 F.gelu default, transformers "gelu" activation
SAID_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

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
