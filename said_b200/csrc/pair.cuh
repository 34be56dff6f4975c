// The "P16" pair format: activations that feed a tensor-core contraction are stored as two fp16 planes per row,
//   row r = [hi(C) | lo(C)],  hi = fp16(x),  lo = fp16(x - hi)      (4 bytes per element, like fp32; ~22 significant bits)
// so that the consuming GEMM (gemm_h.cuh) can TMA-load ready-made operands and form x*w as hi*hi + lo*hi + hi*lo.
// The split is done ONCE, by the kernel that produces the tensor.  Range: |x| < 65504 (fp16); below |x| = 2^-3 the lo plane
// is an fp16 subnormal, so the absolute error of the pair is <= 2^-25 there and <= 2^-22 |x| above.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace said {

// hi/lo split of four fp32 values into packed fp16 pairs.  |x| must stay below 65504 (fp16 max): `amax` accumulates the
// largest magnitude seen so that the producing kernel can raise the engine's overflow flag.
SAID_DEVINL void split_pair4(const float4& x, uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01);
    hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01);
    lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}
SAID_DEVINL float amax4(float m, const float4& x) { return fmaxf(fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w))); }
constexpr float P16_LIMIT = 65000.0f;
// store four consecutive elements (columns c..c+3, c % 4 == 0) of row `row` of a pair tensor with C columns
SAID_DEVINL void store_pair4(__half* base, long long row, int C, int c, const float4& x) {
    uint2 hi, lo;
    split_pair4(x, hi, lo);
    __half* p = base + row * (2LL * C) + c;
    *reinterpret_cast<uint2*>(p) = hi;
    *reinterpret_cast<uint2*>(p + C) = lo;
}
SAID_DEVINL void store_pair4_zero(__half* base, long long row, int C, int c) {
    __half* p = base + row * (2LL * C) + c;
    *reinterpret_cast<uint2*>(p) = make_uint2(0u, 0u);
    *reinterpret_cast<uint2*>(p + C) = make_uint2(0u, 0u);
}


}  // namespace said
