// psg_epi.cuh -- epilogue helpers shared by the fused tcgen05 kernels (sa_fused.cu, chain_fused.cu).
//
// The fused kernels are instruction-issue bound in their epilogues (profiles/r1_notes.md), so the
// per-element work is kept to: one FADD (bias), one FMNMX (ReLU), two integer ops for the ReLU bit.
// ReLU bits are packed MSB-first: column i of a 32-column group is bit (31 - i) of its word, which
// is what a funnel-shift accumulation produces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// v[0..NC) <- relu(v + bias[0..NC)); returns the ReLU bits, column 0 in the MSB.  bias 16-byte aligned.
template <int NC>
__device__ __forceinline__ unsigned psg_relu_bias_bits(float *v, const float *__restrict__ bias)
{
    unsigned w = 0;
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + q);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float y = fmaxf(v[4 * q + j] + bb[j], 0.f);
            v[4 * q + j] = y;
            // y >= 0: (0 - bits(y)) is negative exactly when y > 0; its sign bit is shifted into w
            w = __funnelshift_l((unsigned)(0 - (int)__float_as_uint(y)), w, 1);
        }
    }
    if (NC < 32) w <<= (32 - NC);
    return w;
}

// v[i] <- bit(i) ? v[i] : 0 for the NC columns whose bits start at the MSB of w
template <int NC>
__device__ __forceinline__ void psg_apply_bits(float *v, unsigned w)
{
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = ((int)(w << i) < 0) ? v[i] : 0.f;
}
