// psg_loss.cuh -- per-point loss gradients w.r.t. the logits z (the network's log_softmax head is
// folded in: d/dz of a function of log_softmax(z)).  Shared by the stand-alone kernels in
// elementwise.cu and the fused head of fp_head_fused.cu so both paths have one arithmetic.
// Reference: nontarget.py:26,34 / target.py:27,38 (cross-entropy), nontarget.py:120-128 and
// target.py:149-168 (C&W f).
#pragma once
#include <cuda_runtime.h>

constexpr int kPsgMaxCls = 16;

// Every loop below runs over the compile-time bound kPsgMaxCls with a `c < ncls` predicate and no array is
// ever indexed by a run-time value: the rows stay in registers.  (A local-memory array costs an L2 round
// trip per access in the fused kernels, whose shared-memory carve-out leaves next to no L1.)

// p = softmax(v[0..ncls)); mx / lse such that log_softmax = (v - mx) - lse
__device__ __forceinline__ void psg_softmax_row(const float *v, int ncls, float *p, float &mx, float &lse)
{
    mx = v[0];
#pragma unroll
    for (int c = 1; c < kPsgMaxCls; ++c) if (c < ncls) mx = fmaxf(mx, v[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) { p[c] = c < ncls ? expf(v[c] - mx) : 0.f; if (c < ncls) s += p[c]; }
    lse = logf(s);
    const float inv = 1.0f / s;
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) p[c] *= inv;
}

// generic upstream gradient on the log-probabilities: dz = dlogp - softmax * sum(dlogp)
__device__ __forceinline__ void psg_dz_generic(const float *v, int ncls, const float *__restrict__ dlogp_row, float *o)
{
    float p[kPsgMaxCls], mx, lse, d[kPsgMaxCls];
    psg_softmax_row(v, ncls, p, mx, lse);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) { d[c] = c < ncls ? dlogp_row[c] : 0.f; if (c < ncls) s += d[c]; }
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) o[c] = c < ncls ? d[c] - p[c] * s : 0.f;
}

// cross-entropy on the log-probabilities (log_softmax is idempotent): dz = (softmax - onehot) * scale
__device__ __forceinline__ void psg_dz_ce_row(const float *v, int ncls, int y, float scale, float *o)
{
    float p[kPsgMaxCls], mx, lse;
    psg_softmax_row(v, ncls, p, mx, lse);
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) o[c] = c < ncls ? (p[c] - (c == y ? 1.f : 0.f)) * scale : 0.f;
}

// C&W f = clamp(sign * (p_y - max_{c != y} p_c), min = -kappa); returns f, sets hit = (argmax z == y)
__device__ __forceinline__ float psg_dz_cw_row(const float *v, int ncls, int y, float kappa, float sgn, float *o, int &hit)
{
    float p[kPsgMaxCls], mx, lse;
    psg_softmax_row(v, ncls, p, mx, lse);
    int oc = -1; float other = 0.f, py = 0.f;             // (1 - onehot) * p has a 0 at the label
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) {
        if (c < ncls && c != y && p[c] > other) { other = p[c]; oc = c; }
        if (c == y) py = p[c];
    }
    const float val = sgn * (py - other);
    const bool pass = val >= -kappa;
    int best = 0; float bestv = v[0];                     // outputs.max(dim=2)[1]: first arg-max
#pragma unroll
    for (int c = 1; c < kPsgMaxCls; ++c) if (c < ncls && v[c] > bestv) { bestv = v[c]; best = c; }
    hit = best == y ? 1 : 0;
    // g = df/dp ; dz_c = p_c * (g_c - sum_k g_k p_k); `other` is p[oc] whenever oc >= 0
    const float gy = pass ? sgn : 0.f, go = (pass && oc >= 0) ? -sgn : 0.f;
    const float dot = gy * py + (oc >= 0 ? go * other : 0.f);
#pragma unroll
    for (int c = 0; c < kPsgMaxCls; ++c) {
        const float gc = c == y ? gy : (c == oc ? go : 0.f);
        o[c] = c < ncls ? p[c] * (gc - dot) : 0.f;
    }
    return pass ? val : -kappa;
}
