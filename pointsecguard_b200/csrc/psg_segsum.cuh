// psg_segsum.cuh -- ordered segmented sum with a WARP per destination row (backward of the 3-NN interpolation and of the
// grouping gather: autograd's index_put_(accumulate) of pointnet_util.py:126-137, 305-308, made deterministic).
//
//   dst[p*R + r][c] (+)= sum over the bucket's entries e (ascending slot) of w_e * src[p*rows_per_p + slot_e/div][c]
//
// Lane l fetches entry lo + l of the bucket (slot and weight) ONCE -- one L2 round trip for up to 32 entries instead of
// one per four entries and lane -- and the warp then walks the entries in bucket order, four source rows in flight per
// lane and chunk; chunks lane, lane + 32, ... (NC of them) belong to the lane.  The additions happen in the order of
// gather.cu::segsum_kernel, so both produce the same bits.  CG: read sources and the accumulate target past L1
// (ld.global.cg), for callers whose sources were written earlier in the SAME kernel (deep.cu).
#pragma once
#include "psg_common.cuh"

struct PsgSegsumArgs {
    TView src; long long rows_per_p; int div; const float *wgt; const int *offs, *perm; int M, R; long long P;
    int nch, tail; TView dst; int acc; TView rmask; const float *rm; int rm_stride;
};

__device__ __forceinline__ float4 psg_ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }

template <int NC, bool CG>
__device__ __forceinline__ void psg_segsum_warp(const PsgSegsumArgs &p, long long wfirst, long long wstride, int lane)
{
    const long long nrows = p.P * p.R;
    auto ld4 = [](const float *q) { return CG ? psg_ldcg4(q) : *reinterpret_cast<const float4 *>(q); };
    for (long long wid = wfirst; wid < nrows; wid += wstride) {
        const long long pp = wid / p.R;
        const int r = (int)(wid % p.R);
        const int lo = __ldg(p.offs + pp * (p.R + 1) + r), hi = __ldg(p.offs + pp * (p.R + 1) + r + 1);
        const int *pm = p.perm + pp * p.M;
        const float *ww = p.wgt ? p.wgt + pp * p.M : nullptr;
        const long long sbase = pp * p.rows_per_p;
        float4 acc[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int c = lane + 32 * k;
            acc[k] = (p.acc && c < p.nch) ? ld4(p.dst.base + tv_off(p.dst, wid, c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int w0 = lo; w0 < hi; w0 += 32) {
            const int e = w0 + lane;
            const int my_slot = e < hi ? __ldg(pm + e) : -1;
            const float my_w = (ww && my_slot >= 0) ? __ldg(ww + my_slot) : 1.f;
            const int cnt = min(32, hi - w0);
            for (int j0 = 0; j0 < cnt; j0 += 4) {
                int slot[4]; float sc[4]; float4 v[4][NC];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    slot[u] = __shfl_sync(0xffffffffu, my_slot, (j0 + u) & 31);
                    sc[u] = __shfl_sync(0xffffffffu, my_w, (j0 + u) & 31);
                    if (j0 + u >= cnt) slot[u] = -1;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const int c = lane + 32 * k;
                        v[u][k] = (slot[u] < 0 || c >= p.nch) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                : p.rm ? ld4(p.rm + (sbase + slot[u] / p.div) * p.rm_stride + 4 * c)
                                       : ld4(p.src.base + tv_off(p.src, sbase + slot[u] / p.div, c));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (slot[u] < 0) continue;
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const int c = lane + 32 * k;
                        float4 q = v[u][k];
                        if (c == p.nch - 1 && p.tail) {       // keep only the first tail columns of the last chunk
                            if (p.tail < 2) q.y = 0.f;
                            if (p.tail < 3) q.z = 0.f;
                            q.w = 0.f;
                        }
                        if (ww) {
                            acc[k].x = fmaf(q.x, sc[u], acc[k].x); acc[k].y = fmaf(q.y, sc[u], acc[k].y);
                            acc[k].z = fmaf(q.z, sc[u], acc[k].z); acc[k].w = fmaf(q.w, sc[u], acc[k].w);
                        } else {
                            acc[k].x += q.x; acc[k].y += q.y; acc[k].z += q.z; acc[k].w += q.w;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int c = lane + 32 * k;
            if (c >= p.nch) continue;
            float4 o = acc[k];
            if (p.rmask.base) {                          // gradient w.r.t. the pre-activation of a ReLU layer
                const float4 y = tv_ld(p.rmask, wid, c);
                o.x = y.x > 0.f ? o.x : 0.f; o.y = y.y > 0.f ? o.y : 0.f;
                o.z = y.z > 0.f ? o.z : 0.f; o.w = y.w > 0.f ? o.w : 0.f;
            }
            tv_st(p.dst, wid, c, o);
        }
    }
}
