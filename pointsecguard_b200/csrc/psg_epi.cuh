// psg_epi.cuh -- epilogue helpers shared by the fused tcgen05 kernels (sa_fused.cu, chain_fused.cu).
//
// The fused kernels are instruction-issue bound in their epilogues (profiles/r1_notes.md), so the
// per-element work is kept to: one FADD (bias), one FMNMX (ReLU), two integer ops for the ReLU bit.
// ReLU bits are packed MSB-first: column i of a 32-column group is bit (31 - i) of its word, which
// is what a funnel-shift accumulation produces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// v[0..NC) <- relu(v + bias[0..NC)); returns the ReLU bits, column 0 in the MSB.  bias 16-byte aligned, in
// SHARED memory (a broadcast LDS.128 per four columns; a global load here put an L2 round trip into every
// 32-column step of every epilogue -- profiles/r1_notes.md, phase trace).
template <int NC>
__device__ __forceinline__ unsigned psg_relu_bias_bits(float *v, const float *bias)
{
    unsigned w = 0;
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) {
        const float4 b = *(reinterpret_cast<const float4 *>(bias) + q);
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

// ---- neighbourhood max-pool through a shared-memory transpose ----------------------------------------
// A warp owns 32 consecutive tile rows = 32 / K whole neighbourhoods; lane l holds NC (16 or 32) columns of
// row l in y[] (post bias + ReLU, so >= 0).  The warp writes them to its private 4 KB scratch transposed
// ([column][row], XOR-swizzled in units of four rows so that both the scalar writes and the 128-bit reads
// are bank-conflict free), then lane c scans column c over the K rows of each neighbourhood in ascending
// row order with a strict '>' -- torch.max's first-max tie-break (pointnet_util.py:205).  This replaces two
// warp REDUX per pooled element (measured ~50 cycles per column per warp) by ~1 STS + 1/4 LDS.128 + 3 ALU.
// On return lanes [0, NC) hold best[g] / arg[g] of column `lane` for neighbourhood g of the warp.
template <int K, int NC>
__device__ __forceinline__ void psg_pool_transposed(const float *y, float *scratch, int lane, float *best, int *arg)
{
#pragma unroll
    for (int c = 0; c < NC; ++c) scratch[c * 32 + (lane ^ ((c & 7) << 2))] = y[c];
    __syncwarp();
    if (lane < NC) {
        const float4 *s4 = reinterpret_cast<const float4 *>(scratch) + lane * 8;
#pragma unroll
        for (int g = 0; g < 32 / K; ++g) {
            float b = -1.f; int bi = 0;
#pragma unroll
            for (int j = g * (K / 4); j < (g + 1) * (K / 4); ++j) {
                const float4 q = s4[j ^ (lane & 7)];
                const int r0 = 4 * j - g * K;
                if (q.x > b) { b = q.x; bi = r0; }
                if (q.y > b) { b = q.y; bi = r0 + 1; }
                if (q.z > b) { b = q.z; bi = r0 + 2; }
                if (q.w > b) { b = q.w; bi = r0 + 3; }
            }
            best[g] = b; arg[g] = bi;
        }
    }
    __syncwarp();
}

// Segmented variant for COMPACTED rows (compact.cu): the warp's 32 rows are four octets; `vm` bit o says octet o holds rows,
// `sm` bit o that it starts a new neighbourhood whose centroid is gq[o]; a neighbourhood spans 1..4 consecutive octets of
// the slice.  Same transposition, the scan emits (centroid, max, rank of the first max) at every segment end.  All control
// flow depends on the warp-uniform masks only.
template <int NC, class Emit>
__device__ __forceinline__ void psg_pool_segmented(const float *y, float *scratch, int lane, const int (&gq)[4], unsigned sm,
                                                   unsigned vm, Emit emit)
{
#pragma unroll
    for (int c = 0; c < NC; ++c) scratch[c * 32 + (lane ^ ((c & 7) << 2))] = y[c];
    __syncwarp();
    if (lane < NC) {
        const float4 *s4 = reinterpret_cast<const float4 *>(scratch) + lane * 8;
        float b = -1.f;
        int bi = 0, g = -1, first = 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (!((vm >> o) & 1u)) continue;
            if ((sm >> o) & 1u) {
                if (g >= 0) emit(g, b, bi);
                g = gq[o]; b = -1.f; bi = 0; first = 8 * o;
            }
#pragma unroll
            for (int j = 2 * o; j < 2 * o + 2; ++j) {
                const float4 q = s4[j ^ (lane & 7)];
                const int r0 = 4 * j - first;
                if (q.x > b) { b = q.x; bi = r0; }
                if (q.y > b) { b = q.y; bi = r0 + 1; }
                if (q.z > b) { b = q.z; bi = r0 + 2; }
                if (q.w > b) { b = q.w; bi = r0 + 3; }
            }
        }
        if (g >= 0) emit(g, b, bi);
    }
    __syncwarp();
}

// centroids of the four octets of a warp slice + the segment masks described above, from each lane's centroid (-1: empty row)
__device__ __forceinline__ void psg_slice_segments(int g_lane, int (&gq)[4], unsigned &sm, unsigned &vm)
{
#pragma unroll
    for (int o = 0; o < 4; ++o) gq[o] = __shfl_sync(0xffffffffu, g_lane, 8 * o);
    sm = vm = 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        if (gq[o] >= 0) {
            vm |= 1u << o;
            if (o == 0 || gq[o] != gq[o - 1]) sm |= 1u << o;
        }
    }
}

// ---- arg-max scatter (backward of the neighbourhood max-pool) as a warp-cooperative refill ------------------
// dY[(g, k)][c] = (k == arg[g][c] && out[g][c] > 0) ? dOut[g][c] : 0 for `planes` 4-column chunks starting at chunk c0.
// The K rows of a neighbourhood need the SAME three rows (dOut, out, arg) and differ only in the comparison with
// k, so each lane fetches ONE chunk of its neighbourhood (coalesced, one request per array per lane) into a small
// per-warp staging area and every lane then reads the staged chunks back as shared-memory broadcasts -- instead of
// 3 global requests per chunk per lane (the scatter was ~40 % of a backward tile's latency chain).
// stage_d / stage_a: 32 float4 + 32 uchar4 per warp.  put(chunk, value) writes the A operand.
// `rank` = the row's position inside its neighbourhood (lane % K in the padded layout; with compacted rows K = 8 is the
// cooperating octet and the rank runs over the 1..4 octets of the neighbourhood).
template <int K, class Put>
__device__ __forceinline__ void psg_scatter_warp_rank(const TView &dout, const TView &outv, const unsigned char *arg, int argC,
                                                      long long g, bool valid, int lane, int rank, int c0, int planes,
                                                      float4 *stage_d, uchar4 *stage_a, Put put)
{
    const int kk = lane % K, gl = lane / K;
    for (int cb = 0; cb < planes; cb += K) {
        const int c = cb + kk;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        uchar4 am = make_uchar4(255, 255, 255, 255);
        if (valid && c < planes) {
            d = tv_ld(dout, g, c0 + c);
            const float4 o = tv_ld(outv, g, c0 + c);
            am = *reinterpret_cast<const uchar4 *>(arg + g * argC + 4 * (c0 + c));
            d.x = o.x > 0.f ? d.x : 0.f; d.y = o.y > 0.f ? d.y : 0.f; d.z = o.z > 0.f ? d.z : 0.f; d.w = o.w > 0.f ? d.w : 0.f;
        }
        stage_d[lane] = d; stage_a[lane] = am;
        __syncwarp();
        const int n = planes - cb < K ? planes - cb : K;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 dd = stage_d[gl * K + j];
            const uchar4 aa = stage_a[gl * K + j];
            float4 q;
            q.x = aa.x == rank ? dd.x : 0.f; q.y = aa.y == rank ? dd.y : 0.f;
            q.z = aa.z == rank ? dd.z : 0.f; q.w = aa.w == rank ? dd.w : 0.f;
            put(cb + j, q);
        }
        __syncwarp();
    }
}

template <int K, class Put>
__device__ __forceinline__ void psg_scatter_warp(const TView &dout, const TView &outv, const unsigned char *arg, int argC,
                                                 long long g, bool valid, int lane, int c0, int planes, float4 *stage_d,
                                                 uchar4 *stage_a, Put put)
{
    const int kk = lane % K, gl = lane / K;
    for (int cb = 0; cb < planes; cb += K) {
        const int c = cb + kk;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        uchar4 am = make_uchar4(255, 255, 255, 255);
        if (valid && c < planes) {
            d = tv_ld(dout, g, c0 + c);
            const float4 o = tv_ld(outv, g, c0 + c);
            am = *reinterpret_cast<const uchar4 *>(arg + g * argC + 4 * (c0 + c));
            d.x = o.x > 0.f ? d.x : 0.f; d.y = o.y > 0.f ? d.y : 0.f; d.z = o.z > 0.f ? d.z : 0.f; d.w = o.w > 0.f ? d.w : 0.f;
        }
        stage_d[lane] = d; stage_a[lane] = am;
        __syncwarp();
        const int n = planes - cb < K ? planes - cb : K;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 dd = stage_d[gl * K + j];
            const uchar4 aa = stage_a[gl * K + j];
            float4 q;
            q.x = aa.x == kk ? dd.x : 0.f; q.y = aa.y == kk ? dd.y : 0.f;
            q.z = aa.z == kk ? dd.z : 0.f; q.w = aa.w == kk ? dd.w : 0.f;
            put(cb + j, q);
        }
        __syncwarp();
    }
}

// ---- row-major store of a 32-row x 32-column block through a per-warp staging area -----------------------
// Thread l holds 32 columns of row l.  Written directly, every 16-byte store of a warp lands in a different
// row (32 partial sectors per instruction: measured 8500 cycles for a 128-column tile).  Staged through 4 KB
// of shared memory (XOR-swizzled 16-byte pieces, conflict-free both ways) the warp instead writes four whole
// 128-byte row segments per instruction.  dst = address of (row 0 of the warp, column 0 of the block);
// row_ok(i) tells whether warp row i may be written.
template <class RowOk>
__device__ __forceinline__ void psg_store_rm32(const float *v, float4 *stage, int lane, float *dst, long long stride, RowOk row_ok)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) stage[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int jj = lane & 7, rr = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + rr;
        const float4 q = stage[row * 8 + (jj ^ (row & 7))];
        if (row_ok(row)) *reinterpret_cast<float4 *>(dst + row * stride + 4 * jj) = q;
    }
    __syncwarp();
}
