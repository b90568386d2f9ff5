// compact.cu -- compacted neighbourhood rows for the fused set-abstraction kernels (sa_fused.cu).
//
// query_ball_point pads every neighbourhood to nsample slots with copies of its first hit (pointnet_util.py:104-106).
// The copies compute exactly what the first hit computes and can never win the max-pool's first-max tie-break, so the
// shared MLP only has to run on the REAL hits: at SA1 a ball holds 5 points on average out of 32 slots.  This pass turns
// the resident ball-query result of every forward into a compact row list:
//   * a neighbourhood with c real hits takes ceil(c / 8) octets (8-row groups; the tail of the last octet repeats the
//     first hit, as the padding did);
//   * octets are packed next-fit into 32-row warp slices without straddling, so a neighbourhood always lives inside
//     one warp of a 128-row tile and the pooling epilogue can scan it segment by segment;
//   * crow_src / crow_g give every compact row its source point and its centroid, ctiles the tile count per forward,
//     cperm the gather-backward CSR permutation rewritten to compact rows (same bucket order: bit-identical sums).
// Everything downstream (ReLU bits, arg-max ranks, gradient rows) is indexed by compact row.  Results are bit-identical
// to the padded layout; the row count drops ~4x at SA1 and ~2x at SA2 on S3DIS-density blocks.
#include "psg_common.cuh"
#include "psg_internal.h"

namespace {

// real hits of every neighbourhood: the position of the first slot that repeats slot 0 (hits are distinct and ascending)
__global__ void compact_count_kernel(const int *__restrict__ ball, long long rows, int K, int *__restrict__ cnt)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int4 *r4 = reinterpret_cast<const int4 *>(ball + i * K);
    const int first = ball[i * K];
    int c = K;
    for (int q = 0; q < K / 4; ++q) {
        const int4 v = r4[q];
        const int e[4] = {v.x, v.y, v.z, v.w};
        bool done = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * q + j;
            if (k > 0 && e[j] == first && !done) { c = k; done = true; }
        }
        if (done) break;
    }
    cnt[i] = c;
}

// next-fit packing of one problem's neighbourhoods (octets of 8 rows) into 32-row slices; one thread per problem
__global__ void compact_pack_kernel(const int *__restrict__ cnt, int P, int S, int *__restrict__ slot, int *__restrict__ nsl)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int *c = cnt + (long long)p * S;
    int *o = slot + (long long)p * S;
    int slice = 0, fill = 0;
    for (int s = 0; s < S; ++s) {
        const int oct = (c[s] + 7) >> 3;
        if (fill + oct > 4) { ++slice; fill = 0; }
        o[s] = slice * 4 + fill;
        fill += oct;
    }
    nsl[p] = slice + (fill > 0 ? 1 : 0);
}

// per forward: first compact row of each of its B problems, and the forward's tile count
__global__ void compact_scan_kernel(const int *__restrict__ nsl, int T, int B, int *__restrict__ base, int *__restrict__ ctiles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int slices = 0;
    for (int b = 0; b < B; ++b) {
        base[t * B + b] = slices * 32;
        slices += nsl[t * B + b];
    }
    ctiles[t] = (slices + 3) >> 2;
}

__global__ void compact_fill_kernel(const int *__restrict__ ball, const int *__restrict__ cnt, const int *__restrict__ slot,
                                    const int *__restrict__ base, long long rows, int B, int S, int K, long long cap,
                                    int *__restrict__ crow_src, int *__restrict__ crow_g)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * K) return;
    const long long i = e / K;               // neighbourhood (p, s)
    const int k = (int)(e - i * K);
    const int c = cnt[i];
    if (k >= ((c + 7) & ~7)) return;
    const int p = (int)(i / S), s = (int)(i - (long long)p * S);
    const int t = p / B, b = p - t * B;
    const long long row = (long long)t * cap + base[p] + slot[i] * 8 + k;
    crow_src[row] = ball[i * K + (k < c ? k : 0)];
    crow_g[row] = b * S + s;
}

// CSR permutation entries (slot index s * K + k inside the problem) -> compact row inside the forward
__global__ void compact_perm_kernel(const int *__restrict__ perm, const int *__restrict__ slot, const int *__restrict__ base,
                                    long long P, int S, int K, int *__restrict__ cperm)
{
    const long long M = (long long)S * K;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P * M) return;
    const long long p = e / M;
    const int v = perm[e];
    int out = 0;
    if (v >= 0 && v < M) {
        const int s = v / K, k = v - s * K;
        out = base[p] + slot[p * S + s] * 8 + k;
    }
    cperm[e] = out;
}

inline unsigned nb(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

size_t psg_sa_compact_bytes(long long P, int T, int S, int K)
{
    const long long rows = P * S;
    auto a = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
    return a((size_t)rows * 4) * 2 + a((size_t)P * 4) * 2 + a((size_t)T * 4) + a((size_t)rows * K * 4) * 3 + 4096;
}

PsgCompact psg_sa_compact_carve(void *ws, long long P, int T, int S, int K)
{
    PsgCompact c;
    char *p = (char *)ws;
    auto take = [&](size_t bytes) { char *q = p; p += (bytes + 1023) & ~(size_t)1023; return q; };
    const long long rows = P * S;
    c.cnt = (int *)take((size_t)rows * 4);
    c.slot = (int *)take((size_t)rows * 4);
    c.nsl = (int *)take((size_t)P * 4);
    c.base = (int *)take((size_t)P * 4);
    c.ctiles = (int *)take((size_t)T * 4);
    c.crow_src = (int *)take((size_t)rows * K * 4);
    c.crow_g = (int *)take((size_t)rows * K * 4);
    c.cperm = (int *)take((size_t)rows * K * 4);
    c.cap = 0;
    return c;
}

int psg_sa_compact_build(const int *ball, const int *csr_perm, int T, int B, int S, int K, PsgCompact c, cudaStream_t st)
{
    if (!ball || !csr_perm || T <= 0 || B <= 0 || S <= 0 || (K != 16 && K != 32) || !c.cnt) return PSG_EINVAL;
    const long long P = (long long)T * B, rows = P * S;
    const long long cap = (long long)B * S * K;
    compact_count_kernel<<<nb(rows, 256), 256, 0, st>>>(ball, rows, K, c.cnt);
    PSG_LAUNCH_CHECK();
    compact_pack_kernel<<<nb(P, 32), 32, 0, st>>>(c.cnt, (int)P, S, c.slot, c.nsl);
    PSG_LAUNCH_CHECK();
    compact_scan_kernel<<<nb(T, 64), 64, 0, st>>>(c.nsl, T, B, c.base, c.ctiles);
    PSG_LAUNCH_CHECK();
    if (cudaMemsetAsync(c.crow_src, 0xFF, (size_t)T * cap * 4, st) != cudaSuccess) return PSG_ECUDA;
    if (cudaMemsetAsync(c.crow_g, 0xFF, (size_t)T * cap * 4, st) != cudaSuccess) return PSG_ECUDA;
    compact_fill_kernel<<<nb(rows * K, 256), 256, 0, st>>>(ball, c.cnt, c.slot, c.base, rows, B, S, K, cap, c.crow_src, c.crow_g);
    PSG_LAUNCH_CHECK();
    compact_perm_kernel<<<nb(rows * K, 256), 256, 0, st>>>(csr_perm, c.slot, c.base, P, S, K, c.cperm);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}
