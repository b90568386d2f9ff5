// geomgrad.cu -- the geometric part of the input gradient: d cost / d xyz through the centred
// neighbour coordinates of every set-abstraction level and through the inverse-distance weights of
// every feature-propagation level.  (The feature part -- xyz also enters the network as input
// channels 0:3 -- comes out of the ordinary feature backward.)
//
// Reference: autograd of PointNet/models/pointnet_util.py:126-132 (new_xyz = index_points(xyz, fps_idx);
// grouped_xyz_norm = index_points(xyz, idx) - new_xyz) and :301-308 (dists = square_distance; the three
// smallest; weight = (1/(d+1e-8)) / sum; interpolated = sum(index_points(points2, idx) * weight)).
// Sampling / grouping / neighbour indices are integers and carry no gradient, as in the reference.
//
// The colour attacks never need this (SURVEY.md finding 1); it exists for attack fields that include
// coordinates (BASELINE.json configs[2]) and makes get_model's autograd complete on all 9 channels.
// Everything is a deterministic ordered reduction (no float atomics).
#include "psg_common.cuh"
#include "psg_internal.h"

namespace {

inline unsigned nb(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

__device__ __forceinline__ float tl_get(const TView &v, long long row, int col)
{
    return v.base[tv_off(v, row, col >> 2) + (col & 3)];
}

// source side of a set-abstraction level: dxyz_src[p][r] += sum over CSR entries of dG[row][D..D+2]
__global__ void sa_xyz_src_kernel(TView dG, int D, long long rows_per_p, const int *__restrict__ offs,
                                  const int *__restrict__ perm, int M, int R, long long P, float *__restrict__ dxyz_src)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * R) return;
    const long long p = t / R;
    const int r = (int)(t % R);
    const int lo = offs[p * (R + 1) + r], hi = offs[p * (R + 1) + r + 1];
    const int *pm = perm + p * M;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int e = lo; e < hi; ++e) {
        const long long row = p * rows_per_p + pm[e];
        ax += tl_get(dG, row, D); ay += tl_get(dG, row, D + 1); az += tl_get(dG, row, D + 2);
    }
    float *o = dxyz_src + t * 3;
    o[0] += ax; o[1] += ay; o[2] += az;
}

// centre side: dxyz_ctr[p][s] -= sum_k dG[(p,s,k)][D..D+2]   (padded slots hold exact zeros)
__global__ void sa_xyz_ctr_kernel(TView dG, int D, int K, long long groups, float *__restrict__ dxyz_ctr)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int k = 0; k < K; ++k) {
        const long long row = g * K + k;
        ax += tl_get(dG, row, D); ay += tl_get(dG, row, D + 1); az += tl_get(dG, row, D + 2);
    }
    float *o = dxyz_ctr + g * 3;
    o[0] -= ax; o[1] -= ay; o[2] -= az;
}

// xyz_l = xyz_{l-1}[fps_idx]: dxyz_{l-1}[p][fps_idx[s]] += dxyz_l[p][s].  FPS indices repeat only on degenerate clouds
// (every remaining point coincides with a selected one), but then the order of the additions matters for exactness.
// One CTA per problem: pass 1 finds, per source point, the FIRST centroid that selected it and how many did; pass 2
// lets that first centroid's thread add all contributions of its source point in ascending centroid order (a single
// addition in the non-degenerate case).  Same result as a serial walk over the centroids, which is what this replaced
// (one thread per problem, 1024 dependent read-modify-writes: 1.3 ms of the coordinate attack's 3.9 ms step).
__global__ void __launch_bounds__(1024)
fps_xyz_back_kernel(const float *__restrict__ dxyz_l, const int *__restrict__ fps_idx, int S, int R, float *__restrict__ dxyz_src)
{
    extern __shared__ int sh[];
    int *first = sh, *count = sh + R;
    const int p = blockIdx.x;
    for (int r = threadIdx.x; r < R; r += blockDim.x) { first[r] = 0x7fffffff; count[r] = 0; }
    __syncthreads();
    const int *idx = fps_idx + (long long)p * S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int r = idx[s];
        atomicMin(first + r, s);
        atomicAdd(count + r, 1);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int r = idx[s];
        if (first[r] != s) continue;
        float *o = dxyz_src + ((long long)p * R + r) * 3;
        const float *g = dxyz_l + ((long long)p * S + s) * 3;
        float ox = o[0] + g[0], oy = o[1] + g[1], oz = o[2] + g[2];
        if (count[r] > 1) {
            for (int s2 = s + 1; s2 < S; ++s2) {
                if (idx[s2] != r) continue;
                const float *g2 = dxyz_l + ((long long)p * S + s2) * 3;
                ox += g2[0]; oy += g2[1]; oz += g2[2];
            }
        }
        o[0] = ox; o[1] = oy; o[2] = oz;
    }
}

// fallback for source clouds whose two per-point tables do not fit in shared memory: the serial walk
__global__ void fps_xyz_back_serial_kernel(const float *__restrict__ dxyz_l, const int *__restrict__ fps_idx, int S, int R, int P,
                                           float *__restrict__ dxyz_src)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    for (int s = 0; s < S; ++s) {
        const int r = fps_idx[(long long)p * S + s];
        const float *g = dxyz_l + ((long long)p * S + s) * 3;
        float *o = dxyz_src + ((long long)p * R + r) * 3;
        o[0] += g[0]; o[1] += g[1]; o[2] += g[2];
    }
}

// feature-propagation level: one warp per fine point.
//   dw_j = <dI[n], f[idx_j]>;  w = r / sum(r), r_j = 1 / (d_j + 1e-8)
//   dr_j = (dw_j - sum_k w_k dw_k) / sum(r);  dd_j = -r_j^2 dr_j
//   d_j = -2 x1.x2_j + |x1|^2 + |x2_j|^2  =>  dd_j/dx1 = 2 (x1 - x2_j),  dd_j/dx2_j = 2 (x2_j - x1)
// Fine-side gradient is row-local; the coarse-side vectors go to `ctmp` [rows*3][3] for the CSR reduce.
__global__ void __launch_bounds__(256)
fp_xyz_kernel(TView dI, TView coarse, int C2, const float *__restrict__ xyz1, long long stride1, int nclouds1,
              const float *__restrict__ xyz2, int S, const int *__restrict__ nn_idx, long long rows, int N,
              float *__restrict__ dxyz1, float *__restrict__ ctmp)
{
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= rows) return;
    const long long p = wid / N;
    const int n = (int)(wid % N);
    const int *ii = nn_idx + wid * 3;
    const int i0 = ii[0], i1 = ii[1], i2 = ii[2];
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    for (int c = lane; c < C2 / 4; c += 32) {
        const float4 g = tv_ld(dI, wid, c);
        const float4 a = tv_ld(coarse, p * S + i0, c), b = tv_ld(coarse, p * S + i1, c), e = tv_ld(coarse, p * S + i2, c);
        d0 += g.x * a.x + g.y * a.y + g.z * a.z + g.w * a.w;
        d1 += g.x * b.x + g.y * b.y + g.z * b.z + g.w * b.w;
        d2 += g.x * e.x + g.y * e.y + g.z * e.z + g.w * e.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if (lane != 0) return;
    const float *q = xyz1 + (long long)(p % nclouds1) * stride1 + (long long)n * 3;
    const float qx = q[0], qy = q[1], qz = q[2], qn = psg_sqnorm(qx, qy, qz);
    const float *c2 = xyz2 + p * S * 3;
    const int idx[3] = {i0, i1, i2};
    const float dw[3] = {d0, d1, d2};
    float cx[3], cy[3], cz[3], r[3];
    float rs = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        cx[j] = c2[idx[j] * 3]; cy[j] = c2[idx[j] * 3 + 1]; cz[j] = c2[idx[j] * 3 + 2];
        const float d = psg_sqdist(qx, qy, qz, qn, cx[j], cy[j], cz[j], psg_sqnorm(cx[j], cy[j], cz[j]));
        r[j] = 1.0f / (d + 1e-8f);
        rs += r[j];
    }
    float wd = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) wd += (r[j] / rs) * dw[j];
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float dr = (dw[j] - wd) / rs;
        const float dd = -r[j] * r[j] * dr;
        const float vx = 2.f * (qx - cx[j]) * dd, vy = 2.f * (qy - cy[j]) * dd, vz = 2.f * (qz - cz[j]) * dd;
        gx += vx; gy += vy; gz += vz;
        float *o = ctmp + (wid * 3 + j) * 3;
        o[0] = -vx; o[1] = -vy; o[2] = -vz;
    }
    float *o = dxyz1 + wid * 3;
    o[0] += gx; o[1] += gy; o[2] += gz;
}

// dxyz2[p][r] += sum over CSR entries (ascending slot) of ctmp[p*M + slot]
__global__ void fp_xyz_coarse_kernel(const float *__restrict__ ctmp, const int *__restrict__ offs, const int *__restrict__ perm,
                                     int M, int R, long long P, float *__restrict__ dxyz2)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P * R) return;
    const long long p = t / R;
    const int r = (int)(t % R);
    const int lo = offs[p * (R + 1) + r], hi = offs[p * (R + 1) + r + 1];
    const int *pm = perm + p * M;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int e = lo; e < hi; ++e) {
        const float *v = ctmp + (p * M + pm[e]) * 3;
        ax += v[0]; ay += v[1]; az += v[2];
    }
    float *o = dxyz2 + t * 3;
    o[0] += ax; o[1] += ay; o[2] += az;
}

// dfeat0[row][0..2] += dxyz0[row]   (xyz is also input channels 0:3)
__global__ void add_xyz_to_feat_kernel(const float *__restrict__ dxyz0, long long rows, TView dfeat0)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float4 v = tv_ld(dfeat0, row, 0);
    v.x += dxyz0[row * 3]; v.y += dxyz0[row * 3 + 1]; v.z += dxyz0[row * 3 + 2];
    tv_st(dfeat0, row, 0, v);
}

}  // namespace

int psg_sa_xyz_backward(TView dG, int D, int K, long long groups_per_p, long long P, const int *offs, const int *perm, int R,
                        float *dxyz_src, float *dxyz_ctr, cudaStream_t st)
{
    const long long rows_per_p = groups_per_p * K;
    sa_xyz_src_kernel<<<nb(P * R, 256), 256, 0, st>>>(dG, D, rows_per_p, offs, perm, (int)rows_per_p, R, P, dxyz_src);
    sa_xyz_ctr_kernel<<<nb(P * groups_per_p, 256), 256, 0, st>>>(dG, D, K, P * groups_per_p, dxyz_ctr);
    PSG_LAUNCH_CHECK();
    ++g_psg_launch_count;
    return PSG_OK;
}

int psg_fps_xyz_backward(const float *dxyz_l, const int *fps_idx, int S, int R, int P, float *dxyz_src, cudaStream_t st)
{
    const size_t smem = (size_t)2 * R * sizeof(int);
    if (smem <= 200 * 1024) {
        static PsgDeviceOnce attr_once;
        if (attr_once.need()) {
            if (cudaFuncSetAttribute(fps_xyz_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
                return PSG_ECUDA;
            attr_once.mark();
        }
        fps_xyz_back_kernel<<<P, S < 1024 ? ((S + 31) / 32) * 32 : 1024, smem, st>>>(dxyz_l, fps_idx, S, R, dxyz_src);
    } else {
        fps_xyz_back_serial_kernel<<<nb(P, 64), 64, 0, st>>>(dxyz_l, fps_idx, S, R, P, dxyz_src);
    }
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_fp_xyz_backward(TView dI, TView coarse, int C2, const float *xyz1, long long stride1, int nclouds1, const float *xyz2,
                        int S, const int *nn_idx, const int *offs, const int *perm, long long P, int N, float *dxyz1,
                        float *dxyz2, float *ctmp, cudaStream_t st)
{
    const long long rows = P * N;
    fp_xyz_kernel<<<nb(rows * 32, 256), 256, 0, st>>>(dI, coarse, C2, xyz1, stride1, nclouds1, xyz2, S, nn_idx, rows, N, dxyz1,
                                                      ctmp);
    fp_xyz_coarse_kernel<<<nb(P * S, 256), 256, 0, st>>>(ctmp, offs, perm, N * 3, S, P, dxyz2);
    PSG_LAUNCH_CHECK();
    ++g_psg_launch_count;
    return PSG_OK;
}

int psg_add_xyz_to_feat(const float *dxyz0, long long rows, TView dfeat0, cudaStream_t st)
{
    add_xyz_to_feat_kernel<<<nb(rows, 256), 256, 0, st>>>(dxyz0, rows, dfeat0);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}
