// wgrad_tc.cu -- weight gradient of a 1x1 conv on the tensor cores: dW[n][k] = sum_r dZ[r][n] * X[r][k]  (TF32, fp32 accumulate)
//
// Reference: autograd of the nn.Conv1d / nn.Conv2d(1x1) layers in PointNet/models/pointnet_util.py:200-203, :317-319 during
// train_semseg.py:174 (loss.backward()).  The dgrad GEMMs contract over channels and read T-layout tiles as K-MAJOR operands;
// the weight gradient contracts over ROWS, and the very same tiles are then MN-MAJOR operands: a chunk plane [128 rows][4
// channels] is sixteen 8-row x 16-byte core matrices stacked along K (rows), the next four channels are the next plane --
// the canonical no-swizzle MN-major layout with LBO = 128 B (next eight rows) and SBO = 2048 B (next plane).  So no transpose
// is ever materialised: one bulk copy brings the dZ planes of a 128-row tile (A: M = output channels), one the X planes
// (B: N = input channels), sixteen tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8 rows each) with both operands flagged
// MN-major accumulate into TMEM over all the row tiles of the CTA's split, and the partial tile is written for the ordered
// split reduction of train.cu.  Two-stage bulk-copy ring, one producer thread, one MMA thread, four epilogue warps.
#include "psg_common.cuh"
#include "psg_internal.h"
#include "psg_tc.cuh"

namespace {

constexpr int kBN = 64;                        // input channels per CTA (MMA N)
constexpr int kAStage = 128 * 512;             // 128 output channels x 128 rows x 4 B
constexpr int kBStage = kBN * 512;
constexpr int kStageBytes = kAStage + kBStage; // 96 KB
constexpr int kStages = 2;
constexpr int kSmem = kStages * kStageBytes + 1024;

struct WgradTcArgs {
    TView dz, a1, a2;
    int dz_planes;             // planes (4 output channels each) the dz tensor really has
    int k1chunks, k2chunks;    // input planes of the two sources
    int ntiles, tiles_per_split;
    float *partial;            // [splits][npad][kpad]
    int npad, kpad;
};

// instruction descriptor: D = F32, A = B = TF32, both operands MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t idesc_tf32_mn(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(192) wgrad_tc_kernel(WgradTcArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_full[kStages], bar_empty[kStages], bar_acc;
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *base = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, mt = blockIdx.y, nt = blockIdx.z;
    const int m_plane0 = mt * 32, n_plane0 = nt * (kBN / 4);
    const int a_planes = min(32, p.dz_planes - m_plane0);                   // planes that exist (the rest of the A stage stays zero)
    const int kplanes = p.k1chunks + p.k2chunks;
    const int b_planes = min(kBN / 4, kplanes - n_plane0);
    const int t0 = split * p.tiles_per_split, t1 = min(p.ntiles, t0 + p.tiles_per_split);

    // zero both stages once: planes past the tensors' widths are never written by the copies
    for (int e = threadIdx.x; e < kStages * kStageBytes / 16; e += blockDim.x) reinterpret_cast<float4 *>(base)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::mbar_init(tc::smem_u32(&bar_acc), 1);
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 64);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- producer ----
            // the N tile may straddle the two sources of a feature-propagation concat: planes [n_plane0, n_plane0 + b_planes)
            const int b1 = max(0, min(b_planes, p.k1chunks - n_plane0));            // planes taken from a1
            const int b2 = b_planes - b1;                                           // ... from a2
            const uint32_t bytes = (uint32_t)(a_planes * 2048 + b_planes * 2048);
            int it = 0;
            for (int t = t0; t < t1; ++t, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                const uint32_t full = tc::smem_u32(&bar_full[s]);
                if (it >= kStages) tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ph ^ 1u);
                tc::mbar_expect_tx(full, bytes);
                const uint32_t sA = sbase + s * kStageBytes, sB = sA + kAStage;
                const long long row0 = (long long)t * 128;
                tc::bulk_g2s(sA, p.dz.base + tv_off(p.dz, row0, m_plane0), (uint32_t)(a_planes * 2048), full);
                if (b1 > 0) tc::bulk_g2s(sB, p.a1.base + tv_off(p.a1, row0, n_plane0), (uint32_t)(b1 * 2048), full);
                if (b2 > 0) tc::bulk_g2s(sB + b1 * 2048, p.a2.base + tv_off(p.a2, row0, n_plane0 + b1 - p.k1chunks), (uint32_t)(b2 * 2048), full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer: 16 K-steps of eight rows per 128-row tile ----
            const uint32_t idesc = idesc_tf32_mn(128, kBN);
            uint32_t acc = 0;
            int it = 0;
            for (int t = t0; t < t1; ++t, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), ph);
                tc::fence_after_sync();
                const uint32_t sA = sbase + s * kStageBytes, sB = sA + kAStage;
#pragma unroll 4
                for (int kk = 0; kk < 16; ++kk) {
                    const uint64_t ad = tc::smem_desc(sA + kk * 128, 128, 2048);
                    const uint64_t bd = tc::smem_desc(sB + kk * 128, 128, 2048);
                    tc::mma_tf32(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
                tc::mma_commit(tc::smem_u32(&bar_empty[s]));
            }
            tc::mma_commit(tc::smem_u32(&bar_acc));
        }
    } else {
        // ---- epilogue: lane = output channel, columns = input channels of the tile ----
        const int q = warp & 3;
        tc::mbar_wait(tc::smem_u32(&bar_acc), 0);
        tc::fence_after_sync();
        const int n = mt * 128 + q * 32 + lane;
        float *o = p.partial + ((size_t)split * p.npad + n) * p.kpad + nt * kBN;
        for (int c16 = 0; c16 < kBN; c16 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c16, v);
            if (n < p.npad && t1 > t0) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (nt * kBN + c16 + 4 * c < p.kpad)
                        *reinterpret_cast<float4 *>(o + c16 + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 64);
}

}  // namespace

// partial[split][npad][kpad] (npad = cout rounded to 128, kpad = cin rounded to 64); returns the split count through *splits_out
int psg_wgrad_tc(TView dz, int dz_wchunks, int cout, TView a1, int k1, TView a2, int k2, long long rows, float *partial,
                 int npad, int kpad, int *splits_out, cudaStream_t st)
{
    if (rows <= 0 || rows % 128 || k1 % 4 || k2 % 4 || npad % 128 || kpad % kBN) return PSG_EUNSUPPORTED;
    static PsgDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return PSG_ECUDA;
        attr_once.mark();
    }
    WgradTcArgs p;
    p.dz = dz; p.a1 = a1; p.a2 = a2;
    p.dz_planes = dz_wchunks;
    p.k1chunks = k1 / 4; p.k2chunks = k2 / 4;
    p.ntiles = (int)(rows / 128);
    p.npad = npad; p.kpad = kpad;
    p.partial = partial;
    const int mt = npad / 128, nt = kpad / kBN;
    int splits = (296 + mt * nt - 1) / (mt * nt);
    if (splits > p.ntiles) splits = p.ntiles;
    if (splits > 600) splits = 600;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (p.ntiles + splits - 1) / splits;
    splits = (p.ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
    wgrad_tc_kernel<<<dim3(splits, mt, nt), 192, kSmem, st>>>(p);
    PSG_LAUNCH_CHECK();
    *splits_out = splits;
    return PSG_OK;
}
