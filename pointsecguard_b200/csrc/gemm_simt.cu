// gemm_simt.cu -- FP32 CUDA-core GEMM for the shared 1x1-conv MLPs ("exact" mode).
//
//   Out[M, Nout] = epilogue( [A1 | A2][M, K] * W[Nout, K]^T )        all operands in T-layout
//
// This is the fp32-faithful path the parity tests are pinned on (rtol 1e-3 against the fp32 CPU
// oracle); the tcgen05 kernel in gemm_tc.cu has the same interface and operand layout and replaces
// it when PSG_MLP=tf32 is selected.
//
// Tile 128 x 64 x 16, 256 threads, 8 x 4 outputs per thread, 3-stage cp.async pipeline.  Because a
// T-layout chunk plane is [128 rows][4 k] the global->shared copies are whole 16-byte pieces with
// consecutive threads on consecutive rows, and the inner product reads float4 = four k-steps of
// one row per LDS.128 (rows ty+16i -> two distinct addresses per warp, i.e. broadcast).
//
// Epilogues: bias+ReLU (forward conv+folded-BN+ReLU, pointnet_util.py:201-203, :317-319),
// bias only (conv2, pointnet2_sem_seg.py:37), ReLU-mask (dgrad: dX = (dY * W) . [Y_prev > 0]),
// none (first-layer dgrad).
#include "psg_common.cuh"
#include "psg_internal.h"

namespace {

constexpr int BM = 128, BN = 64, BKC = 4;       // BKC = 16-byte chunks per k-step (16 floats)
constexpr int STAGES = 3;
constexpr int A_STAGE = BKC * BM * 4;           // floats
constexpr int W_STAGE = BKC * BN * 4;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int EPI>
__global__ void __launch_bounds__(256, 2) gemm_simt_kernel(PsgGemmArgs g)
{
    extern __shared__ __align__(16) float smem[];
    float *As = smem;                               // [STAGES][BKC][BM][4]
    float *Ws = smem + STAGES * A_STAGE;            // [STAGES][BKC][BN][4]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int kchunks = g.k1chunks + g.k2chunks;
    const int ksteps = kchunks / BKC;

    auto load_stage = [&](int stage, int ks) {
        const int cbase = ks * BKC;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e = tid + h * 256;            // 0..511
            const int cc = e >> 7, r = e & 127;
            const int c = cbase + cc;
            const float *src = (c < g.k1chunks) ? g.A1.base + tv_off(g.A1, row0 + r, c)
                                                : g.A2.base + tv_off(g.A2, row0 + r, c - g.k1chunks);
            cp_async16(As + stage * A_STAGE + (cc * BM + r) * 4, src);
        }
        {
            const int cc = tid >> 6, n = tid & 63;
            const float *src = g.W + ((size_t)(cbase + cc) * g.Nw + n0 + n) * 4;
            cp_async16(Ws + stage * W_STAGE + (cc * BN + n) * 4, src);
        }
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ksteps) load_stage(s, s);
        cp_async_commit();
    }
    for (int ks = 0; ks < ksteps; ++ks) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = ks + STAGES - 1;
            if (nk < ksteps) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const float *a_s = As + (ks % STAGES) * A_STAGE;
        const float *w_s = Ws + (ks % STAGES) * W_STAGE;
#pragma unroll
        for (int cc = 0; cc < BKC; ++cc) {
            float4 a[8], w[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(a_s + (cc * BM + ty + 16 * i) * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4 *>(w_s + (cc * BN + tx + 16 * j) * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns rows ty+16i, columns n0 + tx + 16j
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx + 16 * j;
        if (col >= g.nout_pad) continue;
        const float bias = (EPI == PSG_EPI_BIAS_RELU || EPI == PSG_EPI_BIAS) ? g.bias[col] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long row = row0 + ty + 16 * i;
            float v = acc[i][j];
            if (EPI == PSG_EPI_BIAS_RELU) { v += bias; v = v > 0.f ? v : 0.f; }
            if (EPI == PSG_EPI_BIAS) v += bias;
            if (EPI == PSG_EPI_MASK) {
                const float y = g.Mask.base[tv_off(g.Mask, row, col >> 2) + (col & 3)];
                v = y > 0.f ? v : 0.f;
            }
            g.Out.base[tv_off(g.Out, row, col >> 2) + (col & 3)] = v;
        }
    }
}

template <int EPI>
int launch(const PsgGemmArgs &g, cudaStream_t st)
{
    const size_t smem = (size_t)STAGES * (A_STAGE + W_STAGE) * sizeof(float);   // 36 KB
    dim3 grid((unsigned)g.mtiles, (unsigned)((g.nout_pad + BN - 1) / BN));
    gemm_simt_kernel<EPI><<<grid, 256, smem, st>>>(g);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

}  // namespace

int psg_gemm_simt(const PsgGemmArgs &g, cudaStream_t st)
{
    if ((g.k1chunks + g.k2chunks) % BKC || g.k1chunks % BKC || g.mtiles <= 0 || g.nout_pad % 4) return PSG_EINVAL;
    if (g.Nw % BN || g.Nw < g.nout_pad) return PSG_EINVAL;
    switch (g.epi) {
    case PSG_EPI_BIAS_RELU: return launch<PSG_EPI_BIAS_RELU>(g, st);
    case PSG_EPI_BIAS: return launch<PSG_EPI_BIAS>(g, st);
    case PSG_EPI_MASK: return launch<PSG_EPI_MASK>(g, st);
    case PSG_EPI_NONE: return launch<PSG_EPI_NONE>(g, st);
    }
    return PSG_EINVAL;
}
