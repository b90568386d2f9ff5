// deep.cu -- the deep levels (FP4, FP3 and the interpolation backwards around them) as ONE persistent kernel per direction.
//
// Reference: PointNet/models/pointnet_util.py:282-320 (PointNetFeaturePropagation: 3-NN interpolation, concat, conv + BN +
// ReLU chain) for fp4 / fp3 of pointnet2_sem_seg.py:36-37, and the autograd of the same (dgrad GEMMs, index_put_(accumulate)
// of the interpolation).
//
// At B = 16 these levels hold 1024 (fp4) and 4096 (fp3) rows: four 128 x N GEMMs of a few hundred MFLOP, two interpolations
// and three segmented sums per direction, each a launch of 5-18 us whose work is 1-3 us (profiles/r2_launches_summary.csv:
// 19 % of the attack step for 3 % of its FLOPs).  What they cost is the dependent-launch chain, not the math.  Here the
// chain is a PHASE LIST executed by one grid of persistent CTAs (one per SM) with a grid barrier between phases:
//   * GEMM phase: the body of gemm_tc.cu (bulk-copy operand ring -> tcgen05.mma.kind::tf32 -> TMEM -> epilogue), one
//     128 x bn output tile per CTA; mbarrier phases, the TMEM allocation and the ring carry over from phase to phase;
//   * interpolation phase (gather.cu::interp_kernel) and segmented-sum phase (gather.cu::segsum_kernel) as grid-stride
//     loops -- the segmented sum with a warp per destination row, the bucket's permutation entries and weights fetched by
//     the lanes in parallel (one L2 round trip for the whole bucket instead of one per four entries).
// The arithmetic of every phase is the stand-alone kernel's, in the same order: results are bit-identical to the per-launch
// path (tests/test_gpu_deep.py), which stays selectable with psg_set_option("deep", 0).
//
// Grid barrier: a monotonic counter in the engine's workspace; the host tracks how many arrivals earlier launches left
// behind and passes the base, so no reset launch sits in the stream.  All CTAs are co-resident by construction (grid <=
// number of SMs, one CTA per SM); the spin gives up after ~2 s and raises the error word next to the counter instead of
// hanging the device if that assumption is ever violated (two engines driven from different streams without psg's
// sm_cap split).
#include <cstdio>
#include <cstring>
#include "psg_common.cuh"
#include "psg_internal.h"
#include "psg_tc.cuh"
#include "psg_segsum.cuh"

namespace {

constexpr int kMaxPhases = PSG_DEEP_MAX_PHASES;
constexpr int kMaxStages = 12;
constexpr int kBlk = 16;                      // chunk planes (4 floats of K each) per stage (gemm_tc.cu: the producer's per-stage cost bounds the K loop)
constexpr int kABytes = kBlk * 2048;          // 32 KB
constexpr int kThreads = 192;
constexpr int kRingBytes = 200 * 1024;
constexpr int kSmemBytes = kRingBytes + 1024;

enum { DP_GEMM = 0, DP_INTERP = 1, DP_SEGSUM = 2 };

struct DeepPhase {
    int kind;
    // DP_GEMM
    PsgGemmArgs g; int bn, ntn;
    // DP_INTERP
    TView i_src; int i_S; const int *i_idx; const float *i_w; long long i_rows; int i_N, i_nch; TView i_out;
    // DP_SEGSUM
    PsgSegsumArgs s;
};

struct DeepArgs {
    DeepPhase ph[kMaxPhases];
    int nph, nst, bstage;       // phases; operand-ring stages; bytes of one B stage (the widest bn of the list)
    unsigned *ctr;              // [0] barrier arrivals (monotonic), [1] error word
    unsigned base;              // arrivals left behind by earlier launches
    long long *trace;           // debug: globaltimer stamps of CTA 0 (psg_debug_trace), or null
};

__device__ __forceinline__ long long gtimer()
{
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}


__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// every CTA's global writes of the phase -> visible to every CTA's reads (generic and async proxy) of the next
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned target)
{
    __threadfence();
    tc::fence_async_all();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        while ((int)(ld_acquire(ctr) - target) < 0) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(ctr + 1, 1u); break; }     // ~2 s: never hang the device
        }
    }
    __syncthreads();
    tc::fence_async_all();
}

// ---- one 128 x BN output tile (gemm_tc.cu's roles; `it0` = ring uses so far, identical in every role) ----------------
__device__ __forceinline__ void gemm_item(const DeepPhase &p, int mt, int nt, int nst, int bstage, uint32_t sA, uint32_t sB,
                                          unsigned long long *bar_full, unsigned long long *bar_empty, unsigned long long *bar_acc,
                                          uint32_t tmem, uint32_t it0, uint32_t accph, int warp, int lane)
{
    const PsgGemmArgs &g = p.g;
    const long long row0 = (long long)mt * 128;
    const int n0 = nt * p.bn;
    const int BN = min(p.bn, g.nout_pad - n0);
    const int s0n = (g.k1chunks + kBlk - 1) / kBlk, s1n = (g.k2chunks + kBlk - 1) / kBlk, total = s0n + s1n;
    if (warp == 0) {
        // ---- producer: lane 0 the barrier bookkeeping + the A planes (one copy: whole planes are contiguous), lanes
        // 1..n one weight plane each ----
        for (int i = 0; i < total; ++i) {
            const int srcI = i < s0n ? 0 : 1;
            const int c = (srcI == 0 ? i : i - s0n) * kBlk;
            const int kglob = srcI == 0 ? 0 : g.k1chunks;
            const int n = min(kBlk, (srcI == 0 ? g.k1chunks : g.k2chunks) - c);
            const TView &src = srcI == 0 ? g.A1 : g.A2;
            const uint32_t it = it0 + (uint32_t)i;
            const int s = (int)(it % (uint32_t)nst);
            const uint32_t ph = (it / (uint32_t)nst) & 1u;
            const uint32_t full = tc::smem_u32(&bar_full[s]);
            if (lane == 0) {
                tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ph ^ 1u);
                tc::mbar_expect_tx(full, (uint32_t)(n * 2048 + n * BN * 16));
            }
            __syncwarp();
            if (lane == 0) tc::bulk_g2s(sA + s * kABytes, src.base + tv_off(src, row0, c), (uint32_t)(n * 2048), full);
            else if (lane <= n) {
                const int j = lane - 1;
                tc::bulk_g2s(sB + s * bstage + j * BN * 16, g.W + ((size_t)(kglob + c + j) * g.Nw + n0) * 4, (uint32_t)(BN * 16), full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_tf32(128, BN);
            uint32_t acc = 0;
            for (int i = 0; i < total; ++i) {
                const int srcI = i < s0n ? 0 : 1;
                const int c = (srcI == 0 ? i : i - s0n) * kBlk;
                const int n = min(kBlk, (srcI == 0 ? g.k1chunks : g.k2chunks) - c);
                const uint32_t it = it0 + (uint32_t)i;
                const int s = (int)(it % (uint32_t)nst);
                const uint32_t ph = (it / (uint32_t)nst) & 1u;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), ph);
                tc::fence_after_sync();
                for (int j = 0; j < n; j += 2) {
                    const uint64_t ad = tc::smem_desc(sA + s * kABytes + j * 2048, 2048, 128);
                    const uint64_t bd = tc::smem_desc(sB + s * bstage + j * BN * 16, (uint32_t)(BN * 16), 128);
                    tc::mma_tf32(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
                tc::mma_commit(tc::smem_u32(&bar_empty[s]));
            }
            tc::mma_commit(tc::smem_u32(bar_acc));
        }
    } else {
        // ---- epilogue: warp q owns TMEM lanes [32q, 32q + 32) ----
        const int q = warp & 3;
        const long long row = row0 + q * 32 + lane;
        tc::mbar_wait(tc::smem_u32(bar_acc), accph);
        tc::fence_after_sync();
        for (int c16 = 0; c16 < BN; c16 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c16, v);
            const int col = n0 + c16;
            if (g.epi == PSG_EPI_BIAS_RELU || g.epi == PSG_EPI_BIAS) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] += __ldg(g.bias + col + i);
                    if (g.epi == PSG_EPI_BIAS_RELU) v[i] = fmaxf(v[i], 0.f);
                }
            }
            if (g.epi == PSG_EPI_MASK) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 y = tv_ld(g.Mask, row, (col >> 2) + c);      // forward activations: constant during this kernel
                    v[4 * c + 0] = y.x > 0.f ? v[4 * c + 0] : 0.f;
                    v[4 * c + 1] = y.y > 0.f ? v[4 * c + 1] : 0.f;
                    v[4 * c + 2] = y.z > 0.f ? v[4 * c + 2] : 0.f;
                    v[4 * c + 3] = y.w > 0.f ? v[4 * c + 3] : 0.f;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                tv_st(g.Out, row, (col >> 2) + c, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
            if (g.Out2.base && col + 16 <= g.out2_cols) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tv_st(g.Out2, row, (col >> 2) + c, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
            }
        }
    }
}

// ---- out[row][c] = (f[i0] w0 + f[i1] w1) + f[i2] w2 (gather.cu::interp_kernel); sources written by an earlier phase of
// this kernel are read past L1 (ld.global.cg) ----
__device__ __forceinline__ void phase_interp(const DeepPhase &p)
{
    const long long total = p.i_rows * p.i_nch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += stride * U) {
        float4 a[U], b[U], d[U]; float w0[U], w1[U], w2[U]; long long row[U]; int ch[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long t = t0 + u * stride;
            const bool in = t < total;
            row[u] = in ? t % p.i_rows : 0;
            ch[u] = in ? (int)(t / p.i_rows) : -1;
            const long long base = (row[u] / p.i_N) * p.i_S;
            const int *ii = p.i_idx + row[u] * 3;
            const float *ww = p.i_w + row[u] * 3;
            const int c = in ? ch[u] : 0;
            a[u] = psg_ldcg4(p.i_src.base + tv_off(p.i_src, base + __ldg(ii), c));
            b[u] = psg_ldcg4(p.i_src.base + tv_off(p.i_src, base + __ldg(ii + 1), c));
            d[u] = psg_ldcg4(p.i_src.base + tv_off(p.i_src, base + __ldg(ii + 2), c));
            w0[u] = __ldg(ww); w1[u] = __ldg(ww + 1); w2[u] = __ldg(ww + 2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ch[u] < 0) continue;
            float4 r;
            r.x = __fadd_rn(__fadd_rn(__fmul_rn(a[u].x, w0[u]), __fmul_rn(b[u].x, w1[u])), __fmul_rn(d[u].x, w2[u]));
            r.y = __fadd_rn(__fadd_rn(__fmul_rn(a[u].y, w0[u]), __fmul_rn(b[u].y, w1[u])), __fmul_rn(d[u].y, w2[u]));
            r.z = __fadd_rn(__fadd_rn(__fmul_rn(a[u].z, w0[u]), __fmul_rn(b[u].z, w1[u])), __fmul_rn(d[u].z, w2[u]));
            r.w = __fadd_rn(__fadd_rn(__fmul_rn(a[u].w, w0[u]), __fmul_rn(b[u].w, w1[u])), __fmul_rn(d[u].w, w2[u]));
            tv_st(p.i_out, row[u], ch[u], r);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) deep_kernel(const __grid_constant__ DeepArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_full[kMaxStages], bar_empty[kMaxStages], bar_acc;
    __shared__ uint32_t tmem_slot;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = sbase, sB = sbase + a.nst * kABytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long *tr = (a.trace && blockIdx.x == 0 && threadIdx.x == 64) ? a.trace : nullptr;
    if (tr) { tr[0] = gtimer(); tr[100] = 0xDEE9; tr[101] = a.nph; }
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nst; ++s) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::mbar_init(tc::smem_u32(&bar_acc), 1);
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    if (tr) tr[1] = gtimer();
    const uint32_t tmem = tmem_slot;
    uint32_t it = 0, accph = 0;
    unsigned target = a.base;
    for (int pi = 0; pi < a.nph; ++pi) {
        const DeepPhase &p = a.ph[pi];
        if (p.kind == DP_GEMM) {
            const int nitems = p.g.mtiles * p.ntn;
            const int total = (p.g.k1chunks + kBlk - 1) / kBlk + (p.g.k2chunks + kBlk - 1) / kBlk;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                gemm_item(p, item / p.ntn, item % p.ntn, a.nst, a.bstage, sA, sB, bar_full, bar_empty, &bar_acc, tmem, it, accph, warp, lane);
                it += (uint32_t)total;
                accph ^= 1u;
                tc::fence_before_sync();
                __syncthreads();            // the accumulator is free again (and every role is done with the item)
                tc::fence_after_sync();
            }
        } else if (p.kind == DP_INTERP) {
            phase_interp(p);
        } else {
            const int nc = (p.s.nch + 31) / 32;
            const long long w0 = (long long)blockIdx.x * (blockDim.x >> 5) + warp, ws = (long long)gridDim.x * (blockDim.x >> 5);
            if (nc <= 1) psg_segsum_warp<1, true>(p.s, w0, ws, lane);
            else if (nc == 2) psg_segsum_warp<2, true>(p.s, w0, ws, lane);
            else psg_segsum_warp<4, true>(p.s, w0, ws, lane);
        }
        if (tr) { tr[2 + 2 * pi] = gtimer(); tr[102 + pi] = p.kind; }
        if (pi + 1 < a.nph) {
            target += gridDim.x;
            grid_barrier(a.ctr, target);
        }
        if (tr) tr[3 + 2 * pi] = gtimer();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 128);
}

DeepArgs g_rec;
bool g_recording = false;
int g_sms = 0;
int g_bn_min = 32, g_items_min = 96;     // column-tile rule (psg_set_option "deep_bn_min" / "deep_items")

}  // namespace

void psg_deep_tune(int bn_min, int items_min)
{
    if (bn_min >= 16) g_bn_min = bn_min;
    if (items_min > 0) g_items_min = items_min;
}
void psg_deep_begin()
{
    memset(&g_rec, 0, sizeof(g_rec));
    g_recording = true;
}
bool psg_deep_recording() { return g_recording; }
int psg_deep_phases() { return g_recording ? g_rec.nph : 0; }
void psg_deep_cancel() { g_recording = false; g_rec.nph = 0; }

bool psg_deep_can_gemm(const PsgGemmArgs &g)
{
    return g_recording && g_rec.nph < kMaxPhases && g.k1chunks > 0 && g.k1chunks % 4 == 0 && g.k2chunks % 4 == 0 && g.mtiles > 0 &&
           g.nout_pad % 16 == 0 && g.Nw >= g.nout_pad;
}
int psg_deep_add_gemm(const PsgGemmArgs &g)
{
    if (!psg_deep_can_gemm(g)) return PSG_EUNSUPPORTED;
    DeepPhase &p = g_rec.ph[g_rec.nph++];
    p.kind = DP_GEMM; p.g = g;
    // few row tiles: narrower column tiles spread the layer over more SMs (gemm_tc.cu's rule)
    int bn = 128;
    while (bn > g_bn_min && g.mtiles * ((g.nout_pad + bn - 1) / bn) < g_items_min) bn >>= 1;
    p.bn = bn; p.ntn = (g.nout_pad + bn - 1) / bn;
    return PSG_OK;
}
int psg_deep_add_interp(TView feats, int S, const int *idx, const float *w, long long P, int N, int nch, TView out)
{
    if (!g_recording || g_rec.nph >= kMaxPhases) return PSG_EUNSUPPORTED;
    DeepPhase &p = g_rec.ph[g_rec.nph++];
    p.kind = DP_INTERP;
    p.i_src = feats; p.i_S = S; p.i_idx = idx; p.i_w = w; p.i_rows = P * N; p.i_N = N; p.i_nch = nch; p.i_out = out;
    return PSG_OK;
}
bool psg_deep_can_segsum(int ncols) { return g_recording && g_rec.nph < kMaxPhases && (ncols + 3) / 4 <= 128; }
int psg_deep_add_segsum(TView src, long long src_rows_per_p, int div, const float *wgt, const int *offs, const int *perm, int M, int R,
                        long long P, int ncols, TView dst, int accumulate, const TView *relu_mask, const float *src_rm, int rm_stride)
{
    if (!psg_deep_can_segsum(ncols)) return PSG_EUNSUPPORTED;
    DeepPhase &p = g_rec.ph[g_rec.nph++];
    p.kind = DP_SEGSUM;
    p.s.src = src; p.s.rows_per_p = src_rows_per_p; p.s.div = div; p.s.wgt = wgt; p.s.offs = offs; p.s.perm = perm; p.s.M = M; p.s.R = R;
    p.s.P = P; p.s.nch = (ncols + 3) / 4; p.s.tail = ncols & 3; p.s.dst = dst; p.s.acc = accumulate;
    p.s.rmask = relu_mask ? *relu_mask : TView{nullptr, 0, 0}; p.s.rm = src_rm; p.s.rm_stride = rm_stride;
    return PSG_OK;
}

// launch the recorded phases; `ctr` = two words of device memory owned by the engine (zeroed when it was bound),
// `*epoch` = the host's count of the arrivals earlier launches left in ctr[0]
int psg_deep_flush(unsigned *ctr, unsigned *epoch, cudaStream_t st)
{
    g_recording = false;
    DeepArgs &a = g_rec;
    if (a.nph == 0) return PSG_OK;
    if (!ctr || !epoch) return PSG_EINVAL;
    if (g_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return PSG_ECUDA;
    }
    static PsgDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(deep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return PSG_ECUDA;
        attr_once.mark();
    }
    int bnmax = 32;
    for (int i = 0; i < a.nph; ++i)
        if (a.ph[i].kind == DP_GEMM && a.ph[i].bn > bnmax) bnmax = a.ph[i].bn;
    a.bstage = kBlk * bnmax * 16;
    a.nst = kRingBytes / (kABytes + a.bstage);
    if (a.nst > kMaxStages) a.nst = kMaxStages;
    const int grid = (g_psg_sm_cap > 0 && g_psg_sm_cap < g_sms) ? g_psg_sm_cap : g_sms;
    a.ctr = ctr; a.base = *epoch;
    a.trace = psg_tile_trace_slot();
    *epoch += (unsigned)(a.nph - 1) * (unsigned)grid;
    if (psg_launch_pdl(deep_kernel, dim3((unsigned)grid), dim3(kThreads), (size_t)kSmemBytes, st, 1, a) != cudaSuccess) return PSG_ECUDA;
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}
