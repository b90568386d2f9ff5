// chain_fused.cu -- the finest feature-propagation level and the classification head as ONE kernel:
//   3-NN inverse-distance interpolation -> fp1 MLP (3 x 128) -> conv1+bn1+ReLU -> conv2 -> logits
//   [-> loss gradient -> conv2^T -> conv1^T -> fp1^T x 3 -> gradient w.r.t. the interpolated rows]
//
// Reference: PointNet/models/pointnet_util.py:305-319 (interpolate + MLP), pointnet2_sem_seg.py:34-39
// (fp1, conv1, bn1, drop1 (identity in eval), conv2, log_softmax), the attack costs of
// nontarget.py:34,120-128 / target.py:38,149-168, and autograd of all of it.
//
// Every one of these layers is row-local (row = point), so a 128-row tile runs the whole forward
// AND the whole backward on-chip: activations ping through one shared-memory A-operand buffer
// (UMMA K-major, no swizzle) and a TMEM accumulator, the ReLU masks the backward needs are bits in
// shared memory, the loss gradient is computed in registers from the 13 logits of the thread's own
// row.  HBM sees only the interpolation gather (L2-resident coarse features) and the 128-column
// gradient rows handed to the interpolation backward.  Ten GEMMs per tile and no activation traffic,
// against 10 separate GEMM launches + 1.1 KB/row of activation round trips in the unfused path.
//
// The 9 weight matrices (580 KB) do not fit in shared memory: a producer thread streams them
// through a ring of 32 KB stages with 1-D bulk copies (TMA), and each stage is consumed by BOTH tiles
// in flight before it is released, halving the L2 -> SM weight traffic.
#include "psg_common.cuh"
#include "psg_internal.h"
#include "psg_loss.cuh"
#include "psg_tc.cuh"

namespace {

constexpr int NG = 2;                       // tiles in flight per CTA
constexpr int kWorkers = 128;
constexpr int kThreads = NG * 128 + 64;     // workers + MMA warp + weight-producer warp
constexpr int kStageBytes = 32 * 1024;
constexpr int kStages = 2;
constexpr int kMaskSlots = 4;
constexpr int kABytes = 128 * 512;          // A operand: up to 128 columns

enum { EPI_RELU_SAVE = 0, EPI_HEAD = 1, EPI_MASK = 2, EPI_STORE = 3 };

struct ChainOp {
    const float *w;        // packed [planes][n][4], contiguous
    const float *bias;
    int planes, n, nstages, epi, slot;
};

struct ChainArgs {
    ChainOp ops[PSG_CHAIN_MAX_OPS];
    int nops;
    // source: 3-NN interpolation of `src` rows
    TView src; int S; const int *nn_idx; const float *nn_w; int Nf; int kin;
    long long rows; int ntiles;
    // head / loss
    int backward;          // 0: stop at the logits (stored to zout); 1: loss gradient + backward chain
    int ncls, loss_kind, target;
    const int *labels; float scale, kappa; const float *dlogp;
    float *loss_rows; unsigned char *hit;
    TView zout, dI;
};

__device__ __forceinline__ float4 *plane_ptr(unsigned char *buf, int chunk, int row)
{
    return reinterpret_cast<float4 *>(buf) + (size_t)chunk * 128 + row;
}

__global__ void __launch_bounds__(kThreads) chain_kernel(const __grid_constant__ ChainArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_in[NG], bar_acc[NG], bar_full[kStages], bar_empty[kStages];
    __shared__ uint32_t tmem_slot;

    const uint32_t s0 = tc::smem_u32(smem_raw);
    const uint32_t sbase = (s0 + 1023u) & ~1023u;
    unsigned char *base = smem_raw + (sbase - s0);
    // [A g0][A g1][W stage 0][W stage 1][mask bits g0][mask bits g1]
    const uint32_t sA0 = sbase, sW = sbase + NG * kABytes;
    unsigned char *pA0 = base;
    unsigned *pMask = reinterpret_cast<unsigned *>(base + NG * kABytes + kStages * kStageBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int g = 0; g < NG; ++g) { tc::mbar_init(tc::smem_u32(&bar_in[g]), kWorkers); tc::mbar_init(tc::smem_u32(&bar_acc[g]), 1); }
        for (int s = 0; s < kStages; ++s) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1); }
        tc::fence_mbar_init();
    }
    if (warp == NG * 4) tc::tmem_alloc(tc::smem_u32(&tmem_slot), NG * 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int tstride = gridDim.x * NG;

    if (warp == NG * 4 + 1) {
        // ---------------- weight producer ----------------
        if (lane == 0) {
            int it = 0;
            for (int tile0 = blockIdx.x * NG; tile0 < a.ntiles; tile0 += tstride) {
                for (int o = 0; o < a.nops; ++o) {
                    const ChainOp &op = a.ops[o];
                    const uint32_t bytes = (uint32_t)(op.planes / op.nstages) * op.n * 16;
                    for (int s = 0; s < op.nstages; ++s, ++it) {
                        const int slot = it % kStages;
                        const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                        tc::mbar_wait(tc::smem_u32(&bar_empty[slot]), ph ^ 1u);
                        const uint32_t full = tc::smem_u32(&bar_full[slot]);
                        tc::mbar_expect_tx(full, bytes);
                        tc::bulk_g2s(sW + slot * kStageBytes, reinterpret_cast<const unsigned char *>(op.w) + (size_t)s * bytes,
                                     bytes, full);
                    }
                }
            }
        }
    } else if (warp == NG * 4) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            int it = 0;
            uint32_t ph_in[NG] = {0u, 0u};
            for (int tile0 = blockIdx.x * NG; tile0 < a.ntiles; tile0 += tstride) {
                for (int o = 0; o < a.nops; ++o) {
                    const ChainOp &op = a.ops[o];
                    const int pps = op.planes / op.nstages;
                    const uint32_t idesc = tc::idesc_tf32(128, op.n);
                    for (int s = 0; s < op.nstages; ++s, ++it) {
                        const int slot = it % kStages;
                        tc::mbar_wait(tc::smem_u32(&bar_full[slot]), (uint32_t)(it / kStages) & 1u);
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            if (tile0 + g >= a.ntiles) continue;
                            if (s == 0) { tc::mbar_wait(tc::smem_u32(&bar_in[g]), ph_in[g]); ph_in[g] ^= 1u; }
                            tc::fence_after_sync();
                            const uint32_t sA = sA0 + g * kABytes + (uint32_t)(s * pps) * 2048u;
                            const uint32_t sB = sW + slot * kStageBytes;
                            for (int j = 0; j < pps; j += 2) {
                                const uint64_t ad = tc::smem_desc(sA + j * 2048, 2048, 128);
                                const uint64_t bd = tc::smem_desc(sB + j * op.n * 16, (uint32_t)(op.n * 16), 128);
                                tc::mma_tf32(tmem + g * 128, ad, bd, idesc, (s > 0 || j > 0) ? 1u : 0u);
                            }
                            if (s == op.nstages - 1) tc::mma_commit(tc::smem_u32(&bar_acc[g]));
                        }
                        tc::mma_commit(tc::smem_u32(&bar_empty[slot]));     // stage free once both tiles consumed it
                    }
                }
            }
        }
    } else {
        // ---------------- workers: thread = tile row = TMEM lane ----------------
        const int grp = warp >> 2, wq = warp & 3;
        const int r = threadIdx.x & 127;
        unsigned char *pA = pA0 + (size_t)grp * kABytes;
        unsigned *mbits = pMask + (size_t)grp * kMaskSlots * 4 * 128;
        const uint32_t b_in = tc::smem_u32(&bar_in[grp]), b_acc = tc::smem_u32(&bar_acc[grp]);
        const uint32_t tl = tmem + grp * 128 + ((uint32_t)(wq * 32) << 16);
        uint32_t ph = 0;
        for (int tile = blockIdx.x * NG + grp; tile < a.ntiles; tile += tstride) {
            const long long row = (long long)tile * 128 + r;
            const bool valid = row < a.rows;
            // ---- interpolation: (f[i0] w0 + f[i1] w1) + f[i2] w2, products rounded separately ----
            {
                const long long rr = valid ? row : 0;
                const long long p = rr / a.Nf;
                const int *ii = a.nn_idx + rr * 3;
                const float *ww = a.nn_w + rr * 3;
                const long long r0 = p * a.S + ii[0], r1 = p * a.S + ii[1], r2 = p * a.S + ii[2];
                const float w0 = valid ? ww[0] : 0.f, w1 = valid ? ww[1] : 0.f, w2 = valid ? ww[2] : 0.f;
#pragma unroll 4
                for (int c = 0; c < a.kin / 4; ++c) {
                    const float4 x = tv_ld(a.src, r0, c), y = tv_ld(a.src, r1, c), z = tv_ld(a.src, r2, c);
                    float4 q;
                    q.x = __fadd_rn(__fadd_rn(__fmul_rn(x.x, w0), __fmul_rn(y.x, w1)), __fmul_rn(z.x, w2));
                    q.y = __fadd_rn(__fadd_rn(__fmul_rn(x.y, w0), __fmul_rn(y.y, w1)), __fmul_rn(z.y, w2));
                    q.z = __fadd_rn(__fadd_rn(__fmul_rn(x.z, w0), __fmul_rn(y.z, w1)), __fmul_rn(z.z, w2));
                    q.w = __fadd_rn(__fadd_rn(__fmul_rn(x.w, w0), __fmul_rn(y.w, w1)), __fmul_rn(z.w, w2));
                    *plane_ptr(pA, c, r) = q;
                }
            }
            tc::fence_async_smem();
            tc::mbar_arrive(b_in);
            for (int o = 0; o < a.nops; ++o) {
                const ChainOp &op = a.ops[o];
                tc::mbar_wait(b_acc, ph); ph ^= 1u; tc::fence_after_sync();
                if (op.epi == EPI_RELU_SAVE) {
                    unsigned bits = 0;
                    unsigned *mslot = mbits + (size_t)op.slot * 4 * 128;
                    for (int c16 = 0; c16 < op.n; c16 += 16) {
                        float v[16];
                        tc::tmem_ld16(tl + (uint32_t)c16, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            v[i] = fmaxf(v[i] + __ldg(op.bias + c16 + i), 0.f);
                            bits |= (v[i] > 0.f ? 1u : 0u) << ((c16 & 16) + i);
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *plane_ptr(pA, (c16 >> 2) + c, r) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                        if ((c16 & 16) || c16 + 16 >= op.n) { mslot[(c16 >> 5) * 128 + r] = bits; bits = 0; }
                    }
                } else if (op.epi == EPI_MASK) {
                    unsigned bits = 0;
                    const unsigned *mslot = mbits + (size_t)op.slot * 4 * 128;
                    for (int c16 = 0; c16 < op.n; c16 += 16) {
                        if ((c16 & 16) == 0) bits = mslot[(c16 >> 5) * 128 + r];
                        float v[16];
                        tc::tmem_ld16(tl + (uint32_t)c16, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (!((bits >> ((c16 & 16) + i)) & 1u)) v[i] = 0.f;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *plane_ptr(pA, (c16 >> 2) + c, r) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    }
                } else if (op.epi == EPI_HEAD) {
                    float v[16], dz[16];
                    tc::tmem_ld16(tl, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += __ldg(op.bias + i);
                    if (!a.backward) {
                        if (valid) {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                tv_st(a.zout, row, c, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
                        }
                    } else {
                        const long long rr = valid ? row : 0;
                        if (a.loss_kind == 0) {
                            psg_dz_generic(v, a.ncls, a.dlogp + rr * a.ncls, dz);
                        } else if (a.loss_kind == 1) {
                            psg_dz_ce_row(v, a.ncls, a.target >= 0 ? a.target : a.labels[rr], a.scale, dz);
                        } else {
                            int h;
                            const float f = psg_dz_cw_row(v, a.ncls, a.target >= 0 ? a.target : a.labels[rr], a.kappa, a.scale, dz, h);
                            if (valid && a.loss_rows) a.loss_rows[row] = f;
                            if (valid && a.hit) a.hit[row] = (unsigned char)h;
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float4 q = make_float4(dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
                            if (!valid) q = make_float4(0.f, 0.f, 0.f, 0.f);
                            *plane_ptr(pA, c, r) = q;
                        }
                    }
                } else {   // EPI_STORE
                    for (int c16 = 0; c16 < op.n; c16 += 16) {
                        float v[16];
                        tc::tmem_ld16(tl + (uint32_t)c16, v);
                        if (valid) {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                tv_st(a.dI, row, (c16 >> 2) + c, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
                        }
                    }
                }
                tc::fence_before_sync();
                if (o + 1 < a.nops) {
                    tc::fence_async_smem();
                    tc::mbar_arrive(b_in);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == NG * 4) tc::tmem_dealloc(tmem, NG * 128);
}

int g_sms = 0;

}  // namespace

// One description of a chain: forward layers (folded conv + bias + ReLU), head layer (bias only),
// and -- when `backward` -- the same layers' dgrad weights in reverse.
int psg_chain_fused(const PsgChain &c, cudaStream_t st)
{
    if (c.nlayers < 1 || c.nlayers > 4 || c.kin % 16 || c.kin > 128 || c.ncls > kPsgMaxCls) return PSG_EUNSUPPORTED;
    ChainArgs a;
    int no = 0;
    int kprev = c.kin;
    for (int j = 0; j < c.nlayers; ++j) {          // hidden layers: bias + ReLU, bits saved in slot j
        if (c.n[j] % 16 || c.n[j] > 128 || c.nwf[j] != c.n[j]) return PSG_EUNSUPPORTED;
        ChainOp &o = a.ops[no++];
        o.w = c.wf[j]; o.bias = c.bias[j]; o.planes = kprev / 4; o.n = c.n[j];
        o.nstages = (o.planes * o.n * 16 + kStageBytes - 1) / kStageBytes; o.epi = EPI_RELU_SAVE; o.slot = j;
        if (o.planes % (2 * o.nstages)) return PSG_EUNSUPPORTED;
        kprev = c.n[j];
    }
    {                                               // head: N = packed width (zero-padded columns)
        if (c.head_nwf > 128 || c.head_nwf % 16) return PSG_EUNSUPPORTED;
        ChainOp &o = a.ops[no++];
        o.w = c.head_wf; o.bias = c.head_bias; o.planes = kprev / 4; o.n = c.head_nwf;
        o.nstages = (o.planes * o.n * 16 + kStageBytes - 1) / kStageBytes; o.epi = EPI_HEAD; o.slot = 0;
        if (o.planes % (2 * o.nstages)) return PSG_EUNSUPPORTED;
    }
    if (c.backward) {
        {                                           // d hidden_last = dz * W_head, masked by the last hidden layer's bits
            if (c.head_nwb != kprev) return PSG_EUNSUPPORTED;
            ChainOp &o = a.ops[no++];
            o.w = c.head_wb; o.bias = nullptr; o.planes = 16 / 4; o.n = kprev; o.nstages = 1; o.epi = EPI_MASK;
            o.slot = c.nlayers - 1;
        }
        for (int j = c.nlayers - 1; j >= 0; --j) {
            const int nin = j > 0 ? c.n[j - 1] : c.kin;
            if (c.nwb[j] != nin) return PSG_EUNSUPPORTED;
            ChainOp &o = a.ops[no++];
            o.w = c.wb[j]; o.bias = nullptr; o.planes = c.n[j] / 4; o.n = nin;
            o.nstages = (o.planes * o.n * 16 + kStageBytes - 1) / kStageBytes;
            if (o.planes % (2 * o.nstages)) return PSG_EUNSUPPORTED;
            o.epi = j > 0 ? EPI_MASK : EPI_STORE; o.slot = j - 1;
        }
    }
    a.nops = no;
    a.src = c.src; a.S = c.S; a.nn_idx = c.nn_idx; a.nn_w = c.nn_w; a.Nf = c.Nf; a.kin = c.kin;
    a.rows = c.rows; a.ntiles = (int)((c.rows + 127) / 128);
    a.backward = c.backward; a.ncls = c.ncls; a.loss_kind = c.loss_kind; a.target = c.target; a.labels = c.labels;
    a.scale = c.scale; a.kappa = c.kappa; a.dlogp = c.dlogp; a.loss_rows = c.loss_rows; a.hit = c.hit;
    a.zout = c.zout; a.dI = c.dI;
    const size_t smem = (size_t)NG * kABytes + (size_t)kStages * kStageBytes + (size_t)NG * kMaskSlots * 4 * 128 * 4 + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSG_ECUDA;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return PSG_ECUDA;
        attr_done = true;
    }
    const int want = (a.ntiles + NG - 1) / NG;
    const int grid = want < g_sms ? want : g_sms;
    chain_kernel<<<grid, kThreads, smem, st>>>(a);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}
