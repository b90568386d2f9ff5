// streambench.cu -- measurement kernel (not on the product path): how fast can an SM pull L2-resident bytes into shared
// memory with bulk async copies, alone, with every SM pulling, and through cluster multicast?  The deep levels of the
// network (few rows, large weights) are bound by exactly this stream, so their ceiling is measured,
// not assumed.  tools/l2_stream_bench.py drives it; psg_debug_l2_stream is declared in include/psg_b200.h.
//
// Every CTA runs an 8-stage ring of 16 KB stages.  mode 0: CTA i streams its own region; mode 1: every CTA streams
// region 0 (the pattern of weight streaming: all CTAs want the same bytes at about the same time).  cluster > 1: the CTAs
// of a cluster want the SAME bytes (the region of the cluster) and CTA q issues stage s only if s % cluster == q, with
// .multicast::cluster to all peers; a peer that has seen a stage land arrives on the issuer's "empty" barrier, which
// the issuer waits on (cluster arrivals) before it reuses the stage.
#include <cstdio>
#include "psg_common.cuh"
#include "psg_tc.cuh"
#include "../../include/psg_b200.h"

namespace {

constexpr int kMaxStages = 16;
int g_stages = 8, g_stage_bytes = 16 * 1024, g_rings = 1;      // psg_set_option "stream_stages" / "stream_stage_bytes"

__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar, uint16_t mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(256, 1) stream_kernel(const unsigned char *buf, long long region_bytes, int mode, int passes, int kStages, int kStageBytes, int rings)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar_full[8 * kMaxStages], bar_empty[8 * kMaxStages];
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int q = CS > 1 ? (int)tc::cluster_ctarank() : 0;
    const int cl = (int)blockIdx.x / CS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages * rings; ++s) { tc::mbar_init(tc::smem_u32(&bar_full[s]), 1); tc::mbar_init(tc::smem_u32(&bar_empty[s]), CS); }
        tc::fence_mbar_init();
    }
    __syncthreads();
    if (CS > 1) tc::cluster_sync_all();
    const unsigned char *region = buf + (mode == 1 ? 0 : (long long)cl * region_bytes);
    const int nst = (int)(region_bytes / kStageBytes) * passes;
    const int per_pass = (int)(region_bytes / kStageBytes);
    const int ring = threadIdx.x >> 5;
    if (CS == 1 && (threadIdx.x & 31) == 0 && ring < rings) {
        // `rings` independent rings per CTA (one warp each, producer + consumer in one thread): ring w streams every
        // rings-th stage of the region through its own slots and barriers
        unsigned long long *bf = bar_full + ring * kStages, *be = bar_empty + ring * kStages;
        const uint32_t sb = sbase + (uint32_t)ring * kStages * kStageBytes;
        const int mine = nst / rings;
        for (int it = 0; it < mine + kStages; ++it) {
            const int jt = it - kStages;
            if (jt >= 0) {
                const int s = jt % kStages;
                tc::mbar_wait(tc::smem_u32(&bf[s]), (uint32_t)(jt / kStages) & 1u);
                tc::mbar_arrive(tc::smem_u32(&be[s]));
            }
            if (it < mine) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                const uint32_t full = tc::smem_u32(&bf[s]);
                tc::mbar_expect_tx(full, kStageBytes);
                if (it >= kStages) tc::mbar_wait(tc::smem_u32(&be[s]), ph ^ 1u);
                tc::bulk_g2s(sb + s * kStageBytes, region + (long long)((it * rings + ring) % per_pass) * kStageBytes, kStageBytes, full);
            }
        }
    } else if (CS > 1 && threadIdx.x == 0) {
        // producer + consumer in one thread: keep kStages stages in flight
        for (int it = 0; it < nst + kStages; ++it) {
            const int jt = it - kStages;                           // first consume the stage issued kStages iterations ago ...
            if (jt >= 0) {
                const int s = jt % kStages;
                const uint32_t ph = (uint32_t)(jt / kStages) & 1u;
                tc::mbar_wait(tc::smem_u32(&bar_full[s]), ph);
                if (CS == 1) tc::mbar_arrive(tc::smem_u32(&bar_empty[s]));
                else tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&bar_empty[s]), (uint32_t)(jt % CS)));
            }
            if (it < nst) {                                        // ... then refill the same slot
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                const uint32_t full = tc::smem_u32(&bar_full[s]);
                tc::mbar_expect_tx(full, kStageBytes);           // every CTA expects the stage, whoever issues it
                if (CS == 1) {
                    if (it >= kStages) tc::mbar_wait(tc::smem_u32(&bar_empty[s]), ph ^ 1u);
                    tc::bulk_g2s(sbase + s * kStageBytes, region + (long long)(it % per_pass) * kStageBytes, kStageBytes, full);
                } else if (it % CS == q) {
                    if (it >= kStages) tc::mbar_wait_cluster(tc::smem_u32(&bar_empty[s]), ph ^ 1u);
                    bulk_g2s_mc(sbase + s * kStageBytes, region + (long long)(it % per_pass) * kStageBytes, kStageBytes, full,
                                (uint16_t)((1u << CS) - 1u));
                }
            }
        }
    }
    __syncthreads();
    if (CS > 1) tc::cluster_sync_all();
}

template <int CS>
int launch(const void *buf, long long region_bytes, int mode, int passes, int ctas, cudaStream_t st)
{
    const int kStages = g_stages, kStageBytes = g_stage_bytes;
    const int rings = CS == 1 ? g_rings : 1;
    const size_t smem = (size_t)rings * kStages * kStageBytes + 1024;
    if (smem > 200 * 1024) return PSG_EINVAL;
    static PsgDeviceOnce once;
    if (once.need()) {
        if (cudaFuncSetAttribute(stream_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return PSG_ECUDA;
        once.mark();
    }
    cudaLaunchConfig_t cfg;
    cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, stream_kernel<CS>, (const unsigned char *)buf, region_bytes, mode, passes, kStages, kStageBytes, rings) != cudaSuccess) return PSG_ECUDA;
    return PSG_OK;
}

}  // namespace

void psg_stream_tune(int stages, int stage_bytes, int rings)
{
    if (rings >= 1 && rings <= 8) g_rings = rings;
    if (stages >= 1 && stages <= kMaxStages) g_stages = stages;
    if (stage_bytes >= 1024 && stage_bytes % 1024 == 0 && (long long)g_stages * stage_bytes <= 192 * 1024) g_stage_bytes = stage_bytes;
}

// NOTE on the consume step with clusters: a CTA waits on its own full barrier for a stage that a PEER issued; it must
// have armed expect_tx before the peer's copy can complete, which the single-threaded loop above guarantees only per
// CTA, not across CTAs -- a peer's bytes may land before this CTA armed the barrier.  complete_tx before expect_tx is
// legal (the transaction count goes transiently negative), so no ordering is needed.
extern "C" int psg_debug_l2_stream(const void *buf, int64_t region_bytes, int mode, int cluster, int passes, int ctas, psg_stream_t stream)
{
    const int kStageBytes = g_stage_bytes;
    if (!buf || region_bytes < kStageBytes || region_bytes % kStageBytes || passes < 1 || ctas < 1 || ctas % (cluster > 0 ? cluster : 1)) return PSG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (cluster) {
    case 1: return launch<1>(buf, region_bytes, mode, passes, ctas, st);
    case 2: return launch<2>(buf, region_bytes, mode, passes, ctas, st);
    case 4: return launch<4>(buf, region_bytes, mode, passes, ctas, st);
    case 8: return launch<8>(buf, region_bytes, mode, passes, ctas, st);
    }
    return PSG_EINVAL;
}
