// psg_tmap.cuh -- 2-D tensor maps (TMA descriptors) over the packed weights [K/4][Nw][4].
//
// A weight stage of a column-sliced GEMM is `box_planes` K-chunks x `box_n` output columns: in the packed
// layout that is box_planes separate runs of box_n * 16 bytes, Nw * 16 bytes apart.  Issued as one 1-D bulk
// copy per run, the copies themselves become the bottleneck (~65 ns each through the SM's copy engine: 72
// copies = 5 us for a K = 256 layer, measured).  One tiled TMA copy per stage fetches the whole box and lays
// it down densely as [plane][box_n][4] -- exactly the UMMA K-major no-swizzle B operand.
//
// The map is declared over 8-byte elements so that a 128-column slice (2 KB per plane) stays within the
// 256-element limit of the innermost box dimension.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// host: out <- map over w [planes][nw][4] floats with boxes of [box_planes][box_n][4]; false if unsupported
static inline bool psg_weight_tmap(CUtensorMap *out, const float *w, long long planes, int nw, int box_planes, int box_n)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    if (!fn || box_n * 2 > 256 || box_planes > 256 || box_n < 1 || box_planes < 1 || ((uintptr_t)w & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)nw * 2, (cuuint64_t)planes};
    const cuuint64_t strides[1] = {(cuuint64_t)nw * 16};
    const cuuint32_t box[2] = {(cuuint32_t)box_n * 2, (cuuint32_t)box_planes};
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<float *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// device: box at (first output column n0, first plane p0) -> dst (shared), completion on mbarrier `bar`
__device__ __forceinline__ void psg_tmap_load(uint32_t dst_smem, const CUtensorMap *map, int n0, int p0, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
                 "l"(map), "r"(n0 * 2), "r"(p0), "r"(bar)
                 : "memory");
}
#endif
