// slicer.cu -- whole-scene block slicer.
//
// Reference: PointNet/data_utils/S3DISDataLoader.py:124-175, ScannetDatasetWholeScene.__getitem__:
// a room [P,6+] of float64 points is cut into overlapping block_size x block_size columns on a
// stride grid; per column `np.where` collects the member points in ascending order, the list is
// padded to a multiple of block_points by np.random.choice, shuffled, gathered and normalised.
// In the reference that is a Python loop of numpy passes over the whole room per column plus
// O(columns^2) vstack copies.
//
// Here the room stays resident in HBM and the work is four launches:
//   scene_minmax        room bounding box (the reference's np.amin / np.amax)
//   cell_count / scan   members per (column, 1024-point chunk), exclusive scan per column
//   cell_fill           ordered compaction: sel[column] = ascending member indices (= np.where)
//   scene_gather        rows -> [x - cx, y - cy, z, rgb / 255, xyz / room_max], label, weight, index
// The random draws (choice + shuffle) depend only on the member COUNTS, so the host makes them on
// numpy's global generator in the reference's order and ships positions, not points
// (pointsecguard_b200/data_utils/S3DISDataLoader.py).  All arithmetic is float64 single operations
// (one subtraction or one division per value), so the result equals numpy's bit for bit.
#include "psg_common.cuh"
#include "psg_internal.h"

namespace {

constexpr int kChunk = 1024;      // points per CTA = threads per CTA

__device__ __forceinline__ double warp_min(double v)
{
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// partial[block][6] = (min x, y, z, max x, y, z) of a grid-strided slice; the last block to finish folds the partials
__global__ void __launch_bounds__(256) scene_minmax_kernel(const double *__restrict__ pts, long long P, int ld,
                                                            double *__restrict__ partial, unsigned *__restrict__ ticket,
                                                            double *__restrict__ out6)
{
    __shared__ double sm[6][8];
    __shared__ bool last;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < P; i += (long long)gridDim.x * 256) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = pts[i * ld + c];
            lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double a = warp_min(lo[c]), b = warp_max(hi[c]);
        if (lane == 0) { sm[c][warp] = a; sm[3 + c][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, sm[threadIdx.x][w]) : fmax(v, sm[threadIdx.x][w]);
        partial[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < 6) {
        double v = partial[threadIdx.x];
        for (unsigned b = 1; b < gridDim.x; ++b) {
            const double q = partial[(size_t)b * 6 + threadIdx.x];
            v = threadIdx.x < 3 ? fmin(v, q) : fmax(v, q);
        }
        out6[threadIdx.x] = v;
        if (threadIdx.x == 0) *ticket = 0u;          // ready for the next call
    }
}

// membership test of S3DISDataLoader.py:143-145 against the column's padded bounds (lo_x, hi_x, lo_y, hi_y),
// which the host computes with the reference's own float64 expressions
__device__ __forceinline__ bool in_cell(double x, double y, const double4 b)
{
    return x >= b.x && x <= b.y && y >= b.z && y <= b.w;
}

// FILL = false: counts[cell][chunk] = members of `cell` among the chunk's 1024 points
// FILL = true : sel[cell_off[cell] + base[cell][chunk] + rank within chunk] = point index, ascending
// One CTA per chunk walks all columns; a column whose bounds miss the chunk's bounding box costs one
// uniform branch (S3DIS rooms are stored object by object, so most chunks touch few columns).
template <bool FILL>
__global__ void __launch_bounds__(kChunk) cell_pass_kernel(const double *__restrict__ pts, long long P, int ld,
                                                            const double4 *__restrict__ bounds, int ncell, int nchunk,
                                                            int *__restrict__ counts, const long long *__restrict__ cell_off,
                                                            int *__restrict__ sel)
{
    __shared__ int wc[2][32];
    __shared__ double bb[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * kChunk + threadIdx.x;
    const bool live = i < P;
    const double x = live ? pts[i * ld] : 0.0, y = live ? pts[i * ld + 1] : 0.0;
    {
        const double a = warp_min(live ? x : INFINITY), b = warp_max(live ? x : -INFINITY);
        const double c = warp_min(live ? y : INFINITY), d = warp_max(live ? y : -INFINITY);
        if (lane == 0) { bb[0][warp] = a; bb[1][warp] = b; bb[2][warp] = c; bb[3][warp] = d; }
    }
    __syncthreads();
    const double xmin = warp_min(bb[0][lane]), xmax = warp_max(bb[1][lane]);
    const double ymin = warp_min(bb[2][lane]), ymax = warp_max(bb[3][lane]);
    int par = 0;
    for (int c = 0; c < ncell; ++c) {
        const double2 b01 = __ldg(reinterpret_cast<const double2 *>(bounds + c)), b23 = __ldg(reinterpret_cast<const double2 *>(bounds + c) + 1);
        const double4 b = make_double4(b01.x, b01.y, b23.x, b23.y);
        if (xmax < b.x || xmin > b.y || ymax < b.z || ymin > b.w) {          // uniform: no member in this chunk
            if (!FILL && threadIdx.x == 0) counts[(size_t)c * nchunk + blockIdx.x] = 0;
            continue;
        }
        const bool f = live && in_cell(x, y, b);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wc[par][warp] = __popc(m);
        __syncthreads();
        const int v = wc[par][lane];
        if (!FILL) {
            const int tot = __reduce_add_sync(0xffffffffu, v);
            if (threadIdx.x == 0) counts[(size_t)c * nchunk + blockIdx.x] = tot;
        } else {
            const int before = __reduce_add_sync(0xffffffffu, lane < warp ? v : 0);
            if (f) sel[cell_off[c] + counts[(size_t)c * nchunk + blockIdx.x] + before + __popc(m & ((1u << lane) - 1u))] = (int)i;
        }
        par ^= 1;       // two buffers: a warp reaches the next-but-one column only after everyone read this one
    }
}

// in place: counts[cell][:] -> exclusive scan over chunks; totals[cell] = members of the column
__global__ void __launch_bounds__(32) cell_scan_kernel(int *__restrict__ counts, int ncell, int nchunk, int *__restrict__ totals)
{
    const int c = blockIdx.x, lane = threadIdx.x;
    int *row = counts + (size_t)c * nchunk;
    int run = 0;
    for (int b = 0; b < nchunk; b += 32) {
        const int v = b + lane < nchunk ? row[b + lane] : 0;
        int s = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (b + lane < nchunk) row[b + lane] = run + s - v;
        run += __shfl_sync(0xffffffffu, s, 31);
    }
    if (lane == 0) totals[c] = run;
}

// one thread per output row (S3DISDataLoader.py:155-166)
__global__ void __launch_bounds__(256) scene_gather_kernel(const double *__restrict__ pts, int ld, int label_col,
                                                            const int *__restrict__ sel, const long long *__restrict__ cell_off,
                                                            const int *__restrict__ block_cell, const int *__restrict__ row_pos,
                                                            const double *__restrict__ centre, const double *__restrict__ room_max,
                                                            const float *__restrict__ labelweights, int ncls, long long rows,
                                                            int block_points, double *__restrict__ data, float *__restrict__ data32,
                                                            long long *__restrict__ label, double *__restrict__ smpw,
                                                            long long *__restrict__ index)
{
    const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    const int cell = block_cell[r / block_points];
    const int src = sel[cell_off[cell] + row_pos[r]];
    const double *p = pts + (long long)src * ld;
    const double x = p[0], y = p[1], z = p[2];
    double o[9];
    o[0] = x - centre[2 * cell];            // data_batch[:, 0] - (s_x + block_size / 2.0)
    o[1] = y - centre[2 * cell + 1];
    o[2] = z;
    o[3] = p[3] / 255.0; o[4] = p[4] / 255.0; o[5] = p[5] / 255.0;
    o[6] = x / room_max[0]; o[7] = y / room_max[1]; o[8] = z / room_max[2];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        if (data) data[r * 9 + c] = o[c];
        if (data32) data32[r * 9 + c] = (float)o[c];       // torch.Tensor(float64 ndarray): round to nearest
    }
    const long long lab = (long long)p[label_col];          // .astype(int): truncation
    if (label) label[r] = lab;
    if (smpw) smpw[r] = (lab >= 0 && lab < ncls) ? (double)labelweights[lab] : 0.0;
    if (index) index[r] = src;
}

}  // namespace

size_t psg_scene_minmax_ws_bytes() { return (size_t)296 * 6 * sizeof(double) + 16; }

int psg_scene_minmax_k(const double *pts, long long P, int ld, double *out6, void *ws, cudaStream_t st)
{
    double *partial = (double *)ws;
    unsigned *ticket = (unsigned *)((char *)ws + (size_t)296 * 6 * sizeof(double));
    long long want = (P + 255) / 256;
    const int grid = (int)(want < 296 ? want : 296);
    scene_minmax_kernel<<<grid, 256, 0, st>>>(pts, P, ld, partial, ticket, out6);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_scene_count_k(const double *pts, long long P, int ld, const double *bounds, int ncell, int *counts, int *totals,
                      cudaStream_t st)
{
    const int nchunk = (int)((P + kChunk - 1) / kChunk);
    cell_pass_kernel<false><<<nchunk, kChunk, 0, st>>>(pts, P, ld, (const double4 *)bounds, ncell, nchunk, counts, nullptr, nullptr);
    PSG_LAUNCH_CHECK();
    cell_scan_kernel<<<ncell, 32, 0, st>>>(counts, ncell, nchunk, totals);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_scene_fill_k(const double *pts, long long P, int ld, const double *bounds, int ncell, const int *counts,
                     const long long *cell_off, int *sel, cudaStream_t st)
{
    const int nchunk = (int)((P + kChunk - 1) / kChunk);
    cell_pass_kernel<true><<<nchunk, kChunk, 0, st>>>(pts, P, ld, (const double4 *)bounds, ncell, nchunk, const_cast<int *>(counts),
                                                      cell_off, sel);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_scene_gather_k(const double *pts, int ld, int label_col, const int *sel, const long long *cell_off, const int *block_cell,
                       const int *row_pos, const double *centre, const double *room_max, const float *labelweights, int ncls,
                       long long rows, int block_points, double *data, float *data32, long long *label, double *smpw,
                       long long *index, cudaStream_t st)
{
    scene_gather_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(pts, ld, label_col, sel, cell_off, block_cell, row_pos, centre,
                                                                       room_max, labelweights, ncls, rows, block_points, data, data32,
                                                                       label, smpw, index);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}
