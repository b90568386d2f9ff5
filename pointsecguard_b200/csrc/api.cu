// api.cu -- C-ABI entry points of the stand-alone primitives and T-layout building blocks
// (include/psg_b200.h).  Thin argument checks over the launchers in the other translation units.
#include "../../include/psg_b200.h"
#include "psg_common.cuh"
#include "psg_internal.h"

static inline TView mkview(const float *base, int wchunks, int c0) { return TView{(float *)base, wchunks, c0}; }

extern "C" size_t psg_fps_workspace(int P, int N) { return psg_fps_workspace_bytes(P, N); }

extern "C" int psg_fps(const float *xyz, int nclouds, int P, int N, int npoint, const int32_t *start, int32_t *out_idx,
                       float *out_xyz, void *workspace, size_t workspace_bytes, psg_stream_t stream)
{
    if (!xyz || !start || !out_idx || nclouds <= 0 || P <= 0 || N <= 0 || npoint <= 0) return PSG_EINVAL;
    return psg_fps_launch(xyz, (long long)N * 3, nclouds, P, N, npoint, start, out_idx, out_xyz, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

extern "C" int psg_square_distance(const float *src, const float *dst, int B, int N, int M, float *out,
                                   psg_stream_t stream)
{
    if (!src || !dst || !out) return PSG_EINVAL;
    return psg_square_distance_launch(src, dst, B, N, M, out, (cudaStream_t)stream);
}

extern "C" int psg_ball_query(const float *xyz, int nclouds, int P, int N, const float *new_xyz, int S, int nr,
                              const double *radius_host, const int *nsample_host, int32_t *out0, int32_t *out1,
                              psg_stream_t stream)
{
    if (!xyz || !new_xyz || !radius_host || !nsample_host || !out0 || (nr == 2 && !out1)) return PSG_EINVAL;
    for (int i = 0; i < nr && i < 2; ++i)
        if (nsample_host[i] <= 0) return PSG_EINVAL;
    return psg_ball_query_launch(xyz, (long long)N * 3, nclouds, P, N, new_xyz, S, nr, radius_host, nsample_host, out0,
                                 out1, (cudaStream_t)stream);
}

extern "C" size_t psg_ball_grid_workspace(int nclouds, int N) { return psg_ballgrid_workspace_bytes(nclouds, N); }

extern "C" int psg_ball_query_grid(const float *xyz, int nclouds, int P, int N, const float *new_xyz, int S, int nr,
                                   const double *radius_host, const int *nsample_host, int32_t *out0, int32_t *out1,
                                   void *workspace, size_t workspace_bytes, psg_stream_t stream)
{
    if (!xyz || !new_xyz || !radius_host || !nsample_host || !out0 || (nr == 2 && !out1) || !workspace) return PSG_EINVAL;
    if (nr < 1 || nr > 2) return PSG_EINVAL;
    for (int i = 0; i < nr; ++i)
        if (nsample_host[i] <= 0 || !(radius_host[i] > 0.0)) return PSG_EINVAL;
    if (workspace_bytes < psg_ballgrid_workspace_bytes(nclouds, N)) return PSG_EWORKSPACE;
    const double rmax = nr == 2 && radius_host[1] > radius_host[0] ? radius_host[1] : radius_host[0];
    int rc = psg_ballgrid_build(xyz, (long long)N * 3, nclouds, N, rmax, workspace, (cudaStream_t)stream);
    if (rc != PSG_OK) return rc;
    return psg_ballgrid_query(workspace, xyz, (long long)N * 3, nclouds, P, N, new_xyz, S, nr, radius_host, nsample_host, out0,
                              out1, (cudaStream_t)stream);
}

extern "C" int psg_three_nn(const float *xyz1, int nclouds1, int P, int N, const float *xyz2, int S, int32_t *idx,
                            float *w, float *d2, psg_stream_t stream)
{
    if (!xyz1 || !xyz2 || !idx || nclouds1 <= 0) return PSG_EINVAL;
    return psg_three_nn_launch(xyz1, (long long)N * 3, nclouds1, P, N, xyz2, S, idx, w, d2, (cudaStream_t)stream);
}

extern "C" int psg_index_points(const float *points, const int64_t *idx, int B, int N, int C, int64_t M, float *out,
                                psg_stream_t stream)
{
    if (!points || !idx || !out || B <= 0 || N <= 0 || C <= 0 || M <= 0) return PSG_EINVAL;
    return psg_index_points_rm(points, (const long long *)idx, B, N, C, M, out, (cudaStream_t)stream);
}

extern "C" int psg_pack_channels_first(const float *x, int64_t sb, int64_t sc, int64_t sn, int B, int C, int N,
                                       float *t_base, int t_wchunks, int t_c0, float *xyz_out, psg_stream_t stream)
{
    if (!x || !t_base || B <= 0 || C <= 0 || N <= 0 || (t_wchunks - t_c0) * 4 < C) return PSG_EINVAL;
    return psg_pack_cf(x, sb, sc, sn, B, C, N, mkview(t_base, t_wchunks, t_c0), round_up(C, 16), xyz_out,
                       (cudaStream_t)stream);
}

extern "C" int psg_unpack_channels_first(const float *t_base, int t_wchunks, int t_c0, int B, int C, int N, float *y,
                                         int accumulate, psg_stream_t stream)
{
    if (!t_base || !y || B <= 0 || C <= 0 || N <= 0) return PSG_EINVAL;
    return psg_unpack_cf(mkview(t_base, t_wchunks, t_c0), B, C, N, y, accumulate, (cudaStream_t)stream);
}

extern "C" int psg_group_points(const float *feats_base, int feats_wchunks, int D, const float *xyz, int nclouds, int Nsrc,
                                const float *new_xyz, const int32_t *idx, int P, int S, int K, float *out_base,
                                int out_cpad, psg_stream_t stream)
{
    if (!feats_base || !xyz || !new_xyz || !idx || !out_base || out_cpad % 16 || out_cpad < D + 3) return PSG_EINVAL;
    return psg_group(mkview(feats_base, feats_wchunks, 0), D, xyz, (long long)Nsrc * 3, nclouds, Nsrc, new_xyz, idx, P, S,
                     K, mkview(out_base, out_cpad / 4, 0), out_cpad, (cudaStream_t)stream);
}

extern "C" int psg_group_max(const float *in_base, int in_wchunks, int64_t groups, int K, int C, float *out_base,
                             int out_wchunks, int out_c0, uint8_t *argmax, psg_stream_t stream)
{
    if (!in_base || !out_base || !argmax || C % 4) return PSG_EINVAL;
    return psg_maxpool(mkview(in_base, in_wchunks, 0), groups, K, C, mkview(out_base, out_wchunks, out_c0), argmax,
                       (cudaStream_t)stream);
}

extern "C" int psg_group_max_backward(const float *dout_base, int dout_wchunks, int dout_c0, const float *out_base,
                                      int out_wchunks, int out_c0, const uint8_t *argmax, int64_t groups, int K, int C,
                                      float *dy_base, int dy_wchunks, psg_stream_t stream)
{
    if (!dout_base || !out_base || !argmax || !dy_base || C % 4) return PSG_EINVAL;
    return psg_maxpool_bwd(mkview(dout_base, dout_wchunks, dout_c0), mkview(out_base, out_wchunks, out_c0), argmax, groups,
                           K, C, mkview(dy_base, dy_wchunks, 0), (cudaStream_t)stream);
}

extern "C" int psg_interpolate(const float *feats_base, int feats_wchunks, int S, const int32_t *idx, const float *w,
                               int64_t P, int N, int ncols, float *out_base, int out_wchunks, int out_c0,
                               psg_stream_t stream)
{
    if (!feats_base || !idx || !w || !out_base || ncols % 4) return PSG_EINVAL;
    return psg_interp(mkview(feats_base, feats_wchunks, 0), S, idx, w, P, N, ncols / 4,
                      mkview(out_base, out_wchunks, out_c0), (cudaStream_t)stream);
}

extern "C" size_t psg_csr_workspace(int64_t P, int M, int R) { return psg_csr_scratch_bytes(P, M, R); }

extern "C" int psg_csr_build_by_source(const int32_t *keys, int64_t P, int M, int R, int pad_group, int32_t *offsets,
                                       int32_t *perm, void *workspace, psg_stream_t stream)
{
    if (!keys || !offsets || !perm || !workspace || P <= 0 || M <= 0 || R <= 0 || pad_group < 0) return PSG_EINVAL;
    return psg_csr_build(keys, P, M, R, pad_group, offsets, perm, workspace, (cudaStream_t)stream);
}

extern "C" int psg_segment_sum(const float *src_base, int src_wchunks, int src_c0, int64_t src_rows_per_problem, int div,
                               const float *weights, const int32_t *offsets, const int32_t *perm, int M, int R, int64_t P,
                               int ncols, float *dst_base, int dst_wchunks, int dst_c0, int accumulate,
                               psg_stream_t stream)
{
    if (!src_base || !offsets || !perm || !dst_base || div <= 0) return PSG_EINVAL;
    return psg_segsum(mkview(src_base, src_wchunks, src_c0), src_rows_per_problem, div, weights, offsets, perm, M, R, P,
                      ncols, mkview(dst_base, dst_wchunks, dst_c0), accumulate, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int psg_confusion_matrix(const float *logp, const int32_t *labels, const uint8_t *mask, int target,
                                    int64_t rows, int ncls, int64_t *conf, psg_stream_t stream)
{
    if (!logp || !labels || !conf || rows <= 0) return PSG_EINVAL;
    return psg_confusion(logp, labels, mask, target, rows, ncls, (long long *)conf, (cudaStream_t)stream);
}

extern "C" int psg_add_vote(const float *logp, const int64_t *point_idx, const float *weight, int64_t rows, int ncls,
                            float *pool, int64_t pool_rows, psg_stream_t stream)
{
    if (!logp || !point_idx || !pool || rows <= 0 || ncls < 1 || pool_rows <= 0) return PSG_EINVAL;
    return psg_add_vote_k(logp, (const long long *)point_idx, weight, rows, ncls, pool, pool_rows, (cudaStream_t)stream);
}

// ---- whole-scene block slicer (slicer.cu) ----
extern "C" size_t psg_scene_minmax_workspace(void) { return psg_scene_minmax_ws_bytes(); }

extern "C" int psg_scene_minmax(const double *points, int64_t P, int ld, double *out6, void *workspace, size_t workspace_bytes,
                                psg_stream_t stream)
{
    if (!points || !out6 || P <= 0 || ld < 3) return PSG_EINVAL;
    if (!workspace || workspace_bytes < psg_scene_minmax_ws_bytes()) return PSG_EWORKSPACE;
    return psg_scene_minmax_k(points, P, ld, out6, workspace, (cudaStream_t)stream);
}

extern "C" int64_t psg_scene_chunks(int64_t P) { return (P + 1023) / 1024; }

extern "C" int psg_scene_cell_counts(const double *points, int64_t P, int ld, const double *cell_bounds, int ncell,
                                     int32_t *counts, int32_t *totals, psg_stream_t stream)
{
    if (!points || !cell_bounds || !counts || !totals || P <= 0 || ld < 2 || ncell <= 0) return PSG_EINVAL;
    return psg_scene_count_k(points, P, ld, cell_bounds, ncell, counts, totals, (cudaStream_t)stream);
}

extern "C" int psg_scene_cell_fill(const double *points, int64_t P, int ld, const double *cell_bounds, int ncell,
                                   const int32_t *counts, const int64_t *cell_offset, int32_t *sel, psg_stream_t stream)
{
    if (!points || !cell_bounds || !counts || !cell_offset || !sel || P <= 0 || ld < 2 || ncell <= 0) return PSG_EINVAL;
    return psg_scene_fill_k(points, P, ld, cell_bounds, ncell, counts, (const long long *)cell_offset, sel, (cudaStream_t)stream);
}

extern "C" int psg_scene_gather(const double *points, int ld, int label_col, const int32_t *sel, const int64_t *cell_offset,
                                const int32_t *block_cell, const int32_t *row_pos, const double *centre, const double *room_max,
                                const float *labelweights, int ncls, int64_t rows, int block_points, double *data, float *data32,
                                int64_t *label, double *smpw, int64_t *index, psg_stream_t stream)
{
    if (!points || !sel || !cell_offset || !block_cell || !row_pos || !centre || !room_max || !labelweights) return PSG_EINVAL;
    if (ld < 6 || label_col < 0 || label_col >= ld || rows <= 0 || block_points <= 0 || ncls <= 0) return PSG_EINVAL;
    return psg_scene_gather_k(points, ld, label_col, sel, (const long long *)cell_offset, block_cell, row_pos, centre, room_max,
                              labelweights, ncls, rows, block_points, data, data32, (long long *)label, smpw, (long long *)index,
                              (cudaStream_t)stream);
}
