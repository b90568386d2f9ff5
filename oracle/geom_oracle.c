/*
 * oracle/geom_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Plain-C restatement of the geometric primitives on the PointNet++ sem-seg attack hot path of
 * C0ldstudy/PointSecGuard, written from the behaviour of the reference's Python
 * (PointNet/models/pointnet_util.py) with the floating-point operation order that the reference's
 * stock PyTorch-CPU execution produces (re-verified by oracle/make_golden.py
 * against the reference itself in the build container):
 *
 *   - farthest point sampling           pointnet_util.py:63-84
 *   - expansion-form squared distance   pointnet_util.py:19-40
 *   - ball query                        pointnet_util.py:87-107
 *   - 3 nearest neighbours + weights    pointnet_util.py:301-307
 *
 * Compile with -ffp-contract=off so that only the explicit fmaf() calls below fuse.
 * Parity status: PINNED -- checked bit-for-bit against tests/golden/geom_*.npz, which were produced
 * by importing and executing the unmodified reference (see oracle/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pointnet_util.py:80  dist = torch.sum((xyz - centroid) ** 2, -1)
 * torch.sum over the 3-wide last dim adds left to right; every product is rounded on its own. */
static inline float fps_dist(const float *p, const float *c)
{
    float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

/* pointnet_util.py:63-84.  xyz [B,N,3], start [B] (the torch.randint draw of line 75, made by the
 * caller so that the CPU generator stream stays the caller's), out [B,npoint]. */
void oracle_fps(const float *xyz, const int64_t *start, int B, int N, int npoint, int64_t *out)
{
    float *mind = (float *)malloc(sizeof(float) * (size_t)N);
    for (int b = 0; b < B; ++b) {
        const float *pts = xyz + (size_t)b * N * 3;
        for (int i = 0; i < N; ++i) mind[i] = 1e10f;          /* :74 */
        int64_t far = start[b];
        for (int it = 0; it < npoint; ++it) {
            out[(size_t)b * npoint + it] = far;                 /* :78, recorded before the update */
            const float *c = pts + far * 3;
            float best = -1.0f; int64_t besti = 0;
            for (int i = 0; i < N; ++i) {
                float d = fps_dist(pts + (size_t)i * 3, c);
                if (d < mind[i]) mind[i] = d;                   /* :81-82 */
                if (i == 0 || mind[i] > best) { best = mind[i]; besti = i; }   /* :83 first max wins */
            }
            far = besti;
        }
    }
    free(mind);
}

/* |p|^2 as torch.sum(p ** 2, -1) does it (pointnet_util.py:38-39). */
static inline float sqnorm(const float *p)
{
    float xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
    float s = xx + yy;
    return s + zz;
}

/* pointnet_util.py:37-39.  -2 * (src . dst) with the K=3 sgemm accumulating x, then y, then z with
 * fused multiply-adds; then "+= |src|^2", then "+= |dst|^2". */
static inline float sqdist(const float *s, float sn, const float *d, float dn)
{
    float dot = s[0] * d[0];
    dot = fmaf(s[1], d[1], dot);
    dot = fmaf(s[2], d[2], dot);
    float r = -2.0f * dot;
    r = r + sn;
    return r + dn;
}

/* square_distance(src [B,N,3], dst [B,M,3]) -> out [B,N,M]. */
void oracle_square_distance(const float *src, const float *dst, int B, int N, int M, float *out)
{
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            const float *s = src + ((size_t)b * N + i) * 3;
            float sn = sqnorm(s);
            for (int j = 0; j < M; ++j) {
                const float *d = dst + ((size_t)b * M + j) * 3;
                out[((size_t)b * N + i) * M + j] = sqdist(s, sn, d, sqnorm(d));
            }
        }
}

/* pointnet_util.py:87-107.  Rows of square_distance are the query centroids (src = new_xyz).
 * Keep indices whose d2 is NOT > r*r (:102), ascending (:103), first nsample, pad with the first
 * hit (:104-106).  radius*radius is evaluated in double by Python and compared against the float32
 * tensor, i.e. the tensor element is compared with the double threshold rounded to... PyTorch
 * compares in float32 after casting the Python scalar to the tensor dtype, so r2 is (float)(r*r).
 * A centroid with no hit at all (cannot happen when new_xyz is a subset of xyz unless d2 of the
 * point to itself rounds above r2) leaves N in every slot, exactly as the reference would. */
void oracle_ball_query(const float *xyz, const float *new_xyz, int B, int N, int S,
                       double radius, int nsample, int64_t *out)
{
    const float r2 = (float)(radius * radius);
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < S; ++s) {
            const float *q = new_xyz + ((size_t)b * S + s) * 3;
            float qn = sqnorm(q);
            int64_t *o = out + ((size_t)b * S + s) * nsample;
            int cnt = 0;
            for (int i = 0; i < N && cnt < nsample; ++i) {
                const float *p = xyz + ((size_t)b * N + i) * 3;
                float d = sqdist(q, qn, p, sqnorm(p));
                if (!(d > r2)) o[cnt++] = i;
            }
            int64_t first = cnt ? o[0] : (int64_t)N;
            for (; cnt < nsample; ++cnt) o[cnt] = first;
        }
}

/* pointnet_util.py:301-307.  Rows are the fine points (src = xyz1), columns the coarse points
 * (dst = xyz2); the three smallest d2 in ascending order, equal keys keep ascending column index
 * (stable sort).  weight = (1/(d2+1e-8)) / sum.  Needs S >= 3. */
void oracle_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S,
                     int64_t *idx, float *d2, float *w)
{
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            const float *q = xyz1 + ((size_t)b * N + i) * 3;
            float qn = sqnorm(q);
            float bd[3] = {INFINITY, INFINITY, INFINITY};
            int64_t bi[3] = {-1, -1, -1};
            for (int j = 0; j < S; ++j) {
                const float *p = xyz2 + ((size_t)b * S + j) * 3;
                float d = sqdist(q, qn, p, sqnorm(p));
                if (d < bd[2] || bi[2] < 0) {
                    int k = 2;
                    while (k > 0 && (bi[k - 1] < 0 || d < bd[k - 1])) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; --k; }
                    bd[k] = d; bi[k] = j;
                }
            }
            size_t o = ((size_t)b * N + i) * 3;
            float r0 = 1.0f / (bd[0] + 1e-8f), r1 = 1.0f / (bd[1] + 1e-8f), r2 = 1.0f / (bd[2] + 1e-8f);
            float nrm = (r0 + r1) + r2;                         /* torch.sum over 3 elements, in order */
            for (int k = 0; k < 3; ++k) { idx[o + k] = bi[k]; d2[o + k] = bd[k]; }
            w[o + 0] = r0 / nrm; w[o + 1] = r1 / nrm; w[o + 2] = r2 / nrm;
        }
}
