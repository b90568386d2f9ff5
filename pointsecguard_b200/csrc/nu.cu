// nu.cu -- kernels of the norm-unbounded (C&W-style) colour attacks.
// Reference: PointNet/attacks/torchattacks/attacks/nontarget.py:52-135 (NU_attack) and
// target.py:62-175 (tar_NU_attack):
//   w = atanh(2 c - 1);  per step:  c = (tanh(w) + 1) / 2 -> image -> model -> cost
//   cost = sum clamp(p_y - max_other p, -kappa) + c_ * sum(k smallest colour distances, block 0)
//          + c_ * sum (adv - image)^2 ;  Adam on w ; early exit on an accuracy threshold.
//
// One step is a fixed launch sequence with no host round trip: the accuracy test of the reference
// (`.item()` every step) is evaluated on the device by nu_reduce_kernel, which latches a `done`
// flag; once it is set the two kernels that change attack state (build_adv, adam) become no-ops,
// so the returned image is exactly the image of the step that triggered the exit.
#include "psg_common.cuh"
#include "psg_internal.h"

namespace {

inline unsigned nb(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// w = 0.5 * log((1 + x) / (1 - x)), x = 2 c - 1   (nontarget.py:111-117; +-inf at c = 0 / 1, Q6)
// (general field: channel j of the field lives in the box [lo_j, hi_j]; the reference's colours are [0, 1])
__global__ void nu_init_kernel(const float *__restrict__ images, int B, int C, int N, PsgNuField fld, float *__restrict__ w,
                               float *__restrict__ m, float *__restrict__ v)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * fld.nc * N) return;
    const int n = (int)(t % N);
    const int j = (int)((t / N) % fld.nc);
    const int b = (int)(t / ((long long)fld.nc * N));
    float c = images[((long long)b * C + fld.c0 + j) * N + n];
    if (fld.lo[j] != 0.f || fld.hi[j] != 1.f) c = (c - fld.lo[j]) / (fld.hi[j] - fld.lo[j]);
    const float x = __fsub_rn(__fmul_rn(c, 2.f), 1.f);
    w[t] = 0.5f * logf(__fdiv_rn(__fadd_rn(1.f, x), __fsub_rn(1.f, x)));
    m[t] = 0.f;
    v[t] = 0.f;
}

// adv = base with the (masked) colours replaced by (tanh(w) + 1) / 2; also the model input
// (feats0 colour columns) and the per-point L2 term sum_ch (adv - image)^2 (nontarget.py:73-78).
__global__ void nu_build_adv_kernel(const float *__restrict__ w, const float *__restrict__ base,
                                    const float *__restrict__ images, const unsigned char *__restrict__ mask,
                                    int B, int C, int N, PsgNuField fld, TView feats0, float *__restrict__ adv,
                                    float *__restrict__ l2_rows, const int *__restrict__ status)
{
    if (status[0]) return;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * N) return;
    const int b = (int)(row / N), n = (int)(row % N);
    const bool on = mask ? mask[row] != 0 : true;
    float l2 = 0.f;
    for (int ch = 0; ch < C; ++ch) {
        const long long o = ((long long)b * C + ch) * N + n;
        float val = base[o];
        if (on && ch >= fld.c0 && ch < fld.c0 + fld.nc) {
            const int j = ch - fld.c0;
            val = 0.5f * (tanhf(w[((long long)b * fld.nc + j) * N + n]) + 1.f);
            if (fld.lo[j] != 0.f || fld.hi[j] != 1.f) val = fld.lo[j] + (fld.hi[j] - fld.lo[j]) * val;
        }
        adv[o] = val;
        feats0.base[tv_off(feats0, row, ch >> 2) + (ch & 3)] = val;
        const float d = val - images[o];
        l2 = fmaf(d, d, l2);
    }
    l2_rows[row] = l2;
}

// Smoothness term (nontarget.py:130-135, Q8): for every point i of block 0 the K smallest
// distances between its adversarial colour and the ORIGINAL colours of all points of the block
// (cdist's matmul form: sqrt(max(0, -2 a.b + |a|^2 + |b|^2))), their sum, and the gradient
// sum_k (a_i - p_jk) / d_k with 0 at d = 0 (_euclidean_dist_backward).
// TPR lanes share one row: lane q scans candidates q, q + TPR, ... and keeps its own ascending top-K (strict '<':
// among equal distances the lower index stays first), then the TPR lists are merged by (distance, index), which is
// exactly the order the one-thread scan of all candidates produces -- sums and gradients are accumulated in that
// order, so the result does not depend on TPR.  (One thread per row left 32 CTAs of serial 4096-candidate scans:
// 0.45 ms of a 1.9 ms step; the sorted insertion runs for the whole warp whenever one lane needs it.)
constexpr int kSmoothTPR = 8;
template <int K>
__global__ void __launch_bounds__(256)
nu_smooth_kernel(const float *__restrict__ adv, const float *__restrict__ images, int C, int N,
                 float *__restrict__ rows_out, float *__restrict__ grad_out)
{
    constexpr int TPR = kSmoothTPR;
    extern __shared__ float sm[];
    float *px = sm, *py = sm + N, *pz = sm + 2 * N, *pn = sm + 3 * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float x = images[(3LL) * N + j], y = images[(4LL) * N + j], z = images[(5LL) * N + j];
        px[j] = x; py[j] = y; pz[j] = z;
        pn[j] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    }
    __syncthreads();
    const int q = threadIdx.x % TPR;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / TPR;          // row; whole TPR-groups are in or out together
    const bool live = i < N;
    const int ii = live ? i : N - 1;
    const float ax = adv[3LL * N + ii], ay = adv[4LL * N + ii], az = adv[5LL * N + ii];
    const float an = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
    const float mx = -2.f * ax, my = -2.f * ay, mz = -2.f * az;
    float best[K]; int bj[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { best[k] = INFINITY; bj[k] = 0x7fffffff; }
    for (int j = q; j < N; j += TPR) {
        float acc = __fmul_rn(mx, px[j]);
        acc = __fmaf_rn(my, py[j], acc);
        acc = __fmaf_rn(mz, pz[j], acc);
        acc = __fadd_rn(acc, an);
        acc = __fadd_rn(acc, pn[j]);
        float d2 = fmaxf(acc, 0.f);
        if (d2 < best[K - 1]) {
            // insert keeping ascending order; strict '<' keeps the lower index first among ties
            float cv = d2; int cj = j;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (cv < best[k]) { float tv_ = best[k]; int tj = bj[k]; best[k] = cv; bj[k] = cj; cv = tv_; cj = tj; }
            }
        }
    }
    // ---- K-way merge of the TPR lists, smallest (distance, index) first; every lane of the group accumulates ----
    float s = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        float v = best[0]; int j = bj[0];
#pragma unroll
        for (int o = 1; o < TPR; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oj = __shfl_xor_sync(0xffffffffu, j, o);
            if (ov < v || (ov == v && oj < j)) { v = ov; j = oj; }
        }
        if (j == bj[0] && v == best[0]) {                 // this lane's head won: pop it (indices are unique per lane set)
#pragma unroll
            for (int t = 0; t + 1 < K; ++t) { best[t] = best[t + 1]; bj[t] = bj[t + 1]; }
            best[K - 1] = INFINITY; bj[K - 1] = 0x7fffffff;
        }
        if (j == 0x7fffffff) continue;                    // fewer than K candidates in the block
        const float d = sqrtf(v);
        s += d;
        if (d > 0.f) {
            const float r = 1.f / d;
            gx += (ax - px[j]) * r; gy += (ay - py[j]) * r; gz += (az - pz[j]) * r;
        }
    }
    if (live && q == 0) {
        rows_out[i] = s;
        grad_out[i] = gx; grad_out[N + i] = gy; grad_out[2 * N + i] = gz;
    }
}

// One CTA: cost[step] = sum f + c * sum smooth + c * sum L2 (fixed summation order, double
// accumulators), the hit count of the accuracy test, and the early-exit latch
// (nontarget.py:86-96, target.py:96-121).
//   status[0] done flag, status[1] step at which it was set, status[2] last hit count
__global__ void __launch_bounds__(1024)
nu_reduce_kernel(const float *__restrict__ f_rows, const float *__restrict__ l2_rows,
                 const float *__restrict__ smooth_rows, const unsigned char *__restrict__ hit,
                 const unsigned char *__restrict__ mask, long long rows, int nsmooth, float c, int step,
                 double acc_denom, double thr, int exit_above, int count_masked_only,
                 float *__restrict__ cost, int *__restrict__ status)
{
    __shared__ double sf[32], sl[32], ss[32];
    __shared__ unsigned int sc[32];
    if (status[0]) return;
    double f = 0.0, l = 0.0, s = 0.0;
    unsigned int cnt = 0;
    // one CTA sums everything in a fixed order; the loads of eight strides are issued before the first is consumed
    // (one row per L2 round trip made this kernel 0.2 ms at 32 x 4096 rows), the additions keep their order
    constexpr int U = 8;
    for (long long r0 = threadIdx.x; r0 < rows; r0 += (long long)U * blockDim.x) {
        float fv[U], lv[U]; unsigned char hv[U], mv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + (long long)u * blockDim.x;
            const bool in = r < rows;
            fv[u] = in ? f_rows[r] : 0.f;
            lv[u] = in ? l2_rows[r] : 0.f;
            hv[u] = in ? hit[r] : (unsigned char)0;
            mv[u] = (in && count_masked_only) ? (mask ? mask[r] : (unsigned char)0) : (unsigned char)1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r0 + (long long)u * blockDim.x >= rows) break;
            f += (double)fv[u];
            l += (double)lv[u];
            if (mv[u] && hv[u]) ++cnt;
        }
    }
    for (int r0 = threadIdx.x; r0 < nsmooth; r0 += U * blockDim.x) {
        float sv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) sv[u] = r0 + u * (int)blockDim.x < nsmooth ? smooth_rows[r0 + u * blockDim.x] : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r0 + u * (int)blockDim.x < nsmooth) s += (double)sv[u];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        f += __shfl_down_sync(0xffffffffu, f, o);
        l += __shfl_down_sync(0xffffffffu, l, o);
        s += __shfl_down_sync(0xffffffffu, s, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { sf[warp] = f; sl[warp] = l; ss[warp] = s; sc[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double F = 0.0, Lq = 0.0, Sq = 0.0; unsigned int Cn = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { F += sf[i]; Lq += sl[i]; Sq += ss[i]; Cn += sc[i]; }
        // float arithmetic of the reference: cost = f + c * S + c * L2 (each sum a float32 scalar)
        const float cs = __fmul_rn(c, (float)Sq), cl = __fmul_rn(c, (float)Lq);
        cost[step] = __fadd_rn(__fadd_rn((float)F, cs), cl);
        status[2] = (int)Cn;
        const double acc = (double)Cn / acc_denom;
        const bool ex = exit_above ? (acc > thr) : (acc < thr);
        if (ex) { status[0] = 1; status[1] = step; }
    }
}

// gradient assembly + torch.optim.Adam step on w (nontarget.py:89-91):
//   g_c = dcost/dcolour = model grad + c * 2 (colour - image colour) + [block 0] c * smooth grad
//   g_w = g_c * (1 - tanh(w)^2) / 2
//   m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g^2; w -= step_size * m / (sqrt(v) / bc2_sqrt + eps)
__global__ void nu_adam_kernel(float *__restrict__ w, float *__restrict__ m, float *__restrict__ v, TView grad0,
                               const float *__restrict__ adv, const float *__restrict__ images,
                               const float *__restrict__ smooth_grad, const unsigned char *__restrict__ mask, int B,
                               int C, int N, PsgNuField fld, float c, float step_size, float bc2_sqrt, float beta1, float beta2,
                               float eps, int reset, const int *__restrict__ status)
{
    if (status[0]) return;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (long long)B * N) return;
    if (mask && !mask[row]) return;
    const int b = (int)(row / N), n = (int)(row % N);
    for (int j = 0; j < fld.nc; ++j) {
        const int ch = fld.c0 + j;
        const long long wi = ((long long)b * fld.nc + j) * N + n;
        const long long o = ((long long)b * C + ch) * N + n;
        float g = grad0.base[tv_off(grad0, row, ch >> 2) + (ch & 3)];
        g = fmaf(c, 2.f * (adv[o] - images[o]), g);
        if (b == 0 && ch >= 3 && ch < 6) g = fmaf(c, smooth_grad[(long long)(ch - 3) * N + n], g);
        const float wv = w[wi];
        const float t = tanhf(wv);
        float gw = g * (0.5f * (1.f - t * t));
        if (fld.lo[j] != 0.f || fld.hi[j] != 1.f) gw *= (fld.hi[j] - fld.lo[j]);
        float mm = reset ? 0.f : m[wi], vv = reset ? 0.f : v[wi];
        mm = mm + (gw - mm) * (1.f - beta1);
        vv = vv * beta2 + (1.f - beta2) * gw * gw;
        m[wi] = mm; v[wi] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        w[wi] = wv - step_size * (mm / denom);
    }
}

// image = clamp(image, lo, hi) on every channel (target.py:132, Q4)
__global__ void clamp_kernel(float *__restrict__ x, long long n, float lo, float hi)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) x[t] = fminf(fmaxf(x[t], lo), hi);
}

}  // namespace

int psg_nu_init_k(const float *images, int B, int C, int N, PsgNuField fld, float *w, float *m, float *v, cudaStream_t st)
{
    nu_init_kernel<<<nb((long long)B * fld.nc * N, 256), 256, 0, st>>>(images, B, C, N, fld, w, m, v);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_nu_build_adv_k(const float *w, const float *base, const float *images, const unsigned char *mask, int B, int C,
                     int N, PsgNuField fld, TView feats0, float *adv, float *l2_rows, const int *status, cudaStream_t st)
{
    nu_build_adv_kernel<<<nb((long long)B * N, 256), 256, 0, st>>>(w, base, images, mask, B, C, N, fld, feats0, adv, l2_rows,
                                                                  status);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_nu_smooth_k(const float *adv0, const float *images0, int C, int N, int k, float *rows_out, float *grad_out,
                  cudaStream_t st)
{
    const size_t smem = (size_t)4 * N * sizeof(float);
    if (smem > 200 * 1024) return PSG_EUNSUPPORTED;
    static PsgDeviceOnce attr_once;
    if (attr_once.need()) {
        if (cudaFuncSetAttribute(nu_smooth_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(nu_smooth_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return PSG_ECUDA;
        attr_once.mark();
    }
    const unsigned grid = nb((long long)N * kSmoothTPR, 256);
    if (k == 5) nu_smooth_kernel<5><<<grid, 256, smem, st>>>(adv0, images0, C, N, rows_out, grad_out);
    else if (k == 10) nu_smooth_kernel<10><<<grid, 256, smem, st>>>(adv0, images0, C, N, rows_out, grad_out);
    else return PSG_EUNSUPPORTED;
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_nu_reduce_k(const float *f_rows, const float *l2_rows, const float *smooth_rows, const unsigned char *hit,
                  const unsigned char *mask, long long rows, int nsmooth, float c, int step, double acc_denom, double thr,
                  int exit_above, int count_masked_only, float *cost, int *status, cudaStream_t st)
{
    nu_reduce_kernel<<<1, 1024, 0, st>>>(f_rows, l2_rows, smooth_rows, hit, mask, rows, nsmooth, c, step, acc_denom, thr,
                                         exit_above, count_masked_only, cost, status);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_nu_adam_k(float *w, float *m, float *v, TView grad0, const float *adv, const float *images,
                const float *smooth_grad, const unsigned char *mask, int B, int C, int N, PsgNuField fld, float c, float step_size,
                float bc2_sqrt, float beta1, float beta2, float eps, int reset, const int *status, cudaStream_t st)
{
    nu_adam_kernel<<<nb((long long)B * N, 256), 256, 0, st>>>(w, m, v, grad0, adv, images, smooth_grad, mask, B, C, N, fld, c,
                                                            step_size, bc2_sqrt, beta1, beta2, eps, reset, status);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}

int psg_clamp_k(float *x, long long n, float lo, float hi, cudaStream_t st)
{
    clamp_kernel<<<nb(n, 256), 256, 0, st>>>(x, n, lo, hi);
    PSG_LAUNCH_CHECK();
    return PSG_OK;
}
