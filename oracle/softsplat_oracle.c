/*
 * oracle/softsplat_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Sequential CPU restatement of the DiffCodec motion-compensation hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (CUDA) path never does.
 *
 * Every function cites the reference lines it restates
 * (paths relative to the upstream repository root):
 *
 *   orc_splat_fwd_*      controlnet/softsplat.py:290-335  (kernel softsplat_out)
 *   orc_splat_ingrad_*   controlnet/softsplat.py:376-423  (kernel softsplat_ingrad)
 *   orc_splat_flowgrad_* controlnet/softsplat.py:447-512  (kernel softsplat_flowgrad)
 *   orc_backwarp_*       cmp/models/modules/warp.py:9-25  (grid_sample, bilinear,
 *                        zeros padding; both align_corners conventions)
 *
 * Canonical summation order (what the optional deterministic CUDA mode must
 * reproduce bit-for-bit): the linear element index runs n, c, y, x exactly as
 * the reference's grid-stride loop enumerates it (softsplat.py:291-294), and
 * for each element the four corners are visited NW, NE, SW, SE
 * (softsplat.py:320-334). All tensors are dense NCHW.
 *
 * Rounding model: compiled with -ffp-contract=off. The forward kernel has no
 * contractible multiply-add (the product is rounded, then atomically added).
 * The two gather kernels are compiled by NVRTC with its default -fmad=true, so
 * `acc += g * w` contracts to one fma; we spell that as fma()/fmaf().
 *
 * Parity pin: the reference ships no test or golden vector for this path
 * (SURVEY.md section 4). This restatement is pinned instead against the
 * reference's own kernel text executed sequentially on the CPU
 * (oracle/ref_emulation.py -> tests/golden/ref_emu_*.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* float -> int conversion with CUDA cvt.rzi.s32 semantics (saturating; the
 * reference's `(int) floor(x)` on device). NaN never reaches here. */
static inline int sat_int_d(double v) {
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (-2147483647 - 1);
    return (int)v;
}

#define DEFINE_ORACLE(T, SUF, FLOOR, FMA, ISFINITE)                                                \
                                                                                                   \
/* softsplat.py:290-335. out must be zero-initialised by the caller (softsplat.py:281). */        \
void orc_splat_fwd_##SUF(const T* in, const T* flow, T* out, int N, int C, int H, int W) {         \
    const size_t plane = (size_t)H * (size_t)W;                                                    \
    for (int n = 0; n < N; ++n)                                                                    \
    for (int c = 0; c < C; ++c)                                                                    \
    for (int y = 0; y < H; ++y)                                                                    \
    for (int x = 0; x < W; ++x) {                                                                  \
        const T fx = (T)x + flow[((size_t)n * 2 + 0) * plane + (size_t)y * W + x];                 \
        const T fy = (T)y + flow[((size_t)n * 2 + 1) * plane + (size_t)y * W + x];                 \
        if (!ISFINITE(fx) || !ISFINITE(fy)) continue;            /* :301-302 */                    \
        const T v = in[((size_t)n * C + c) * plane + (size_t)y * W + x];                           \
        const int nwx = sat_int_d((double)FLOOR(fx)), nwy = sat_int_d((double)FLOOR(fy));          \
        /* wrap-around like the device's int +1 (INT_MAX + 1 is UB in C; use unsigned) */          \
        const int sex = (int)((unsigned)nwx + 1u), sey = (int)((unsigned)nwy + 1u);                \
        const T wnw = ((T)sex - fx) * ((T)sey - fy);             /* :315 */                        \
        const T wne = (fx - (T)nwx) * ((T)sey - fy);             /* :316 */                        \
        const T wsw = ((T)sex - fx) * (fy - (T)nwy);             /* :317 */                        \
        const T wse = (fx - (T)nwx) * (fy - (T)nwy);             /* :318 */                        \
        T* o = out + ((size_t)n * C + c) * plane;                                                  \
        if (nwx >= 0 && nwx < W && nwy >= 0 && nwy < H) o[(size_t)nwy * W + nwx] += v * wnw;       \
        if (sex >= 0 && sex < W && nwy >= 0 && nwy < H) o[(size_t)nwy * W + sex] += v * wne;       \
        if (nwx >= 0 && nwx < W && sey >= 0 && sey < H) o[(size_t)sey * W + nwx] += v * wsw;       \
        if (sex >= 0 && sex < W && sey >= 0 && sey < H) o[(size_t)sey * W + sex] += v * wse;       \
    }                                                                                              \
}                                                                                                  \
                                                                                                   \
/* softsplat.py:376-423. ingrad must be zero-initialised (softsplat.py:364): elements whose    */  \
/* flow is non-finite are left untouched by the kernel (early return at :389-390).             */  \
void orc_splat_ingrad_##SUF(const T* flow, const T* outgrad, T* ingrad,                            \
                            int N, int C, int H, int W) {                                          \
    const size_t plane = (size_t)H * (size_t)W;                                                    \
    for (int n = 0; n < N; ++n)                                                                    \
    for (int c = 0; c < C; ++c)                                                                    \
    for (int y = 0; y < H; ++y)                                                                    \
    for (int x = 0; x < W; ++x) {                                                                  \
        const T fx = (T)x + flow[((size_t)n * 2 + 0) * plane + (size_t)y * W + x];                 \
        const T fy = (T)y + flow[((size_t)n * 2 + 1) * plane + (size_t)y * W + x];                 \
        if (!ISFINITE(fx) || !ISFINITE(fy)) continue;                                              \
        const int nwx = sat_int_d((double)FLOOR(fx)), nwy = sat_int_d((double)FLOOR(fy));          \
        const int sex = (int)((unsigned)nwx + 1u), sey = (int)((unsigned)nwy + 1u);                \
        const T wnw = ((T)sex - fx) * ((T)sey - fy);                                               \
        const T wne = (fx - (T)nwx) * ((T)sey - fy);                                               \
        const T wsw = ((T)sex - fx) * (fy - (T)nwy);                                               \
        const T wse = (fx - (T)nwx) * (fy - (T)nwy);                                               \
        const T* g = outgrad + ((size_t)n * C + c) * plane;                                        \
        T acc = (T)0;                                                                              \
        if (nwx >= 0 && nwx < W && nwy >= 0 && nwy < H) acc = FMA(g[(size_t)nwy * W + nwx], wnw, acc); \
        if (sex >= 0 && sex < W && nwy >= 0 && nwy < H) acc = FMA(g[(size_t)nwy * W + sex], wne, acc); \
        if (nwx >= 0 && nwx < W && sey >= 0 && sey < H) acc = FMA(g[(size_t)sey * W + nwx], wsw, acc); \
        if (sex >= 0 && sex < W && sey >= 0 && sey < H) acc = FMA(g[(size_t)sey * W + sex], wse, acc); \
        ingrad[((size_t)n * C + c) * plane + (size_t)y * W + x] = acc;                             \
    }                                                                                              \
}                                                                                                  \
                                                                                                   \
/* softsplat.py:447-512. flowgrad must be zero-initialised (softsplat.py:365). */                  \
void orc_splat_flowgrad_##SUF(const T* in, const T* flow, const T* outgrad, T* flowgrad,           \
                              int N, int C, int H, int W) {                                        \
    const size_t plane = (size_t)H * (size_t)W;                                                    \
    for (int n = 0; n < N; ++n)                                                                    \
    for (int d = 0; d < 2; ++d)                                                                    \
    for (int y = 0; y < H; ++y)                                                                    \
    for (int x = 0; x < W; ++x) {                                                                  \
        const T fx = (T)x + flow[((size_t)n * 2 + 0) * plane + (size_t)y * W + x];                 \
        const T fy = (T)y + flow[((size_t)n * 2 + 1) * plane + (size_t)y * W + x];                 \
        if (!ISFINITE(fx) || !ISFINITE(fy)) continue;                                              \
        const int nwx = sat_int_d((double)FLOOR(fx)), nwy = sat_int_d((double)FLOOR(fy));          \
        const int sex = (int)((unsigned)nwx + 1u), sey = (int)((unsigned)nwy + 1u);                \
        T dnw, dne, dsw, dse;                                                                      \
        if (d == 0) {                                            /* :477-481 */                    \
            dnw = ((T)-1) * ((T)sey - fy);                                                         \
            dne = ((T)+1) * ((T)sey - fy);                                                         \
            dsw = ((T)-1) * (fy - (T)nwy);                                                         \
            dse = ((T)+1) * (fy - (T)nwy);                                                         \
        } else {                                                 /* :483-487 */                    \
            dnw = ((T)sex - fx) * ((T)-1);                                                         \
            dne = (fx - (T)nwx) * ((T)-1);                                                         \
            dsw = ((T)sex - fx) * ((T)+1);                                                         \
            dse = (fx - (T)nwx) * ((T)+1);                                                         \
        }                                                                                          \
        T acc = (T)0;                                                                              \
        for (int c = 0; c < C; ++c) {                            /* :491-509 */                    \
            const T v = in[((size_t)n * C + c) * plane + (size_t)y * W + x];                       \
            const T* g = outgrad + ((size_t)n * C + c) * plane;                                    \
            if (nwx >= 0 && nwx < W && nwy >= 0 && nwy < H) acc = FMA(g[(size_t)nwy * W + nwx] * v, dnw, acc); \
            if (sex >= 0 && sex < W && nwy >= 0 && nwy < H) acc = FMA(g[(size_t)nwy * W + sex] * v, dne, acc); \
            if (nwx >= 0 && nwx < W && sey >= 0 && sey < H) acc = FMA(g[(size_t)sey * W + nwx] * v, dsw, acc); \
            if (sex >= 0 && sex < W && sey >= 0 && sey < H) acc = FMA(g[(size_t)sey * W + sex] * v, dse, acc); \
        }                                                                                          \
        flowgrad[((size_t)n * 2 + d) * plane + (size_t)y * W + x] = acc;                           \
    }                                                                                              \
}                                                                                                  \
                                                                                                   \
/* cmp/models/modules/warp.py:9-25 as executed by F.grid_sample(bilinear, zeros padding).      */  \
/* The layer builds grid = linspace(-1,1,W)[x] + flow_x/((W-1)/2) (and likewise y); grid_sample */  \
/* un-normalises with align_corners=False (torch default since 1.3): s = ((g+1)*W-1)/2, or     */  \
/* with align_corners=True: s = (g+1)/2*(W-1). We take the already-built grid value so that the */  \
/* float rounding of the grid construction is done by the caller exactly as the layer does it. */  \
void orc_grid_sample_##SUF(const T* image, const T* grid /* [N,H,W,2] */, T* out,                  \
                           int N, int C, int H, int W, int align_corners) {                        \
    const size_t plane = (size_t)H * (size_t)W;                                                    \
    for (int n = 0; n < N; ++n)                                                                    \
    for (int y = 0; y < H; ++y)                                                                    \
    for (int x = 0; x < W; ++x) {                                                                  \
        const T gx = grid[(((size_t)n * H + y) * W + x) * 2 + 0];                                  \
        const T gy = grid[(((size_t)n * H + y) * W + x) * 2 + 1];                                  \
        T sx, sy;                                                                                  \
        if (align_corners) { sx = ((gx + (T)1) / (T)2) * (T)(W - 1); sy = ((gy + (T)1) / (T)2) * (T)(H - 1); } \
        else { sx = ((gx + (T)1) * (T)W - (T)1) / (T)2; sy = ((gy + (T)1) * (T)H - (T)1) / (T)2; } \
        const T x0f = FLOOR(sx), y0f = FLOOR(sy);                                                  \
        const T ax = sx - x0f, ay = sy - y0f;                                                      \
        const int x0 = sat_int_d((double)x0f), y0 = sat_int_d((double)y0f);                        \
        const int x1 = (int)((unsigned)x0 + 1u), y1 = (int)((unsigned)y0 + 1u);                    \
        const T wnw = ((T)1 - ax) * ((T)1 - ay), wne = ax * ((T)1 - ay);                           \
        const T wsw = ((T)1 - ax) * ay, wse = ax * ay;                                             \
        const int ok = ISFINITE(sx) && ISFINITE(sy);                                               \
        for (int c = 0; c < C; ++c) {                                                              \
            const T* im = image + ((size_t)n * C + c) * plane;                                     \
            T acc = (T)0;                                                                          \
            if (ok) {                                                                              \
                if (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H) acc += im[(size_t)y0 * W + x0] * wnw;  \
                if (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H) acc += im[(size_t)y0 * W + x1] * wne;  \
                if (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H) acc += im[(size_t)y1 * W + x0] * wsw;  \
                if (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H) acc += im[(size_t)y1 * W + x1] * wse;  \
            }                                                                                      \
            out[((size_t)n * C + c) * plane + (size_t)y * W + x] = acc;                            \
        }                                                                                          \
    }                                                                                              \
}

#define ISFINITE_F(x) (isfinite(x))
DEFINE_ORACLE(float, f32, floorf, fmaf, ISFINITE_F)
DEFINE_ORACLE(double, f64, floor, fma, ISFINITE_F)

/*
 * Frame-parallel forward for the CPU baseline (bench.py cpu_baseline / --impl
 * reference): identical arithmetic and per-frame order to orc_splat_fwd_f32;
 * frames are independent (a splat never crosses n), so threading over n keeps
 * every output bit-identical to the sequential call. Plain pthreads (this
 * image's gcc has no libgomp).
 */
typedef struct {
    const float* in; const float* flow; float* out;
    int N, C, H, W, tid, nthreads;
} orc_mt_job;

static void* orc_mt_worker(void* arg) {
    const orc_mt_job* j = (const orc_mt_job*)arg;
    const size_t plane = (size_t)j->H * (size_t)j->W;
    for (int n = j->tid; n < j->N; n += j->nthreads)
        orc_splat_fwd_f32(j->in + (size_t)n * j->C * plane, j->flow + (size_t)n * 2 * plane,
                          j->out + (size_t)n * j->C * plane, 1, j->C, j->H, j->W);
    return NULL;
}

/* exp() of the deterministic mode: the same IEEE double operations, in the same order, as exp_det() in the CUDA
 * library (csrc/dcb_common.cuh): both sides round (float)exp_det((double)m) identically. Compiled with
 * -ffp-contract=off; every fused operation is an explicit fma(). */
double orc_exp_det(double x) {
    if (!(x == x)) return x;
    if (x > 709.0) return INFINITY;
    if (x < -745.0) return 0.0;
    const double k = rint(x * 1.4426950408889634);
    double r = fma(-k, 6.93147180369123816490e-01, x);
    r = fma(-k, 1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return scalbn(p, (int)k);
}

void orc_exp_det_f32(const float* m, float* out, long long n) {
    for (long long i = 0; i < n; ++i) out[i] = (float)orc_exp_det((double)m[i]);
}

void orc_exp_det_f64(const double* m, double* out, long long n) {
    for (long long i = 0; i < n; ++i) out[i] = orc_exp_det(m[i]);
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

void orc_splat_fwd_f32_mt(const float* in, const float* flow, float* out,
                          int N, int C, int H, int W, int threads) {
    if (threads <= 0) threads = orc_max_threads();
    if (threads > N) threads = N;
    if (threads > 256) threads = 256;
    if (threads <= 1) { orc_splat_fwd_f32(in, flow, out, N, C, H, W); return; }
    pthread_t th[256];
    orc_mt_job jobs[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (orc_mt_job){in, flow, out, N, C, H, W, t, threads};
        pthread_create(&th[t], NULL, orc_mt_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
