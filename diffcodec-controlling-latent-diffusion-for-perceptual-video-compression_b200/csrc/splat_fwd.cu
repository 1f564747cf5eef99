// splat_fwd.cu -- forward splatting (K1 scatter + K2 normalise) for sm_100a.
//
// Replaces controlnet/softsplat.py:232-274 (mode wrapper: cat / exp / mul pre-ops, eps + divide
// post-ops) and :277-355 (softsplat_func.forward, kernel `softsplat_out`) of the reference.
//
// Design (not a port of the reference's one-thread-per-element kernel):
//   * one thread per SOURCE PIXEL: flow is read once and the four bilinear weights are computed
//     once per pixel (the reference recomputes them for every channel);
//   * the appended channel (1 | m | exp(m)) and the in*g(m) product live in registers: the
//     concatenated tensor of softsplat.py:241-247 is never materialised;
//   * C+1 <= 4 (frames, flows, SD latents): accumulators are pixel-interleaved [N,H,W,4] fp32 and
//     every corner is ONE 16-byte `red.global.add.v4.f32` (4 L2 reductions per pixel instead of
//     4*(C+1)); otherwise accumulators are planar [N,C+1,H,W] and a warp's reds of one corner and
//     one channel fall on one or two 128-byte lines;
//   * accumulation is always fp32 (fp64 for fp64 tensors); bf16 is rounded once, at the output;
//   * K2 fuses the eps rule, the divide, the optional (1 - mask) product, the cast, the save of
//     the normaliser for backward, and re-zeroes the accumulators so that the next call needs no
//     memset (DCB_FLAG_WS_CLEAN).
#include "dcb_common.cuh"

#include <stdlib.h>

namespace dcb {

constexpr int kThreads = 256;

struct FwdArgs {
    View in, flow, metric, mask;
    void* acc;       // accumulators (workspace, or `out` itself for SUM when T == accumulator type)
    void* out;       // [N,C,H,W] contiguous
    void* norm;      // optional [N,1,H,W] accumulator-typed
    unsigned total;  // N*H*W
    unsigned HW;
    int N, C, H, W;
    int Cacc;        // channels held by the accumulators: C (SUM) or C+1
    int cgroup;      // planar scatter: channels per blockIdx.y
    int mode, eps;
    int rezero;      // K2 leaves the accumulators zeroed
};

template <class A> __device__ __forceinline__ A weight_of(int mode, A m) {
    // g(m) of softsplat.py:240-247: 1 (avg), m (linear), exp(m) (soft)
    return mode == DCB_MODE_SOFT ? exp_t(m) : (mode == DCB_MODE_LINEAR ? m : (A)1);
}

// ---------------------------------------------------------------------------------------------
// K1a: planar accumulators, any C. grid = (ceil(total/256), ceil(Cacc/cgroup)).
// ---------------------------------------------------------------------------------------------
template <class T, class TF>
__global__ void __launch_bounds__(kThreads) k_scatter_planar(const FwdArgs a) {
    using A = typename Acc<T>::type;
    const unsigned p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);

    const TF* fp = (const TF*)a.flow.p + n * a.flow.sN + y * a.flow.sH + x * a.flow.sW;
    const Foot<A> f = make_foot<A>(x, y, ld<A>(fp), ld<A>(fp + a.flow.sC));
    if (!f.finite) return;                                        // softsplat.py:301-302

    const int W = a.W, H = a.H;
    const int x1 = (int)((unsigned)f.x0 + 1u), y1 = (int)((unsigned)f.y0 + 1u);
    const bool vx0 = (unsigned)f.x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W;
    const bool vy0 = (unsigned)f.y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
    const bool bnw = vx0 && vy0, bne = vx1 && vy0, bsw = vx0 && vy1, bse = vx1 && vy1;
    if (!(bnw || bne || bsw || bse)) return;
    // offsets are only formed for in-range corners; H*W < 2^31 is checked by the host
    const long long onw = (long long)f.y0 * W + f.x0;

    A g = (A)1;
    if (a.mode == DCB_MODE_LINEAR || a.mode == DCB_MODE_SOFT) {
        const T* mp = (const T*)a.metric.p + n * a.metric.sN + y * a.metric.sH + x * a.metric.sW;
        g = weight_of<A>(a.mode, ld<A>(mp));
    }

    const int c0 = blockIdx.y * a.cgroup;
    const int c1 = min(a.Cacc, c0 + a.cgroup);
    const T* ip = (const T*)a.in.p + n * a.in.sN + y * a.in.sH + x * a.in.sW + c0 * a.in.sC;
    A* acc = (A*)a.acc + ((long long)n * a.Cacc + c0) * a.HW + onw;
    for (int c = c0; c < c1; ++c, ip += a.in.sC, acc += a.HW) {
        A v;
        if (c < a.C) {
            v = ld<A>(ip);
            if (a.mode >= DCB_MODE_LINEAR) v = mul_rn(v, g);      // in * m  /  in * exp(m)
        } else {
            v = g;                                                // appended channel
        }
        if (bnw) red_add(acc, mul_rn(v, f.wnw));
        if (bne) red_add(acc + 1, mul_rn(v, f.wne));
        if (bsw) red_add(acc + W, mul_rn(v, f.wsw));
        if (bse) red_add(acc + W + 1, mul_rn(v, f.wse));
    }
}

// ---------------------------------------------------------------------------------------------
// K1b: pixel-interleaved accumulators [N,H,W,4] fp32, Cacc <= 4. grid = ceil(total/256).
// ---------------------------------------------------------------------------------------------
template <class T, class TF>
__global__ void __launch_bounds__(kThreads) k_scatter_vec4(const FwdArgs a) {
    const unsigned p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);

    const TF* fp = (const TF*)a.flow.p + n * a.flow.sN + y * a.flow.sH + x * a.flow.sW;
    const Foot<float> f = make_foot<float>(x, y, ld<float>(fp), ld<float>(fp + a.flow.sC));
    if (!f.finite) return;

    const int W = a.W, H = a.H;
    const int x1 = (int)((unsigned)f.x0 + 1u), y1 = (int)((unsigned)f.y0 + 1u);
    const bool vx0 = (unsigned)f.x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W;
    const bool vy0 = (unsigned)f.y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
    if (!((vx0 || vx1) && (vy0 || vy1))) return;

    float g = 1.f;
    if (a.mode == DCB_MODE_LINEAR || a.mode == DCB_MODE_SOFT) {
        const T* mp = (const T*)a.metric.p + n * a.metric.sN + y * a.metric.sH + x * a.metric.sW;
        g = weight_of<float>(a.mode, ld<float>(mp));
    }
    const T* ip = (const T*)a.in.p + n * a.in.sN + y * a.in.sH + x * a.in.sW;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        v[c] = 0.f;
        if (c < a.C) {
            v[c] = ld<float>(ip + c * a.in.sC);
            if (a.mode >= DCB_MODE_LINEAR) v[c] = mul_rn(v[c], g);
        } else if (c == a.C && a.mode != DCB_MODE_SUM) {
            v[c] = g;
        }
    }
    float* acc = (float*)a.acc + ((long long)n * a.HW + (long long)f.y0 * W + f.x0) * 4;
    if (vx0 && vy0) red_add_v4(acc, mul_rn(v[0], f.wnw), mul_rn(v[1], f.wnw), mul_rn(v[2], f.wnw), mul_rn(v[3], f.wnw));
    if (vx1 && vy0) red_add_v4(acc + 4, mul_rn(v[0], f.wne), mul_rn(v[1], f.wne), mul_rn(v[2], f.wne), mul_rn(v[3], f.wne));
    if (vx0 && vy1) red_add_v4(acc + 4 * W, mul_rn(v[0], f.wsw), mul_rn(v[1], f.wsw), mul_rn(v[2], f.wsw), mul_rn(v[3], f.wsw));
    if (vx1 && vy1) red_add_v4(acc + 4 * W + 4, mul_rn(v[0], f.wse), mul_rn(v[1], f.wse), mul_rn(v[2], f.wse), mul_rn(v[3], f.wse));
}

// normaliser rule of softsplat.py:256-266
template <class A> __device__ __forceinline__ A apply_eps(A d, int eps) {
    if (eps == DCB_EPS_ADD) return add_rn(d, (A)0.0000001);
    if (eps == DCB_EPS_ZERO) return d == (A)0 ? (A)1 : d;
    return d < (A)0.0000001 ? (A)0.0000001 : d;      // clip(1e-7, None); NaN propagates like torch.clip
}

// ---------------------------------------------------------------------------------------------
// K2: normalise + cast (+ mask, + save normaliser, + re-zero). One thread per TARGET pixel.
// VEC4 selects the accumulator layout. For SUM the "normaliser" is absent: plain cast/copy.
// ---------------------------------------------------------------------------------------------
template <class T, bool VEC4>
__global__ void __launch_bounds__(kThreads) k_normalize(const FwdArgs a) {
    using A = typename Acc<T>::type;
    const unsigned p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const bool normalised = a.mode != DCB_MODE_SUM;

    A keep = (A)1;
    if (a.mask.p) {
        const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
        const T* mp = (const T*)a.mask.p + n * a.mask.sN + y * a.mask.sH + x * a.mask.sW;
        keep = sub_rn((A)1, ld<A>(mp));                           // control_utils.py:69-70
    }
    T* out = (T*)a.out + (long long)n * a.C * a.HW + r;

    if (VEC4) {
        float4* ap = (float4*)a.acc + p;
        const float4 s = *ap;
        const float sv[4] = {s.x, s.y, s.z, s.w};
        float d = 1.f;
        if (normalised) {
            d = apply_eps<float>(a.C == 3 ? s.w : (a.C == 2 ? s.z : (a.C == 1 ? s.y : s.x)), a.eps);
            if (a.norm) ((float*)a.norm)[p] = d;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < a.C) {
                float o = normalised ? sv[c] / d : sv[c];
                if (a.mask.p) o = mul_rn(o, (float)keep);
                st<T, A>(out + (long long)c * a.HW, (A)o);
            }
        }
        if (a.rezero) *ap = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        A* ap = (A*)a.acc + (long long)n * a.Cacc * a.HW + r;
        A d = (A)1;
        if (normalised) {
            d = apply_eps<A>(ap[(long long)a.C * a.HW], a.eps);
            if (a.norm) ((A*)a.norm)[p] = d;
            if (a.rezero) ap[(long long)a.C * a.HW] = (A)0;
        }
        for (int c = 0; c < a.C; ++c) {
            A* q = ap + (long long)c * a.HW;
            A o = normalised ? *q / d : *q;
            if (a.mask.p) o = mul_rn(o, keep);
            st<T, A>(out + (long long)c * a.HW, o);
            if (a.rezero) *q = (A)0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
// implemented in splat_pipe.cu
long long pipe_workspace(long long N, long long H, long long W);
int splat_pipe_impl(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric, const DcbTensor* out,
                    const DcbTensor* norm, const DcbTensor* mask, void* ws, int mode, int eps, bool ws_clean,
                    cudaStream_t st, bool ones_metric = false, const DcbTensor* mask_out = nullptr);
bool pipe_supported(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric);
// implemented in splat_planar.cu
long long planar_workspace(long long N, long long C, long long H, long long W, int dtype, int mode);
int splat_planar_impl(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric, const DcbTensor* out,
                      const DcbTensor* norm, const DcbTensor* mask, void* ws, int mode, int eps, bool ws_clean,
                      cudaStream_t st);

// C+1 <= 4 channels in fp32 / bf16: the persistent pipelined kernel (splat_pipe.cu)
static bool use_pipe(int dtype, int mode, long long C) {
    static const bool disabled = getenv("DCB_NO_PIPE") != nullptr;      // debugging / A-B measurements only
    if (disabled || dtype == DCB_F64) return false;
    return C + (mode == DCB_MODE_SUM ? 0 : 1) <= 4;
}

long long splat_fwd_workspace(long long N, long long C, long long H, long long W, int dtype, int mode) {
    const long long cacc = C + (mode == DCB_MODE_SUM ? 0 : 1);
    const long long esz = dtype == DCB_F64 ? 8 : 4;
    static const bool no_pipe = getenv("DCB_NO_PIPE") != nullptr;
    if (use_pipe(dtype, mode, C)) return pipe_workspace(N, H, W);
    if (!no_pipe && dtype != DCB_F64) return planar_workspace(N, C, H, W, dtype, mode);
    if (mode == DCB_MODE_SUM && dtype != DCB_BF16) return 0;      // planar reds go straight into `out`
    if (dtype != DCB_F64 && cacc <= 4) return align_up(N * H * W * 16, 256);   // (DCB_NO_PIPE) vec4 accumulators
    return align_up(N * cacc * H * W * esz, 256);
}

template <class T, class TF>
static int launch_fwd(FwdArgs& a, bool vec4, bool ws_clean, bool acc_is_out, long long acc_bytes, cudaStream_t st) {
    const unsigned blocks = (a.total + kThreads - 1) / kThreads;
    if (!ws_clean) DCB_CHECK_CUDA(cudaMemsetAsync(a.acc, 0, (size_t)acc_bytes, st));
    if (vec4) {
        k_scatter_vec4<T, TF><<<blocks, kThreads, 0, st>>>(a);
        DCB_CHECK_LAUNCH("k_scatter_vec4");
    } else {
        // split channels over blockIdx.y when there are too few pixels to fill 148 SMs
        int groups = 1;
        const long long want = 148LL * 8 * kThreads;              // ~8 resident CTAs per SM
        while (groups < a.Cacc && (long long)a.total * groups < want && groups < 64) groups *= 2;
        a.cgroup = (a.Cacc + groups - 1) / groups;
        dim3 grid(blocks, (a.Cacc + a.cgroup - 1) / a.cgroup);
        k_scatter_planar<T, TF><<<grid, kThreads, 0, st>>>(a);
        DCB_CHECK_LAUNCH("k_scatter_planar");
    }
    if (!acc_is_out) {
        if (vec4) k_normalize<T, true><<<blocks, kThreads, 0, st>>>(a);
        else k_normalize<T, false><<<blocks, kThreads, 0, st>>>(a);
        DCB_CHECK_LAUNCH("k_normalize");
    }
    return DCB_OK;
}

// epilogue-only entry used by the deterministic path (det.cu): planar accumulators -> out
int normalize_planar_launch(int dtype, const DcbTensor* in, const DcbTensor* out, const DcbTensor* norm,
                            const DcbTensor* mask, void* acc, int mode, int eps, cudaStream_t st) {
    FwdArgs a;
    a.in = make_view(in);
    a.flow = make_view(nullptr);
    a.metric = make_view(nullptr);
    a.mask = make_view(mask);
    a.N = (int)in->size[0]; a.C = (int)in->size[1]; a.H = (int)in->size[2]; a.W = (int)in->size[3];
    a.HW = (unsigned)(in->size[2] * in->size[3]);
    a.total = (unsigned)(in->size[0] * in->size[2] * in->size[3]);
    a.Cacc = a.C + (mode == DCB_MODE_SUM ? 0 : 1);
    a.cgroup = a.Cacc;
    a.mode = mode; a.eps = eps;
    a.acc = acc;
    a.out = out->ptr;
    a.norm = norm ? norm->ptr : nullptr;
    a.rezero = 0;
    const unsigned blocks = (a.total + kThreads - 1) / kThreads;
    switch (dtype) {
        case DCB_F32: k_normalize<float, false><<<blocks, kThreads, 0, st>>>(a); break;
        case DCB_F64: k_normalize<double, false><<<blocks, kThreads, 0, st>>>(a); break;
        case DCB_BF16: k_normalize<__nv_bfloat16, false><<<blocks, kThreads, 0, st>>>(a); break;
        default: return set_error(DCB_E_DTYPE, "normalize: unsupported dtype %d", dtype);
    }
    DCB_CHECK_LAUNCH("k_normalize");
    return DCB_OK;
}

int splat_fwd_impl(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric, const DcbTensor* out,
                   const DcbTensor* norm, const DcbTensor* mask, void* ws, long long ws_bytes, int mode, int eps,
                   int flags, cudaStream_t st) {
    FwdArgs a;
    a.in = make_view(in);
    a.flow = make_view(flow);
    a.metric = make_view(metric);
    a.mask = make_view(mask);
    a.N = (int)in->size[0]; a.C = (int)in->size[1]; a.H = (int)in->size[2]; a.W = (int)in->size[3];
    a.HW = (unsigned)(in->size[2] * in->size[3]);
    a.total = (unsigned)(in->size[0] * in->size[2] * in->size[3]);
    a.Cacc = a.C + (mode == DCB_MODE_SUM ? 0 : 1);
    a.cgroup = a.Cacc;
    a.mode = mode; a.eps = eps;
    a.out = out->ptr;
    a.norm = norm ? norm->ptr : nullptr;
    if (a.total == 0 || a.C == 0) return DCB_OK;

    if (use_pipe(in->dtype, mode, a.C) && pipe_supported(in, flow, metric)) {
        const long long need = pipe_workspace(a.N, a.H, a.W);
        if (!ws || ws_bytes < need || ((uintptr_t)ws & 255))
            return set_error(DCB_E_WORKSPACE, "splat_fwd: workspace of %lld bytes (256 B aligned) required, got %lld", need, ws_bytes);
        return splat_pipe_impl(in, flow, metric, out, norm, mask, ws, mode, eps, (flags & DCB_FLAG_WS_CLEAN) != 0, st);
    }
    if (in->dtype != DCB_F64 && getenv("DCB_NO_PIPE") == nullptr) {
        const long long need = planar_workspace(a.N, a.C, a.H, a.W, in->dtype, mode);
        if (need > 0 && (!ws || ws_bytes < need || ((uintptr_t)ws & 255)))
            return set_error(DCB_E_WORKSPACE, "splat_fwd: workspace of %lld bytes (256 B aligned) required, got %lld", need, ws_bytes);
        return splat_planar_impl(in, flow, metric, out, norm, mask, ws, mode, eps, (flags & DCB_FLAG_WS_CLEAN) != 0, st);
    }
    // the one-kernel-per-stage paths below: fp64, and A-B measurements (DCB_NO_PIPE=1)
    const bool vec4 = in->dtype != DCB_F64 && a.Cacc <= 4 && !(mode == DCB_MODE_SUM && in->dtype == DCB_F32);
    const bool acc_is_out = (mode == DCB_MODE_SUM && in->dtype != DCB_BF16 && !mask);
    long long acc_bytes;
    if (acc_is_out) {
        a.acc = out->ptr;
        acc_bytes = (long long)a.total * a.C * elem_size(in->dtype);
    } else {
        acc_bytes = vec4 ? (long long)a.total * 16 : (long long)a.total * a.Cacc * (in->dtype == DCB_F64 ? 8 : 4);
        if (!ws || ws_bytes < acc_bytes || ((uintptr_t)ws & 255))
            return set_error(DCB_E_WORKSPACE, "splat_fwd: workspace of %lld bytes (256 B aligned) required, got %lld",
                             acc_bytes, ws_bytes);
        a.acc = ws;
    }
    const bool ws_clean = (flags & DCB_FLAG_WS_CLEAN) && !acc_is_out;
    a.rezero = ws_clean ? 1 : 0;

    const bool flow_f32 = flow->dtype == DCB_F32;
    switch (in->dtype) {
        case DCB_F32: return launch_fwd<float, float>(a, vec4, ws_clean, acc_is_out, acc_bytes, st);
        case DCB_F64: return launch_fwd<double, double>(a, vec4, ws_clean, acc_is_out, acc_bytes, st);
        case DCB_BF16:
            return flow_f32 ? launch_fwd<__nv_bfloat16, float>(a, vec4, ws_clean, acc_is_out, acc_bytes, st)
                            : launch_fwd<__nv_bfloat16, __nv_bfloat16>(a, vec4, ws_clean, acc_is_out, acc_bytes, st);
    }
    return set_error(DCB_E_DTYPE, "splat_fwd: unsupported dtype %d", in->dtype);
}

}  // namespace dcb
