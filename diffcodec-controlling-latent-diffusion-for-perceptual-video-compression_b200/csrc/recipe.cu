// recipe.cu -- occlusion mask and the fused conditioning builder (K5) for sm_100a.
//
// dcb_occlusion_mask replaces compute_mask(), controlnet/control_utils.py:11-17.
// dcb_residual_fused replaces the arithmetic of ResidueDataset.__getitem__
// (controlnet/dataset.py:233-265) and WarpingDatasetWrapper.__getitem__
// (controlnet/residual_utils.py:159-199): in the reference that is 4 softsplat() calls (each
// cat + memset + scatter + 3 post-ops), 2 norms, 2 compares and ~10 more eager kernels per frame;
// here it is ONE scatter launch and ONE epilogue launch for a whole batch of frames:
//
//   scatter (per source pixel, flows read once):
//     footprint of flow1 -> accA[y,x] += w * (img_r*e, img_g*e, img_b*e, e)      (red.v4)
//                           accB[y,x] += w * (flow2_x*e, flow2_y*e)              (red.v2)
//     footprint of flow2 -> accC[y,x] += w * (flow1_x*e, flow1_y*e, e, 0)        (red.v4)
//     (e = exp(1): the reference splats with an all-ones metric in 'soft' mode; the flow2 splat by
//      flow1 shares the footprint AND the normaliser channel of the image splat)
//   epilogue (per target pixel): three normalisations, both masks, the fusion weights, the
//     optional double-hole fill and the residual, written once; accumulators re-zeroed.
#include "dcb_common.cuh"

namespace dcb {

constexpr float kExp1 = 2.7182817459106445f;   // expf(1.0f), what tenMetric.exp() yields for ones

struct RecipeArgs {
    View img, flow1, flow2, gt;
    float* accA;     // [N*H*W][4]
    float* accB;     // [N*H*W][2]
    float* accC;     // [N*H*W][4]
    void* fused;     // [N,C,H,W]
    void* residual;  // [N,C,H,W]
    void* occ_fwd;   // [N,1,H,W] or null
    void* occ_bwd;   // [N,1,H,W] or null
    void* mask;      // dcb_occlusion_mask output
    unsigned total, HW;
    int N, C, H, W;
    int variant, rezero;
};

struct Corners {
    bool b[4];
    long long base;  // y0 * W + x0
};

__device__ __forceinline__ Corners corners_of(const Foot<float>& f, int W, int H) {
    Corners c;
    const int x1 = (int)((unsigned)f.x0 + 1u), y1 = (int)((unsigned)f.y0 + 1u);
    const bool vx0 = (unsigned)f.x0 < (unsigned)W, vx1 = (unsigned)x1 < (unsigned)W;
    const bool vy0 = (unsigned)f.y0 < (unsigned)H, vy1 = (unsigned)y1 < (unsigned)H;
    c.b[0] = f.finite && vx0 && vy0; c.b[1] = f.finite && vx1 && vy0;
    c.b[2] = f.finite && vx0 && vy1; c.b[3] = f.finite && vx1 && vy1;
    c.base = (long long)f.y0 * W + f.x0;
    return c;
}

// scatter of one 2-channel flow field (+ normaliser) by another: the splat inside compute_mask
template <class T>
__global__ void __launch_bounds__(256) k_mask_scatter(const RecipeArgs a) {
    const unsigned p = blockIdx.x * 256 + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
    const T* ap = (const T*)a.flow1.p + n * a.flow1.sN + y * a.flow1.sH + x * a.flow1.sW;   // splatted
    const T* bp = (const T*)a.flow2.p + n * a.flow2.sN + y * a.flow2.sH + x * a.flow2.sW;   // motion
    const float ax = ld<float>(ap), ay = ld<float>(ap + a.flow1.sC);
    const Foot<float> f = make_foot<float>(x, y, ld<float>(bp), ld<float>(bp + a.flow2.sC));
    const Corners c = corners_of(f, a.W, a.H);
    const float vx = mul_rn(ax, kExp1), vy = mul_rn(ay, kExp1);
    float* acc = a.accC + ((long long)n * a.HW + c.base) * 4;
    const long long off[4] = {0, 4, 4ll * a.W, 4ll * a.W + 4};
    const float w[4] = {f.wnw, f.wne, f.wsw, f.wse};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (c.b[k]) red_add_v4(acc + off[k], mul_rn(vx, w[k]), mul_rn(vy, w[k]), mul_rn(kExp1, w[k]), 0.f);
}

__device__ __forceinline__ float occlusion(float wx, float wy, float d, float mx, float my) {
    // control_utils.py:15-16: ||flow + warped||_2 > 0.3
    const float n = add_rn(d, 0.0000001f);
    const float ex = add_rn(mx, wx / n), ey = add_rn(my, wy / n);
    return sqrtf(add_rn(mul_rn(ex, ex), mul_rn(ey, ey))) > 0.3f ? 1.f : 0.f;
}

template <class T>
__global__ void __launch_bounds__(256) k_mask_epilogue(const RecipeArgs a) {
    const unsigned p = blockIdx.x * 256 + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
    float4* ap = (float4*)a.accC + p;
    const float4 s = *ap;
    const T* bp = (const T*)a.flow2.p + n * a.flow2.sN + y * a.flow2.sH + x * a.flow2.sW;
    st<T, float>((T*)a.mask + p, occlusion(s.x, s.y, s.z, ld<float>(bp), ld<float>(bp + a.flow2.sC)));
    if (a.rezero) *ap = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <class T>
__global__ void __launch_bounds__(256) k_recipe_scatter(const RecipeArgs a) {
    const unsigned p = blockIdx.x * 256 + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
    const T* p1 = (const T*)a.flow1.p + n * a.flow1.sN + y * a.flow1.sH + x * a.flow1.sW;
    const T* p2 = (const T*)a.flow2.p + n * a.flow2.sN + y * a.flow2.sH + x * a.flow2.sW;
    const float f1x = ld<float>(p1), f1y = ld<float>(p1 + a.flow1.sC);
    const float f2x = ld<float>(p2), f2y = ld<float>(p2 + a.flow2.sC);
    const T* ip = (const T*)a.img.p + n * a.img.sN + y * a.img.sH + x * a.img.sW;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = c < a.C ? mul_rn(ld<float>(ip + c * a.img.sC), kExp1) : 0.f;

    const float w2x = mul_rn(f2x, kExp1), w2y = mul_rn(f2y, kExp1);
    const float w1x = mul_rn(f1x, kExp1), w1y = mul_rn(f1y, kExp1);
    {   // footprint of flow1: image (+ normaliser) and flow2
        const Foot<float> f = make_foot<float>(x, y, f1x, f1y);
        const Corners c = corners_of(f, a.W, a.H);
        float* accA = a.accA + ((long long)n * a.HW + c.base) * 4;
        float* accB = a.accB + ((long long)n * a.HW + c.base) * 2;
        const long long off[4] = {0, 1, a.W, (long long)a.W + 1};
        const float w[4] = {f.wnw, f.wne, f.wsw, f.wse};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c.b[k]) {
                red_add_v4(accA + off[k] * 4, mul_rn(v[0], w[k]), mul_rn(v[1], w[k]), mul_rn(v[2], w[k]), mul_rn(kExp1, w[k]));
                red_add_v2(accB + off[k] * 2, mul_rn(w2x, w[k]), mul_rn(w2y, w[k]));
            }
    }
    {   // footprint of flow2: flow1 (+ its own normaliser)
        const Foot<float> f = make_foot<float>(x, y, f2x, f2y);
        const Corners c = corners_of(f, a.W, a.H);
        float* accC = a.accC + ((long long)n * a.HW + c.base) * 4;
        const long long off[4] = {0, 1, a.W, (long long)a.W + 1};
        const float w[4] = {f.wnw, f.wne, f.wsw, f.wse};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c.b[k]) red_add_v4(accC + off[k] * 4, mul_rn(w1x, w[k]), mul_rn(w1y, w[k]), mul_rn(kExp1, w[k]), 0.f);
    }
}

template <class T>
__global__ void __launch_bounds__(256) k_recipe_epilogue(const RecipeArgs a) {
    const unsigned p = blockIdx.x * 256 + threadIdx.x;
    if (p >= a.total) return;
    const unsigned n = p / a.HW, r = p - n * a.HW;
    const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
    float4* pa = (float4*)a.accA + p;
    float2* pb = (float2*)a.accB + p;
    float4* pc = (float4*)a.accC + p;
    const float4 sa = *pa;
    const float2 sb = *pb;
    const float4 sc = *pc;
    if (a.rezero) {
        *pa = make_float4(0.f, 0.f, 0.f, 0.f);
        *pb = make_float2(0.f, 0.f);
        *pc = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const T* p1 = (const T*)a.flow1.p + n * a.flow1.sN + y * a.flow1.sH + x * a.flow1.sW;
    const T* p2 = (const T*)a.flow2.p + n * a.flow2.sN + y * a.flow2.sH + x * a.flow2.sW;
    // occ_fwd = compute_mask(flow1, flow2): flow1 splatted by flow2 (accC), compared with flow2
    const float of = occlusion(sc.x, sc.y, sc.z, ld<float>(p2), ld<float>(p2 + a.flow2.sC));
    // occ_bwd = compute_mask(flow2, flow1): flow2 splatted by flow1 (accB, normaliser of accA)
    const float ob = occlusion(sb.x, sb.y, sa.w, ld<float>(p1), ld<float>(p1 + a.flow1.sC));
    if (a.occ_fwd) st<T, float>((T*)a.occ_fwd + p, of);
    if (a.occ_bwd) st<T, float>((T*)a.occ_bwd + p, ob);

    float w0, w1;
    if (a.variant == DCB_RECIPE_DATASET) {        // dataset.py:255-259: masks are the confidences
        const float ws = add_rn(add_rn(of, ob), 0.000001f);
        w0 = of / ws; w1 = ob / ws;
    } else {                                      // residual_utils.py:181-185: ones are the confidences
        const float ws = add_rn(2.f, 0.000001f);
        w0 = 1.f / ws; w1 = w0;
    }
    const bool hole = a.variant == DCB_RECIPE_WRAPPER && add_rn(of, ob) > 1.5f;   // residual_utils.py:190-193
    const float d = add_rn(sa.w, 0.0000001f);
    const float sv[3] = {sa.x, sa.y, sa.z};
    const T* gp = (const T*)a.gt.p + n * a.gt.sN + y * a.gt.sH + x * a.gt.sW;
    T* fo = (T*)a.fused + (long long)n * a.C * a.HW + r;
    T* ro = (T*)a.residual + (long long)n * a.C * a.HW + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c < a.C) {
            float warped = sv[c] / d;
            // the reference stores warped1/warped2 in T before fusing; keep that rounding for bf16
            st<T, float>(fo + (long long)c * a.HW, warped);
            warped = ld<float>(fo + (long long)c * a.HW);
            float fused = add_rn(mul_rn(w0, warped), mul_rn(w1, warped));          // warped1 == warped2 (B-6)
            if (hole) fused = mul_rn(0.5f, add_rn(warped, warped));
            st<T, float>(fo + (long long)c * a.HW, fused);
            fused = ld<float>(fo + (long long)c * a.HW);
            st<T, float>(ro + (long long)c * a.HW, sub_rn(ld<float>(gp + c * a.gt.sC), fused));
        }
    }
}

// ---------------------------------------------------------------------------------------------
long long mask_workspace(long long N, long long H, long long W) { return align_up(N * H * W * 16, 256); }
long long recipe_workspace(long long N, long long H, long long W) { return align_up(N * H * W * 40, 256); }

static void fill(RecipeArgs& a, const DcbTensor* t, int C) {
    a.N = (int)t->size[0]; a.C = C; a.H = (int)t->size[2]; a.W = (int)t->size[3];
    a.HW = (unsigned)(t->size[2] * t->size[3]);
    a.total = (unsigned)(t->size[0] * t->size[2] * t->size[3]);
}

template <class T> static int launch_mask(const RecipeArgs& a, cudaStream_t st) {
    const unsigned blocks = (a.total + 255) / 256;
    k_mask_scatter<T><<<blocks, 256, 0, st>>>(a);
    DCB_CHECK_LAUNCH("k_mask_scatter");
    k_mask_epilogue<T><<<blocks, 256, 0, st>>>(a);
    DCB_CHECK_LAUNCH("k_mask_epilogue");
    return DCB_OK;
}

int occlusion_mask_impl(const DcbTensor* flow_a, const DcbTensor* flow_b, const DcbTensor* mask, void* ws,
                        long long ws_bytes, int flags, cudaStream_t st) {
    RecipeArgs a{};
    a.flow1 = make_view(flow_a); a.flow2 = make_view(flow_b);
    a.mask = mask->ptr;
    fill(a, flow_a, 2);
    if (a.total == 0) return DCB_OK;
    const long long need = mask_workspace(a.N, a.H, a.W);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255))
        return set_error(DCB_E_WORKSPACE, "occlusion_mask: workspace of %lld bytes required, got %lld", need, ws_bytes);
    a.accC = (float*)ws;
    a.rezero = (flags & DCB_FLAG_WS_CLEAN) ? 1 : 0;
    if (!a.rezero) DCB_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)a.total * 16, st));
    switch (flow_a->dtype) {
        case DCB_F32: return launch_mask<float>(a, st);
        case DCB_BF16: return launch_mask<__nv_bfloat16>(a, st);
    }
    return set_error(DCB_E_DTYPE, "occlusion_mask: F32 or BF16 only, got %d", flow_a->dtype);
}

template <class T> static int launch_recipe(const RecipeArgs& a, cudaStream_t st) {
    const unsigned blocks = (a.total + 255) / 256;
    k_recipe_scatter<T><<<blocks, 256, 0, st>>>(a);
    DCB_CHECK_LAUNCH("k_recipe_scatter");
    k_recipe_epilogue<T><<<blocks, 256, 0, st>>>(a);
    DCB_CHECK_LAUNCH("k_recipe_epilogue");
    return DCB_OK;
}

int residual_fused_impl(const DcbTensor* image1, const DcbTensor* flow1, const DcbTensor* flow2, const DcbTensor* gt,
                        const DcbTensor* fused, const DcbTensor* residual, const DcbTensor* occ_fwd,
                        const DcbTensor* occ_bwd, void* ws, long long ws_bytes, int variant, int flags,
                        cudaStream_t st) {
    RecipeArgs a{};
    a.img = make_view(image1); a.flow1 = make_view(flow1); a.flow2 = make_view(flow2); a.gt = make_view(gt);
    a.fused = fused->ptr; a.residual = residual->ptr;
    a.occ_fwd = occ_fwd ? occ_fwd->ptr : nullptr;
    a.occ_bwd = occ_bwd ? occ_bwd->ptr : nullptr;
    a.variant = variant;
    fill(a, image1, (int)image1->size[1]);
    if (a.total == 0) return DCB_OK;
    const long long need = recipe_workspace(a.N, a.H, a.W);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255))
        return set_error(DCB_E_WORKSPACE, "residual_fused: workspace of %lld bytes required, got %lld", need, ws_bytes);
    a.accA = (float*)ws;
    a.accC = a.accA + (size_t)a.total * 4;
    a.accB = a.accC + (size_t)a.total * 4;
    a.rezero = (flags & DCB_FLAG_WS_CLEAN) ? 1 : 0;
    if (!a.rezero) DCB_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)a.total * 40, st));
    switch (image1->dtype) {
        case DCB_F32: return launch_recipe<float>(a, st);
        case DCB_BF16: return launch_recipe<__nv_bfloat16>(a, st);
    }
    return set_error(DCB_E_DTYPE, "residual_fused: F32 or BF16 only, got %d", image1->dtype);
}

}  // namespace dcb
