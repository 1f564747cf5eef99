// block.cu -- the motion-compensation part of one bi-directional conditioning block as ONE C call per pyramid scale
// (SURVEY.md section 8, row f-1), sm_100a.
//
// Replaces, per scale, controlnet/extractors.py:289-310 (and :181-205): two compute_mask() calls
// (control_utils.py:11-17), two FeatureWarperSoftsplat splats with (1 - mask) (control_utils.py:61-72), the confidence
// fusion and the double-hole fill with its `holes.any()` host sync (extractors.py:298-310) -- in the reference ~45 eager
// kernels, 4 NVRTC cache lookups and one device->host stall per scale; in round 1 five library calls from Python. Here
// the host enters the library once. Scales whose accumulators fit the L2 (every pyramid level of the live consumer) run
// as TWO launches (pyramid.cu: all four scatter jobs, then masks + normalisation + fusion); bigger tensors run the
// kernels of recipe.cu / splat_*.cu / fuse.cu back to back on the caller's stream.
#include "dcb_common.cuh"

namespace dcb {

long long mask_workspace(long long N, long long H, long long W);
long long splat_fwd_workspace(long long N, long long C, long long H, long long W, int dtype, int mode);
long long splat_bwd_workspace(long long N, long long C, long long H, long long W, int dtype, int mode);
int occlusion_mask_impl(const DcbTensor*, const DcbTensor*, const DcbTensor*, void*, long long, int, cudaStream_t);
int splat_fwd_impl(const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, void*,
                   long long, int, int, int, cudaStream_t);
int splat_bwd_impl(const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*,
                   const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, void*, long long, int, int, cudaStream_t);
int bidir_fuse_fwd_impl(const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*,
                        const DcbTensor*, cudaStream_t);
int bidir_fuse_bwd_impl(const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*,
                        const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, const DcbTensor*, cudaStream_t);

// pyramid.cu: the two-launch form (every scatter job in one launch, everything else in a second)
long long pyramid_workspace(const DcbPyramidLevel* lv, int n);
int bidir_pyramid_fwd_impl(const DcbPyramidLevel* lv, int n, void* ws, int flags, cudaStream_t st);

// One scale goes through the two-launch pyramid kernels when the accumulators of BOTH directions fit the L2 together
// (every pyramid level of the live consumer does); bigger tensors keep the launch sequence below (frame-group pipeline
// or per-target lists per splat).
constexpr long long kFusedLevelBytes = 48ll << 20;
static long long fused_level_bytes(long long N, long long C, long long H, long long W, int dtype) {
    DcbTensor t = {};
    t.dtype = dtype; t.size[0] = N; t.size[1] = C; t.size[2] = H; t.size[3] = W;
    DcbPyramidLevel lv = {};
    lv.first = &t;
    return pyramid_workspace(&lv, 1);
}
static bool block_is_fused(long long N, long long C, long long H, long long W, int dtype) {
    return (dtype == DCB_F32 || dtype == DCB_BF16) && fused_level_bytes(N, C, H, W, dtype) <= kFusedLevelBytes;
}

static DcbTensor contiguous_like(const DcbTensor* t, long long C, int dtype, void* ptr) {
    DcbTensor r = *t;
    r.ptr = ptr; r.dtype = dtype;
    r.size[1] = C;
    r.stride[3] = 1; r.stride[2] = t->size[3]; r.stride[1] = t->size[2] * t->size[3]; r.stride[0] = C * r.stride[1];
    return r;
}

long long block_fwd_acc_bytes(long long N, long long C, long long H, long long W, int dtype) {
    const long long a = mask_workspace(N, H, W), b = splat_fwd_workspace(N, C, H, W, dtype, DCB_MODE_SOFT);
    const long long c = block_is_fused(N, C, H, W, dtype) ? fused_level_bytes(N, C, H, W, dtype) : 0;
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// scratch of the forward: whatever of {warped_f, warped_b, occ_f, occ_b} the caller does not want back
long long block_fwd_scratch_bytes(long long N, long long C, long long H, long long W, int dtype) {
    const long long s = elem_size(dtype);
    return 2 * align_up(N * C * H * W * s, 256) + 2 * align_up(N * H * W * s, 256);
}

// scratch of the backward: splat backward scalars + gradients w.r.t. both warped maps + both fusion-confidence gradients
long long block_bwd_scratch_bytes(long long N, long long C, long long H, long long W, int dtype) {
    const long long s = elem_size(dtype);
    return align_up(splat_bwd_workspace(N, C, H, W, dtype, DCB_MODE_SOFT), 256) + 2 * align_up(N * C * H * W * s, 256) + 2 * align_up(N * H * W * s, 256);
}

int bidir_block_fwd_impl(const DcbTensor* first, const DcbTensor* last, const DcbTensor* flow_f, const DcbTensor* flow_b,
                         const DcbTensor* metric_f, const DcbTensor* metric_b, const DcbTensor* fused, const DcbTensor* warped_f,
                         const DcbTensor* warped_b, const DcbTensor* norm_f, const DcbTensor* norm_b, const DcbTensor* occ_f,
                         const DcbTensor* occ_b, void* ws_acc, long long acc_bytes, void* ws_scratch, long long scratch_bytes, int flags,
                         cudaStream_t st) {
    const long long N = first->size[0], C = first->size[1], H = first->size[2], W = first->size[3];
    if (N * C * H * W == 0) return DCB_OK;
    const int dt = first->dtype;
    if (block_is_fused(N, C, H, W, dt)) {
        // two launches for the scale: outputs the caller does not ask for are never stored
        const long long need = block_fwd_acc_bytes(N, C, H, W, dt);
        if (!ws_acc || acc_bytes < need || ((uintptr_t)ws_acc & 255))
            return set_error(DCB_E_WORKSPACE, "bidir_block_fwd: accumulator workspace of %lld bytes (256 B aligned) required, got %lld", need, acc_bytes);
        DcbPyramidLevel lv = {};
        lv.first = first; lv.last = last; lv.flow_f = flow_f; lv.flow_b = flow_b; lv.metric_f = metric_f; lv.metric_b = metric_b;
        lv.fused = fused->ptr;
        lv.warped_f = warped_f ? warped_f->ptr : nullptr; lv.warped_b = warped_b ? warped_b->ptr : nullptr;
        lv.norm_f = norm_f ? norm_f->ptr : nullptr; lv.norm_b = norm_b ? norm_b->ptr : nullptr;
        lv.occ_f = occ_f ? occ_f->ptr : nullptr; lv.occ_b = occ_b ? occ_b->ptr : nullptr;
        return bidir_pyramid_fwd_impl(&lv, 1, ws_acc, flags & DCB_FLAG_WS_CLEAN, st);
    }
    const long long s = elem_size(dt), big = align_up(N * C * H * W * s, 256), plane = align_up(N * H * W * s, 256);
    char* sc = (char*)ws_scratch;
    long long used = 0;
    DcbTensor t_wf, t_wb, t_of, t_ob;
    auto take = [&](const DcbTensor* given, DcbTensor& slot, long long Cx, long long bytes) -> const DcbTensor* {
        if (given) return given;
        slot = contiguous_like(first, Cx, dt, sc + used);
        used += bytes;
        return &slot;
    };
    const DcbTensor* wf = take(warped_f, t_wf, C, big);
    const DcbTensor* wb = take(warped_b, t_wb, C, big);
    const DcbTensor* of = take(occ_f, t_of, 1, plane);
    const DcbTensor* ob = take(occ_b, t_ob, 1, plane);
    if (used > 0 && (!ws_scratch || scratch_bytes < used || ((uintptr_t)ws_scratch & 255)))
        return set_error(DCB_E_WORKSPACE, "bidir_block_fwd: scratch of %lld bytes (256 B aligned) required, got %lld", used, scratch_bytes);
    const long long need = block_fwd_acc_bytes(N, C, H, W, dt);
    if (need > 0 && (!ws_acc || acc_bytes < need || ((uintptr_t)ws_acc & 255)))
        return set_error(DCB_E_WORKSPACE, "bidir_block_fwd: accumulator workspace of %lld bytes (256 B aligned) required, got %lld", need, acc_bytes);
    const int clean = flags & DCB_FLAG_WS_CLEAN;
    // occ_fwd = compute_mask(flow_f, flow_b), occ_bwd = compute_mask(flow_b, flow_f)      extractors.py:290-291
    int rc = occlusion_mask_impl(flow_f, flow_b, of, ws_acc, acc_bytes, clean, st);
    if (rc != DCB_OK) return rc;
    rc = occlusion_mask_impl(flow_b, flow_f, ob, ws_acc, acc_bytes, DCB_FLAG_WS_CLEAN, st);
    if (rc != DCB_OK) return rc;
    // warped = softsplat(feat, flow, metric, 'soft') * (1 - mask)                          control_utils.py:62-70
    rc = splat_fwd_impl(first, flow_f, metric_f, wf, norm_f, of, ws_acc, acc_bytes, DCB_MODE_SOFT, DCB_EPS_ADD, DCB_FLAG_WS_CLEAN, st);
    if (rc != DCB_OK) return rc;
    rc = splat_fwd_impl(last, flow_b, metric_b, wb, norm_b, ob, ws_acc, acc_bytes, DCB_MODE_SOFT, DCB_EPS_ADD, DCB_FLAG_WS_CLEAN, st);
    if (rc != DCB_OK) return rc;
    // confidence fusion + double-hole fill, no host sync                                   extractors.py:298-310
    return bidir_fuse_fwd_impl(wf, wb, metric_f, metric_b, of, ob, fused, st);
}

template <class T> __global__ void __launch_bounds__(256) k_add_inplace(T* dst, const T* src, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) st<T, float>(dst + i, add_rn(ld<float>(dst + i), ld<float>(src + i)));
}

static int add_inplace(int dt, void* dst, const void* src, long long n, cudaStream_t st) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dt == DCB_F32) k_add_inplace<float><<<blocks, 256, 0, st>>>((float*)dst, (const float*)src, n);
    else k_add_inplace<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n);
    DCB_CHECK_LAUNCH("k_add_inplace");
    return DCB_OK;
}

int bidir_block_bwd_impl(const DcbTensor* g, const DcbTensor* first, const DcbTensor* last, const DcbTensor* flow_f,
                         const DcbTensor* flow_b, const DcbTensor* metric_f, const DcbTensor* metric_b, const DcbTensor* warped_f,
                         const DcbTensor* warped_b, const DcbTensor* norm_f, const DcbTensor* norm_b, const DcbTensor* occ_f,
                         const DcbTensor* occ_b, const DcbTensor* g_first, const DcbTensor* g_last, const DcbTensor* g_metric_f,
                         const DcbTensor* g_metric_b, void* ws, long long ws_bytes, cudaStream_t st) {
    const long long N = first->size[0], C = first->size[1], H = first->size[2], W = first->size[3];
    if (N * C * H * W == 0) return DCB_OK;
    const int dt = first->dtype;
    const long long s = elem_size(dt), big = align_up(N * C * H * W * s, 256), plane = align_up(N * H * W * s, 256);
    const long long bwd = align_up(splat_bwd_workspace(N, C, H, W, dt, DCB_MODE_SOFT), 256);
    const long long need = bwd + 2 * big + 2 * plane;
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255))
        return set_error(DCB_E_WORKSPACE, "bidir_block_bwd: scratch of %lld bytes (256 B aligned) required, got %lld", need, ws_bytes);
    char* sc = (char*)ws;
    const DcbTensor gA = contiguous_like(first, C, dt, sc + bwd), gB = contiguous_like(first, C, dt, sc + bwd + big);
    const DcbTensor gca = contiguous_like(first, 1, dt, sc + bwd + 2 * big), gcb = contiguous_like(first, 1, dt, sc + bwd + 2 * big + plane);
    const bool wf = g_first || g_metric_f, wb = g_last || g_metric_b;
    // fusion backward: gradients w.r.t. both warped maps and, when the metric is learned, both confidences
    int rc = bidir_fuse_bwd_impl(g, warped_f, warped_b, metric_f, metric_b, occ_f, occ_b, wf ? &gA : nullptr, wb ? &gB : nullptr,
                                 g_metric_f ? &gca : nullptr, g_metric_b ? &gcb : nullptr, st);
    if (rc != DCB_OK) return rc;
    // splat backward of each direction (flows carry no gradient in the extractors: extractors.py:286-287 are no_grad inputs)
    if (wf) {
        rc = splat_bwd_impl(&gA, first, flow_f, metric_f, warped_f, norm_f, occ_f, g_first, nullptr, g_metric_f, sc, bwd, DCB_MODE_SOFT, DCB_EPS_ADD, st);
        if (rc != DCB_OK) return rc;
        if (g_metric_f) { rc = add_inplace(dt, g_metric_f->ptr, gca.ptr, N * H * W, st); if (rc != DCB_OK) return rc; }
    }
    if (wb) {
        rc = splat_bwd_impl(&gB, last, flow_b, metric_b, warped_b, norm_b, occ_b, g_last, nullptr, g_metric_b, sc, bwd, DCB_MODE_SOFT, DCB_EPS_ADD, st);
        if (rc != DCB_OK) return rc;
        if (g_metric_b) { rc = add_inplace(dt, g_metric_b->ptr, gcb.ptr, N * H * W, st); if (rc != DCB_OK) return rc; }
    }
    return DCB_OK;
}

}  // namespace dcb
