// red_bench.cu -- building-block measurements for the splat design on B200 (round 1).
// Not product code: a standalone binary that times candidate scatter / normalise kernels so that
// design choices in csrc/ are made from numbers (results are summarised in profiles/*.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o red_bench red_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 1080, W = 1920, HW = H * W;

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) b[i] = a[i];
}

// flow generator: pattern 0 = zero, 1 = smooth (amplitude ~8 px, wavelength ~100-300 px), 2 = hash-random +-32 px
__global__ void k_make_flow(float* flow, int frames, int pattern) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= frames * HW) return;
    int n = p / HW, r = p - n * HW, y = r / W, x = r - y * W;
    float fx = 0.f, fy = 0.f;
    if (pattern == 1) {
        fx = 5.f * sinf(x * 0.021f + y * 0.013f + n) + 3.f * cosf(y * 0.047f - x * 0.009f);
        fy = 4.f * cosf(x * 0.017f - y * 0.019f + 2 * n) + 3.f * sinf(x * 0.031f + 0.5f);
    } else if (pattern == 2) {
        unsigned h = (unsigned)p * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        fx = ((h & 0xffff) / 65535.f - 0.5f) * 64.f;
        fy = (((h >> 16) & 0xffff) / 65535.f - 0.5f) * 64.f;
    }
    flow[(size_t)n * 2 * HW + r] = fx;
    flow[(size_t)n * 2 * HW + HW + r] = fy;
}

struct Foot { int x0, y0; float w[4]; bool b[4]; };
__device__ __forceinline__ Foot foot(int x, int y, float fx, float fy) {
    Foot f;
    float px = x + fx, py = y + fy;
    float x0f = floorf(px), y0f = floorf(py);
    f.x0 = (int)x0f; f.y0 = (int)y0f;
    float dx = px - x0f, dy = py - y0f, ex = (x0f + 1.f) - px, ey = (y0f + 1.f) - py;
    f.w[0] = ex * ey; f.w[1] = dx * ey; f.w[2] = ex * dy; f.w[3] = dx * dy;
    bool vx0 = (unsigned)f.x0 < (unsigned)W, vx1 = (unsigned)(f.x0 + 1) < (unsigned)W;
    bool vy0 = (unsigned)f.y0 < (unsigned)H, vy1 = (unsigned)(f.y0 + 1) < (unsigned)H;
    f.b[0] = vx0 && vy0; f.b[1] = vx1 && vy0; f.b[2] = vx0 && vy1; f.b[3] = vx1 && vy1;
    return f;
}

// V1: soft-mode scatter, 3 planar channels + metric + flow in, interleaved float4 accumulators, red.v4
__global__ void __launch_bounds__(256) k_scatter_v4(const float* __restrict__ in, const float* __restrict__ metric,
                                                    const float* __restrict__ flow, float* acc, int total) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= total) return;
    int n = p / HW, r = p - n * HW, y = r / W, x = r - y * W;
    const float* fl = flow + (size_t)n * 2 * HW + r;
    Foot f = foot(x, y, fl[0], fl[HW]);
    float g = expf(metric[p]);
    const float* ip = in + (size_t)n * 3 * HW + r;
    float v0 = ip[0] * g, v1 = ip[HW] * g, v2 = ip[2 * HW] * g;
    float* a = acc + ((size_t)n * HW + (size_t)f.y0 * W + f.x0) * 4;
    const int off[4] = {0, 4, 4 * W, 4 * W + 4};
#pragma unroll
    for (int k = 0; k < 4; ++k) if (f.b[k]) red4(a + off[k], v0 * f.w[k], v1 * f.w[k], v2 * f.w[k], g * f.w[k]);
}

// V2: same, planar fp32 accumulators [N,4,H,W], scalar reds (16 per pixel)
__global__ void __launch_bounds__(256) k_scatter_planar(const float* __restrict__ in, const float* __restrict__ metric,
                                                        const float* __restrict__ flow, float* acc, int total) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= total) return;
    int n = p / HW, r = p - n * HW, y = r / W, x = r - y * W;
    const float* fl = flow + (size_t)n * 2 * HW + r;
    Foot f = foot(x, y, fl[0], fl[HW]);
    float g = expf(metric[p]);
    const float* ip = in + (size_t)n * 3 * HW + r;
    float v[4] = {ip[0] * g, ip[HW] * g, ip[2 * HW] * g, g};
    float* a = acc + (size_t)n * 4 * HW + (size_t)f.y0 * W + f.x0;
    const int off[4] = {0, 1, W, W + 1};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) if (f.b[k]) atomicAdd(a + (size_t)c * HW + off[k], v[c] * f.w[k]);
}

// V3: red.v4 with warp-level merging of horizontally adjacent contributions:
// lane i's NE/SE corner coincides with lane i+1's NW/SW corner when the flow is smooth, so the
// pair is summed through a shuffle and issued once (2-3 reds per pixel instead of 4).
__global__ void __launch_bounds__(256) k_scatter_v4_merge(const float* __restrict__ in, const float* __restrict__ metric,
                                                          const float* __restrict__ flow, float* acc, int total) {
    int p = blockIdx.x * 256 + threadIdx.x;
    bool live = p < total;
    int pc = live ? p : total - 1;
    int n = pc / HW, r = pc - n * HW, y = r / W, x = r - y * W;
    const float* fl = flow + (size_t)n * 2 * HW + r;
    Foot f = foot(x, y, fl[0], fl[HW]);
    float g = expf(metric[pc]);
    const float* ip = in + (size_t)n * 3 * HW + r;
    float v[4] = {ip[0] * g, ip[HW] * g, ip[2 * HW] * g, g};
    if (!live) { f.b[0] = f.b[1] = f.b[2] = f.b[3] = false; }
    const unsigned lane = threadIdx.x & 31;
    // my west column (x0) contributions: rows y0 (k=0) and y0+1 (k=2); east column (x0+1): k=1, k=3
    float westN[4], westS[4], eastN[4], eastS[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        westN[c] = f.b[0] ? v[c] * f.w[0] : 0.f; eastN[c] = f.b[1] ? v[c] * f.w[1] : 0.f;
        westS[c] = f.b[2] ? v[c] * f.w[2] : 0.f; eastS[c] = f.b[3] ? v[c] * f.w[3] : 0.f;
    }
    // does my left neighbour's east column equal my west column (same frame, same row footprint)?
    int lx0 = __shfl_up_sync(0xffffffffu, f.x0, 1), ly0 = __shfl_up_sync(0xffffffffu, f.y0, 1), ln = __shfl_up_sync(0xffffffffu, n, 1);
    bool take = lane > 0 && ln == n && lx0 + 1 == f.x0 && ly0 == f.y0;
    // the right neighbour tells me whether it took my east column
    bool given = __shfl_down_sync(0xffffffffu, (int)take, 1) && lane < 31;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float en = __shfl_up_sync(0xffffffffu, eastN[c], 1), es = __shfl_up_sync(0xffffffffu, eastS[c], 1);
        if (take) { westN[c] += en; westS[c] += es; }
    }
    bool vx0 = (unsigned)f.x0 < (unsigned)W, vy0 = (unsigned)f.y0 < (unsigned)H, vy1 = (unsigned)(f.y0 + 1) < (unsigned)H;
    bool vx1 = (unsigned)(f.x0 + 1) < (unsigned)W;
    float* a = acc + ((size_t)n * HW + (size_t)f.y0 * W + f.x0) * 4;
    if (live || take) {
        if (vx0 && vy0) red4(a, westN[0], westN[1], westN[2], westN[3]);
        if (vx0 && vy1) red4(a + 4 * W, westS[0], westS[1], westS[2], westS[3]);
    }
    if (live && !given) {
        if (vx1 && vy0) red4(a + 4, eastN[0], eastN[1], eastN[2], eastN[3]);
        if (vx1 && vy1) red4(a + 4 * W + 4, eastS[0], eastS[1], eastS[2], eastS[3]);
    }
}

// normalise: float4 acc -> 3 planar outputs, re-zero
__global__ void __launch_bounds__(256) k_norm_v4(float4* acc, float* __restrict__ out, int total, int rezero) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= total) return;
    int n = p / HW, r = p - n * HW;
    float4 s = acc[p];
    float d = s.w + 1e-7f;
    float* o = out + (size_t)n * 3 * HW + r;
    o[0] = s.x / d; o[HW] = s.y / d; o[2 * HW] = s.z / d;
    if (rezero) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(256) k_norm_planar(float* acc, float* __restrict__ out, int total, int rezero) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= total) return;
    int n = p / HW, r = p - n * HW;
    float* a = acc + (size_t)n * 4 * HW + r;
    float d = a[3 * HW] + 1e-7f;
    float* o = out + (size_t)n * 3 * HW + r;
    o[0] = a[0] / d; o[HW] = a[HW] / d; o[2 * HW] = a[2 * HW] / d;
    if (rezero) { a[0] = 0.f; a[HW] = 0.f; a[2 * HW] = 0.f; a[3 * HW] = 0.f; }
}

// pure red throughput probes: every thread issues one red.v4 / 4 scalar reds to its own (coalesced) slot
__global__ void __launch_bounds__(256) k_probe_red4(float* acc, int total) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p < total) red4(acc + (size_t)p * 4, 1.f, 2.f, 3.f, 4.f);
}
__global__ void __launch_bounds__(256) k_probe_red1x4(float* acc, int total) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p < total) { int n = p / HW, r = p - n * HW; float* a = acc + (size_t)n * 4 * HW + r;
        atomicAdd(a, 1.f); atomicAdd(a + HW, 2.f); atomicAdd(a + 2 * HW, 3.f); atomicAdd(a + 3 * HW, 4.f); }
}
__global__ void __launch_bounds__(256) k_probe_st4(float* acc, int total) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p < total) ((float4*)acc)[p] = make_float4(1.f, 2.f, 3.f, 4.f);
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    template <class F> float run(F f, int warm, int iters) {
        for (int i = 0; i < warm; ++i) f(i);
        CK(cudaDeviceSynchronize());
        std::vector<float> t;
        for (int i = 0; i < iters; ++i) {
            CK(cudaEventRecord(a)); f(i + warm); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); t.push_back(ms);
        }
        std::sort(t.begin(), t.end());
        return t[t.size() / 2];
    }
};

int main(int argc, char** argv) {
    int F = argc > 1 ? atoi(argv[1]) : 16;        // frames per launch
    int POOL = argc > 2 ? atoi(argv[2]) : 4;      // rotating input sets (defeats L2 reuse of inputs)
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s, %d SMs, L2 %d MB; F=%d frames/launch, pool=%d\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20, F, POOL);
    const int total = F * HW;
    const size_t in_elems = (size_t)F * 3 * HW, flow_elems = (size_t)F * 2 * HW, met_elems = (size_t)F * HW;
    float *in, *metric, *flow[3], *acc, *out;
    CK(cudaMalloc(&in, in_elems * 4 * POOL)); CK(cudaMalloc(&metric, met_elems * 4 * POOL));
    for (int i = 0; i < 3; ++i) CK(cudaMalloc(&flow[i], flow_elems * 4 * POOL));
    CK(cudaMalloc(&acc, (size_t)total * 16)); CK(cudaMalloc(&out, in_elems * 4 * POOL));
    CK(cudaMemset(in, 0, in_elems * 4 * POOL)); CK(cudaMemset(metric, 0, met_elems * 4 * POOL));
    for (int pat = 0; pat < 3; ++pat)
        for (int s = 0; s < POOL; ++s) k_make_flow<<<(total + 255) / 256, 256>>>(flow[pat] + s * flow_elems, F, pat);
    CK(cudaDeviceSynchronize());
    Timer T;
    const int blocks = (total + 255) / 256;
    const double px = (double)total;
    const char* pname[3] = {"zero", "smooth8", "rand32"};

    {   // calibration copy: same bytes as the algorithmic traffic of one soft forward (36 B/px): 18 B/px each way
        size_t n4 = (size_t)total * 18 / 16;
        float ms = T.run([&](int i) { k_copy<<<148 * 16, 512>>>((const float4*)in, (float4*)out, n4); }, 3, 20);
        printf("copy 36B/px-equivalent          : %8.1f us  %7.1f GB/s\n", ms * 1e3, px * 36 / ms / 1e6);
        n4 = in_elems * POOL / 4;
        ms = T.run([&](int i) { k_copy<<<148 * 16, 512>>>((const float4*)in, (float4*)out, n4); }, 3, 10);
        printf("copy big (%5.2f GB r+w)         : %8.1f us  %7.1f GB/s\n", n4 * 32 / 1e9, ms * 1e3, n4 * 32.0 / ms / 1e6);
    }
    {   // raw reduction / store probes on the accumulator buffer
        float ms = T.run([&](int i) { k_probe_st4<<<blocks, 256>>>(acc, total); }, 3, 20);
        printf("probe st.v4 (16 B/px)           : %8.1f us  %7.1f Gpx/s\n", ms * 1e3, px / ms / 1e6);
        ms = T.run([&](int i) { k_probe_red4<<<blocks, 256>>>(acc, total); }, 3, 20);
        printf("probe red.v4 x1 (coalesced)     : %8.1f us  %7.1f Gred/s\n", ms * 1e3, px / ms / 1e6);
        ms = T.run([&](int i) { k_probe_red1x4<<<blocks, 256>>>(acc, total); }, 3, 20);
        printf("probe red.f32 x4 planar         : %8.1f us  %7.1f Gred/s\n", ms * 1e3, px * 4 / ms / 1e6);
        if (F > 1) {
            ms = T.run([&](int i) { k_probe_red4<<<(HW + 255) / 256, 256>>>(acc, HW); }, 3, 20);
            printf("probe red.v4 x1, 1 frame (L2)   : %8.1f us  %7.1f Gred/s\n", ms * 1e3, (double)HW / ms / 1e6);
        }
    }
    for (int pat = 0; pat < 3; ++pat) {
        printf("--- flow pattern %s ---\n", pname[pat]);
        auto slot = [&](int i) { return (size_t)(i % POOL); };
        float ms;
        CK(cudaMemset(acc, 0, (size_t)total * 16));
        ms = T.run([&](int i) { size_t s = slot(i); k_scatter_v4<<<blocks, 256>>>(in + s * in_elems, metric + s * met_elems, flow[pat] + s * flow_elems, acc, total); }, 3, 20);
        printf("scatter red.v4   (all frames)   : %8.1f us  %7.1f Mpx/s\n", ms * 1e3, px / ms / 1e3);
        ms = T.run([&](int i) { size_t s = slot(i); k_scatter_v4_merge<<<blocks, 256>>>(in + s * in_elems, metric + s * met_elems, flow[pat] + s * flow_elems, acc, total); }, 3, 20);
        printf("scatter red.v4 + warp merge     : %8.1f us  %7.1f Mpx/s\n", ms * 1e3, px / ms / 1e3);
        ms = T.run([&](int i) { size_t s = slot(i); k_scatter_planar<<<blocks, 256>>>(in + s * in_elems, metric + s * met_elems, flow[pat] + s * flow_elems, acc, total); }, 3, 20);
        printf("scatter red.f32 planar          : %8.1f us  %7.1f Mpx/s\n", ms * 1e3, px / ms / 1e3);
        ms = T.run([&](int i) { size_t s = slot(i); k_norm_v4<<<blocks, 256>>>((float4*)acc, out + s * in_elems, total, 1); }, 3, 20);
        printf("normalise v4 + rezero           : %8.1f us  %7.1f Mpx/s\n", ms * 1e3, px / ms / 1e3);
        // whole forward, all frames per launch (accumulators stream through HBM when F is large)
        ms = T.run([&](int i) { size_t s = slot(i);
            k_scatter_v4<<<blocks, 256>>>(in + s * in_elems, metric + s * met_elems, flow[pat] + s * flow_elems, acc, total);
            k_norm_v4<<<blocks, 256>>>((float4*)acc, out + s * in_elems, total, 1); }, 3, 20);
        printf("fwd = scatter.v4 + norm, 1 wave : %8.1f us  %7.1f Mpx/s  %6.1f GB/s alg\n", ms * 1e3, px / ms / 1e3, px * 36 / ms / 1e6);
        ms = T.run([&](int i) { size_t s = slot(i);
            k_scatter_v4_merge<<<blocks, 256>>>(in + s * in_elems, metric + s * met_elems, flow[pat] + s * flow_elems, acc, total);
            k_norm_v4<<<blocks, 256>>>((float4*)acc, out + s * in_elems, total, 1); }, 3, 20);
        printf("fwd = merge.v4   + norm, 1 wave : %8.1f us  %7.1f Mpx/s  %6.1f GB/s alg\n", ms * 1e3, px / ms / 1e3, px * 36 / ms / 1e6);
        // whole forward in L2-sized waves of `wave` frames: accumulators (wave*33 MB) stay in L2
        for (int wave : {1, 2}) {
            if (wave > F) continue;
            const int wtotal = wave * HW, wblocks = (wtotal + 255) / 256;
            ms = T.run([&](int i) { size_t s = slot(i);
                for (int f0 = 0; f0 + wave <= F; f0 += wave) {
                    k_scatter_v4<<<wblocks, 256>>>(in + s * in_elems + (size_t)f0 * 3 * HW, metric + s * met_elems + (size_t)f0 * HW,
                                                   flow[pat] + s * flow_elems + (size_t)f0 * 2 * HW, acc, wtotal);
                    k_norm_v4<<<wblocks, 256>>>((float4*)acc, out + s * in_elems + (size_t)f0 * 3 * HW, wtotal, 1);
                } }, 3, 20);
            printf("fwd in waves of %d frame(s)      : %8.1f us  %7.1f Mpx/s  %6.1f GB/s alg\n", wave, ms * 1e3, px / ms / 1e3, px * 36 / ms / 1e6);
            ms = T.run([&](int i) { size_t s = slot(i);
                for (int f0 = 0; f0 + wave <= F; f0 += wave) {
                    k_scatter_v4_merge<<<wblocks, 256>>>(in + s * in_elems + (size_t)f0 * 3 * HW, metric + s * met_elems + (size_t)f0 * HW,
                                                   flow[pat] + s * flow_elems + (size_t)f0 * 2 * HW, acc, wtotal);
                    k_norm_v4<<<wblocks, 256>>>((float4*)acc, out + s * in_elems + (size_t)f0 * 3 * HW, wtotal, 1);
                } }, 3, 20);
            printf("merge fwd in waves of %d frame(s): %8.1f us  %7.1f Mpx/s  %6.1f GB/s alg\n", wave, ms * 1e3, px / ms / 1e3, px * 36 / ms / 1e6);
        }
        ms = T.run([&](int i) { size_t s = slot(i);
            k_scatter_planar<<<blocks, 256>>>(in + s * in_elems, metric + s * met_elems, flow[pat] + s * flow_elems, acc, total);
            k_norm_planar<<<blocks, 256>>>(acc, out + s * in_elems, total, 1); }, 3, 20);
        printf("fwd = planar     + norm, 1 wave : %8.1f us  %7.1f Mpx/s  %6.1f GB/s alg\n", ms * 1e3, px / ms / 1e3, px * 36 / ms / 1e6);
    }
    printf("done\n");
    return 0;
}
