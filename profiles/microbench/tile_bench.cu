// tile_bench.cu -- round-2 probes (not product code): can target cells be aggregated ON CHIP before they reach L2?
//
// The forward splat is bound by the reductions it sends to L2 (profiles/r01/NOTES.md sections 2 and 6; profiles/r02/NOTES.md
// sections 8 and 12). This binary times, one 1080p frame per launch into ONE L2-resident accumulator (the regime of the step
// pipeline), on a smooth, a rough (|dflow/dx| ~ 0.25, like bench.py) and a random flow:
//     naive 4 x red.v4 | east-merge by shuffle (what the product does, plus a vertical carry)
//   probe 1  k_window: a PRIVATE shared-memory window per warp and strip, fp32 shared atomics (CAS loops on sm_100a), every
//            touched cell flushed once as full sectors                                              -> loses 2-3x
//   probe 2  k_slide:  a window that SLIDES down a 32-column strip (ring of rows), plain ld/st.shared made conflict-free by
//            ranking lanes with match.any, ring rows flushed once when they leave the window       -> loses 1.7-2.1x
//   probe 3  k_bulk_red / k_lane_red: the flush alone -- cp.reduce.async.bulk (TMA reduce-add) of staged rows against one
//            red.v4 per lane over the same cells                                                   -> both ~4 TB/s of payload
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tile_bench tile_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 1080, W = 1920, HW = H * W;

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// pattern 0 smooth (|d/dx| ~ 0.03), 1 rough (~0.25, amplitude 8 px), 2 hash-random +-32 px
__global__ void k_make_flow(float* flow, int frames, int pattern) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= frames * HW) return;
    int n = p / HW, r = p - n * HW, y = r / W, x = r - y * W;
    float fx, fy;
    if (pattern == 0) {
        fx = 5.f * sinf(x * 0.0042f + y * 0.0026f + n) + 3.f * cosf(y * 0.0094f - x * 0.0018f);
        fy = 4.f * cosf(x * 0.0034f - y * 0.0038f + 2 * n) + 3.f * sinf(x * 0.0062f + 0.5f);
    } else if (pattern == 1) {
        fx = 5.f * sinf(x * 0.045f + y * 0.027f + n) + 3.f * cosf(y * 0.09f - x * 0.02f);
        fy = 4.f * cosf(x * 0.037f - y * 0.041f + 2 * n) + 3.f * sinf(x * 0.066f + 0.5f);
    } else {
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

// one frame; in [3,H,W], metric [H,W], flow [2,H,W]; acc [H*W][4]
__global__ void __launch_bounds__(256) k_naive(const float* __restrict__ in, const float* __restrict__ metric,
                                               const float* __restrict__ flow, float* acc) {
    int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= HW) return;
    int y = r / W, x = r - y * W;
    Foot f = foot(x, y, __ldcs(flow + r), __ldcs(flow + HW + r));
    float g = expf(__ldcs(metric + r));
    float v0 = __ldcs(in + r) * g, v1 = __ldcs(in + HW + r) * g, v2 = __ldcs(in + 2 * HW + r) * g;
    float* a = acc + ((size_t)f.y0 * W + f.x0) * 4;
    const int off[4] = {0, 4, 4 * W, 4 * W + 4};
#pragma unroll
    for (int k = 0; k < 4; ++k) if (f.b[k]) red4(a + off[k], v0 * f.w[k], v1 * f.w[k], v2 * f.w[k], g * f.w[k]);
}

__global__ void __launch_bounds__(256) k_merge_east(const float* __restrict__ in, const float* __restrict__ metric,
                                                    const float* __restrict__ flow, float* acc) {
    int r = blockIdx.x * 256 + threadIdx.x;
    bool live = r < HW;
    int rc = live ? r : HW - 1;
    int y = rc / W, x = rc - y * W;
    Foot f = foot(x, y, __ldcs(flow + rc), __ldcs(flow + HW + rc));
    float g = expf(__ldcs(metric + rc));
    float v[4] = {__ldcs(in + rc) * g, __ldcs(in + HW + rc) * g, __ldcs(in + 2 * HW + rc) * g, g};
    if (!live) f.b[0] = f.b[1] = f.b[2] = f.b[3] = false;
    const unsigned lane = threadIdx.x & 31;
    float westN[4], westS[4], eastN[4], eastS[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        westN[c] = f.b[0] ? v[c] * f.w[0] : 0.f; eastN[c] = f.b[1] ? v[c] * f.w[1] : 0.f;
        westS[c] = f.b[2] ? v[c] * f.w[2] : 0.f; eastS[c] = f.b[3] ? v[c] * f.w[3] : 0.f;
    }
    int lx0 = __shfl_up_sync(0xffffffffu, f.x0, 1), ly0 = __shfl_up_sync(0xffffffffu, f.y0, 1);
    bool take = lane > 0 && lx0 + 1 == f.x0 && ly0 == f.y0;
    bool given = __shfl_down_sync(0xffffffffu, (int)take, 1) && lane < 31;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float en = __shfl_up_sync(0xffffffffu, eastN[c], 1), es = __shfl_up_sync(0xffffffffu, eastS[c], 1);
        if (take) { westN[c] += en; westS[c] += es; }
    }
    bool vx0 = (unsigned)f.x0 < (unsigned)W, vy0 = (unsigned)f.y0 < (unsigned)H, vy1 = (unsigned)(f.y0 + 1) < (unsigned)H;
    bool vx1 = (unsigned)(f.x0 + 1) < (unsigned)W;
    float* a = acc + ((size_t)f.y0 * W + f.x0) * 4;
    if (live || take) {
        if (vx0 && vy0) red4(a, westN[0], westN[1], westN[2], westN[3]);
        if (vx0 && vy1) red4(a + 4 * W, westS[0], westS[1], westS[2], westS[3]);
    }
    if (live && !given) {
        if (vx1 && vy0) red4(a + 4, eastN[0], eastN[1], eastN[2], eastN[3]);
        if (vx1 && vy1) red4(a + 4 * W + 4, eastS[0], eastS[1], eastS[2], eastS[3]);
    }
}

// one warp per SX x SY strip (SX * SY / 32 pixels per lane), private WX x WY window of float4 cells kept as
// four planes (bank-conflict free for neighbouring cells)
template <int SX, int SY, int WX, int WY>
__global__ void __launch_bounds__(32) k_window(const float* __restrict__ in, const float* __restrict__ metric,
                                               const float* __restrict__ flow, float* acc, unsigned long long* spilled) {
    constexpr int CELLS = WX * WY, PER = SX * SY / 32, LY = 32 / SX;      // lanes cover SX columns x LY rows per pass
    __shared__ float win[4 * CELLS];
    const int lane = threadIdx.x;
    const int tiles_x = (W + SX - 1) / SX;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int lx = lane % SX, ly = lane / SX;
    for (int i = lane; i < 4 * CELLS; i += 32) win[i] = 0.f;
    Foot f[PER];
    float v[PER][4];
    int minx = 1 << 30, miny = 1 << 30;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
        const int x = tx * SX + lx, y = ty * SY + p * LY + ly;
        const bool in_img = x < W && y < H;
        const int r = in_img ? y * W + x : 0;
        f[p] = foot(x, y, __ldcs(flow + r), __ldcs(flow + HW + r));
        const float g = expf(__ldcs(metric + r));
        v[p][0] = __ldcs(in + r) * g; v[p][1] = __ldcs(in + HW + r) * g; v[p][2] = __ldcs(in + 2 * HW + r) * g; v[p][3] = g;
        if (!in_img) f[p].b[0] = f[p].b[1] = f[p].b[2] = f[p].b[3] = false;
        if (f[p].b[0] || f[p].b[1] || f[p].b[2] || f[p].b[3]) { minx = min(minx, f[p].x0); miny = min(miny, f[p].y0); }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, d));
        miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, d));
    }
    if (minx < 0) minx = 0;                                        // a corner at x0 = -1 is out of the frame anyway
    if (miny < 0) miny = 0;
    __syncwarp();
    unsigned nspill = 0;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!f[p].b[k]) continue;
            const int gx = f[p].x0 + (k & 1), gy = f[p].y0 + (k >> 1);
            const int cx = gx - minx, cy = gy - miny;
            const float wk = f[p].w[k];
            if ((unsigned)cx < (unsigned)WX && (unsigned)cy < (unsigned)WY) {
                float* c = win + cy * WX + cx;
                atomicAdd(c, v[p][0] * wk); atomicAdd(c + CELLS, v[p][1] * wk);
                atomicAdd(c + 2 * CELLS, v[p][2] * wk); atomicAdd(c + 3 * CELLS, v[p][3] * wk);
            } else {
                red4(acc + ((size_t)gy * W + gx) * 4, v[p][0] * wk, v[p][1] * wk, v[p][2] * wk, v[p][3] * wk);
                ++nspill;
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < CELLS; i += 32) {
        const float a0 = win[i], a1 = win[CELLS + i], a2 = win[2 * CELLS + i], a3 = win[3 * CELLS + i];
        if (a3 != 0.f) {                                           // the weight channel is positive wherever anything landed
            const int gx = minx + i % WX, gy = miny + i / WX;
            red4(acc + ((size_t)gy * W + gx) * 4, a0, a1, a2, a3);
        }
    }
    if (spilled && nspill) atomicAdd(spilled, (unsigned long long)nspill);
}


// ---------------------------------------------------------------------------------------------------------------------
// round-2 probe 2: SLIDING window.  One warp walks down a 32-column strip of SEG rows; its private window is a ring of R
// rows x WXC float4 cells that follows the walk (target rows [j, j + 2 MY] relative to the strip's landing row are live
// while source row j is processed).  Contributions are added with PLAIN ld.shared / st.shared: lanes whose footprints
// share a cell-key are ranked with match.any and take turns, so no instruction ever writes one address twice (no CAS
// loops).  A ring row that falls out of the window is flushed once: consecutive lanes red consecutive non-empty cells
// (full sectors), and zero them.  Pieces that land outside the window go straight to L2 (counted in `spilled`).
// ---------------------------------------------------------------------------------------------------------------------
template <int SEG, int MX, int MY, int R, int KR>
__global__ void __launch_bounds__(32) k_slide(const float* __restrict__ in, const float* __restrict__ metric,
                                              const float* __restrict__ flow, float* acc, unsigned long long* spilled) {
    constexpr int WXC = 32 + 2 * MX;
    static_assert((R & (R - 1)) == 0 && R >= 2 * MY + 2, "ring: power of two, at least one spare row");
    static_assert(SEG % KR == 0, "whole load groups");
    __shared__ float4 ring[R * WXC];
    const int lane = threadIdx.x;
    const unsigned full = 0xffffffffu;
    const int tiles_x = (W + 31) / 32;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x = tx * 32 + lane, y_first = ty * SEG;
    const bool xin = x < W;
    const int xs = xin ? x : W - 1;
    for (int i = lane; i < R * WXC; i += 32) ring[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // window offset: the flow in the middle of the strip
    int cx, cy;
    {
        const int ym = min(y_first + SEG / 2, H - 1), xm = min(tx * 32 + 16, W - 1);
        cx = (int)rintf(__ldg(flow + ym * W + xm)); cy = (int)rintf(__ldg(flow + HW + ym * W + xm));
    }
    const int colbase = tx * 32 + cx - MX;          // frame column of window column 0
    const int row0 = y_first + cy - MY;             // frame row of ring time 0
    unsigned nspill = 0;
    __syncwarp();

    auto flush = [&](int t) {                       // ring time t -> frame row row0 + t
        const int gy = row0 + t;
        float4* rowp = ring + (t & (R - 1)) * WXC;
#pragma unroll
        for (int it = 0; it < (WXC + 31) / 32; ++it) {
            const int c = lane + 32 * it;
            if (c < WXC) {
                const float4 v = rowp[c];
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
                    red4(acc + ((size_t)gy * W + (colbase + c)) * 4, v.x, v.y, v.z, v.w);
                    rowp[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    };

#pragma unroll 1
    for (int g = 0; g < SEG / KR; ++g) {
        const int yb = y_first + g * KR;
        if (yb >= H) break;
        float fxv[KR], fyv[KR], mv[KR], iv[KR][3];
#pragma unroll
        for (int r = 0; r < KR; ++r) {
            const int y = min(yb + r, H - 1);
            const int o = y * W + xs;
            fxv[r] = __ldcs(flow + o); fyv[r] = __ldcs(flow + HW + o); mv[r] = __ldcs(metric + o);
            iv[r][0] = __ldcs(in + o); iv[r][1] = __ldcs(in + HW + o); iv[r][2] = __ldcs(in + 2 * HW + o);
        }
#pragma unroll
        for (int r = 0; r < KR; ++r) {
            const int y = yb + r, j = g * KR + r;
            const bool live = xin && y < H;
            Foot f = foot(x, y, fxv[r], fyv[r]);
            const float gexp = expf(mv[r]);
            const float v[4] = {iv[r][0] * gexp, iv[r][1] * gexp, iv[r][2] * gexp, gexp};
            const bool any = live && (f.b[0] || f.b[1] || f.b[2] || f.b[3]);
            const int col = f.x0 - colbase, tr = f.y0 - row0;          // window column, ring time of the NW corner
            const bool inwin = any && col >= 0 && col <= WXC - 2 && tr >= j && tr + 1 <= j + 2 * MY;
            if (any && !inwin) {                                        // outside the window: straight to L2
                float* a = acc + ((size_t)f.y0 * W + f.x0) * 4;
                const int off[4] = {0, 4, 4 * W, 4 * W + 4};
#pragma unroll
                for (int k = 0; k < 4; ++k) if (f.b[k]) red4(a + off[k], v[0] * f.w[k], v[1] * f.w[k], v[2] * f.w[k], v[3] * f.w[k]);
                ++nspill;
            }
            // lanes that share a footprint take turns
            const int key = inwin ? (tr - j) * 64 + col : 4096 + lane;
            const unsigned same = __match_any_sync(full, key);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const int rounds = __reduce_max_sync(full, inwin ? rank : 0);
            float4* cN = ring + (tr & (R - 1)) * WXC + col;
            float4* cS = ring + ((tr + 1) & (R - 1)) * WXC + col;
            for (int rd = 0; rd <= rounds; ++rd) {
                const bool mine = inwin && rank == rd;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4* c = ((k >> 1) ? cS : cN) + (k & 1);
                    if (mine && f.b[k]) {
                        float4 a = *c;
                        a.x += v[0] * f.w[k]; a.y += v[1] * f.w[k]; a.z += v[2] * f.w[k]; a.w += v[3] * f.w[k];
                        *c = a;
                    }
                    __syncwarp();
                }
            }
            if (y < H) flush(j);                                        // ring time j is out of reach of every later row
            __syncwarp();
        }
    }
    const int rows_done = min(SEG, H - y_first);
    for (int t = rows_done; t < rows_done + 2 * MY + 1; ++t) flush(t);
    if (spilled && nspill) atomicAdd(spilled, (unsigned long long)nspill);
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    template <class F> float run(F f, int warm, int iters) {
        for (int i = 0; i < warm; ++i) f(i);
        CK(cudaEventRecord(a));
        for (int i = 0; i < iters; ++i) f(i);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        return ms / iters;
    }
};


// ---------------------------------------------------------------------------------------------------------------------
// round-2 probe 3: what does the FLUSH of a staged tile cost if the TMA unit does it?  One elected lane reduces a row of
// BYTES of shared memory into the L2-resident accumulator with cp.reduce.async.bulk (add.f32) -- whole lines, no per-lane
// address -- against one red.v4 per lane over the same cells (k_lane_red).  Both touch every cell of the frame exactly once.
// ---------------------------------------------------------------------------------------------------------------------
template <int BYTES>
__global__ void __launch_bounds__(32) k_bulk_red(float* acc, int rows_per_warp, size_t total_bytes) {
    __shared__ __align__(128) float buf[BYTES / 4];
    const int lane = threadIdx.x;
    for (int i = lane; i < BYTES / 4; i += 32) buf[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(buf);
        char* dst = (char*)acc + (size_t)blockIdx.x * rows_per_warp * BYTES;
        for (int r = 0; r < rows_per_warp; ++r) {
            if ((size_t)(dst - (char*)acc) + BYTES <= total_bytes)
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(BYTES) : "memory");
            dst += BYTES;
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) k_lane_red(float* acc) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r < HW) red4(acc + (size_t)r * 4, 1.f, 1.f, 1.f, 1.f);
}

template <int BYTES> static void bench_bulk(Timer& T, float* acc) {
    const size_t total = (size_t)HW * 16;
    const int rows_per_warp = 16;
    const int warps = (int)((total + (size_t)BYTES * rows_per_warp - 1) / ((size_t)BYTES * rows_per_warp));
    float ms = T.run([&](int) { k_bulk_red<BYTES><<<warps, 32>>>(acc, rows_per_warp, total); }, 4, 32);
    printf("  cp.reduce.async.bulk add.f32, %4d-byte rows: %7.1f us/frame  (%.0f GB/s of reduced payload)\n", BYTES, ms * 1e3, total / ms / 1e6);
}

template <int SX, int SY, int WX, int WY>
static void bench_window(Timer& T, const char* name, const float* in, const float* metric, const float* flow, float* acc,
                         unsigned long long* spilled, int POOL) {
    const int strips = ((W + SX - 1) / SX) * ((H + SY - 1) / SY);
    CK(cudaMemset(spilled, 0, 8));
    float ms = T.run([&](int i) { size_t s = i % POOL;
        k_window<SX, SY, WX, WY><<<strips, 32>>>(in + s * 3 * HW, metric + s * HW, flow + s * 2 * HW, acc, spilled); }, 4, 32);
    unsigned long long sp; CK(cudaMemcpy(&sp, spilled, 8, cudaMemcpyDeviceToHost));
    printf("  window %-22s: %7.1f us/frame   %.2f %% of corners spilled past the window\n", name, ms * 1e3, 100.0 * sp / 36.0 / (4.0 * HW));
}


template <int SEG, int MX, int MY, int R, int KR>
static void bench_slide(Timer& T, const char* name, const float* in, const float* metric, const float* flow, float* acc,
                        unsigned long long* spilled, int POOL) {
    const int strips = ((W + 31) / 32) * ((H + SEG - 1) / SEG);
    CK(cudaMemset(spilled, 0, 8));
    float ms = T.run([&](int i) { size_t s = i % POOL;
        k_slide<SEG, MX, MY, R, KR><<<strips, 32>>>(in + s * 3 * HW, metric + s * HW, flow + s * 2 * HW, acc, spilled); }, 4, 32);
    unsigned long long sp; CK(cudaMemcpy(&sp, spilled, 8, cudaMemcpyDeviceToHost));
    printf("  slide  %-22s: %7.1f us/frame   %.2f %% of pixels spilled past the window\n", name, ms * 1e3, 100.0 * sp / 36.0 / (1.0 * HW));
}

// correctness of k_slide against k_naive on one frame (max |difference| relative to max |value|)
template <int SEG, int MX, int MY, int R, int KR>
static void check_slide(const float* in, const float* metric, const float* flow, float* acc, float* acc2) {
    const int strips = ((W + 31) / 32) * ((H + SEG - 1) / SEG);
    CK(cudaMemset(acc, 0, (size_t)HW * 16)); CK(cudaMemset(acc2, 0, (size_t)HW * 16));
    k_naive<<<(HW + 255) / 256, 256>>>(in, metric, flow, acc);
    k_slide<SEG, MX, MY, R, KR><<<strips, 32>>>(in, metric, flow, acc2, nullptr);
    CK(cudaDeviceSynchronize());
    float* a = (float*)malloc((size_t)HW * 16); float* b = (float*)malloc((size_t)HW * 16);
    CK(cudaMemcpy(a, acc, (size_t)HW * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b, acc2, (size_t)HW * 16, cudaMemcpyDeviceToHost));
    double md = 0, mx = 0;
    for (size_t i = 0; i < (size_t)HW * 4; ++i) { md = fmax(md, fabs((double)a[i] - b[i])); mx = fmax(mx, fabs((double)a[i])); }
    printf("  check slide vs naive: max |diff| %.3g of max %.3g\n", md, mx);
    free(a); free(b);
    CK(cudaMemset(acc, 0, (size_t)HW * 16));
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("# device %s, %d SMs; one 1080p frame per launch into one 33 MB accumulator (L2-resident)\n", prop.name, prop.multiProcessorCount);
    const int POOL = 16;
    float *in, *metric, *flow[3], *acc; unsigned long long* spilled;
    CK(cudaMalloc(&in, (size_t)POOL * 3 * HW * 4)); CK(cudaMalloc(&metric, (size_t)POOL * HW * 4));
    for (int i = 0; i < 3; ++i) CK(cudaMalloc(&flow[i], (size_t)POOL * 2 * HW * 4));
    CK(cudaMalloc(&acc, (size_t)HW * 16)); CK(cudaMalloc(&spilled, 8));
    CK(cudaMemset(in, 0, (size_t)POOL * 3 * HW * 4)); CK(cudaMemset(metric, 0, (size_t)POOL * HW * 4)); CK(cudaMemset(acc, 0, (size_t)HW * 16));
    for (int pat = 0; pat < 3; ++pat) k_make_flow<<<(POOL * HW + 255) / 256, 256>>>(flow[pat], POOL, pat);
    k_make_flow<<<(POOL * HW + 255) / 256, 256>>>(metric, POOL / 2, 0);          // metric / input values: any smooth non-zero pattern
    k_make_flow<<<(POOL * HW + 255) / 256, 256>>>(in, POOL, 1);
    CK(cudaDeviceSynchronize());
    float* acc2; CK(cudaMalloc(&acc2, (size_t)HW * 16));
    Timer T;
    {
        printf("--- flush of a staged tile: every cell of the frame reduced once ---\n");
        float ms = T.run([&](int) { k_lane_red<<<(HW + 255) / 256, 256>>>(acc); }, 4, 32);
        printf("  one red.v4 per lane, consecutive cells        : %7.1f us/frame  (%.0f GB/s of reduced payload)\n", ms * 1e3, (double)HW * 16 / ms / 1e6);
        bench_bulk<256>(T, acc); bench_bulk<768>(T, acc); bench_bulk<2048>(T, acc); bench_bulk<8192>(T, acc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(acc, 0, (size_t)HW * 16));
    }
    const char* pname[3] = {"smooth (|d/dx| ~ 0.03)", "rough (|d/dx| ~ 0.25)", "random +-32 px"};
    const int blocks = (HW + 255) / 256;
    for (int pat = 0; pat < 3; ++pat) {
        printf("--- %s ---\n", pname[pat]);
        float ms = T.run([&](int i) { size_t s = i % POOL; k_naive<<<blocks, 256>>>(in + s * 3 * HW, metric + s * HW, flow[pat] + s * 2 * HW, acc); }, 4, 32);
        printf("  naive 4 x red.v4              : %7.1f us/frame\n", ms * 1e3);
        ms = T.run([&](int i) { size_t s = i % POOL; k_merge_east<<<blocks, 256>>>(in + s * 3 * HW, metric + s * HW, flow[pat] + s * 2 * HW, acc); }, 4, 32);
        printf("  east merge by shuffle         : %7.1f us/frame\n", ms * 1e3);
        check_slide<32, 8, 6, 16, 4>(in, metric, flow[pat], acc, acc2);
        bench_slide<32, 8, 6, 16, 4>(T, "32x32, mx 8 my 6", in, metric, flow[pat], acc, spilled, POOL);
        bench_slide<64, 8, 6, 16, 4>(T, "32x64, mx 8 my 6", in, metric, flow[pat], acc, spilled, POOL);
        bench_slide<32, 8, 3, 8, 4>(T, "32x32, mx 8 my 3", in, metric, flow[pat], acc, spilled, POOL);
        bench_slide<32, 4, 3, 8, 4>(T, "32x32, mx 4 my 3", in, metric, flow[pat], acc, spilled, POOL);
        bench_slide<32, 8, 6, 16, 8>(T, "32x32, mx 8 my 6, kr 8", in, metric, flow[pat], acc, spilled, POOL);
        bench_window<32, 4, 40, 8>(T, "32x4 strip, 40x8", in, metric, flow[pat], acc, spilled, POOL);
        bench_window<32, 4, 48, 12>(T, "32x4 strip, 48x12", in, metric, flow[pat], acc, spilled, POOL);
        bench_window<32, 4, 48, 20>(T, "32x4 strip, 48x20", in, metric, flow[pat], acc, spilled, POOL);
        bench_window<16, 8, 24, 16>(T, "16x8 strip, 24x16", in, metric, flow[pat], acc, spilled, POOL);
        bench_window<16, 8, 32, 20>(T, "16x8 strip, 32x20", in, metric, flow[pat], acc, spilled, POOL);
        bench_window<8, 16, 16, 24>(T, "8x16 strip, 16x24", in, metric, flow[pat], acc, spilled, POOL);
    }
    printf("done\n");
    return 0;
}
