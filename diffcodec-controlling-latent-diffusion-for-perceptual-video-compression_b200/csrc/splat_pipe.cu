// splat_pipe.cu -- the forward splat for C+1 <= 4 channels (frames, flows, SD latents) as ONE
// persistent, software-pipelined kernel (sm_100a).
//
// Why (measured on B200, profiles/r01/): L2 reductions retire at most one 32-byte sector per
// slice per clock (~360 G sectors/s), so the reference's 4 corner adds per pixel (~3 sectors/px on
// realistic flow) cap a scatter at ~120 Gpx/s, and fp32 accumulators that round-trip through HBM
// triple the DRAM traffic (only ~1.5 frames of 1080p accumulators stay L2-resident next to the
// streaming inputs). This kernel
//
//  1. stages each 256 x 8 source tile (flow, metric, channels) in SHARED MEMORY with TMA bulk
//     copies (cp.async.bulk + mbarrier, two stages per CTA): the loads of the next tile are in
//     flight while the current one is scattered, with no registers or issue slots spent on them;
//  2. merges corner contributions in REGISTERS before they reach L2: a warp owns 32 columns x 8
//     rows; the east column of lane i is handed to lane i+1 by shuffle when their footprints abut,
//     the south row of a pixel is carried to the next row of the same thread when they abut
//     vertically. Smooth flow -> ~1.2 `red.global.add.v4.f32` per pixel instead of 4; any flow
//     stays correct (unmatched pieces are simply issued alone);
//  3. keeps the fp32 accumulators L2-RESIDENT: frames go in order through a ring of `ring`
//     frame-sized accumulators; scatter tiles S(f) and normalise chunks N(f) are work items drawn
//     from one atomic ticket in the order  S0 | N0,S1 | N1,S2 | ...  so that every dependency
//     (N(f) after all of S(f); S(f+ring) after all of N(f)) points at EARLIER tickets: waiting
//     CTAs only ever wait for CTAs that are already running -> no co-residency requirement, no
//     cooperative launch, no deadlock;
//  4. is ONE launch for any number of frames: no memset (N re-zeroes what it read), no launch
//     gaps; normalise of frame f overlaps scatter of frame f+1 on the same SMs.
//
// Replaces controlnet/softsplat.py:240-270 (pre/post ops) + :281-345 (zero-init + softsplat_out).
#include "dcb_common.cuh"

#include <stdlib.h>

namespace dcb {

constexpr int kPipeThreads = 256;
constexpr int kTW = 256, kTH = 8;           // source tile: one column per thread, 8 rows per warp
constexpr int kPlaneElems = kTW * kTH;
constexpr int kMaxPlanes = 6;               // flow_x, flow_y, up to 4 value planes (C values [+ metric])
constexpr int kStageBytes = kMaxPlanes * kPlaneElems * 4;
constexpr int kChunk = 2048;                // target pixels per normalise item
constexpr int kCtrlWords = 64;              // [0] ticket, [1] exit count; then done_s[N], done_n[N]

struct PipeArgs {
    View in, flow, metric, mask;
    float* acc;              // ring * HW * 4 floats, all-zero on entry and on exit
    unsigned* ctrl;
    void* out;               // [N,C,H,W]
    void* norm;              // [N,1,H,W] fp32 or null
    int N, C, H, W;
    unsigned HW;
    int eps;
    int ring;
    int tiles_x, ts, tn;     // scatter tiles per row / per frame, normalise chunks per frame
    unsigned total_items;
    int bulk;                // inputs are row-contiguous and 16-byte aligned: TMA bulk staging
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk async copy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy (TMA, no tensor map): 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_count(const unsigned* p, unsigned want) {
    if (threadIdx.x == 0) {
        unsigned ns = 32;
        while (ld_acquire(p) < want) {
            __nanosleep(ns);
            if (ns < 512) ns *= 2;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void signal_done(unsigned* p) {
    // bar.sync orders every thread's reds / stores before thread 0's gpu-scope fence (the fence is
    // cumulative), so one MEMBAR per CTA publishes the whole item -- the grid-sync idiom
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// work-item decoding: tickets in the order S0 .. S(L-1) | N0,S(L) | N1,S(L+1) | ... | N tail
// ---------------------------------------------------------------------------------------------
struct Item { bool scatter; int frame, idx; };

__device__ __forceinline__ Item decode(const PipeArgs& a, unsigned t) {
    Item it;
    const int L = a.ring - 1;                                            // normalise lags scatter by L frames
    const unsigned n1 = (unsigned)min(a.N, L) * a.ts;                    // S-only groups
    const unsigned per = (unsigned)(a.ts + a.tn);
    const unsigned n2 = (unsigned)max(a.N - L, 0) * per;                 // groups holding N(q) and S(q+L)
    if (t < n1) { it.scatter = true; it.frame = t / a.ts; it.idx = t - it.frame * a.ts; }
    else if (t - n1 < n2) {
        t -= n1;
        // With L > 0 the normalise items go first, which keeps every dependency one stage away;
        // with L == 0 the scatter items must precede the normalise items that wait for them.
        const unsigned q = t / per, r = t - q * per;
        const unsigned first = L > 0 ? (unsigned)a.tn : (unsigned)a.ts;
        const bool in_first = r < first;
        it.scatter = (L > 0) ? !in_first : in_first;
        it.idx = (int)(in_first ? r : r - first);
        it.frame = it.scatter ? (int)q + L : (int)q;
    } else {
        t -= n1 + n2;
        const unsigned i = t / a.tn;
        it.scatter = false; it.frame = max(a.N - L, 0) + (int)i; it.idx = (int)(t - i * a.tn);
    }
    return it;
}

// staged planes: 0 flow_x, 1 flow_y, 2.. values, then metric
template <int MODE> __device__ __forceinline__ int plane_count(int C) { return 2 + C + (MODE >= DCB_MODE_LINEAR ? 1 : 0); }

// ---------------------------------------------------------------------------------------------
// staging
// ---------------------------------------------------------------------------------------------
template <class T, class TF, int MODE>
__device__ __forceinline__ void stage_issue_bulk(const PipeArgs& a, int frame, int tile, unsigned char* stage,
                                                 unsigned long long* bar, int lane, unsigned long long pol) {
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    const int x0 = tx * kTW, y0 = ty * kTH;
    const int cols = min(kTW, a.W - x0), rows = min(kTH, a.H - y0);
    const int planes = plane_count<MODE>(a.C);
    const unsigned bytes_v = (unsigned)cols * sizeof(T), bytes_f = (unsigned)cols * sizeof(TF);
    if (lane == 0) mbar_expect_tx(bar, (unsigned)rows * (2u * bytes_f + (unsigned)(planes - 2) * bytes_v));
    __syncwarp();
    const int copies = planes * rows;
    for (int i = lane; i < copies; i += 32) {
        const int p = i / rows, r = i - p * rows;
        const long long y = y0 + r;
        unsigned char* dst = stage + (size_t)p * (kPlaneElems * 4);
        if (p < 2) {
            const TF* src = (const TF*)a.flow.p + frame * a.flow.sN + p * a.flow.sC + y * a.flow.sH + x0;
            bulk_g2s(dst + (size_t)r * kTW * sizeof(TF), src, bytes_f, bar, pol);
        } else if (p < 2 + a.C) {
            const T* src = (const T*)a.in.p + frame * a.in.sN + (p - 2) * a.in.sC + y * a.in.sH + x0;
            bulk_g2s(dst + (size_t)r * kTW * sizeof(T), src, bytes_v, bar, pol);
        } else {
            const T* src = (const T*)a.metric.p + frame * a.metric.sN + y * a.metric.sH + x0;
            bulk_g2s(dst + (size_t)r * kTW * sizeof(T), src, bytes_v, bar, pol);
        }
    }
}

// any strides / alignment: all threads copy the tile with ordinary loads (synchronous)
template <class T, class TF, int MODE>
__device__ __forceinline__ void stage_generic(const PipeArgs& a, int frame, int tile, unsigned char* stage) {
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    const int x = tx * kTW + threadIdx.x, y0 = ty * kTH;
    const int rows = min(kTH, a.H - y0);
    if (x < a.W) {
        for (int r = 0; r < rows; ++r) {
            const long long y = y0 + r;
            const int e = r * kTW + threadIdx.x;
            const TF* fp = (const TF*)a.flow.p + frame * a.flow.sN + y * a.flow.sH + x * a.flow.sW;
            ((TF*)stage)[e] = fp[0];
            ((TF*)(stage + kPlaneElems * 4))[e] = fp[a.flow.sC];
            const T* ip = (const T*)a.in.p + frame * a.in.sN + y * a.in.sH + x * a.in.sW;
            for (int c = 0; c < a.C; ++c) ((T*)(stage + (size_t)(2 + c) * kPlaneElems * 4))[e] = ip[c * a.in.sC];
            if (MODE >= DCB_MODE_LINEAR)
                ((T*)(stage + (size_t)(2 + a.C) * kPlaneElems * 4))[e] =
                    ((const T*)a.metric.p)[frame * a.metric.sN + y * a.metric.sH + x * a.metric.sW];
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// scatter of one staged tile
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red4_if(bool p, float* acc, int off, const float (&v)[4]) {
    if (p) red_add_v4(acc + (long long)off * 4, v[0], v[1], v[2], v[3]);
}

template <class T, class TF, int MODE>
__device__ __forceinline__ void scatter_tile(const PipeArgs& a, int tile, const unsigned char* stage, float* acc) {
    const int lane = threadIdx.x & 31;
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    const int x = tx * kTW + threadIdx.x, yb = ty * kTH;
    const int W = a.W, H = a.H, C = a.C;
    const int rows = min(kTH, H - yb);
    const bool xin = x < W;
    const unsigned full = 0xffffffffu;
    const int kDead = -7;
    const int pitch = W + 2;                       // key = (y0 + 1) * pitch + (x0 + 1) identifies a footprint
    const int CA = C + (MODE != DCB_MODE_SUM ? 1 : 0);

    const TF* sfx = (const TF*)stage + threadIdx.x;
    const TF* sfy = (const TF*)(stage + kPlaneElems * 4) + threadIdx.x;
    const T* sv = (const T*)(stage + 2 * kPlaneElems * 4) + threadIdx.x;
    const T* sm = (const T*)(stage + (size_t)(2 + C) * kPlaneElems * 4) + threadIdx.x;
    constexpr int kPlaneT = kPlaneElems * 4 / sizeof(T);   // plane pitch in elements of T

    float pend[4] = {0.f, 0.f, 0.f, 0.f};
    int pend_key = kDead, pend_off = 0;
    bool pend_ok = false;

#pragma unroll
    for (int r = 0; r < kTH; ++r) {
        const bool in_img = xin && r < rows;
        float flx = 0.f, fly = 0.f;
        if (in_img) { flx = ld<float>(sfx + r * kTW); fly = ld<float>(sfy + r * kTW); }
        const float fx = add_rn((float)x, flx), fy = add_rn((float)(yb + r), fly);   // softsplat.py:298-299
        const float x0f = floorf(fx), y0f = floorf(fy);
        const int x0 = __float2int_rz(x0f), y0 = __float2int_rz(y0f);
        // finite landing point (softsplat.py:301-302) with at least one corner inside the frame
        const bool alive = in_img && fabsf(fx) < 3.0e38f && fabsf(fy) < 3.0e38f &&
                           ((unsigned)x0 + 1u) <= (unsigned)W && ((unsigned)y0 + 1u) <= (unsigned)H;
        const float ex = sub_rn(add_rn(x0f, 1.f), fx), ey = sub_rn(add_rn(y0f, 1.f), fy);  // softsplat.py:315-318
        const float dx = sub_rn(fx, x0f), dy = sub_rn(fy, y0f);
        const float wnw = mul_rn(ex, ey), wne = mul_rn(dx, ey), wsw = mul_rn(ex, dy), wse = mul_rn(dx, dy);

        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (in_img) {
            float g = 1.f;
            if (MODE == DCB_MODE_LINEAR) g = ld<float>(sm + r * kTW);
            if (MODE == DCB_MODE_SOFT) g = expf(ld<float>(sm + r * kTW));
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < C) {
                    const float t = ld<float>(sv + c * kPlaneT + r * kTW);
                    v[c] = (MODE >= DCB_MODE_LINEAR) ? mul_rn(t, g) : t;                  // softsplat.py:244,247
                } else if (c == C && MODE != DCB_MODE_SUM) {
                    v[c] = g;                                                             // appended channel
                }
            }
        }
        float nw[4], ne[4], sw[4], se[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            nw[c] = mul_rn(v[c], wnw); ne[c] = mul_rn(v[c], wne);
            sw[c] = mul_rn(v[c], wsw); se[c] = mul_rn(v[c], wse);
        }
        const int key = alive ? (y0 + 1) * pitch + (x0 + 1) : kDead;
        const int off = y0 * W + x0;
        const bool vx0 = x0 >= 0, vx1 = x0 < W - 1, vy0 = y0 >= 0, vy1 = y0 < H - 1;

        // ---- horizontal hand-over: my east column goes to lane + 1 if our footprints abut ----
        const int lkey = __shfl_up_sync(full, key, 1);
        const bool take = lane > 0 && alive && lkey != kDead && lkey + 1 == key;
        const bool given = (__shfl_down_sync(full, (int)take, 1) != 0) && lane < 31;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < CA) {
                const float en = __shfl_up_sync(full, ne[c], 1), es = __shfl_up_sync(full, se[c], 1);
                if (take) { nw[c] = add_rn(nw[c], en); sw[c] = add_rn(sw[c], es); }
            }
        }
        const bool east = alive && !given && vx1;
        red4_if(east && vy0, acc, off + 1, ne);
        red4_if(east && vy1, acc, off + W + 1, se);
        // ---- vertical carry: the previous row's south piece joins my north piece if they abut ----
        const bool join = pend_key == key && alive;            // kDead never equals a live key
        if (join) {
#pragma unroll
            for (int c = 0; c < 4; ++c) nw[c] = add_rn(nw[c], pend[c]);
        }
        red4_if(pend_ok && !join, acc, pend_off, pend);
        red4_if(alive && vx0 && vy0, acc, off, nw);
#pragma unroll
        for (int c = 0; c < 4; ++c) pend[c] = sw[c];
        pend_key = alive ? key + pitch : kDead;
        pend_off = off + W;
        pend_ok = alive && vx0 && vy1;
    }
    red4_if(pend_ok, acc, pend_off, pend);
}

// ---------------------------------------------------------------------------------------------
// normalise chunk: eps rule, divide, (1 - mask), cast, save normaliser, re-zero
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float eps_rule(float d, int eps) {
    if (eps == DCB_EPS_ADD) return add_rn(d, 0.0000001f);                // softsplat.py:257,260
    if (eps == DCB_EPS_ZERO) return d == 0.f ? 1.f : d;                  // :263
    return d < 0.0000001f ? 0.0000001f : d;                              // :266
}

template <class T, int MODE>
__device__ __forceinline__ void normalize_chunk(const PipeArgs& a, int frame, int chunk, float* acc) {
    const unsigned base = (unsigned)chunk * kChunk + threadIdx.x;
    T* out = (T*)a.out + (long long)frame * a.C * a.HW;
    constexpr int kPer = kChunk / kPipeThreads;                          // 8 pixels per thread
    float4 s[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {                                     // all L2 reads in flight first
        const unsigned r = base + i * kPipeThreads;
        s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < a.HW) s[i] = __ldcg((const float4*)acc + r);             // written by other SMs' reds: L2 is the point of coherence
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const unsigned r = base + i * kPipeThreads;
        if (r >= a.HW) break;
        __stcg((float4*)acc + r, make_float4(0.f, 0.f, 0.f, 0.f));       // accumulators leave the kernel all-zero
        const float sv[4] = {s[i].x, s[i].y, s[i].z, s[i].w};
        float d = 1.f;
        if (MODE != DCB_MODE_SUM) {
            d = eps_rule(a.C == 3 ? sv[3] : (a.C == 2 ? sv[2] : (a.C == 1 ? sv[1] : sv[0])), a.eps);
            if (a.norm) __stcs((float*)a.norm + (long long)frame * a.HW + r, d);
        }
        float keep = 1.f;
        if (a.mask.p) {
            const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
            const T* mp = (const T*)a.mask.p + frame * a.mask.sN + (long long)y * a.mask.sH + (long long)x * a.mask.sW;
            keep = sub_rn(1.f, ld<float>(mp));                           // control_utils.py:69-70
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < a.C) {
                float o = (MODE != DCB_MODE_SUM) ? sv[c] / d : sv[c];    // true division, as the reference (softsplat.py:270)
                if (a.mask.p) o = mul_rn(o, keep);
                st<T, float>(out + (long long)c * a.HW + r, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <class T, class TF, int MODE>
__global__ void __launch_bounds__(kPipeThreads, 2) k_splat_pipe(const __grid_constant__ PipeArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];               // 2 stages of kStageBytes
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ unsigned s_tk[2];
    __shared__ unsigned s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* done_s = a.ctrl + kCtrlWords;
    unsigned* done_n = done_s + a.N;
    unsigned long long pol = 0;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0 owns the ticket counter and the TMA issue; the ticket of the NEXT item is drawn (and
    // its tile put in flight) before the current item is processed
    auto fetch = [&](int slot) {
        unsigned t = 0;
        if (lane == 0) { t = atomicAdd(a.ctrl, 1u); s_tk[slot] = t; }
        t = __shfl_sync(0xffffffffu, t, 0);
        if (a.bulk && t < a.total_items) {
            const Item it = decode(a, t);
            if (it.scatter) stage_issue_bulk<T, TF, MODE>(a, it.frame, it.idx, smem + (size_t)slot * kStageBytes, &s_bar[slot], lane, pol);
        }
    };
    if (warp == 0) { pol = policy_evict_first(); fetch(0); }
    __syncthreads();

    unsigned parity[2] = {0u, 0u};
    for (unsigned n = 0;; ++n) {
        const int slot = n & 1;
        const unsigned t = s_tk[slot];
        if (t >= a.total_items) break;
        if (warp == 0) fetch(slot ^ 1);                                  // stage slot^1 was released by the barrier that ended item n-1
        const Item it = decode(a, t);
        float* acc = a.acc + (size_t)(it.frame % a.ring) * a.HW * 4;
        if (it.scatter) {
            unsigned char* stage = smem + (size_t)slot * kStageBytes;
            if (a.bulk) { mbar_wait(&s_bar[slot], parity[slot]); parity[slot] ^= 1u; }
            else stage_generic<T, TF, MODE>(a, it.frame, it.idx, stage);
            if (it.frame >= a.ring) wait_count(done_n + (it.frame - a.ring), (unsigned)a.tn);   // ring slot is free again
            scatter_tile<T, TF, MODE>(a, it.idx, stage, acc);
            signal_done(done_s + it.frame);
        } else {
            wait_count(done_s + it.frame, (unsigned)a.ts);               // every source of the frame has landed
            normalize_chunk<T, MODE>(a, it.frame, it.idx, acc);
            signal_done(done_n + it.frame);
        }
    }
    // The last CTA to leave puts the control block back to all-zero, so the whole workspace
    // (accumulators AND control words) is clean again when the kernel ends.
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.ctrl + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int i = threadIdx.x; i < kCtrlWords + 2 * a.N; i += kPipeThreads) a.ctrl[i] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRing = 3;

static int pipe_ring(long long N) {
    static int cap = 0;
    if (cap == 0) {                                   // DCB_PIPE_RING=1|2|3: experiments only
        const char* e = getenv("DCB_PIPE_RING");
        cap = e ? atoi(e) : 2;
        if (cap < 1 || cap > kMaxRing) cap = 2;
    }
    return N >= cap ? cap : (int)(N < 1 ? 1 : N);
}

// sized for the largest ring so that the workspace query does not depend on the environment
long long pipe_acc_bytes(long long N, long long H, long long W) {
    const long long slots = N >= kMaxRing ? kMaxRing : (N < 1 ? 1 : N);
    return align_up(slots * H * W * 16, 256);
}

long long pipe_workspace(long long N, long long H, long long W) {
    return pipe_acc_bytes(N, H, W) + align_up((kCtrlWords + 2 * N) * 4, 256);
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <class T> static bool bulk_ok(const DcbTensor* t, long long W) {
    if (!t) return true;
    const long long es = sizeof(T);
    return t->stride[3] == 1 && aligned16(t->ptr) && (t->stride[0] * es) % 16 == 0 && (t->stride[1] * es) % 16 == 0 &&
           (t->stride[2] * es) % 16 == 0 && (W * es) % 16 == 0;
}

template <class T, class TF, int MODE> static int launch_pipe_mode(const PipeArgs& a, cudaStream_t st) {
    static int grid_cap = 0;
    const size_t smem = 2 * (size_t)kStageBytes;
    if (grid_cap == 0) {
        DCB_CHECK_CUDA(cudaFuncSetAttribute(k_splat_pipe<T, TF, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_splat_pipe<T, TF, MODE>, kPipeThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        grid_cap = device_sm_count() * per_sm;
    }
    const int grid = (int)(a.total_items < (unsigned)grid_cap ? a.total_items : (unsigned)grid_cap);
    k_splat_pipe<T, TF, MODE><<<grid, kPipeThreads, smem, st>>>(a);
    DCB_CHECK_LAUNCH("k_splat_pipe");
    return DCB_OK;
}

template <class T, class TF> static int launch_pipe(const PipeArgs& a, int mode, cudaStream_t st) {
    switch (mode) {
        case DCB_MODE_SUM: return launch_pipe_mode<T, TF, DCB_MODE_SUM>(a, st);
        case DCB_MODE_AVG: return launch_pipe_mode<T, TF, DCB_MODE_AVG>(a, st);
        case DCB_MODE_LINEAR: return launch_pipe_mode<T, TF, DCB_MODE_LINEAR>(a, st);
        default: return launch_pipe_mode<T, TF, DCB_MODE_SOFT>(a, st);
    }
}

// Preconditions (checked by the caller): C + (mode != SUM) <= 4, dtype F32/BF16, workspace >= pipe_workspace().
int splat_pipe_impl(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric, const DcbTensor* out,
                    const DcbTensor* norm, const DcbTensor* mask, void* ws, int mode, int eps, bool ws_clean,
                    cudaStream_t st) {
    PipeArgs a;
    a.in = make_view(in); a.flow = make_view(flow); a.metric = make_view(metric); a.mask = make_view(mask);
    a.N = (int)in->size[0]; a.C = (int)in->size[1]; a.H = (int)in->size[2]; a.W = (int)in->size[3];
    a.HW = (unsigned)(in->size[2] * in->size[3]);
    a.eps = eps;
    a.ring = pipe_ring(a.N);
    a.tiles_x = (a.W + kTW - 1) / kTW;
    a.ts = a.tiles_x * ((a.H + kTH - 1) / kTH);
    a.tn = (int)((a.HW + kChunk - 1) / kChunk);
    a.total_items = (unsigned)((long long)a.N * (a.ts + a.tn));
    a.out = out->ptr;
    a.norm = norm ? norm->ptr : nullptr;
    const long long acc_bytes = pipe_acc_bytes(a.N, a.H, a.W);
    a.acc = (float*)ws;
    a.ctrl = (unsigned*)((char*)ws + acc_bytes);
    if (!ws_clean) DCB_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)pipe_workspace(a.N, a.H, a.W), st));
    const bool ff = flow->dtype == DCB_F32;
    if (in->dtype == DCB_F32) {
        a.bulk = bulk_ok<float>(in, a.W) && bulk_ok<float>(flow, a.W) && bulk_ok<float>(metric, a.W);
        return launch_pipe<float, float>(a, mode, st);
    }
    if (in->dtype == DCB_BF16) {
        a.bulk = bulk_ok<__nv_bfloat16>(in, a.W) && bulk_ok<__nv_bfloat16>(metric, a.W) &&
                 (ff ? bulk_ok<float>(flow, a.W) : bulk_ok<__nv_bfloat16>(flow, a.W));
        return ff ? launch_pipe<__nv_bfloat16, float>(a, mode, st) : launch_pipe<__nv_bfloat16, __nv_bfloat16>(a, mode, st);
    }
    return set_error(DCB_E_DTYPE, "splat_pipe: unsupported dtype %d", in->dtype);
}

}  // namespace dcb
