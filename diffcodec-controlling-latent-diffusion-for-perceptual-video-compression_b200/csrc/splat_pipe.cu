// splat_pipe.cu -- the forward splat for C+1 <= 4 channels (frames, flows, SD latents) as ONE
// persistent, software-pipelined kernel (sm_100a).
//
// Why (measured on B200, profiles/r01/red_bench_F16_v0.log): L2 reductions retire ~5 TB/s of
// payload however they are vectorised, so a scatter that issues the reference's 4 corner adds per
// pixel (64 B of red payload for 36 B of compulsory HBM traffic) is atomics-bound at ~80 Gpx/s,
// and fp32 accumulators that round-trip through HBM triple the DRAM traffic. This kernel
//
//  1. merges corner contributions in REGISTERS before they reach L2: a warp owns 32 columns x
//     ROWS rows of source pixels; the east column of lane i is handed to lane i+1 by shuffle when
//     their footprints abut (x0+1 == x0', y0 == y0'), the south row of a pixel is carried to the
//     next row of the same thread when they abut vertically. Smooth flow -> ~1.2 `red.v4` per
//     pixel instead of 4; any flow stays correct (unmatched pieces are simply issued alone);
//     pieces whose four products are all +-0 (integer flows) are dropped -- adding +-0 to an
//     accumulator that starts at +0 never changes a bit;
//  2. keeps the fp32 accumulators L2-RESIDENT: frames are processed in order through a ring of
//     `ring` frame-sized accumulators; scatter tiles S(f) and normalise chunks N(f) are work items
//     drawn from one atomic ticket in the order  S0 | S1 | N0,S2 | N1,S3 | ...  so that every
//     dependency (N(f) after all of S(f); S(f+ring) after all of N(f)) points at EARLIER tickets:
//     waiting CTAs only ever wait for CTAs that are already running -> no co-residency
//     requirement, no cooperative launch, no deadlock;
//  3. is ONE launch for any number of frames: no memset (N re-zeroes what it read), no launch
//     gaps, normalise of frame f overlaps scatter of frame f+2 on the same SMs.
//
// Replaces controlnet/softsplat.py:240-270 (pre/post ops) + :281-345 (zero-init + softsplat_out).
#include "dcb_common.cuh"

namespace dcb {

constexpr int kPipeThreads = 256;
constexpr int kRows = 4;                    // rows loaded at once by a warp
constexpr int kPasses = 2;                  // consecutive row groups per warp (the vertical carry spans them)
constexpr int kTileW = 128, kTileH = 2 * kRows * kPasses;
constexpr int kChunk = 2048;                // target pixels per normalise item
constexpr int kCtrlWords = 64;              // ticket + padding, then done_s[N], done_n[N]

struct PipeArgs {
    View in, flow, metric, mask;
    float* acc;              // ring * HW * 4 floats, all-zero on entry and on exit
    unsigned* ctrl;          // [0] ticket; [1] exit count; [kCtrlWords + f] done_s; [kCtrlWords + N + f] done_n
    void* out;               // [N,C,H,W]
    void* norm;              // [N,1,H,W] fp32 or null
    int N, C, H, W;
    unsigned HW;
    int mode, eps;
    int ring;
    int tiles_x, ts, tn;     // scatter tiles per row / per frame, normalise chunks per frame
    unsigned total_items;
};

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
            if (ns < 1024) ns *= 2;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void signal_done(unsigned* p) {
    // bar.sync orders every thread's reds / stores before thread 0's gpu-scope fence (the fence is
    // cumulative), so one MEMBAR per CTA publishes the whole tile -- the grid-sync idiom
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p, 1u);
    }
}

__device__ __forceinline__ void emit(float* acc, int W, int H, int ty, int tx, const float (&v)[4]) {
    if ((unsigned)tx < (unsigned)W && (unsigned)ty < (unsigned)H &&
        (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f || v[3] != 0.f))
        red_add_v4(acc + ((long long)ty * W + tx) * 4, v[0], v[1], v[2], v[3]);
}

template <class T, class TF>
__device__ __forceinline__ void scatter_tile(const PipeArgs& a, int frame, int tile, float* acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    const int x = tx * kTileW + (warp & 3) * 32 + lane;
    const int W = a.W, H = a.H;
    const bool xin = x < W;
    const unsigned full = 0xffffffffu;
    float pend[4] = {0.f, 0.f, 0.f, 0.f};
    int pend_x = 0, pend_y = 0;
    bool pend_valid = false;

#pragma unroll 1
  for (int pass = 0; pass < kPasses; ++pass) {
    const int y0row = ty * kTileH + ((warp >> 2) * kPasses + pass) * kRows;
    if (y0row >= H) break;                                               // warp-uniform

    // ---- issue every load of the warp's ROWS rows up front (memory-level parallelism) ----
    float fxv[kRows], fyv[kRows], mv[kRows], iv[kRows][3];
    const TF* fbase = (const TF*)a.flow.p + (long long)frame * a.flow.sN + (long long)x * a.flow.sW;
    const T* ibase = (const T*)a.in.p + (long long)frame * a.in.sN + (long long)x * a.in.sW;
    const T* mbase = a.metric.p ? (const T*)a.metric.p + (long long)frame * a.metric.sN + (long long)x * a.metric.sW : nullptr;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int y = y0row + r;
        const bool in_img = xin && y < H;
        fxv[r] = fyv[r] = 0.f; mv[r] = 0.f;
        iv[r][0] = iv[r][1] = iv[r][2] = 0.f;
        if (in_img) {
            const TF* fp = fbase + (long long)y * a.flow.sH;
            fxv[r] = ld<float>(fp); fyv[r] = ld<float>(fp + a.flow.sC);
            if (mbase) mv[r] = ld<float>(mbase + (long long)y * a.metric.sH);
            const T* ip = ibase + (long long)y * a.in.sH;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < a.C) iv[r][c] = ld<float>(ip + c * a.in.sC);
            if (a.C == 4) mv[r] = ld<float>(ip + 3 * a.in.sC);          // SUM with 4 channels: 4th value rides in mv
        }
    }

#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int y = y0row + r;
        const Foot<float> f = make_foot<float>(x, y, fxv[r], fyv[r]);
        // alive: inside the image, finite landing point (softsplat.py:301-302), some corner in range
        const bool alive = xin && y < H && f.finite && f.x0 >= -1 && f.x0 < W && f.y0 >= -1 && f.y0 < H;
        float v[4];
        {
            float g = 1.f;
            if (a.mode == DCB_MODE_LINEAR) g = mv[r];
            else if (a.mode == DCB_MODE_SOFT) g = expf(mv[r]);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = (a.mode >= DCB_MODE_LINEAR) ? mul_rn(iv[r][c], g) : iv[r][c];
            v[3] = 0.f;
            if (a.mode == DCB_MODE_SUM) { if (a.C == 4) v[3] = mv[r]; }
            else {                                                       // appended channel sits at index C
                if (a.C == 3) v[3] = g; else if (a.C == 2) v[2] = g; else if (a.C == 1) v[1] = g; else v[0] = g;
            }
        }
        float nw[4], ne[4], sw[4], se[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            nw[c] = alive ? mul_rn(v[c], f.wnw) : 0.f;
            ne[c] = alive ? mul_rn(v[c], f.wne) : 0.f;
            sw[c] = alive ? mul_rn(v[c], f.wsw) : 0.f;
            se[c] = alive ? mul_rn(v[c], f.wse) : 0.f;
        }
        // ---- horizontal hand-over: my east column goes to lane+1 if our footprints abut ----
        const int lx0 = __shfl_up_sync(full, f.x0, 1), ly0 = __shfl_up_sync(full, f.y0, 1);
        const int lalive = __shfl_up_sync(full, (int)alive, 1);
        const bool take = lane > 0 && alive && lalive && (lx0 + 1 == f.x0) && (ly0 == f.y0);
        const bool given = (__shfl_down_sync(full, (int)take, 1) != 0) && lane < 31;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float en = __shfl_up_sync(full, ne[c], 1), es = __shfl_up_sync(full, se[c], 1);
            if (take) { nw[c] = add_rn(nw[c], en); sw[c] = add_rn(sw[c], es); }
        }
        if (alive && !given) {
            emit(acc, W, H, f.y0, f.x0 + 1, ne);
            emit(acc, W, H, f.y0 + 1, f.x0 + 1, se);
        }
        // ---- vertical carry: the previous row's south piece joins my north piece if they abut ----
        if (pend_valid) {
            if (alive && pend_x == f.x0 && pend_y == f.y0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) nw[c] = add_rn(nw[c], pend[c]);
            } else {
                emit(acc, W, H, pend_y, pend_x, pend);
            }
        }
        if (alive) emit(acc, W, H, f.y0, f.x0, nw);
#pragma unroll
        for (int c = 0; c < 4; ++c) pend[c] = sw[c];
        pend_x = f.x0; pend_y = f.y0 + 1; pend_valid = alive;
    }
  }
    if (pend_valid) emit(acc, W, H, pend_y, pend_x, pend);
}

__device__ __forceinline__ float eps_rule(float d, int eps) {
    if (eps == DCB_EPS_ADD) return add_rn(d, 0.0000001f);                // softsplat.py:257,260
    if (eps == DCB_EPS_ZERO) return d == 0.f ? 1.f : d;                  // :263
    return d < 0.0000001f ? 0.0000001f : d;                              // :266
}

template <class T>
__device__ __forceinline__ void normalize_chunk(const PipeArgs& a, int frame, int chunk, float* acc) {
    const unsigned base = (unsigned)chunk * kChunk;
    T* out = (T*)a.out + (long long)frame * a.C * a.HW;
    const bool normalised = a.mode != DCB_MODE_SUM;
    constexpr int kPer = kChunk / kPipeThreads;                          // 8 pixels per thread
    // all loads first (8 independent L2 reads in flight per thread), then the arithmetic
    float4 s[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const unsigned r = base + i * kPipeThreads + threadIdx.x;
        s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < a.HW) s[i] = __ldcg((const float4*)acc + r);             // written by other SMs' reds: L2 is the point of coherence
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const unsigned r = base + i * kPipeThreads + threadIdx.x;
        if (r >= a.HW) break;
        __stcg((float4*)acc + r, make_float4(0.f, 0.f, 0.f, 0.f));       // accumulators leave the kernel all-zero
        const float sv[4] = {s[i].x, s[i].y, s[i].z, s[i].w};
        float d = 1.f;
        if (normalised) {
            d = eps_rule(a.C == 3 ? sv[3] : (a.C == 2 ? sv[2] : (a.C == 1 ? sv[1] : sv[0])), a.eps);
            if (a.norm) __stcs((float*)a.norm + (long long)frame * a.HW + r, d);
        }
        float keep = 1.f;
        if (a.mask.p) {
            const int y = (int)(r / (unsigned)a.W), x = (int)(r - (unsigned)y * (unsigned)a.W);
            const T* mp = (const T*)a.mask.p + (long long)frame * a.mask.sN + (long long)y * a.mask.sH + (long long)x * a.mask.sW;
            keep = sub_rn(1.f, ld<float>(mp));                           // control_utils.py:69-70
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < a.C) {
                float o = normalised ? sv[c] / d : sv[c];
                if (a.mask.p) o = mul_rn(o, keep);
                st<T, float>(out + (long long)c * a.HW + r, o);
            }
        }
    }
}

template <class T, class TF>
__global__ void __launch_bounds__(kPipeThreads, 4) k_splat_pipe(const PipeArgs a) {
    __shared__ unsigned s_ticket;
    unsigned* done_s = a.ctrl + kCtrlWords;
    unsigned* done_n = done_s + a.N;
    const int L = a.ring - 1;                                            // normalise lags scatter by L frames
    const unsigned n1 = (unsigned)min(a.N, L) * a.ts;                    // S-only groups
    const unsigned per = (unsigned)(a.ts + a.tn);
    const unsigned n2 = (unsigned)max(a.N - L, 0) * per;                 // N(g-L), S(g) groups
    unsigned next = 0;
    if (threadIdx.x == 0) next = atomicAdd(a.ctrl, 1u);
    for (;;) {
        if (threadIdx.x == 0) s_ticket = next;
        __syncthreads();
        unsigned t = s_ticket;
        __syncthreads();
        if (t >= a.total_items) break;
        if (threadIdx.x == 0) next = atomicAdd(a.ctrl, 1u);              // prefetch: its latency hides behind this item
        bool is_scatter;
        int frame, idx;
        if (t < n1) { is_scatter = true; frame = t / a.ts; idx = t - frame * a.ts; }
        else if (t - n1 < n2) {
            t -= n1;
            // group q holds N(q) and S(q+L). With L > 0 the normalise items go first, which keeps every
            // dependency at least one full stage away (S(q+L) waits for N(q+L-ring) = N(q-1)); with
            // L == 0 (one frame) the scatter items must precede the normalise items that wait for them
            const unsigned q = t / per, r = t - q * per;
            const unsigned first = L > 0 ? (unsigned)a.tn : (unsigned)a.ts;
            const bool in_first = r < first;
            is_scatter = (L > 0) ? !in_first : in_first;
            idx = (int)(in_first ? r : r - first);
            frame = is_scatter ? (int)q + L : (int)q;
        } else {
            t -= n1 + n2;
            const unsigned i = t / a.tn;
            is_scatter = false; frame = max(a.N - L, 0) + (int)i; idx = (int)(t - i * a.tn);
        }
        float* acc = a.acc + (size_t)(frame % a.ring) * a.HW * 4;
        if (is_scatter) {
            if (frame >= a.ring) wait_count(done_n + (frame - a.ring), (unsigned)a.tn);   // ring slot is free again
            scatter_tile<T, TF>(a, frame, idx, acc);
            signal_done(done_s + frame);
        } else {
            wait_count(done_s + frame, (unsigned)a.ts);                  // every source of the frame has landed
            normalize_chunk<T>(a, frame, idx, acc);
            signal_done(done_n + frame);
        }
    }
    // The last CTA to leave puts the control block back to all-zero, so the whole workspace
    // (accumulators AND control words) is clean again when the kernel ends.
    __shared__ unsigned s_last;
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
static int pipe_ring(long long N) { return N >= 3 ? 3 : (int)(N < 1 ? 1 : N); }

long long pipe_acc_bytes(long long N, long long H, long long W) { return align_up((long long)pipe_ring(N) * H * W * 16, 256); }

long long pipe_workspace(long long N, long long H, long long W) {
    return pipe_acc_bytes(N, H, W) + align_up((kCtrlWords + 2 * N) * 4, 256);
}

template <class T, class TF> static int pipe_grid() {
    static int grid = 0;
    if (grid == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_splat_pipe<T, TF>, kPipeThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        grid = device_sm_count() * per_sm;
    }
    return grid;
}

template <class T, class TF> static int launch_pipe(const PipeArgs& a, cudaStream_t st) {
    const long long items = a.total_items;
    const int grid = (int)(items < pipe_grid<T, TF>() ? items : pipe_grid<T, TF>());
    k_splat_pipe<T, TF><<<grid, kPipeThreads, 0, st>>>(a);
    DCB_CHECK_LAUNCH("k_splat_pipe");
    return DCB_OK;
}

// Preconditions (checked by the caller): C + (mode != SUM) <= 4, dtype F32/BF16, workspace >= pipe_workspace().
int splat_pipe_impl(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric, const DcbTensor* out,
                    const DcbTensor* norm, const DcbTensor* mask, void* ws, int mode, int eps, bool ws_clean,
                    cudaStream_t st) {
    PipeArgs a;
    a.in = make_view(in); a.flow = make_view(flow); a.metric = make_view(metric); a.mask = make_view(mask);
    a.N = (int)in->size[0]; a.C = (int)in->size[1]; a.H = (int)in->size[2]; a.W = (int)in->size[3];
    a.HW = (unsigned)(in->size[2] * in->size[3]);
    a.mode = mode; a.eps = eps;
    a.ring = pipe_ring(a.N);
    a.tiles_x = (a.W + kTileW - 1) / kTileW;
    a.ts = a.tiles_x * ((a.H + kTileH - 1) / kTileH);
    a.tn = (int)((a.HW + kChunk - 1) / kChunk);
    a.total_items = (unsigned)((long long)a.N * (a.ts + a.tn));
    a.out = out->ptr;
    a.norm = norm ? norm->ptr : nullptr;
    const long long acc_bytes = pipe_acc_bytes(a.N, a.H, a.W);
    a.acc = (float*)ws;
    a.ctrl = (unsigned*)((char*)ws + acc_bytes);
    if (!ws_clean) DCB_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)pipe_workspace(a.N, a.H, a.W), st));
    const bool ff = flow->dtype == DCB_F32;
    if (in->dtype == DCB_F32) return launch_pipe<float, float>(a, st);
    if (in->dtype == DCB_BF16)
        return ff ? launch_pipe<__nv_bfloat16, float>(a, st) : launch_pipe<__nv_bfloat16, __nv_bfloat16>(a, st);
    return set_error(DCB_E_DTYPE, "splat_pipe: unsupported dtype %d", in->dtype);
}

}  // namespace dcb
