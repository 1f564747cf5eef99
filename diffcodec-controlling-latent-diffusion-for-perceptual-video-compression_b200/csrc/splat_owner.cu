// splat_owner.cu -- forward splat with TARGET-TILE OWNERSHIP (sm_100a): the accumulators never
// leave the SM.
//
// Round 1's pipeline (splat_pipe.cu) accumulates with `red.global.add.v4.f32` into fp32 cells in
// L2 and normalises them in a second pass; ncu showed that design bound by what it moves through
// L2 (reduction sectors + accumulator read + re-zero: ~320 MB per 75 MB frame) and by its
// instruction count (profiles/r01/NOTES.md section 6). Here a CTA OWNS a
// 64 x 32 tile of TARGET cells:
//
//   1. a pre-pass (k_strip_box) reads the flow once and records, per 32 x 4 source strip, the
//      bounding box of the cells its pixels land on, and per strip row the box of the row;
//   2. the owner CTA lists the strips whose box meets its tile (two-level test, a few dozen
//      candidates), re-reads their flow (L2 hits: the pre-pass of the same frame group just ran)
//      and, for the pixels that land in the tile, the input channels and the metric;
//   3. contributions are accumulated into a shared-memory window with PLAIN ld/st.shared -- fp32
//      shared atomics are CAS loops on sm_100a (measured 2-3x slower, profiles/r01/NOTES.md
//      section 8). Conflict freedom is by construction: every pixel writes its id into
//      owner[cell(x0, y0)]; the pixels that read their own id back ("winners") have pairwise
//      distinct north-west cells, hence pairwise distinct NW, NE, SW and SE cells, so the four
//      corners are added in four barrier-separated phases without any atomic. The few losers
//      (two sources with the same integer landing cell: only where the flow compresses) add with
//      shared CAS atomics in a fifth phase;
//   4. the epilogue normalises straight out of shared memory (eps rule, reciprocal, (1 - mask),
//      cast, saved normaliser, or the occlusion test) and writes `out` exactly once.
//
// Global traffic = flow (twice, second time from L2) + inputs (x ~1.05 tile-border overlap) +
// outputs; no accumulator, no memset, no workspace-clean protocol, no normalise pass. Frames of
// up to 512 strips (latents, pyramid levels) skip the pre-pass: each CTA derives the boxes of
// its frame itself, so the whole call is ONE launch.
//
// Replaces controlnet/softsplat.py:240-270 (pre/post ops) + :281-345 (zero-init + softsplat_out).
#include "dcb_common.cuh"

#include <stdlib.h>

namespace dcb {

#ifndef DCB_OW_THREADS
#define DCB_OW_THREADS 64
#endif
#ifndef DCB_OW_TW
#define DCB_OW_TW 64
#endif
#ifndef DCB_OW_TH
#define DCB_OW_TH 8
#endif
#ifndef DCB_OW_MINCTAS
#define DCB_OW_MINCTAS 12
#endif
constexpr int kOT = DCB_OW_THREADS;      // threads per CTA
constexpr int kOWarps = kOT / 32;
constexpr int kTW = DCB_OW_TW, kTH = DCB_OW_TH;   // target tile owned by a CTA
static_assert(kOT % kTW == 0 && (kTW & (kTW - 1)) == 0, "epilogue mapping: a row of the tile is a power-of-two slice of the CTA");
constexpr int kPW = kTW + 1, kPH = kTH + 1;   // window with one guard column / row: cell(x0 - tx0 + 1, y0 - ty0 + 1)
constexpr int kCells = (kPH + 1) * kPW + 7;   // a hit pixel's four corners never leave the window: no per-corner bounds tests
constexpr int kSR = 4;                   // rows of a source strip (32 x 4 pixels, one warp)
constexpr int kCandCap = 512;            // candidate strips per pass
constexpr int kRowCap = 256;             // candidate strip rows per pass
constexpr int kSelfStrips = 512;         // frames with at most this many strips need no pre-pass
constexpr float kExp1o = 2.7182817459106445f;   // expf(1.0f)

struct OwnerArgs {
    View in, flow, metric, mask;
    void* out;               // [N,C,H,W]
    void* norm;              // [N,1,H,W] fp32 or null
    short4* strip_box;       // [N][strip_rows][strips_x]: (x0 min, x0 max, y0 min, y0 max) of the strip's pixels
    short4* row_box;         // [N][strip_rows]: (y0 min, y0 max, max reach to the left, max reach to the right)
    int N, C, H, W;
    unsigned HW;
    int eps, ones, epi;      // epi: 0 normalise, 1 occlusion mask
    int tiles_x, tiles_y, strips_x, strip_rows;
    int frame0;              // first frame of this launch (frame groups)
    int self_box;            // every CTA derives the boxes of its frame itself (single launch)
    View epi_flow;
    void* mask_out;
};

__device__ __forceinline__ int floor_div32(int a) { return a >> 5; }      // arithmetic shift = floor for negatives

// landing cell of one source pixel (softsplat.py:298-302); ok = finite
__device__ __forceinline__ bool landing(int x, int y, float flx, float fly, float& fx, float& fy, float& x0f, float& y0f, int& x0, int& y0) {
    fx = add_rn((float)x, flx); fy = add_rn((float)y, fly);
    x0f = floorf(fx); y0f = floorf(fy);
    x0 = __float2int_rz(x0f); y0 = __float2int_rz(y0f);
    return fabsf(fx) < 3.0e38f && fabsf(fy) < 3.0e38f;      // false for NaN and +-Inf
}

// box of one 32 x kSR strip: warp-collective, result in every lane
template <class TF>
__device__ __forceinline__ void strip_box(const View& flow, int frame, int row, int sx, int lane, int H, int W,
                                          int& xmin, int& xmax, int& ymin, int& ymax) {
    const int x = sx * 32 + lane;
    xmin = ymin = 32767; xmax = ymax = -32768;
    if (x < W) {
        const TF* fb = (const TF*)flow.p + (long long)frame * flow.sN + (long long)x * flow.sW;
        float flx[kSR], fly[kSR];
#pragma unroll
        for (int r = 0; r < kSR; ++r) {
            const int y = row * kSR + r;
            flx[r] = fly[r] = __int_as_float(0x7fc00000);
            if (y < H) { const TF* fp = fb + (long long)y * flow.sH; flx[r] = ld<float>(fp); fly[r] = ld<float>(fp + flow.sC); }
        }
#pragma unroll
        for (int r = 0; r < kSR; ++r) {
            float fx, fy, x0f, y0f; int x0, y0;
            const bool ok = landing(x, row * kSR + r, flx[r], fly[r], fx, fy, x0f, y0f, x0, y0);
            if (ok && x0 >= -1 && x0 < W && y0 >= -1 && y0 < H) {          // at least one corner inside the frame
                xmin = min(xmin, x0); xmax = max(xmax, x0); ymin = min(ymin, y0); ymax = max(ymax, y0);
            }
        }
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
}

__device__ __forceinline__ short clamp16(int v) { return (short)max(-32768, min(32767, v)); }

// boxes of one strip row: CTA-collective (NW warps). sbox: NW * 4 ints of shared scratch.
template <class TF, int NW>
__device__ __forceinline__ void row_boxes(const View& flow, int frame, int row, int H, int W, int strips_x,
                                          short4* strip_out, short4* row_out, int* sbox) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int rymin = 32767, rymax = -32768, rdl = -32768, rdr = -32768;
#pragma unroll 2
    for (int sx = warp; sx < strips_x; sx += NW) {
        int xmin, xmax, ymin, ymax;
        strip_box<TF>(flow, frame, row, sx, lane, H, W, xmin, xmax, ymin, ymax);
        if (lane == 0) strip_out[sx] = make_short4((short)xmin, (short)xmax, (short)ymin, (short)ymax);
        if (xmin <= xmax) {
            rymin = min(rymin, ymin); rymax = max(rymax, ymax);
            rdl = max(rdl, sx * 32 - xmin); rdr = max(rdr, xmax - (sx * 32 + 31));
        }
    }
    if (lane == 0) { sbox[warp * 4 + 0] = rymin; sbox[warp * 4 + 1] = rymax; sbox[warp * 4 + 2] = rdl; sbox[warp * 4 + 3] = rdr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NW; ++w) {
            rymin = min(rymin, sbox[w * 4 + 0]); rymax = max(rymax, sbox[w * 4 + 1]);
            rdl = max(rdl, sbox[w * 4 + 2]); rdr = max(rdr, sbox[w * 4 + 3]);
        }
        *row_out = make_short4(clamp16(rymin), clamp16(rymax), clamp16(rdl), clamp16(rdr));
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// pre-pass: one CTA per (strip row, frame)
// ---------------------------------------------------------------------------------------------
constexpr int kPreThreads = 1024;        // a whole strip row of a 1080p frame (60 strips) is in flight at once

template <class TF>
__global__ void __launch_bounds__(kPreThreads) k_strip_box(const __grid_constant__ OwnerArgs a) {
    // The first pre-pass of a call waits for whatever precedes it in the stream (the producer of the flow). Later
    // ones follow an owner launch of the same call, which only triggers after ITS wait: by induction everything
    // older is complete, and the owner launch itself writes nothing this kernel reads -- so they start while it drains.
    if (a.frame0 == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ int sbox[kPreThreads / 32 * 4];
    const int row = blockIdx.x, f = a.frame0 + blockIdx.y;
    const size_t rb = (size_t)f * a.strip_rows + row;
    row_boxes<TF, kPreThreads / 32>(a.flow, f, row, a.H, a.W, a.strips_x, a.strip_box + rb * a.strips_x, a.row_box + rb, sbox);
}

// ---------------------------------------------------------------------------------------------
// owner kernel. NV = float4 vectors per cell (CA <= 4 * NV accumulated channels).
// ---------------------------------------------------------------------------------------------
constexpr int kBatch = 4 * kOT;          // records per accumulate batch: four per thread
constexpr int kQ = 2 * kBatch;           // record ring (power of two): < kBatch left over + one scan step of <= kOWarps * 128 = kBatch hits
static_assert((kQ & (kQ - 1)) == 0, "ring size");
constexpr int kDummy = kCells;           // scratch cells for idle lanes (dummy + 1, + kPW, + kPW + 1 are scratch too)
constexpr int kWinCells = kCells + kPW + 2;

template <int NV> struct OwnerSmem {
    float4 win[NV][kWinCells];           // the tile's accumulators (+ guard row / column, + scratch cells)
    unsigned short owner[kWinCells + 2];
    unsigned q_yx[kQ];                   // hit records: source pixel (y << 16 | x) ...
    float q_fx[kQ], q_fy[kQ];            // ... and its landing point
    int cand[kCandCap];                  // candidate strips: row << 12 | sx  (strips_x <= 4096)
    int rows[kRowCap * 2];               // candidate strip rows: (row, first sx | count << 16)
    int n_cand, n_rows, q_tail, n_lost, sbox[kOWarps * 4];
};

// ---- stage A: one warp scans one 32 x kSR strip and appends the pixels that land in the tile to the ring ----
template <class TF, int NV>
__device__ __forceinline__ void scan_strip(const OwnerArgs& a, OwnerSmem<NV>& s, int f, int cd, int tx0, int ty0, int lane) {
    const int row = cd >> 12, sx = cd & 0xfff;
    const int x = sx * 32 + lane;
    const TF* fb = (const TF*)a.flow.p + (long long)f * a.flow.sN + x * (int)a.flow.sW;
    const int f_sH = (int)a.flow.sH, f_sC = (int)a.flow.sC;
    float flx[kSR], fly[kSR];
#pragma unroll
    for (int r = 0; r < kSR; ++r) {
        const int y = row * kSR + r;
        flx[r] = fly[r] = __int_as_float(0x7fc00000);
        if (x < a.W && y < a.H) { const TF* fp = fb + y * f_sH; flx[r] = ld<float>(fp); fly[r] = ld<float>(fp + f_sC); }
    }
    float fx[kSR], fy[kSR];
    unsigned bal[kSR];
    int total = 0;
#pragma unroll
    for (int r = 0; r < kSR; ++r) {
        float x0f, y0f; int x0, y0;
        const bool ok = landing(x, row * kSR + r, flx[r], fly[r], fx[r], fy[r], x0f, y0f, x0, y0);
        // NW corner in [tx0 - 1, tx0 + kTW - 1] x [ty0 - 1, ty0 + kTH - 1]: at least one corner is a cell of this tile
        const bool hit = ok && (unsigned)(x0 - tx0 + 1) <= (unsigned)kTW && (unsigned)(y0 - ty0 + 1) <= (unsigned)kTH;
        bal[r] = __ballot_sync(0xffffffffu, hit);
        total += __popc(bal[r]);
    }
    if (total == 0) return;                                              // warp-uniform
    int base = 0;
    if (lane == 0) base = atomicAdd(&s.q_tail, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSR; ++r) {
        if (bal[r] >> lane & 1u) {
            const int i = (base + __popc(bal[r] & lt)) & (kQ - 1);
            s.q_yx[i] = (unsigned)(row * kSR + r) << 16 | (unsigned)x; s.q_fx[i] = fx[r]; s.q_fy[i] = fy[r];
        }
        base += __popc(bal[r]);
    }
}

// ---- stage B: one accumulate batch over records [head, head + nb) of the ring; returns nothing, re-queues losers ----
template <class T, int MODE, int CA, int NV>
__device__ __forceinline__ void accumulate_batch(const OwnerArgs& a, OwnerSmem<NV>& s, int f, int head, int nb, int tx0, int ty0, int tid, bool last) {
    constexpr int C = CA - (MODE != DCB_MODE_SUM ? 1 : 0);
    constexpr int R = kBatch / kOT;
    const int lane = tid & 31;
    const T* ibase = (const T*)a.in.p + (long long)f * a.in.sN;
    const T* mbase = (MODE >= DCB_MODE_LINEAR && !a.ones) ? (const T*)a.metric.p + (long long)f * a.metric.sN : nullptr;
    const int i_sH = (int)a.in.sH, i_sW = (int)a.in.sW, i_sC = (int)a.in.sC, m_sH = (int)a.metric.sH, m_sW = (int)a.metric.sW;
    int key[R];
    unsigned yx[R];
    float fx[R], fy[R], v[R][CA], mv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {                                        // every load of the batch in flight before the first use
        const int k = j * kOT + tid;
        key[j] = -1; yx[j] = 0; fx[j] = fy[j] = 0.f; mv[j] = 0.f;
#pragma unroll
        for (int c = 0; c < CA; ++c) v[j][c] = 0.f;
        if (k < nb) {
            const int i = (head + k) & (kQ - 1);
            yx[j] = s.q_yx[i]; fx[j] = s.q_fx[i]; fy[j] = s.q_fy[i];
            const int y = (int)(yx[j] >> 16), x = (int)(yx[j] & 0xffffu);
            const T* ip = ibase + (y * i_sH + x * i_sW);
#pragma unroll
            for (int c = 0; c < C; ++c) v[j][c] = ld_stream(ip + c * i_sC);
            if (MODE >= DCB_MODE_LINEAR && !a.ones) mv[j] = ld_stream(mbase + (y * m_sH + x * m_sW));
            key[j] = 0;
        }
    }
    float ex[R], ey[R], dx[R], dy[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const float x0f = floorf(fx[j]), y0f = floorf(fy[j]);
        // softsplat.py:315-318
        ex[j] = sub_rn(add_rn(x0f, 1.f), fx[j]); ey[j] = sub_rn(add_rn(y0f, 1.f), fy[j]);
        dx[j] = sub_rn(fx[j], x0f); dy[j] = sub_rn(fy[j], y0f);
        if (key[j] == 0) key[j] = (__float2int_rz(y0f) - ty0 + 1) * kPW + (__float2int_rz(x0f) - tx0 + 1);
    }
    // ---- claim: one winner per landing cell ----
#pragma unroll
    for (int j = 0; j < R; ++j)
        if (key[j] >= 0) s.owner[key[j]] = (unsigned short)(j * kOT + tid);
    if (tid == 0) s.n_lost = 0;
    __syncthreads();
    // pre-op while the claim settles: g = 1 | m | exp(m); in * g; appended channel (softsplat.py:240-247)
#pragma unroll
    for (int j = 0; j < R; ++j) {
        float g = 1.f;
        if (MODE == DCB_MODE_LINEAR) g = a.ones ? 1.f : mv[j];
        if (MODE == DCB_MODE_SOFT) g = a.ones ? kExp1o : expf(mv[j]);
        if (MODE >= DCB_MODE_LINEAR) {
#pragma unroll
            for (int c = 0; c < C; ++c) v[j][c] = mul_rn(v[j][c], g);
        }
        if (MODE != DCB_MODE_SUM) v[j][C] = key[j] >= 0 ? g : 0.f;
    }
    bool lost[R];
    int n_lost = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        lost[j] = key[j] >= 0 && s.owner[key[j]] != (unsigned short)(j * kOT + tid);
        n_lost += lost[j] ? 1 : 0;
    }
    {   // how many records lost their claim (decides between re-queueing and the CAS fallback)
        int w = n_lost;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0 && w > 0) atomicAdd(&s.n_lost, w);
    }
    // ---- four corner phases: the winners' cells are pairwise distinct within a phase; idle lanes add 0 to scratch cells ----
    int cell[R];
#pragma unroll
    for (int j = 0; j < R; ++j) cell[j] = (key[j] >= 0 && !lost[j]) ? key[j] : kDummy;
#pragma unroll
    for (int corner = 0; corner < 4; ++corner) {
        const int off = (corner & 1) + (corner >> 1) * kPW;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const float w = mul_rn((corner & 1) ? dx[j] : ex[j], (corner >> 1) ? dy[j] : ey[j]);
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                float4 cv = s.win[q][cell[j] + off];
                if (4 * q + 0 < CA) cv.x = fma_rn(v[j][4 * q + 0 < CA ? 4 * q + 0 : 0], w, cv.x);
                if (4 * q + 1 < CA) cv.y = fma_rn(v[j][4 * q + 1 < CA ? 4 * q + 1 : 0], w, cv.y);
                if (4 * q + 2 < CA) cv.z = fma_rn(v[j][4 * q + 2 < CA ? 4 * q + 2 : 0], w, cv.z);
                if (4 * q + 3 < CA) cv.w = fma_rn(v[j][4 * q + 3 < CA ? 4 * q + 3 : 0], w, cv.w);
                s.win[q][cell[j] + off] = cv;
            }
        }
        __syncthreads();
    }
    // ---- losers (another source of this batch shares their landing cell) ----
    const int all_lost = s.n_lost;
    if (all_lost == 0) return;                                           // CTA-uniform
    if (!last && all_lost * 4 <= nb * 3) {
        // back into the ring: they are retried, densely packed, with the next batch (the slots of this batch are free)
        const unsigned bl = __ballot_sync(0xffffffffu, n_lost > 0);
        if (bl) {
            int mine = n_lost, pre = 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, mine, o); if (lane >= o) mine += t; }
            pre = mine - n_lost;                                         // exclusive prefix of this lane
            int base = 0;
            if (lane == 31) base = atomicAdd(&s.q_tail, mine);
            base = __shfl_sync(0xffffffffu, base, 31) + pre;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (lost[j]) { const int i = base++ & (kQ - 1); s.q_yx[i] = yx[j]; s.q_fx[i] = fx[j]; s.q_fy[i] = fy[j]; }
            }
        }
    } else {
        // the last, partial batch of a tile (re-queueing its losers would cost a whole batch each time they halve), or nearly
        // everything collides (e.g. a whole frame flowing into one cell): shared CAS atomics finish the job regardless
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (lost[j]) {
#pragma unroll
                for (int corner = 0; corner < 4; ++corner) {
                    const int off = (corner & 1) + (corner >> 1) * kPW;
                    const float w = mul_rn((corner & 1) ? dx[j] : ex[j], (corner >> 1) ? dy[j] : ey[j]);
#pragma unroll
                    for (int c = 0; c < CA; ++c)
                        atomicAdd(reinterpret_cast<float*>(&s.win[c >> 2][key[j] + off]) + (c & 3), mul_rn(v[j][c], w));
                }
            }
        }
    }
    __syncthreads();
}

template <class T, class TF, int MODE, int CA>
__global__ void __launch_bounds__(kOT, (CA <= 4 ? DCB_OW_MINCTAS : (DCB_OW_MINCTAS + 1) / 2)) k_splat_owner(const __grid_constant__ OwnerArgs a) {
    constexpr int NV = (CA + 3) / 4;
    constexpr int C = CA - (MODE != DCB_MODE_SUM ? 1 : 0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OwnerSmem<NV>& s = *reinterpret_cast<OwnerSmem<NV>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H;

    const int tiles = a.tiles_x * a.tiles_y;
    const int f = a.frame0 + blockIdx.x / tiles, t = blockIdx.x % tiles;
    const int tx0 = (t % a.tiles_x) * kTW, ty0 = (t / a.tiles_x) * kTH;

    // Programmatic dependent launch. Wait FIRST (everything before this launch in the stream is then complete:
    // the pre-pass that wrote the boxes, and whoever produced the inputs), then let the next launch start:
    // the pre-pass of the next frame group skips its own wait and relies on exactly this order.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const short4* strip_box_g = a.strip_box + (size_t)f * a.strip_rows * a.strips_x;
    const short4* row_box_g = a.row_box + (size_t)f * a.strip_rows;
    if (a.self_box) {
        // small frame: derive every box of the frame here, into the (not yet used) window
        short4* sb = reinterpret_cast<short4*>(&s.win[0][0]);
        short4* rbx = sb + kSelfStrips;
        for (int row = 0; row < a.strip_rows; ++row)
            row_boxes<TF, kOWarps>(a.flow, f, row, H, W, a.strips_x, sb + row * a.strips_x, rbx + row, s.sbox);
        strip_box_g = sb; row_box_g = rbx;            // generic pointers into shared memory
    }

    bool win_clean = false;
    int head = 0;                                     // records consumed so far (CTA-uniform); s.q_tail counts the appended ones
    if (tid == 0) s.q_tail = 0;
    for (int row0 = 0; row0 < a.strip_rows; row0 += kRowCap) {
        // ---- level 1: strip rows whose box meets the tile ----
        if (tid == 0) s.n_rows = 0;
        __syncthreads();
        for (int row = row0 + tid; row < min(row0 + kRowCap, a.strip_rows); row += kOT) {
            const short4 rb = row_box_g[row];         // (y0 min, y0 max, reach left, reach right)
            if (rb.y >= ty0 - 1 && rb.x <= ty0 + kTH - 1) {
                int lo = floor_div32(tx0 - 32 - rb.w + 31), hi = floor_div32(tx0 + kTW - 1 + rb.z);
                lo = max(lo, 0); hi = min(hi, a.strips_x - 1);
                if (lo <= hi) {
                    const int k = atomicAdd(&s.n_rows, 1);
                    s.rows[2 * k] = row; s.rows[2 * k + 1] = lo | ((hi - lo + 1) << 16);
                }
            }
        }
        __syncthreads();
        const int n_rows = s.n_rows;
        int e = 0;
        while (e < n_rows) {                                           // CTA-uniform
            // ---- level 2: the strips of as many candidate rows as fit the list ----
            int e_end = e, total = 0;
            while (e_end < n_rows) {
                const int cnt = s.rows[2 * e_end + 1] >> 16;
                if (e_end > e && total + cnt > kCandCap) break;
                total += cnt; ++e_end;
            }
            if (tid == 0) s.n_cand = 0;
            __syncthreads();
            for (int k = e; k < e_end; ++k) {
                const int row = s.rows[2 * k], pk = s.rows[2 * k + 1], lo = pk & 0xffff, cnt = pk >> 16;
                for (int i = tid; i < cnt; i += kOT) {
                    const short4 b = strip_box_g[(size_t)row * a.strips_x + lo + i];
                    if (b.y >= tx0 - 1 && b.x <= tx0 + kTW - 1 && b.w >= ty0 - 1 && b.z <= ty0 + kTH - 1) {
                        const int q = atomicAdd(&s.n_cand, 1);
                        if (q < kCandCap) s.cand[q] = (row << 12) | (lo + i);
                    }
                }
            }
            __syncthreads();
            const int n_cand = min(s.n_cand, kCandCap);
            e = e_end;
            if (!win_clean) {                                          // after the last use of the self-derived boxes
                for (int i = tid; i < kWinCells * NV; i += kOT) (&s.win[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                win_clean = true;
                __syncthreads();
            }
            // ---- scan the candidates kOWarps strips at a time; accumulate whenever a full batch of hits is queued ----
            for (int base = 0; base < n_cand; base += kOWarps) {
                if (base + warp < n_cand) scan_strip<TF, NV>(a, s, f, s.cand[base + warp], tx0, ty0, lane);
                __syncthreads();
                // at most kBatch - 1 records may be left before the next scan step (which appends up to 1024): ring of 2048
                while (s.q_tail - head >= kBatch) {                    // CTA-uniform: q_tail only moves between barriers
                    accumulate_batch<T, MODE, CA, NV>(a, s, f, head, kBatch, tx0, ty0, tid, false);
                    head += kBatch;
                    __syncthreads();
                }
            }
        }
    }
    if (!win_clean) {     // no candidate at all: the tile is a hole
        for (int i = tid; i < kWinCells * NV; i += kOT) (&s.win[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // ---- drain the ring (re-queued losers keep arriving until every record has won a claim) ----
    for (;;) {
        const int avail = s.q_tail - head;                            // CTA-uniform: read between barriers
        if (avail <= 0) break;
        const int nb = min(avail, kBatch);
        __syncthreads();                                               // everybody has read q_tail before it moves again
        accumulate_batch<T, MODE, CA, NV>(a, s, f, head, nb, tx0, ty0, tid, avail <= kBatch);
        head += nb;
        __syncthreads();
    }

    // ---- epilogue: normalise out of shared memory, write once ----
    const int cx = tid & (kTW - 1);
    const int x = tx0 + cx;
    if (x >= W) return;
    T* outp = (T*)a.out + (long long)f * C * a.HW;
#pragma unroll 2
    for (int cy = tid / kTW; cy < kTH; cy += kOT / kTW) {
        const int y = ty0 + cy;
        if (y >= H) break;
        const unsigned r = (unsigned)y * (unsigned)W + (unsigned)x;
        float sv[4 * NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 cell = s.win[q][(cy + 1) * kPW + cx + 1];
            sv[4 * q] = cell.x; sv[4 * q + 1] = cell.y; sv[4 * q + 2] = cell.z; sv[4 * q + 3] = cell.w;
        }
        if (a.epi == 1) {
            // occlusion test of compute_mask (control_utils.py:15-16): accumulators hold (x*e, y*e, e)
            const T* fp = (const T*)a.epi_flow.p + (long long)f * a.epi_flow.sN + (long long)y * a.epi_flow.sH + (long long)x * a.epi_flow.sW;
            const float n = add_rn(sv[2], 0.0000001f);
            const float qx = add_rn(ld<float>(fp), sv[0] / n), qy = add_rn(ld<float>(fp + a.epi_flow.sC), sv[1] / n);
            st<T, float>((T*)a.mask_out + (long long)f * a.HW + r, sqrtf(add_rn(mul_rn(qx, qx), mul_rn(qy, qy))) > 0.3f ? 1.f : 0.f);
            continue;
        }
        float scale = 1.f;
        bool scaled = false;
        if (MODE != DCB_MODE_SUM) {
            float d = sv[C];
            // softsplat.py:256-266
            if (a.eps == DCB_EPS_ADD) d = add_rn(d, 0.0000001f);
            else if (a.eps == DCB_EPS_ZERO) d = (d == 0.f) ? 1.f : d;
            else d = (d < 0.0000001f) ? 0.0000001f : d;
            scale = __frcp_rn(d);            // one correctly rounded reciprocal + C multiplies (<= 1 ulp from softsplat.py:270)
            scaled = true;
            if (a.norm) __stcs((float*)a.norm + (long long)f * a.HW + r, d);
        }
        if (a.mask.p) {
            const T* mp = (const T*)a.mask.p + (long long)f * a.mask.sN + (long long)y * a.mask.sH + (long long)x * a.mask.sW;
            scale = mul_rn(scale, sub_rn(1.f, ld<float>(mp)));       // control_utils.py:69-70
            scaled = true;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) st_stream(outp + (size_t)c * a.HW + r, scaled ? mul_rn(sv[c], scale) : sv[c]);
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static long long g_owner_group_bytes = 48ll << 20;      // flow bytes per frame group (pre-pass + owner launch pair): stays in L2 between the two

void owner_set_group_bytes(long long b) { g_owner_group_bytes = b > 0 ? b : (48ll << 20); }

static void owner_geometry(long long H, long long W, long long& strips_x, long long& strip_rows) {
    strips_x = (W + 31) / 32; strip_rows = (H + kSR - 1) / kSR;
}

bool owner_supported(long long C, int mode, long long H, long long W) {
    const long long ca = C + (mode == DCB_MODE_SUM ? 0 : 1);
    long long sx, sr; owner_geometry(H, W, sx, sr);
    return ca >= 1 && ca <= 4 && H < 32000 && W < 32000 && sx <= kCandCap && sx < 4096 && sr < (1 << 19);
}

long long owner_workspace(long long N, long long H, long long W) {
    long long sx, sr; owner_geometry(H, W, sx, sr);
    if (sx * sr <= kSelfStrips && sr <= kRowCap) return 0;
    return align_up(N * sr * (sx + 1) * 8, 256);
}

template <class T, class TF, int MODE, int CA> static int launch_owner(OwnerArgs& a, cudaStream_t st) {
    constexpr int NV = (CA + 3) / 4;
    const size_t smem = sizeof(OwnerSmem<NV>);
    auto kern = k_splat_owner<T, TF, MODE, CA>;
    static bool attr_done = false;                    // per instantiation
    if (!attr_done) {
        DCB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const long long flow_bytes = (long long)a.HW * 2 * (long long)sizeof(TF);
    long long G = a.self_box ? a.N : g_owner_group_bytes / (flow_bytes > 0 ? flow_bytes : 1);
    if (G < 1) G = 1;
    const long long tiles = (long long)a.tiles_x * a.tiles_y;
    const long long max_frames = 0x7fffffffll / (tiles > 0 ? tiles : 1);     // gridDim.x limit
    if (G > max_frames) G = max_frames;
    if (G > 65535) G = 65535;                                                 // gridDim.y of the pre-pass
    for (long long f0 = 0; f0 < a.N; f0 += G) {
        const int frames = (int)(a.N - f0 < G ? a.N - f0 : G);
        a.frame0 = (int)f0;
        if (!a.self_box) {
            DCB_CHECK_CUDA(launch_pdl(k_strip_box<TF>, dim3((unsigned)a.strip_rows, (unsigned)frames), dim3(kPreThreads), 0, st, a));
            count_launch();
        }
        DCB_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)(tiles * frames)), dim3(kOT), smem, st, a));
        count_launch();
    }
    return DCB_OK;
}

template <class T, class TF, int MODE> static int owner_mode(OwnerArgs& a, cudaStream_t st) {
    const int ca = a.C + (MODE != DCB_MODE_SUM ? 1 : 0);
    switch (ca) {
        case 1: if (MODE == DCB_MODE_SUM) return launch_owner<T, TF, DCB_MODE_SUM, 1>(a, st); break;
        case 2: return launch_owner<T, TF, MODE, 2>(a, st);
        case 3: return launch_owner<T, TF, MODE, 3>(a, st);
        case 4: return launch_owner<T, TF, MODE, 4>(a, st);
    }
    return set_error(DCB_E_LIMIT, "splat_owner: %d accumulated channels", ca);
}

template <class T, class TF> static int owner_dtype(OwnerArgs& a, int mode, cudaStream_t st) {
    switch (mode) {
        case DCB_MODE_SUM: return owner_mode<T, TF, DCB_MODE_SUM>(a, st);
        case DCB_MODE_AVG: return owner_mode<T, TF, DCB_MODE_AVG>(a, st);
        case DCB_MODE_LINEAR: return owner_mode<T, TF, DCB_MODE_LINEAR>(a, st);
        default: return owner_mode<T, TF, DCB_MODE_SOFT>(a, st);
    }
}

// Preconditions (checked by the caller): owner_supported(), dtype F32/BF16, workspace >= owner_workspace().
int splat_owner_impl(const DcbTensor* in, const DcbTensor* flow, const DcbTensor* metric, const DcbTensor* out,
                     const DcbTensor* norm, const DcbTensor* mask, void* ws, int mode, int eps, cudaStream_t st,
                     bool ones_metric, const DcbTensor* mask_out) {
    OwnerArgs a;
    a.ones = ones_metric ? 1 : 0;
    a.epi = mask_out ? 1 : 0;
    a.epi_flow = make_view(mask_out ? flow : nullptr);
    a.mask_out = mask_out ? mask_out->ptr : nullptr;
    a.in = make_view(in); a.flow = make_view(flow); a.metric = make_view(metric); a.mask = make_view(mask);
    a.N = (int)in->size[0]; a.C = (int)in->size[1]; a.H = (int)in->size[2]; a.W = (int)in->size[3];
    a.HW = (unsigned)(in->size[2] * in->size[3]);
    a.eps = eps;
    a.tiles_x = (a.W + kTW - 1) / kTW; a.tiles_y = (a.H + kTH - 1) / kTH;
    long long sx, sr; owner_geometry(a.H, a.W, sx, sr);
    a.strips_x = (int)sx; a.strip_rows = (int)sr;
    a.self_box = owner_workspace(a.N, a.H, a.W) == 0 ? 1 : 0;
    a.strip_box = (short4*)ws;
    a.row_box = a.strip_box + (size_t)a.N * sr * sx;
    a.out = out ? out->ptr : nullptr;
    a.norm = norm ? norm->ptr : nullptr;
    a.frame0 = 0;
    const bool ff = flow->dtype == DCB_F32;
    if (in->dtype == DCB_F32) return owner_dtype<float, float>(a, mode, st);
    if (in->dtype == DCB_BF16)
        return ff ? owner_dtype<__nv_bfloat16, float>(a, mode, st) : owner_dtype<__nv_bfloat16, __nv_bfloat16>(a, mode, st);
    return set_error(DCB_E_DTYPE, "splat_owner: unsupported dtype %d", in->dtype);
}

}  // namespace dcb
