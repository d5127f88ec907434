// Tiled DCNv3 backward (sm_100a): two kernels, both deterministic.
//
//  bwd_gather_kernel  -- grad_offset and grad_mask.  Same structure as the tiled forward: one CTA per
//      output tile x group chunk, the chunk's input box staged by one TMA load, every lane owns one
//      (pixel, group), gathers the 4 corner slabs of every tap (conflict-free rotated LDS.128) and forms
//      the four dot products <grad_out, I_k> with packed FFMA2; from them the two gradients
//      (SURVEY.md App. A.2).  No cross-lane reduction, no atomics.
//
//  bwd_scatter_kernel -- grad_x, OWNER COMPUTES.  One CTA = one 16x16 tile J of INPUT pixels x one group
//      chunk.  Shared memory holds an int32 accumulator for exactly the cells of J.  The CTA walks every
//      output pixel whose taps can plausibly reach J (home pixels + a margin) and adds
//          q = floor(G[c] * Wk / 2^32) + b        G = grad_out, Wk = mask * bilinear weight (fixed point),
//                                                 b = parity bit that makes the truncation unbiased
//      with shared-memory integer atomics (ATOMS.ADD; measured 31.6 updates/clk/SM when conflict free,
//      tools/microbench2.cu; rows and cells of the accumulator have odd pitches so that neighbouring
//      pixels fall into different banks).  Integer addition is associative: the result is bitwise
//      independent of warp scheduling.  Every cell of J is owned by exactly one CTA: grad_x is written
//      once, there is no cross-CTA reduction and no float atomic anywhere.
//
//  Overflow / far taps -- every landing also adds ceil(|Wk|*1024) to a per-(cell, group) counter; while
//      that sum stays <= 8*1024 the int32 accumulator provably cannot wrap.  Cells that exceed it
//      ("hot", only with adversarial inputs) are zeroed at flush and recomputed exactly by
//      redo_hot_kernel with 64-bit integer global atomics; taps of home pixels that land in a tile
//      whose owner does not visit the pixel (|offset| beyond the margin) take the same 64-bit side
//      path directly.  The side buffer uses the SAME integers q and is merged by merge_far_kernel; which
//      path a contribution takes is a deterministic function of the inputs.
//
// The common scale comes from max|grad_out| (amax_go_kernel): G = round(go * 2^eg), 2^29 <= max|G| < 2^30.
#include "dcnv3_kernels.h"
#include "dcnv3_tiled.cuh"

namespace dcnv3 {

constexpr int kSG = 2;             // groups per scatter CTA (lane = pixel * kSG + group, 16 pixels per warp)
constexpr int kSCell = kSG * kGC;   // accumulator ints per cell: exactly 32 banks wide, so the bank of an
                                   // update depends only on (group, channel) and never on the cell
constexpr int kBudget = 8 * 1024;  // sum of ceil(|Wk| * 1024) allowed in the int32 accumulator
constexpr int kWShift = 29;        // Wk fixed point: round(Wk * 2^29), |Wk| < 4

struct BwdGeom {
    int tj, tj_log2;       // input tile edge in cells (16 or 32) and its log2
    int pitch;             // accumulator row pitch in cells
    int tiles_x, tiles_y;  // tiles of tj x tj un-padded input pixels
    int chunks;            // G / kSG
    int margin;            // cells by which the scatter kernel looks beyond J for source pixels
    int acc_ints, wsum_ints;
};

struct FarWs {
    WsHeader* hd;
    unsigned char* dirty;         // [N][H][W][G]  (pixel, group) has side-buffer contributions
    int* redo;                    // [N][chunks][tiles_y][tiles_x]  scatter CTA found hot cells
    unsigned long long* acc64;    // [N][H][W][C] fixed point, zero outside a call
};

// un-padded nominal input column of output row h (may be -1 at the border), integer arithmetic so
// that every CTA agrees exactly:  floor((2h+3) * (W_in-2) / (2 H_in)) - 1
// (32-bit: tiled_applicable() limits H, W to 16384, so (2h+3)*(W_in-2) < 2^31)
__device__ __forceinline__ int nominal_ux(const KParams& q, int h) {
    return (int)((unsigned)((2 * h + 3) * (q.win - 2)) / (unsigned)(2 * q.hin)) - q.pw;
}
__device__ __forceinline__ int nominal_uy(const KParams& q, int w) {
    return (int)((unsigned)((2 * w + 3) * (q.hin - 2)) / (unsigned)(2 * q.win)) - q.ph;
}
// smallest index i in [0, n] with nominal(i) >= a   (nominal is non-decreasing)
template <typename F>
__device__ __forceinline__ int first_ge(F nominal, int n, int a, int num_scale, int den) {
    // floor((2i+3)*S / (2D)) - 1 >= a   <=>   i >= ((a+1)*2D - 3S) / (2S); closed form, then fix up
    const int num = 2 * (a + 1) * den - 3 * num_scale;
    int r = num <= 0 ? 0 : (int)((unsigned)(num + 2 * num_scale - 1) / (unsigned)(2 * num_scale));
    if (r > n) r = n;
    while (r > 0 && nominal(r - 1) >= a) --r;
    while (r < n && nominal(r) < a) ++r;
    return r;
}

struct Range { int lo, hi; };

// source range along one axis for tile index j: every i with
//   nominal(i) in [j*tj - margin, j*tj + tj + margin)   or   home(i) == j
template <typename F>
__device__ __forceinline__ Range window_range(F nominal, int n, int j, int ntiles, int tj, int margin, int num,
                                              int den) {
    Range r;
    r.lo = (j == 0) ? 0 : first_ge(nominal, n, j * tj - margin, num, den);
    r.hi = (j == ntiles - 1) ? n : first_ge(nominal, n, j * tj + tj + margin, num, den);
    return r;
}
template <typename F>
__device__ __forceinline__ Range home_range(F nominal, int n, int j, int ntiles, int tj, int num, int den) {
    Range r;
    r.lo = (j == 0) ? 0 : first_ge(nominal, n, j * tj, num, den);
    r.hi = (j == ntiles - 1) ? n : first_ge(nominal, n, j * tj + tj, num, den);
    return r;
}
// tile indices j whose owner visits a source index with nominal value u: [lo, hi]  (the complement is
// "far").  Mirrors window_range: u in [j*tj - margin, j*tj + tj + margin), the first tile also takes
// u < 0 and the last one u >= extent.
__device__ __forceinline__ Range covering_tiles(int u, int ntiles, int tj, int tj_log2, int margin) {
    Range r;
    int lo = u - tj - margin + 1;  // need j*tj >= lo  (ceil division)
    lo = lo <= 0 ? 0 : (lo + tj - 1) >> tj_log2;
    int hi = u + margin;           // need j*tj <= hi  (floor division)
    hi = hi < 0 ? 0 : hi >> tj_log2;
    r.lo = min(lo, ntiles - 1);
    r.hi = min(hi, ntiles - 1);
    return r;
}

// called by threads 0..3 of the CTA, one range each
__device__ __forceinline__ void tile_ranges(const KParams& q, const BwdGeom& bg, int jx, int jy, Range& home_h,
                                            Range& home_w, Range& win_h, Range& win_w) {
    auto nx = [&](int h) { return nominal_ux(q, h); };
    auto ny = [&](int w) { return nominal_uy(q, w); };
    // output rows h walk along input x, output columns w along input y
    if (threadIdx.x == 0) home_h = home_range(nx, q.ho, jx, bg.tiles_x, bg.tj, q.win - 2, q.hin);
    if (threadIdx.x == 1) home_w = home_range(ny, q.wo, jy, bg.tiles_y, bg.tj, q.hin - 2, q.win);
    if (threadIdx.x == 2) win_h = window_range(nx, q.ho, jx, bg.tiles_x, bg.tj, bg.margin, q.win - 2, q.hin);
    if (threadIdx.x == 3) win_w = window_range(ny, q.wo, jy, bg.tiles_y, bg.tj, bg.margin, q.hin - 2, q.win);
}

template <typename T>
__device__ __forceinline__ const T* global_slab_b(const T* x, const KParams& q, int n, int yp, int xp, int g) {
    const int y = yp - q.ph, xx = xp - q.pw;
    if (y < 0 || y >= q.h || xx < 0 || xx >= q.w) return nullptr;
    return x + ((((size_t)n * q.h + y) * q.w + xx) * q.G + g) * kGC;
}

// per-warp staging bytes of the gather kernel: results + (bf16 only) a separate fp32 park
template <typename T>
constexpr int kGatherStageBytes = RowStage<T>::BYTES + (sizeof(T) == 4 ? 0 : 32 * kTaps * 4);

// grad_out of one (pixel, group) in fixed point: G[c] = round(go[c] * 2^eg)
// ... XOR-permuted by the lane's pixel index: G[c] holds channel c ^ rot.  At step c every lane updates
// channel c ^ rot, so the 32 lanes of one ATOMS (16 pixels x 2 groups) hit 32 different banks
// whatever cells they land on.
template <typename T>
__device__ __forceinline__ void load_fixed_point_go(const T* go, float sg, int rot, int (&G)[16]) {
    using C = Chunk<T>;  // only for the 16-byte piece geometry of T
    f2 gf[8];
#pragma unroll
    for (int pc = 0; pc < C::NPIECE; ++pc) load_piece<T>(go + pc * C::CH_PER_PIECE, gf + pc * C::PAIRS);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        G[2 * c] = __float2int_rn(lo_of(gf[c]) * sg);
        G[2 * c + 1] = __float2int_rn(hi_of(gf[c]) * sg);
    }
#pragma unroll
    for (int b = 1; b < 16; b <<= 1) {  // v[i] <- v[i ^ rot]: four conditional butterfly stages
        const bool sw = rot & b;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((i & b) == 0) {
                const int lo = G[i], hi = G[i | b];
                G[i] = sw ? hi : lo;
                G[i | b] = sw ? lo : hi;
            }
        }
    }
}

// shared-memory integer add without return value, 32-bit shared address
__device__ __forceinline__ void red_shared_add(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// =====================================================================================================
// grad_offset / grad_mask
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256, 2)
bwd_gather_kernel(const __grid_constant__ CUtensorMap xmap, const T* __restrict__ x,
                  const T* __restrict__ offset, const T* __restrict__ mask, const T* __restrict__ grad_out,
                  T* __restrict__ grad_offset, T* __restrict__ grad_mask, WsHeader* __restrict__ hd, const KParams q,
                  const TileGeom tg) {
    using C = Chunk<T>;
    using RS = RowStage<T>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    int b = blockIdx.x;
    const int tx = b % tg.tiles_w; b /= tg.tiles_w;
    const int ty = b % tg.tiles_h; b /= tg.tiles_h;
    const int chunk = b % tg.chunks;
    const int n = b / tg.chunks;
    const int h0 = ty * tg.th, w0 = tx * tg.tw;
    const int th = min(tg.th, q.ho - h0), tw = min(tg.tw, q.wo - w0);
    const int cx0 = max(0, min((int)floorf(nominal_x(q, h0)) - tg.halo_x, q.win - tg.bw));
    const int cy0 = max(0, min((int)floorf(nominal_y(q, w0)) - tg.halo_y, q.hin - tg.bh));

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(tg.bw * tg.bh * kCellBytes));
        tma_load_4d(smem, &xmap, &bar, chunk * C::GQ * kGC, cx0 - q.pw, cy0 - q.ph, n);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g_l = lane % C::GQ, px_l = lane / C::GQ;
    const int g = min(chunk * C::GQ + g_l, q.G - 1);  // phantom groups of a trailing chunk shadow the last one
    const int ng = min(C::GQ, q.G - chunk * C::GQ);   // real groups in this chunk
    const int rot = Slab<T>::rot_of(px_l);
    const unsigned char* sbase = smem + g_l * (kGC * (int)sizeof(T));
    // per-warp staging slot for the results (stored coalesced once per row segment) and, for the fused
    // soft-max path, an fp32 park for dL/dm_p (it aliases the mask slot when T is fp32)
    unsigned char* st = smem + (size_t)tg.bw * tg.bh * kCellBytes + warp * kGatherStageBytes<T>;
    float* park = sizeof(T) == 4 ? reinterpret_cast<float*>(st + RS::OFF_BYTES)
                                 : reinterpret_cast<float*>(st + RS::BYTES);
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const int colblocks = (tw + C::PXW - 1) / C::PXW;
    bool waited = false;
    float amax = 0.f;  // max |grad_out| seen by this thread: the fixed-point scale of the scatter kernel

    // one warp iteration = PXW consecutive pixels of one output row
    for (int it = warp; it < th * colblocks; it += nwarps) {
        const int h = h0 + it / colblocks, wb = w0 + (it % colblocks) * C::PXW;
        const int npx = min(C::PXW, w0 + tw - wb);
        const int w = wb + min(px_l, npx - 1);  // idle lanes shadow the last pixel (their slots are never stored)
        const size_t pix0 = ((size_t)n * q.ho + h) * q.wo + wb;
        const size_t pg = (pix0 + min(px_l, npx - 1)) * q.G + g;
        const T* offp = offset + pg * 18;
        const T* mskp = mask + pg * 9;
        f2 go[8];
#pragma unroll
        for (int pc = 0; pc < C::NPIECE; ++pc)
            load_piece<T>(grad_out + pg * kGC + Slab<T>::chan_of(pc, rot), go + pc * C::PAIRS);
#pragma unroll
        for (int c = 0; c < 8; ++c) amax = fmaxf(amax, fmaxf(fabsf(lo_of(go[c])), fabsf(hi_of(go[c]))));
        float mx = 0.f, inv_sum = 1.f;
        if (logits) softmax_stats9<T>(mskp, mx, inv_sum);
        float ref0, ref1;
        ref_point(q, h, w, ref0, ref1);
        float ox, oy, ml, ox2, oy2, ml2;
        load_tap_inputs<T>(offp, mskp, 0, ox, oy, ml);
        load_tap_inputs<T>(offp, mskp, 1, ox2, oy2, ml2);
        if (!waited) {
            mbar_wait(&bar, 0);
            waited = true;
        }
        float gm_dot_m = 0.f;
#pragma unroll 1
        for (int p = 0; p < kTaps; ++p) {
            const float cx = ox, cy = oy, cm = ml;
            ox = ox2; oy = oy2; ml = ml2;
            if (p + 2 < kTaps) load_tap_inputs<T>(offp, mskp, p + 2, ox2, oy2, ml2);  // two taps ahead
            const Tap t = make_tap(q, ref0, ref1, p, cx, cy);
            const int bx = t.x0 - cx0, by = t.y0 - cy0;
            const bool inbox = bx >= 0 && bx + 1 < tg.bw && by >= 0 && by + 1 < tg.bh;
            const float mm = logits ? expf(cm - mx) * inv_sum : cm;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;  // <go, I_k>: a=(y0,x0) b=(y1,x0) c=(y0,x1) d=(y1,x1)
            if (__builtin_expect(t.alive && !inbox, 0)) {
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const T* src = global_slab_b(x, q, n, t.y0 + (k & 1), t.x0 + (k >> 1), g);
                    if (src == nullptr) continue;
                    f2 dk2 = 0ull;
#pragma unroll
                    for (int pc = 0; pc < C::NPIECE; ++pc) {
                        f2 v[C::PAIRS];
                        load_piece<T>(src + Slab<T>::chan_of(pc, rot), v);
#pragma unroll
                        for (int j = 0; j < C::PAIRS; ++j) ffma2v(dk2, v[j], go[pc * C::PAIRS + j]);
                    }
                    const float dk = lo_of(dk2) + hi_of(dk2);
                    if (k == 0) d0 = dk; else if (k == 1) d1 = dk; else if (k == 2) d2 = dk; else d3 = dk;
                }
            } else {
                const unsigned char* a = sbase + (size_t)(t.alive ? by * tg.bw + bx : 0) * kCellBytes;
                f2 va[8], vb[8], e0 = 0ull, e1 = 0ull, e2 = 0ull, e3 = 0ull;
                Slab<T>::load(a, rot, va);
                Slab<T>::load(a + (size_t)tg.bw * kCellBytes, rot, vb);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2v(e0, va[c], go[c]);
                Slab<T>::load(a + kCellBytes, rot, va);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2v(e1, vb[c], go[c]);
                Slab<T>::load(a + (size_t)(tg.bw + 1) * kCellBytes, rot, vb);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2v(e2, va[c], go[c]);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2v(e3, vb[c], go[c]);
                d0 = lo_of(e0) + hi_of(e0);
                d1 = lo_of(e1) + hi_of(e1);
                d2 = lo_of(e2) + hi_of(e2);
                d3 = lo_of(e3) + hi_of(e3);
            }
            // dead taps have all four deltas zero => zero gradients
            const float g_m = t.dx1 * t.dy1 * d0 + t.dx1 * t.dy0 * d1 + t.dx0 * t.dy1 * d2 + t.dx0 * t.dy0 * d3;
            const float gxq = mm * (t.dy1 * (d2 - d0) + t.dy0 * (d3 - d1));
            const float gyq = mm * (t.dx1 * (d1 - d0) + t.dx0 * (d3 - d2));
            gm_dot_m += g_m * mm;
            // results go to the lane's staging slot (lane stride 72 / 36 bytes: conflict free)
            if (sizeof(T) == 4) {
                *reinterpret_cast<float2*>(st + lane * RS::LANE_OFF + p * 8) = make_float2(gxq * q.fx, gyq * q.fy);
            } else {
                *reinterpret_cast<unsigned*>(st + lane * RS::LANE_OFF + p * 4) = pack_bf16x2(gxq * q.fx, gyq * q.fy);
            }
            if (logits || sizeof(T) == 4) park[lane * kTaps + p] = g_m;
            else *reinterpret_cast<__nv_bfloat16*>(st + RS::OFF_BYTES + lane * RS::LANE_MSK + p * 2) = __float2bfloat16_rn(g_m);
        }
        if (logits) {
            // softmax Jacobian needs sum_p m_p*dL/dm_p: second sweep over this lane's own 9 values
#pragma unroll 1
            for (int p = 0; p < kTaps; ++p) {
                const float mm = expf(Elem<T>::ld(mskp + p) - mx) * inv_sum;
                const float v = mm * (park[lane * kTaps + p] - gm_dot_m);
                if (sizeof(T) == 4) park[lane * kTaps + p] = v;
                else *reinterpret_cast<__nv_bfloat16*>(st + RS::OFF_BYTES + lane * RS::LANE_MSK + p * 2) = __float2bfloat16_rn(v);
            }
        }
        RS::store_off_msk(st, grad_offset + (pix0 * q.G + chunk * C::GQ) * 18,
                          grad_mask + (pix0 * q.G + chunk * C::GQ) * 9, q.G, npx, ng, lane);
    }
    if (!waited) mbar_wait(&bar, 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) atomicMax(&hd->amax_go_bits, __float_as_uint(amax));
}

// =====================================================================================================
// grad_x
// =====================================================================================================
// home pixels first, then the four margin bands of the window
__device__ __forceinline__ void rect_of(int rect, Range wh, Range ww, Range hh, Range hw, Range& rh, Range& rw) {
    if (rect == 0) { rh = hh; rw = hw; }                                      // home
    else if (rect == 1) { rh.lo = wh.lo; rh.hi = hh.lo; rw = ww; }            // band below the home rows
    else if (rect == 2) { rh.lo = hh.hi; rh.hi = wh.hi; rw = ww; }            // band above
    else if (rect == 3) { rh = hh; rw.lo = ww.lo; rw.hi = hw.lo; }            // left band
    else { rh = hh; rw.lo = hw.hi; rw.hi = ww.hi; }                           // right band
}

// The scatter walk over the source pixels of tile (jx, jy).
//   MODE 0 (scatter kernel): landings inside J -> int32 shared atomics + weight counters; far landings of
//                            home pixels -> 64-bit side buffer
//   MODE 1 (redo, pass 1)  : weight counters only
//   MODE 2 (redo, pass 2)  : landings on hot cells of J -> 64-bit side buffer
// Source pixels are walked rectangle by rectangle -- the home pixels first, then the four margin bands
// -- so that a warp holds either home pixels (every tap lands) or margin pixels (almost none does).
template <typename T, int MODE>
__device__ __forceinline__ void scatter_pass(int* acc, int* wsum, const T* __restrict__ offset,
                                             const T* __restrict__ mask, const T* __restrict__ grad_out,
                                             const FarWs& ws, const KParams& q, const BwdGeom& bg, int n, int chunk,
                                             int jx, int jy, Range wh, Range ww, Range hh, Range hw, int eg,
                                             int* next_item) {
    constexpr int PXW = 32 / kSG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g_l = lane % kSG, px_l = lane / kSG;
    const int g = chunk * kSG + g_l;
    if (g >= q.G) return;  // phantom group of a trailing chunk (no warp-level synchronisation in this walk)
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const int ux0 = jx << bg.tj_log2, uy0 = jy << bg.tj_log2;
    const int tjw = min(bg.tj, q.w - ux0), tjh = min(bg.tj, q.h - uy0);
    const float sg = ldexpf(1.0f, eg);  // G = round(go * 2^eg), |G| < 2^30
    const size_t img_pixels = (size_t)q.h * q.w;
    const uint32_t acc_s = MODE == 0 ? smem_u32(acc) : 0u;
    int deal = 0;  // how many blocks have been dealt so far (mod nwarps)
    (void)next_item;
#pragma unroll 1
    for (int rect = 0; rect < 5; ++rect) {
        Range rh, rw;
        rect_of(rect, wh, ww, hh, hw, rh, rw);
        const int nw = rw.hi - rw.lo, npix = (rh.hi - rh.lo) * nw;
        const bool is_home = MODE == 0 && rect == 0;
        // blocks of PXW pixels are dealt round-robin, continuing across rectangles so that the warps
        // that got one block fewer in a rectangle get the first ones of the next
        for (int p0 = ((warp + nwarps - deal) % nwarps) * PXW; p0 < npix; p0 += nwarps * PXW) {
            const int pix = p0 + px_l;
            if (pix >= npix) continue;
            const int h = rh.lo + pix / nw, w = rw.lo + pix % nw;
            const size_t pg = (((size_t)n * q.ho + h) * q.wo + w) * q.G + g;
            const T* offp = offset + pg * 18;
            const T* mskp = mask + pg * 9;
            float mx = 0.f, inv_sum = 1.f;
            if (logits) softmax_stats9<T>(mskp, mx, inv_sum);
            float ref0, ref1;
            ref_point(q, h, w, ref0, ref1);
            // un-padded cell range [lo, hi] covered by the tiles whose owners visit this pixel themselves:
            // a landing of a home pixel outside it is "far" and takes the side path
            int cvx_lo = 0, cvx_hi = 0, cvy_lo = 0, cvy_hi = 0;
            if (is_home) {
                const Range cx = covering_tiles(nominal_ux(q, h), bg.tiles_x, bg.tj, bg.tj_log2, bg.margin);
                const Range cy = covering_tiles(nominal_uy(q, w), bg.tiles_y, bg.tj, bg.tj_log2, bg.margin);
                cvx_lo = cx.lo << bg.tj_log2; cvx_hi = ((cx.hi + 1) << bg.tj_log2) - 1;
                cvy_lo = cy.lo << bg.tj_log2; cvy_hi = ((cy.hi + 1) << bg.tj_log2) - 1;
            }
            int G[16];
            bool have_g = false;
            if (is_home) {  // every tap of a home pixel lands: convert grad_out once, with the whole warp converged
                load_fixed_point_go<T>(grad_out + pg * kGC, sg, px_l, G);
                have_g = true;
            }
            float ox, oy, ml;
            load_tap_inputs<T>(offp, mskp, 0, ox, oy, ml);
#pragma unroll 1
            for (int p = 0; p < kTaps; ++p) {
                const float cxo = ox, cyo = oy, cm = ml;
                if (p + 1 < kTaps) load_tap_inputs<T>(offp, mskp, p + 1, ox, oy, ml);
                // one axis at a time: pixels of the margin are mostly rejected after the first coordinate
                const Axis axx = axis_x(q, ref0, p, cxo);
                if (!axx.alive) continue;
                const int lx = axx.i0 - q.pw - ux0;  // corner (y0,x0) relative to J
                const bool col0 = lx >= 0 && lx < tjw, col1 = lx + 1 >= 0 && lx + 1 < tjw;
                if (!is_home && !(col0 || col1)) continue;
                const Axis axy = axis_y(q, ref1, p, cyo);
                if (!axy.alive) continue;
                const int ly = axy.i0 - q.ph - uy0;
                const bool row0 = ly >= 0 && ly < tjh, row1 = ly + 1 >= 0 && ly + 1 < tjh;
                if (!is_home && !(row0 || row1)) continue;
                const float mm = logits ? expf(cm - mx) * inv_sum : cm;
                const int par = (p ^ h ^ w) & 1;
                // can a corner of this tap be far?  (only then is the out-of-tile part looked at)
                const int ax0 = lx + ux0, ay0 = ly + uy0;
                const bool far_possible = is_home && (ax0 < cvx_lo || ax0 + 1 > cvx_hi || ay0 < cvy_lo || ay0 + 1 > cvy_hi);
                unsigned side_mask = 0;  // corners that need the 64-bit side path
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // a b c d = (y0,x0) (y1,x0) (y0,x1) (y1,x1)
                    const bool in_tile = ((k >> 1) ? col1 : col0) && ((k & 1) ? row1 : row0);
                    const float wf = ((k >> 1) ? axx.d0 : axx.d1) * ((k & 1) ? axy.d0 : axy.d1) * mm;
                    if (wf == 0.f) continue;
                    if (!in_tile) {
                        if (far_possible) side_mask |= 1u << k;  // exact far test below
                        continue;
                    }
                    const int cell = (ly + (k & 1)) * bg.pitch + lx + (k >> 1);
                    const int cellg = cell * kSG + g_l;
                    if (MODE == 2) {
                        if (wsum[cellg] > kBudget) side_mask |= 1u << k;
                        continue;
                    }
                    // a raw mask beyond the fixed-point range (|Wk| >= 3.9) makes the cell hot by itself
                    const int wb = __float2int_ru(fabsf(wf) * 1024.f);
                    atomicAdd(&wsum[cellg], wb < 3994 ? wb : kBudget + 1);
                    if (MODE == 1) continue;
                    if (!have_g) {
                        load_fixed_point_go<T>(grad_out + pg * kGC, sg, px_l, G);
                        have_g = true;
                    }
                    const int wq = __float2int_rn(fminf(fmaxf(wf, -3.9f), 3.9f) * (float)(1 << kWShift));
                    const int bb = par ^ (k & 1) ^ (k >> 1);
                    // slab base is 64-byte aligned: base + ((c ^ rot) * 4) == (base ^ rot*4) ^ c*4
                    const uint32_t dx = (acc_s + (uint32_t)(cell * kSCell + g_l * kGC) * 4u) ^ ((uint32_t)px_l << 2);
#pragma unroll
                    for (int c = 0; c < 16; ++c) red_shared_add(dx ^ (c << 2), __mulhi(G[c], wq) + bb);
                }
                if (__builtin_expect(side_mask != 0, 0)) {
#pragma unroll 1
                    for (int k = 0; k < 4; ++k) {
                        if (!((side_mask >> k) & 1)) continue;
                        const int ax = lx + (k >> 1) + ux0, ay = ly + (k & 1) + uy0;  // un-padded image coords
                        if (MODE == 0) {
                            if (ax < 0 || ax >= q.w || ay < 0 || ay >= q.h) continue;  // zero ring: gradient dropped
                            // the owner of that tile visits this pixel itself unless the tap is far
                            if (ax >= cvx_lo && ax <= cvx_hi && ay >= cvy_lo && ay <= cvy_hi) continue;
                        }
                        if (!have_g) {
                            load_fixed_point_go<T>(grad_out + pg * kGC, sg, px_l, G);
                            have_g = true;
                        }
                        const float wf = ((k >> 1) ? axx.d0 : axx.d1) * ((k & 1) ? axy.d0 : axy.d1) * mm;
                        // same integer q as the shared-memory path; |Wk| >= 4 (raw masks only) is pre-shifted
                        int sh = 0;
                        if (!(fabsf(wf) < 3.9f))
                            sh = min(max((int)((__float_as_uint(wf) >> 23) & 0xffu) - 128, 0), 30);
                        const int wq = __float2int_rn(ldexpf(wf, kWShift - sh));
                        const int bb = par ^ (k & 1) ^ (k >> 1);
                        unsigned long long* dst =
                            ws.acc64 + (((size_t)n * img_pixels + (size_t)ay * q.w + ax) * q.G + g) * kGC;
#pragma unroll
                        for (int c = 0; c < 16; ++c)  // G[c] holds channel c ^ px_l
                            atomicAdd(dst + (c ^ px_l), (unsigned long long)(((long long)__mulhi(G[c], wq) << sh) + bb));
                        ws.dirty[((size_t)n * img_pixels + (size_t)ay * q.w + ax) * q.G + g] = 1;
                    }
                }
            }
        }
        deal = (deal + (npix + PXW - 1) / PXW) % nwarps;
    }
}

template <typename T>
__global__ void __launch_bounds__(640, 1)
bwd_scatter_kernel(const T* __restrict__ offset, const T* __restrict__ mask, const T* __restrict__ grad_out,
                   T* __restrict__ grad_x, const FarWs ws, const KParams q, const BwdGeom bg) {
    extern __shared__ __align__(16) int acc[];  // [acc_ints] accumulator + [wsum_ints] weight counters
    __shared__ Range s_home_h, s_home_w, s_win_h, s_win_w;
    __shared__ int s_next;
    int* wsum = acc + bg.acc_ints;

    int b = blockIdx.x;
    const int jx = b % bg.tiles_x; b /= bg.tiles_x;
    const int jy = b % bg.tiles_y; b /= bg.tiles_y;
    const int chunk = b % bg.chunks;
    const int n = b / bg.chunks;
    const int ux0 = jx << bg.tj_log2, uy0 = jy << bg.tj_log2;
    const int tjw = min(bg.tj, q.w - ux0), tjh = min(bg.tj, q.h - uy0);

    if (threadIdx.x < 4) tile_ranges(q, bg, jx, jy, s_home_h, s_home_w, s_win_h, s_win_w);
    if (threadIdx.x == 0) s_next = 0;
    const int eg = 30 - fixed_exponent_raw(ws.hd);
    for (int i = threadIdx.x; i < bg.acc_ints + bg.wsum_ints; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    scatter_pass<T, 0>(acc, wsum, offset, mask, grad_out, ws, q, bg, n, chunk, jx, jy, s_win_h, s_win_w, s_home_h,
                       s_home_w, eg, &s_next);
    __syncthreads();

    // ---- flush: J is written exactly once ----
    const float inv_s = ldexpf(1.0f, -(eg + kWShift - 32));  // q = value * 2^(eg + kWShift - 32)
    constexpr int CH = kSG * kGC;    // channels of this chunk per cell
    constexpr int QPC = CH / 4;      // 4-channel pieces per cell
    const int ncell = tjw * tjh;
    bool any_hot = false;
    for (int i = threadIdx.x; i < ncell * QPC; i += blockDim.x) {
        const int cl = i / QPC, piece = i % QPC;
        const int cy = cl / tjw, cx = cl % tjw;
        const int cell = cy * bg.pitch + cx;
        const int gl = (piece * 4) / kGC;
        if (chunk * kSG + gl >= q.G) continue;  // phantom group
        // a hot (cell, group) may have wrapped: it is zeroed here and recomputed by redo_hot_kernel
        const bool hot = wsum[cell * kSG + gl] > kBudget;
        any_hot |= hot;
        const int* src = acc + cell * kSCell + piece * 4;
        const size_t gpix = (size_t)n * q.h * q.w + (size_t)(uy0 + cy) * q.w + (ux0 + cx);
        const size_t gidx = gpix * ((size_t)q.G * kGC) + (size_t)chunk * CH + piece * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = hot ? 0.f : (float)src[j] * inv_s;
        if (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(grad_x) + gidx) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            uint2 r;
            r.x = pack_bf16x2(v[0], v[1]);
            r.y = pack_bf16x2(v[2], v[3]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(grad_x) + gidx) = r;
        }
    }
    if (any_hot) ws.redo[blockIdx.x] = 1;
}

// Exact recomputation of the hot cells of one tile (see header): pass 1 rebuilds the weight counters,
// pass 2 adds the landings on hot cells to the 64-bit side buffer.  Exits at once unless the scatter
// kernel flagged the CTA -- which only adversarial inputs make it do.
template <typename T>
__global__ void __launch_bounds__(256)
redo_hot_kernel(const T* __restrict__ offset, const T* __restrict__ mask, const T* __restrict__ grad_out,
                const FarWs ws, const KParams q, const BwdGeom bg) {
    if (ws.redo[blockIdx.x] == 0) return;
    extern __shared__ __align__(16) int wsum[];  // [wsum_ints]
    __shared__ Range s_home_h, s_home_w, s_win_h, s_win_w;
    __shared__ int s_next;
    int b = blockIdx.x;
    const int jx = b % bg.tiles_x; b /= bg.tiles_x;
    const int jy = b % bg.tiles_y; b /= bg.tiles_y;
    const int chunk = b % bg.chunks;
    const int n = b / bg.chunks;
    if (threadIdx.x < 4) tile_ranges(q, bg, jx, jy, s_home_h, s_home_w, s_win_h, s_win_w);
    if (threadIdx.x == 0) s_next = 0;
    for (int i = threadIdx.x; i < bg.wsum_ints; i += blockDim.x) wsum[i] = 0;
    __syncthreads();
    const int eg = 30 - fixed_exponent_raw(ws.hd);
    scatter_pass<T, 1>(nullptr, wsum, offset, mask, grad_out, ws, q, bg, n, chunk, jx, jy, s_win_h, s_win_w, s_home_h,
                       s_home_w, eg, &s_next);
    __syncthreads();
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    scatter_pass<T, 2>(nullptr, wsum, offset, mask, grad_out, ws, q, bg, n, chunk, jx, jy, s_win_h, s_win_w, s_home_h,
                       s_home_w, eg, &s_next);
    __syncthreads();
    if (threadIdx.x == 0) ws.redo[blockIdx.x] = 0;
}

// grad_x += side buffer for the (pixel, group)s flagged in the dirty map; leaves the side buffer and
// the map zeroed for the next call.  Each warp scans 128 map bytes; flagged entries are then handled by
// half-warps, one lane per channel (the 16 channels of a group are contiguous in both buffers).
template <typename T>
__global__ void __launch_bounds__(256)
merge_far_kernel(T* __restrict__ grad_x, const FarWs ws, const KParams q, size_t count) {
    const int lane = threadIdx.x & 31;
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    unsigned flags = 0;
    if (i4 < count) flags = *reinterpret_cast<const unsigned*>(ws.dirty + i4);  // map is padded to 256 bytes
    unsigned any = __ballot_sync(0xffffffffu, flags != 0);
    if (any == 0) return;
    const double inv_s = ldexp(1.0, -(30 - fixed_exponent_raw(ws.hd) + kWShift - 32));
    const size_t warp_base = i4 - (size_t)lane * 4;
    while (any) {
        const int src = __ffs(any) - 1;
        any &= any - 1;
        const unsigned f = __shfl_sync(0xffffffffu, flags, src);
        // the up to four flagged entries of lane `src`: two per pass, one half-warp each
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int j = pass * 2 + (lane >> 4);
            const size_t e = warp_base + (size_t)src * 4 + j;
            if (((f >> (8 * j)) & 0xffu) != 0 && e < count) {
                const size_t idx = e * kGC + (lane & 15);
                const long long v = (long long)ws.acc64[idx];
                if (v != 0) {
                    // |v| can exceed 2^24: go through double so that the exact total is rounded once
                    Elem<T>::st(grad_x + idx, Elem<T>::ld_plain(grad_x + idx) + (float)((double)v * inv_s));
                    ws.acc64[idx] = 0ull;
                }
            }
        }
    }
    if (flags != 0) *reinterpret_cast<unsigned*>(ws.dirty + i4) = 0u;
}

// ---- host side -----------------------------------------------------------------------------------
static BwdGeom make_bwd_geom(const KParams& q) {
    BwdGeom bg;
    // 32x32 tiles keep the margin overhead low; small images use 16x16 so that the grid still fills the GPU
    bg.tj = (q.w >= 32 && q.h >= 32) ? 32 : 16;
    bg.tj_log2 = bg.tj == 32 ? 5 : 4;
    bg.pitch = bg.tj;
    bg.tiles_x = (q.w + bg.tj - 1) / bg.tj;
    bg.tiles_y = (q.h + bg.tj - 1) / bg.tj;
    bg.chunks = (q.G + kSG - 1) / kSG;
    // source pixels are searched up to ~3 offset units (+1 for the tap grid) beyond the tile
    const float r = fmaxf(q.wm2_f / q.win_f, q.hm2_f / q.hin_f) * q.scale;
    bg.margin = min((int)ceilf((1.0f + 3.0f) * r), 12);
    bg.acc_ints = bg.tj * bg.pitch * kSCell;
    bg.wsum_ints = bg.tj * bg.pitch * kSG;
    return bg;
}

static size_t flag_bytes(size_t count) { return (count * sizeof(int) + 255) / 256 * 256; }
static size_t dirty_bytes(const KParams& q) { return ((size_t)q.n * q.h * q.w * q.G + 255) / 256 * 256; }

size_t bwd_tiled_workspace_bytes(const KParams& q) {
    const size_t tiles = (size_t)q.n * ((q.w + 15) / 16) * ((q.h + 15) / 16);  // upper bound (16x16 tiles)
    const size_t chunks = (size_t)(q.G + 1) / 2;
    return sizeof(WsHeader) + dirty_bytes(q) + flag_bytes(tiles * chunks) +
           sizeof(long long) * (size_t)q.n * q.h * q.w * q.G * q.gc;
}

template <typename T>
static cudaError_t launch_bwd_tiled_t(const void* x, const void* offset, const void* mask, const void* grad_out,
                                      void* grad_x, void* grad_offset, void* grad_mask, void* wsp,
                                      const KParams& q, int dtype, bool ws_clean, cudaStream_t st) {
    const BwdGeom bg = make_bwd_geom(q);
    const size_t tiles = (size_t)q.n * bg.tiles_x * bg.tiles_y;
    const size_t tiles_ub = (size_t)q.n * ((q.w + 15) / 16) * ((q.h + 15) / 16);
    const size_t chunks_ub = (size_t)(q.G + 1) / 2;
    FarWs ws;
    char* base = (char*)wsp;
    ws.hd = (WsHeader*)base;
    ws.dirty = (unsigned char*)(base + sizeof(WsHeader));
    ws.redo = (int*)(base + sizeof(WsHeader) + dirty_bytes(q));
    ws.acc64 = (unsigned long long*)(base + sizeof(WsHeader) + dirty_bytes(q) + flag_bytes(tiles_ub * chunks_ub));
    // flags and side buffer must be zero on entry; redo / merge kernels leave them zero on exit
    cudaError_t e = cudaMemsetAsync(wsp, 0, ws_clean ? sizeof(WsHeader) : bwd_tiled_workspace_bytes(q), st);
    if (e != cudaSuccess) return e;

    // ---- grad_offset / grad_mask ----
    const int max_cells = kMaxBoxBytes / kCellBytes;
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, max_cells);
    if (tg.bw * tg.bh > max_cells) return cudaErrorInvalidConfiguration;
    CUtensorMap map;
    if (!make_x_tensor_map(&map, x, q, dtype, tg.bw, tg.bh)) return cudaErrorNotSupported;
    const size_t smem_b = (size_t)(bg.acc_ints + bg.wsum_ints) * sizeof(int);
    const int threads_b = bg.tj == 32 ? 640 : 256;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        e = cudaFuncSetAttribute(bwd_gather_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kMaxBoxBytes + 8 * kGatherStageBytes<T>);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(bwd_scatter_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (32 * 32 * (kSCell + kSG)) * (int)sizeof(int));
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    KernelTiming& kt = kernel_timing();
    if (kt.enabled) cudaEventRecord(kt.ev[0], st);
    const unsigned grid_a = (unsigned)((size_t)q.n * tg.chunks * tg.tiles_h * tg.tiles_w);
    // (also leaves max|grad_out| in the workspace header: the fixed-point scale of the scatter kernel)
    bwd_gather_kernel<T><<<grid_a, 256, (size_t)tg.bw * tg.bh * kCellBytes + 8 * kGatherStageBytes<T>, st>>>(
        map, (const T*)x, (const T*)offset, (const T*)mask, (const T*)grad_out, (T*)grad_offset, (T*)grad_mask, ws.hd,
        q, tg);

    // ---- grad_x ----
    if (kt.enabled) cudaEventRecord(kt.ev[1], st);
    const unsigned grid_b = (unsigned)(tiles * bg.chunks);
    bwd_scatter_kernel<T><<<grid_b, threads_b, smem_b, st>>>((const T*)offset, (const T*)mask, (const T*)grad_out,
                                                             (T*)grad_x, ws, q, bg);
    if (kt.enabled) cudaEventRecord(kt.ev[2], st);
    redo_hot_kernel<T><<<grid_b, 256, bg.wsum_ints * sizeof(int), st>>>((const T*)offset, (const T*)mask,
                                                                        (const T*)grad_out, ws, q, bg);
    if (kt.enabled) cudaEventRecord(kt.ev[3], st);
    const size_t npg = (size_t)q.n * q.h * q.w * q.G;
    merge_far_kernel<T><<<(unsigned)((npg / 4 + 256) / 256), 256, 0, st>>>((T*)grad_x, ws, q, npg);
    if (kt.enabled) cudaEventRecord(kt.ev[4], st);
    count_launch(4);
    return cudaGetLastError();
}

cudaError_t launch_bwd_tiled(const void* x, const void* offset, const void* mask, const void* grad_out,
                             void* grad_x, void* grad_offset, void* grad_mask, void* ws, const KParams& q,
                             int dtype, bool ws_clean, cudaStream_t st) {
    return dtype == DCNV3_F32
               ? launch_bwd_tiled_t<float>(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, q, dtype, ws_clean, st)
               : launch_bwd_tiled_t<__nv_bfloat16>(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, q, dtype, ws_clean, st);
}

}  // namespace dcnv3
