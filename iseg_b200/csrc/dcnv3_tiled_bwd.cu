// Tiled DCNv3 backward (sm_100a): two kernels, both deterministic.
//
//  bwd_gather_kernel  -- grad_offset and grad_mask.  Same structure as the tiled forward: one CTA per
//      output tile x group chunk, the chunk's input box staged by one TMA load, every lane owns one
//      (pixel, group), gathers the 4 corner slabs of every tap (conflict-free rotated LDS.128) and forms
//      the four dot products <grad_out, I_k> with packed FFMA2; from them the two gradients
//      (SURVEY.md App. A.2).  No cross-lane reduction, no atomics.
//
//  bwd_scatter_kernel -- grad_x.  One CTA = one TJ x TJ tile J of INPUT pixels x one chunk of two groups.
//      It visits exactly the output pixels whose nominal sampling position lies in J (its HOME pixels;
//      every output pixel is home to exactly one tile) and adds, for each of their 9 x 4 landings,
//          q = round(G[c] * Wk / 2^32)            G = grad_out, Wk = mask * bilinear weight (fixed point)
//      to an int32 accumulator BOX in shared memory that covers J plus a ring of a few cells (and the
//      zero ring of tf.pad where J touches the image border, so that landings there need no test).
//      The adds are shared-memory integer atomics (ATOMS.ADD; 31.6 updates/clk/SM when conflict free,
//      tools/microbench2.cu): a lane owns one (pixel, group) and walks its 16 channels in an order
//      XOR-rotated by its pixel index, so the 32 lanes of every ATOMS hit 32 different banks whatever
//      cells they land on; the four corners of a tap share one set of rotated addresses (the box pitch
//      is a compile-time constant, the other three corners are immediate offsets).
//      Integer addition is associative: the sums are bitwise independent of warp scheduling.
//      At the end the cells of J are converted and stored once (plain stores, no float atomic anywhere);
//      the ring cells -- they belong to neighbouring tiles -- are added to a 64-bit fixed-point side
//      buffer with integer global atomics and flagged in a dirty map; merge_far_kernel folds the side
//      buffer into grad_x.
//
//  Overflow / far taps -- every tap also adds ceil(|mask| * 1025) to a counter of its anchor cell; the
//      weight that reached a cell is bounded by the four counters around it, and while that bound stays
//      <= 15*1024 the int32 accumulator provably cannot wrap.  Cells that exceed it ("hot", only with
//      adversarial inputs) are skipped at flush and recomputed exactly by redo_hot_kernel straight into
//      the 64-bit side buffer; landings beyond the box (|offset| larger than the ring) take the same
//      side path directly.  The side buffer uses the SAME integers q; which path a contribution takes is
//      a deterministic function of the inputs.
//
// The common scale comes from max|grad_out| (left in the workspace header by bwd_gather_kernel):
// G = round(go * 2^eg), 2^29 <= max|G| < 2^30.
#include <algorithm>
#include <type_traits>

#include "dcnv3_kernels.h"
#include "dcnv3_tiled.cuh"

namespace dcnv3 {

#ifndef DCNV3_SCATTER_THREADS
#define DCNV3_SCATTER_THREADS 512  // 16 warps x 128 registers: a 32x32 tile is 64 blocks = 4 full rounds
#endif
#ifndef DCNV3_SCATTER_BATCH
#define DCNV3_SCATTER_BATCH 3      // taps whose coordinate arithmetic is done together before their ATOMS runs (0: per-tap loop)
#endif

constexpr int kSG = 2;             // groups per scatter CTA (lane = pixel * kSG + group, 16 pixels per warp)
constexpr int kSCell = kSG * kGC;   // accumulator ints per cell: exactly 32 banks wide, so the bank of an
                                   // update depends only on (group, channel) and never on the cell
constexpr int kBudget = 15 * 1024;  // 1024 * (sum of |Wk|) allowed in one int32 accumulator: |q| <= 2^27 |Wk|
constexpr int kWShift = 29;        // Wk fixed point: round(Wk * 2^29), |Wk| < 4
constexpr int kRingLo = 4;         // largest ring below / above a tile (cells); box pitch = TJ + 9
constexpr int kRingHi = 5;

// compile-time shape of the scatter kernel for a tile edge of TJ input cells
template <int TJ>
struct ScatterShape {
    static constexpr int PITCH = TJ + kRingLo + kRingHi;   // accumulator row pitch in cells (41 / 25)
    // pitch of the weight counters: one extra column, and odd, so that the 16 pixels of a warp -- they walk
    // down a column of cells -- spread their counters over all banks
    static constexpr int WPITCH = (PITCH + 1) | 1;
    static constexpr int THREADS = TJ == 32 ? DCNV3_SCATTER_THREADS : 256;  // batched walk: 16 warps x 128 registers
    static constexpr int THREADS_PER_TAP = TJ == 32 ? 640 : 256;             // per-tap walk: 20 warps x 96 registers
    static constexpr int MIN_CTAS = TJ == 32 ? 1 : 2;
};

struct BwdGeom {
    int tj, tj_log2;       // input tile edge in cells (16 or 32) and its log2
    int tiles_x, tiles_y;  // tiles of tj x tj un-padded input pixels
    int chunks;            // ceil(G / kSG)
    int ring_lo, ring_hi;  // ring of box cells kept below / above the tile (clipped to the image + zero ring)
    int box_rows;          // rows of the largest box (<= tj + ring_lo + ring_hi)
    int narrow;            // 1: the ring serves less than |offset| <= 3 (offset_scale > 1): per-tap walk
    // launch order of the tiles of an image, largest number of home pixels first (the tiles differ by up to 20 %: with
    // 3.5 waves of CTAs the last, partial wave should hold the small ones); identity when there are more than 64 tiles
    int ordered;
    unsigned char order_x[64], order_y[64];
};

// CTA index -> (image, group chunk, tile): tiles in launch order are the slow index, so that all (image, chunk) pairs of
// the largest tile come first
struct ScatterCta { int n, chunk, jx, jy; };
__device__ __forceinline__ ScatterCta decode_scatter_cta(const BwdGeom& bg, int n_images, int b) {
    ScatterCta c;
    if (bg.ordered) {
        const int pairs = n_images * bg.chunks;
        const int rank = b / pairs, nc = b - rank * pairs;
        c.jx = bg.order_x[rank]; c.jy = bg.order_y[rank];
        c.chunk = nc % bg.chunks; c.n = nc / bg.chunks;
    } else {
        c.jx = b % bg.tiles_x; b /= bg.tiles_x;
        c.jy = b % bg.tiles_y; b /= bg.tiles_y;
        c.chunk = b % bg.chunks; c.n = b / bg.chunks;
    }
    return c;
}

struct FarWs {
    WsHeader* hd;
    ImgMax* img_max;              // [N]  per-image max |grad_out| (bit pattern), written by the gather kernel
    unsigned char* dirty;         // [N][H][W][G]  (pixel, group) has side-buffer contributions
    int* redo;                    // [N][chunks][tiles_y][tiles_x]  scatter CTA found hot cells
    unsigned long long* acc64;    // [N][H][W][C] fixed point, zero outside a call
};

// Transposed copy of the per-pixel side inputs, written by the gather kernel (it has them staged in shared
// memory anyway) and read by the scatter kernel.  In the reference layout a lane's 18 offsets sit 72 bytes from its
// neighbour's, so each of the scatter kernel's per-tap loads touches 16-18 cache lines -- measured as a quarter of
// that kernel's time (profiles/r02_scatter_ablation.md).  Layout here: [image][pair of groups][tap][pixel][2 groups],
// offsets as (x, y) pairs, so the 32 lanes of a scatter warp (16 consecutive pixels x 2 groups) read 256 / 128
// contiguous bytes per tap.  Lives in the scratch tail of the workspace (no zero-on-entry contract).
template <typename T>
struct SideT {
    using Pair = typename std::conditional<sizeof(T) == 4, float2, unsigned>::type;
    Pair* off;   // [N][CS][9][H*W][2]
    T* msk;      // [N][CS][9][H*W][2]
    int cs;      // pairs of groups: ceil(G / 2)
    int hw;      // pixels per image
    __device__ __forceinline__ size_t index(int n, int pair, int pix, int gl) const {
        return (((size_t)n * cs + pair) * kTaps * hw + pix) * 2 + gl;  // tap 0; tap p is p * 2 * hw further
    }
};
template <typename T>
__device__ __forceinline__ void side_t_store(const SideT<T>& s, size_t idx, float ox, float oy, float ml);
template <>
__device__ __forceinline__ void side_t_store<float>(const SideT<float>& s, size_t idx, float ox, float oy, float ml) {
    s.off[idx] = make_float2(ox, oy);
    s.msk[idx] = ml;
}
template <>
__device__ __forceinline__ void side_t_store<__nv_bfloat16>(const SideT<__nv_bfloat16>& s, size_t idx, float ox, float oy, float ml) {
    s.off[idx] = (__float_as_uint(ox) >> 16) | (__float_as_uint(oy) & 0xffff0000u);  // (both are exact bf16 values)
    s.msk[idx] = __float2bfloat16_rn(ml);
}
template <typename T>
__device__ __forceinline__ void side_t_load(const SideT<T>& s, size_t idx, float& ox, float& oy, float& ml);
template <>
__device__ __forceinline__ void side_t_load<float>(const SideT<float>& s, size_t idx, float& ox, float& oy, float& ml) {
    const float2 o = __ldcg(s.off + idx);   // (written by the previous kernel of the stream: not through the
    ox = o.x; oy = o.y;                     //  non-coherent path)
    ml = __ldcg(s.msk + idx);
}
template <>
__device__ __forceinline__ void side_t_load<__nv_bfloat16>(const SideT<__nv_bfloat16>& s, size_t idx, float& ox, float& oy, float& ml) {
    const unsigned r = __ldcg(s.off + idx);
    ox = __uint_as_float(r << 16); oy = __uint_as_float(r & 0xffff0000u);
    ml = __uint_as_float((unsigned)__ldcg(reinterpret_cast<const unsigned short*>(s.msk) + idx) << 16);
}
// fp32 only: for bf16 the per-lane loads touch half as many lines to begin with, and the gather kernel's stores of
// the copy (lane = pixel x 4 groups there: 8-byte runs) cost more than the scatter kernel gains (profiles/r02_ab.md)
template <typename T>
constexpr bool kUseSideT = sizeof(T) == 4;
static size_t side_t_bytes(const KParams& q, int dtype) {
    if (dtype != DCNV3_F32) return 0;
    const size_t es = dtype == DCNV3_F32 ? 4 : 2;
    const size_t entries = (size_t)q.n * ((q.G + 1) / 2) * kTaps * q.h * q.w * 2;
    return (entries * 2 * es + 255) / 256 * 256 + (entries * es + 255) / 256 * 256;
}

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

// home range along one axis for tile index j: every i with nominal(i) in [j*tj, j*tj + tj); the first
// tile also takes nominal < 0 and the last one nominal >= extent
template <typename F>
__device__ __forceinline__ Range home_range(F nominal, int n, int j, int ntiles, int tj, int num, int den) {
    Range r;
    r.lo = (j == 0) ? 0 : first_ge(nominal, n, j * tj, num, den);
    r.hi = (j == ntiles - 1) ? n : first_ge(nominal, n, j * tj + tj, num, den);
    return r;
}

// called by threads 0 and 1 of the CTA, one range each
__device__ __forceinline__ void tile_ranges(const KParams& q, const BwdGeom& bg, int jx, int jy, Range& home_h,
                                            Range& home_w) {
    auto nx = [&](int h) { return nominal_ux(q, h); };
    auto ny = [&](int w) { return nominal_uy(q, w); };
    // output rows h walk along input x, output columns w along input y
    if (threadIdx.x == 0) home_h = home_range(nx, q.ho, jx, bg.tiles_x, bg.tj, q.win - 2, q.hin);
    if (threadIdx.x == 1) home_w = home_range(ny, q.wo, jy, bg.tiles_y, bg.tj, q.hin - 2, q.win);
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
static_assert(kGatherStageBytes<float> % 128 == 0 && kGatherStageBytes<__nv_bfloat16> % 128 == 0, "TMA slots are 128-byte aligned");

// The kernel that runs last in a backward call leaves the workspace prefix zeroed for the next call: its
// last CTA to finish clears the per-image maxima and the counter itself (the launch needs no memset node,
// which would break the programmatic-launch chain).
__device__ __forceinline__ void finalize_workspace(const FarWs& ws, int n_images) {
    __shared__ bool s_last;
    __syncthreads();  // every thread of this CTA is done reading the maxima
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&ws.hd->done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    unsigned* m = reinterpret_cast<unsigned*>(ws.img_max);
    for (int i = threadIdx.x; i < 2 * n_images; i += blockDim.x) m[i] = 0u;
    if (threadIdx.x == 0) {
        ws.hd->any_redo = 0u;
        ws.hd->done = 0u;
    }
}

// grad_out of one (pixel, group) in fixed point: G[c] = round(go[c] * 2^eg)
// ... XOR-permuted by the lane's pixel index: G[c] holds channel c ^ rot.  At step c every lane updates
// channel c ^ rot, so the 32 lanes of one ATOMS (16 pixels x 2 groups) hit 32 different banks
// whatever cells they land on.
template <typename T>
__device__ __forceinline__ void load_go(const T* go, f2 (&gf)[8]) {
    using C = Chunk<T>;  // only for the 16-byte piece geometry of T
#pragma unroll
    for (int pc = 0; pc < C::NPIECE; ++pc) load_piece<T>(go + pc * C::CH_PER_PIECE, gf + pc * C::PAIRS);
}
// (oms = 1 - centre feature scale of the (pixel, group), 1 without the blend: the core's output gradient is
//  RN(grad_out * oms), the product whose maximum the gather kernel took for the scale sg)
__device__ __forceinline__ void fixed_point_go(const f2 (&gf)[8], float oms, float sg, int rot, int (&G)[16]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        G[2 * c] = __float2int_rn(__fmul_rn(lo_of(gf[c]), oms) * sg);
        G[2 * c + 1] = __float2int_rn(__fmul_rn(hi_of(gf[c]), oms) * sg);
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
template <typename T>
__device__ __forceinline__ void load_fixed_point_go(const T* go, float oms, float sg, int rot, int (&G)[16]) {
    f2 gf[8];
    load_go<T>(go, gf);
    fixed_point_go(gf, oms, sg, rot, G);
}

// =====================================================================================================
// grad_offset / grad_mask
// =====================================================================================================
template <typename T, bool STAGED, bool BLEND>
__global__ void __launch_bounds__(kTiledWarps * 32, 2)
bwd_gather_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap offmap,
                  const __grid_constant__ CUtensorMap goffmap, const T* __restrict__ x,
                  const T* __restrict__ offset, const T* __restrict__ mask, const T* __restrict__ grad_out,
                  T* __restrict__ grad_offset, T* __restrict__ grad_mask, ImgMax* __restrict__ img_max,
                  const SideT<T> side_t, const KParams q, const TileGeom tg) {
    using C = Chunk<T>;
    using RS = RowStage<T>;
    constexpr bool MIXED = sizeof(T) == 2 && DCNV3_BF16_MIXED;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t sbar[kTiledWarps];
    pdl_launch_dependents();

    const int box_bytes = tg.bw * tg.bh * kCellBytes;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
#pragma unroll
        for (int i = 0; i < kTiledWarps; ++i) mbar_init(&sbar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();
    const TileCtx ctx = decode_tile<C::PXW>(q, tg, blockIdx.x);
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)box_bytes);
        tma_load_4d(smem, &xmap, &bar, ctx.chunk * C::GQ * kGC, ctx.cx0 - q.pw, ctx.cy0 - q.ph, ctx.n);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g_l = lane % C::GQ, px_l = lane / C::GQ;
    const int rot = Slab<T>::rot_of(px_l);
    // The warp's slot: side inputs in (STAGED), results out -- every result overwrites the input value of the
    // same tap (identical layout), and the slot is stored once per row segment.  For the fused soft-max path an
    // fp32 park holds dL/dm_p (it aliases the mask part when T is fp32).
    unsigned char* st = smem + (size_t)box_bytes + warp * kGatherStageBytes<T>;
    float* park = sizeof(T) == 4 ? reinterpret_cast<float*>(st + RS::OFF_BYTES)
                                 : reinterpret_cast<float*>(st + RS::BYTES);
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    uint32_t sphase = 0;

    auto request = [&](const TileCtx& c, int it) {
        const int h = c.h0 + it / c.colblocks, wb = c.w0 + (it % c.colblocks) * C::PXW;
        const size_t pix0 = ((size_t)c.n * q.ho + h) * q.wo + wb;
        RS::request(st, &sbar[warp], &offmap, mask + (pix0 * q.G + c.chunk * C::GQ) * 9, q.G, c.chunk, wb, c.n * q.ho + h,
                    min(C::PXW, c.w0 + c.tw - wb), lane);
    };
    if (STAGED && warp < ctx.nit) request(ctx, warp);  // the warp's first row segment

    {
        const unsigned char* sbase = smem + g_l * (kGC * (int)sizeof(T));
        const int n = ctx.n, chunk = ctx.chunk, cx0 = ctx.cx0, cy0 = ctx.cy0;
        const int g = min(chunk * C::GQ + g_l, q.G - 1);  // phantom groups of a trailing chunk shadow the last one
        const int ng = min(C::GQ, q.G - chunk * C::GQ);   // real groups in this chunk
        bool waited = false;
        unsigned amax = 0u;  // max |grad_out| bits seen by this thread in this tile: the image's fixed-point scale

        // one warp iteration = PXW consecutive pixels of one output row
        for (int it = warp; it < ctx.nit; it += kTiledWarps) {
            const int h = ctx.h0 + it / ctx.colblocks, wb = ctx.w0 + (it % ctx.colblocks) * C::PXW;
            const int npx = min(C::PXW, ctx.w0 + ctx.tw - wb);
            const int w = wb + min(px_l, npx - 1);  // idle lanes shadow the last pixel (their slots are never stored)
            const size_t pix0 = ((size_t)n * q.ho + h) * q.wo + wb;
            // (half groups -- KParams::gsh -- never run the staged instance: it stays free of their arithmetic)
            const size_t pg = (pix0 + min(px_l, npx - 1)) * q.G + g, ps = STAGED ? pg : side_entry(q, pix0 + min(px_l, npx - 1), g);
            const T* offp = offset + ps * 18;
            const T* mskp = mask + ps * 9;
            // bf16: grad_out stays packed (gw) and meets the packed slabs in FHFMA; fp32: pairs (go)
            f2 go[MIXED ? 1 : 8];
            unsigned gw[MIXED ? 8 : 1];
            if constexpr (MIXED) {
#pragma unroll
                for (int pc = 0; pc < C::NPIECE; ++pc) {
                    const uint4 r = __ldg(reinterpret_cast<const uint4*>(grad_out + pg * kGC + Slab<T>::chan_of(pc, rot)));
                    gw[4 * pc + 0] = r.x; gw[4 * pc + 1] = r.y; gw[4 * pc + 2] = r.z; gw[4 * pc + 3] = r.w;
                }
            } else {
#pragma unroll
                for (int pc = 0; pc < C::NPIECE; ++pc)
                    load_piece<T>(grad_out + pg * kGC + Slab<T>::chan_of(pc, rot), go + pc * C::PAIRS);
            }
            auto go_pair = [&](int c) -> f2 {
                if constexpr (MIXED) return bf16x2_to_f2(gw[c]);
                else return go[c];
            };
            float ref0, ref1;
            ref_point(q, h, w, ref0, ref1);
            // centre-feature-scale blend around the op (dcn_v3.py:146): the core's output gradient is go * (1 - s)
            float cs = 0.f, oms = 1.f;
            if (BLEND) {
                cs = Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + ps);
                oms = __fsub_rn(1.0f, cs);
            }
            // this lane's entries of the transposed side copy (real pixels and real groups only)
            const bool t_on = kUseSideT<T> && side_t.off != nullptr && px_l < npx && chunk * C::GQ + g_l < q.G;
            size_t t_idx = side_t.index(n, chunk * (C::GQ / 2) + (g_l >> 1), h * q.wo + w, g_l & 1);
            if (STAGED) {
                cp_async_wait_all();
                __syncwarp();
                mbar_wait(&sbar[warp], sphase);
                sphase ^= 1;
            }
            float mx = 0.f, inv_sum = 1.f;
            if (logits) {
                if (STAGED) RS::softmax_stats(st, lane, mx, inv_sum);
                else softmax_stats9<T>(mskp, mx, inv_sum);
            }
            float ox, oy, ml, ox2 = 0.f, oy2 = 0.f, ml2 = 0.f;
            if (STAGED) {
                RS::tap(st, lane, 0, ox, oy, ml);
            } else {
                load_tap_inputs<T>(offp, mskp, 0, ox, oy, ml);
                load_tap_inputs<T>(offp, mskp, 1, ox2, oy2, ml2);
            }
            if (!waited) {
                mbar_wait(&bar, 0);
                waited = true;
            }
            float gm_dot_m = 0.f;
#pragma unroll 1
            for (int p = 0; p < kTaps; ++p) {
                const float cx = ox, cy = oy, cm = ml;
                if (t_on) side_t_store<T>(side_t, t_idx, cx, cy, cm);
                t_idx += 2 * (size_t)side_t.hw;
                if (STAGED) {
                    if (p + 1 < kTaps) RS::tap(st, lane, p + 1, ox, oy, ml);  // (read before tap p's results land)
                } else {
                    ox = ox2; oy = oy2; ml = ml2;
                    if (p + 2 < kTaps) load_tap_inputs<T>(offp, mskp, p + 2, ox2, oy2, ml2);  // two taps ahead
                }
                const Tap t = make_tap_live(q, ref0, ref1, p, cx, cy);
                const int bx = t.x0 - cx0, by = t.y0 - cy0;
                const bool inbox = bx >= 0 && bx + 1 < tg.bw && by >= 0 && by + 1 < tg.bh;
                const float mm = logits ? expf(cm - mx) * inv_sum : cm;
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;  // <go, I_k>: a=(y0,x0) b=(y1,x0) c=(y0,x1) d=(y1,x1)
                if (__builtin_expect(t.alive && !inbox, 0)) {
#pragma unroll 1
                    for (int kk = 0; kk < 4; ++kk) {
                        const T* src = global_slab_b(x, q, n, t.y0 + (kk & 1), t.x0 + (kk >> 1), g);
                        if (src == nullptr) continue;
                        f2 dk2 = 0ull;
#pragma unroll
                        for (int pc = 0; pc < C::NPIECE; ++pc) {
                            f2 v[C::PAIRS];
                            load_piece<T>(src + Slab<T>::chan_of(pc, rot), v);
#pragma unroll
                            for (int j = 0; j < C::PAIRS; ++j) ffma2v(dk2, v[j], go_pair(pc * C::PAIRS + j));
                        }
                        const float dk = lo_of(dk2) + hi_of(dk2);
                        if (kk == 0) d0 = dk; else if (kk == 1) d1 = dk; else if (kk == 2) d2 = dk; else d3 = dk;
                    }
                } else if constexpr (MIXED) {
                    // two chains per corner (even / odd channels), the association of the FFMA2 form below
                    const unsigned char* a = sbase + (size_t)(t.alive ? by * tg.bw + bx : 0) * kCellBytes;
                    unsigned va[8], vb[8];
                    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                    Slab<T>::load_packed(a, rot, va);
                    Slab<T>::load_packed(a + (size_t)tg.bw * kCellBytes, rot, vb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 0>(e0, va[c], gw[c]); fhfma_x<1, 1>(o0, va[c], gw[c]); }
                    Slab<T>::load_packed(a + kCellBytes, rot, va);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 0>(e1, vb[c], gw[c]); fhfma_x<1, 1>(o1, vb[c], gw[c]); }
                    Slab<T>::load_packed(a + (size_t)(tg.bw + 1) * kCellBytes, rot, vb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 0>(e2, va[c], gw[c]); fhfma_x<1, 1>(o2, va[c], gw[c]); }
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 0>(e3, vb[c], gw[c]); fhfma_x<1, 1>(o3, vb[c], gw[c]); }
                    d0 = e0 + o0;
                    d1 = e1 + o1;
                    d2 = e2 + o2;
                    d3 = e3 + o3;
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
                if (!STAGED && q.gsh) {  // half groups: the group's dot products are the sums over its two halves (adjacent lanes)
                    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
                    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
                    d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
                    d3 += __shfl_xor_sync(0xffffffffu, d3, 1);
                }
                // dead taps have all four deltas zero => zero gradients
                // (the dot products are taken with the unscaled grad_out; (1 - s) is applied to the results)
                const float g_m = t.dx1 * t.dy1 * d0 + t.dx1 * t.dy0 * d1 + t.dx0 * t.dy1 * d2 + t.dx0 * t.dy0 * d3;
                const float gxq = mm * oms * (t.dy1 * (d2 - d0) + t.dy0 * (d3 - d1));
                const float gyq = mm * oms * (t.dx1 * (d1 - d0) + t.dx0 * (d3 - d2));
                gm_dot_m += g_m * mm;  // = <grad_out, core output> once all taps are in
                // results go to the lane's slot positions (lane stride 72 / 36 bytes: conflict free)
                if (sizeof(T) == 4) {
                    *reinterpret_cast<float2*>(st + lane * RS::LANE_OFF + p * 8) = make_float2(gxq * q.fx, gyq * q.fy);
                } else {
                    *reinterpret_cast<unsigned*>(st + lane * RS::LANE_OFF + p * 4) = pack_bf16x2(gxq * q.fx, gyq * q.fy);
                }
                if (logits) park[lane * kTaps + p] = g_m;
                else if (sizeof(T) == 4) park[lane * kTaps + p] = g_m * oms;
                else *reinterpret_cast<__nv_bfloat16*>(st + RS::OFF_BYTES + lane * RS::LANE_MSK + p * 2) = __float2bfloat16_rn(g_m * oms);
            }
            // (max |grad_out| is taken here, after the taps: the first use of the freshly loaded grad_out is then
            //  the first tap's dot products, behind its coordinate arithmetic and shared-memory loads)
            // ... of what the scatter kernel will convert: RN(grad_out * (1 - s)), the same product there
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const f2 gc2 = go_pair(c);
                amax = max(amax, max(abs_bits(__fmul_rn(lo_of(gc2), oms)), abs_bits(__fmul_rn(hi_of(gc2), oms))));
            }
            if (BLEND) {
                // d s = sum_c grad_out[c] * (x_proj[c] - core[c]) = <grad_out, x_proj> - sum_p m_p * dL/dm_p
                f2 dx2 = 0ull;
#pragma unroll
                for (int pc = 0; pc < C::NPIECE; ++pc) {
                    f2 xo[C::PAIRS];
                    load_piece<T>(x + pg * kGC + Slab<T>::chan_of(pc, rot), xo);
#pragma unroll
                    for (int j = 0; j < C::PAIRS; ++j) ffma2v(dx2, xo[j], go_pair(pc * C::PAIRS + j));
                }
                float dxs = lo_of(dx2) + hi_of(dx2);
                bool writer = px_l < npx && chunk * C::GQ + g_l < q.G;
                if (!STAGED && q.gsh) {  // half groups: <grad_out, x_proj> over both halves, stored once
                    dxs += __shfl_xor_sync(0xffffffffu, dxs, 1);
                    writer = writer && (g_l & 1) == 0;
                }
                if (writer) Elem<T>::st(reinterpret_cast<T*>(q.grad_cfs) + ps, dxs - gm_dot_m);
            }
            if (logits) {
                // softmax Jacobian needs sum_p m_p*dL/dm_p: second sweep over this lane's own 9 values.  The logits
                // are still in the slot when T is bf16 (the park is a separate area); for fp32 the park has
                // overwritten them and they are read again from global memory (L1 / L2 hits)
#pragma unroll 1
                for (int p = 0; p < kTaps; ++p) {
                    const float lg = (STAGED && sizeof(T) == 2) ? RS::mask_at(st, lane, p) : Elem<T>::ld(mskp + p);
                    const float mm = expf(lg - mx) * inv_sum;
                    const float v = mm * oms * (park[lane * kTaps + p] - gm_dot_m);
                    if (sizeof(T) == 4) park[lane * kTaps + p] = v;
                    else *reinterpret_cast<__nv_bfloat16*>(st + RS::OFF_BYTES + lane * RS::LANE_MSK + p * 2) = __float2bfloat16_rn(v);
                }
            }
            if (STAGED) {
                RS::store_results(st, &goffmap, grad_mask + (pix0 * q.G + chunk * C::GQ) * 9, q.G, chunk, wb, n * q.ho + h,
                                  npx, lane);
                if (it + kTiledWarps < ctx.nit) request(ctx, it + kTiledWarps);  // the slot is free again
            } else {
                if (q.gsh)  // both halves of a group hold the same results: the even half's are stored
                    RS::store_off_msk_halves(st, grad_offset + side_entry(q, pix0, chunk * C::GQ) * 18,
                                             grad_mask + side_entry(q, pix0, chunk * C::GQ) * 9, q.G >> 1, npx, ng >> 1, lane);
                else
                    RS::store_off_msk(st, grad_offset + (pix0 * q.G + chunk * C::GQ) * 18,
                                      grad_mask + (pix0 * q.G + chunk * C::GQ) * 9, q.G, npx, ng, lane);
            }
        }
        if (!waited) mbar_wait(&bar, 0);  // never leave with a TMA in flight
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0 && amax != 0u) atomicMax(&img_max[n].go_bits, amax);
    }
}

// =====================================================================================================
// grad_x
// =====================================================================================================
// geometry of one scatter CTA (the same in the scatter and the redo kernel)
struct TileBox {
    int ux0, uy0, tjw, tjh;  // tile J in un-padded input coordinates
    int bx0, by0, bw, bh;    // accumulator box: un-padded origin (-1 = the zero ring of tf.pad) and extent
};
__device__ __forceinline__ TileBox make_box(const KParams& q, const BwdGeom& bg, int jx, int jy) {
    TileBox b;
    b.ux0 = jx << bg.tj_log2;
    b.uy0 = jy << bg.tj_log2;
    b.tjw = min(bg.tj, q.w - b.ux0);
    b.tjh = min(bg.tj, q.h - b.uy0);
    b.bx0 = max(b.ux0 - bg.ring_lo, -1);
    b.by0 = max(b.uy0 - bg.ring_lo, -1);
    b.bw = min(b.ux0 + b.tjw + bg.ring_hi, q.w + 1) - b.bx0;
    b.bh = min(b.uy0 + b.tjh + bg.ring_hi, q.h + 1) - b.by0;
    return b;
}

// shared-memory integer add without return value: 32-bit shared address + immediate byte offset
template <int OFF>
__device__ __forceinline__ void red_shared_add(uint32_t addr, int v) {
    asm volatile("red.shared.add.s32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

// Weight counters.  Every tap adds ceil(|m| * 1025) -- an upper bound of 1024 * (sum of its four |Wk|) --
// to ONE counter, that of its anchor cell (y0, x0); the counter array has one extra row and column so
// that anchors one cell outside the box can be counted.  The weight that reached cell (y, x) is then at
// most the sum of the four counters anchored at (y, x), (y-1, x), (y, x-1), (y-1, x-1).  A raw mask beyond
// the fixed-point range (|m| >= 3.9) makes the four cells hot by itself.
__device__ __forceinline__ int weight_units(float mm) {
    const int wb = __float2int_ru(fabsf(mm) * 1025.f);
    return wb < 3994 ? wb : kBudget + 1;
}
template <int WP>
__device__ __forceinline__ bool cell_is_hot(const int* wsum, int cy, int cx, int gl) {
    const int* c = wsum + (cy * WP + cx) * kSG + gl;  // anchor (cy-1, cx-1) lives at index [cy][cx]
    return (long long)c[0] + c[kSG] + c[WP * kSG] + c[WP * kSG + kSG] > kBudget;
}
// one contribution in fixed point: round(G * Wk / 2^32), ties up (a single IMAD.HI with a constant addend)
__device__ __forceinline__ int qmul(int g, int wq) {
    return (int)(((long long)g * wq + 0x80000000ll) >> 32);
}
// (|wf| < 3.9 whenever the tap's cells are not hot -- weight_units() -- so the conversion cannot saturate there;
//  for hot cells it may, deterministically, and those accumulators are discarded and recomputed)
__device__ __forceinline__ int weight_fixed(float wf) { return __float2int_rn(wf * (float)(1 << kWShift)); }

// one landing into the 64-bit side buffer: the same integers q as the shared-memory path; |Wk| >= 3.9
// (raw masks only) is pre-shifted so that the product still fits
__device__ __forceinline__ void side_add(const FarWs& ws, size_t cellg, const int (&G)[16], int px_l, float wf) {
    int sh = 0;
    if (!(fabsf(wf) < 3.9f)) sh = min(max((int)((__float_as_uint(wf) >> 23) & 0xffu) - 128, 0), 30);
    const int wq = __float2int_rn(ldexpf(wf, kWShift - sh));
    unsigned long long* dst = ws.acc64 + cellg * kGC;
#pragma unroll
    for (int c = 0; c < 16; ++c)  // G[c] holds channel c ^ px_l
        atomicAdd(dst + (c ^ px_l), (unsigned long long)((long long)qmul(G[c], wq) << sh));
    ws.dirty[cellg] = 1;
}

// The walk of the redo kernel over the home pixels of one tile (same enumeration of blocks and taps as the scatter
// kernel's walk below, per-tap loop, inputs straight from the reference-layout tensors):
//   MODE 1 (pass 1): rebuilds the weight counters exactly as the scatter kernel left them
//   MODE 2 (pass 2): landings on hot cells of the box -> 64-bit side buffer
template <typename T, int MODE, int TJ, bool SPLIT = true>  // SPLIT: half groups possible (KParams::gsh read at run time)
__device__ __forceinline__ void redo_walk(int* wsum, const T* __restrict__ offset, const T* __restrict__ mask,
                                          const T* __restrict__ grad_out, const FarWs& ws, const KParams& q,
                                          const TileBox& box, int n, int chunk, Range hh, Range hw, int eg) {
    static_assert(MODE == 1 || MODE == 2, "redo passes only");
    constexpr int WP = ScatterShape<TJ>::WPITCH;  // pitch of the weight counters
    constexpr int PXW = 32 / kSG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g_l = lane % kSG, px_l = lane / kSG;
    const int g = chunk * kSG + g_l;
    if (g >= q.G) return;  // phantom group of a trailing chunk (no warp-level synchronisation in this walk)
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const float sg = ldexpf(1.0f, eg);  // G = round(go * 2^eg), |G| < 2^30
    const size_t img_pixels = (size_t)q.h * q.w;
    const uint32_t wsum_s = smem_u32(wsum) + (uint32_t)g_l * 4u;  // the lane's counter of anchor (-1, -1)
    const int nw = hw.hi - hw.lo, npix = (hh.hi - hh.lo) * nw;
    const int nblocks = (npix + PXW - 1) / PXW;
#pragma unroll 1
    for (int blk = warp; blk < nblocks; blk += nwarps) {
        const int pix = blk * PXW + px_l;
        if (pix >= npix) continue;
        const int row = pix / nw;
        const int h = hh.lo + row, w = hw.lo + (pix - row * nw);
        const size_t pixel = ((size_t)n * q.ho + h) * q.wo + w;
        const size_t pg = pixel * q.G + g, ps = SPLIT ? side_entry(q, pixel, g) : pg;
        const T* offp = offset + ps * 18;
        const T* mskp = mask + ps * 9;
        int G[16];
        bool have_g = false;
        float mx = 0.f, inv_sum = 1.f;
        if (logits) softmax_stats9<T>(mskp, mx, inv_sum);
        float ref0, ref1;
        ref_point(q, h, w, ref0, ref1);
#pragma unroll 1
        for (int p = 0; p < kTaps; ++p) {
            float ox, oy, ml;
            load_tap_inputs<T>(offp, mskp, p, ox, oy, ml);
            const Axis axx = axis_x_live(q, ref0, p, ox);
            const Axis axy = axis_y_live(q, ref1, p, oy);
            if (!(axx.alive && axy.alive)) continue;  // a clipped corner pair coincides: contributes exactly 0
            const int lx = axx.i0 - q.pw - box.bx0;   // corner (y0,x0) relative to the box
            const int ly = axy.i0 - q.ph - box.by0;
            const float mm = logits ? expf(ml - mx) * inv_sum : ml;
            if (mm == 0.f) continue;
            const bool touches = (unsigned)(lx + 1) <= (unsigned)box.bw && (unsigned)(ly + 1) <= (unsigned)box.bh;
            if (!touches) continue;  // no corner in the box: counted nowhere, and its cells are not this box's
            if (MODE == 1) {
                red_shared_add<(WP + 1) * kSG * 4>(wsum_s + (uint32_t)(ly * WP + lx) * (kSG * 4u), weight_units(mm));
            }
#pragma unroll 1
            for (int k = 0; k < (MODE == 2 ? 4 : 0); ++k) {
                const float wf = ((k >> 1) ? axx.d0 : axx.d1) * ((k & 1) ? axy.d0 : axy.d1) * mm;
                if (wf == 0.f) continue;
                const int cx = lx + (k >> 1), cy = ly + (k & 1);
                const int ax = cx + box.bx0, ay = cy + box.by0;  // un-padded image coordinates
                if (!((unsigned)cx < (unsigned)box.bw && (unsigned)cy < (unsigned)box.bh)) continue;
                if (!(ax >= 0 && ax < q.w && ay >= 0 && ay < q.h) || !cell_is_hot<WP>(wsum, cy, cx, g_l)) continue;
                if (!have_g) {
                    const float oms = q.cfs != nullptr ? __fsub_rn(1.0f, Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + ps)) : 1.0f;
                    load_fixed_point_go<T>(grad_out + pg * kGC, oms, sg, px_l, G);
                    have_g = true;
                }
                side_add(ws, ((size_t)n * img_pixels + (size_t)ay * q.w + ax) * q.G + g, G, px_l, wf);
            }
        }
    }
}

// ---- per-tap walk of the scatter kernel (round 1), kept for configurations whose ring cannot serve |offset| <= 3 --
// With offset_scale 2 (InternImage-L) the 4 / 5-cell ring only covers |offset| <= 1 and a third of the taps leave the
// box at least partly: they go corner by corner, and beyond the box through 64-bit global atomics.  There the
// batched walk below loses (its runs are built for taps that are wholly inside; 1643 vs 979 us at 160x160 C160 G10
// bf16, profiles/r02_scatter_study.md): this loop handles a tap's corners right where its coordinates are, with 20
// warps to hide the global atomics.  Inputs straight from the reference-layout tensors, two taps ahead.
// A work item is one block of 16 pixels x 2 groups; blocks are dealt round-robin to the warps, and the
// blocks of the last, incomplete round are split by taps over the warps that would otherwise idle.
template <typename T, int MODE, int TJ, bool BLEND>
__device__ __forceinline__ void scatter_walk_per_tap(int* acc, int* wsum, const T* __restrict__ offset,
                                             const T* __restrict__ mask, const T* __restrict__ grad_out,
                                             const FarWs& ws, const KParams& q, const TileBox& box, int n, int chunk,
                                             Range hh, Range hw, int eg) {
    constexpr int PITCH = ScatterShape<TJ>::PITCH;
    constexpr int WP = ScatterShape<TJ>::WPITCH;  // pitch of the weight counters
    constexpr int PXW = 32 / kSG;
    constexpr int ROWB = PITCH * kSCell * 4;  // bytes between accumulator rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g_l = lane % kSG, px_l = lane / kSG;
    const int g = chunk * kSG + g_l;
    if (g >= q.G) return;  // phantom group of a trailing chunk (no warp-level synchronisation in this walk)
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const float sg = ldexpf(1.0f, eg);  // G = round(go * 2^eg), |G| < 2^30
    const size_t img_pixels = (size_t)q.h * q.w;
    // the lane's slab inside cell 0, pre-rotated: slab bases are 64-byte aligned, so
    // base + ((c ^ rot) * 4) == (base ^ rot*4) ^ c*4
    uint32_t acc_s = MODE == 0 ? ((smem_u32(acc) + (uint32_t)g_l * (kGC * 4u)) ^ ((uint32_t)px_l << 2)) : 0u;
    // the lane's counter of anchor (-1, -1)
    uint32_t wsum_s = MODE != 2 ? smem_u32(wsum) + (uint32_t)g_l * 4u : 0u;
    // opaque to the compiler: otherwise it re-derives both from %tid inside the tap loop (S2R latency)
    asm volatile("" : "+r"(acc_s), "+r"(wsum_s));
    const int nw = hw.hi - hw.lo, npix = (hh.hi - hh.lo) * nw;
    const int nblocks = (npix + PXW - 1) / PXW;
    const int full_rounds = nblocks / nwarps, rest = nblocks - full_rounds * nwarps;
    const int parts = rest ? min(kTaps, nwarps / rest) : 1;  // warps per block of the last round
#pragma unroll 1
    for (int round = 0; round <= full_rounds; ++round) {
        int blk = round * nwarps + warp, p_lo = 0, p_hi = kTaps;
        if (round == full_rounds) {
            if (warp >= rest * parts) break;
            const int part = warp % parts;
            blk = round * nwarps + warp / parts;
            p_lo = part * kTaps / parts;
            p_hi = (part + 1) * kTaps / parts;
        }
        const int pix = blk * PXW + px_l;
        if (pix >= npix) continue;
        const int row = pix / nw;
        const int h = hh.lo + row, w = hw.lo + (pix - row * nw);
        const size_t pixel = ((size_t)n * q.ho + h) * q.wo + w;
        const size_t pg = pixel * q.G + g, ps = side_entry(q, pixel, g);
        const T* offp = offset + ps * 18;
        const T* mskp = mask + ps * 9;
        float ox, oy, ml, ox2, oy2, ml2;
        load_tap_inputs<T>(offp, mskp, p_lo, ox, oy, ml);
        load_tap_inputs<T>(offp, mskp, min(p_lo + 1, kTaps - 1), ox2, oy2, ml2);
        int G[16];
        f2 gf[8];
        bool have_g = false;
        if (MODE == 0) load_go<T>(grad_out + pg * kGC, gf);
        float mx = 0.f, inv_sum = 1.f;
        if (logits) softmax_stats9<T>(mskp, mx, inv_sum);
        float ref0, ref1;
        ref_point(q, h, w, ref0, ref1);
        const float oms = BLEND ? __fsub_rn(1.0f, Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + ps)) : 1.0f;
        if (MODE == 0) {  // every tap of a home pixel lands: convert grad_out once, with the whole warp converged
            fixed_point_go(gf, oms, sg, px_l, G);
            have_g = true;
        }
#pragma unroll 1
        for (int p = p_lo; p < p_hi; ++p) {
            const float cxo = ox, cyo = oy, cm = ml;
            ox = ox2; oy = oy2; ml = ml2;
            if (p + 2 < kTaps) load_tap_inputs<T>(offp, mskp, p + 2, ox2, oy2, ml2);  // two taps ahead
            const Axis axx = axis_x_live(q, ref0, p, cxo);
            const Axis axy = axis_y_live(q, ref1, p, cyo);
            if (!(axx.alive && axy.alive)) continue;  // a clipped corner pair coincides: contributes exactly 0
            const int lx = axx.i0 - q.pw - box.bx0;   // corner (y0,x0) relative to the box
            const int ly = axy.i0 - q.ph - box.by0;
            const float mm = logits ? expf(cm - mx) * inv_sum : cm;
            if (mm == 0.f) continue;
            if (MODE == 0 &&
                __builtin_expect((unsigned)lx < (unsigned)(box.bw - 1) && (unsigned)ly < (unsigned)(box.bh - 1), 1)) {
                // all four corners a b c d = (y0,x0) (y1,x0) (y0,x1) (y1,x1) lie in the box
                const int cell = ly * PITCH + lx;
                red_shared_add<(WP + 1) * kSG * 4>(wsum_s + (uint32_t)(ly * WP + lx) * (kSG * 4u), weight_units(mm));
                const int wqa = weight_fixed(axx.d1 * axy.d1 * mm), wqb = weight_fixed(axx.d1 * axy.d0 * mm);
                const int wqc = weight_fixed(axx.d0 * axy.d1 * mm), wqd = weight_fixed(axx.d0 * axy.d0 * mm);
                const uint32_t base = acc_s + (uint32_t)cell * (kSCell * 4u);  // rotation bits stay put: cell*128
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const uint32_t a = base ^ (uint32_t)(c << 2);
                    red_shared_add<0>(a, qmul(G[c], wqa));
                    red_shared_add<ROWB>(a, qmul(G[c], wqb));
                    red_shared_add<kSCell * 4>(a, qmul(G[c], wqc));
                    red_shared_add<ROWB + kSCell * 4>(a, qmul(G[c], wqd));
                }
                continue;
            }
            // ---- corner by corner: part of the patch leaves the box (or one of the redo passes) ----
            const bool touches = (unsigned)(lx + 1) <= (unsigned)box.bw && (unsigned)(ly + 1) <= (unsigned)box.bh;
            if (MODE != 2 && touches)  // some corner lies in the box
                red_shared_add<(WP + 1) * kSG * 4>(wsum_s + (uint32_t)(ly * WP + lx) * (kSG * 4u), weight_units(mm));
#pragma unroll 1
            for (int k = 0; k < (MODE == 1 || (MODE == 2 && !touches) ? 0 : 4); ++k) {
                const float wf = ((k >> 1) ? axx.d0 : axx.d1) * ((k & 1) ? axy.d0 : axy.d1) * mm;
                if (wf == 0.f) continue;
                const int cx = lx + (k >> 1), cy = ly + (k & 1);
                const int ax = cx + box.bx0, ay = cy + box.by0;  // un-padded image coordinates
                const bool in_image = ax >= 0 && ax < q.w && ay >= 0 && ay < q.h;
                const size_t cellg = ((size_t)n * img_pixels + (size_t)ay * q.w + ax) * q.G + g;
                if ((unsigned)cx < (unsigned)box.bw && (unsigned)cy < (unsigned)box.bh) {
                    if (MODE == 2) {
                        if (in_image && cell_is_hot<WP>(wsum, cy, cx, g_l)) {
                            if (!have_g) {
                                load_fixed_point_go<T>(grad_out + pg * kGC, oms, sg, px_l, G);
                                have_g = true;
                            }
                            side_add(ws, cellg, G, px_l, wf);
                        }
                        continue;
                    }
                    const int wq = weight_fixed(wf);
                    const uint32_t base = acc_s + (uint32_t)(cy * PITCH + cx) * (kSCell * 4u);
#pragma unroll
                    for (int c = 0; c < 16; ++c) red_shared_add<0>(base ^ (uint32_t)(c << 2), qmul(G[c], wq));
                } else if (MODE == 0 && in_image) {  // beyond the ring; outside the image the gradient is dropped
                    side_add(ws, cellg, G, px_l, wf);
                }
            }
        }
    }
}

// ---- batched walk of the scatter kernel ---------------------------------------------------------------------
// Inside a tap the coordinate arithmetic (a dependent chain of ~25 fp32 operations per axis, fed by global loads)
// and the 64 ATOMS of its four landings alternate, and because every warp of the CTA runs the same instruction
// stream at the same pace they alternate in step: while the warps sit in the chain the ATOMS pipe -- the resource
// the kernel is bound by -- idles (ncu, round 2: 83 % busy inside the ATOMS runs, 30 % over the kernel).  Here the
// chains of TB taps are computed together, branch free, so that they overlap one another (instruction-level
// parallelism instead of one exposed latency per tap), their results stay in registers as TapRec, and the ATOMS
// runs then follow back to back; the next block's inputs are requested before the last run.  Taps that are not
// wholly inside the box (rare) are noted in a bit mask and handled after the runs by slow_tap().
struct TapRec {
    uint32_t base;           // lane's rotated slab address inside the anchor cell (y0, x0)
    int wqa, wqb, wqc, wqd;  // fixed-point weights of the corners (y0,x0) (y1,x0) (y0,x1) (y1,x1)
};

template <int ROWB>
__device__ __forceinline__ void atoms_run(const TapRec& r, const int (&G)[16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const uint32_t a = r.base ^ (uint32_t)(c << 2);
        red_shared_add<0>(a, qmul(G[c], r.wqa));
        red_shared_add<ROWB>(a, qmul(G[c], r.wqb));
        red_shared_add<kSCell * 4>(a, qmul(G[c], r.wqc));
        red_shared_add<ROWB + kSCell * 4>(a, qmul(G[c], r.wqd));
    }
}

// a tap whose 2x2 patch leaves the box: corner by corner -- shared atomics where the corner is in the box, the
// 64-bit side buffer where it lies beyond it (outside the image the gradient is dropped)
template <typename T, int TJ>
__device__ __forceinline__ void slow_tap(uint32_t acc_s, uint32_t wsum_s, const T* __restrict__ offp, const T* __restrict__ mskp,
                                      const FarWs& ws, const KParams& q, const TileBox& box, int n, int g, int px_l, int p,
                                      float ref0, float ref1, float mx, float inv_sum, const int (&G)[16]) {
    constexpr int PITCH = ScatterShape<TJ>::PITCH;
    constexpr int WP = ScatterShape<TJ>::WPITCH;
    float ox, oy, ml;
    load_tap_inputs<T>(offp, mskp, p, ox, oy, ml);
    const Axis axx = axis_x_live(q, ref0, p, ox);
    const Axis axy = axis_y_live(q, ref1, p, oy);
    const int lx = axx.i0 - q.pw - box.bx0, ly = axy.i0 - q.ph - box.by0;
    const float mm = (q.flags & DCNV3_FLAG_MASK_LOGITS) ? expf(ml - mx) * inv_sum : ml;
    const bool touches = (unsigned)(lx + 1) <= (unsigned)box.bw && (unsigned)(ly + 1) <= (unsigned)box.bh;
    if (touches)  // some corner lies in the box
        red_shared_add<(WP + 1) * kSG * 4>(wsum_s + (uint32_t)(ly * WP + lx) * (kSG * 4u), weight_units(mm));
    const size_t img_pixels = (size_t)q.h * q.w;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const float wf = ((k >> 1) ? axx.d0 : axx.d1) * ((k & 1) ? axy.d0 : axy.d1) * mm;
        if (wf == 0.f) continue;
        const int cx = lx + (k >> 1), cy = ly + (k & 1);
        const int ax = cx + box.bx0, ay = cy + box.by0;  // un-padded image coordinates
        if ((unsigned)cx < (unsigned)box.bw && (unsigned)cy < (unsigned)box.bh) {
            const int wq = weight_fixed(wf);
            const uint32_t base = acc_s + (uint32_t)(cy * PITCH + cx) * (kSCell * 4u);
#pragma unroll
            for (int c = 0; c < 16; ++c) red_shared_add<0>(base ^ (uint32_t)(c << 2), qmul(G[c], wq));
        } else if (ax >= 0 && ax < q.w && ay >= 0 && ay < q.h) {
            side_add(ws, ((size_t)n * img_pixels + (size_t)ay * q.w + ax) * q.G + g, G, px_l, wf);
        }
    }
}

template <typename T, int TJ, int TB, bool BLEND>
__device__ __forceinline__ void scatter_walk_batched(int* acc, int* wsum, const T* __restrict__ offset,
                                                     const T* __restrict__ mask, const T* __restrict__ grad_out,
                                                     const SideT<T>& side_t, const FarWs& ws, const KParams& q,
                                                     const TileBox& box, int n, int chunk, Range hh, Range hw, int eg) {
    constexpr int PITCH = ScatterShape<TJ>::PITCH;
    constexpr int WP = ScatterShape<TJ>::WPITCH;
    constexpr int PXW = 32 / kSG;
    constexpr int ROWB = PITCH * kSCell * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g_l = lane % kSG, px_l = lane / kSG;
    const int g = chunk * kSG + g_l;
    if (g >= q.G) return;  // phantom group of a trailing chunk (no warp-level synchronisation in this walk)
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const float sg = ldexpf(1.0f, eg);
    uint32_t acc_s = (smem_u32(acc) + (uint32_t)g_l * (kGC * 4u)) ^ ((uint32_t)px_l << 2);
    uint32_t wsum_s = smem_u32(wsum) + (uint32_t)g_l * 4u;
    asm volatile("" : "+r"(acc_s), "+r"(wsum_s));
    const int nw = hw.hi - hw.lo, npix = (hh.hi - hh.lo) * nw;
    const int nblocks = (npix + PXW - 1) / PXW;
    const int full_rounds = nblocks / nwarps, rest = nblocks - full_rounds * nwarps;
    const int parts = rest ? min(kTaps, nwarps / rest) : 1;  // warps per block of the last round
    const int bx_off = q.pw + box.bx0, by_off = q.ph + box.by0;
    const unsigned bw1 = (unsigned)(box.bw - 1), bh1 = (unsigned)(box.bh - 1);

    // the work item of a round: one block of 16 pixels x 2 groups (all 9 taps), or -- last, incomplete round -- a
    // range of its taps.  Returns false when this lane has nothing to do in that round.
    struct Item { int h, w, p_lo, p_hi; size_t pg; bool valid; };
    auto item_of = [&](int round) {
        Item it;
        it.valid = false; it.h = it.w = 0; it.pg = 0; it.p_lo = 0; it.p_hi = kTaps;
        if (round > full_rounds) return it;
        int blk = round * nwarps + warp;
        if (round == full_rounds) {
            if (warp >= rest * parts) return it;
            const int part = warp % parts;
            blk = round * nwarps + warp / parts;
            it.p_lo = part * kTaps / parts;
            it.p_hi = (part + 1) * kTaps / parts;
        }
        const int pix = blk * PXW + px_l;
        if (pix >= npix) return it;
        const int row = pix / nw;
        it.h = hh.lo + row; it.w = hw.lo + (pix - row * nw);
        it.pg = (((size_t)n * q.ho + it.h) * q.wo + it.w) * q.G + g;
        it.valid = true;
        return it;
    };
    struct Inputs { float ox[kTaps], oy[kTaps], ml[kTaps]; f2 gf[8]; float oms; };
    auto load_inputs = [&](const Item& it, Inputs& in) {
        // from the transposed copy the gather kernel has just written: consecutive lanes, consecutive entries
        if (kUseSideT<T>) {
            const size_t t_idx = side_t.index(n, chunk, it.h * q.wo + it.w, g_l);
#pragma unroll
            for (int p = 0; p < kTaps; ++p) side_t_load<T>(side_t, t_idx + (size_t)p * 2 * side_t.hw, in.ox[p], in.oy[p], in.ml[p]);
        } else {
            const T* offp = offset + it.pg * 18;
            const T* mskp = mask + it.pg * 9;
#pragma unroll
            for (int p = 0; p < kTaps; ++p) load_tap_inputs<T>(offp, mskp, p, in.ox[p], in.oy[p], in.ml[p]);
        }
        load_go<T>(grad_out + it.pg * kGC, in.gf);
        in.oms = BLEND ? __fsub_rn(1.0f, Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + it.pg)) : 1.0f;
    };

    Item cur = item_of(0);
    Inputs in;
    if (cur.valid) load_inputs(cur, in);
#pragma unroll 1
    for (int round = 0; round <= full_rounds; ++round) {
        const Item nxt = item_of(round + 1);
        bool requested = false;
        if (cur.valid) {
            int G[16];
            fixed_point_go(in.gf, in.oms, sg, px_l, G);
            float mx = 0.f, inv_sum = 1.f;
            if (logits) {
                mx = -INFINITY;
#pragma unroll
                for (int p = 0; p < kTaps; ++p) mx = fmaxf(mx, in.ml[p]);
                float s = 0.f;
#pragma unroll
                for (int p = 0; p < kTaps; ++p) s += expf(in.ml[p] - mx);
                inv_sum = 1.0f / s;
            }
            float ref0, ref1;
            ref_point(q, cur.h, cur.w, ref0, ref1);
            unsigned slow = 0u;
#pragma unroll
            for (int b0 = 0; b0 < kTaps; b0 += TB) {
                // a warp of the last, split round owns only some of the taps: batches wholly outside its range are
                // skipped (warp-uniform)
                if (!(b0 < cur.p_hi && b0 + TB > cur.p_lo)) {
                    if (b0 + TB >= kTaps && nxt.valid) {  // (the request that normally rides under tap 8's run)
                        load_inputs(nxt, in);
                        requested = true;
                    }
                    continue;
                }
                TapRec rec[TB];
                unsigned fast = 0u;
                // ---- the chains of TB taps, branch free ----
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    const int p = b0 + t;
                    if (p >= kTaps) continue;
                    const Axis axx = axis_x_live(q, ref0, p, in.ox[p]);
                    const Axis axy = axis_y_live(q, ref1, p, in.oy[p]);
                    const int lx = axx.i0 - bx_off, ly = axy.i0 - by_off;  // corner (y0,x0) relative to the box
                    const float mm = logits ? expf(in.ml[p] - mx) * inv_sum : in.ml[p];
                    // alive: no clipped corner pair coincides (else the tap contributes exactly 0)
                    const bool on = axx.alive && axy.alive && mm != 0.f && p >= cur.p_lo && p < cur.p_hi;
                    const bool inbox = (unsigned)lx < bw1 && (unsigned)ly < bh1;  // all four corners lie in the box
                    rec[t].wqa = weight_fixed(axx.d1 * axy.d1 * mm);
                    rec[t].wqb = weight_fixed(axx.d1 * axy.d0 * mm);
                    rec[t].wqc = weight_fixed(axx.d0 * axy.d1 * mm);
                    rec[t].wqd = weight_fixed(axx.d0 * axy.d0 * mm);
                    rec[t].base = acc_s + (uint32_t)(ly * PITCH + lx) * (kSCell * 4u);  // rotation bits stay put: cell*128
                    if (on && inbox) {
                        fast |= 1u << t;
                        red_shared_add<(WP + 1) * kSG * 4>(wsum_s + (uint32_t)(ly * WP + lx) * (kSG * 4u), weight_units(mm));
                    }
                    if (on && !inbox) slow |= 1u << p;
                }
                // ---- their ATOMS runs ----
#pragma unroll
                for (int t = 0; t < TB; ++t) {
                    const int p = b0 + t;
                    if (p >= kTaps) continue;
                    if (p == kTaps - 1 && nxt.valid) {     // the next block's inputs travel under the last run (this
                        load_inputs(nxt, in);           //  block's have all been consumed by now)
                        requested = true;
                    }
                    if (fast & (1u << t)) atoms_run<ROWB>(rec[t], G);
                }
            }
            if (__builtin_expect(slow != 0u, 0)) {
                const T* offp = offset + cur.pg * 18;
                const T* mskp = mask + cur.pg * 9;
                do {
                    const int p = __ffs(slow) - 1;
                    slow &= slow - 1;
                    slow_tap<T, TJ>(acc_s, wsum_s, offp, mskp, ws, q, box, n, g, px_l, p, ref0, ref1, mx, inv_sum, G);
                } while (slow != 0u);
            }
        }
        if (nxt.valid && !requested) load_inputs(nxt, in);
        cur = nxt;
    }
}

// Whole-image tiles (the image fits one scatter tile: every InternImage stage from 32x32 down): nobody else
// contributes to the tile's cells -- no ring, no far landings, no merge launch -- so the CTA that found hot cells
// recomputes them itself, right after its flush: the landings on hot cells go to the 64-bit side buffer (pass 2 of the
// redo; the weight counters of pass 1 are still in shared memory), and are then converted and added to what the flush
// left in grad_x (0, or the blend's direct term).  Leaves the side buffer zero again.
template <typename T, int TJ, bool SPLIT = true>
__device__ __forceinline__ void redo_own_tile(int* wsum, const T* __restrict__ offset, const T* __restrict__ mask,
                                              const T* __restrict__ grad_out, T* __restrict__ grad_x, const FarWs& ws,
                                              const KParams& q, const TileBox& box, int n, int chunk, Range hh, Range hw,
                                              int eg) {
    constexpr int PITCH = ScatterShape<TJ>::PITCH;
    constexpr int WP = ScatterShape<TJ>::WPITCH;
    redo_walk<T, 2, TJ, SPLIT>(wsum, offset, mask, grad_out, ws, q, box, n, chunk, hh, hw, eg);
    __threadfence();
    __syncthreads();
    const size_t img_pixels = (size_t)q.h * q.w;
    const double inv_d = ldexp(1.0, -(eg + kWShift - 32));
    for (int i = threadIdx.x; i < box.bh * (PITCH * kSCell); i += blockDim.x) {
        const int cy = i / (PITCH * kSCell), r = i - cy * (PITCH * kSCell);
        const int cx = r / kSCell, ch = r % kSCell, gl = ch / kGC;
        const int ax = box.bx0 + cx, ay = box.by0 + cy, g = chunk * kSG + gl;
        if (cx >= box.bw || ax < 0 || ax >= q.w || ay < 0 || ay >= q.h || g >= q.G) continue;
        if (!cell_is_hot<WP>(wsum, cy, cx, gl)) continue;
        const size_t cellg = ((size_t)n * img_pixels + (size_t)(ay * q.w + ax)) * q.G + g;
        const size_t idx = cellg * kGC + ch % kGC;
        const long long v = (long long)__ldcg(ws.acc64 + idx);
        // |v| can exceed 2^24: go through double so that the exact total is rounded once
        Elem<T>::st(grad_x + idx, __fadd_rn(Elem<T>::ld_plain(grad_x + idx), (float)((double)v * inv_d)));
        ws.acc64[idx] = 0ull;
        ws.dirty[cellg] = 0;
    }
}

#ifdef DCNV3_SCATTER_PROFILE
// debug build only (tools/scatter_phases.py): per-CTA phase clocks
__device__ long long* g_scatter_prof = nullptr;
#define PROF_MARK(k) do { if (g_scatter_prof && threadIdx.x == 0) g_scatter_prof[(size_t)blockIdx.x * 16 + (k)] = clock64(); } while (0)
#else
#define PROF_MARK(k) do { } while (0)
#endif

template <typename T, int TJ, bool BLEND, bool PER_TAP>
__global__ void __launch_bounds__(PER_TAP ? ScatterShape<TJ>::THREADS_PER_TAP : ScatterShape<TJ>::THREADS, ScatterShape<TJ>::MIN_CTAS)
bwd_scatter_kernel(const T* __restrict__ offset, const T* __restrict__ mask, const T* __restrict__ grad_out,
                   T* __restrict__ grad_x, const SideT<T> side_t, const FarWs ws, const KParams q, const BwdGeom bg) {
    constexpr int PITCH = ScatterShape<TJ>::PITCH;
    constexpr int WP = ScatterShape<TJ>::WPITCH;
    // [box_rows][PITCH][kSCell] accumulator + [box_rows + 1][WP][kSG] weight counters
    extern __shared__ __align__(128) int acc[];
    __shared__ Range s_home_h, s_home_w;
    const int acc_ints = bg.box_rows * PITCH * kSCell;
    int* wsum = acc + acc_ints;

    const ScatterCta cta = decode_scatter_cta(bg, q.n, blockIdx.x);
    const int jx = cta.jx, jy = cta.jy, chunk = cta.chunk, n = cta.n;
    const TileBox box = make_box(q, bg, jx, jy);

    constexpr bool blend = BLEND;
    pdl_launch_dependents();
#ifdef DCNV3_SCATTER_PROFILE
    __shared__ long long s_wmin, s_wmax;
    if (threadIdx.x == 0) { s_wmin = 0x7fffffffffffffffll; s_wmax = 0; }
    PROF_MARK(0);
    if (g_scatter_prof && threadIdx.x == 0) { unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); g_scatter_prof[(size_t)blockIdx.x * 16 + 8] = smid; }
#endif
    // prologue without global memory traffic (overlaps the tail of the gather kernel): home ranges, zeroed box
    if (threadIdx.x < 2) tile_ranges(q, bg, jx, jy, s_home_h, s_home_w);
    {
        int2* z = reinterpret_cast<int2*>(acc);  // both parts have an even number of ints
        for (int i = threadIdx.x; i < (acc_ints + (bg.box_rows + 1) * WP * kSG) / 2; i += blockDim.x)
            z[i] = make_int2(0, 0);
    }
    __syncthreads();
    PROF_MARK(1);
    pdl_wait();  // the gather kernel has left the image's max |grad_out| in the workspace
    PROF_MARK(2);
    const unsigned go_bits = ws.img_max[n].go_bits;
    const bool nonfinite = bits_nonfinite(go_bits);  // NaN / Inf in this image's grad_out: its grad_x is NaN
    const int eg = 30 - fixed_exponent_raw(go_bits);
    if (!nonfinite) {
        if (PER_TAP)
            scatter_walk_per_tap<T, 0, TJ, BLEND>(acc, wsum, offset, mask, grad_out, ws, q, box, n, chunk, s_home_h, s_home_w, eg);
        else
            scatter_walk_batched<T, TJ, DCNV3_SCATTER_BATCH, BLEND>(acc, wsum, offset, mask, grad_out, side_t, ws, q, box, n, chunk, s_home_h, s_home_w, eg);
    }
#ifdef DCNV3_SCATTER_PROFILE
    if ((threadIdx.x & 31) == 0) { const long long t = clock64(); atomicMin(&s_wmin, t); atomicMax(&s_wmax, t); }
#endif
    __syncthreads();
    PROF_MARK(3);
#ifdef DCNV3_SCATTER_PROFILE
    if (g_scatter_prof && threadIdx.x == 0) { g_scatter_prof[(size_t)blockIdx.x * 16 + 6] = s_wmin; g_scatter_prof[(size_t)blockIdx.x * 16 + 7] = s_wmax; }
#endif

    // ---- flush: the cells of J are written exactly once; ring cells go to the side buffer ----
    // work item = 32 consecutive 4-channel pieces of one box row (4 cells x 2 groups), dealt round-robin to the warps
    const float inv_s = ldexpf(1.0f, -(eg + kWShift - 32));  // q = value * 2^(eg + kWShift - 32)
    constexpr int QPC = kSCell / 4;  // 4-channel pieces per cell
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int segs = (box.bw * QPC + 31) >> 5;
    const int piece = lane & (QPC - 1), gl = piece >> 2;  // (QPC = 8 divides 32: a lane keeps its piece)
    const int g = chunk * kSG + gl;
    const float qnan = __int_as_float(0x7fc00000);
    bool any_hot = false;
    // Can any cell be hot at all?  A cell's bound is the sum of four counters, so it needs one above kBudget / 4.
    // Almost never: then the flush runs without the per-cell test, a warp per box row and with 32-bit indexing.
    int cmax = 0;
    if (!PER_TAP)
        for (int i = threadIdx.x; i < (bg.box_rows + 1) * WP * kSG; i += blockDim.x) cmax = max(cmax, wsum[i]);
    // (the per-tap instance keeps the round-1 kernel body: at its 96 registers every addition to the kernel -- this
    //  flush, the in-kernel redo below -- was measured to slow its walk down by 15-30 %, profiles/r02_scatter_study.md)
    if (!PER_TAP && !__syncthreads_or(cmax > kBudget / 4 || nonfinite)) {
        const int row_elems = q.w * q.G * kGC;                // grad_x elements per image row (< 2^31: tiled_applicable)
        const int npieces = box.bw * QPC;
        for (int cy = warp; cy < box.bh; cy += nwarps) {
            const int ay = box.by0 + cy;
            if (ay < 0 || ay >= q.h) continue;               // zero ring of tf.pad: gradient dropped (Pad-grad)
            const size_t grow = (size_t)n * q.h + ay;
            T* gx_row = grad_x + grow * row_elems + chunk * kSCell;
            unsigned long long* side_row = ws.acc64 + (grow * q.w * q.G + chunk * kSG) * kGC;
            unsigned char* dirty_row = ws.dirty + grow * q.w * q.G + chunk * kSG;
            const bool row_owned = (unsigned)(ay - box.uy0) < (unsigned)box.tjh;
            const int* arow = acc + cy * (PITCH * kSCell);
            for (int idx = lane; idx < npieces; idx += 32) {
                const int cx = idx >> 3, ax = box.bx0 + cx;  // (QPC = 8; the lane keeps its piece: 32 % 8 == 0)
                if (ax < 0 || ax >= q.w || g >= q.G) continue;
                const int4 v = *reinterpret_cast<const int4*>(arow + idx * 4);
                if (row_owned && (unsigned)(ax - box.ux0) < (unsigned)box.tjw) {
                    const int e = ax * (q.G * kGC) + piece * 4;
                    float f0 = (float)v.x * inv_s, f1 = (float)v.y * inv_s, f2_ = (float)v.z * inv_s, f3 = (float)v.w * inv_s;
                    if (blend) {  // the blend's direct path: d x_proj += grad_out * s at the pixel itself
                        const float cs = Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + (PER_TAP ? side_entry(q, grow * q.w + ax, g) : (grow * q.w + ax) * q.G + g));
                        const float4 g4 = Elem<T>::ld4(grad_out + grow * row_elems + chunk * kSCell + e);
                        f0 = __fadd_rn(f0, __fmul_rn(g4.x, cs)); f1 = __fadd_rn(f1, __fmul_rn(g4.y, cs));
                        f2_ = __fadd_rn(f2_, __fmul_rn(g4.z, cs)); f3 = __fadd_rn(f3, __fmul_rn(g4.w, cs));
                    }
                    if (sizeof(T) == 4) {
                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(gx_row) + e) = make_float4(f0, f1, f2_, f3);
                    } else {
                        uint2 r2;
                        r2.x = pack_bf16x2(f0, f1);
                        r2.y = pack_bf16x2(f2_, f3);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(gx_row) + e) = r2;
                    }
                } else if ((v.x | v.y | v.z | v.w) != 0) {
                    unsigned long long* dst = side_row + (ax * q.G + gl) * kGC + (piece & 3) * 4;
                    if (v.x) atomicAdd(dst + 0, (unsigned long long)(long long)v.x);
                    if (v.y) atomicAdd(dst + 1, (unsigned long long)(long long)v.y);
                    if (v.z) atomicAdd(dst + 2, (unsigned long long)(long long)v.z);
                    if (v.w) atomicAdd(dst + 3, (unsigned long long)(long long)v.w);
                    dirty_row[ax * q.G + gl] = 1;
                }
            }
        }
        __syncthreads();
        PROF_MARK(4);
        if (bg.tiles_x * bg.tiles_y == 1) finalize_workspace(ws, q.n);  // whole-image tile: last kernel of the call
        return;
    }
    for (int item = warp; item < box.bh * segs; item += nwarps) {
        const int cy = item / segs, cx = ((item - cy * segs) << 2) + (lane >> 3);
        const int ax = box.bx0 + cx, ay = box.by0 + cy;
        // beyond the box row; zero ring: gradient dropped (Pad-grad); phantom group
        if (cx >= box.bw || ax < 0 || ax >= q.w || ay < 0 || ay >= q.h || g >= q.G) continue;
        // a hot (cell, group) may have wrapped: it is skipped here and recomputed by redo_hot_kernel
        const bool hot = cell_is_hot<WP>(wsum, cy, cx, gl);
        any_hot |= hot;
        const int4 v = *reinterpret_cast<const int4*>(acc + (cy * PITCH + cx) * kSCell + piece * 4);
        const size_t gpix = ((size_t)n * q.h + ay) * q.w + ax;
        if ((unsigned)(ax - box.ux0) < (unsigned)box.tjw && (unsigned)(ay - box.uy0) < (unsigned)box.tjh) {
            const size_t gidx = gpix * ((size_t)q.G * kGC) + (size_t)(chunk * kSCell + piece * 4);
            const bool drop = hot || nonfinite;
            const float fill = nonfinite ? qnan : 0.f;
            float f0 = drop ? fill : (float)v.x * inv_s, f1 = drop ? fill : (float)v.y * inv_s;
            float f2_ = drop ? fill : (float)v.z * inv_s, f3 = drop ? fill : (float)v.w * inv_s;
            if (blend) {  // the blend's direct path (hot cells: the redo / merge kernels add to this)
                const float cs = Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + (PER_TAP ? side_entry(q, gpix, g) : gpix * q.G + g));
                const float4 g4 = Elem<T>::ld4(grad_out + gidx);
                f0 = __fadd_rn(f0, __fmul_rn(g4.x, cs)); f1 = __fadd_rn(f1, __fmul_rn(g4.y, cs));
                f2_ = __fadd_rn(f2_, __fmul_rn(g4.z, cs)); f3 = __fadd_rn(f3, __fmul_rn(g4.w, cs));
            }
            if (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(grad_x) + gidx) = make_float4(f0, f1, f2_, f3);
            } else {
                uint2 r2;
                r2.x = pack_bf16x2(f0, f1);
                r2.y = pack_bf16x2(f2_, f3);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(grad_x) + gidx) = r2;
            }
        } else if (!hot && (v.x | v.y | v.z | v.w) != 0) {
            const size_t cellg = gpix * q.G + g;
            unsigned long long* dst = ws.acc64 + cellg * kGC + (piece * 4) % kGC;
            if (v.x) atomicAdd(dst + 0, (unsigned long long)(long long)v.x);
            if (v.y) atomicAdd(dst + 1, (unsigned long long)(long long)v.y);
            if (v.z) atomicAdd(dst + 2, (unsigned long long)(long long)v.z);
            if (v.w) atomicAdd(dst + 3, (unsigned long long)(long long)v.w);
            ws.dirty[cellg] = 1;
        }
    }
    if (!PER_TAP && bg.tiles_x * bg.tiles_y == 1) {
        // whole-image tile: hot cells are redone here and this kernel is the last one of the call
        if (__syncthreads_or(any_hot) && !nonfinite)
            redo_own_tile<T, TJ, false>(wsum, offset, mask, grad_out, grad_x, ws, q, box, n, chunk, s_home_h, s_home_w, eg);
        finalize_workspace(ws, q.n);
    } else if (any_hot) {
        ws.redo[blockIdx.x] = 1;
        ws.hd->any_redo = 1u;  // (benign race: every writer stores the same value)
    }
#ifdef DCNV3_SCATTER_PROFILE
    __syncthreads();
    PROF_MARK(4);
#endif
}
#ifdef DCNV3_SCATTER_PROFILE
extern "C" int dcnv3_debug_scatter_profile(long long* buf) {
    return (int)cudaMemcpyToSymbol(g_scatter_prof, &buf, sizeof(buf));
}
#endif

// Exact recomputation of the hot cells of the flagged boxes (whole-image tiles of the batched scatter instance do it
// inside the scatter kernel, redo_own_tile, and need no launch of this kernel): pass 1 rebuilds the weight counters,
// pass 2 adds the landings on hot cells to the 64-bit side buffer, which merge_far_kernel then folds into grad_x (or,
// for a whole-image tile, this CTA itself).
// Launched with a handful of CTAs: unless some scatter CTA raised any_redo -- which only adversarial inputs make it
// do -- they only read that word and leave; otherwise they share the flagged tiles among themselves.
template <typename T, int TJ>
__global__ void __launch_bounds__(256)
redo_hot_kernel(const T* __restrict__ offset, const T* __restrict__ mask, const T* __restrict__ grad_out,
                T* __restrict__ grad_x, const FarWs ws, const KParams q, const BwdGeom bg, const int n_tiles, const int is_last) {
    constexpr int WP = ScatterShape<TJ>::WPITCH;
    extern __shared__ __align__(16) int wsum[];  // [box_rows + 1][WP][kSG]
    __shared__ Range s_home_h, s_home_w;
    pdl_launch_dependents();
    pdl_wait();
    for (int tile = blockIdx.x; ws.hd->any_redo != 0u && tile < n_tiles; tile += gridDim.x) {  // (uniform over the grid)
        if (ws.redo[tile] == 0) continue;  // (uniform per CTA)
        const ScatterCta cta = decode_scatter_cta(bg, q.n, tile);
        const int jx = cta.jx, jy = cta.jy, chunk = cta.chunk, n = cta.n;
        const TileBox box = make_box(q, bg, jx, jy);
        __syncthreads();  // the previous tile's counters and ranges are no longer read
        if (threadIdx.x < 2) tile_ranges(q, bg, jx, jy, s_home_h, s_home_w);
        for (int i = threadIdx.x; i < (bg.box_rows + 1) * WP * kSG; i += blockDim.x) wsum[i] = 0;
        __syncthreads();
        const int eg = 30 - fixed_exponent_raw(ws.img_max[n].go_bits);
        redo_walk<T, 1, TJ>(wsum, offset, mask, grad_out, ws, q, box, n, chunk, s_home_h, s_home_w, eg);
        __syncthreads();
        if (bg.tiles_x * bg.tiles_y == 1)  // whole-image tile of the per-tap instance: pass 2 + conversion (no merge launch)
            redo_own_tile<T, TJ>(wsum, offset, mask, grad_out, grad_x, ws, q, box, n, chunk, s_home_h, s_home_w, eg);
        else
            redo_walk<T, 2, TJ>(wsum, offset, mask, grad_out, ws, q, box, n, chunk, s_home_h, s_home_w, eg);
        __syncthreads();
        if (threadIdx.x == 0) ws.redo[tile] = 0;
    }
    if (is_last) finalize_workspace(ws, q.n);
}

// grad_x += side buffer for the (pixel, group)s flagged in the dirty map; leaves the side buffer and
// the map zeroed for the next call.  A warp scans 32 map bytes; the flagged entries are then taken eight
// at a time, four lanes per entry and four channels per lane (coalesced 128-byte runs of the side buffer).
template <typename T>
__global__ void __launch_bounds__(256)
merge_far_kernel(T* __restrict__ grad_x, const FarWs ws, const KParams q, size_t count) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) - lane;
    const size_t e = base + lane;
    const unsigned char flag = e < count ? ws.dirty[e] : (unsigned char)0;  // map is padded to 256 bytes
    const unsigned m = __ballot_sync(0xffffffffu, flag != 0);
    if (m != 0) {
        const size_t per_image = (size_t)q.h * q.w * q.G;  // (pixel, group) entries of one image
        const int nset = __popc(m);
        for (int k0 = 0; k0 < nset; k0 += 8) {
            const int k = k0 + (lane >> 2);
            if (k < nset) {
                const size_t ent = base + __fns(m, 0, k + 1);
                const size_t idx = ent * kGC + (lane & 3) * 4;
                // the image's own fixed-point scale (a warp's 32 entries may straddle two images)
                const double inv_s = ldexp(1.0, -(30 - fixed_exponent_raw(ws.img_max[ent / per_image].go_bits) + kWShift - 32));
                longlong4 v;
                const longlong2 v01 = *reinterpret_cast<const longlong2*>(ws.acc64 + idx);
                const longlong2 v23 = *reinterpret_cast<const longlong2*>(ws.acc64 + idx + 2);
                v.x = v01.x; v.y = v01.y; v.z = v23.x; v.w = v23.y;
                // |v| can exceed 2^24: go through double so that the exact total is rounded once
                const float a0 = (float)((double)v.x * inv_s), a1 = (float)((double)v.y * inv_s);
                const float a2 = (float)((double)v.z * inv_s), a3 = (float)((double)v.w * inv_s);
                float4 gx = Elem<T>::ld4_plain(grad_x + idx);
                gx.x += a0; gx.y += a1; gx.z += a2; gx.w += a3;
                Elem<T>::st4(grad_x + idx, gx);
                *reinterpret_cast<ulonglong2*>(ws.acc64 + idx) = make_ulonglong2(0ull, 0ull);
                *reinterpret_cast<ulonglong2*>(ws.acc64 + idx + 2) = make_ulonglong2(0ull, 0ull);
            }
        }
        if (flag != 0) ws.dirty[e] = 0;
    }
    finalize_workspace(ws, q.n);  // always the last kernel of a call when it is launched
}

// ---- host side -----------------------------------------------------------------------------------
static BwdGeom make_bwd_geom(const KParams& q) {
    BwdGeom bg;
    // 32x32 tiles keep the ring small; images that fit a 16x16 tile use that (smaller box, more CTAs per SM)
    bg.tj = (q.w > 16 || q.h > 16) ? 32 : 16;
    bg.tj_log2 = bg.tj == 32 ? 5 : 4;
    bg.tiles_x = (q.w + bg.tj - 1) / bg.tj;
    bg.tiles_y = (q.h + bg.tj - 1) / bg.tj;
    bg.chunks = (q.G + kSG - 1) / kSG;
    // A tap lands within [nominal - ceil(D), nominal + floor(D) + 2] of its pixel's nominal cell, D = (1 +
    // |offset|) * scale * (dim-2)/dim.  The ring serves |offset| <= reach (reference units) from shared
    // memory; reach starts at 3 and shrinks until the box fits the compile-time pitch.
    const float r = fmaxf(q.wm2_f / q.win_f, q.hm2_f / q.hin_f) * q.scale;
    for (float reach = 3.0f;; reach *= 0.75f) {
        const float d = (1.0f + reach) * r;
        bg.ring_lo = max((int)ceilf(d), 1);
        bg.ring_hi = (int)floorf(d) + 2;
        if ((bg.ring_lo <= kRingLo && bg.ring_hi <= kRingHi) || reach < 1e-3f) break;
    }
    bg.narrow = 0;
    {   // did the loop above have to shrink the reach?
        const float d3 = 4.0f * r;
        bg.narrow = !((int)ceilf(d3) <= kRingLo && (int)floorf(d3) + 2 <= kRingHi);
    }
    bg.ring_lo = min(bg.ring_lo, kRingLo);
    bg.ring_hi = min(bg.ring_hi, kRingHi);
    bg.ordered = 0;
    if (bg.tiles_x * bg.tiles_y > 1 && bg.tiles_x * bg.tiles_y <= 64) {
        // home pixels per tile along each axis: output rows h whose nominal input column falls into tile jx (the device's
        // nominal_ux / nominal_uy, same integer arithmetic), likewise output columns w and tile jy
        int cx[64] = {0}, cy[64] = {0};
        for (int h = 0; h < q.ho; ++h) {
            const int u = (int)((unsigned)((2 * h + 3) * (q.win - 2)) / (unsigned)(2 * q.hin)) - q.pw;
            cx[std::min(std::max(u >> bg.tj_log2, 0), bg.tiles_x - 1)]++;
        }
        for (int w = 0; w < q.wo; ++w) {
            const int u = (int)((unsigned)((2 * w + 3) * (q.hin - 2)) / (unsigned)(2 * q.win)) - q.ph;
            cy[std::min(std::max(u >> bg.tj_log2, 0), bg.tiles_y - 1)]++;
        }
        int idx[64];
        const int nt = bg.tiles_x * bg.tiles_y;
        for (int i = 0; i < nt; ++i) idx[i] = i;
        std::stable_sort(idx, idx + nt, [&](int a, int b) {
            return cx[a % bg.tiles_x] * cy[a / bg.tiles_x] > cx[b % bg.tiles_x] * cy[b / bg.tiles_x];
        });
        for (int i = 0; i < nt; ++i) {
            bg.order_x[i] = (unsigned char)(idx[i] % bg.tiles_x);
            bg.order_y[i] = (unsigned char)(idx[i] / bg.tiles_x);
        }
        bg.ordered = 1;
    }
    bg.box_rows = 0;
    for (int jy = 0; jy < bg.tiles_y; ++jy) {
        const int uy0 = jy * bg.tj, tjh = min(bg.tj, q.h - uy0);
        const int by0 = max(uy0 - bg.ring_lo, -1);
        bg.box_rows = max(bg.box_rows, min(uy0 + tjh + bg.ring_hi, q.h + 1) - by0);
    }
    return bg;
}

static size_t flag_bytes(size_t count) { return (count * sizeof(int) + 255) / 256 * 256; }
static size_t dirty_bytes(const KParams& q) { return ((size_t)q.n * q.h * q.w * q.G + 255) / 256 * 256; }

size_t bwd_tiled_scratch_bytes(const KParams& q, int dtype) { return side_t_bytes(tiled_view(q), dtype); }

size_t bwd_tiled_workspace_bytes(const KParams& q_in) {
    const KParams q = tiled_view(q_in);
    const size_t tiles = (size_t)q.n * ((q.w + 15) / 16) * ((q.h + 15) / 16);  // upper bound (16x16 tiles)
    const size_t chunks = (size_t)(q.G + 1) / 2;
    return sizeof(WsHeader) + img_max_bytes(q.n) + dirty_bytes(q) + flag_bytes(tiles * chunks) +
           sizeof(long long) * (size_t)q.n * q.h * q.w * q.G * q.gc;
}

// accumulator [rows][PITCH][kSCell] + weight counters [rows + 1][WPITCH][kSG]
template <int TJ>
static size_t scatter_smem_bytes(int rows) {
    using S = ScatterShape<TJ>;
    return ((size_t)rows * S::PITCH * kSCell + (size_t)(rows + 1) * S::WPITCH * kSG) * sizeof(int);
}

template <typename T, int TJ>
static cudaError_t launch_scatter(const T* offset, const T* mask, const T* grad_out, T* grad_x, const SideT<T>& side_t,
                                  const FarWs& ws, const KParams& q, const BwdGeom& bg, unsigned grid, bool redo_is_last,
                                  cudaStream_t st) {
    using S = ScatterShape<TJ>;
    const size_t smem = scatter_smem_bytes<TJ>(bg.box_rows);
    // (two instances: the centre-feature-scale blend costs the plain op nothing)
    // (instances: with / without the centre-feature-scale blend, so that it costs the plain op nothing; batched walk,
    //  or the per-tap walk where the ring is narrow -- the blend is read at run time there)
    //  bf16 always takes the per-tap walk: its per-lane loads touch half as many lines, there is no transposed copy to
    //  read, and the batched walk measured 1 % slower on the InternImage-T step and 17 % slower on InternImage-L)
    const bool blend = q.cfs != nullptr;
    void (*kernel)(const T*, const T*, const T*, T*, SideT<T>, FarWs, KParams, BwdGeom) =
        blend ? bwd_scatter_kernel<T, TJ, true, true> : bwd_scatter_kernel<T, TJ, false, true>;
    bool per_tap = true;
    if constexpr (sizeof(T) == 4) {
        per_tap = bg.narrow != 0 || q.gsh != 0;  // (half groups: only the per-tap walk and the redo kernel index them)
        if (!per_tap) kernel = blend ? bwd_scatter_kernel<T, TJ, true, false> : bwd_scatter_kernel<T, TJ, false, false>;
    }
    const unsigned threads = per_tap ? S::THREADS_PER_TAP : S::THREADS;
    cudaError_t e = ensure_max_smem((const void*)kernel, (int)scatter_smem_bytes<TJ>(S::PITCH));
    if (e != cudaSuccess) return e;
    KernelTiming& kt = kernel_timing();
    e = launch_pdl(kernel, grid, threads, smem, st, offset, mask, grad_out, grad_x, side_t, ws, q, bg);
    if (e != cudaSuccess) return e;
    if (kt.enabled) cudaEventRecord(kt.ev[2], st);
    // whole-image tiles of the batched instance: the scatter kernel redoes its own hot cells and finalises the workspace
    if (redo_is_last && !per_tap) return cudaSuccess;
    count_launch(1);
    const unsigned redo_grid = grid < 32u ? grid : 32u;  // normally they only read one word and leave
    return launch_pdl(redo_hot_kernel<T, TJ>, redo_grid, 256, (size_t)(bg.box_rows + 1) * S::WPITCH * kSG * sizeof(int), st,
                      offset, mask, grad_out, grad_x, ws, q, bg, (int)grid, (int)redo_is_last);
}

template <typename T, bool STAGED>
static cudaError_t launch_gather_variant(const CUtensorMap& map, const CUtensorMap& offmap, const CUtensorMap& goffmap,
                                         const void* x, const void* offset, const void* mask, const void* grad_out,
                                         void* grad_offset, void* grad_mask, ImgMax* img_max, const SideT<T>& side_t,
                                         const KParams& q, const TileGeom& tg, cudaStream_t st) {
    auto kernel = q.cfs != nullptr ? bwd_gather_kernel<T, STAGED, true> : bwd_gather_kernel<T, STAGED, false>;
    cudaError_t e = ensure_max_smem((const void*)kernel, kMaxBoxBytes + kTiledWarps * kGatherStageBytes<T>);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((size_t)q.n * tg.chunks * tg.tiles_h * tg.tiles_w);
    // (also leaves the per-image max |grad_out| in the workspace: the fixed-point scale of the scatter kernel)
    return launch_pdl(kernel, grid, kTiledWarps * 32,
                      (size_t)tg.bw * tg.bh * kCellBytes + kTiledWarps * kGatherStageBytes<T>, st, map, offmap, goffmap,
                      (const T*)x, (const T*)offset, (const T*)mask, (const T*)grad_out, (T*)grad_offset, (T*)grad_mask,
                      img_max, side_t, q, tg);
}

template <typename T>
static cudaError_t launch_bwd_tiled_t(const void* x, const void* offset, const void* mask, const void* grad_out,
                                      void* grad_x, void* grad_offset, void* grad_mask, void* wsp, void* scratch,
                                      const KParams& q, int dtype, bool ws_clean, cudaStream_t st) {
    const BwdGeom bg = make_bwd_geom(q);
    SideT<T> side_t;
    {
        const size_t entries = (size_t)q.n * ((q.G + 1) / 2) * kTaps * q.h * q.w * 2;
        // (no copy where the per-tap walk runs: it reads the reference-layout tensors)
        side_t.off = (kUseSideT<T> && !bg.narrow && !q.gsh) ? (typename SideT<T>::Pair*)scratch : nullptr;
        side_t.msk = (T*)((char*)scratch + (entries * 2 * sizeof(T) + 255) / 256 * 256);
        side_t.cs = (q.G + 1) / 2;
        side_t.hw = q.h * q.w;
    }
    const size_t tiles = (size_t)q.n * bg.tiles_x * bg.tiles_y;
    const size_t tiles_ub = (size_t)q.n * ((q.w + 15) / 16) * ((q.h + 15) / 16);
    const size_t chunks_ub = (size_t)(q.G + 1) / 2;
    FarWs ws;
    char* base = (char*)wsp;
    ws.hd = (WsHeader*)base; base += sizeof(WsHeader);
    ws.img_max = (ImgMax*)base; base += img_max_bytes(q.n);
    ws.dirty = (unsigned char*)base; base += dirty_bytes(q);
    ws.redo = (int*)base; base += flag_bytes(tiles_ub * chunks_ub);
    ws.acc64 = (unsigned long long*)base;
    // The whole workspace must be zero on entry.  Every call leaves it zero on exit (the last kernel's last CTA
    // clears the prefix, the redo / merge kernels the rest), so a caller that keeps it says so and no memset
    // node interrupts the chain of programmatically dependent launches.
    cudaError_t e;
    if (!ws_clean && (e = cudaMemsetAsync(wsp, 0, bwd_tiled_workspace_bytes(q), st)) != cudaSuccess) return e;

    // ---- grad_offset / grad_mask ----
    const int max_cells = kMaxBoxBytes / kCellBytes;
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, max_cells);
    if (tg.bw * tg.bh > max_cells) return cudaErrorInvalidConfiguration;
    CUtensorMap map, offmap, goffmap;
    if (!make_x_tensor_map(&map, x, q, dtype, tg.bw, tg.bh)) return cudaErrorNotSupported;
    const bool staged = side_stageable(q, dtype);
    if (staged && !(make_side_tensor_map(&offmap, offset, q, dtype, 18) &&
                    make_side_tensor_map(&goffmap, grad_offset, q, dtype, 18)))
        return cudaErrorNotSupported;
    if (!staged) offmap = goffmap = map;  // unused by the kernel variant
    KernelTiming& kt = kernel_timing();
    if (kt.enabled) cudaEventRecord(kt.ev[0], st);
    e = staged ? launch_gather_variant<T, true>(map, offmap, goffmap, x, offset, mask, grad_out, grad_offset, grad_mask,
                                                ws.img_max, side_t, q, tg, st)
               : launch_gather_variant<T, false>(map, offmap, goffmap, x, offset, mask, grad_out, grad_offset, grad_mask,
                                                 ws.img_max, side_t, q, tg, st);
    if (e != cudaSuccess) return e;

    // ---- grad_x ----
    if (kt.enabled) cudaEventRecord(kt.ev[1], st);
    const unsigned grid_b = (unsigned)(tiles * bg.chunks);
    // a single tile owns every cell of its image: no ring, no far landings, nothing to merge
    const bool merge = bg.tiles_x * bg.tiles_y > 1;
    e = bg.tj == 32 ? launch_scatter<T, 32>((const T*)offset, (const T*)mask, (const T*)grad_out, (T*)grad_x, side_t, ws, q,
                                            bg, grid_b, !merge, st)
                    : launch_scatter<T, 16>((const T*)offset, (const T*)mask, (const T*)grad_out, (T*)grad_x, side_t, ws, q,
                                            bg, grid_b, !merge, st);
    if (e != cudaSuccess) return e;
    if (kt.enabled) cudaEventRecord(kt.ev[3], st);
    if (merge) {
        const size_t npg = (size_t)q.n * q.h * q.w * q.G;
        e = launch_pdl(merge_far_kernel<T>, (unsigned)((npg + 255) / 256), 256, 0, st, (T*)grad_x, ws, q, npg);
        if (e != cudaSuccess) return e;
    }
    if (kt.enabled) cudaEventRecord(kt.ev[4], st);
    count_launch(merge ? 3 : 2);  // gather, scatter (+ merge); the redo launch counts itself
    return cudaGetLastError();
}

void bwd_tiled_plan(const KParams& q_in, int dtype, int out[16]) {
    const KParams q = tiled_view(q_in);
    gather_tiled_plan(q, dtype, kTiledWarps * (dtype == DCNV3_F32 ? kGatherStageBytes<float> : kGatherStageBytes<__nv_bfloat16>), out);
    const BwdGeom bg = make_bwd_geom(q);
    out[8] = bg.tj; out[9] = bg.ring_lo; out[10] = bg.ring_hi; out[11] = bg.box_rows;
    out[12] = (int)((long long)q.n * bg.tiles_x * bg.tiles_y * bg.chunks);
    out[13] = (int)(bg.tj == 32 ? scatter_smem_bytes<32>(bg.box_rows) : scatter_smem_bytes<16>(bg.box_rows));
    out[14] = bg.tj == 32 ? ((bg.narrow || dtype != DCNV3_F32) ? ScatterShape<32>::THREADS_PER_TAP : ScatterShape<32>::THREADS)
                          : ScatterShape<16>::THREADS;
    out[15] = bg.tiles_x * bg.tiles_y > 1;
}

cudaError_t launch_bwd_tiled(const void* x, const void* offset, const void* mask, const void* grad_out,
                             void* grad_x, void* grad_offset, void* grad_mask, void* ws, void* scratch, const KParams& q_in,
                             int dtype, bool ws_clean, cudaStream_t st) {
    const KParams q = tiled_view(q_in);
    return dtype == DCNV3_F32
               ? launch_bwd_tiled_t<float>(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, scratch, q, dtype, ws_clean, st)
               : launch_bwd_tiled_t<__nv_bfloat16>(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, scratch, q, dtype, ws_clean, st);
}

}  // namespace dcnv3
