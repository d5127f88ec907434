// Tiled DCNv3 forward (sm_100a).  One CTA = one output tile x one group chunk of one image.
//   1. an elected thread issues ONE TMA box load (cp.async.bulk.tensor.4d) of the chunk's input slab
//      covering the tile's predictable footprint (+halo) into shared memory; out-of-image cells are
//      zero-filled by the TMA unit, i.e. the zero ring of op.py:46 is never materialised;
//   2. every lane owns one (pixel, group): 18 offsets + 9 mask values in registers, 9 taps x 4
//      corners gathered as conflict-free LDS.128 (see dcnv3_tiled.cuh), fp32 accumulation in the
//      reference's tap order (utils.py:195-206);
//   3. taps whose 2x2 patch leaves the staged box (large learned offsets) fall back to L2/global
//      loads -- correctness never depends on the halo, only speed does.
// Algorithmic HBM bytes per (pixel, group): x 16 + out 16 + offset 18 + mask 9 elements.
#include "dcnv3_kernels.h"
#include "dcnv3_tiled.cuh"

namespace dcnv3 {

template <typename T>
__device__ __forceinline__ const T* global_slab(const T* x, const KParams& q, int n, int yp, int xp, int g) {
    const int y = yp - q.ph, xx = xp - q.pw;
    if (y < 0 || y >= q.h || xx < 0 || xx >= q.w) return nullptr;
    return x + ((((size_t)n * q.h + y) * q.w + xx) * q.G + g) * kGC;
}

template <typename T>
__global__ void __launch_bounds__(256, 2)
fwd_tiled_kernel(const __grid_constant__ CUtensorMap xmap, const T* __restrict__ x,
                 const T* __restrict__ offset, const T* __restrict__ mask, T* __restrict__ out,
                 const KParams q, const TileGeom tg) {
    using C = Chunk<T>;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    // ---- which tile ----
    int b = blockIdx.x;
    const int tx = b % tg.tiles_w; b /= tg.tiles_w;
    const int ty = b % tg.tiles_h; b /= tg.tiles_h;
    const int chunk = b % tg.chunks;
    const int n = b / tg.chunks;
    const int h0 = ty * tg.th, w0 = tx * tg.tw;
    const int th = min(tg.th, q.ho - h0), tw = min(tg.tw, q.wo - w0);
    // box origin in padded coordinates (output rows h walk along x, columns w along y)
    // (kept inside the padded image: cells beyond it can only belong to dead taps)
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

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g_l = lane % C::GQ, px_l = lane / C::GQ;
    const bool real_g = chunk * C::GQ + g_l < q.G;          // false for the phantom groups of a trailing chunk
    const int g = min(chunk * C::GQ + g_l, q.G - 1);        // (phantom lanes shadow the last group, never store)
    const int rot = Slab<T>::rot_of(px_l);
    const unsigned char* sbase = smem + g_l * (kGC * (int)sizeof(T));
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const int npix = th * tw;
    bool waited = false;
    (void)th;

    for (int p0 = warp * C::PXW; p0 < npix; p0 += (blockDim.x >> 5) * C::PXW) {
        const bool valid = real_g && p0 + px_l < npix;
        const int pix = min(p0 + px_l, npix - 1);  // idle lanes shadow the last pixel (loads stay in bounds)
        const int h = h0 + pix / tw, w = w0 + pix % tw;
        const size_t pg = (((size_t)n * q.ho + h) * q.wo + w) * q.G + g;
        const T* offp = offset + pg * 18;
        const T* mskp = mask + pg * 9;
        float mx = 0.f, inv_sum = 1.f;
        if (logits) softmax_stats9<T>(mskp, mx, inv_sum);
        float ref0, ref1;
        ref_point(q, h, w, ref0, ref1);
        f2 acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0ull;
        float ox, oy, ml, ox2, oy2, ml2;
        load_tap_inputs<T>(offp, mskp, 0, ox, oy, ml);
        load_tap_inputs<T>(offp, mskp, 1, ox2, oy2, ml2);
        if (!waited) {  // the box is needed from here on
            mbar_wait(&bar, 0);
            waited = true;
        }
#pragma unroll 1
        for (int p = 0; p < kTaps; ++p) {
            const float cx = ox, cy = oy, cm = ml;
            ox = ox2; oy = oy2; ml = ml2;
            if (p + 2 < kTaps) load_tap_inputs<T>(offp, mskp, p + 2, ox2, oy2, ml2);  // prefetch two taps ahead
            const Tap t = make_tap(q, ref0, ref1, p, cx, cy);
            const int bx = t.x0 - cx0, by = t.y0 - cy0;
            const bool inbox = bx >= 0 && bx + 1 < tg.bw && by >= 0 && by + 1 < tg.bh;
            const float mm = t.alive ? (logits ? expf(cm - mx) * inv_sum : cm) : 0.f;
            const float wa = t.dx1 * t.dy1 * mm, wb = t.dx1 * t.dy0 * mm;   // (y0,x0) (y1,x0)
            const float wc = t.dx0 * t.dy1 * mm, wd = t.dx0 * t.dy0 * mm;   // (y0,x1) (y1,x1)
            if (__builtin_expect(t.alive && !inbox, 0)) {
                // rare: patch outside the staged box -> straight from global memory
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const T* src = global_slab(x, q, n, t.y0 + (k & 1), t.x0 + (k >> 1), g);
                    if (src == nullptr) continue;
                    const float wk = (k & 1) ? ((k >> 1) ? wd : wb) : ((k >> 1) ? wc : wa);
#pragma unroll
                    for (int pc = 0; pc < C::NPIECE; ++pc) {
                        f2 v[C::PAIRS];
                        load_piece<T>(src + Slab<T>::chan_of(pc, rot), v);
#pragma unroll
                        for (int j = 0; j < C::PAIRS; ++j) ffma2s(acc[pc * C::PAIRS + j], v[j], wk);
                    }
                }
            } else {
                // dead taps carry zero weights and read cell 0
                const unsigned char* a = sbase + (size_t)(t.alive ? by * tg.bw + bx : 0) * kCellBytes;
                f2 va[8], vb[8];
                Slab<T>::load(a, rot, va);
                Slab<T>::load(a + (size_t)tg.bw * kCellBytes, rot, vb);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2s(acc[c], va[c], wa);
                Slab<T>::load(a + kCellBytes, rot, va);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2s(acc[c], vb[c], wb);
                Slab<T>::load(a + (size_t)(tg.bw + 1) * kCellBytes, rot, vb);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2s(acc[c], va[c], wc);
#pragma unroll
                for (int c = 0; c < 8; ++c) ffma2s(acc[c], vb[c], wd);
            }
        }
        if (valid) {
            T* dst = out + pg * kGC;
#pragma unroll
            for (int pc = 0; pc < C::NPIECE; ++pc)
                store_piece<T>(dst + Slab<T>::chan_of(pc, rot), acc + pc * C::PAIRS);
        }
    }
    if (!waited) mbar_wait(&bar, 0);  // never leave with a TMA in flight
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool make_x_tensor_map(CUtensorMap* map, const void* x, const KParams& q, int dtype, int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t es = dtype == DCNV3_F32 ? 4 : 2;
    const int gq = dtype == DCNV3_F32 ? 2 : 4;
    const cuuint64_t C = (cuuint64_t)q.G * q.gc;
    const cuuint64_t dims[4] = {C, (cuuint64_t)q.w, (cuuint64_t)q.h, (cuuint64_t)q.n};
    const cuuint64_t strides[3] = {C * es, (cuuint64_t)q.w * C * es, (cuuint64_t)q.h * q.w * C * es};
    const cuuint32_t box[4] = {(cuuint32_t)(gq * kGC), (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(map, dtype == DCNV3_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Box geometry: the nominal footprint of a th x tw output tile spans (th-1)*ax columns and (tw-1)*ay
// rows (ax = (W_in-2)/H_in, ay = (H_in-2)/W_in); taps add (1 + |offset|)*scale*(dim-2)/dim on each side
// and the bilinear patch one more cell.  `reach` is the |offset| (reference units) served from shared
// memory; beyond it the global fallback takes over.
static TileGeom geom_at(const KParams& q, int dtype, int th, int tw, float reach) {
    TileGeom tg;
    tg.th = min(th, q.ho);
    tg.tw = min(tw, q.wo);
    tg.tiles_h = (q.ho + tg.th - 1) / tg.th;
    tg.tiles_w = (q.wo + tg.tw - 1) / tg.tw;
    const int gq = dtype == DCNV3_F32 ? 2 : 4;
    tg.chunks = (q.G + gq - 1) / gq;  // a trailing partial chunk reads zero-filled phantom groups (TMA OOB)
    const float ax = q.wm2_f / q.hin_f, ay = q.hm2_f / q.win_f;
    const float rx = q.wm2_f / q.win_f, ry = q.hm2_f / q.hin_f;
    tg.halo_x = (int)ceilf((1.0f + reach) * fabsf(q.scale) * rx);
    tg.halo_y = (int)ceilf((1.0f + reach) * fabsf(q.scale) * ry);
    tg.bw = (int)ceilf((tg.th - 1) * ax) + 2 * tg.halo_x + 2;
    tg.bh = (int)ceilf((tg.tw - 1) * ay) + 2 * tg.halo_y + 2;
    // never stage more than the padded image itself
    tg.bw = min(tg.bw, min(q.win, 256));
    tg.bh = min(tg.bh, min(q.hin, 256));
    return tg;
}

// 16 x 16 output tiles with the largest reach <= `reach` whose box fits.  The reference pairs output rows
// with input columns, so on strongly non-square images a 16-row tile spans 16*W/H columns,
// and a large offset_scale widens every halo: when the square tile cannot keep a reach of 1 offset unit,
// tile shape (rows down to 1, columns up to 128) and reach are chosen together by a simple cost model --
// pixels per staged cell, discounted by the share of taps (offsets ~ N(0,1)) that would miss the box and
// take the ~10x slower global path.  If nothing fits, bw * bh > max_cells in the result and the caller
// uses the generic kernels.
TileGeom make_geom(const KParams& q, int dtype, int th, int tw, float reach, int max_cells) {
    TileGeom tg;
    for (float r = reach; r >= 1.0f; r *= 0.75f) {
        tg = geom_at(q, dtype, th, tw, r);
        if (tg.bw * tg.bh <= max_cells) return tg;
    }
    TileGeom best = geom_at(q, dtype, th, tw, reach);
    float best_score = -1.0f;
    for (int h = 16; h >= 1; h >>= 1)
        for (int w = 16; w <= 128; w <<= 1)
            for (float r = reach; r >= 0.05f; r *= 0.75f) {
                const TileGeom c = geom_at(q, dtype, h, w, r);
                if (c.bw * c.bh > max_cells) continue;
                const float miss = 1.0f - erff(r * 0.70710678f);
                const float score = (float)(c.th * c.tw) / (float)(c.bw * c.bh) / (1.0f + 10.0f * miss);
                if (score > best_score) {
                    best_score = score;
                    best = c;
                }
                break;  // smaller reaches of this shape only score lower
            }
    return best;
}

// Tiled kernels serve the InternImage configuration only -- and only images whose tile boxes fit shared
// memory (everything but extreme aspect ratios at large offset_scale); the rest runs the generic kernels.
bool tiled_applicable(const KParams& q, int dtype) {
    // any group count: the last chunk may be partly empty (phantom groups are masked)
    if (!(q.P == 9 && q.kh == 3 && q.sh == 1 && q.sw == 1 && q.dh == 1 && q.dw == 1 && q.gc == kGC && q.ho == q.h &&
          q.wo == q.w && q.ph == 1 && q.pw == 1 && q.scale > 0.f && q.scale <= 16.f && q.h <= 16384 && q.w <= 16384))
        return false;
    const int fwd_cells = kFwdBoxBytes / kCellBytes, bwd_cells = kMaxBoxBytes / kCellBytes;
    const TileGeom a = make_geom(q, dtype, 16, 16, 3.0f, fwd_cells), b = make_geom(q, dtype, 16, 16, 3.0f, bwd_cells);
    return a.bw * a.bh <= fwd_cells && b.bw * b.bh <= bwd_cells;
}

template <typename T>
static cudaError_t launch_fwd_tiled_t(const void* x, const void* offset, const void* mask, void* out,
                                      const KParams& q, int dtype, cudaStream_t st) {
    const int max_cells = kFwdBoxBytes / kCellBytes;  // two CTAs per SM
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, max_cells);
    if (tg.bw * tg.bh > max_cells) return cudaErrorInvalidConfiguration;
    CUtensorMap map;
    if (!make_x_tensor_map(&map, x, q, dtype, tg.bw, tg.bh)) return cudaErrorNotSupported;
    const size_t smem = (size_t)tg.bw * tg.bh * kCellBytes;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {  // per device: the attribute lives in the context
        cudaError_t e = cudaFuncSetAttribute(fwd_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kFwdBoxBytes);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const unsigned grid = (unsigned)((size_t)q.n * tg.chunks * tg.tiles_h * tg.tiles_w);
    fwd_tiled_kernel<T><<<grid, 256, smem, st>>>(map, (const T*)x, (const T*)offset, (const T*)mask, (T*)out, q, tg);
    count_launch(1);
    return cudaGetLastError();
}

static void geom_numbers(const KParams& q, const TileGeom& tg, long long smem, int out[8]) {
    out[0] = tg.th; out[1] = tg.tw; out[2] = tg.bw; out[3] = tg.bh; out[4] = tg.halo_x; out[5] = tg.halo_y;
    out[6] = (int)((long long)q.n * tg.chunks * tg.tiles_h * tg.tiles_w);
    out[7] = (int)smem;
}

void fwd_tiled_plan(const KParams& q, int dtype, int out[8]) {
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, kFwdBoxBytes / kCellBytes);
    geom_numbers(q, tg, (long long)tg.bw * tg.bh * kCellBytes, out);
}

void gather_tiled_plan(const KParams& q, int dtype, int stage_bytes, int out[8]) {
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, kMaxBoxBytes / kCellBytes);
    geom_numbers(q, tg, (long long)tg.bw * tg.bh * kCellBytes + stage_bytes, out);
}

cudaError_t launch_fwd_tiled(const void* x, const void* offset, const void* mask, void* out,
                             const KParams& q, int dtype, cudaStream_t st) {
    return dtype == DCNV3_F32 ? launch_fwd_tiled_t<float>(x, offset, mask, out, q, dtype, st)
                              : launch_fwd_tiled_t<__nv_bfloat16>(x, offset, mask, out, q, dtype, st);
}

}  // namespace dcnv3
