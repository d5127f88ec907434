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
#include <mutex>
#include <unordered_map>

#include "dcnv3_kernels.h"
#include "dcnv3_tiled.cuh"

namespace dcnv3 {

template <typename T>
__device__ __forceinline__ const T* global_slab(const T* x, const KParams& q, int n, int yp, int xp, int g) {
    const int y = yp - q.ph, xx = xp - q.pw;
    if (y < 0 || y >= q.h || xx < 0 || xx >= q.w) return nullptr;
    return x + ((((size_t)n * q.h + y) * q.w + xx) * q.G + g) * kGC;
}

template <typename T, bool STAGED>
__global__ void __launch_bounds__(kTiledWarps * 32, 2)
fwd_tiled_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap offmap,
                 const T* __restrict__ x, const T* __restrict__ offset, const T* __restrict__ mask,
                 T* __restrict__ out, const KParams q, const TileGeom tg) {
    using C = Chunk<T>;
    using RS = RowStage<T>;
    constexpr bool WMIX = sizeof(T) == 2 && DCNV3_BF16_FWD_W;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t sbar[kTiledWarps];
    pdl_launch_dependents();  // the next kernel of the stream may start its prologue (it waits before reading)

    const int box_bytes = tg.bw * tg.bh * kCellBytes;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
#pragma unroll
        for (int i = 0; i < kTiledWarps; ++i) mbar_init(&sbar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();  // from here on global memory is read: everything the previous kernels wrote is visible
    const TileCtx ctx = decode_tile<C::PXW>(q, tg, blockIdx.x);
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)box_bytes);
        tma_load_4d(smem, &xmap, &bar, ctx.chunk * C::GQ * kGC, ctx.cx0 - q.pw, ctx.cy0 - q.ph, ctx.n);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g_l = lane % C::GQ, px_l = lane / C::GQ;
    const int rot = Slab<T>::rot_of(px_l);
    unsigned char* st = smem + (size_t)box_bytes + warp * RS::BYTES;  // the warp's side slot
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    uint32_t sphase = 0;

    // side inputs of one warp iteration (PXW consecutive pixels of one output row) -> the warp's slot
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
        const bool real_g = chunk * C::GQ + g_l < q.G;    // false for the phantom groups of a trailing chunk
        const int g = min(chunk * C::GQ + g_l, q.G - 1);  // (phantom lanes shadow the last group, never store)
        bool waited = false;
        for (int it = warp; it < ctx.nit; it += kTiledWarps) {
            const int h = ctx.h0 + it / ctx.colblocks, wb = ctx.w0 + (it % ctx.colblocks) * C::PXW;
            const int npx = min(C::PXW, ctx.w0 + ctx.tw - wb);
            const bool valid = real_g && px_l < npx;
            const int w = wb + min(px_l, npx - 1);  // idle lanes shadow the last pixel (loads stay in bounds)
            const size_t pixel = ((size_t)n * q.ho + h) * q.wo + w;
            const size_t pg = pixel * q.G + g, ps = STAGED ? pg : side_entry(q, pixel, g);  // (half groups are never staged)
            const T* offp = offset + ps * 18;
            const T* mskp = mask + ps * 9;
            float ref0, ref1;
            ref_point(q, h, w, ref0, ref1);
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
            f2 acc[8];
            float facc[WMIX ? 16 : 1];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0ull;
            if constexpr (WMIX) {
#pragma unroll
                for (int c = 0; c < 16; ++c) facc[c] = 0.f;
            }
            float ox, oy, ml, ox2 = 0.f, oy2 = 0.f, ml2 = 0.f;
            if (STAGED) {
                RS::tap(st, lane, 0, ox, oy, ml);
            } else {
                load_tap_inputs<T>(offp, mskp, 0, ox, oy, ml);
                load_tap_inputs<T>(offp, mskp, 1, ox2, oy2, ml2);
            }
            if (!waited) {  // the box is needed from here on
                mbar_wait(&bar, 0);
                waited = true;
            }
#pragma unroll 1
            for (int p = 0; p < kTaps; ++p) {
                const float cx = ox, cy = oy, cm = ml;
                if (STAGED) {
                    if (p + 1 < kTaps) RS::tap(st, lane, p + 1, ox, oy, ml);  // shared memory: one tap ahead is enough
                } else {
                    ox = ox2; oy = oy2; ml = ml2;
                    if (p + 2 < kTaps) load_tap_inputs<T>(offp, mskp, p + 2, ox2, oy2, ml2);  // prefetch two taps ahead
                }
                const Tap t = make_tap_live(q, ref0, ref1, p, cx, cy);
                const int bx = t.x0 - cx0, by = t.y0 - cy0;
                const bool inbox = bx >= 0 && bx + 1 < tg.bw && by >= 0 && by + 1 < tg.bh;
                const float mm = t.alive ? (logits ? expf(cm - mx) * inv_sum : cm) : 0.f;
                const float wa = t.dx1 * t.dy1 * mm, wb_ = t.dx1 * t.dy0 * mm;  // (y0,x0) (y1,x0)
                const float wc = t.dx0 * t.dy1 * mm, wd = t.dx0 * t.dy0 * mm;   // (y0,x1) (y1,x1)
                if (__builtin_expect(t.alive && !inbox, 0)) {
                    // rare: patch outside the staged box -> straight from global memory
#pragma unroll 1
                    for (int kk = 0; kk < 4; ++kk) {
                        const T* src = global_slab(x, q, n, t.y0 + (kk & 1), t.x0 + (kk >> 1), g);
                        if (src == nullptr) continue;
                        const float wk = (kk & 1) ? ((kk >> 1) ? wd : wb_) : ((kk >> 1) ? wc : wa);
#pragma unroll
                        for (int pc = 0; pc < C::NPIECE; ++pc) {
                            f2 v[C::PAIRS];
                            load_piece<T>(src + Slab<T>::chan_of(pc, rot), v);
#pragma unroll
                            for (int j = 0; j < C::PAIRS; ++j) ffma2s(acc[pc * C::PAIRS + j], v[j], wk);
                        }
                    }
                } else if constexpr (WMIX) {
                    // bf16 weights meet the packed slabs in FHFMA (fp32 accumulators)
                    const unsigned char* a = sbase + (size_t)(t.alive ? by * tg.bw + bx : 0) * kCellBytes;
                    unsigned va[8], vb[8];
                    const unsigned wab = pack_bf16x2(wa, wb_), wcd = pack_bf16x2(wc, wd);
                    Slab<T>::load_packed(a, rot, va);
                    Slab<T>::load_packed(a + (size_t)tg.bw * kCellBytes, rot, vb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 0>(facc[2 * c], va[c], wab); fhfma_x<1, 0>(facc[2 * c + 1], va[c], wab); }
                    Slab<T>::load_packed(a + kCellBytes, rot, va);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 1>(facc[2 * c], vb[c], wab); fhfma_x<1, 1>(facc[2 * c + 1], vb[c], wab); }
                    Slab<T>::load_packed(a + (size_t)(tg.bw + 1) * kCellBytes, rot, vb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 0>(facc[2 * c], va[c], wcd); fhfma_x<1, 0>(facc[2 * c + 1], va[c], wcd); }
#pragma unroll
                    for (int c = 0; c < 8; ++c) { fhfma_x<0, 1>(facc[2 * c], vb[c], wcd); fhfma_x<1, 1>(facc[2 * c + 1], vb[c], wcd); }
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
                    for (int c = 0; c < 8; ++c) ffma2s(acc[c], vb[c], wb_);
                    Slab<T>::load(a + (size_t)(tg.bw + 1) * kCellBytes, rot, vb);
#pragma unroll
                    for (int c = 0; c < 8; ++c) ffma2s(acc[c], va[c], wc);
#pragma unroll
                    for (int c = 0; c < 8; ++c) ffma2s(acc[c], vb[c], wd);
                }
            }
            if constexpr (WMIX) {  // (acc holds what the rare out-of-box taps added)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    acc[c] = pack2(lo_of(acc[c]) + facc[2 * c], hi_of(acc[c]) + facc[2 * c + 1]);
            }
            if (STAGED) {
                // every lane has consumed its slot values (they fed the arithmetic above): refill for the next iteration
                __syncwarp();
                if (it + kTiledWarps < ctx.nit) request(ctx, it + kTiledWarps);
            }
            if (valid) {
                if (q.cfs != nullptr) {
                    // centre-feature-scale blend (dcn_v3.py:146): core * (1 - s) + x_proj * s, the reference's three
                    // separately rounded operations; x_proj is this pixel's own slab of x (ho == h, wo == w here)
                    const float s = Elem<T>::ld(reinterpret_cast<const T*>(q.cfs) + ps);
                    const float oms = __fsub_rn(1.0f, s);
#pragma unroll
                    for (int pc = 0; pc < C::NPIECE; ++pc) {
                        f2 xo[C::PAIRS];
                        load_piece<T>(x + pg * kGC + Slab<T>::chan_of(pc, rot), xo);
#pragma unroll
                        for (int j = 0; j < C::PAIRS; ++j) {
                            const f2 a = acc[pc * C::PAIRS + j];
                            acc[pc * C::PAIRS + j] =
                                pack2(__fadd_rn(__fmul_rn(lo_of(a), oms), __fmul_rn(lo_of(xo[j]), s)),
                                      __fadd_rn(__fmul_rn(hi_of(a), oms), __fmul_rn(hi_of(xo[j]), s)));
                        }
                    }
                }
                T* dst = out + pg * kGC;
#pragma unroll
                for (int pc = 0; pc < C::NPIECE; ++pc)
                    store_piece<T>(dst + Slab<T>::chan_of(pc, rot), acc + pc * C::PAIRS);
            }
        }
        if (!waited) mbar_wait(&bar, 0);  // never leave with a TMA in flight
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static const EncodeTiledFn fn = [] {  // (thread-safe: initialised once)
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)p;
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

cudaError_t ensure_max_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::unordered_map<const void*, unsigned long long> done;  // kernel -> devices (bit mask) already set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    unsigned long long& m = done[kernel];
    if (dev < 64 && (m >> dev) & 1ull) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);  // lives in the device's context
    if (e == cudaSuccess && dev < 64) m |= 1ull << dev;
    return e;
}

// Encoding a tensor map costs a driver call (~1 us); layers call with the same tensors over and over, so the
// encoded maps are kept in a small per-thread table keyed by everything that went into them.
struct MapKey {
    const void* base;
    unsigned long long d0, d1, d2, d3;
    unsigned b0, b1, b2, es;
    bool operator==(const MapKey& o) const {
        return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && b0 == o.b0 && b1 == o.b1 &&
               b2 == o.b2 && es == o.es;
    }
};
struct MapSlot {
    MapKey key;
    CUtensorMap map;
    bool used;
};
static bool encode_cached(CUtensorMap* out, const MapKey& k, int rank, CUtensorMapDataType dt, const cuuint64_t* dims,
                          const cuuint64_t* strides, const cuuint32_t* box) {
    constexpr int kSlots = 128;
    static thread_local MapSlot table[kSlots];
    size_t hsh = (size_t)k.base * 0x9E3779B97F4A7C15ull;
    hsh ^= (k.d0 * 31 + k.d1) * 0xC2B2AE3D27D4EB4Full + k.d2 * 1315423911ull + k.d3 * 2654435761ull + k.b1 * 97 + k.b2 * 89 + k.b0;
    MapSlot& sl = table[(hsh >> 17) % kSlots];
    if (sl.used && sl.key == k) {
        *out = sl.map;
        return true;
    }
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(k.base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    sl.key = k;
    sl.map = *out;
    sl.used = true;
    return true;
}

bool make_x_tensor_map(CUtensorMap* map, const void* x, const KParams& q, int dtype, int bw, int bh) {
    const cuuint64_t es = dtype == DCNV3_F32 ? 4 : 2;
    const int gq = dtype == DCNV3_F32 ? 2 : 4;
    const cuuint64_t C = (cuuint64_t)q.G * q.gc;
    const cuuint64_t dims[4] = {C, (cuuint64_t)q.w, (cuuint64_t)q.h, (cuuint64_t)q.n};
    const cuuint64_t strides[3] = {C * es, (cuuint64_t)q.w * C * es, (cuuint64_t)q.h * q.w * C * es};
    const cuuint32_t box[4] = {(cuuint32_t)(gq * kGC), (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const MapKey k = {x, dims[0], dims[1], dims[2], dims[3], box[0], box[1], box[2], (unsigned)es};
    return encode_cached(map, k, 4, dtype == DCNV3_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         dims, strides, box);
}

// offsets / grad_offset as [N*Ho rows][Wo][G*18]: box = the 18 values of the GQ groups of a chunk for the PXW
// pixels of one warp iteration.  (The mask's 72-byte runs are not a legal box width; see RowStage.)
bool side_stageable(const KParams& q, int dtype) {
    const int gq = dtype == DCNV3_F32 ? 2 : 4;
    return q.gsh == 0 && q.G % gq == 0 && (long long)q.n * q.ho < (1ll << 31);  // (half groups: 72-byte runs, no TMA box)
}
bool make_side_tensor_map(CUtensorMap* map, const void* base, const KParams& q, int dtype, int per_group) {
    const cuuint64_t es = dtype == DCNV3_F32 ? 4 : 2;
    const int gq = dtype == DCNV3_F32 ? 2 : 4, pxw = 32 / gq;
    const cuuint64_t ch = (cuuint64_t)q.G * per_group;
    const cuuint64_t dims[3] = {ch, (cuuint64_t)q.wo, (cuuint64_t)q.n * q.ho};
    const cuuint64_t strides[2] = {ch * es, (cuuint64_t)q.wo * ch * es};
    const cuuint32_t box[3] = {(cuuint32_t)(gq * per_group), (cuuint32_t)pxw, 1};
    const MapKey k = {base, dims[0], dims[1], dims[2], 0, box[0], box[1], box[2], (unsigned)es};
    return encode_cached(map, k, 3, dtype == DCNV3_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         dims, strides, box);
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
    {   // 8-row tiles when 16x16 tiles would give fewer than four waves of CTAs (148 SMs x 2): the small InternImage
        // stages then run 3.5 instead of 1.7 waves (-3 % gather fp32, -3.4 % bf16 step, profiles/r02_ab.md)
        const int gq = dtype == DCNV3_F32 ? 2 : 4;
        const long long ctas = (long long)q.n * ((q.G + gq - 1) / gq) * ((q.ho + 15) / 16) * ((q.wo + 15) / 16);
        if (th == 16 && ctas < 4 * 296) th = 8;
    }
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

// forward box budget: with staged side inputs the per-warp slots share the CTA's shared memory
static int fwd_box_bytes(const KParams& q, int dtype) { return side_stageable(q, dtype) ? kMaxBoxBytes : kFwdBoxBytes; }

// Tiled kernels serve the InternImage configuration only -- and only images whose tile boxes fit shared
// memory (everything but extreme aspect ratios at large offset_scale); the rest runs the generic kernels.
bool tiled_applicable(const KParams& q_in, int dtype) {
    const KParams q = tiled_view(q_in);  // (32 channels per group run as half groups)
    // any group count: the last chunk may be partly empty (phantom groups are masked)
    if (!(q.P == 9 && q.kh == 3 && q.sh == 1 && q.sw == 1 && q.dh == 1 && q.dw == 1 && q.gc == kGC && q.ho == q.h &&
          q.wo == q.w && q.ph == 1 && q.pw == 1 && q.scale > 0.f && q.scale <= 16.f && q.h <= 16384 && q.w <= 16384 &&
          (long long)q.w * q.G * q.gc < (1ll << 30)))  // (32-bit element offsets inside an image row)
        return false;
    const int fwd_cells = fwd_box_bytes(q, dtype) / kCellBytes, bwd_cells = kMaxBoxBytes / kCellBytes;
    const TileGeom a = make_geom(q, dtype, 16, 16, 3.0f, fwd_cells), b = make_geom(q, dtype, 16, 16, 3.0f, bwd_cells);
    return a.bw * a.bh <= fwd_cells && b.bw * b.bh <= bwd_cells;
}

template <typename T, bool STAGED>
static cudaError_t launch_fwd_variant(const CUtensorMap& map, const CUtensorMap& offmap, const void* x, const void* offset,
                                      const void* mask, void* out, const KParams& q, const TileGeom& tg, cudaStream_t st) {
    const size_t smem = (size_t)tg.bw * tg.bh * kCellBytes + (STAGED ? kTiledWarps * RowStage<T>::BYTES : 0);
    cudaError_t e = ensure_max_smem((const void*)fwd_tiled_kernel<T, STAGED>,
                                    STAGED ? kMaxBoxBytes + kTiledWarps * RowStage<T>::BYTES : kFwdBoxBytes);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((size_t)q.n * tg.chunks * tg.tiles_h * tg.tiles_w);
    return launch_pdl(fwd_tiled_kernel<T, STAGED>, grid, kTiledWarps * 32, smem, st, map, offmap, (const T*)x,
                      (const T*)offset, (const T*)mask, (T*)out, q, tg);
}

template <typename T>
static cudaError_t launch_fwd_tiled_t(const void* x, const void* offset, const void* mask, void* out,
                                      const KParams& q, int dtype, cudaStream_t st) {
    const int max_cells = fwd_box_bytes(q, dtype) / kCellBytes;  // two CTAs per SM
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, max_cells);
    if (tg.bw * tg.bh > max_cells) return cudaErrorInvalidConfiguration;
    CUtensorMap map, offmap;
    if (!make_x_tensor_map(&map, x, q, dtype, tg.bw, tg.bh)) return cudaErrorNotSupported;
    const bool staged = side_stageable(q, dtype);
    if (staged && !make_side_tensor_map(&offmap, offset, q, dtype, 18)) return cudaErrorNotSupported;
    if (!staged) offmap = map;  // unused by the kernel variant
    cudaError_t e = staged ? launch_fwd_variant<T, true>(map, offmap, x, offset, mask, out, q, tg, st)
                           : launch_fwd_variant<T, false>(map, offmap, x, offset, mask, out, q, tg, st);
    if (e != cudaSuccess) return e;
    count_launch(1);
    return cudaGetLastError();
}

static void geom_numbers(const KParams& q, const TileGeom& tg, long long smem, int out[8]) {
    out[0] = tg.th; out[1] = tg.tw; out[2] = tg.bw; out[3] = tg.bh; out[4] = tg.halo_x; out[5] = tg.halo_y;
    out[6] = (int)((long long)q.n * tg.chunks * tg.tiles_h * tg.tiles_w);
    out[7] = (int)smem;
}

void fwd_tiled_plan(const KParams& q_in, int dtype, int out[8]) {
    const KParams q = tiled_view(q_in);
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, fwd_box_bytes(q, dtype) / kCellBytes);
    const int slot = dtype == DCNV3_F32 ? RowStage<float>::BYTES : RowStage<__nv_bfloat16>::BYTES;
    geom_numbers(q, tg, (long long)tg.bw * tg.bh * kCellBytes + (side_stageable(q, dtype) ? kTiledWarps * slot : 0), out);
}

void gather_tiled_plan(const KParams& q, int dtype, int stage_bytes, int out[8]) {
    const TileGeom tg = make_geom(q, dtype, 16, 16, 3.0f, kMaxBoxBytes / kCellBytes);
    geom_numbers(q, tg, (long long)tg.bw * tg.bh * kCellBytes + stage_bytes, out);
}

cudaError_t launch_fwd_tiled(const void* x, const void* offset, const void* mask, void* out,
                             const KParams& q_in, int dtype, cudaStream_t st) {
    const KParams q = tiled_view(q_in);
    return dtype == DCNV3_F32 ? launch_fwd_tiled_t<float>(x, offset, mask, out, q, dtype, st)
                              : launch_fwd_tiled_t<__nv_bfloat16>(x, offset, mask, out, q, dtype, st);
}

}  // namespace dcnv3
