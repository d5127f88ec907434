// Shared pieces of the tiled (shared-memory staged) DCNv3 kernels: 3x3 taps, stride 1, dilation 1,
// 16 channels per group -- the only configuration iSeg's InternImage ever instantiates
// (reference backbones/intern_image/intern_image_layer.py:62-74).
//
// Shared-memory cell = the 128 contiguous bytes of one pixel's group chunk:
//     fp32 : GQ = 2 groups x 16 ch x 4 B        bf16 : GQ = 4 groups x 16 ch x 2 B
// A warp works on 32/GQ pixels x GQ groups, lane = px*GQ + g, every lane owning all 16 channels of
// its (pixel, group).  The 16-byte pieces of a corner slab are visited in a lane-rotated order so
// that the 8 lanes of every LDS.128 phase hit 8 different 16-byte bank groups whatever cells they
// gather from (measured 126-128 B/clk/SM, tools/microbench.cu; 26 B/clk/SM without the rotation).
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "dcnv3_common.cuh"

namespace dcnv3 {

constexpr int kTaps = 9;
constexpr int kGC = 16;
constexpr int kCellBytes = 128;
constexpr int kMaxBoxBytes = 82 * 1024;  // staged input box of the forward / gather kernels; with the per-warp side slots two CTAs fit one SM
constexpr int kFwdBoxBytes = 100 * 1024; // forward box when the side inputs are not staged (no slots): two CTAs per SM
// Forward / gather launch shape: 8-warp CTAs, one tile each, two CTAs per SM (one CTA's box load runs under the
// other's arithmetic).  A persistent variant -- one 16-warp CTA per SM walking over tiles with two input boxes, the
// next tile's TMA load under the current tile's arithmetic -- was measured and lost by 3-8 % on every InternImage-T
// shape but 64x64 (profiles/r02_ab.md), so it is not kept.
constexpr int kTiledWarps = 8;

template <typename T>
struct Chunk {
    static constexpr int GQ = kCellBytes / (kGC * (int)sizeof(T));  // groups per cell: 2 (fp32) / 4 (bf16)
    static constexpr int PXW = 32 / GQ;                            // pixels per warp iteration
    static constexpr int NPIECE = kGC * (int)sizeof(T) / 16;       // 16-byte pieces per slab: 4 / 2
    static constexpr int CH_PER_PIECE = 16 / (int)sizeof(T);       // 4 / 8
    static constexpr int PAIRS = CH_PER_PIECE / 2;                 // packed fp32 pairs per piece: 2 / 4
};

// Geometry of one launch (host-computed, identical for every CTA).
struct TileGeom {
    int th, tw;            // output tile (rows h, columns w)
    int tiles_h, tiles_w;  // tiles per image
    int bw, bh;            // staged box, in cells: bw columns (x) by bh rows (y)
    int halo_x, halo_y;    // cells kept on each side of the nominal footprint
    int chunks;            // G / GQ
};

// One tile of the forward / gather kernels, decoded from its linear index (x fastest, then y, chunk, image).
struct TileCtx {
    int n, chunk, h0, w0, th, tw;  // image, group chunk, output tile origin and clipped extent
    int cx0, cy0;                  // origin of the staged input box (padded coordinates)
    int colblocks, nit;            // row segments per tile row / per tile (one warp iteration each)
};

// host helpers implemented in dcnv3_tiled_fwd.cu
TileGeom make_geom(const KParams& q, int dtype, int th, int tw, float reach, int max_cells);
bool make_x_tensor_map(CUtensorMap* map, const void* x, const KParams& q, int dtype, int bw, int bh);
// [N*Ho][Wo][G*per_group] view of offset / mask (and their gradients), box = one warp iteration of one chunk
bool side_stageable(const KParams& q, int dtype);
bool make_side_tensor_map(CUtensorMap* map, const void* base, const KParams& q, int dtype, int per_group);
// raises a kernel's dynamic shared-memory limit once per (kernel, device); safe to call from any thread
cudaError_t ensure_max_smem(const void* kernel, int bytes);

// Launch with programmatic stream serialisation: the kernel may be scheduled while its predecessor in the
// stream is still draining (its CTAs then wait in pdl_wait() before they touch global memory).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                     cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
void gather_tiled_plan(const KParams& q, int dtype, int stage_bytes, int out[8]);

// Nominal sampling position of output row h / column w (zero offset, centre tap), reference
// arithmetic collapsed: xq = ((h + 1.5) / H_in) * (W_in - 2)  -- note H_in under h: the reference
// pairs ref_y with the x coordinate (utils.py:52), so output rows walk along input columns.
__device__ __forceinline__ float nominal_x(const KParams& q, int h) {
    return ((float)h + q.y0c) / q.hin_f * q.wm2_f;
}
__device__ __forceinline__ float nominal_y(const KParams& q, int w) {
    return ((float)w + q.x0c) / q.win_f * q.hm2_f;
}

template <int PXW>
__device__ __forceinline__ TileCtx decode_tile(const KParams& q, const TileGeom& tg, int t) {
    TileCtx c;
    const int tx = t % tg.tiles_w; t /= tg.tiles_w;
    const int ty = t % tg.tiles_h; t /= tg.tiles_h;
    c.chunk = t % tg.chunks;
    c.n = t / tg.chunks;
    c.h0 = ty * tg.th;
    c.w0 = tx * tg.tw;
    c.th = min(tg.th, q.ho - c.h0);
    c.tw = min(tg.tw, q.wo - c.w0);
    // box origin in padded coordinates (output rows h walk along x, columns w along y); kept inside the padded
    // image: cells beyond it can only belong to dead taps
    c.cx0 = max(0, min((int)floorf(nominal_x(q, c.h0)) - tg.halo_x, q.win - tg.bw));
    c.cy0 = max(0, min((int)floorf(nominal_y(q, c.w0)) - tg.halo_y, q.hin - tg.bh));
    c.colblocks = (c.tw + PXW - 1) / PXW;
    c.nit = c.th * c.colblocks;
    return c;
}

// ---- mbarrier / TMA (sm_90+ PTX; SASS: SYNCS / UTMALDG) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!done);
}
// 4-D tiled load: coordinates (channel, x, y, n); out-of-bounds elements are zero-filled, which is
// exactly the zero ring tf.pad adds (op.py:46) when x/y are given in un-padded coordinates.
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// 3-D tiled load / store of the per-pixel side tensors, viewed as [N*Ho rows][Wo][channels]: coordinates
// (channel, w, row).  Loads zero-fill and stores drop whatever lies beyond Wo, so a row segment that is cut
// by the right image edge needs no special casing.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// 8-byte asynchronous global -> shared copy (SASS LDGSTS)
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- programmatic dependent launch (sm_90+): the next kernel of the stream may start its prologue while
//      this one drains; it must not touch global memory before pdl_wait() ------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- packed fp32 pairs: Blackwell issues two fp32 FMAs per instruction (PTX fma.rn.f32x2, SASS FFMA2) ----
typedef unsigned long long f2;  // {lo, hi} = two consecutive channels

__device__ __forceinline__ f2 pack2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f2 a) { return __uint_as_float((unsigned)a); }
__device__ __forceinline__ float hi_of(f2 a) { return __uint_as_float((unsigned)(a >> 32)); }
// acc += v * {w, w}
__device__ __forceinline__ void ffma2s(f2& acc, f2 v, float w) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(pack2(w, w)));
}
// acc += v * g
__device__ __forceinline__ void ffma2v(f2& acc, f2 v, f2 g) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(g));
}

// ---- one corner slab -> 8 packed pairs (16 channels) in the lane's rotated channel order ---------
// fp32: piece k of the loop is physical quad (k + px) & 3; bf16: physical half (k + px) & 1.
template <typename T>
struct Slab;

template <>
struct Slab<float> {
    // base = the lane's group slab inside a cell of the staged box; rot = px & 3
    static __device__ __forceinline__ void load(const unsigned char* base, int rot, f2 (&v)[8]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const ulonglong2 r = *reinterpret_cast<const ulonglong2*>(base + (((k + rot) & 3) << 4));
            v[2 * k] = r.x;
            v[2 * k + 1] = r.y;
        }
    }
    static __device__ __forceinline__ int rot_of(int px) { return px & 3; }
    // element offset of rotated piece k inside the 16 channels
    static __device__ __forceinline__ int chan_of(int k, int rot) { return ((k + rot) & 3) * 4; }
};

__device__ __forceinline__ f2 bf16x2_to_f2(unsigned r) {
    return pack2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
}

template <>
struct Slab<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const unsigned char* base, int rot, f2 (&v)[8]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint4 r = *reinterpret_cast<const uint4*>(base + (((k + rot) & 1) << 4));
            v[4 * k + 0] = bf16x2_to_f2(r.x);
            v[4 * k + 1] = bf16x2_to_f2(r.y);
            v[4 * k + 2] = bf16x2_to_f2(r.z);
            v[4 * k + 3] = bf16x2_to_f2(r.w);
        }
    }
    // the same 16 channels left packed: word 4k+j = channels (2j, 2j+1) of rotated piece k
    static __device__ __forceinline__ void load_packed(const unsigned char* base, int rot, unsigned (&v)[8]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint4 r = *reinterpret_cast<const uint4*>(base + (((k + rot) & 1) << 4));
            v[4 * k + 0] = r.x; v[4 * k + 1] = r.y; v[4 * k + 2] = r.z; v[4 * k + 3] = r.w;
        }
    }
    static __device__ __forceinline__ int rot_of(int px) { return px & 1; }
    static __device__ __forceinline__ int chan_of(int k, int rot) { return ((k + rot) & 1) * 8; }
};

// ---- bf16 operands straight from their packed words (PTX fma.rn.f32.bf16, SASS FHFMA.BF16, sm_100+): the
//      product of two bf16 values is exact in fp32 and the accumulator is fp32, so a packed slab needs no
//      shift / mask instructions to become fp32 pairs first (those were 28 % of the bf16 forward's instructions).
//      Gather kernel: slab x grad_out, both bf16 tensors -- bit-identical to the unpacked FFMA2 form.
//      Forward: the second operand is the tap's corner weight (bilinear weight x mask, computed in fp32), rounded
//      ONCE to bf16 -- the operand precision of a bf16 tensor-core product, and finer than the reference, which
//      under mixed_bfloat16 rounds the bilinear weight, the mask product and every partial sum to bf16
//      (utils.py:195-206).  Measured against the fp32 oracle on bf16 inputs: rms 1.66e-3 -> 2.19e-3 of rms(out),
//      max 3.5e-3 -> 4.6e-3 of max|out| (the output's own rounding to bf16 is the 1.66e-3); bar 1e-2.
//      -DDCNV3_BF16_MIXED=0 / -DDCNV3_BF16_FWD_W=0 rebuild the unpacked forms (profiles/r02_ab.md has the A/B).
#ifndef DCNV3_BF16_MIXED
#define DCNV3_BF16_MIXED 1
#endif
#ifndef DCNV3_BF16_FWD_W
#define DCNV3_BF16_FWD_W 1
#endif

// global 16-byte piece <-> packed pairs (PAIRS = 2 for fp32, 4 for bf16)
template <typename T>
__device__ __forceinline__ void load_piece(const T* p, f2* v);
template <>
__device__ __forceinline__ void load_piece<float>(const float* p, f2* v) {
    const ulonglong2 r = __ldg(reinterpret_cast<const ulonglong2*>(p));
    v[0] = r.x; v[1] = r.y;
}
template <>
__device__ __forceinline__ void load_piece<__nv_bfloat16>(const __nv_bfloat16* p, f2* v) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = bf16x2_to_f2(r.x); v[1] = bf16x2_to_f2(r.y); v[2] = bf16x2_to_f2(r.z); v[3] = bf16x2_to_f2(r.w);
}
__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&t);
}
template <typename T>
__device__ __forceinline__ void store_piece(T* p, const f2* v);
template <>
__device__ __forceinline__ void store_piece<float>(float* p, const f2* v) {
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v[0], v[1]);
}
template <>
__device__ __forceinline__ void store_piece<__nv_bfloat16>(__nv_bfloat16* p, const f2* v) {
    uint4 r;
    r.x = pack_bf16x2(lo_of(v[0]), hi_of(v[0])); r.y = pack_bf16x2(lo_of(v[1]), hi_of(v[1]));
    r.z = pack_bf16x2(lo_of(v[2]), hi_of(v[2])); r.w = pack_bf16x2(lo_of(v[3]), hi_of(v[3]));
    *reinterpret_cast<uint4*>(p) = r;
}

// ---- per-warp side slot ---------------------------------------------------------------------------
// A warp iteration works on PXW consecutive pixels of ONE output row (x GQ groups = 32 lanes).  Their
// offsets / mask values are [pixel][GQ groups][18 | 9] runs of 144 / 72 contiguous bytes per pixel in
// global memory (pixel stride G*18 / G*9 elements).  Letting every lane fetch its own 72+36 bytes makes
// each load instruction touch ~18 cache lines (a third of the L1 data-pipe wavefronts of the round-1
// kernels); instead the offsets arrive as one TMA box [PXW pixels][144 B] and the mask as 8-byte cp.async pieces, densely
// in a private shared-memory slot, and every lane reads its own values back with conflict-free LDS (lane
// stride 72 B for 8-byte loads, 36 B for 4-byte).  The gather kernel writes grad_offset / grad_mask over the
// values it has consumed (identical layout) and hands the slot to a TMA store.
// Needs 16-byte aligned offset runs and 8-byte aligned mask runs, i.e. G % GQ == 0 (side_stageable); other
// group counts keep the per-lane loads and the cooperative store below.
template <typename T>
struct RowStage {
    using C = Chunk<T>;
    static constexpr int OFF_PX = C::GQ * 18 * (int)sizeof(T);  // 144 bytes per pixel
    static constexpr int MSK_PX = C::GQ * 9 * (int)sizeof(T);   // 72
    static constexpr int OFF_BYTES = C::PXW * OFF_PX;           // 2304 (fp32) / 1152 (bf16)
    static constexpr int MSK_BYTES = C::PXW * MSK_PX;           // 1152 / 576
    static constexpr int BYTES = (OFF_BYTES + MSK_BYTES + 127) / 128 * 128;  // per warp: 3456 / 1792 (TMA: 128-byte aligned)
    static constexpr int LANE_OFF = 18 * (int)sizeof(T);        // 72 / 36: a lane's own offsets
    static constexpr int LANE_MSK = 9 * (int)sizeof(T);         // 36 / 18

    // Fill the slot with the side inputs of the PXW pixels starting at column w of row `row` (= n*Ho + h):
    // lane 0 asks the TMA unit for the offsets box (arrives on `bar`); the mask runs are 72 bytes per pixel --
    // not a legal TMA box width -- and travel as 8-byte cp.async pieces, consecutive lanes taking consecutive
    // pieces (completion: cp_async_wait_all + __syncwarp).  msk_row = first pixel's run of this chunk.
    static __device__ __forceinline__ void request(unsigned char* st, uint64_t* bar, const CUtensorMap* offmap,
                                                   const T* msk_row, int G, int chunk, int w, int row, int npx,
                                                   int lane) {
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)OFF_BYTES);
            tma_load_3d(st, offmap, bar, chunk * C::GQ * 18, w, row);
        }
        const int n = npx * 9;
        const size_t msk_stride = (size_t)G * 9 * sizeof(T);
#pragma unroll
        for (int i = 0; i < (C::PXW * 9 + 31) / 32; ++i) {
            const int c = lane + 32 * i;
            if (c < n) {
                const int px = (c * 7282) >> 16, r = c - px * 9;  // c / 9 for c < 288
                cp_async_8(st + OFF_BYTES + px * MSK_PX + r * 8,
                           reinterpret_cast<const unsigned char*>(msk_row) + px * msk_stride + r * 8);
            }
        }
    }
    // results written over the consumed inputs: grad_offset leaves through a TMA store (one lane; returns once
    // the slot may be overwritten), grad_mask through 8-byte pieces, consecutive lanes -> consecutive pieces
    static __device__ __forceinline__ void store_results(const unsigned char* st, const CUtensorMap* goffmap,
                                                         T* gmsk_row, int G, int chunk, int w, int row, int npx,
                                                         int lane) {
        fence_proxy_async();  // this lane's generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            tma_store_3d(goffmap, st, chunk * C::GQ * 18, w, row);
            bulk_commit();
        }
        const int n = npx * 9;
        const size_t msk_stride = (size_t)G * 9 * sizeof(T);
#pragma unroll
        for (int i = 0; i < (C::PXW * 9 + 31) / 32; ++i) {
            const int c = lane + 32 * i;
            if (c < n) {
                const int px = (c * 7282) >> 16, r = c - px * 9;
                *(reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(gmsk_row) + px * msk_stride) + r) =
                    *reinterpret_cast<const uint2*>(st + OFF_BYTES + px * MSK_PX + r * 8);
            }
        }
        if (lane == 0) bulk_wait_read();
        __syncwarp();
    }
    // the lane's tap p
    static __device__ __forceinline__ void tap(const unsigned char* st, int lane, int p, float& ox, float& oy,
                                               float& ml) {
        if (sizeof(T) == 4) {
            const float2 o = *reinterpret_cast<const float2*>(st + lane * LANE_OFF + p * 8);
            ox = o.x; oy = o.y;
            ml = *reinterpret_cast<const float*>(st + OFF_BYTES + lane * LANE_MSK + p * 4);
        } else {
            const unsigned r = *reinterpret_cast<const unsigned*>(st + lane * LANE_OFF + p * 4);
            ox = __uint_as_float(r << 16); oy = __uint_as_float(r & 0xffff0000u);
            ml = __uint_as_float((unsigned)*reinterpret_cast<const unsigned short*>(st + OFF_BYTES + lane * LANE_MSK + p * 2) << 16);
        }
    }
    static __device__ __forceinline__ float mask_at(const unsigned char* st, int lane, int p) {
        if (sizeof(T) == 4) return *reinterpret_cast<const float*>(st + OFF_BYTES + lane * LANE_MSK + p * 4);
        return __uint_as_float((unsigned)*reinterpret_cast<const unsigned short*>(st + OFF_BYTES + lane * LANE_MSK + p * 2) << 16);
    }
    // softmax over the lane's 9 staged logits (dcn_v3.py:120-123): max and 1/sum
    static __device__ __forceinline__ void softmax_stats(const unsigned char* st, int lane, float& mx, float& inv_sum) {
        mx = -INFINITY;
#pragma unroll
        for (int p = 0; p < kTaps; ++p) mx = fmaxf(mx, mask_at(st, lane, p));
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < kTaps; ++p) s += expf(mask_at(st, lane, p) - mx);
        inv_sum = 1.0f / s;
    }
    // cooperative, coalesced store of staged offset-shaped (144 B / pixel) and mask-shaped (72 B / pixel)
    // results, e.g. grad_offset / grad_mask (group counts the TMA store cannot serve)
    static __device__ __forceinline__ void store_off_msk(const unsigned char* st, T* off_row, T* msk_row, int G,
                                                         int npx, int ng, int lane) {
        __syncwarp();
        const size_t off_stride = (size_t)G * 18 * sizeof(T), msk_stride = (size_t)G * 9 * sizeof(T);
        if (G % C::GQ == 0) {  // every run is complete and 16-byte aligned
            const int n = npx * 9;
#pragma unroll
            for (int i = 0; i < (C::PXW * 9 + 31) / 32; ++i) {
                const int c = lane + 32 * i;
                if (c < n) {
                    const int px = (c * 7282) >> 16, r = c - px * 9;  // c / 9 for c < 288
                    *(reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(off_row) + px * off_stride) + r) =
                        *reinterpret_cast<const uint4*>(st + px * OFF_PX + r * 16);
                    *(reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(msk_row) + px * msk_stride) + r) =
                        *reinterpret_cast<const uint2*>(st + OFF_BYTES + px * MSK_PX + r * 8);
                }
            }
        } else {
            // group count not a multiple of the chunk: runs are only element aligned and the trailing
            // chunk has fewer real groups -> element-wise copy of the real part of every run
            const int ro = ng * 18, rm = ng * 9;  // elements per pixel
            for (int c = lane; c < npx * ro; c += 32) {
                const int px = c / ro, r = c - px * ro;
                reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(off_row) + px * off_stride)[r] =
                    reinterpret_cast<const T*>(st + px * OFF_PX)[r];
            }
            for (int c = lane; c < npx * rm; c += 32) {
                const int px = c / rm, r = c - px * rm;
                reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(msk_row) + px * msk_stride)[r] =
                    reinterpret_cast<const T*>(st + OFF_BYTES + px * MSK_PX)[r];
            }
        }
        __syncwarp();
    }
    // The same for half groups (KParams::gsh = 1): slot entries 2r and 2r+1 of a pixel hold identical results of group
    // r; the even one is stored.  G / ng = whole groups per pixel / in this chunk; off_row / msk_row = the first
    // pixel's run of this chunk's whole groups.
    static __device__ __forceinline__ void store_off_msk_halves(const unsigned char* st, T* off_row, T* msk_row, int G,
                                                                int npx, int ng, int lane) {
        __syncwarp();
        const size_t off_stride = (size_t)G * 18 * sizeof(T), msk_stride = (size_t)G * 9 * sizeof(T);
        const int ro = ng * 18, rm = ng * 9;  // elements per pixel
        for (int c = lane; c < npx * ro; c += 32) {
            const int px = c / ro, r = c - px * ro, gr = r / 18, e = r - gr * 18;
            reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(off_row) + px * off_stride)[r] =
                reinterpret_cast<const T*>(st + px * OFF_PX + 2 * gr * LANE_OFF)[e];
        }
        for (int c = lane; c < npx * rm; c += 32) {
            const int px = c / rm, r = c - px * rm, gr = r / 9, e = r - gr * 9;
            reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(msk_row) + px * msk_stride)[r] =
                reinterpret_cast<const T*>(st + OFF_BYTES + px * MSK_PX + 2 * gr * LANE_MSK)[e];
        }
        __syncwarp();
    }
};

// Per-tap inputs of one (pixel, group): the offset pair and the mask value (or logit) of tap p.
template <typename T>
__device__ __forceinline__ void load_tap_inputs(const T* off, const T* msk, int p, float& ox, float& oy, float& ml);
template <>
__device__ __forceinline__ void load_tap_inputs<float>(const float* off, const float* msk, int p, float& ox,
                                                       float& oy, float& ml) {
    const float2 r = __ldg(reinterpret_cast<const float2*>(off) + p);
    ox = r.x; oy = r.y;
    ml = __ldg(msk + p);
}
template <>
__device__ __forceinline__ void load_tap_inputs<__nv_bfloat16>(const __nv_bfloat16* off, const __nv_bfloat16* msk,
                                                               int p, float& ox, float& oy, float& ml) {
    const unsigned r = __ldg(reinterpret_cast<const unsigned*>(off) + p);  // 2 bf16, 4-byte aligned
    ox = __uint_as_float(r << 16); oy = __uint_as_float(r & 0xffff0000u);
    ml = __bfloat162float(__ldg(msk + p));
}

// softmax over the 9 tap logits of one (pixel, group) (dcn_v3.py:120-123): max and 1/sum
template <typename T>
__device__ __forceinline__ void softmax_stats9(const T* msk, float& mx, float& inv_sum) {
    mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < kTaps; ++p) mx = fmaxf(mx, Elem<T>::ld(msk + p));
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < kTaps; ++p) s += expf(Elem<T>::ld(msk + p) - mx);
    inv_sum = 1.0f / s;
}

}  // namespace dcnv3
