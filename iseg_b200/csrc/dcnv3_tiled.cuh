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

template <typename T>
struct Chunk {
    static constexpr int GQ = kCellBytes / (kGC * (int)sizeof(T));  // groups per cell: 2 (fp32) / 4 (bf16)
    static constexpr int PXW = 32 / GQ;                            // pixels per warp iteration
    static constexpr int NPIECE = kGC * (int)sizeof(T) / 16;       // 16-byte pieces per slab: 4 / 2
    static constexpr int CH_PER_PIECE = 16 / (int)sizeof(T);       // 4 / 8
};

// Geometry of one launch (host-computed, identical for every CTA).
struct TileGeom {
    int th, tw;            // output tile (rows h, columns w)
    int tiles_h, tiles_w;  // tiles per image
    int bw, bh;            // staged box, in cells: bw columns (x) by bh rows (y)
    int halo_x, halo_y;    // cells kept on each side of the nominal footprint
    int chunks;            // G / GQ
};

// Nominal sampling position of output row h / column w (zero offset, centre tap), reference
// arithmetic collapsed: xq = ((h + 1.5) / H_in) * (W_in - 2)  -- note H_in under h: the reference
// pairs ref_y with the x coordinate (utils.py:52), so output rows walk along input columns.
__device__ __forceinline__ float nominal_x(const KParams& q, int h) {
    return ((float)h + q.y0c) / q.hin_f * q.wm2_f;
}
__device__ __forceinline__ float nominal_y(const KParams& q, int w) {
    return ((float)w + q.x0c) / q.win_f * q.hm2_f;
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

// ---- one corner slab -> 16 fp32 values in the lane's rotated channel order ------------------------
// fp32: piece k of the loop is physical quad (k + px) & 3; bf16: physical half (k + px) & 1.
template <typename T>
struct Slab;

template <>
struct Slab<float> {
    // base = the lane's group slab inside a cell of the staged box; rot = px & 3
    static __device__ __forceinline__ void load(const unsigned char* base, int rot, float (&v)[16]) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 r = *reinterpret_cast<const float4*>(base + (((k + rot) & 3) << 4));
            v[4 * k] = r.x; v[4 * k + 1] = r.y; v[4 * k + 2] = r.z; v[4 * k + 3] = r.w;
        }
    }
    static __device__ __forceinline__ int rot_of(int px) { return px & 3; }
    // element offset of rotated piece k inside the 16 channels
    static __device__ __forceinline__ int chan_of(int k, int rot) { return ((k + rot) & 3) * 4; }
};

template <>
struct Slab<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const unsigned char* base, int rot, float (&v)[16]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint4 rr = *reinterpret_cast<const uint4*>(base + (((k + rot) & 1) << 4));
            const uint32_t r0 = rr.x, r1 = rr.y, r2 = rr.z, r3 = rr.w;
            v[8 * k + 0] = __uint_as_float(r0 << 16);
            v[8 * k + 1] = __uint_as_float(r0 & 0xffff0000u);
            v[8 * k + 2] = __uint_as_float(r1 << 16);
            v[8 * k + 3] = __uint_as_float(r1 & 0xffff0000u);
            v[8 * k + 4] = __uint_as_float(r2 << 16);
            v[8 * k + 5] = __uint_as_float(r2 & 0xffff0000u);
            v[8 * k + 6] = __uint_as_float(r3 << 16);
            v[8 * k + 7] = __uint_as_float(r3 & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ int rot_of(int px) { return px & 1; }
    static __device__ __forceinline__ int chan_of(int k, int rot) { return ((k + rot) & 1) * 8; }
};

// global 16-byte piece <-> fp32 registers
template <typename T>
__device__ __forceinline__ void load_piece(const T* p, float* v);
template <>
__device__ __forceinline__ void load_piece<float>(const float* p, float* v) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <>
__device__ __forceinline__ void load_piece<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
    v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void store_piece(T* p, const float* v);
template <>
__device__ __forceinline__ void store_piece<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&t);
}
template <>
__device__ __forceinline__ void store_piece<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    uint4 r;
    r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
    r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = r;
}

// The lane's 18 offsets and 9 mask values (one (pixel, group)) -> registers.
template <typename T>
__device__ __forceinline__ void load_offsets_mask(const T* off, const T* msk, float (&o)[18], float (&m)[9]);
template <>
__device__ __forceinline__ void load_offsets_mask<float>(const float* off, const float* msk, float (&o)[18],
                                                         float (&m)[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(off) + i);
        o[2 * i] = r.x; o[2 * i + 1] = r.y;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = __ldg(msk + i);
}
template <>
__device__ __forceinline__ void load_offsets_mask<__nv_bfloat16>(const __nv_bfloat16* off, const __nv_bfloat16* msk,
                                                                 float (&o)[18], float (&m)[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {  // 18 bf16 = 36 B per (pixel, group): 4-byte aligned pairs
        const unsigned r = __ldg(reinterpret_cast<const unsigned*>(off) + i);
        o[2 * i] = __uint_as_float(r << 16); o[2 * i + 1] = __uint_as_float(r & 0xffff0000u);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = __bfloat162float(__ldg(msk + i));
}

// softmax over the 9 taps, in registers (dcn_v3.py:120-123)
__device__ __forceinline__ void softmax9(float (&m)[9]) {
    float mx = m[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) mx = fmaxf(mx, m[i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { m[i] = expf(m[i] - mx); s += m[i]; }
    const float inv = 1.0f / s;
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] *= inv;
}

}  // namespace dcnv3
