// Shared device-side definitions for the DCNv3 kernels (sm_100a).
//
// The coordinate arithmetic mirrors, operation by operation and in fp32 round-to-nearest without
// FMA contraction, what the reference computes in
//   layers/dcn_v3/utils.py:14-58   get_reference_points   (ref_y lands in channel 0, :52)
//   layers/dcn_v3/utils.py:65-103  generate_dilation_grids (tap p = i*kh + j)
//   layers/dcn_v3/op.py:77-87      loc = ref + grid*s + offset*s/[W_in,H_in];  g = 2*loc - 1
//   layers/dcn_v3/utils.py:142-166 q = 0.5*((g+1)*(max-1)); floor; clip; deltas from clipped corners
// so that floor()/clip decisions are bit-identical to the oracle's.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcnv3_b200.h"

#define DCNV3_MAX_TAPS 49  // kh*kw <= 7*7

namespace dcnv3 {

// Kernel-side parameter block (passed by value, lives in constant bank 0).
struct KParams {
    int n, h, w, ho, wo, G, gc, P, kh;
    int sh, sw, ph, pw, dh, dw;
    int hin, win;          // padded extent
    float hin_f, win_f;    // (float)H_in, (float)W_in
    float rhin_f, rwin_f;  // RN(1/H_in), RN(1/W_in): seeds of the correctly rounded divisions
    float hm2_f, wm2_f;    // (float)(H_in-2), (float)(W_in-2)   (utils.py:142-143: max-1)
    float y0c, x0c;        // (d*(k-1))//2 + 0.5                (utils.py:29-33)
    float scale;           // offset_scale
    float fx, fy;          // d xq/d offset0 = (W_in-2)*s/W_in ; d yq/d offset1 = (H_in-2)*s/H_in
    unsigned flags;
    // Tiled kernels on 32 channels per group (InternImage-H): every group is run as two 16-channel half groups
    // that share the group's offsets and mask.  gsh = 1 then, G counts the HALF groups (x / out / grad_x are
    // addressed with it unchanged: channel = g*16 + k) and the side tensors are indexed with g >> gsh among
    // G >> gsh groups per pixel (side_entry below).  0 everywhere else.
    int gsh;
    // centre-feature-scale blend fused around the op (reference layers/dcn_v3/dcn_v3.py:138-146; tiled kernels
    // only): out = core * (1 - s) + x * s with s = cfs[n, h, w, g] broadcast over the group's channels.
    // NULL = plain op.  grad_cfs is written by the gather kernel.
    const void* cfs;
    void* grad_cfs;
    float gs0[DCNV3_MAX_TAPS];  // (dx_p / W_in) * s   -- grid*offset_scale, channel 0 (op.py:82)
    float gs1[DCNV3_MAX_TAPS];  // (dy_p / H_in) * s
};

// (pixel, group) entry of offset / mask / their gradients for pixel index `pixel` and kernel-side group g
__host__ __device__ __forceinline__ size_t side_entry(const KParams& q, size_t pixel, int g) {
    return pixel * (size_t)(q.G >> q.gsh) + (size_t)(g >> q.gsh);
}
// the view of the parameters the tiled kernels run on (see KParams::gsh)
static inline KParams tiled_view(const KParams& q) {
    KParams v = q;
    if (q.gc == 32 && q.gsh == 0) { v.G = q.G * 2; v.gc = 16; v.gsh = 1; }
    return v;
}

struct Tap {
    int x0, y0;        // clipped lower corner (x1 = x0+1, y1 = y0+1 whenever the tap is alive)
    float dx0, dx1, dy0, dy1;
    bool alive;        // false <=> a clipped corner pair coincides => the tap contributes exactly 0
};

// a / b, correctly rounded, from r = RN(1/b): q0 = RN(a*r); q = RN(q0 + (a - q0*b)*r)  (the tail of
// the IEEE division sequence without its special-case checks; b is a small positive integer here).
__device__ __forceinline__ float div_rn(float a, float b, float r) {
    const float q0 = __fmul_rn(a, r);
    const float e = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(e, r, q0);
}

// One normalised coordinate -> pixel coordinate, exactly as op.py:82-87 + utils.py:142.
__device__ __forceinline__ float pixel_coord(float ref, float gs, float off, float scale, float dim_f,
                                             float rdim_f, float dim_m2_f) {
    float loc = __fadd_rn(ref, gs);
    loc = __fadd_rn(loc, div_rn(__fmul_rn(off, scale), dim_f, rdim_f));
    const float g = __fsub_rn(__fmul_rn(2.0f, loc), 1.0f);
    return __fmul_rn(0.5f, __fmul_rn(__fadd_rn(g, 1.0f), dim_m2_f));
}

// Reference point of output pixel (h, w): (h*sh + y0c)/H_in goes to channel 0 (paired with W_in
// further down -- the reference's transposed sampling), (w*sw + x0c)/W_in to channel 1.
__device__ __forceinline__ void ref_point(const KParams& q, int h, int w, float& ref0, float& ref1) {
    ref0 = __fdiv_rn(__fadd_rn((float)(h * q.sh), q.y0c), q.hin_f);
    ref1 = __fdiv_rn(__fadd_rn((float)(w * q.sw), q.x0c), q.win_f);
}

// One axis of a tap: clipped lower corner index, the two deltas from the clipped corners
// (utils.py:146-166) and whether the corner pair is distinct.  Everything stays in the float domain
// except the final index conversion, so huge or non-finite coordinates need no special casing: they
// fail the range test and the tap is dead (contributes exactly 0).
struct Axis {
    int i0;
    float d0, d1;
    bool alive;
};
__device__ __forceinline__ Axis make_axis(float coord, float max_f /* dim-1 */) {
    Axis a;
    const float f = floorf(coord);
    a.alive = (f >= 0.0f) && (f < max_f);          // clipped corners coincide otherwise
    const float c0 = fminf(fmaxf(f, 0.0f), max_f);
    const float c1 = fminf(fmaxf(f + 1.0f, 0.0f), max_f);
    a.i0 = a.alive ? (int)c0 : 0;
    a.d0 = a.alive ? __fsub_rn(coord, c0) : 0.0f;
    a.d1 = a.alive ? __fsub_rn(c1, coord) : 0.0f;
    return a;
}
// The same for callers that skip dead taps altogether: for a live axis the clips are identities
// (0 <= f and f + 1 <= dim-1), so the deltas are formed from f directly -- bit-identical to make_axis.
__device__ __forceinline__ Axis make_axis_live(float coord, float max_f /* dim-1 */) {
    Axis a;
    const float f = floorf(coord);
    a.alive = (f >= 0.0f) && (f < max_f);
    a.i0 = (int)f;
    a.d0 = __fsub_rn(coord, f);
    a.d1 = __fsub_rn(f + 1.0f, coord);
    return a;
}
__device__ __forceinline__ Axis axis_x_live(const KParams& q, float ref0, int p, float offx) {
    return make_axis_live(pixel_coord(ref0, q.gs0[p], offx, q.scale, q.win_f, q.rwin_f, q.wm2_f), q.wm2_f + 1.0f);
}
__device__ __forceinline__ Axis axis_y_live(const KParams& q, float ref1, int p, float offy) {
    return make_axis_live(pixel_coord(ref1, q.gs1[p], offy, q.scale, q.hin_f, q.rhin_f, q.hm2_f), q.hm2_f + 1.0f);
}
__device__ __forceinline__ Axis axis_x(const KParams& q, float ref0, int p, float offx) {
    return make_axis(pixel_coord(ref0, q.gs0[p], offx, q.scale, q.win_f, q.rwin_f, q.wm2_f), q.wm2_f + 1.0f);
}
__device__ __forceinline__ Axis axis_y(const KParams& q, float ref1, int p, float offy) {
    return make_axis(pixel_coord(ref1, q.gs1[p], offy, q.scale, q.hin_f, q.rhin_f, q.hm2_f), q.hm2_f + 1.0f);
}

__device__ __forceinline__ Tap make_tap(const KParams& q, float ref0, float ref1, int p, float offx,
                                        float offy) {
    const Axis ax = axis_x(q, ref0, p, offx), ay = axis_y(q, ref1, p, offy);
    Tap t;
    t.alive = ax.alive && ay.alive;
    t.x0 = ax.i0;
    t.y0 = ay.i0;
    t.dx0 = t.alive ? ax.d0 : 0.0f;
    t.dx1 = t.alive ? ax.d1 : 0.0f;
    t.dy0 = t.alive ? ay.d0 : 0.0f;
    t.dy1 = t.alive ? ay.d1 : 0.0f;
    return t;
}

// The same tap for kernels that stage the zero ring (pad = 1): a live axis needs no clips (make_axis_live); a dead tap
// gets zero deltas by selection, never by multiplication (its coordinates may be huge or NaN).
__device__ __forceinline__ Tap make_tap_live(const KParams& q, float ref0, float ref1, int p, float offx, float offy) {
    const Axis ax = axis_x_live(q, ref0, p, offx), ay = axis_y_live(q, ref1, p, offy);
    Tap t;
    t.alive = ax.alive && ay.alive;
    t.x0 = t.alive ? ax.i0 : 0;
    t.y0 = t.alive ? ay.i0 : 0;
    t.dx0 = t.alive ? ax.d0 : 0.0f;
    t.dx1 = t.alive ? ax.d1 : 0.0f;
    t.dy0 = t.alive ? ay.d0 : 0.0f;
    t.dy1 = t.alive ? ay.d1 : 0.0f;
    return t;
}

// ---- reference-dtype (bfloat16) coordinate arithmetic --------------------------------------------------
// Under the mixed_bfloat16 policy the reference computes reference points, grids, sampling locations,
// pixel coordinates and bilinear weights in x.dtype = bfloat16 (op.py:62,72,80-87; utils.py:130,140-172):
// every primitive rounds its result to bf16.  DCNV3_FLAG_REF_DTYPE reproduces that operation by operation
// (rb = round to bf16) so that the SAME cells are sampled with the SAME weights; the default bf16 path
// keeps fp32 coordinates (better numerics, different results: coordinates near 130 have a bf16 step of 1).
__device__ __forceinline__ float rb(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// one axis: pos = h*stride + start (utils.py:29-36, fp32 linspace value), ref_dim = the extent the
// reference point is divided by (utils.py:49-50), disp = the tap's integer displacement (utils.py:77-87),
// dim = the extent this coordinate is paired with (op.py:77: [W_in, H_in])
__device__ __forceinline__ Axis axis_refdtype(float pos, float ref_dim, int disp, float dim, float off, float scale) {
    const float sc = rb(scale);                                              // python float -> tensor of x.dtype
    const float ref = rb(__fdiv_rn(rb(pos), rb(ref_dim)));                   // utils.py:40-50
    const float gs = rb(rb(__fdiv_rn(rb((float)disp), rb(dim))) * sc);       // utils.py:91-95, op.py:82
    float loc = rb(ref + gs);                                                // op.py:82
    loc = rb(loc + rb(__fdiv_rn(rb(off * sc), rb(dim))));                    // op.py:85
    const float g = rb(rb(2.0f * loc) - 1.0f);                               // op.py:87
    const float coord = rb(0.5f * rb(rb(g + 1.0f) * rb(dim - 2.0f)));        // utils.py:142-143
    Axis a;
    const float f = floorf(coord), max_f = dim - 1.0f;
    a.alive = (f >= 0.0f) && (f < max_f);                                    // clipped corners coincide otherwise
    a.i0 = a.alive ? (int)f : 0;
    const float c0 = rb(f), c1 = rb(f + 1.0f);                               // utils.py:158-161: int -> x.dtype
    a.d0 = a.alive ? rb(coord - c0) : 0.0f;                                  // utils.py:163-166
    a.d1 = a.alive ? rb(c1 - coord) : 0.0f;
    return a;
}
__device__ __forceinline__ Tap make_tap_refdtype(const KParams& q, int h, int w, int p, float offx, float offy) {
    const int kw = q.P / q.kh, i = p / q.kh, j = p % q.kh;                   // utils.py:77-101: p = i*kh + j
    const int dxp = -((q.dw * (kw - 1)) / 2) + i * q.dw, dyp = -((q.dh * (q.kh - 1)) / 2) + j * q.dh;
    // channel 0: ref_y (divided by H_in) + x displacement / W_in + offset / W_in -> x coordinate
    const Axis ax = axis_refdtype((float)(h * q.sh) + q.y0c, q.hin_f, dxp, q.win_f, offx, q.scale);
    const Axis ay = axis_refdtype((float)(w * q.sw) + q.x0c, q.win_f, dyp, q.hin_f, offy, q.scale);
    Tap t;
    t.alive = ax.alive && ay.alive;
    t.x0 = ax.i0; t.y0 = ay.i0;
    t.dx0 = t.alive ? ax.d0 : 0.0f; t.dx1 = t.alive ? ax.d1 : 0.0f;
    t.dy0 = t.alive ? ay.d0 : 0.0f; t.dy1 = t.alive ? ay.d1 : 0.0f;
    return t;
}

// acc += bf16(half HV of v) * bf16(half HW of w), fp32 accumulator: PTX fma.rn.f32.bf16 (sm_100+), SASS FHFMA.BF16.
// The product of two bf16 values is exact in fp32, so packed bf16 pairs need no shift / mask to fp32 first.
// HV / HW = 0 / 1 pick the low / high half of v and of w
template <int HV, int HW>
__device__ __forceinline__ void fhfma_x(float& acc, unsigned v, unsigned w) {
    if (HV == 0 && HW == 0)
        asm("{\n.reg .b16 a, b, c, d;\nmov.b32 {a, b}, %1;\nmov.b32 {c, d}, %2;\nfma.rn.f32.bf16 %0, a, c, %0;\n}" : "+f"(acc) : "r"(v), "r"(w));
    else if (HV == 1 && HW == 0)
        asm("{\n.reg .b16 a, b, c, d;\nmov.b32 {a, b}, %1;\nmov.b32 {c, d}, %2;\nfma.rn.f32.bf16 %0, b, c, %0;\n}" : "+f"(acc) : "r"(v), "r"(w));
    else if (HV == 0 && HW == 1)
        asm("{\n.reg .b16 a, b, c, d;\nmov.b32 {a, b}, %1;\nmov.b32 {c, d}, %2;\nfma.rn.f32.bf16 %0, a, d, %0;\n}" : "+f"(acc) : "r"(v), "r"(w));
    else
        asm("{\n.reg .b16 a, b, c, d;\nmov.b32 {a, b}, %1;\nmov.b32 {c, d}, %2;\nfma.rn.f32.bf16 %0, b, d, %0;\n}" : "+f"(acc) : "r"(v), "r"(w));
}

// ---- element access -----------------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float ld_plain(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ float4 ld4(const float* p) {
        return __ldg(reinterpret_cast<const float4*>(p));
    }
    static __device__ __forceinline__ float4 ld4_plain(const float* p) {
        return *reinterpret_cast<const float4*>(p);
    }
    static __device__ __forceinline__ void st4(float* p, float4 v) {
        *reinterpret_cast<float4*>(p) = v;
    }
};
template <>
struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
        return __bfloat162float(__ldg(p));
    }
    static __device__ __forceinline__ float ld_plain(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) {
        *p = __float2bfloat16_rn(v);
    }
    static __device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        float4 v;
        v.x = __uint_as_float(r.x << 16);
        v.y = __uint_as_float(r.x & 0xffff0000u);
        v.z = __uint_as_float(r.y << 16);
        v.w = __uint_as_float(r.y & 0xffff0000u);
        return v;
    }
    static __device__ __forceinline__ float4 ld4_plain(const __nv_bfloat16* p) {
        const uint2 r = *reinterpret_cast<const uint2*>(p);
        float4 v;
        v.x = __uint_as_float(r.x << 16);
        v.y = __uint_as_float(r.x & 0xffff0000u);
        v.z = __uint_as_float(r.y << 16);
        v.w = __uint_as_float(r.y & 0xffff0000u);
        return v;
    }
    static __device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
        const __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned*>(&a);
        r.y = *reinterpret_cast<const unsigned*>(&b);
        *reinterpret_cast<uint2*>(p) = r;
    }
};

// ---- fixed-point accumulation (order-independent => bitwise reproducible scatter) ----------------
// Workspace prefix shared by the generic and the tiled backward:
//   [WsHeader, 256 B][per-image maxima: 2 x N unsigned, padded to 256 B]
// The fixed-point scale of an image comes from the maximum |grad_out| of THAT image only, so an image's
// grad_x never depends on what else is in the batch.  Maxima are kept as the bit pattern of |v| and compared
// as unsigned integers: order independent, and NaN > Inf > every finite value, so a non-finite grad_out is
// seen (and the image's grad_x becomes NaN) instead of being swallowed by fmaxf / float -> int conversion.
struct WsHeader {
    unsigned done;      // tiled path: CTAs of the last kernel that have finished (reset by the last one)
    unsigned any_redo;  // tiled path: some scatter CTA found overflow-suspect cells (reset by the last kernel)
    unsigned pad[62];
};
struct ImgMax {
    unsigned go_bits;   // max |grad_out| of the image, float bits
    unsigned m_bits;    // max |mask| of the image (generic path, raw masks), float bits
};
__host__ __device__ __forceinline__ size_t img_max_bytes(int n) {
    return ((size_t)n * sizeof(ImgMax) + 255) / 256 * 256;
}
__device__ __forceinline__ unsigned abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ bool bits_nonfinite(unsigned b) { return b >= 0x7f800000u; }

// Exponent e such that every contribution |v| <= amax_go*amax_m maps to |v * 2^e| <= 2^46, leaving
// 2^16 accumulations of headroom in an int64.
__device__ __forceinline__ int fixed_exponent(const ImgMax& im, bool mask_is_prob) {
    const float a = __uint_as_float(im.go_bits);
    const float m = mask_is_prob ? 1.0f : __uint_as_float(im.m_bits);
    const float bound = a * m;
    if (!(bound > 0.0f) || !(bound < 3.0e38f)) return 0;
    int ex;
    frexpf(bound, &ex);  // bound = f * 2^ex, f in [0.5, 1)  =>  bound < 2^ex
    int e = 46 - ex;
    return max(min(e, 120), -120);
}

// frexp exponent of max|grad_out| (amax < 2^ex), clamped so that every derived power of two is a
// normal float; 30 (=> unit scale) when the maximum is 0 or not finite.
__device__ __forceinline__ int fixed_exponent_raw(unsigned go_bits) {
    const float a = __uint_as_float(go_bits);
    if (!(a > 0.0f) || !(a < 3.0e38f)) return 30;
    int ex;
    frexpf(a, &ex);
    return max(min(ex, 120), -90);
}

__device__ __forceinline__ long long to_fixed(float v, int e) {
    return __float2ll_rn(ldexpf(v, e));  // exact scaling, exact conversion (24-bit mantissa)
}

}  // namespace dcnv3
