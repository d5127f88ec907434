// Generic DCNv3 kernels: any kernel size / stride / dilation / padding / group width the reference
// accepts.  Correctness-first path; the k=3/s=1/d=1 InternImage shapes are
// served by the tiled kernels in dcnv3_tiled.cu when they apply.
//
// forward : one thread per (pixel, group, channel chunk); 4 corner loads per tap straight from the
//           un-padded NHWC tensor (the zero ring of op.py:46 is never materialised).
// backward: one thread per (pixel, group); grad_offset / grad_mask need no cross-thread reduction;
//           grad_x contributions are scattered as 64-bit fixed-point integer atomics (integer
//           addition is associative => bitwise reproducible, no float atomics), then converted.
#include "dcnv3_kernels.h"

namespace dcnv3 {

template <typename T>
__device__ __forceinline__ const T* slab_ptr(const T* x, const KParams& q, int n, int yp, int xp,
                                             int g) {
    const int y = yp - q.ph, xx = xp - q.pw;
    if (y < 0 || y >= q.h || xx < 0 || xx >= q.w) return nullptr;
    return x + ((((size_t)n * q.h + y) * q.w + xx) * q.G + g) * q.gc;
}

// softmax over the P logits of one (pixel, group): returns max and 1/sum (dcn_v3.py:120-123)
template <typename T>
__device__ __forceinline__ void softmax_stats(const T* logits, int P, float& mx, float& inv_sum) {
    mx = -INFINITY;
    for (int p = 0; p < P; ++p) mx = fmaxf(mx, Elem<T>::ld(logits + p));
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += expf(Elem<T>::ld(logits + p) - mx);
    inv_sum = 1.0f / s;
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
fwd_generic_kernel(const T* __restrict__ x, const T* __restrict__ offset, const T* __restrict__ mask,
                   T* __restrict__ out, const KParams q) {
    const int cq_n = q.gc / VEC;
    const size_t total = (size_t)q.n * q.ho * q.wo * q.G * cq_n;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cq = (int)(idx % cq_n);
    const size_t pg = idx / cq_n;
    const int g = (int)(pg % q.G);
    const size_t pix = pg / q.G;
    const int w = (int)(pix % q.wo);
    const int h = (int)((pix / q.wo) % q.ho);
    const int n = (int)(pix / ((size_t)q.wo * q.ho));
    float ref0, ref1;
    ref_point(q, h, w, ref0, ref1);
    const T* off = offset + pg * q.P * 2;
    const T* msk = mask + pg * q.P;
    float mx = 0.f, inv_sum = 1.f;
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    if (logits) softmax_stats(msk, q.P, mx, inv_sum);
    float acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
    const int c0 = cq * VEC;
    // bf16 only: every intermediate rounded to bf16 in the reference's op order (dcnv3_common.cuh)
    const bool refdt = sizeof(T) == 2 && (q.flags & DCNV3_FLAG_REF_DTYPE);
    for (int p = 0; p < q.P; ++p) {
        const float ox = Elem<T>::ld(off + 2 * p), oy = Elem<T>::ld(off + 2 * p + 1);
        const Tap t = refdt ? make_tap_refdtype(q, h, w, p, ox, oy) : make_tap(q, ref0, ref1, p, ox, oy);
        if (!t.alive) continue;
        float m = Elem<T>::ld(msk + p);
        if (logits) m = expf(m - mx) * inv_sum;
        float wgt[4] = {t.dx1 * t.dy1, t.dx1 * t.dy0, t.dx0 * t.dy1, t.dx0 * t.dy0};
        if (refdt) {
            // utils.py:169-206 in bf16: weights, the four products, their sum (fp32 accumulator, rounded once),
            // the mask product and the running output are each rounded to bf16
#pragma unroll
            for (int k = 0; k < 4; ++k) wgt[k] = rb(wgt[k]);
            if (logits) m = rb(m);
            float s[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) s[c] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const T* src = slab_ptr(x, q, n, t.y0 + (k & 1), t.x0 + (k >> 1), g);
                if (src == nullptr) continue;
#pragma unroll
                for (int c = 0; c < VEC; ++c) s[c] += rb(Elem<T>::ld(src + c0 + c) * wgt[k]);
            }
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] = rb(acc[c] + rb(rb(s[c]) * m));
            continue;
        }
        float s[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) s[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // a b c d = (y0,x0) (y1,x0) (y0,x1) (y1,x1), utils.py:177-178
            const T* src = slab_ptr(x, q, n, t.y0 + (k & 1), t.x0 + (k >> 1), g);
            if (src == nullptr) continue;
            if (VEC == 4) {
                const float4 v = Elem<T>::ld4(src + c0);
                s[0] += v.x * wgt[k];
                s[1 % VEC] += v.y * wgt[k];
                s[2 % VEC] += v.z * wgt[k];
                s[3 % VEC] += v.w * wgt[k];
            } else {
#pragma unroll
                for (int c = 0; c < VEC; ++c) s[c] += Elem<T>::ld(src + c0 + c) * wgt[k];
            }
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] += s[c] * m;
    }
    T* o = out + pg * q.gc + c0;
    if (VEC == 4) {
        Elem<T>::st4(o, make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]));
    } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) Elem<T>::st(o + c, acc[c]);
    }
}

// per image: max |grad_out| and max |mask| -> workspace (atomicMax on the bit pattern of |v| is order
// independent and lets NaN / Inf through; dcnv3_common.cuh).  grid = (blocks per image, N).
template <typename T>
__global__ void __launch_bounds__(256)
amax_kernel(const T* __restrict__ grad_out, size_t n_go, const T* __restrict__ mask, size_t n_m,
            ImgMax* __restrict__ img_max, int n_images) {
  for (int n = blockIdx.y; n < n_images; n += gridDim.y) {
    unsigned a = 0u, b = 0u;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const T* go = grad_out + (size_t)n * n_go;
    const T* mk = mask + (size_t)n * n_m;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_go; i += stride)
        a = max(a, abs_bits(Elem<T>::ld(go + i)));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_m; i += stride)
        b = max(b, abs_bits(Elem<T>::ld(mk + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&img_max[n].go_bits, a);
        atomicMax(&img_max[n].m_bits, b);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
bwd_generic_kernel(const T* __restrict__ x, const T* __restrict__ offset, const T* __restrict__ mask,
                   const T* __restrict__ grad_out, T* __restrict__ grad_offset,
                   T* __restrict__ grad_mask, const ImgMax* __restrict__ img_max,
                   unsigned long long* __restrict__ acc64, const KParams q) {
    const size_t total = (size_t)q.n * q.ho * q.wo * q.G;
    const size_t pg = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pg >= total) return;
    const int g = (int)(pg % q.G);
    const size_t pix = pg / q.G;
    const int w = (int)(pix % q.wo);
    const int h = (int)((pix / q.wo) % q.ho);
    const int n = (int)(pix / ((size_t)q.wo * q.ho));
    const bool logits = q.flags & DCNV3_FLAG_MASK_LOGITS;
    const int e = fixed_exponent(img_max[n], logits);  // the image's own scale: batch invariant
    float ref0, ref1;
    ref_point(q, h, w, ref0, ref1);
    const T* off = offset + pg * q.P * 2;
    const T* msk = mask + pg * q.P;
    const T* go = grad_out + pg * q.gc;
    T* goff = grad_offset + pg * q.P * 2;
    T* gmsk = grad_mask + pg * q.P;
    float mx = 0.f, inv_sum = 1.f;
    if (logits) softmax_stats(msk, q.P, mx, inv_sum);
    float gm_dot_m = 0.f;       // sum_p m_p * dL/dm_p, for the softmax Jacobian
    float gm_local[DCNV3_MAX_TAPS];
    // bf16 only: sampling cells and weights as the reference computes them in bf16 (dcnv3_common.cuh); the
    // gradient arithmetic itself stays in fp32
    const bool refdt = sizeof(T) == 2 && (q.flags & DCNV3_FLAG_REF_DTYPE);
    for (int p = 0; p < q.P; ++p) {
        const float ox = Elem<T>::ld(off + 2 * p), oy = Elem<T>::ld(off + 2 * p + 1);
        const Tap t = refdt ? make_tap_refdtype(q, h, w, p, ox, oy) : make_tap(q, ref0, ref1, p, ox, oy);
        float m = Elem<T>::ld(msk + p);
        if (logits) m = expf(m - mx) * inv_sum;
        float gm = 0.f, gxq = 0.f, gyq = 0.f;
        if (t.alive) {
            const float wgt[4] = {t.dx1 * t.dy1, t.dx1 * t.dy0, t.dx0 * t.dy1, t.dx0 * t.dy0};
            float dot[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yp = t.y0 + (k & 1), xp = t.x0 + (k >> 1);
                const T* src = slab_ptr(x, q, n, yp, xp, g);
                float d = 0.f;
                if (src != nullptr) {
                    const float mw = m * wgt[k];
                    unsigned long long* dst =
                        acc64 + ((((size_t)n * q.h + (yp - q.ph)) * q.w + (xp - q.pw)) * q.G + g) * q.gc;
                    for (int c = 0; c < q.gc; ++c) {
                        const float gv = Elem<T>::ld(go + c);
                        d += gv * Elem<T>::ld(src + c);
                        atomicAdd(dst + c, (unsigned long long)to_fixed(gv * mw, e));
                    }
                }
                dot[k] = d;
            }
            gm = wgt[0] * dot[0] + wgt[1] * dot[1] + wgt[2] * dot[2] + wgt[3] * dot[3];
            gxq = m * (t.dy1 * (dot[2] - dot[0]) + t.dy0 * (dot[3] - dot[1]));
            gyq = m * (t.dx1 * (dot[1] - dot[0]) + t.dx0 * (dot[3] - dot[2]));
        }
        Elem<T>::st(goff + 2 * p, gxq * q.fx);
        Elem<T>::st(goff + 2 * p + 1, gyq * q.fy);
        if (logits) {
            gm_local[p] = gm;
            gm_dot_m += gm * m;
        } else {
            Elem<T>::st(gmsk + p, gm);
        }
    }
    if (logits) {
        for (int p = 0; p < q.P; ++p) {
            const float m = expf(Elem<T>::ld(msk + p) - mx) * inv_sum;
            Elem<T>::st(gmsk + p, m * (gm_local[p] - gm_dot_m));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
fixed_to_float_kernel(long long* __restrict__ acc64, const ImgMax* __restrict__ img_max,
                      T* __restrict__ grad_x, size_t per_image, unsigned flags, int n_images) {
  // grid = (blocks per image, min(N, 65535))
  for (int n = blockIdx.y; n < n_images; n += gridDim.y) {
    const ImgMax im = img_max[n];
    const int e = fixed_exponent(im, flags & DCNV3_FLAG_MASK_LOGITS);
    // a non-finite grad_out (or raw mask) cannot be represented in fixed point: the image's grad_x is NaN
    const bool bad = bits_nonfinite(im.go_bits) || (!(flags & DCNV3_FLAG_MASK_LOGITS) && bits_nonfinite(im.m_bits));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    long long* a = acc64 + (size_t)n * per_image;
    T* gx = grad_x + (size_t)n * per_image;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += stride) {
        Elem<T>::st(gx + i, bad ? __int_as_float(0x7fc00000) : (float)ldexp((double)a[i], -e));
        a[i] = 0;  // the workspace is left all-zero (DCNV3_FLAG_WORKSPACE_ZEROED contract, shared with the tiled path)
    }
  }
}

// ---- launchers -----------------------------------------------------------------------------------
static inline unsigned blocks_for(size_t total, int threads) {
    return (unsigned)((total + threads - 1) / threads);
}

template <typename T>
cudaError_t launch_fwd_generic_t(const void* x, const void* offset, const void* mask, void* out,
                                 const KParams& q, cudaStream_t st) {
    const size_t pg = (size_t)q.n * q.ho * q.wo * q.G;
    if (pg == 0) return cudaSuccess;
    const bool vec = (q.gc % 4 == 0);
    if (vec) {
        const size_t total = pg * (q.gc / 4);
        fwd_generic_kernel<T, 4><<<blocks_for(total, 256), 256, 0, st>>>(
            (const T*)x, (const T*)offset, (const T*)mask, (T*)out, q);
    } else {
        const size_t total = pg * q.gc;
        fwd_generic_kernel<T, 1><<<blocks_for(total, 256), 256, 0, st>>>(
            (const T*)x, (const T*)offset, (const T*)mask, (T*)out, q);
    }
    count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_fwd_generic(const void* x, const void* offset, const void* mask, void* out,
                               const KParams& q, int dtype, cudaStream_t st) {
    return dtype == DCNV3_F32 ? launch_fwd_generic_t<float>(x, offset, mask, out, q, st)
                              : launch_fwd_generic_t<__nv_bfloat16>(x, offset, mask, out, q, st);
}

size_t bwd_generic_workspace_bytes(const KParams& q) {
    return sizeof(WsHeader) + img_max_bytes(q.n) + sizeof(long long) * (size_t)q.n * q.h * q.w * q.G * q.gc;
}

template <typename T>
cudaError_t launch_bwd_generic_t(const void* x, const void* offset, const void* mask,
                                 const void* grad_out, void* grad_x, void* grad_offset,
                                 void* grad_mask, void* ws, const KParams& q, bool ws_clean, cudaStream_t st) {
    const size_t n_x = (size_t)q.n * q.h * q.w * q.G * q.gc;
    const size_t pg = (size_t)q.n * q.ho * q.wo * q.G;
    // (header and per-image maxima are rewritten by every call; the accumulators are re-zeroed by
    //  fixed_to_float_kernel)
    const size_t prefix = sizeof(WsHeader) + img_max_bytes(q.n);
    cudaError_t err = cudaSuccess;
    if (!ws_clean && (err = cudaMemsetAsync(ws, 0, bwd_generic_workspace_bytes(q), st)) != cudaSuccess) return err;
    ImgMax* img_max = (ImgMax*)((char*)ws + sizeof(WsHeader));
    unsigned long long* acc = (unsigned long long*)((char*)ws + prefix);
    if (pg > 0) {
        const size_t n_go = pg / q.n * q.gc, n_m = pg / q.n * q.P;  // per image
        const unsigned nb = (unsigned)max((size_t)1, min((size_t)148 * 8 / q.n + 1, (n_go + 255) / 256));
        amax_kernel<T><<<dim3(nb, min(q.n, 65535)), 256, 0, st>>>((const T*)grad_out, n_go, (const T*)mask,
                                                                  (q.flags & DCNV3_FLAG_MASK_LOGITS) ? 0 : n_m, img_max, q.n);
        bwd_generic_kernel<T><<<blocks_for(pg, 128), 128, 0, st>>>(
            (const T*)x, (const T*)offset, (const T*)mask, (const T*)grad_out, (T*)grad_offset,
            (T*)grad_mask, img_max, acc, q);
        count_launch(2);
    }
    if (n_x > 0) {
        const size_t per_image = n_x / q.n;
        const unsigned nb = (unsigned)max((size_t)1, min((size_t)148 * 16 / q.n + 1, (per_image + 255) / 256));
        fixed_to_float_kernel<T><<<dim3(nb, min(q.n, 65535)), 256, 0, st>>>((long long*)acc, img_max, (T*)grad_x, per_image,
                                                                            q.flags, q.n);
        count_launch(1);
    }
    // leave the prefix (per-image maxima) zeroed as well: the workspace is all-zero between calls
    if ((err = cudaMemsetAsync(ws, 0, prefix, st)) != cudaSuccess) return err;
    return cudaGetLastError();
}

cudaError_t launch_bwd_generic(const void* x, const void* offset, const void* mask,
                               const void* grad_out, void* grad_x, void* grad_offset, void* grad_mask,
                               void* ws, const KParams& q, int dtype, bool ws_clean, cudaStream_t st) {
    return dtype == DCNV3_F32
               ? launch_bwd_generic_t<float>(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, q, ws_clean, st)
               : launch_bwd_generic_t<__nv_bfloat16>(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, q, ws_clean, st);
}

// =====================================================================================================
// Sibling gather op: the sampling + aggregation of iSeg's deformable multi-head
// self-attention, reference layers/deformable_multihead_self_attention.py:102-175 (_bilinear_sample) followed by
// :233-235 (sum over the points, weighted by the attention weights) -- fused, so that the [N,H,W,heads,P,C]
// intermediate of the reference never exists:
//     out[n,h,w,hd,:] = sum_p attn[n,h,w,hd,p] * bilinear(value[n,:,:,hd,:], y[n,h,w,hd,p], x[n,h,w,hd,p])
// Conventions are that function's, not DCNv3's: absolute pixel coordinates, the four neighbour INDICES clamped to
// the image (:128-131), the weights taken from the unclamped fractional parts (:133-136: wy1 = y - floor(y)).
// Same skeleton as the generic DCNv3 kernels: thread per (pixel, head[, 4 channels]) forward; thread per
// (pixel, head, point) backward with grad_value accumulated in 64-bit fixed point (integer global atomics: bitwise
// reproducible, scale per image) and converted by fixed_to_float_kernel.
// =====================================================================================================
struct DaParams {
    int n, h, w, heads, points, c;   // c = channels per head
};

struct DaTap {
    int y0, x0, y1, x1;              // clamped neighbour indices
    float wy0, wy1, wx0, wx1;
};
__device__ __forceinline__ DaTap da_tap(const DaParams& q, float y, float x) {
    DaTap t;
    const float fy = floorf(y), fx = floorf(x);
    t.wy1 = __fsub_rn(y, fy); t.wx1 = __fsub_rn(x, fx);          // :133-134
    t.wy0 = __fsub_rn(1.0f, t.wy1); t.wx0 = __fsub_rn(1.0f, t.wx1);  // :135-136
    // cast to int then clip (:128-131); the float is clamped first so that huge or non-finite coordinates stay defined
    const float hy = (float)(q.h - 1), hx = (float)(q.w - 1);
    t.y0 = (int)fminf(fmaxf(fy, 0.f), hy); t.y1 = (int)fminf(fmaxf(fy + 1.0f, 0.f), hy);
    t.x0 = (int)fminf(fmaxf(fx, 0.f), hx); t.x1 = (int)fminf(fmaxf(fx + 1.0f, 0.f), hx);
    return t;
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
deform_attn_fwd_kernel(const T* __restrict__ value, const T* __restrict__ ys, const T* __restrict__ xs,
                       const T* __restrict__ attn, T* __restrict__ out, const DaParams q) {
    const int cpt = q.c / VEC;
    const size_t total = (size_t)q.n * q.h * q.w * q.heads * cpt;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cq = (int)(idx % cpt) * VEC;
    const size_t ph = idx / cpt;                      // (pixel, head)
    const int hd = (int)(ph % q.heads);
    const size_t n = ph / ((size_t)q.heads * q.w * q.h);
    const size_t row = (size_t)q.w * q.heads * q.c, img = (size_t)q.h * row;
    const T* vbase = value + n * img + (size_t)hd * q.c + cq;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int p = 0; p < q.points; ++p) {
        const size_t pi = ph * q.points + p;
        const DaTap t = da_tap(q, Elem<T>::ld(ys + pi), Elem<T>::ld(xs + pi));
        const float a = Elem<T>::ld(attn + pi);
        const float w00 = __fmul_rn(t.wy0, t.wx0), w01 = __fmul_rn(t.wy0, t.wx1);   // :162-165
        const float w10 = __fmul_rn(t.wy1, t.wx0), w11 = __fmul_rn(t.wy1, t.wx1);
        const T* p00 = vbase + (size_t)t.y0 * row + (size_t)t.x0 * q.heads * q.c;
        const T* p01 = vbase + (size_t)t.y0 * row + (size_t)t.x1 * q.heads * q.c;
        const T* p10 = vbase + (size_t)t.y1 * row + (size_t)t.x0 * q.heads * q.c;
        const T* p11 = vbase + (size_t)t.y1 * row + (size_t)t.x1 * q.heads * q.c;
        float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
        if (VEC == 4) {
            const float4 a4 = Elem<T>::ld4(p00), b4 = Elem<T>::ld4(p01), c4 = Elem<T>::ld4(p10), d4 = Elem<T>::ld4(p11);
            v00[0] = a4.x; v00[1] = a4.y; v00[2] = a4.z; v00[3] = a4.w;
            v01[0] = b4.x; v01[1] = b4.y; v01[2] = b4.z; v01[3] = b4.w;
            v10[0] = c4.x; v10[1] = c4.y; v10[2] = c4.z; v10[3] = c4.w;
            v11[0] = d4.x; v11[1] = d4.y; v11[2] = d4.z; v11[3] = d4.w;
        } else {
            v00[0] = Elem<T>::ld(p00); v01[0] = Elem<T>::ld(p01); v10[0] = Elem<T>::ld(p10); v11[0] = Elem<T>::ld(p11);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            // :167 (left to right), then x attention weight and the sum over the points (:234-235)
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, v00[v]), __fmul_rn(w01, v01[v])), __fmul_rn(w10, v10[v])),
                                      __fmul_rn(w11, v11[v]));
            acc[v] = __fadd_rn(acc[v], __fmul_rn(s, a));
        }
    }
    T* o = out + ph * q.c + cq;
    if (VEC == 4) Elem<T>::st4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
    else Elem<T>::st(o, acc[0]);
}

// A group of LPP lanes (a power of two <= 32, about the channels per head) works on one (pixel, head, point): lane = channel,
// so that the four corner reads and the four 64-bit atomic adds of a channel step are contiguous runs (one 256-byte run per
// corner for 32 channels instead of 32 scattered words); the three dot products are reduced by shuffles inside the group.
template <typename T>
__global__ void __launch_bounds__(128)
deform_attn_bwd_kernel(const T* __restrict__ value, const T* __restrict__ ys, const T* __restrict__ xs,
                       const T* __restrict__ attn, const T* __restrict__ grad_out, T* __restrict__ grad_y,
                       T* __restrict__ grad_x, T* __restrict__ grad_attn, const ImgMax* __restrict__ img_max,
                       unsigned long long* __restrict__ acc64, const DaParams q, const int lpp) {
    const size_t total = (size_t)q.n * q.h * q.w * q.heads * q.points;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = tid / lpp < total;              // (whole groups are in or out: blockDim is a multiple of lpp;
    const size_t pi = valid ? tid / lpp : total - 1;    //  lanes past the end shadow the last point and write nothing,
    const int cl = (int)(tid % lpp);                    //  so that the full-warp shuffles below stay well defined)
    const size_t ph = pi / q.points;
    const int hd = (int)(ph % q.heads);
    const size_t n = ph / ((size_t)q.heads * q.w * q.h);
    const int e = fixed_exponent(img_max[n], false);   // the image's own scale: max|grad_out| * max|attn|
    const size_t row = (size_t)q.w * q.heads * q.c, img = (size_t)q.h * row;
    const DaTap t = da_tap(q, Elem<T>::ld(ys + pi), Elem<T>::ld(xs + pi));
    const float a = Elem<T>::ld(attn + pi);
    const float w00 = __fmul_rn(t.wy0, t.wx0), w01 = __fmul_rn(t.wy0, t.wx1);
    const float w10 = __fmul_rn(t.wy1, t.wx0), w11 = __fmul_rn(t.wy1, t.wx1);
    const size_t o00 = n * img + (size_t)t.y0 * row + ((size_t)t.x0 * q.heads + hd) * q.c;
    const size_t o01 = n * img + (size_t)t.y0 * row + ((size_t)t.x1 * q.heads + hd) * q.c;
    const size_t o10 = n * img + (size_t)t.y1 * row + ((size_t)t.x0 * q.heads + hd) * q.c;
    const size_t o11 = n * img + (size_t)t.y1 * row + ((size_t)t.x1 * q.heads + hd) * q.c;
    const T* go = grad_out + ph * q.c;
    float s = 0.f, gy = 0.f, gx = 0.f;
    for (int c = valid ? cl : q.c; c < q.c; c += lpp) {
        const float g = Elem<T>::ld(go + c);
        const float v00 = Elem<T>::ld(value + o00 + c), v01 = Elem<T>::ld(value + o01 + c);
        const float v10 = Elem<T>::ld(value + o10 + c), v11 = Elem<T>::ld(value + o11 + c);
        s += g * (w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11);
        gy += g * ((v10 - v00) * t.wx0 + (v11 - v01) * t.wx1);   // d/dy: wy1 = y - floor(y), wy0 = 1 - wy1
        gx += g * ((v01 - v00) * t.wy0 + (v11 - v10) * t.wy1);
        const float ga = g * a;
        atomicAdd(acc64 + o00 + c, (unsigned long long)to_fixed(ga * w00, e));
        atomicAdd(acc64 + o01 + c, (unsigned long long)to_fixed(ga * w01, e));
        atomicAdd(acc64 + o10 + c, (unsigned long long)to_fixed(ga * w10, e));
        atomicAdd(acc64 + o11 + c, (unsigned long long)to_fixed(ga * w11, e));
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) {           // (fixed tree: the sums do not depend on scheduling)
        s += __shfl_xor_sync(0xffffffffu, s, o);
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
    }
    if (cl == 0 && valid) {
        Elem<T>::st(grad_attn + pi, s);
        Elem<T>::st(grad_y + pi, a * gy);
        Elem<T>::st(grad_x + pi, a * gx);
    }
}

static int lanes_per_point(int c) {
    int l = 1;
    while (l < c && l < 32) l <<= 1;
    return l;
}

size_t deform_attn_workspace_bytes(int n, int h, int w, int heads, int c) {
    return sizeof(WsHeader) + img_max_bytes(n) + sizeof(long long) * (size_t)n * h * w * heads * c;
}

template <typename T>
static cudaError_t launch_deform_attn_fwd_t(const void* value, const void* ys, const void* xs, const void* attn, void* out,
                                            const DaParams& q, cudaStream_t st) {
    const size_t ph = (size_t)q.n * q.h * q.w * q.heads;
    if (ph == 0) return cudaSuccess;
    if (q.c % 4 == 0)
        deform_attn_fwd_kernel<T, 4><<<blocks_for(ph * (q.c / 4), 256), 256, 0, st>>>((const T*)value, (const T*)ys, (const T*)xs,
                                                                                    (const T*)attn, (T*)out, q);
    else
        deform_attn_fwd_kernel<T, 1><<<blocks_for(ph * q.c, 256), 256, 0, st>>>((const T*)value, (const T*)ys, (const T*)xs,
                                                                              (const T*)attn, (T*)out, q);
    count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_deform_attn_fwd(const void* value, const void* ys, const void* xs, const void* attn, void* out, int n, int h,
                                   int w, int heads, int points, int c, int dtype, cudaStream_t st) {
    const DaParams q = {n, h, w, heads, points, c};
    return dtype == DCNV3_F32 ? launch_deform_attn_fwd_t<float>(value, ys, xs, attn, out, q, st)
                              : launch_deform_attn_fwd_t<__nv_bfloat16>(value, ys, xs, attn, out, q, st);
}

template <typename T>
static cudaError_t launch_deform_attn_bwd_t(const void* value, const void* ys, const void* xs, const void* attn,
                                            const void* grad_out, void* grad_value, void* grad_y, void* grad_x, void* grad_attn,
                                            void* ws, const DaParams& q, cudaStream_t st) {
    const size_t n_v = (size_t)q.n * q.h * q.w * q.heads * q.c, n_p = (size_t)q.n * q.h * q.w * q.heads * q.points;
    if (n_v == 0) return cudaSuccess;
    const size_t prefix = sizeof(WsHeader) + img_max_bytes(q.n);
    ImgMax* img_max = (ImgMax*)((char*)ws + sizeof(WsHeader));
    unsigned long long* acc = (unsigned long long*)((char*)ws + prefix);
    const size_t per_image = n_v / q.n, pts_per_image = n_p / q.n;
    const unsigned nb = (unsigned)max((size_t)1, min((size_t)148 * 8 / q.n + 1, (per_image + 255) / 256));
    // max |grad_out| and max |attn| per image -> the image's fixed-point scale
    amax_kernel<T><<<dim3(nb, min(q.n, 65535)), 256, 0, st>>>((const T*)grad_out, per_image, (const T*)attn, pts_per_image, img_max, q.n);
    if (n_p > 0) {
        const int lpp = lanes_per_point(q.c);
        deform_attn_bwd_kernel<T><<<blocks_for(n_p * lpp, 128), 128, 0, st>>>((const T*)value, (const T*)ys, (const T*)xs,
                                                                              (const T*)attn, (const T*)grad_out, (T*)grad_y,
                                                                              (T*)grad_x, (T*)grad_attn, img_max, acc, q, lpp);
    }
    const unsigned nb2 = (unsigned)max((size_t)1, min((size_t)148 * 16 / q.n + 1, (per_image + 255) / 256));
    fixed_to_float_kernel<T><<<dim3(nb2, min(q.n, 65535)), 256, 0, st>>>((long long*)acc, img_max, (T*)grad_value, per_image, 0u, q.n);
    count_launch(3);
    cudaError_t err = cudaMemsetAsync(ws, 0, prefix, st);  // leave the workspace all-zero
    return err != cudaSuccess ? err : cudaGetLastError();
}

cudaError_t launch_deform_attn_bwd(const void* value, const void* ys, const void* xs, const void* attn, const void* grad_out,
                                   void* grad_value, void* grad_y, void* grad_x, void* grad_attn, void* ws, int n, int h, int w,
                                   int heads, int points, int c, int dtype, cudaStream_t st) {
    const DaParams q = {n, h, w, heads, points, c};
    return dtype == DCNV3_F32
               ? launch_deform_attn_bwd_t<float>(value, ys, xs, attn, grad_out, grad_value, grad_y, grad_x, grad_attn, ws, q, st)
               : launch_deform_attn_bwd_t<__nv_bfloat16>(value, ys, xs, attn, grad_out, grad_value, grad_y, grad_x, grad_attn, ws, q, st);
}

// =====================================================================================================
// Sibling gather op: the sampling stage of iSeg's DCNv2, reference layers/dcn_v2.py:137-247
// (everything between the offset convolution :128-135 and the contraction with the kernel :249-271):
//     map[n,i,j,k,:] = mask[n,i,j,k] * bilinear(pad(x)[n], i + ph + py_k + oy[n,i,j,k], j + pw + px_k + ox[n,i,j,k])
// with that function's conventions: coordinates in the zero-padded image (:219), tap k = row-major over (py, px)
// (:110-111), all four neighbour indices AND the coordinate itself clipped to [0, H+1] x [0, W+1] (:165-175), weights
// from the clipped values (:193-211).  The padded image is never materialised (out-of-image neighbours read as 0).
// =====================================================================================================
struct D2Params {
    int n, h, w, c, kh, kw;
};
struct D2Tap {
    int y0, x0, y1, x1;            // clipped neighbour indices, padded coordinates
    float d0y, d1y, d0x, d1x;      // :193-194 from the clipped coordinate
    float iny, inx;                // 1 where the coordinate passes the clip (gradient of tf.clip_by_value)
};
__device__ __forceinline__ D2Tap d2_tap(const D2Params& q, int i, int j, int k, float oy, float ox) {
    const int ph = (q.kh - 1) / 2, pw = (q.kw - 1) / 2;
    const float gy = __fadd_rn((float)(i + ph + (k / q.kw - ph)), oy);   // :157-163 (integer sums, cast, + offset)
    const float gx = __fadd_rn((float)(j + pw + (k % q.kw - pw)), ox);
    const float hb = (float)(q.h + 1), wb = (float)(q.w + 1);             // :139-142
    const float fy = floorf(gy), fx = floorf(gx);
    const float y1 = fminf(fmaxf(fy + 1.0f, 0.f), hb), x1 = fminf(fmaxf(fx + 1.0f, 0.f), wb);   // :166-170
    const float y0 = fminf(fmaxf(fy, 0.f), hb), x0 = fminf(fmaxf(fx, 0.f), wb);               // :174
    const float gyc = fminf(fmaxf(gy, 0.f), hb), gxc = fminf(fmaxf(gx, 0.f), wb);             // :177
    D2Tap t;
    t.y0 = (int)y0; t.y1 = (int)y1; t.x0 = (int)x0; t.x1 = (int)x1;
    t.d0y = __fsub_rn(gyc, y0); t.d1y = __fsub_rn(y1, gyc);
    t.d0x = __fsub_rn(gxc, x0); t.d1x = __fsub_rn(x1, gxc);
    t.iny = (gy >= 0.f && gy <= hb) ? 1.f : 0.f;
    t.inx = (gx >= 0.f && gx <= wb) ? 1.f : 0.f;
    return t;
}
// element offset of padded pixel (yp, xp) in the un-padded image, or -1 in the zero border (or beyond it)
__device__ __forceinline__ long long d2_pixel(const D2Params& q, size_t n, int yp, int xp) {
    const int y = yp - (q.kh - 1) / 2, x = xp - (q.kw - 1) / 2;
    if (y < 0 || y >= q.h || x < 0 || x >= q.w) return -1;
    return (long long)(((n * q.h + y) * q.w + x) * (size_t)q.c);
}

template <typename T>
__global__ void __launch_bounds__(256)
dcnv2_sample_fwd_kernel(const T* __restrict__ x, const T* __restrict__ offs, const T* __restrict__ mask, T* __restrict__ out,
                        const D2Params q) {
    const int ks = q.kh * q.kw;
    const size_t total = (size_t)q.n * q.h * q.w * ks * q.c;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % q.c);
    const size_t pk = idx / q.c;                     // (pixel, tap)
    const int k = (int)(pk % ks);
    const size_t pix = pk / ks;
    const int j = (int)(pix % q.w), i = (int)((pix / q.w) % q.h);
    const size_t n = pix / ((size_t)q.w * q.h);
    const D2Tap t = d2_tap(q, i, j, k, Elem<T>::ld(offs + pk * 2), Elem<T>::ld(offs + pk * 2 + 1));
    auto v = [&](int yp, int xp) {
        const long long o = d2_pixel(q, n, yp, xp);
        return o < 0 ? 0.f : Elem<T>::ld(x + o + c);
    };
    // :193-211 weights, :231-236 [1,4] x [4,C] product in the order of the index list (:180-184), then x mask (:238)
    float s = __fmul_rn(__fmul_rn(t.d0y, t.d0x), v(t.y1, t.x1));
    s = __fadd_rn(s, __fmul_rn(__fmul_rn(t.d0y, t.d1x), v(t.y1, t.x0)));
    s = __fadd_rn(s, __fmul_rn(__fmul_rn(t.d1y, t.d0x), v(t.y0, t.x1)));
    s = __fadd_rn(s, __fmul_rn(__fmul_rn(t.d1y, t.d1x), v(t.y0, t.x0)));
    Elem<T>::st(out + idx, __fmul_rn(s, Elem<T>::ld(mask + pk)));
}

// (lane group per (pixel, tap), lane = channel: see deform_attn_bwd_kernel)
template <typename T>
__global__ void __launch_bounds__(128)
dcnv2_sample_bwd_kernel(const T* __restrict__ x, const T* __restrict__ offs, const T* __restrict__ mask,
                        const T* __restrict__ grad_out, T* __restrict__ grad_offs, T* __restrict__ grad_mask,
                        const ImgMax* __restrict__ img_max, unsigned long long* __restrict__ acc64, const D2Params q,
                        const int lpp) {
    const int ks = q.kh * q.kw;
    const size_t total = (size_t)q.n * q.h * q.w * ks;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = tid / lpp < total;
    const size_t pk = valid ? tid / lpp : total - 1;
    const int cl = (int)(tid % lpp);
    const int k = (int)(pk % ks);
    const size_t pix = pk / ks;
    const int j = (int)(pix % q.w), i = (int)((pix / q.w) % q.h);
    const size_t n = pix / ((size_t)q.w * q.h);
    const int e = fixed_exponent(img_max[n], false);
    const D2Tap t = d2_tap(q, i, j, k, Elem<T>::ld(offs + pk * 2), Elem<T>::ld(offs + pk * 2 + 1));
    const float m = Elem<T>::ld(mask + pk);
    const long long o11 = d2_pixel(q, n, t.y1, t.x1), o10 = d2_pixel(q, n, t.y1, t.x0);
    const long long o01 = d2_pixel(q, n, t.y0, t.x1), o00 = d2_pixel(q, n, t.y0, t.x0);
    const float w11 = t.d0y * t.d0x, w10 = t.d0y * t.d1x, w01 = t.d1y * t.d0x, w00 = t.d1y * t.d1x;
    const T* go = grad_out + pk * q.c;
    float gs = 0.f, gy = 0.f, gx = 0.f;
    for (int c = valid ? cl : q.c; c < q.c; c += lpp) {
        const float g = Elem<T>::ld(go + c);
        const float v11 = o11 < 0 ? 0.f : Elem<T>::ld(x + o11 + c), v10 = o10 < 0 ? 0.f : Elem<T>::ld(x + o10 + c);
        const float v01 = o01 < 0 ? 0.f : Elem<T>::ld(x + o01 + c), v00 = o00 < 0 ? 0.f : Elem<T>::ld(x + o00 + c);
        gs += g * (w11 * v11 + w10 * v10 + w01 * v01 + w00 * v00);
        gy += g * ((t.d0x * v11 + t.d1x * v10) - (t.d0x * v01 + t.d1x * v00));   // d/d(clipped y): d0y up, d1y down
        gx += g * ((t.d0y * v11 + t.d1y * v01) - (t.d0y * v10 + t.d1y * v00));
        const float gm = g * m;
        if (o11 >= 0) atomicAdd(acc64 + o11 + c, (unsigned long long)to_fixed(gm * w11, e));
        if (o10 >= 0) atomicAdd(acc64 + o10 + c, (unsigned long long)to_fixed(gm * w10, e));
        if (o01 >= 0) atomicAdd(acc64 + o01 + c, (unsigned long long)to_fixed(gm * w01, e));
        if (o00 >= 0) atomicAdd(acc64 + o00 + c, (unsigned long long)to_fixed(gm * w00, e));
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(0xffffffffu, gs, o);
        gy += __shfl_xor_sync(0xffffffffu, gy, o);
        gx += __shfl_xor_sync(0xffffffffu, gx, o);
    }
    if (cl == 0 && valid) {
        Elem<T>::st(grad_mask + pk, gs);
        Elem<T>::st(grad_offs + pk * 2, m * gy * t.iny);
        Elem<T>::st(grad_offs + pk * 2 + 1, m * gx * t.inx);
    }
}

size_t dcnv2_sample_workspace_bytes(int n, int h, int w, int c) {
    return sizeof(WsHeader) + img_max_bytes(n) + sizeof(long long) * (size_t)n * h * w * c;
}

template <typename T>
static cudaError_t launch_dcnv2_fwd_t(const void* x, const void* offs, const void* mask, void* out, const D2Params& q,
                                      cudaStream_t st) {
    const size_t total = (size_t)q.n * q.h * q.w * q.kh * q.kw * q.c;
    if (total == 0) return cudaSuccess;
    dcnv2_sample_fwd_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>((const T*)x, (const T*)offs, (const T*)mask, (T*)out, q);
    count_launch(1);
    return cudaGetLastError();
}
cudaError_t launch_dcnv2_sample_fwd(const void* x, const void* offs, const void* mask, void* out, int n, int h, int w, int c,
                                    int kh, int kw, int dtype, cudaStream_t st) {
    const D2Params q = {n, h, w, c, kh, kw};
    return dtype == DCNV3_F32 ? launch_dcnv2_fwd_t<float>(x, offs, mask, out, q, st)
                              : launch_dcnv2_fwd_t<__nv_bfloat16>(x, offs, mask, out, q, st);
}

template <typename T>
static cudaError_t launch_dcnv2_bwd_t(const void* x, const void* offs, const void* mask, const void* grad_out, void* grad_x,
                                      void* grad_offs, void* grad_mask, void* ws, const D2Params& q, cudaStream_t st) {
    const int ks = q.kh * q.kw;
    const size_t n_x = (size_t)q.n * q.h * q.w * q.c, n_pk = (size_t)q.n * q.h * q.w * ks;
    if (n_x == 0) return cudaSuccess;
    const size_t prefix = sizeof(WsHeader) + img_max_bytes(q.n);
    ImgMax* img_max = (ImgMax*)((char*)ws + sizeof(WsHeader));
    unsigned long long* acc = (unsigned long long*)((char*)ws + prefix);
    const size_t per_image = n_x / q.n, go_per_image = per_image * ks, m_per_image = n_pk / q.n;
    const unsigned nb = (unsigned)max((size_t)1, min((size_t)148 * 8 / q.n + 1, (go_per_image + 255) / 256));
    amax_kernel<T><<<dim3(nb, min(q.n, 65535)), 256, 0, st>>>((const T*)grad_out, go_per_image, (const T*)mask, m_per_image, img_max, q.n);
    const int lpp = lanes_per_point(q.c);
    dcnv2_sample_bwd_kernel<T><<<blocks_for(n_pk * lpp, 128), 128, 0, st>>>((const T*)x, (const T*)offs, (const T*)mask,
                                                                            (const T*)grad_out, (T*)grad_offs, (T*)grad_mask,
                                                                            img_max, acc, q, lpp);
    const unsigned nb2 = (unsigned)max((size_t)1, min((size_t)148 * 16 / q.n + 1, (per_image + 255) / 256));
    fixed_to_float_kernel<T><<<dim3(nb2, min(q.n, 65535)), 256, 0, st>>>((long long*)acc, img_max, (T*)grad_x, per_image, 0u, q.n);
    count_launch(3);
    cudaError_t err = cudaMemsetAsync(ws, 0, prefix, st);
    return err != cudaSuccess ? err : cudaGetLastError();
}
cudaError_t launch_dcnv2_sample_bwd(const void* x, const void* offs, const void* mask, const void* grad_out, void* grad_x,
                                    void* grad_offs, void* grad_mask, void* ws, int n, int h, int w, int c, int kh, int kw,
                                    int dtype, cudaStream_t st) {
    const D2Params q = {n, h, w, c, kh, kw};
    return dtype == DCNV3_F32 ? launch_dcnv2_bwd_t<float>(x, offs, mask, grad_out, grad_x, grad_offs, grad_mask, ws, q, st)
                              : launch_dcnv2_bwd_t<__nv_bfloat16>(x, offs, mask, grad_out, grad_x, grad_offs, grad_mask, ws, q, st);
}

}  // namespace dcnv3
