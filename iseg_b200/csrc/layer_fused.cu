// Inference fast path of the layer around the DCNv3 op: the element-wise chains the
// reference leaves to XLA, each as ONE pass over HBM.
//
//   dwconv_ln_act_kernel -- x1 = act(LayerNorm(DepthwiseConv2D(x) + bias))        reference layers/dcn_v3/dcn_v3.py:115-117
//                           (the branch that feeds the offset / mask / centre-scale projections)
//   ln_join_kernel       -- the joins of InternImageLayer (backbones/intern_image/intern_image_layer.py:126-172):
//        mode 0 (pre-norm):       z = r + gamma * y ;           also LayerNorm(z) for the sub-layer that follows  (:161-170)
//        mode 1 (post-norm):      z = r + gamma * LayerNorm(y)                                                    (:127-138)
//                                 (gamma = NULL: res-post-norm, :145-155)
//        mode 2:                  LayerNorm(y) only
//
// One warp per pixel, channels across the lanes in 4-channel pieces, statistics in fp32 from the values as they are
// stored (a bf16 tensor is normalised from its bf16-rounded sums, like the unfused chain), two-pass variance.
// Algorithmic bytes: (1 read + 1 write) * C * element size per pixel (+ the second output of mode 0).  Measured on the
// InternImage-B forward (bf16, batch 32, ncu): the joins run at 41 % (112 channels: a 224-byte row per warp bounds the
// bytes in flight) to 65 % of the HBM rate; the depthwise branch is bound by instruction issue, not by memory (75 % of
// the issue slots, ~600 warp instructions per pixel: nine taps, LayerNorm and four exact-erf GELUs for every 8 bytes a
// lane moves) at 5-10 x its memory time -- still 1.4 x faster than cuDNN depthwise conv + LayerNorm + GELU as three
// passes.  The 3x3 neighbourhood is served by L1: a CTA walks an 8-row x 16-column patch.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcnv3_b200.h"
#include "dcnv3_common.cuh"
#include "dcnv3_kernels.h"

namespace dcnv3 {

constexpr int kLnWarps = 8;
constexpr int kMaxLnChannels = 4096;  // per-warp fp32 row in shared memory: 16 KB x 8 warps

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ float round_to(float v);
template <>
__device__ __forceinline__ float round_to<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// mean and 1/sqrt(var + eps) of the warp's row (C floats in shared memory), two passes
__device__ __forceinline__ void row_stats(const float* row, int C, int lane, float eps, float& mean, float& rstd) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += row[c];
    mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = row[c] - mean;
        q += d * d;
    }
    rstd = rsqrtf(warp_sum(q) / (float)C + eps);
}

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_join_kernel(const T* __restrict__ y, const T* __restrict__ r, const T* __restrict__ gamma,
               const T* __restrict__ lw, const T* __restrict__ lb, T* __restrict__ out_sum, T* __restrict__ out_norm,
               long long rows, int C, float eps, int mode) {
    extern __shared__ float srow[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* row = srow + (size_t)warp * C;
    for (long long p = (long long)blockIdx.x * kLnWarps + warp; p < rows; p += (long long)gridDim.x * kLnWarps) {
        const T* yp = y + p * C;
        const T* rp = r ? r + p * C : nullptr;
        if (mode == 0) {
            // z = r + gamma * y  (products and sums rounded to T like the unfused chain), then LayerNorm(z)
            for (int c = lane * 4; c < C; c += 128) {
                const float4 a = Elem<T>::ld4(yp + c), b = Elem<T>::ld4(rp + c);
                float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
                if (gamma) g = Elem<T>::ld4(gamma + c);
                float4 z;
                z.x = round_to<T>(b.x + round_to<T>(a.x * g.x)); z.y = round_to<T>(b.y + round_to<T>(a.y * g.y));
                z.z = round_to<T>(b.z + round_to<T>(a.z * g.z)); z.w = round_to<T>(b.w + round_to<T>(a.w * g.w));
                *reinterpret_cast<float4*>(row + c) = z;
                Elem<T>::st4(out_sum + p * C + c, z);
            }
            __syncwarp();
            if (out_norm != nullptr) {
                float mean, rstd;
                row_stats(row, C, lane, eps, mean, rstd);
                for (int c = lane * 4; c < C; c += 128) {
                    const float4 z = *reinterpret_cast<const float4*>(row + c);
                    const float4 w4 = Elem<T>::ld4(lw + c), b4 = Elem<T>::ld4(lb + c);
                    Elem<T>::st4(out_norm + p * C + c,
                                 make_float4((z.x - mean) * rstd * w4.x + b4.x, (z.y - mean) * rstd * w4.y + b4.y,
                                             (z.z - mean) * rstd * w4.z + b4.z, (z.w - mean) * rstd * w4.w + b4.w));
                }
            }
        } else {
            for (int c = lane * 4; c < C; c += 128) *reinterpret_cast<float4*>(row + c) = Elem<T>::ld4(yp + c);
            __syncwarp();
            float mean, rstd;
            row_stats(row, C, lane, eps, mean, rstd);
            for (int c = lane * 4; c < C; c += 128) {
                const float4 v = *reinterpret_cast<const float4*>(row + c);
                const float4 w4 = Elem<T>::ld4(lw + c), b4 = Elem<T>::ld4(lb + c);
                float4 n4 = make_float4(round_to<T>((v.x - mean) * rstd * w4.x + b4.x), round_to<T>((v.y - mean) * rstd * w4.y + b4.y),
                                        round_to<T>((v.z - mean) * rstd * w4.z + b4.z), round_to<T>((v.w - mean) * rstd * w4.w + b4.w));
                if (mode == 1) {
                    const float4 b = Elem<T>::ld4(rp + c);
                    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (gamma) g = Elem<T>::ld4(gamma + c);
                    n4 = make_float4(b.x + round_to<T>(n4.x * g.x), b.y + round_to<T>(n4.y * g.y),
                                     b.z + round_to<T>(n4.z * g.z), b.w + round_to<T>(n4.w * g.w));
                }
                Elem<T>::st4(out_sum + p * C + c, n4);
            }
        }
        __syncwarp();  // the row buffer is reused by the warp's next pixel
    }
}

// weights: [k*k][C] (tap-major, channels contiguous), tap t = ky * k + kx; zero padding pad_lo before / (k-1-pad_lo) after
template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
dwconv_ln_act_kernel(const T* __restrict__ x, const T* __restrict__ wt, const T* __restrict__ bias,
                     const T* __restrict__ lw, const T* __restrict__ lb, T* __restrict__ out, int N, int H, int W, int C,
                     int k, int pad_lo, float eps, int act) {
    extern __shared__ float srow[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* row = srow + (size_t)warp * C;
    const long long rows = (long long)N * H * W;
    for (long long p = (long long)blockIdx.x * kLnWarps + warp; p < rows; p += (long long)gridDim.x * kLnWarps) {
        const int wq = (int)(p % W), hq = (int)((p / W) % H);
        const long long n = p / ((long long)W * H);
        for (int c = lane * 4; c < C; c += 128) {
            float4 acc = bias ? Elem<T>::ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int ky = 0; ky < k; ++ky) {
                const int yy = hq + ky - pad_lo;
                if (yy < 0 || yy >= H) continue;
                for (int kx = 0; kx < k; ++kx) {
                    const int xx = wq + kx - pad_lo;
                    if (xx < 0 || xx >= W) continue;
                    const float4 v = Elem<T>::ld4(x + ((n * H + yy) * W + xx) * C + c);
                    const float4 w4 = Elem<T>::ld4(wt + (size_t)(ky * k + kx) * C + c);
                    acc.x = fmaf(v.x, w4.x, acc.x); acc.y = fmaf(v.y, w4.y, acc.y);
                    acc.z = fmaf(v.z, w4.z, acc.z); acc.w = fmaf(v.w, w4.w, acc.w);
                }
            }
            // (the unfused chain stores the convolution result in T before normalising it)
            *reinterpret_cast<float4*>(row + c) =
                make_float4(round_to<T>(acc.x), round_to<T>(acc.y), round_to<T>(acc.z), round_to<T>(acc.w));
        }
        __syncwarp();
        float mean, rstd;
        row_stats(row, C, lane, eps, mean, rstd);
        for (int c = lane * 4; c < C; c += 128) {
            const float4 v = *reinterpret_cast<const float4*>(row + c);
            const float4 w4 = Elem<T>::ld4(lw + c), b4 = Elem<T>::ld4(lb + c);
            float4 o = make_float4(round_to<T>((v.x - mean) * rstd * w4.x + b4.x), round_to<T>((v.y - mean) * rstd * w4.y + b4.y),
                                   round_to<T>((v.z - mean) * rstd * w4.z + b4.z), round_to<T>((v.w - mean) * rstd * w4.w + b4.w));
            if (act == 1) o = make_float4(gelu_erf(o.x), gelu_erf(o.y), gelu_erf(o.z), gelu_erf(o.w));
            Elem<T>::st4(out + p * C + c, o);
        }
        __syncwarp();
    }
}

static unsigned ln_grid(long long rows) {
    const long long want = (rows + kLnWarps - 1) / kLnWarps;
    const long long cap = 148ll * 8;  // 8 CTAs of 8 warps per SM: the grid-stride loop takes the rest
    return (unsigned)(want < cap ? want : cap);
}

// ---- register-resident variants (channels <= 128 * QPL, QPL = 4-channel pieces per lane) ---------------------------------
// The row-buffer kernels above loop over a row's pieces with a run-time trip count, one dependent load at a time: latency
// bound (ncu on the InternImage-B forward: 25 % of the HBM rate).  Here the pieces of a pixel live in registers, the trip
// counts are compile-time constants, and every global load of a pixel is in flight before the first use.
template <int QPL>
__device__ __forceinline__ void reg_stats(const float4 (&v)[QPL], const bool (&on)[QPL], int C, float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < QPL; ++i)
        if (on[i]) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < QPL; ++i)
        if (on[i]) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    rstd = rsqrtf(warp_sum(q) / (float)C + eps);
}
template <typename T>
__device__ __forceinline__ float4 norm4(float4 v, float mean, float rstd, float4 w, float4 b) {
    return make_float4(round_to<T>((v.x - mean) * rstd * w.x + b.x), round_to<T>((v.y - mean) * rstd * w.y + b.y),
                       round_to<T>((v.z - mean) * rstd * w.z + b.z), round_to<T>((v.w - mean) * rstd * w.w + b.w));
}

template <typename T, int QPL, int R>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_join_reg_kernel(const T* __restrict__ y, const T* __restrict__ r, const T* __restrict__ gamma,
                   const T* __restrict__ lw, const T* __restrict__ lb, T* __restrict__ out_sum, T* __restrict__ out_norm,
                   long long rows, int C, float eps, int mode) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr bool CACHE = QPL <= 2;   // the lane's channels are the same for every row: few pieces -> parameters in registers
    constexpr int NC = CACHE ? QPL : 1;
    bool on[QPL];
    float4 cg[NC], cw[NC], cb[NC];
#pragma unroll
    for (int i = 0; i < QPL; ++i) on[i] = (lane + 32 * i) * 4 < C;
    auto param = [&](const T* base, int i, float fill) {
        return (on[i] && base) ? Elem<T>::ld4(base + (lane + 32 * i) * 4) : make_float4(fill, fill, fill, fill);
    };
    if (CACHE) {
#pragma unroll
        for (int i = 0; i < NC; ++i) { cg[i] = param(gamma, i, 1.f); cw[i] = param(lw, i, 1.f); cb[i] = param(lb, i, 0.f); }
    }
    // (parameters at use: from the registers above, or L1-resident loads)
    auto G4 = [&](int i) { return CACHE ? cg[i % NC] : param(gamma, i, 1.f); };
    auto W4 = [&](int i) { return CACHE ? cw[i % NC] : param(lw, i, 1.f); };
    auto B4 = [&](int i) { return CACHE ? cb[i % NC] : param(lb, i, 0.f); };
    const long long stride = (long long)gridDim.x * kLnWarps;
    // R rows per warp iteration, their loads all first (R = 1 is what is launched: 4 / 2 rows for the narrow 112- / 224-
    // channel rows cost 74 registers and measured 18-70 % slower on B200 than one row at full occupancy)
    for (long long p0 = (long long)blockIdx.x * kLnWarps + warp; p0 < rows; p0 += stride * R) {
        float4 a[R][QPL], b[R][QPL];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const long long p = p0 + k * stride;
#pragma unroll
            for (int i = 0; i < QPL; ++i) {
                const int c = (lane + 32 * i) * 4;
                const bool ld = on[i] && p < rows;
                a[k][i] = ld ? Elem<T>::ld4(y + p * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                b[k][i] = (ld && mode != 2) ? Elem<T>::ld4(r + p * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const long long p = p0 + k * stride;
            const bool live = p < rows;   // (warp-uniform; dead rows still take part in the shuffles)
            float mean, rstd;
            if (mode == 0) {
#pragma unroll
                for (int i = 0; i < QPL; ++i) {
                    const float4 g4 = G4(i);
                    a[k][i] = make_float4(round_to<T>(b[k][i].x + round_to<T>(a[k][i].x * g4.x)), round_to<T>(b[k][i].y + round_to<T>(a[k][i].y * g4.y)),
                                          round_to<T>(b[k][i].z + round_to<T>(a[k][i].z * g4.z)), round_to<T>(b[k][i].w + round_to<T>(a[k][i].w * g4.w)));
                    if (on[i] && live) Elem<T>::st4(out_sum + p * C + (lane + 32 * i) * 4, a[k][i]);
                }
                if (out_norm != nullptr) {
                    reg_stats<QPL>(a[k], on, C, eps, mean, rstd);
#pragma unroll
                    for (int i = 0; i < QPL; ++i)
                        if (on[i] && live) Elem<T>::st4(out_norm + p * C + (lane + 32 * i) * 4, norm4<T>(a[k][i], mean, rstd, W4(i), B4(i)));
                }
            } else {
                reg_stats<QPL>(a[k], on, C, eps, mean, rstd);
#pragma unroll
                for (int i = 0; i < QPL; ++i) {
                    float4 n4 = norm4<T>(a[k][i], mean, rstd, W4(i), B4(i));
                    if (mode == 1) {
                        const float4 g4 = G4(i);
                        n4 = make_float4(b[k][i].x + round_to<T>(n4.x * g4.x), b[k][i].y + round_to<T>(n4.y * g4.y),
                                         b[k][i].z + round_to<T>(n4.z * g4.z), b[k][i].w + round_to<T>(n4.w * g4.w));
                    }
                    if (on[i] && live) Elem<T>::st4(out_sum + p * C + (lane + 32 * i) * 4, n4);
                }
            }
        }
    }
}

// 3x3 depthwise convolution, weights [9][C] staged as fp32 in shared memory once per CTA (every pixel reuses them)
template <typename T, int QPL>
__global__ void __launch_bounds__(kLnWarps * 32)
dwconv3_ln_act_reg_kernel(const T* __restrict__ x, const T* __restrict__ wt, const T* __restrict__ bias,
                          const T* __restrict__ lw, const T* __restrict__ lb, T* __restrict__ out, int N, int H, int W, int C,
                          int pad_lo, float eps, int act, int PW) {
    extern __shared__ float swt[];   // [9][C]: fp32, or (bf16) the packed values themselves
    constexpr bool PACKED = sizeof(T) == 2;  // bf16 x bf16 straight into FHFMA.BF16: no unpacking of either operand
    if constexpr (PACKED) {
        for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) reinterpret_cast<T*>(swt)[i] = wt[i];
    } else {
        for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) swt[i] = Elem<T>::ld(wt + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool on[QPL];
#pragma unroll
    for (int i = 0; i < QPL; ++i) on[i] = (lane + 32 * i) * 4 < C;
    auto param = [&](const T* base, int i, float fill) {   // (L1-resident after the first pixel)
        return (on[i] && base) ? Elem<T>::ld4(base + (lane + 32 * i) * 4) : make_float4(fill, fill, fill, fill);
    };
    // A CTA works on a patch of 8 rows (one per warp) x 16 columns, column by column: the row above / below a warp's is
    // being read by the neighbouring warps at the same moment, and its own previous columns a moment ago, so eight of the
    // nine neighbour loads hit L1 (warp-per-pixel over scattered pixels made all nine go to L2: 8 % of the HBM rate)
    // (PW = 16 columns; narrower for small images so that every SM still gets patches)
    const int tiles_x = (W + PW - 1) / PW, tiles_y = (H + kLnWarps - 1) / kLnWarps;
    const long long ntiles = (long long)N * tiles_y * tiles_x;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int xi = 0; xi < PW; ++xi) {
        const int tx = (int)(tile % tiles_x), ty = (int)((tile / tiles_x) % tiles_y);
        const long long n = tile / ((long long)tiles_x * tiles_y);
        const int hq = ty * kLnWarps + warp, wq = tx * PW + xi;
        if (hq >= H || wq >= W) continue;   // (warp-uniform)
        const long long p = (n * H + hq) * W + wq;
        bool in[9];
        long long base[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {   // (warp-uniform border tests)
            const int yy = hq + t / 3 - pad_lo, xx = wq + t % 3 - pad_lo;
            in[t] = yy >= 0 && yy < H && xx >= 0 && xx < W;
            base[t] = ((n * H + yy) * W + xx) * C;
        }
        float4 acc[QPL];
#pragma unroll
        for (int i = 0; i < QPL; ++i) {
            const int c = (lane + 32 * i) * 4;
            if constexpr (PACKED) {
                uint2 v[9];
#pragma unroll
                for (int t = 0; t < 9; ++t)   // the nine loads of a piece first, then its arithmetic
                    v[t] = (in[t] && on[i]) ? __ldg(reinterpret_cast<const uint2*>(x + base[t] + c)) : make_uint2(0u, 0u);
                acc[i] = param(bias, i, 0.f);
#pragma unroll
                for (int t = 0; t < 9; ++t) {   // (the same products and the same order as the fp32 form below: exact either way)
                    const uint2 k2 = on[i] ? *reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(swt) + t * C + c) : make_uint2(0u, 0u);
                    fhfma_x<0, 0>(acc[i].x, v[t].x, k2.x); fhfma_x<1, 1>(acc[i].y, v[t].x, k2.x);
                    fhfma_x<0, 0>(acc[i].z, v[t].y, k2.y); fhfma_x<1, 1>(acc[i].w, v[t].y, k2.y);
                }
            } else {
                float4 v[9];
#pragma unroll
                for (int t = 0; t < 9; ++t)   // the nine loads of a piece first, then its arithmetic
                    v[t] = (in[t] && on[i]) ? Elem<T>::ld4(x + base[t] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                acc[i] = param(bias, i, 0.f);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float4 k4 = on[i] ? *reinterpret_cast<const float4*>(swt + t * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc[i].x = fmaf(v[t].x, k4.x, acc[i].x); acc[i].y = fmaf(v[t].y, k4.y, acc[i].y);
                    acc[i].z = fmaf(v[t].z, k4.z, acc[i].z); acc[i].w = fmaf(v[t].w, k4.w, acc[i].w);
                }
            }
            // (the unfused chain stores the convolution result in T before normalising it)
            acc[i] = make_float4(round_to<T>(acc[i].x), round_to<T>(acc[i].y), round_to<T>(acc[i].z), round_to<T>(acc[i].w));
        }
        float mean, rstd;
        reg_stats<QPL>(acc, on, C, eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < QPL; ++i) {
            float4 o = norm4<T>(acc[i], mean, rstd, param(lw, i, 1.f), param(lb, i, 0.f));
            if (act == 1) o = make_float4(gelu_erf(o.x), gelu_erf(o.y), gelu_erf(o.z), gelu_erf(o.w));
            if (on[i]) Elem<T>::st4(out + p * C + (lane + 32 * i) * 4, o);
        }
    }
}

template <typename T, int QPL, int R>
static cudaError_t launch_ln_join_reg(const void* y, const void* r, const void* gamma, const void* lw, const void* lb, void* out_sum,
                                      void* out_norm, long long rows, int C, float eps, int mode, cudaStream_t st) {
    ln_join_reg_kernel<T, QPL, R><<<ln_grid(rows), kLnWarps * 32, 0, st>>>((const T*)y, (const T*)r, (const T*)gamma, (const T*)lw,
                                                                        (const T*)lb, (T*)out_sum, (T*)out_norm, rows, C, eps, mode);
    return cudaGetLastError();
}
template <typename T>
static bool try_ln_join_reg(const void* y, const void* r, const void* gamma, const void* lw, const void* lb, void* out_sum,
                            void* out_norm, long long rows, int C, float eps, int mode, cudaStream_t st, cudaError_t& e) {
    if (C <= 128) e = launch_ln_join_reg<T, 1, 1>(y, r, gamma, lw, lb, out_sum, out_norm, rows, C, eps, mode, st);
    else if (C <= 256) e = launch_ln_join_reg<T, 2, 1>(y, r, gamma, lw, lb, out_sum, out_norm, rows, C, eps, mode, st);
    else if (C <= 512) e = launch_ln_join_reg<T, 4, 1>(y, r, gamma, lw, lb, out_sum, out_norm, rows, C, eps, mode, st);
    else if (C <= 1024) e = launch_ln_join_reg<T, 8, 1>(y, r, gamma, lw, lb, out_sum, out_norm, rows, C, eps, mode, st);
    else return false;
    return true;
}
template <typename T, int QPL>
static cudaError_t launch_dwconv3_reg(const void* x, const void* wt, const void* bias, const void* lw, const void* lb, void* out,
                                      int N, int H, int W, int C, int pad_lo, float eps, int act, cudaStream_t st) {
    const size_t smem = (size_t)9 * C * sizeof(float);
    int pw = 16;
    long long ntiles = 0;
    for (;; pw >>= 1) {   // at least four patches per SM, down to single columns
        ntiles = (long long)N * ((H + kLnWarps - 1) / kLnWarps) * ((W + pw - 1) / pw);
        if (ntiles >= 148ll * 4 || pw == 1) break;
    }
    const unsigned grid = (unsigned)(ntiles < 148ll * 8 ? ntiles : 148ll * 8);
    dwconv3_ln_act_reg_kernel<T, QPL><<<grid, kLnWarps * 32, smem, st>>>(
        (const T*)x, (const T*)wt, (const T*)bias, (const T*)lw, (const T*)lb, (T*)out, N, H, W, C, pad_lo, eps, act, pw);
    return cudaGetLastError();
}
template <typename T>
static bool try_dwconv3_reg(const void* x, const void* wt, const void* bias, const void* lw, const void* lb, void* out, int N, int H,
                            int W, int C, int k, int pad_lo, float eps, int act, cudaStream_t st, cudaError_t& e) {
    if (k != 3 || C > 1024) return false;   // (weights in shared memory: 36 KB at 1024 channels)
    if (C <= 128) e = launch_dwconv3_reg<T, 1>(x, wt, bias, lw, lb, out, N, H, W, C, pad_lo, eps, act, st);
    else if (C <= 256) e = launch_dwconv3_reg<T, 2>(x, wt, bias, lw, lb, out, N, H, W, C, pad_lo, eps, act, st);
    else if (C <= 512) e = launch_dwconv3_reg<T, 4>(x, wt, bias, lw, lb, out, N, H, W, C, pad_lo, eps, act, st);
    else e = launch_dwconv3_reg<T, 8>(x, wt, bias, lw, lb, out, N, H, W, C, pad_lo, eps, act, st);
    return true;
}

template <typename K>
static cudaError_t ensure_row_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return ensure_max_smem((const void*)kernel, (int)bytes);
}

cudaError_t launch_ln_join(const void* y, const void* r, const void* gamma, const void* lw, const void* lb, void* out_sum,
                           void* out_norm, long long rows, int C, float eps, int mode, int dtype, cudaStream_t st) {
    const size_t smem = (size_t)kLnWarps * C * sizeof(float);
    cudaError_t e;
    if (dtype == DCNV3_F32 ? try_ln_join_reg<float>(y, r, gamma, lw, lb, out_sum, out_norm, rows, C, eps, mode, st, e)
                           : try_ln_join_reg<__nv_bfloat16>(y, r, gamma, lw, lb, out_sum, out_norm, rows, C, eps, mode, st, e)) {
        count_launch(1);
        return e;
    }
    if (dtype == DCNV3_F32) {
        if ((e = ensure_row_smem(ln_join_kernel<float>, smem)) != cudaSuccess) return e;
        ln_join_kernel<float><<<ln_grid(rows), kLnWarps * 32, smem, st>>>(
            (const float*)y, (const float*)r, (const float*)gamma, (const float*)lw, (const float*)lb, (float*)out_sum,
            (float*)out_norm, rows, C, eps, mode);
    } else {
        using B = __nv_bfloat16;
        if ((e = ensure_row_smem(ln_join_kernel<B>, smem)) != cudaSuccess) return e;
        ln_join_kernel<B><<<ln_grid(rows), kLnWarps * 32, smem, st>>>((const B*)y, (const B*)r, (const B*)gamma, (const B*)lw,
                                                                     (const B*)lb, (B*)out_sum, (B*)out_norm, rows, C, eps, mode);
    }
    count_launch(1);
    return cudaGetLastError();
}

cudaError_t launch_dwconv_ln_act(const void* x, const void* wt, const void* bias, const void* lw, const void* lb, void* out,
                                 int N, int H, int W, int C, int k, int pad_lo, float eps, int act, int dtype, cudaStream_t st) {
    const size_t smem = (size_t)kLnWarps * C * sizeof(float);
    const long long rows = (long long)N * H * W;
    cudaError_t e;
    if (dtype == DCNV3_F32 ? try_dwconv3_reg<float>(x, wt, bias, lw, lb, out, N, H, W, C, k, pad_lo, eps, act, st, e)
                           : try_dwconv3_reg<__nv_bfloat16>(x, wt, bias, lw, lb, out, N, H, W, C, k, pad_lo, eps, act, st, e)) {
        count_launch(1);
        return e;
    }
    if (dtype == DCNV3_F32) {
        if ((e = ensure_row_smem(dwconv_ln_act_kernel<float>, smem)) != cudaSuccess) return e;
        dwconv_ln_act_kernel<float><<<ln_grid(rows), kLnWarps * 32, smem, st>>>(
            (const float*)x, (const float*)wt, (const float*)bias, (const float*)lw, (const float*)lb, (float*)out, N, H, W, C, k,
            pad_lo, eps, act);
    } else {
        using B = __nv_bfloat16;
        if ((e = ensure_row_smem(dwconv_ln_act_kernel<B>, smem)) != cudaSuccess) return e;
        dwconv_ln_act_kernel<B><<<ln_grid(rows), kLnWarps * 32, smem, st>>>((const B*)x, (const B*)wt, (const B*)bias, (const B*)lw,
                                                                           (const B*)lb, (B*)out, N, H, W, C, k, pad_lo, eps, act);
    }
    count_launch(1);
    return cudaGetLastError();
}

int max_ln_channels() { return kMaxLnChannels; }

}  // namespace dcnv3
