/*
 * dcnv3_b200 -- C ABI of the B200-native DCNv3 core operator (forward + backward).
 *
 * Drop-in boundary for edwardyehuang/iSeg's
 *     layers/dcn_v3/op.py:16   dcnv3_op(x, offset, mask, kernel_size, strides, padding,
 *                                       dilation_rate, groups, group_channels, offset_scale)
 * (its only caller: layers/dcn_v3/dcn_v3.py:125) and for the gradient TF autodiff derives from that
 * function (no source).  Tensor layouts are the reference's (op.py:89-107):
 *     x          [N, H,  W,  G*gc]                      NHWC, dense, row-major
 *     offset     [N, Ho, Wo, G*P*2]   element (g*P+p)*2+{0,1}; channel 0 is the "x"/W_in coordinate
 *     mask       [N, Ho, Wo, G*P]     already soft-maxed over P by the layer (dcn_v3.py:120-123),
 *                                     or raw logits when DCNV3_FLAG_MASK_LOGITS is set
 *     out        [N, Ho, Wo, G*gc]
 * P = kh*kw, tap p = i*kh + j with i the W-direction displacement (utils.py:77-101).
 * "SAME"/"VALID" (op.py:29-39) are resolved to pad_h/pad_w by the host wrapper.
 *
 * Conventions: every entry point returns 0 or a negative dcnv3_status; nothing throws across the
 * ABI; dcnv3_last_error() gives a thread-local message for the last failure on the calling thread.
 * No user-visible memory is allocated: outputs and the backward workspace are caller-owned device
 * buffers.  All work is enqueued on the caller's CUDA stream (cudaStream_t passed as void*); the
 * calls are asynchronous except the *_host variants.  The library is re-entrant: any number of host threads
 * may call it concurrently on their own streams / devices (tests/test_gpu_parity.py::test_two_threads_two_streams);
 * the only shared mutable state is the per-device scratch of the *_host entry points (one lock per device) and
 * the bench-only kernel timing switch, which is not meant to be flipped while other threads launch.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * DCNV3_ERR_CUDA.
 */
#ifndef DCNV3_B200_H_
#define DCNV3_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "dcnv3_dlpack.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DCNV3_ABI_VERSION 1

typedef enum {
    DCNV3_OK = 0,
    DCNV3_ERR_DTYPE = -1,      /* unsupported / mismatching element type */
    DCNV3_ERR_SHAPE = -2,      /* shapes inconsistent with the parameters (op.py:83 reshape) */
    DCNV3_ERR_LAYOUT = -3,     /* not dense row-major, or misaligned */
    DCNV3_ERR_DEVICE = -4,     /* not CUDA memory / tensors on different devices */
    DCNV3_ERR_CUDA = -5,       /* CUDA runtime error (no device, launch failure, ...) */
    DCNV3_ERR_WORKSPACE = -6,  /* backward workspace missing or too small */
    DCNV3_ERR_ARGUMENT = -7    /* NULL pointer, non-positive size, unsupported parameter */
} dcnv3_status;

typedef enum { DCNV3_F32 = 0, DCNV3_BF16 = 1 } dcnv3_dtype;

/* mask argument holds pre-softmax logits; softmax over P is fused (dcn_v3.py:120-123) and
   grad_mask is the gradient w.r.t. the logits */
#define DCNV3_FLAG_MASK_LOGITS 1u
/* backward only: the caller guarantees that the zero part of the workspace -- its first
   dcnv3_backward_workspace_zero_bytes() bytes -- is all-zero on entry (e.g. it was zeroed once and only ever
   used by dcnv3_backward, which leaves that part zeroed); saves a memset.  The rest of the workspace is
   scratch: never read before it is written, contents undefined afterwards.  A workspace that is reused for a
   call with a LARGER zero part must have the difference re-zeroed first (the scratch of the earlier call lay there). */
#define DCNV3_FLAG_WORKSPACE_ZEROED 4u
/* debugging aid, with DCNV3_FLAG_WORKSPACE_ZEROED: verify that promise before the call (one reduction kernel
   and a stream synchronisation; DCNV3_ERR_WORKSPACE if the workspace is dirty).  A launch failure or an aborted
   CUDA graph in the middle of an earlier backward leaves the workspace dirty; callers that keep a workspace
   should re-zero it after any failed call (iseg_b200/_cabi.py drops its cached one). */
#define DCNV3_FLAG_CHECK_WORKSPACE 16u
/* testing aid: bypass the shared-memory tiled kernels and run the generic kernels */
#define DCNV3_FLAG_FORCE_GENERIC 2u
/* bf16 tensors only (ignored for fp32): reproduce the arithmetic the reference performs under the
   mixed_bfloat16 policy -- reference points, grids, sampling locations, pixel coordinates, bilinear weights
   and the forward accumulation all rounded to bfloat16 after every primitive (op.py:62-87, utils.py:130-206).
   Default (flag clear): coordinates, weights and accumulation in fp32 on the bf16 inputs (the tiled forward rounds
   each finished corner weight -- bilinear weight x mask -- once to bf16 for its bf16 x bf16 -> fp32 product) --
   better numerics, but a different result, because bf16 coordinates near 130 have a step of 1 pixel.
   The flag runs the generic kernels. */
#define DCNV3_FLAG_REF_DTYPE 8u

typedef struct dcnv3_params {
    int32_t n, h, w;                 /* x is [n, h, w, groups*group_channels] */
    int32_t ho, wo;                  /* spatial size of offset / mask / out (op.py:51) */
    int32_t groups, group_channels;
    int32_t kh, kw;                  /* kernel_size  */
    int32_t sh, sw;                  /* strides      */
    int32_t ph, pw;                  /* zero padding added on each side (op.py:34-37,46) */
    int32_t dh, dw;                  /* dilation_rate */
    float offset_scale;
    int32_t dtype;                   /* dcnv3_dtype of all seven tensors */
    uint32_t flags;
} dcnv3_params;

int dcnv3_abi_version(void);
const char* dcnv3_last_error(void);
/* e.g. "dcnv3_b200 abi 1, sm_100a, nvcc 12.9" */
const char* dcnv3_build_info(void);

/* Validates a parameter block the way the reference's reshapes would. */
int dcnv3_check_params(const dcnv3_params* p);

/* Introspection (no GPU needed; tests/test_cabi_cpu.py): how a call with these parameters would be tiled.
   plan25[0]      1 = shared-memory tiled kernels, 0 = generic kernels (the rest is then 0)
   plan25[1..8]   forward:  tile rows, tile cols, box width, box height (cells of 128 B), halo x, halo y,
                  CTAs, dynamic shared memory bytes
   plan25[9..16]  grad_offset / grad_mask kernel: the same eight numbers
   plan25[17..24] grad_x scatter kernel: tile edge (input cells), ring below, ring above, box rows, CTAs,
                  dynamic shared memory bytes, threads per CTA, 1 if the side-buffer merge kernel is launched */
int dcnv3_launch_plan(const dcnv3_params* p, int* plan25);

/* ---- device-pointer entry points (replace op.py:16 and its autodiff gradient) ---- */
int dcnv3_forward(const void* x, const void* offset, const void* mask, void* out,
                  const dcnv3_params* p, void* cuda_stream);

/* bytes of caller-owned device workspace dcnv3_backward needs (256-byte aligned pointer), and how many of them,
   from the start, form the zero part (see DCNV3_FLAG_WORKSPACE_ZEROED); 0 for invalid parameters */
size_t dcnv3_backward_workspace_bytes(const dcnv3_params* p);
size_t dcnv3_backward_workspace_zero_bytes(const dcnv3_params* p);

/* grad_x / grad_offset / grad_mask are fully overwritten.  Bitwise reproducible run to run. */
int dcnv3_backward(const void* x, const void* offset, const void* mask, const void* grad_out,
                   void* grad_x, void* grad_offset, void* grad_mask, void* workspace,
                   size_t workspace_bytes, const dcnv3_params* p, void* cuda_stream);

/* ---- the op with the layer's centre-feature-scale blend fused in (replaces dcn_v3.py:138-146 around op.py:16):
        out = core * (1 - s) + x * s,  s = center_scale[n, h, w, g] broadcast over the group's channels
        (center_scale: [n, h, w, groups], dtype of x; no sigmoid, as in the reference).  The backward also returns
        grad_center_scale [n, h, w, groups]; grad_x includes the direct path grad_out * s.
        Available where the shared-memory tiled kernels run (dcnv3_blend_supported() == 1: 3x3, stride 1,
        dilation 1, SAME, 16 or 32 channels per group -- what InternImage-T/S/B/L and -H instantiate); DCNV3_ERR_ARGUMENT
        otherwise, and the caller applies the blend itself around dcnv3_forward / dcnv3_backward. ---- */
int dcnv3_blend_supported(const dcnv3_params* p);
int dcnv3_forward_blend(const void* x, const void* offset, const void* mask, const void* center_scale, void* out,
                        const dcnv3_params* p, void* cuda_stream);
int dcnv3_backward_blend(const void* x, const void* offset, const void* mask, const void* center_scale,
                         const void* grad_out, void* grad_x, void* grad_offset, void* grad_mask,
                         void* grad_center_scale, void* workspace, size_t workspace_bytes, const dcnv3_params* p,
                         void* cuda_stream);

/* ---- inference fast path of the layers around the op: each element-wise chain as one pass (no gradients) ----
   dcnv3_dwconv_ln_act: out = act(LayerNorm(DepthwiseConv2D(x) + bias)) on NHWC [n,h,w,c] -- the x1 branch of
        DeformableConvolutionV3.call (reference layers/dcn_v3/dcn_v3.py:115-117).  weight_kkc: [k*k][c], tap-major
        (Keras depthwise kernel [k,k,c,1] as stored); stride 1, zero padding pad_lo before and k-1-pad_lo after
        (Keras 'same': pad_lo = (k-1)/2); activation 0 = none, 1 = exact (erf) GELU.
   dcnv3_layer_join: the joins of InternImageLayer.call (reference backbones/intern_image/intern_image_layer.py:126-172)
        over [rows, channels]:
        mode 0: out_sum = residual + gamma * y; out_norm (may be NULL) = LayerNorm(out_sum)        (pre-norm, :161-170)
        mode 1: out_sum = residual + gamma * LayerNorm(y)                      (post-norm :127-138, res-post-norm :145-155)
        mode 2: out_sum = LayerNorm(y)
        gamma may be NULL (no layer scale).  All arrays have the activation dtype; channels % 4 == 0, <= 4096. */
int dcnv3_dwconv_ln_act(const void* x, const void* weight_kkc, const void* bias, const void* ln_weight, const void* ln_bias,
                        void* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t pad_lo, float eps,
                        int32_t activation, int32_t dtype, void* cuda_stream);
int dcnv3_layer_join(const void* y, const void* residual, const void* gamma, const void* ln_weight, const void* ln_bias,
                     void* out_sum, void* out_norm, int64_t rows, int32_t channels, float eps, int32_t mode, int32_t dtype,
                     void* cuda_stream);

/* ---- sibling gather op: the sampling + aggregation of the reference's deformable multi-head self-attention
        (layers/deformable_multihead_self_attention.py:102-175 _bilinear_sample, then :233-235), fused:
            out[n,h,w,hd,:] = sum_p attn[n,h,w,hd,p] * bilinear(value[n,:,:,hd,:], y[n,h,w,hd,p], x[n,h,w,hd,p])
        value / out / grad_value: [n,h,w,heads*head_channels]; y, x, attn and their gradients: [n,h,w,heads*points];
        absolute pixel coordinates, neighbour indices clamped to the image, weights from the fractional parts (that
        function's conventions).  The backward is bitwise reproducible (64-bit fixed-point integer atomics, scale per
        image); its workspace must be all-zero on entry and is left all-zero (DCNV3_FLAG_WORKSPACE_ZEROED as above). ---- */
size_t dcnv3_deform_attn_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t heads, int32_t head_channels);
int dcnv3_deform_attn_forward(const void* value, const void* y, const void* x, const void* attn, void* out, int32_t n,
                              int32_t h, int32_t w, int32_t heads, int32_t points, int32_t head_channels, int32_t dtype,
                              void* cuda_stream);
int dcnv3_deform_attn_backward(const void* value, const void* y, const void* x, const void* attn, const void* grad_out,
                               void* grad_value, void* grad_y, void* grad_x, void* grad_attn, void* workspace,
                               size_t workspace_bytes, int32_t n, int32_t h, int32_t w, int32_t heads, int32_t points,
                               int32_t head_channels, int32_t dtype, uint32_t flags, void* cuda_stream);

/* ---- sibling gather op: the sampling stage of the reference's DCNv2 (layers/dcn_v2.py:137-247, between the offset
        convolution and the contraction with the kernel):
            out[n,i,j,k,:] = mask[n,i,j,k] * bilinear(zero-padded x[n], i + ph + py_k + oy, j + pw + px_k + ox)
        x / grad_x: [n,h,w,channels]; offsets: [n,h,w,kh*kw*2] ((oy, ox) per tap, the first 2*kh*kw channels of the
        offset convolution, :144-146); mask: [n,h,w,kh*kw] (after the sigmoid, :148); out / grad_out: [n,h,w,kh*kw*channels]
        (the `map_all` of :247).  Tap order, clipping of indices and coordinate to [0, h+1] x [0, w+1] and the weights from
        the clipped values are that function's.  Odd kernels 3..15.  Deterministic backward, workspace as above. ---- */
size_t dcnv3_dcnv2_sample_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t channels);
int dcnv3_dcnv2_sample_forward(const void* x, const void* offsets, const void* mask, void* out, int32_t n, int32_t h, int32_t w,
                               int32_t channels, int32_t kh, int32_t kw, int32_t dtype, void* cuda_stream);
int dcnv3_dcnv2_sample_backward(const void* x, const void* offsets, const void* mask, const void* grad_out, void* grad_x,
                                void* grad_offsets, void* grad_mask, void* workspace, size_t workspace_bytes, int32_t n,
                                int32_t h, int32_t w, int32_t channels, int32_t kh, int32_t kw, int32_t dtype, uint32_t flags,
                                void* cuda_stream);

/* ---- DLPack entry points: same calls, tensors described by DLManagedTensor (zero copy);
        shapes, dtype, device and contiguity are taken from / checked against the tensors ---- */
int dcnv3_forward_dlpack(const DLManagedTensor* x, const DLManagedTensor* offset,
                         const DLManagedTensor* mask, DLManagedTensor* out, int kh, int kw, int sh,
                         int sw, int pad_h, int pad_w, int dh, int dw, int groups,
                         int group_channels, float offset_scale, unsigned flags, void* cuda_stream);

int dcnv3_backward_dlpack(const DLManagedTensor* x, const DLManagedTensor* offset,
                          const DLManagedTensor* mask, const DLManagedTensor* grad_out,
                          DLManagedTensor* grad_x, DLManagedTensor* grad_offset,
                          DLManagedTensor* grad_mask, DLManagedTensor* workspace, int kh, int kw,
                          int sh, int sw, int pad_h, int pad_w, int dh, int dw, int groups,
                          int group_channels, float offset_scale, unsigned flags,
                          void* cuda_stream);

/* ---- host-buffer entry points: what a CPU-tensor caller (e.g. the reference's TF CPU path) binds.
        Copies host -> device, runs, copies results back and synchronises.  Device scratch is cached
        per device inside the library.  Pinned host memory makes the copies asynchronous. ---- */
int dcnv3_forward_host(const void* x, const void* offset, const void* mask, void* out,
                       const dcnv3_params* p, int device);

int dcnv3_forward_backward_host(const void* x, const void* offset, const void* mask,
                                const void* grad_out, void* out, void* grad_x, void* grad_offset,
                                void* grad_mask, const dcnv3_params* p, int device);

/* Pipelined variant: enqueues copy-in (on the device's copy-in stream), kernels (on the stream of scratch
   slot `slot`, 0 <= slot < dcnv3_host_slots()) and copy-out (on the device's copy-out stream), chained by
   events, and returns without waiting, so that consecutive calls on different slots overlap H2D, compute
   and D2H at full PCIe rate in both directions.  A slot's scratch is reused as soon as its previous outputs
   have been copied out (the wait happens on the device).  Host buffers must be pinned and stay valid --
   inputs unmodified, outputs unread -- until dcnv3_host_sync(device) returns. */
int dcnv3_forward_backward_host_async(const void* x, const void* offset, const void* mask,
                                      const void* grad_out, void* out, void* grad_x, void* grad_offset,
                                      void* grad_mask, const dcnv3_params* p, int device, int slot);
int dcnv3_host_sync(int device);
int dcnv3_host_slots(void);

/* Releases the per-device scratch used by the *_host entry points. */
int dcnv3_release_host_scratch(void);

/* Measurement aid (bench.py): when enabled, the tiled dcnv3_backward brackets its four kernels
   (gather, scatter, redo, merge) with CUDA events on the launch stream; dcnv3_get_kernel_timing waits for
   the last backward and returns their durations in milliseconds. */
int dcnv3_set_kernel_timing(int enable);
int dcnv3_get_kernel_timing(float* ms4);

/* Number of kernels this library has launched on behalf of the calling process (monotonic). */
uint64_t dcnv3_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DCNV3_B200_H_ */
