// Internal launcher interface between the C ABI (dcnv3_cabi.cu) and the kernel translation units.
#pragma once

#include "dcnv3_common.cuh"

namespace dcnv3 {

void count_launch(unsigned n);

// optional per-kernel timing of the tiled backward (bench.py roofline): events recorded on the launch stream
struct KernelTiming {
    bool enabled = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // before gather, scatter, redo, merge, after
};
KernelTiming& kernel_timing();

// generic path (dcnv3_generic.cu)
cudaError_t launch_fwd_generic(const void* x, const void* offset, const void* mask, void* out,
                               const KParams& q, int dtype, cudaStream_t st);
size_t bwd_generic_workspace_bytes(const KParams& q);
cudaError_t launch_bwd_generic(const void* x, const void* offset, const void* mask,
                               const void* grad_out, void* grad_x, void* grad_offset, void* grad_mask,
                               void* ws, const KParams& q, int dtype, bool ws_clean, cudaStream_t st);

// tiled path (dcnv3_tiled_*.cu): k=3, s=1, d=1, SAME, 16 channels per group (or 32, run as half groups: tiled_view())
bool tiled_applicable(const KParams& q, int dtype);
cudaError_t launch_fwd_tiled(const void* x, const void* offset, const void* mask, void* out,
                             const KParams& q, int dtype, cudaStream_t st);

// launch plans as plain numbers (dcnv3_launch_plan: CPU-side tests of the tiling logic)
void fwd_tiled_plan(const KParams& q, int dtype, int out[8]);   // th tw bw bh halo_x halo_y grid smem_bytes
void bwd_tiled_plan(const KParams& q, int dtype, int out[16]);  // gather: as above; scatter: tj ring_lo ring_hi box_rows grid smem_bytes threads merge
// workspace of the tiled backward = [bwd_tiled_workspace_bytes: zero on entry, left zero on exit][scratch: no contract]
size_t bwd_tiled_workspace_bytes(const KParams& q);
size_t bwd_tiled_scratch_bytes(const KParams& q, int dtype);
cudaError_t launch_bwd_tiled(const void* x, const void* offset, const void* mask, const void* grad_out,
                             void* grad_x, void* grad_offset, void* grad_mask, void* ws, void* scratch, const KParams& q,
                             int dtype, bool ws_clean, cudaStream_t st);

// raises a kernel's dynamic shared-memory limit once per (kernel, device) (dcnv3_tiled_fwd.cu)
cudaError_t ensure_max_smem(const void* kernel, int bytes);

// layer fast path (layer_fused.cu): element-wise chains around the op, one pass each
int max_ln_channels();
cudaError_t launch_ln_join(const void* y, const void* r, const void* gamma, const void* lw, const void* lb, void* out_sum,
                           void* out_norm, long long rows, int C, float eps, int mode, int dtype, cudaStream_t st);
cudaError_t launch_dwconv_ln_act(const void* x, const void* wt, const void* bias, const void* lw, const void* lb, void* out,
                                 int N, int H, int W, int C, int k, int pad_lo, float eps, int act, int dtype, cudaStream_t st);

// sibling gather op (dcnv3_generic.cu): sampling + aggregation of iSeg's deformable multi-head self-attention
size_t deform_attn_workspace_bytes(int n, int h, int w, int heads, int c);
cudaError_t launch_deform_attn_fwd(const void* value, const void* ys, const void* xs, const void* attn, void* out, int n, int h,
                                   int w, int heads, int points, int c, int dtype, cudaStream_t st);
cudaError_t launch_deform_attn_bwd(const void* value, const void* ys, const void* xs, const void* attn, const void* grad_out,
                                   void* grad_value, void* grad_y, void* grad_x, void* grad_attn, void* ws, int n, int h, int w,
                                   int heads, int points, int c, int dtype, cudaStream_t st);

// sibling gather op (dcnv3_generic.cu): the sampling stage of iSeg's DCNv2
size_t dcnv2_sample_workspace_bytes(int n, int h, int w, int c);
cudaError_t launch_dcnv2_sample_fwd(const void* x, const void* offs, const void* mask, void* out, int n, int h, int w, int c,
                                    int kh, int kw, int dtype, cudaStream_t st);
cudaError_t launch_dcnv2_sample_bwd(const void* x, const void* offs, const void* mask, const void* grad_out, void* grad_x,
                                    void* grad_offs, void* grad_mask, void* ws, int n, int h, int w, int c, int kh, int kw,
                                    int dtype, cudaStream_t st);

}  // namespace dcnv3
