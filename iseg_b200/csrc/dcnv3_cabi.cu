// C ABI of dcnv3_b200 (include/dcnv3_b200.h): validation, parameter derivation, dispatch, DLPack and
// host-buffer front ends.  No kernels here.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "dcnv3_kernels.h"

namespace dcnv3 {

static std::atomic<unsigned long long> g_launches{0};
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char t_err[512] = "";

KernelTiming& kernel_timing() {
    static KernelTiming t;
    return t;
}

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

static int cuda_fail(cudaError_t e, const char* what) {
    return fail(DCNV3_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// the reference-dtype emulation exists for bf16 only (for fp32 the reference dtype IS the kernels' arithmetic)
static bool ref_dtype_mode(const dcnv3_params* p) { return p->dtype == DCNV3_BF16 && (p->flags & DCNV3_FLAG_REF_DTYPE); }

static size_t elem_size(int dtype) { return dtype == DCNV3_F32 ? 4 : 2; }

// The envelope the reference enforces through tf.reshape (op.py:83 with utils.py:26-27): the
// reference-point grid must have the offset's spatial size.
static int check(const dcnv3_params* p) {
    if (p == nullptr) return fail(DCNV3_ERR_ARGUMENT, "params is NULL");
    if (p->dtype != DCNV3_F32 && p->dtype != DCNV3_BF16)
        return fail(DCNV3_ERR_DTYPE, "dtype %d not supported (0=f32, 1=bf16)", p->dtype);
    if (p->n < 0 || p->h <= 0 || p->w <= 0 || p->groups <= 0 || p->group_channels <= 0)
        return fail(DCNV3_ERR_SHAPE, "non-positive tensor dimension");
    if (p->kh <= 0 || p->kw <= 0 || p->kh * p->kw > DCNV3_MAX_TAPS)
        return fail(DCNV3_ERR_ARGUMENT, "kernel_size %dx%d unsupported (kh*kw <= %d)", p->kh, p->kw,
                    DCNV3_MAX_TAPS);
    if (p->sh <= 0 || p->sw <= 0 || p->dh <= 0 || p->dw <= 0 || p->ph < 0 || p->pw < 0)
        return fail(DCNV3_ERR_ARGUMENT, "strides/dilation must be positive, padding non-negative");
    const long hin = (long)p->h + 2 * p->ph, win = (long)p->w + 2 * p->pw;
    const long eh = hin - ((long)p->dh * (p->kh - 1) + 1), ew = win - ((long)p->dw * (p->kw - 1) + 1);
    if (eh < 0 || ew < 0) return fail(DCNV3_ERR_SHAPE, "kernel extent exceeds the padded input");
    const long ho = eh / p->sh + 1, wo = ew / p->sw + 1;
    if (ho != p->ho || wo != p->wo)
        return fail(DCNV3_ERR_SHAPE,
                    "offset/mask spatial size %dx%d does not match the reference-point grid %ldx%ld "
                    "(op.py:83 reshape)", p->ho, p->wo, ho, wo);
    const double elems = (double)p->n * hin * win * p->groups * p->group_channels;
    const double offs = (double)p->n * p->ho * p->wo * p->groups * p->kh * p->kw * 2;
    if (elems > 2.0e18 || offs > 2.0e18) return fail(DCNV3_ERR_SHAPE, "tensor too large");
    return DCNV3_OK;
}

static KParams derive(const dcnv3_params* p) {
    KParams q;
    memset(&q, 0, sizeof(q));
    q.n = p->n; q.h = p->h; q.w = p->w; q.ho = p->ho; q.wo = p->wo;
    q.G = p->groups; q.gc = p->group_channels; q.P = p->kh * p->kw; q.kh = p->kh;
    q.sh = p->sh; q.sw = p->sw; q.ph = p->ph; q.pw = p->pw; q.dh = p->dh; q.dw = p->dw;
    q.hin = p->h + 2 * p->ph; q.win = p->w + 2 * p->pw;
    q.hin_f = (float)q.hin; q.win_f = (float)q.win;
    q.rhin_f = 1.0f / q.hin_f; q.rwin_f = 1.0f / q.win_f;
    q.hm2_f = (float)(q.hin - 2); q.wm2_f = (float)(q.win - 2);
    q.y0c = (float)((p->dh * (p->kh - 1)) / 2 + 0.5f);
    q.x0c = (float)((p->dw * (p->kw - 1)) / 2 + 0.5f);
    q.scale = p->offset_scale;
    // volatile: keep each operation a separately rounded fp32 operation on the host as well
    volatile float fx = q.wm2_f * q.scale; fx = fx / q.win_f;
    volatile float fy = q.hm2_f * q.scale; fy = fy / q.hin_f;
    q.fx = fx; q.fy = fy;
    q.flags = p->flags;
    for (int t = 0; t < q.P; ++t) {
        const int i = t / p->kh, j = t % p->kh;  // utils.py:77-101
        volatile float g0 = (float)(-((p->dw * (p->kw - 1)) / 2) + i * p->dw) / q.win_f;
        volatile float g1 = (float)(-((p->dh * (p->kh - 1)) / 2) + j * p->dh) / q.hin_f;
        volatile float a = g0 * q.scale, b = g1 * q.scale;
        q.gs0[t] = a; q.gs1[t] = b;
    }
    return q;
}

static int check_ptr_align(const void* ptr, const char* name, size_t align = 16) {
    if (ptr == nullptr) return fail(DCNV3_ERR_ARGUMENT, "%s is NULL", name);
    if (((uintptr_t)ptr) % align != 0) return fail(DCNV3_ERR_LAYOUT, "%s is not %zu-byte aligned", name, align);
    return 0;
}

// is the centre-feature-scale blend available for these parameters?  (fused into the tiled kernels only)
static bool blend_supported(const dcnv3_params* p) {
    const KParams q = derive(p);
    return tiled_applicable(q, p->dtype) && !(p->flags & DCNV3_FLAG_FORCE_GENERIC) && !ref_dtype_mode(p);
}

static int forward_impl(const void* x, const void* offset, const void* mask, void* out,
                        const dcnv3_params* p, cudaStream_t st, const void* cfs = nullptr, bool blend = false) {
    int rc = check(p);
    if (rc) return rc;
    if (p->n == 0) return DCNV3_OK;  // empty batch: nothing to do (data pointers may be NULL)
    if ((rc = check_ptr_align(x, "x")) || (rc = check_ptr_align(offset, "offset")) ||
        (rc = check_ptr_align(mask, "mask")) || (rc = check_ptr_align(out, "out")))
        return rc;
    KParams q = derive(p);
    if (blend) {
        if ((rc = check_ptr_align(cfs, "center_scale", 4))) return rc;
        if (!blend_supported(p))
            return fail(DCNV3_ERR_ARGUMENT, "the fused centre-feature-scale blend needs the tiled configuration "
                                            "(3x3, stride 1, dilation 1, SAME, 16 or 32 channels per group); see dcnv3_blend_supported");
        q.cfs = cfs;
    }
    const bool tiled = tiled_applicable(q, p->dtype) && !(p->flags & DCNV3_FLAG_FORCE_GENERIC) && !ref_dtype_mode(p);
    cudaError_t e = tiled ? launch_fwd_tiled(x, offset, mask, out, q, p->dtype, st)
                          : launch_fwd_generic(x, offset, mask, out, q, p->dtype, st);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_forward launch");
    return DCNV3_OK;
}

// Backward workspace = [zero part][scratch].  The zero part (fixed-point side buffer, dirty map, flags, per-image
// maxima) must be all-zero on entry and is left all-zero on exit; the scratch tail (the tiled path's transposed
// copy of offset / mask) carries no contract in either direction.
static size_t backward_ws_zero_bytes(const dcnv3_params* p) {
    const KParams q = derive(p);
    const size_t a = bwd_generic_workspace_bytes(q), b = bwd_tiled_workspace_bytes(q);
    return ((a > b ? a : b) + 255) / 256 * 256;
}
static size_t backward_ws_bytes(const dcnv3_params* p) {
    const KParams q = derive(p);
    return backward_ws_zero_bytes(p) + (tiled_applicable(q, p->dtype) ? bwd_tiled_scratch_bytes(q, p->dtype) : 0);
}

// DCNV3_FLAG_CHECK_WORKSPACE: is the workspace really all-zero?  (debug aid: one reduction kernel + a
// stream synchronisation)
__global__ void __launch_bounds__(256) workspace_nonzero_kernel(const uint4* __restrict__ ws, size_t n16, int* flag) {
    unsigned any = 0u;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = ws[i];
        any |= v.x | v.y | v.z | v.w;
    }
    if (__any_sync(0xffffffffu, any != 0u) && (threadIdx.x & 31) == 0) *flag = 1;
}

static int check_workspace_zero(const void* ws, size_t bytes, cudaStream_t st) {
    int* flag = nullptr;
    cudaError_t e = cudaMallocHost(&flag, sizeof(int));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
    *flag = 0;
    workspace_nonzero_kernel<<<148 * 4, 256, 0, st>>>((const uint4*)ws, bytes / 16, flag);
    e = cudaStreamSynchronize(st);
    const int dirty = *flag;
    cudaFreeHost(flag);
    if (e != cudaSuccess) return cuda_fail(e, "workspace check");
    if (dirty)
        return fail(DCNV3_ERR_WORKSPACE, "DCNV3_FLAG_WORKSPACE_ZEROED was given but the workspace is not all-zero "
                                         "(an earlier call failed or something else wrote to it)");
    return 0;
}

static int backward_impl(const void* x, const void* offset, const void* mask, const void* grad_out,
                         void* grad_x, void* grad_offset, void* grad_mask, void* ws, size_t ws_bytes,
                         const dcnv3_params* p, cudaStream_t st, const void* cfs = nullptr, void* grad_cfs = nullptr,
                         bool blend = false) {
    int rc = check(p);
    if (rc) return rc;
    if (p->n == 0) return DCNV3_OK;
    if ((rc = check_ptr_align(x, "x")) || (rc = check_ptr_align(offset, "offset")) ||
        (rc = check_ptr_align(mask, "mask")) || (rc = check_ptr_align(grad_out, "grad_out")) ||
        (rc = check_ptr_align(grad_x, "grad_x")) || (rc = check_ptr_align(grad_offset, "grad_offset")) ||
        (rc = check_ptr_align(grad_mask, "grad_mask")))
        return rc;
    const size_t need = backward_ws_bytes(p);
    if (ws == nullptr || ws_bytes < need)
        return fail(DCNV3_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, ws_bytes);
    if ((rc = check_ptr_align(ws, "workspace", 256))) return rc;
    const size_t zero_bytes = backward_ws_zero_bytes(p);
    if ((p->flags & DCNV3_FLAG_CHECK_WORKSPACE) && (p->flags & DCNV3_FLAG_WORKSPACE_ZEROED) &&
        (rc = check_workspace_zero(ws, zero_bytes, st)))
        return rc;
    KParams q = derive(p);
    if (blend) {
        if ((rc = check_ptr_align(cfs, "center_scale", 4)) || (rc = check_ptr_align(grad_cfs, "grad_center_scale", 4))) return rc;
        if (!blend_supported(p))
            return fail(DCNV3_ERR_ARGUMENT, "the fused centre-feature-scale blend needs the tiled configuration "
                                            "(3x3, stride 1, dilation 1, SAME, 16 or 32 channels per group); see dcnv3_blend_supported");
        q.cfs = cfs;
        q.grad_cfs = grad_cfs;
    }
    const bool tiled = tiled_applicable(q, p->dtype) && !(p->flags & DCNV3_FLAG_FORCE_GENERIC) && !ref_dtype_mode(p);
    cudaError_t e;
    // the whole zero part, whichever path runs: both leave it zero, so the caller's next call may be either
    if (!(p->flags & DCNV3_FLAG_WORKSPACE_ZEROED) && (e = cudaMemsetAsync(ws, 0, zero_bytes, st)) != cudaSuccess)
        return cuda_fail(e, "memset(workspace)");
    e = tiled ? launch_bwd_tiled(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, (char*)ws + zero_bytes, q,
                                 p->dtype, true, st)
              : launch_bwd_generic(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, ws, q, p->dtype, true, st);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_backward launch");
    return DCNV3_OK;
}

// ---- DLPack helpers ------------------------------------------------------------------------------
static int dl_dtype(const DLTensor& t, int* dtype, const char* name) {
    if (t.dtype.lanes != 1) return fail(DCNV3_ERR_DTYPE, "%s: vector lanes unsupported", name);
    if (t.dtype.code == kDLFloat && t.dtype.bits == 32) { *dtype = DCNV3_F32; return 0; }
    if (t.dtype.code == kDLBfloat && t.dtype.bits == 16) { *dtype = DCNV3_BF16; return 0; }
    return fail(DCNV3_ERR_DTYPE, "%s: dtype code %d bits %d unsupported (float32 / bfloat16)", name,
                (int)t.dtype.code, (int)t.dtype.bits);
}

static int dl_check(const DLManagedTensor* m, const char* name, int ndim, int device_id, int dtype,
                    void** data) {
    if (m == nullptr) return fail(DCNV3_ERR_ARGUMENT, "%s is NULL", name);
    const DLTensor& t = m->dl_tensor;
    if (t.device.device_type != kDLCUDA && t.device.device_type != kDLCUDAManaged)
        return fail(DCNV3_ERR_DEVICE, "%s is not a CUDA tensor (device_type %d)", name, t.device.device_type);
    if (device_id >= 0 && t.device.device_id != device_id)
        return fail(DCNV3_ERR_DEVICE, "%s is on device %d, expected %d", name, t.device.device_id, device_id);
    if (t.ndim != ndim) return fail(DCNV3_ERR_SHAPE, "%s has %d dims, expected %d", name, t.ndim, ndim);
    int dt;
    int rc = dl_dtype(t, &dt, name);
    if (rc) return rc;
    if (dtype >= 0 && dt != dtype) return fail(DCNV3_ERR_DTYPE, "%s dtype differs from x", name);
    if (t.strides != nullptr) {  // must be compact row-major (size-1 dims may carry any stride)
        int64_t expect = 1;
        for (int i = t.ndim - 1; i >= 0; --i) {
            if (t.shape[i] != 1 && t.strides[i] != expect)
                return fail(DCNV3_ERR_LAYOUT, "%s is not dense row-major (NHWC contiguous)", name);
            expect *= t.shape[i];
        }
    }
    *data = (char*)t.data + t.byte_offset;
    return 0;
}

static int dl_expect_shape(const DLManagedTensor* m, const char* name, int64_t a, int64_t b, int64_t c,
                           int64_t d) {
    const int64_t* s = m->dl_tensor.shape;
    if (s[0] != a || s[1] != b || s[2] != c || s[3] != d)
        return fail(DCNV3_ERR_SHAPE, "%s has shape [%lld,%lld,%lld,%lld], expected [%lld,%lld,%lld,%lld]",
                    name, (long long)s[0], (long long)s[1], (long long)s[2], (long long)s[3],
                    (long long)a, (long long)b, (long long)c, (long long)d);
    return 0;
}

static int dl_params(const DLManagedTensor* x, const DLManagedTensor* offset, int kh, int kw, int sh,
                     int sw, int ph, int pw, int dh, int dw, int groups, int gc, float scale,
                     unsigned flags, dcnv3_params* p) {
    int dtype;
    int rc = dl_dtype(x->dl_tensor, &dtype, "x");
    if (rc) return rc;
    const int64_t* xs = x->dl_tensor.shape;
    const int64_t* os = offset->dl_tensor.shape;
    for (int i = 0; i < 4; ++i)
        if (xs[i] > 0x7fffffff || os[i] > 0x7fffffff) return fail(DCNV3_ERR_SHAPE, "dimension too large");
    p->n = (int)xs[0]; p->h = (int)xs[1]; p->w = (int)xs[2];
    p->ho = (int)os[1]; p->wo = (int)os[2];
    p->groups = groups; p->group_channels = gc;
    p->kh = kh; p->kw = kw; p->sh = sh; p->sw = sw; p->ph = ph; p->pw = pw; p->dh = dh; p->dw = dw;
    p->offset_scale = scale; p->dtype = dtype; p->flags = flags;
    return 0;
}

// ---- host-buffer scratch -------------------------------------------------------------------------
// Per device: one copy-in stream and one copy-out stream, so that each PCIe direction carries one copy at a
// time at full rate; kHostSlots slots, each with its own compute stream, data scratch and a backward
// workspace that is zeroed when (re)allocated and kept zeroed by dcnv3_backward itself.  Events chain
// copy-in -> kernels -> copy-out per slot; nothing blocks the host until dcnv3_host_sync.
constexpr int kHostSlots = 4;
struct HostScratch {
    void* buf = nullptr;
    size_t bytes = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    size_t ws_clean = 0;                            // bytes [0, ws_clean) are known to be zero (a call leaves its zero part zero
                                                    // and may write anything beyond it)
    cudaStream_t stream = nullptr;                  // kernels of this slot
    // forward inputs resident / grad_out resident / forward done / backward done / outputs copied
    cudaEvent_t ev_in = nullptr, ev_go = nullptr, ev_fwd = nullptr, ev_comp = nullptr, ev_out = nullptr;
};
struct DevicePipes {
    cudaStream_t in = nullptr, out = nullptr;
};
// one lock per device: host threads driving different GPUs (one thread per replica, as under the reference's
// MirroredStrategy) never wait for each other; calls on the same device are enqueued one after the other
static std::mutex g_scratch_mu[64];
static HostScratch g_scratch[64][kHostSlots];
static DevicePipes g_pipes[64];

static int scratch_reserve(int device, int slot, size_t bytes, size_t ws_bytes, HostScratch** out) {
    if (device < 0 || device >= 64) return fail(DCNV3_ERR_DEVICE, "device %d out of range", device);
    if (slot < 0 || slot >= kHostSlots) return fail(DCNV3_ERR_ARGUMENT, "slot %d out of range [0,%d)", slot, kHostSlots);
    HostScratch& s = g_scratch[device][slot];
    DevicePipes& dp = g_pipes[device];
    cudaError_t e;
    if (dp.in == nullptr) {
        if ((e = cudaStreamCreateWithFlags(&dp.in, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&dp.out, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(e, "cudaStreamCreate");
    }
    if (s.stream == nullptr) {
        if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(e, "cudaStreamCreate");
        if ((e = cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&s.ev_go, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&s.ev_fwd, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming)) != cudaSuccess)
            return cuda_fail(e, "cudaEventCreate");
    }
    if (s.bytes < bytes || s.ws_bytes < ws_bytes) {
        // the slot may still be running an earlier asynchronous call (its copy-out comes last)
        if ((e = cudaEventSynchronize(s.ev_out)) != cudaSuccess) return cuda_fail(e, "event synchronize");
        if ((e = cudaStreamSynchronize(s.stream)) != cudaSuccess) return cuda_fail(e, "stream synchronize");
    }
    if (s.bytes < bytes) {
        if (s.buf) cudaFree(s.buf);
        s.buf = nullptr; s.bytes = 0;
        if ((e = cudaMalloc(&s.buf, bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
        s.bytes = bytes;
    }
    if (s.ws_bytes < ws_bytes) {
        if (s.ws) cudaFree(s.ws);
        s.ws = nullptr; s.ws_bytes = 0;
        if ((e = cudaMalloc(&s.ws, ws_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)");
        if ((e = cudaMemsetAsync(s.ws, 0, ws_bytes, s.stream)) != cudaSuccess) return cuda_fail(e, "memset(workspace)");
        s.ws_bytes = ws_bytes;
        s.ws_clean = ws_bytes;
    }
    *out = &s;
    return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace dcnv3

using namespace dcnv3;

extern "C" {

int dcnv3_abi_version(void) { return DCNV3_ABI_VERSION; }

const char* dcnv3_last_error(void) { return t_err; }

const char* dcnv3_build_info(void) {
    static char info[128];
    snprintf(info, sizeof(info), "dcnv3_b200 abi %d, sm_100a, nvcc %d.%d", DCNV3_ABI_VERSION,
             __CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__);
    return info;
}

int dcnv3_check_params(const dcnv3_params* p) { return check(p); }

int dcnv3_launch_plan(const dcnv3_params* p, int* plan25) {
    int rc = check(p);
    if (rc) return rc;
    if (plan25 == nullptr) return fail(DCNV3_ERR_ARGUMENT, "NULL plan buffer");
    const KParams q = derive(p);
    for (int i = 0; i < 25; ++i) plan25[i] = 0;
    plan25[0] = tiled_applicable(q, p->dtype) && !(p->flags & DCNV3_FLAG_FORCE_GENERIC) && !ref_dtype_mode(p);
    if (plan25[0]) {
        fwd_tiled_plan(q, p->dtype, plan25 + 1);
        bwd_tiled_plan(q, p->dtype, plan25 + 9);
    }
    return DCNV3_OK;
}

int dcnv3_forward(const void* x, const void* offset, const void* mask, void* out,
                  const dcnv3_params* p, void* cuda_stream) {
    return forward_impl(x, offset, mask, out, p, (cudaStream_t)cuda_stream);
}

int dcnv3_blend_supported(const dcnv3_params* p) {
    if (check(p) != DCNV3_OK) return 0;
    return blend_supported(p) ? 1 : 0;
}

int dcnv3_forward_blend(const void* x, const void* offset, const void* mask, const void* center_scale, void* out,
                        const dcnv3_params* p, void* cuda_stream) {
    return forward_impl(x, offset, mask, out, p, (cudaStream_t)cuda_stream, center_scale, true);
}

int dcnv3_backward_blend(const void* x, const void* offset, const void* mask, const void* center_scale,
                         const void* grad_out, void* grad_x, void* grad_offset, void* grad_mask,
                         void* grad_center_scale, void* workspace, size_t workspace_bytes, const dcnv3_params* p,
                         void* cuda_stream) {
    return backward_impl(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, workspace, workspace_bytes, p,
                         (cudaStream_t)cuda_stream, center_scale, grad_center_scale, true);
}

int dcnv3_layer_join(const void* y, const void* residual, const void* gamma, const void* ln_weight, const void* ln_bias,
                     void* out_sum, void* out_norm, int64_t rows, int32_t channels, float eps, int32_t mode, int32_t dtype,
                     void* cuda_stream) {
    if (dtype != DCNV3_F32 && dtype != DCNV3_BF16) return fail(DCNV3_ERR_DTYPE, "dtype %d not supported", dtype);
    if (rows < 0 || channels <= 0 || channels % 4 != 0 || channels > max_ln_channels())
        return fail(DCNV3_ERR_SHAPE, "dcnv3_layer_join: %lld rows x %d channels (channels must be a multiple of 4, <= %d)",
                    (long long)rows, channels, max_ln_channels());
    if (mode < 0 || mode > 2) return fail(DCNV3_ERR_ARGUMENT, "dcnv3_layer_join: mode %d", mode);
    if (rows == 0) return DCNV3_OK;
    int rc;
    const size_t al = dtype == DCNV3_F32 ? 16 : 8;
    if ((rc = check_ptr_align(y, "y", al)) || (rc = check_ptr_align(out_sum, "out_sum", al))) return rc;
    if (mode != 2 && (rc = check_ptr_align(residual, "residual", al))) return rc;
    if ((mode != 0 || out_norm != nullptr) && ((rc = check_ptr_align(ln_weight, "ln_weight", al)) || (rc = check_ptr_align(ln_bias, "ln_bias", al))))
        return rc;
    if (gamma != nullptr && (rc = check_ptr_align(gamma, "gamma", al))) return rc;
    if (out_norm != nullptr && (rc = check_ptr_align(out_norm, "out_norm", al))) return rc;
    const cudaError_t e = launch_ln_join(y, residual, gamma, ln_weight, ln_bias, out_sum, out_norm, rows, channels, eps, mode, dtype,
                                         (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_layer_join launch");
    return DCNV3_OK;
}

int dcnv3_dwconv_ln_act(const void* x, const void* weight_kkc, const void* bias, const void* ln_weight, const void* ln_bias,
                        void* out, int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t pad_lo, float eps,
                        int32_t activation, int32_t dtype, void* cuda_stream) {
    if (dtype != DCNV3_F32 && dtype != DCNV3_BF16) return fail(DCNV3_ERR_DTYPE, "dtype %d not supported", dtype);
    if (n < 0 || h <= 0 || w <= 0 || c <= 0 || c % 4 != 0 || c > max_ln_channels() || k <= 0 || k > 15 || pad_lo < 0 || pad_lo >= k)
        return fail(DCNV3_ERR_SHAPE, "dcnv3_dwconv_ln_act: [%d,%d,%d,%d], kernel %d, pad %d (channels: multiple of 4, <= %d)", n, h, w,
                    c, k, pad_lo, max_ln_channels());
    if (activation != 0 && activation != 1) return fail(DCNV3_ERR_ARGUMENT, "activation %d (0 none, 1 gelu)", activation);
    if (n == 0) return DCNV3_OK;
    int rc;
    const size_t al = dtype == DCNV3_F32 ? 16 : 8;
    if ((rc = check_ptr_align(x, "x", al)) || (rc = check_ptr_align(weight_kkc, "weight", al)) || (rc = check_ptr_align(out, "out", al)) ||
        (rc = check_ptr_align(ln_weight, "ln_weight", al)) || (rc = check_ptr_align(ln_bias, "ln_bias", al)))
        return rc;
    if (bias != nullptr && (rc = check_ptr_align(bias, "bias", al))) return rc;
    const cudaError_t e = launch_dwconv_ln_act(x, weight_kkc, bias, ln_weight, ln_bias, out, n, h, w, c, k, pad_lo, eps, activation,
                                               dtype, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_dwconv_ln_act launch");
    return DCNV3_OK;
}

static int check_deform_attn(int32_t n, int32_t h, int32_t w, int32_t heads, int32_t points, int32_t c, int32_t dtype) {
    if (dtype != DCNV3_F32 && dtype != DCNV3_BF16) return fail(DCNV3_ERR_DTYPE, "dtype %d not supported", dtype);
    if (n < 0 || h <= 0 || w <= 0 || heads <= 0 || points <= 0 || c <= 0)
        return fail(DCNV3_ERR_SHAPE, "deform_attn: value [%d,%d,%d,%d*%d], %d points", n, h, w, heads, c, points);
    if ((long long)h * w * heads * ((long long)c > points ? c : points) >= (1ll << 40))
        return fail(DCNV3_ERR_SHAPE, "deform_attn: image too large");
    return 0;
}

size_t dcnv3_deform_attn_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t heads, int32_t head_channels) {
    if (n < 0 || h <= 0 || w <= 0 || heads <= 0 || head_channels <= 0) return 0;
    return (deform_attn_workspace_bytes(n, h, w, heads, head_channels) + 255) / 256 * 256;
}

int dcnv3_deform_attn_forward(const void* value, const void* y, const void* x, const void* attn, void* out, int32_t n,
                              int32_t h, int32_t w, int32_t heads, int32_t points, int32_t head_channels, int32_t dtype,
                              void* cuda_stream) {
    int rc = check_deform_attn(n, h, w, heads, points, head_channels, dtype);
    if (rc) return rc;
    if (n == 0) return DCNV3_OK;
    const size_t al = (head_channels % 4 == 0) ? (dtype == DCNV3_F32 ? 16 : 8) : (dtype == DCNV3_F32 ? 4 : 2);
    if ((rc = check_ptr_align(value, "value", al)) || (rc = check_ptr_align(out, "out", al)) || (rc = check_ptr_align(y, "y", 2)) ||
        (rc = check_ptr_align(x, "x", 2)) || (rc = check_ptr_align(attn, "attn", 2)))
        return rc;
    const cudaError_t e = launch_deform_attn_fwd(value, y, x, attn, out, n, h, w, heads, points, head_channels, dtype,
                                                 (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_deform_attn_forward launch");
    return DCNV3_OK;
}

int dcnv3_deform_attn_backward(const void* value, const void* y, const void* x, const void* attn, const void* grad_out,
                               void* grad_value, void* grad_y, void* grad_x, void* grad_attn, void* workspace,
                               size_t workspace_bytes, int32_t n, int32_t h, int32_t w, int32_t heads, int32_t points,
                               int32_t head_channels, int32_t dtype, uint32_t flags, void* cuda_stream) {
    int rc = check_deform_attn(n, h, w, heads, points, head_channels, dtype);
    if (rc) return rc;
    if (n == 0) return DCNV3_OK;
    if ((rc = check_ptr_align(value, "value", 2)) || (rc = check_ptr_align(y, "y", 2)) || (rc = check_ptr_align(x, "x", 2)) ||
        (rc = check_ptr_align(attn, "attn", 2)) || (rc = check_ptr_align(grad_out, "grad_out", 2)) ||
        (rc = check_ptr_align(grad_value, "grad_value", 2)) || (rc = check_ptr_align(grad_y, "grad_y", 2)) ||
        (rc = check_ptr_align(grad_x, "grad_x", 2)) || (rc = check_ptr_align(grad_attn, "grad_attn", 2)))
        return rc;
    const size_t need = dcnv3_deform_attn_workspace_bytes(n, h, w, heads, head_channels);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(DCNV3_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    if ((rc = check_ptr_align(workspace, "workspace", 256))) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaError_t e;
    if (!(flags & DCNV3_FLAG_WORKSPACE_ZEROED) && (e = cudaMemsetAsync(workspace, 0, need, st)) != cudaSuccess)
        return cuda_fail(e, "memset(workspace)");
    e = launch_deform_attn_bwd(value, y, x, attn, grad_out, grad_value, grad_y, grad_x, grad_attn, workspace, n, h, w, heads, points,
                               head_channels, dtype, st);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_deform_attn_backward launch");
    return DCNV3_OK;
}

static int check_dcnv2(int32_t n, int32_t h, int32_t w, int32_t c, int32_t kh, int32_t kw, int32_t dtype) {
    if (dtype != DCNV3_F32 && dtype != DCNV3_BF16) return fail(DCNV3_ERR_DTYPE, "dtype %d not supported", dtype);
    if (n < 0 || h <= 0 || w <= 0 || c <= 0 || kh < 3 || kw < 3 || kh % 2 == 0 || kw % 2 == 0 || kh > 15 || kw > 15)
        return fail(DCNV3_ERR_SHAPE, "dcnv2_sample: x [%d,%d,%d,%d], kernel %dx%d (odd, 3..15: the clip range of "
                                     "layers/dcn_v2.py:165-177 lies inside the padded image only then)", n, h, w, c, kh, kw);
    if ((long long)h * w * c * kh * kw >= (1ll << 40)) return fail(DCNV3_ERR_SHAPE, "dcnv2_sample: image too large");
    return 0;
}

size_t dcnv3_dcnv2_sample_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t channels) {
    if (n < 0 || h <= 0 || w <= 0 || channels <= 0) return 0;
    return (dcnv2_sample_workspace_bytes(n, h, w, channels) + 255) / 256 * 256;
}

int dcnv3_dcnv2_sample_forward(const void* x, const void* offsets, const void* mask, void* out, int32_t n, int32_t h, int32_t w,
                               int32_t channels, int32_t kh, int32_t kw, int32_t dtype, void* cuda_stream) {
    int rc = check_dcnv2(n, h, w, channels, kh, kw, dtype);
    if (rc) return rc;
    if (n == 0) return DCNV3_OK;
    if ((rc = check_ptr_align(x, "x", 2)) || (rc = check_ptr_align(offsets, "offsets", 2)) || (rc = check_ptr_align(mask, "mask", 2)) ||
        (rc = check_ptr_align(out, "out", 2)))
        return rc;
    const cudaError_t e = launch_dcnv2_sample_fwd(x, offsets, mask, out, n, h, w, channels, kh, kw, dtype, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_dcnv2_sample_forward launch");
    return DCNV3_OK;
}

int dcnv3_dcnv2_sample_backward(const void* x, const void* offsets, const void* mask, const void* grad_out, void* grad_x,
                                void* grad_offsets, void* grad_mask, void* workspace, size_t workspace_bytes, int32_t n,
                                int32_t h, int32_t w, int32_t channels, int32_t kh, int32_t kw, int32_t dtype, uint32_t flags,
                                void* cuda_stream) {
    int rc = check_dcnv2(n, h, w, channels, kh, kw, dtype);
    if (rc) return rc;
    if (n == 0) return DCNV3_OK;
    if ((rc = check_ptr_align(x, "x", 2)) || (rc = check_ptr_align(offsets, "offsets", 2)) || (rc = check_ptr_align(mask, "mask", 2)) ||
        (rc = check_ptr_align(grad_out, "grad_out", 2)) || (rc = check_ptr_align(grad_x, "grad_x", 2)) ||
        (rc = check_ptr_align(grad_offsets, "grad_offsets", 2)) || (rc = check_ptr_align(grad_mask, "grad_mask", 2)))
        return rc;
    const size_t need = dcnv3_dcnv2_sample_workspace_bytes(n, h, w, channels);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(DCNV3_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    if ((rc = check_ptr_align(workspace, "workspace", 256))) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaError_t e;
    if (!(flags & DCNV3_FLAG_WORKSPACE_ZEROED) && (e = cudaMemsetAsync(workspace, 0, need, st)) != cudaSuccess)
        return cuda_fail(e, "memset(workspace)");
    e = launch_dcnv2_sample_bwd(x, offsets, mask, grad_out, grad_x, grad_offsets, grad_mask, workspace, n, h, w, channels, kh, kw,
                                dtype, st);
    if (e != cudaSuccess) return cuda_fail(e, "dcnv3_dcnv2_sample_backward launch");
    return DCNV3_OK;
}

size_t dcnv3_backward_workspace_bytes(const dcnv3_params* p) {
    if (check(p) != DCNV3_OK) return 0;
    return backward_ws_bytes(p);
}

size_t dcnv3_backward_workspace_zero_bytes(const dcnv3_params* p) {
    if (check(p) != DCNV3_OK) return 0;
    return backward_ws_zero_bytes(p);
}

int dcnv3_backward(const void* x, const void* offset, const void* mask, const void* grad_out,
                   void* grad_x, void* grad_offset, void* grad_mask, void* workspace,
                   size_t workspace_bytes, const dcnv3_params* p, void* cuda_stream) {
    return backward_impl(x, offset, mask, grad_out, grad_x, grad_offset, grad_mask, workspace,
                         workspace_bytes, p, (cudaStream_t)cuda_stream);
}

int dcnv3_forward_dlpack(const DLManagedTensor* x, const DLManagedTensor* offset,
                         const DLManagedTensor* mask, DLManagedTensor* out, int kh, int kw, int sh,
                         int sw, int pad_h, int pad_w, int dh, int dw, int groups,
                         int group_channels, float offset_scale, unsigned flags, void* cuda_stream) {
    void *px, *po, *pm, *pout;
    int rc;
    if ((rc = dl_check(x, "x", 4, -1, -1, &px))) return rc;
    const int dev = x->dl_tensor.device.device_id;
    int dtype;
    dl_dtype(x->dl_tensor, &dtype, "x");
    if ((rc = dl_check(offset, "offset", 4, dev, dtype, &po)) || (rc = dl_check(mask, "mask", 4, dev, dtype, &pm)) ||
        (rc = dl_check(out, "out", 4, dev, dtype, &pout)))
        return rc;
    dcnv3_params p;
    if ((rc = dl_params(x, offset, kh, kw, sh, sw, pad_h, pad_w, dh, dw, groups, group_channels,
                        offset_scale, flags, &p)))
        return rc;
    const int64_t C = (int64_t)groups * group_channels, GP = (int64_t)groups * kh * kw;
    if ((rc = dl_expect_shape(x, "x", p.n, p.h, p.w, C)) ||
        (rc = dl_expect_shape(offset, "offset", p.n, p.ho, p.wo, GP * 2)) ||
        (rc = dl_expect_shape(mask, "mask", p.n, p.ho, p.wo, GP)) ||
        (rc = dl_expect_shape(out, "out", p.n, p.ho, p.wo, C)))
        return rc;
    return forward_impl(px, po, pm, pout, &p, (cudaStream_t)cuda_stream);
}

int dcnv3_backward_dlpack(const DLManagedTensor* x, const DLManagedTensor* offset,
                          const DLManagedTensor* mask, const DLManagedTensor* grad_out,
                          DLManagedTensor* grad_x, DLManagedTensor* grad_offset,
                          DLManagedTensor* grad_mask, DLManagedTensor* workspace, int kh, int kw,
                          int sh, int sw, int pad_h, int pad_w, int dh, int dw, int groups,
                          int group_channels, float offset_scale, unsigned flags,
                          void* cuda_stream) {
    void *px, *po, *pm, *pgo, *pgx, *pgoff, *pgm;
    int rc;
    if ((rc = dl_check(x, "x", 4, -1, -1, &px))) return rc;
    const int dev = x->dl_tensor.device.device_id;
    int dtype;
    dl_dtype(x->dl_tensor, &dtype, "x");
    if ((rc = dl_check(offset, "offset", 4, dev, dtype, &po)) || (rc = dl_check(mask, "mask", 4, dev, dtype, &pm)) ||
        (rc = dl_check(grad_out, "grad_out", 4, dev, dtype, &pgo)) ||
        (rc = dl_check(grad_x, "grad_x", 4, dev, dtype, &pgx)) ||
        (rc = dl_check(grad_offset, "grad_offset", 4, dev, dtype, &pgoff)) ||
        (rc = dl_check(grad_mask, "grad_mask", 4, dev, dtype, &pgm)))
        return rc;
    dcnv3_params p;
    if ((rc = dl_params(x, offset, kh, kw, sh, sw, pad_h, pad_w, dh, dw, groups, group_channels,
                        offset_scale, flags, &p)))
        return rc;
    const int64_t C = (int64_t)groups * group_channels, GP = (int64_t)groups * kh * kw;
    if ((rc = dl_expect_shape(x, "x", p.n, p.h, p.w, C)) ||
        (rc = dl_expect_shape(offset, "offset", p.n, p.ho, p.wo, GP * 2)) ||
        (rc = dl_expect_shape(mask, "mask", p.n, p.ho, p.wo, GP)) ||
        (rc = dl_expect_shape(grad_out, "grad_out", p.n, p.ho, p.wo, C)) ||
        (rc = dl_expect_shape(grad_x, "grad_x", p.n, p.h, p.w, C)) ||
        (rc = dl_expect_shape(grad_offset, "grad_offset", p.n, p.ho, p.wo, GP * 2)) ||
        (rc = dl_expect_shape(grad_mask, "grad_mask", p.n, p.ho, p.wo, GP)))
        return rc;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    if (workspace != nullptr) {
        const DLTensor& t = workspace->dl_tensor;
        if (t.device.device_type != kDLCUDA || t.device.device_id != dev)
            return fail(DCNV3_ERR_DEVICE, "workspace must be a CUDA tensor on x's device");
        size_t count = 1;
        for (int i = 0; i < t.ndim; ++i) count *= (size_t)t.shape[i];
        ws = (char*)t.data + t.byte_offset;
        ws_bytes = count * ((t.dtype.bits + 7) / 8) * t.dtype.lanes;
    }
    return backward_impl(px, po, pm, pgo, pgx, pgoff, pgm, ws, ws_bytes, &p, (cudaStream_t)cuda_stream);
}

static int host_run(const void* x, const void* offset, const void* mask, const void* grad_out,
                    void* out, void* grad_x, void* grad_offset, void* grad_mask,
                    const dcnv3_params* p, int device, bool with_backward, int slot, bool wait) {
    int rc = check(p);
    if (rc) return rc;
    if (!x || !offset || !mask || !out || (with_backward && (!grad_out || !grad_x || !grad_offset || !grad_mask)))
        return fail(DCNV3_ERR_ARGUMENT, "NULL host buffer");
    if (p->n == 0) return DCNV3_OK;
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t es = elem_size(p->dtype);
    const size_t C = (size_t)p->groups * p->group_channels, GP = (size_t)p->groups * p->kh * p->kw;
    const size_t b_x = align_up((size_t)p->n * p->h * p->w * C * es, 256);
    const size_t b_o = align_up((size_t)p->n * p->ho * p->wo * C * es, 256);
    const size_t b_off = align_up((size_t)p->n * p->ho * p->wo * GP * 2 * es, 256);
    const size_t b_m = align_up((size_t)p->n * p->ho * p->wo * GP * es, 256);
    const size_t b_ws = with_backward ? align_up(backward_ws_bytes(p), 256) : 0;
    size_t total = b_x + b_off + b_m + b_o;
    if (with_backward) total += b_o + b_x + b_off + b_m;
    if (device < 0 || device >= 64) return fail(DCNV3_ERR_DEVICE, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(g_scratch_mu[device]);
    HostScratch* s;
    if ((rc = scratch_reserve(device, slot, total, b_ws, &s))) return rc;
    char* base = (char*)s->buf;
    char* d_x = base; base += b_x;
    char* d_off = base; base += b_off;
    char* d_m = base; base += b_m;
    char* d_out = base; base += b_o;
    cudaStream_t st = s->stream, sin = g_pipes[device].in, sout = g_pipes[device].out;
#define CK(call, what) if ((e = (call)) != cudaSuccess) return cuda_fail(e, what)
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sin), "H2D copy")
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, sout), "D2H copy")
    const size_t n_x = (size_t)p->n * p->h * p->w * C * es, n_o = (size_t)p->n * p->ho * p->wo * C * es;
    const size_t n_off = (size_t)p->n * p->ho * p->wo * GP * 2 * es, n_m = n_off / 2;
    char* d_go = base; base += with_backward ? b_o : 0;
    char* d_gx = base; base += with_backward ? b_x : 0;
    char* d_goff = base; base += with_backward ? b_off : 0;
    char* d_gm = base;
    // copy-in: after the slot's previous outputs have left the scratch (never-recorded event = no wait)
    CK(cudaStreamWaitEvent(sin, s->ev_out, 0), "cudaStreamWaitEvent");
    H2D(d_x, x, n_x);
    H2D(d_off, offset, n_off);
    H2D(d_m, mask, n_m);
    CK(cudaEventRecord(s->ev_in, sin), "cudaEventRecord");
    if (with_backward) {
        H2D(d_go, grad_out, n_o);
        CK(cudaEventRecord(s->ev_go, sin), "cudaEventRecord");
    }
    // forward as soon as its three inputs are resident; its output leaves while grad_out still arrives
    CK(cudaStreamWaitEvent(st, s->ev_in, 0), "cudaStreamWaitEvent");
    if ((rc = forward_impl(d_x, d_off, d_m, d_out, p, st))) return rc;
    CK(cudaEventRecord(s->ev_fwd, st), "cudaEventRecord");
    CK(cudaStreamWaitEvent(sout, s->ev_fwd, 0), "cudaStreamWaitEvent");
    D2H(out, d_out, n_o);
    if (with_backward) {
        CK(cudaStreamWaitEvent(st, s->ev_go, 0), "cudaStreamWaitEvent");
        dcnv3_params pb = *p;
        // the slot's workspace is zeroed at allocation; a call keeps its zero part zero and scribbles on the scratch
        // behind it, so a later call with a larger zero part re-zeroes the difference
        const size_t zb = backward_ws_zero_bytes(p);
        if (zb > s->ws_clean) CK(cudaMemsetAsync((char*)s->ws + s->ws_clean, 0, zb - s->ws_clean, st), "memset(workspace)");
        s->ws_clean = zb;
        pb.flags |= DCNV3_FLAG_WORKSPACE_ZEROED;
        if ((rc = backward_impl(d_x, d_off, d_m, d_go, d_gx, d_goff, d_gm, s->ws, s->ws_bytes, &pb, st))) return rc;
        CK(cudaEventRecord(s->ev_comp, st), "cudaEventRecord");
        CK(cudaStreamWaitEvent(sout, s->ev_comp, 0), "cudaStreamWaitEvent");
        D2H(grad_x, d_gx, n_x);
        D2H(grad_offset, d_goff, n_off);
        D2H(grad_mask, d_gm, n_m);
    }
    CK(cudaEventRecord(s->ev_out, sout), "cudaEventRecord");
#undef H2D
#undef D2H
    if (wait) CK(cudaStreamSynchronize(sout), "stream synchronize");
#undef CK
    return DCNV3_OK;
}

int dcnv3_forward_host(const void* x, const void* offset, const void* mask, void* out,
                       const dcnv3_params* p, int device) {
    return host_run(x, offset, mask, nullptr, out, nullptr, nullptr, nullptr, p, device, false, 0, true);
}

int dcnv3_forward_backward_host(const void* x, const void* offset, const void* mask,
                                const void* grad_out, void* out, void* grad_x, void* grad_offset,
                                void* grad_mask, const dcnv3_params* p, int device) {
    return host_run(x, offset, mask, grad_out, out, grad_x, grad_offset, grad_mask, p, device, true, 0, true);
}

int dcnv3_forward_backward_host_async(const void* x, const void* offset, const void* mask,
                                      const void* grad_out, void* out, void* grad_x, void* grad_offset,
                                      void* grad_mask, const dcnv3_params* p, int device, int slot) {
    // (the slot's scratch is reused once its previous outputs have been copied out: the copy-in stream
    //  waits for that on the device, the host does not block)
    return host_run(x, offset, mask, grad_out, out, grad_x, grad_offset, grad_mask, p, device, true, slot, false);
}

int dcnv3_host_sync(int device) {
    if (device < 0 || device >= 64) return fail(DCNV3_ERR_DEVICE, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(g_scratch_mu[device]);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if (g_pipes[device].in && (e = cudaStreamSynchronize(g_pipes[device].in)) != cudaSuccess)
        return cuda_fail(e, "stream synchronize");
    for (int k = 0; k < kHostSlots; ++k)
        if (g_scratch[device][k].stream && (e = cudaStreamSynchronize(g_scratch[device][k].stream)) != cudaSuccess)
            return cuda_fail(e, "stream synchronize");
    if (g_pipes[device].out && (e = cudaStreamSynchronize(g_pipes[device].out)) != cudaSuccess)
        return cuda_fail(e, "stream synchronize");
    return DCNV3_OK;
}

int dcnv3_host_slots(void) { return kHostSlots; }

int dcnv3_release_host_scratch(void) {
    for (int d = 0; d < 64; ++d) {
        std::lock_guard<std::mutex> lock(g_scratch_mu[d]);
        DevicePipes& dp = g_pipes[d];
        if (dp.in) {
            cudaSetDevice(d);
            cudaStreamSynchronize(dp.in);
        }
        for (int k = 0; k < kHostSlots; ++k) {
            HostScratch& s = g_scratch[d][k];
            if (s.buf || s.ws || s.stream) {
                cudaSetDevice(d);
                if (s.stream) cudaStreamSynchronize(s.stream);
                if (dp.out) cudaStreamSynchronize(dp.out);
                if (s.buf) cudaFree(s.buf);
                if (s.ws) cudaFree(s.ws);
                if (s.stream) cudaStreamDestroy(s.stream);
                if (s.ev_in) cudaEventDestroy(s.ev_in);
                if (s.ev_go) cudaEventDestroy(s.ev_go);
                if (s.ev_fwd) cudaEventDestroy(s.ev_fwd);
                if (s.ev_comp) cudaEventDestroy(s.ev_comp);
                if (s.ev_out) cudaEventDestroy(s.ev_out);
                s = HostScratch();
            }
        }
        if (dp.in) {
            cudaStreamDestroy(dp.in);
            cudaStreamDestroy(dp.out);
            dp = DevicePipes();
        }
    }
    return DCNV3_OK;
}

int dcnv3_set_kernel_timing(int enable) {
    KernelTiming& t = kernel_timing();
    if (enable && t.ev[0] == nullptr) {
        for (int i = 0; i < 5; ++i) {
            cudaError_t e = cudaEventCreate(&t.ev[i]);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        }
    }
    t.enabled = enable != 0;
    return DCNV3_OK;
}

int dcnv3_get_kernel_timing(float* ms4) {
    KernelTiming& t = kernel_timing();
    if (t.ev[0] == nullptr || ms4 == nullptr) return fail(DCNV3_ERR_ARGUMENT, "kernel timing was never enabled");
    cudaError_t e = cudaEventSynchronize(t.ev[4]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
    for (int i = 0; i < 4; ++i)
        if ((e = cudaEventElapsedTime(&ms4[i], t.ev[i], t.ev[i + 1])) != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime");
    return DCNV3_OK;
}

uint64_t dcnv3_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
