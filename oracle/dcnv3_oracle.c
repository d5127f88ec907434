/*
 * CPU oracle for the DCNv3 core operator -- plain C restatement of the reference algorithm.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, never by the product path (iseg_b200/).
 *
 * Parity pinning: see the header of oracle/dcnv3_oracle.py -- the reference has no golden vectors
 * and real TensorFlow cannot run here ("parity unpinned" against TF); this file is pinned against
 * tests/golden/ (outputs of the reference's own op.py / utils.py executed over a torch stand-in for
 * the TF primitives) by tests/test_oracle.py.
 *
 * Follows, per sampled point (n,h,w,g,p):
 *   reference/layers/dcn_v3/op.py:46      zero padding            -> corner reads outside the image = 0
 *   reference/layers/dcn_v3/utils.py:14   get_reference_points    -> ref0 = (h*sh+y0c)/H_in in channel 0 (:52)
 *   reference/layers/dcn_v3/utils.py:65   generate_dilation_grids -> p = i*kh + j, i = W-direction
 *   reference/layers/dcn_v3/op.py:77-87   loc = ref + grid*s + offset*s/[W_in,H_in];  g = 2*loc-1
 *   reference/layers/dcn_v3/utils.py:142-206  pixel coords, floor, clip, weights from clipped corners,
 *                                          4-corner gather, *mask, accumulate over taps in order
 * Backward: analytic gradient of that graph (floor/cast/clip have zero gradient), grad_x scattered in
 * a fixed order (h, w, p, corner) inside each (n, g) plane -- deterministic for any thread count.
 *
 * Build with -ffp-contract=off so that no FMA is formed and the float operation order is the one
 * written here (matches the numpy restatement bit for bit in the forward pass).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n, h, w, ho, wo, groups, gc;
    int kh, kw, sh, sw, ph, pw, dh, dw;
    float scale;
} dcnv3_oracle_params;

typedef struct {
    int x0, x1, y0, y1;
    float dx0, dx1, dy0, dy1;
} tap_t;

#define TAP_FN(NAME, T, FLOOR)                                                                   \
    static inline void NAME(const dcnv3_oracle_params* q, int hin, int win, int h, int w, int p, \
                            T offx, T offy, int* x0, int* x1, int* y0, int* y1, T* dx0, T* dx1,  \
                            T* dy0, T* dy1) {                                                    \
        const int i = p / q->kh, j = p % q->kh;                                                  \
        const T s = (T)q->scale;                                                                 \
        const T y0c = (T)((q->dh * (q->kh - 1)) / 2 + 0.5f);                                     \
        const T x0c = (T)((q->dw * (q->kw - 1)) / 2 + 0.5f);                                     \
        const T ref0 = ((T)(h * q->sh) + y0c) / (T)hin;                                          \
        const T ref1 = ((T)(w * q->sw) + x0c) / (T)win;                                          \
        const T g0 = (T)(-((q->dw * (q->kw - 1)) / 2) + i * q->dw) / (T)win;                     \
        const T g1 = (T)(-((q->dh * (q->kh - 1)) / 2) + j * q->dh) / (T)hin;                     \
        T loc0 = ref0 + g0 * s;                                                                  \
        T loc1 = ref1 + g1 * s;                                                                  \
        loc0 = loc0 + offx * s / (T)win;                                                         \
        loc1 = loc1 + offy * s / (T)hin;                                                         \
        const T xq = (T)0.5 * (((((T)2 * loc0) - (T)1) + (T)1) * (T)(win - 2));                  \
        const T yq = (T)0.5 * (((((T)2 * loc1) - (T)1) + (T)1) * (T)(hin - 2));                  \
        T fx = FLOOR(xq), fy = FLOOR(yq);                                                        \
        /* clamp before the int conversion so that huge offsets cannot overflow */              \
        if (!(fx > (T)-2)) fx = (T)-2;                                                           \
        if (fx > (T)win) fx = (T)win;                                                            \
        if (!(fy > (T)-2)) fy = (T)-2;                                                           \
        if (fy > (T)hin) fy = (T)hin;                                                            \
        int ix = (int)fx, iy = (int)fy;                                                          \
        *x0 = ix < 0 ? 0 : (ix > win - 1 ? win - 1 : ix);                                        \
        *x1 = ix + 1 < 0 ? 0 : (ix + 1 > win - 1 ? win - 1 : ix + 1);                            \
        *y0 = iy < 0 ? 0 : (iy > hin - 1 ? hin - 1 : iy);                                        \
        *y1 = iy + 1 < 0 ? 0 : (iy + 1 > hin - 1 ? hin - 1 : iy + 1);                            \
        *dx0 = xq - (T)*x0;                                                                      \
        *dx1 = (T)*x1 - xq;                                                                      \
        *dy0 = yq - (T)*y0;                                                                      \
        *dy1 = (T)*y1 - yq;                                                                      \
    }

TAP_FN(tap_f32, float, floorf)

/* padded-image read: rows/cols are in padded coordinates */
static inline const float* slab(const float* x, const dcnv3_oracle_params* q, int n, int yp, int xp,
                                int g) {
    const int y = yp - q->ph, xx = xp - q->pw;
    if (y < 0 || y >= q->h || xx < 0 || xx >= q->w) return NULL;
    return x + ((((size_t)n * q->h + y) * q->w + xx) * q->groups + g) * q->gc;
}

int dcnv3_oracle_forward_f32(const float* x, const float* offset, const float* mask, float* out,
                             const dcnv3_oracle_params* q, int nthreads) {
    const int hin = q->h + 2 * q->ph, win = q->w + 2 * q->pw;
    const int P = q->kh * q->kw, G = q->groups, gc = q->gc;
    const long npix = (long)q->n * q->ho * q->wo;
    if (gc > 256) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long pix = 0; pix < npix; ++pix) {
        const int n = (int)(pix / ((long)q->ho * q->wo));
        const int h = (int)((pix / q->wo) % q->ho), w = (int)(pix % q->wo);
        float acc[256];
        for (int g = 0; g < G; ++g) {
            for (int c = 0; c < gc; ++c) acc[c] = 0.f;
            for (int p = 0; p < P; ++p) {
                const size_t gp = (size_t)pix * G * P + (size_t)g * P + p;
                int x0, x1, y0, y1;
                float dx0, dx1, dy0, dy1;
                tap_f32(q, hin, win, h, w, p, offset[2 * gp], offset[2 * gp + 1], &x0, &x1, &y0, &y1,
                        &dx0, &dx1, &dy0, &dy1);
                const float wa = dx1 * dy1, wb = dx1 * dy0, wc = dx0 * dy1, wd = dx0 * dy0;
                const float* ia = slab(x, q, n, y0, x0, g);
                const float* ib = slab(x, q, n, y1, x0, g);
                const float* ic = slab(x, q, n, y0, x1, g);
                const float* id = slab(x, q, n, y1, x1, g);
                const float m = mask[gp];
                for (int c = 0; c < gc; ++c) {
                    float s = (ia ? ia[c] : 0.f) * wa + (ib ? ib[c] : 0.f) * wb;
                    s = s + (ic ? ic[c] : 0.f) * wc;
                    s = s + (id ? id[c] : 0.f) * wd;
                    acc[c] = acc[c] + s * m;
                }
            }
            float* o = out + ((size_t)pix * G + g) * gc;
            for (int c = 0; c < gc; ++c) o[c] = acc[c];
        }
    }
    return 0;
}

int dcnv3_oracle_backward_f32(const float* x, const float* offset, const float* mask,
                              const float* grad_out, float* grad_x, float* grad_offset,
                              float* grad_mask, const dcnv3_oracle_params* q, int nthreads) {
    const int hin = q->h + 2 * q->ph, win = q->w + 2 * q->pw;
    const int P = q->kh * q->kw, G = q->groups, gc = q->gc;
    const float fx = (float)(win - 2) * q->scale / (float)win;
    const float fy = (float)(hin - 2) * q->scale / (float)hin;
    if (gc > 256) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    memset(grad_x, 0, sizeof(float) * (size_t)q->n * q->h * q->w * G * gc);
    /* one (n, g) plane per task: planes never share a grad_x cell, and inside a plane the order is
       fixed, so the result does not depend on the thread count */
#pragma omp parallel for schedule(dynamic, 1)
    for (int ng = 0; ng < q->n * G; ++ng) {
        const int n = ng / G, g = ng % G;
        for (int h = 0; h < q->ho; ++h)
            for (int w = 0; w < q->wo; ++w) {
                const size_t pix = ((size_t)n * q->ho + h) * q->wo + w;
                const float* go = grad_out + (pix * G + g) * gc;
                for (int p = 0; p < P; ++p) {
                    const size_t gp = pix * G * P + (size_t)g * P + p;
                    int xs[2], ys[2];
                    float dx0, dx1, dy0, dy1;
                    tap_f32(q, hin, win, h, w, p, offset[2 * gp], offset[2 * gp + 1], &xs[0], &xs[1],
                            &ys[0], &ys[1], &dx0, &dx1, &dy0, &dy1);
                    const float m = mask[gp];
                    /* corner order a b c d = (y0,x0) (y1,x0) (y0,x1) (y1,x1) */
                    const float wgt[4] = {dx1 * dy1, dx1 * dy0, dx0 * dy1, dx0 * dy0};
                    float dot[4];
                    for (int k = 0; k < 4; ++k) {
                        const int yy = ys[k & 1], xx = xs[k >> 1];
                        const float* src = slab(x, q, n, yy, xx, g);
                        float d = 0.f;
                        if (src)
                            for (int c = 0; c < gc; ++c) d = d + go[c] * src[c];
                        dot[k] = d;
                        float* dst = (float*)slab(grad_x, q, n, yy, xx, g);
                        if (dst) {
                            const float mw = m * wgt[k];
                            for (int c = 0; c < gc; ++c) dst[c] = dst[c] + go[c] * mw;
                        }
                    }
                    float gm = wgt[0] * dot[0] + wgt[1] * dot[1];
                    gm = gm + wgt[2] * dot[2];
                    gm = gm + wgt[3] * dot[3];
                    grad_mask[gp] = gm;
                    const float gxq = m * (dy1 * (dot[2] - dot[0]) + dy0 * (dot[3] - dot[1]));
                    const float gyq = m * (dx1 * (dot[1] - dot[0]) + dx0 * (dot[3] - dot[2]));
                    grad_offset[2 * gp] = gxq * fx;
                    grad_offset[2 * gp + 1] = gyq * fy;
                }
            }
    }
    return 0;
}

int dcnv3_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
