// On-chip gather / scatter micro-benchmarks that decide the DCNv3 kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
// Each test runs NB CTAs per SM x 148 SMs, 256 threads, ITER inner iterations, and reports the
// aggregate on-chip bandwidth per SM per clock (using the measured SM clock of the run).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int CELLS = 512;          // 512 cells x 256 B = 128 KB? no: per-test layouts below
constexpr int ITER = 2048;

__device__ __forceinline__ unsigned rng(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// mode 0: LDS.128 gather, lane = (pixel, group) with rotated quads, cell layout [cell][4 groups][16 ch] (256 B)
// mode 1: LDS.128 gather, random 64 B slabs, same quad for all lanes (worst case)
// mode 2: LDS.128 gather, random slabs, rotated quads, single-group layout [cell][16 ch] (64 B cells)
// mode 3: RMW float4 (LDS.128 + 4 FADD + STS.128), layout as mode 0
// mode 4: ATOMS.ADD.32 x4 per lane (int fixed point), layout as mode 0
// mode 5: x-pair contiguous: 8 lanes read 128 contiguous bytes (2 cells x 64 B), random pairs
template <int MODE>
__global__ void __launch_bounds__(256) smem_kernel(float* out, long long* cycles, int ncell) {
    extern __shared__ float4 sm[];
    const int nq = ncell * 16;  // float4 count when cell = 256 B
    for (int i = threadIdx.x; i < nq; i += blockDim.x) sm[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    __syncthreads();
    unsigned s = threadIdx.x * 9781u + blockIdx.x * 7919u + 17u;
    const int lane = threadIdx.x & 31;
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        const unsigned r = rng(s);
        if (MODE == 0 || MODE == 3 || MODE == 4) {
            const int g = lane & 3, px = lane >> 2;
            const int cell = r % ncell;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int quad = (q + (g >> 1) + 2 * px) & 3;
                const int idx = cell * 16 + g * 4 + quad;
                if (MODE == 0) {
                    const float4 v = sm[idx];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                } else if (MODE == 3) {
                    float4 v = sm[idx];
                    v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 1.f;
                    sm[idx] = v;
                } else {
                    int* ip = reinterpret_cast<int*>(&sm[idx]);
                    atomicAdd(ip + 0, 1); atomicAdd(ip + 1, 2); atomicAdd(ip + 2, 3); atomicAdd(ip + 3, 4);
                }
            }
        } else if (MODE == 1 || MODE == 2) {
            const int slab = r % (ncell * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int quad = MODE == 1 ? q : ((q + lane) & 3);
                const float4 v = sm[slab * 4 + quad];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        } else if (MODE == 5) {
            // lanes: 4 pixel-groups x (xcorner 2 x quad 4); pair base random per pixel-group
            const unsigned rr = __shfl_sync(0xffffffffu, r, lane & ~7);
            const int pair = rr % (ncell * 4 - 1);
#pragma unroll
            for (int row = 0; row < 4; ++row) {
                const int idx = ((pair + row * 37) % (ncell * 4 - 1)) * 4 + (lane & 7);
                const float4 v = sm[idx];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + sm[threadIdx.x].x;
}

// global-memory gather / atomics on a buffer of `cells` 64 B slabs
// mode 0: LDG.128 gather of random slabs within a small window (L1 resident)   mode 1: window = whole buffer (L2)
// mode 2: red.global.add.v4.f32 random slabs                                   mode 3: atomicAdd u64 x2 per lane
template <int MODE>
__global__ void __launch_bounds__(256) gmem_kernel(float4* buf, float* out, size_t slabs, int window) {
    unsigned s = threadIdx.x * 9781u + blockIdx.x * 7919u + 17u;
    const int lane = threadIdx.x & 31;
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t base = ((size_t)blockIdx.x * 4099u) % (slabs - window);
    for (int it = 0; it < 256; ++it) {
        unsigned r = rng(s);
        r = __shfl_sync(0xffffffffu, r, lane & ~3);  // 4 lanes share a slab (one 64 B pixel-group)
        const size_t slab = base + r % window;
        if (MODE <= 1) {
            const float4 v = __ldg(&buf[slab * 4 + (lane & 3)]);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        } else if (MODE == 2) {
            atomicAdd(&buf[slab * 4 + (lane & 3)], make_float4(1.f, 1.f, 1.f, 1.f));
        } else {
            unsigned long long* p = reinterpret_cast<unsigned long long*>(&buf[slab * 4 + (lane & 3)]);
            atomicAdd(p, 3ull); atomicAdd(p + 1, 5ull);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>
void run_smem(const char* name, int nb_per_sm, int ncell, double bytes_per_lane_iter) {
    const int grid = 148 * nb_per_sm;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, grid * 256 * sizeof(float)));
    CK(cudaMalloc(&cyc, grid * sizeof(long long)));
    const size_t smem = (size_t)ncell * 256;
    CK(cudaFuncSetAttribute(smem_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    smem_kernel<MODE><<<grid, 256, smem>>>(out, cyc, ncell);
    CK(cudaEventRecord(a));
    smem_kernel<MODE><<<grid, 256, smem>>>(out, cyc, ncell);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    long long* h = (long long*)malloc(grid * sizeof(long long));
    CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    const double bytes_cta = 256.0 * ITER * bytes_per_lane_iter;
    // per-SM bytes per clock while nb CTAs are co-resident
    printf("%-46s nb/SM=%d smem=%3zuKB  %.1f us  cta_cycles=%.0f  B/clk/SM=%.1f\n", name, nb_per_sm, smem / 1024,
           ms * 1e3, mean, bytes_cta * nb_per_sm / mean);
    free(h); cudaFree(out); cudaFree(cyc);
}

template <int MODE>
void run_gmem(const char* name, float4* buf, size_t slabs, int window, double bytes_per_lane_iter) {
    const int grid = 148 * 16;
    float* out; CK(cudaMalloc(&out, grid * 256 * sizeof(float)));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    gmem_kernel<MODE><<<grid, 256>>>(buf, out, slabs, window);
    CK(cudaEventRecord(a));
    gmem_kernel<MODE><<<grid, 256>>>(buf, out, slabs, window);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double bytes = (double)grid * 256 * 256 * bytes_per_lane_iter;
    printf("%-46s window=%8d slabs  %.1f us  %.1f GB/s  (%.2f G slab-ops/s)\n", name, window, ms * 1e3,
           bytes / ms * 1e-6, (double)grid * 256 * 256 / 4 / ms * 1e-6);
    cudaFree(out);
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s, %d SMs, clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
    for (int nb : {1, 2, 4}) {
        const int ncell = nb == 1 ? 768 : (nb == 2 ? 384 : 192);
        run_smem<0>("LDS.128 gather (px,g) lanes, rotated quads", nb, ncell, 64);
        run_smem<1>("LDS.128 gather random slabs, same quad", nb, ncell, 64);
        run_smem<2>("LDS.128 gather random slabs, rotated quad", nb, ncell, 64);
        run_smem<5>("LDS.128 x-pair contiguous 128B", nb, ncell, 64);
        run_smem<3>("RMW float4 (LDS+FADD+STS) rotated", nb, ncell, 128);
        run_smem<4>("ATOMS.ADD.32 x4 rotated", nb, ncell, 64);
    }
    const size_t slabs = (size_t)1 << 21;  // 128 MB of 64 B slabs
    float4* buf; CK(cudaMalloc(&buf, slabs * 64)); CK(cudaMemset(buf, 0, slabs * 64));
    run_gmem<0>("LDG.128 gather, L1-resident window", buf, slabs, 512, 16);
    run_gmem<0>("LDG.128 gather, 2 MB window", buf, slabs, 32768, 16);
    run_gmem<1>("LDG.128 gather, whole 128 MB", buf, slabs, (int)(slabs / 2), 16);
    run_gmem<2>("RED.ADD.F32x4, 64 MB window", buf, slabs, (int)(slabs / 2), 16);
    run_gmem<2>("RED.ADD.F32x4, 2 MB window", buf, slabs, 32768, 16);
    run_gmem<3>("ATOMG.ADD.U64 x2, 64 MB window", buf, slabs, (int)(slabs / 2), 16);
    return 0;
}
