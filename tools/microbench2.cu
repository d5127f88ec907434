// ATOMS.ADD.32 throughput under different bank mappings (decides the grad_x scatter design).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ITER = 1024;
__device__ __forceinline__ unsigned rng(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// layout [cell][4 groups][16 ch] ints (256 B cells); lane = (px(8), g(4)); each lane adds its 16 channels
// MODE 0: channel order identical in all lanes (4-way bank conflict)
// MODE 1: channel order rotated so that the 32 lanes of every ATOMS hit 32 distinct banks
// MODE 2: as 1 but pairs of lanes (px, px^1) target the SAME cell (intra-warp address collisions... different g -> none) -> use same cell for all px: 8 lanes per address
// MODE 3: as 1, two limbs (2 ATOMS per channel, second array)
// MODE 4: plain RMW via LDS.128/STS.128 (reference)
// MODE 5: as 1 but using 64-bit vectorised index math removed: ATOMS with immediate offsets
template <int MODE>
__global__ void __launch_bounds__(256) k(int* out, long long* cycles, int ncell) {
    extern __shared__ int sm[];
    const int total = ncell * 64 * (MODE == 3 ? 2 : 1);
    for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    unsigned s = threadIdx.x * 9781u + blockIdx.x * 7919u + 17u;
    const int lane = threadIdx.x & 31;
    const int g = lane & 3, px = lane >> 2;
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        unsigned r = rng(s);
        if (MODE == 2) r = __shfl_sync(0xffffffffu, r, g);
        const int cell = r % ncell;
        int* base = sm + cell * 64 + g * 16;
        if (MODE == 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int quad = (q + (g >> 1) + 2 * px) & 3;
                int4* p = reinterpret_cast<int4*>(base + quad * 4);
                int4 v = *p; v.x += 1; v.y += 2; v.z += 3; v.w += 4; *p = v;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int quad, jj;
                    if (MODE == 0) { quad = q; jj = j; }
                    else { quad = (q + (g >> 1) * 2 + (px & 1)) & 3; jj = (j + (px >> 1)) & 3; }
                    atomicAdd(base + quad * 4 + jj, it + j);
                    if (MODE == 3) atomicAdd(base + ncell * 64 + quad * 4 + jj, it - j);
                }
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

template <int MODE>
void run(const char* name, int nb, int ncell) {
    const int grid = 148 * nb;
    int* out; long long* cyc;
    CK(cudaMalloc(&out, grid * 256 * sizeof(int))); CK(cudaMalloc(&cyc, grid * sizeof(long long)));
    const size_t smem = (size_t)ncell * 256 * (MODE == 3 ? 2 : 1);
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<MODE><<<grid, 256, smem>>>(out, cyc, ncell);
    k<MODE><<<grid, 256, smem>>>(out, cyc, ncell);
    CK(cudaDeviceSynchronize());
    long long* h = (long long*)malloc(grid * sizeof(long long));
    CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    // element updates (one fp32-equivalent accumulator update) per clock per SM
    const double upd = 256.0 * ITER * 16 * nb / mean;
    printf("%-58s nb/SM=%d  cta_cycles=%.0f  updates/clk/SM=%.2f  (cycles per 576-update pixel-group: %.1f)\n", name, nb, mean, upd, 576.0 / upd);
    free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int nb : {1, 2, 4}) {
        const int ncell = nb == 1 ? 384 : (nb == 2 ? 192 : 96);
        run<0>("ATOMS.ADD same channel order (4-way bank conflict)", nb, ncell);
        run<1>("ATOMS.ADD rotated, 32 distinct banks", nb, ncell);
        run<2>("ATOMS.ADD rotated, 8 lanes same cell (diff ch/g -> no addr clash)", nb, ncell);
        run<3>("ATOMS.ADD rotated, two limbs", nb, ncell);
        run<4>("RMW int4 LDS/STS rotated", nb, ncell);
    }
    return 0;
}
