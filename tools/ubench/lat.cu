// Microbenchmark: dependent-load latency (pointer chase) for a working set in L1 / L2 / DRAM, and written-by-another-kernel data.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void chase(const unsigned *p, int steps, unsigned *out, long long *cyc, int nc) {
    unsigned i = threadIdx.x + blockIdx.x * 977;
    long long t0 = clock64();
    for (int s = 0; s < steps; s++) i = nc ? __ldg(p + i) : p[i];
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = i;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void fillk(unsigned *p, size_t n, unsigned stride) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (unsigned)((i + stride) % n);
}
int main() {
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    for (size_t mb : {0ul, 1ul, 16ul, 64ul, 512ul}) {
        size_t n = mb ? mb * (1 << 20) / 4 : 4096;
        unsigned *p; cudaMalloc(&p, n * 4);
        fillk<<<1024, 256>>>(p, n, 12345 * 32 + 32);  // written by a kernel (dirty in L2), big stride
        cudaDeviceSynchronize();
        for (int nc = 0; nc < 2; nc++) {
            for (int rep = 0; rep < 2; rep++) chase<<<1, 1>>>(p, 2000, out, cyc, nc);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("working set %4zu MB (%s): %.0f cycles per dependent load (1 thread)\n", mb, nc ? "ld.nc" : "ld", (double)h / 2000);
        }
        // loaded system: 148*8 CTAs x 128 threads all chasing
        for (int rep = 0; rep < 2; rep++) chase<<<1184, 128>>>(p, 200, out, cyc, 1);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("working set %4zu MB loaded GPU (1184x128 threads, scattered): %.0f cycles per dependent load\n", mb, (double)h / 200);
        cudaFree(p);
    }
    return 0;
}
