// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput and dependent latency on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};"
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}" : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
template <int MODE, int CHAINS>
__global__ void k(float *out, float a, float b, int iters, long long *cyc) {
    float2 v[CHAINS];
    for (int i = 0; i < CHAINS; i++) v[i] = make_float2(threadIdx.x + i, i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }
            else v[i] = fma2(v[i], a2, b2);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int CHAINS>
void run(const char *name, int warps) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<MODE, CHAINS><<<1, warps * 32>>>(out, 1.0001f, 0.5f, iters, cyc);
    k<MODE, CHAINS><<<1, warps * 32>>>(out, 1.0001f, 0.5f, iters, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fp32_per_iter = 2.0 * CHAINS;  // scalar-fma equivalents per thread per iteration
    printf("%-6s chains=%d warps=%2d: %.2f cycles/iter, %.2f cycles per scalar-fma-equivalent per warp, SM rate %.1f fma-lanes/clk\n", name, CHAINS, warps,
           (double)h / iters, (double)h / iters / fp32_per_iter, fp32_per_iter * warps * 32 * iters / (double)h);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) { run<0, 8>("FFMA", w); run<1, 8>("FFMA2", w); }
    run<0, 1>("FFMA", 1); run<1, 1>("FFMA2", 1);
    return 0;
}
