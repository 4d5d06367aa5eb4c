// Microbenchmark: what does it cost to write a 4096 x 4096 RGBA8 framebuffer (64 MiB) in 16 x 16 tiles on B200?
// The floor under the tile ("composite") kernel, by store mechanism. Every variant is timed with CUDA events around the
// launch, after a 256 MiB memset that leaves the L2 full of dirty lines (the condition tools/stage_times.py measures).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_floor store_floor.cu -lcuda && ./store_floor
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int W = 4096, H = 4096, TW = W / 16, TH = H / 16, NT = TW * TH;
constexpr size_t PITCH = (size_t)W * 4;

__global__ void k_empty() {}

// linear 16-byte stores, grid-stride
template <int MODE>
__device__ __forceinline__ void st16(uint4 *p, uint4 v) {
    if (MODE == 0) *p = v;
    if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 3) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 4) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    }
}

template <int MODE>
__global__ void k_linear(uint8_t *fb, const uint32_t *colors) {
    const size_t n = (size_t)W * H / 4;
    const uint32_t c = colors[0];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        st16<MODE>(reinterpret_cast<uint4 *>(fb) + i, make_uint4(c, c, c, c));
}

// the tile kernel's one-colour path: CTA of 128 threads takes 16 consecutive tiles, thread = (tile, 16-byte column),
// walks down the rows (2 row phases of 8 rows)
template <int MODE, int TPC>
__global__ void __launch_bounds__(128) k_tiles(uint8_t *fb, const uint32_t *colors) {
    __shared__ uint32_t col[TPC];
    const uint32_t t0 = blockIdx.x * TPC;
    if (threadIdx.x < TPC) col[threadIdx.x] = colors[t0 + threadIdx.x];
    __syncthreads();
    constexpr int ROW_STEP = 128 / (4 * TPC);
    const uint32_t tt = (threadIdx.x >> 2) % TPC, quarter = threadIdx.x & 3;
    const uint32_t t = t0 + tt, tx = t % TW, ty = t / TW, c = col[tt];
    uint8_t *dst = fb + (size_t)(ty * 16 + threadIdx.x / (4 * TPC)) * PITCH + (size_t)(tx * 16 + quarter * 4) * 4;
#pragma unroll
    for (int r = 0; r < 16 / ROW_STEP; r++, dst += ROW_STEP * PITCH) st16<MODE>(reinterpret_cast<uint4 *>(dst), make_uint4(c, c, c, c));
}

// a warp per tile ROW of the framebuffer tile: 512 threads write one 16-pixel-high band segment? no: warp per tile, lane =
// (row pair, 16-byte column): 2 stores of 16 bytes per lane ... the blend path's store pattern
template <int MODE>
__global__ void __launch_bounds__(128) k_warp_tile(uint8_t *fb, const uint32_t *colors) {
    const uint32_t t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t tx = t % TW, ty = t / TW, c = colors[t];
    uint8_t *dst = fb + (size_t)(ty * 16 + (lane >> 2) * 2) * PITCH + (size_t)(tx * 16 + (lane & 3) * 4) * 4;
    st16<MODE>(reinterpret_cast<uint4 *>(dst), make_uint4(c, c, c, c));
    st16<MODE>(reinterpret_cast<uint4 *>(dst + PITCH), make_uint4(c, c, c, c));
    dst += 8 * 2 * PITCH;  // (dummy second half: lanes cover rows 0..15 via (lane>>2)*2 -> 16 rows: nothing more)
}

// bulk async copies shared -> global (UBLKCP): one thread issues 16 row copies of 64 bytes per tile
template <int TPC>
__global__ void __launch_bounds__(128) k_bulk(uint8_t *fb, const uint32_t *colors) {
    __shared__ __align__(128) uint32_t rows[TPC][16];  // one 64-byte row of every tile's colour
    const uint32_t t0 = blockIdx.x * TPC;
    for (uint32_t i = threadIdx.x; i < TPC * 16; i += 128) rows[i >> 4][i & 15] = colors[t0 + (i >> 4)];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // thread = (tile, row): TPC * 16 copies over 128 threads
    for (uint32_t i = threadIdx.x; i < TPC * 16; i += 128) {
        const uint32_t tt = i >> 4, r = i & 15, t = t0 + tt, tx = t % TW, ty = t / TW;
        uint8_t *dst = fb + (size_t)(ty * 16 + r) * PITCH + (size_t)tx * 64;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(&rows[tt][0]);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 64;" ::"l"(dst), "r"(src) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// TMA tensor stores (UTMASTG): one thread per tile issues ONE 16 x 16-pixel box store from a 1 KiB shared tile
template <int TPC>
__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap map, const uint32_t *colors) {
    __shared__ __align__(128) uint32_t tiles[TPC][256];
    const uint32_t t0 = blockIdx.x * TPC;
    for (uint32_t i = threadIdx.x; i < TPC * 256; i += 128) tiles[i >> 8][i & 255] = colors[t0 + (i >> 8)];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < TPC) {
        const uint32_t t = t0 + threadIdx.x, tx = t % TW, ty = t / TW;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(&tiles[threadIdx.x][0]);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"(tx * 16), "r"(ty * 16), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// TMA with a wide box: one store covers 16 rows x 256 pixels (16 tiles of one tile row) from a 16 KiB shared band
__global__ void __launch_bounds__(128) k_tma_band(const __grid_constant__ CUtensorMap map, const uint32_t *colors) {
    extern __shared__ __align__(128) uint32_t band[];  // [16 rows][256 px]
    const uint32_t t0 = blockIdx.x * 16;
    for (uint32_t i = threadIdx.x; i < 16 * 256; i += 128) band[i] = colors[t0 + ((i & 255) >> 4)];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t tx = t0 % TW, ty = t0 / TW;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(band);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"(tx * 16), "r"(ty * 16), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(void *fb, uint32_t box_w, uint32_t box_h) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[2] = {W, H}, strides[1] = {PITCH};
    cuuint32_t box[2] = {box_w, box_h}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, fb, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return m;
}

static uint8_t *g_flush, *g_fb;
static uint32_t *g_colors;
static cudaEvent_t e0, e1;

template <typename F>
static void bench(const char *name, F launch, bool check = true) {
    std::vector<float> ts;
    for (int it = 0; it < 12; it++) {
        cudaMemsetAsync(g_flush, it, 256u << 20);
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ts.push_back(ms * 1e3f);
    }
    cudaError_t err = cudaGetLastError();
    std::sort(ts.begin() + 2, ts.end());
    const float med = ts[2 + 5];
    // correctness: every pixel equals its tile's colour
    const char *ok = "";
    if (check) {
        std::vector<uint32_t> h((size_t)W * H), col(NT);
        cudaMemcpy(h.data(), g_fb, (size_t)W * H * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(col.data(), g_colors, NT * 4, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int y = 0; y < H; y += 5)
            for (int x = 0; x < W; x += 3) bad += h[(size_t)y * W + x] != col[(y / 16) * TW + x / 16];
        ok = bad ? " WRONG" : " ok";
        cudaMemset(g_fb, 0, (size_t)W * H * 4);
    }
    printf("%-44s %7.1f us (min %.1f)  %6.0f GB/s%s %s\n", name, med, ts[2], 67.108864e6 / med / 1e3, ok,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    cudaMalloc(&g_flush, 256u << 20);
    cudaMalloc(&g_fb, (size_t)W * H * 4);
    cudaMalloc(&g_colors, NT * 4);
    std::vector<uint32_t> col(NT);
    for (int i = 0; i < NT; i++) col[i] = 0xff000000u | (uint32_t)(i * 2654435761u >> 8);
    col[0] = col[1];
    cudaMemcpy(g_colors, col.data(), NT * 4, cudaMemcpyHostToDevice);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d; 64 MiB framebuffer, 65536 tiles; times are event-to-event after a 256 MiB memset (dirty L2)\n", sms);
    bench("empty kernel", [] { k_empty<<<1, 32>>>(); }, false);
    bench("cudaMemsetAsync 64 MiB", [] { cudaMemsetAsync(g_fb, 1, (size_t)W * H * 4); }, false);
    {
        std::vector<uint32_t> one(NT, col[0]);
        cudaMemcpy(g_colors, one.data(), NT * 4, cudaMemcpyHostToDevice);
        bench("linear st.v4 grid 8/SM x 256", [=] { k_linear<0><<<sms * 8, 256>>>(g_fb, g_colors); });
        bench("linear st.cs.v4", [=] { k_linear<1><<<sms * 8, 256>>>(g_fb, g_colors); });
        bench("linear st.cg.v4", [=] { k_linear<2><<<sms * 8, 256>>>(g_fb, g_colors); });
        bench("linear st.wt.v4", [=] { k_linear<3><<<sms * 8, 256>>>(g_fb, g_colors); });
        bench("linear st L2::evict_first", [=] { k_linear<4><<<sms * 8, 256>>>(g_fb, g_colors); });
        bench("linear st.v4 grid 16/SM x 256", [=] { k_linear<0><<<sms * 16, 256>>>(g_fb, g_colors); });
        cudaMemcpy(g_colors, col.data(), NT * 4, cudaMemcpyHostToDevice);
    }
    bench("tiles: 16/CTA, thread=(tile,quarter) st.v4", [] { k_tiles<0, 16><<<NT / 16, 128>>>(g_fb, g_colors); });
    bench("tiles: 16/CTA st.cs", [] { k_tiles<1, 16><<<NT / 16, 128>>>(g_fb, g_colors); });
    bench("tiles: 16/CTA st.cg", [] { k_tiles<2, 16><<<NT / 16, 128>>>(g_fb, g_colors); });
    bench("tiles: 16/CTA L2::evict_first", [] { k_tiles<4, 16><<<NT / 16, 128>>>(g_fb, g_colors); });
    bench("tiles: 32/CTA st.v4", [] { k_tiles<0, 32><<<NT / 32, 128>>>(g_fb, g_colors); });
    bench("tiles: 8/CTA st.v4", [] { k_tiles<0, 8><<<NT / 8, 128>>>(g_fb, g_colors); });
    bench("warp per tile, 2 x st.v4 per lane", [] { k_warp_tile<0><<<NT / 4, 128>>>(g_fb, g_colors); });
    bench("warp per tile, st.cs", [] { k_warp_tile<1><<<NT / 4, 128>>>(g_fb, g_colors); });
    bench("bulk copies smem->global 64 B rows, 16/CTA", [] { k_bulk<16><<<NT / 16, 128>>>(g_fb, g_colors); });
    bench("bulk copies 64 B rows, 32/CTA", [] { k_bulk<32><<<NT / 32, 128>>>(g_fb, g_colors); });
    {
        const CUtensorMap m = make_map(g_fb, 16, 16);
        bench("TMA tensor store 16x16 box, 16 tiles/CTA", [=] { k_tma<16><<<NT / 16, 128>>>(m, g_colors); });
        bench("TMA tensor store 16x16 box, 8 tiles/CTA", [=] { k_tma<8><<<NT / 8, 128>>>(m, g_colors); });
        const CUtensorMap mb = make_map(g_fb, 256, 16);
        cudaFuncSetAttribute(k_tma_band, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        bench("TMA tensor store 256x16 band per CTA", [=] { k_tma_band<<<NT / 16, 128, 16384>>>(mb, g_colors); });
    }
    return 0;
}
