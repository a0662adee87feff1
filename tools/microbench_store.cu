// HBM write-path microbenchmark for sm_100a: how fast can 704-byte rows (Panda FK output: 11 links x 64 B per
// configuration) be streamed out of shared memory, with no arithmetic at all?
//   mode 0: plain st.global.cs.v4, a warp writes 512 contiguous bytes per instruction (the ceiling for LSU stores)
//   mode 1: one cp.async.bulk.tensor.3d store per warp tile (32 rows x 704 B = 22 KB), the k_fk_forward_static_tmap pattern
//   mode 2: one cp.async.bulk (linear, 22 KB) per warp tile: a warp's 32 rows are contiguous in H
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_store tools/microbench_store.cu ; run on a B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int kRowBytes = 704, kTile = 32 * kRowBytes;   // 22528

__global__ void __launch_bounds__(128) k_plain(float4* out, long long n_rows) {
    const long long n4 = n_rows * (kRowBytes / 16);
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) __stcs(out + i, v);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap tmap, float* out, long long n_rows, int warps_per_block) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned tile_s = ((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
    tile_s += warp * kTile;
    for (int i = lane; i < kTile / 16; i += 32)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + i * 16), "f"(1.f), "f"(2.f), "f"(3.f), "f"((float)i) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const long long n_tiles = n_rows / 32, stride = (long long)gridDim.x * warps_per_block;
    for (long long t = (long long)blockIdx.x * warps_per_block + warp; t < n_tiles; t += stride) {
        if (lane == 0) {
            if (MODE == 1)
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(0), "r"(0), "r"((int)(t * 32)), "r"(tile_s) : "memory");
            else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * (kTile / 4)), "r"(tile_s), "n"(kTile) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // as the FK kernel must before refilling the tile
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const long long n_rows = 1 << 20;
    float* out;
    cudaMalloc(&out, n_rows * kRowBytes);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {16, 11, (cuuint64_t)n_rows};
    const cuuint64_t strides[2] = {64, 704};
    const cuuint32_t box[3] = {16, 11, 32}, estr[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto report = [&](const char* name, int warps, int blocks_per_sm, float ms) {
        printf("{\"mode\": \"%s\", \"warps_per_block\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n", name, warps, blocks_per_sm, ms,
               n_rows * (double)kRowBytes / ms / 1e6);
    };
    for (int bps : {4, 8, 16}) {
        for (int it = 0; it < 3; ++it) k_plain<<<148 * bps, 128>>>((float4*)out, n_rows);
        cudaEventRecord(a);
        for (int it = 0; it < 10; ++it) k_plain<<<148 * bps, 128>>>((float4*)out, n_rows);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        report("plain_st_v4", 4, bps, ms / 10);
    }
    for (int mode = 1; mode <= 2; ++mode)
        for (int warps : {1, 2, 4}) {
            const size_t smem = (size_t)warps * kTile + 1024;
            auto k = mode == 1 ? k_tma<1> : k_tma<2>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int bps = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, warps * 32, smem);
            for (int it = 0; it < 3; ++it) k<<<148 * bps, warps * 32, smem>>>(tmap, out, n_rows, warps);
            cudaEventRecord(a);
            for (int it = 0; it < 10; ++it) k<<<148 * bps, warps * 32, smem>>>(tmap, out, n_rows, warps);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            report(mode == 1 ? "tensor_3d_store" : "bulk_linear_store", warps, bps, ms / 10);
        }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
