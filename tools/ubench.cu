// Dev micro-benchmark (not part of the product): isolates per-instruction costs of the conv kernel's building blocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dd3d_b200/csrc tools/ubench.cu -o gpurun_out/ubench
//   mode 0: back-to-back tcgen05.mma (M128 x N x K16, SS) from one thread, `chains` accumulators, commit every `cper` MMAs
//   mode 1: TMA 2-D boxes [rows x 128 B] from a large matrix into an 8-deep ring, waited by the issuing thread
//   mode 2: TMA boxes issued by one thread, consumed (waited + released via mbarrier arrive) by another thread
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace dd3d;

__device__ __forceinline__ bool elect_one_() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// generic K-major smem descriptor: layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B, 0 = none (interleaved)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout) << 61;
    return d;
}

// operand descriptor of K16 slice k (0..3) of a 64-channel k-block stored in layout `lay` at `base`:
//   0: SW128  rows of 128 B, slice = +32 B inside the row, 8-row groups 1024 B apart
//   1: SW32   four sub-tiles [rows][32 B] (rows*32 B each), 8-row groups 256 B apart
//   2: SW64   two sub-tiles [rows][64 B], slice = +32 B inside the row, 8-row groups 512 B apart
//   3: none   core matrices 8 rows x 16 B contiguous; K chunks `rows*16` B apart, 8-row groups 128 B apart
__device__ __forceinline__ uint64_t slice_desc(uint32_t base, int k, int lay, int rows) {
    switch (lay) {
        case 1: return make_desc(base + k * rows * 32, 16, 256, 6);
        case 2: return make_desc(base + (k >> 1) * rows * 64 + (k & 1) * 32, 16, 512, 4);
        case 3: return make_desc(base + k * 2 * rows * 16, rows * 16, 128, 0);
        default: return make_desc(base + k * 32, 16, 1024, 2);
    }
}

__global__ void __launch_bounds__(128, 1)
mma_kernel(int N, int chains, int cper, int iters, int lay_a, int lay_b, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 1) {
        ptx::tmem_alloc(&slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0) {  // whole-warp loop, elect.sync around the issue only (the conv kernel's issue pattern)
        const uint32_t idesc = ptx::make_idesc_bf16(128, N);
        const uint32_t a = ptx::smem_u32(smem), b = a + 16384;
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ad[k] = slice_desc(a, k, lay_a, 128);
            bd[k] = slice_desc(b, k, lay_b, N);
        }
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 4) {  // one "k-block": 4 MMAs + a commit, like the conv kernel
            if (elect_one_()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, ad[k], bd[k], idesc, (i | k) != 0 ? 1u : 0u);
                ptx::umma_commit(&bar);
            }
            __syncwarp();
            phase ^= 1;
        }
        ptx::mbar_wait(&bar, phase ^ 1, 2);
        if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

__global__ void __launch_bounds__(128, 1)
tma_kernel(const __grid_constant__ CUtensorMap map, int rows, int stages, int iters, int split, int rows_total,
           long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[8], empty[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&empty[i], 1);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();
    const int box_bytes = rows * 128;
    const int base_row = (blockIdx.x * 9973) % (rows_total - 4096);
    if (!split) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            // keep `stages` boxes in flight: issue, and wait for the oldest before reusing its slot
            for (int i = 0; i < iters + stages; ++i) {
                const int s = i % stages;
                if (i >= stages) ptx::mbar_wait(&full[s], ((i / stages) - 1) & 1, 1);
                if (i < iters) {
                    ptx::mbar_expect_tx(&full[s], box_bytes);
                    ptx::tma_load_2d(smem + s * box_bytes, &map, &full[s], 0, base_row + (i * rows) % 4096);
                }
            }
            cycles[blockIdx.x] = clock64() - t0;
        }
    } else {
        if (threadIdx.x == 0) {  // producer
            for (int i = 0; i < iters; ++i) {
                const int s = i % stages;
                ptx::mbar_wait(&empty[s], ((i / stages) & 1) ^ 1, 1);
                ptx::mbar_expect_tx(&full[s], box_bytes);
                ptx::tma_load_2d(smem + s * box_bytes, &map, &full[s], 0, base_row + (i * rows) % 4096);
            }
        } else if (threadIdx.x == 32) {  // consumer
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const int s = i % stages;
                ptx::mbar_wait(&full[s], (i / stages) & 1, 2);
                ptx::mbar_arrive(&empty[s]);
            }
            cycles[blockIdx.x] = clock64() - t0;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    long long* d_cycles;
    cudaMalloc(&d_cycles, 148 * 8);
    long long h[148];
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("# mode0: MMA M128xNxK16 from ONE thread per SM, 148 CTAs. ns per MMA (event time / iters)\n");
    const int iters = 20000;
    const char* lname[4] = {"SW128", "SW32", "SW64", "NONE"};
    for (int N : {16, 64, 128, 192, 256}) {
        for (int lay_a : {0, 1, 2, 3}) {
            for (int lay_b : {0, 1}) {
                if (lay_b == 1 && lay_a != 1) continue;
                const int chains = 1, cper = 4;
                mma_kernel<<<148, 128, 50 * 1024>>>(N, chains, cper, iters, lay_a, lay_b, d_cycles);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                mma_kernel<<<148, 128, 50 * 1024>>>(N, chains, cper, iters, lay_a, lay_b, d_cycles);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
                printf("N=%3d A=%-5s B=%-5s : %7.1f ns/MMA  %7.1f cycles/MMA (clock64)  %s\n", N, lname[lay_a], lname[lay_b],
                       ms * 1e6 / iters, (double)h[0] / iters, cudaGetErrorString(err));
            }
        }
    }
    if (getenv("UBENCH_MMA_ONLY")) return 0;
    // TMA
    const int rows_total = 1 << 20;  // 1M rows x 128 B = 128 MB (>= L2)
    void* d_mat;
    cudaMalloc(&d_mat, (size_t)rows_total * 128);
    cudaMemset(d_mat, 0, (size_t)rows_total * 128);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    printf("# mode1/2: TMA boxes [rows x 128 B] (SW128), `stages` in flight, 148 CTAs. ns per box, GB/s per SM\n");
    for (int rows : {16, 128, 256}) {
        CUtensorMap map;
        cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)rows};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_mat, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("encode failed %d\n", (int)r);
            return 1;
        }
        for (int stages : {1, 2, 4, 6}) {
            for (int split : {0, 1}) {
                if (stages * rows * 128 > 190 * 1024) continue;
                const int it = 4000;
                tma_kernel<<<148, 128, 200 * 1024>>>(map, rows, stages, it, split, rows_total, d_cycles);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                tma_kernel<<<148, 128, 200 * 1024>>>(map, rows, stages, it, split, rows_total, d_cycles);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                const double ns = ms * 1e6 / it;
                printf("rows=%3d stages=%d split=%d : %7.1f ns/box  %6.1f GB/s/SM  %s\n", rows, stages, split, ns,
                       rows * 128 / ns, cudaGetErrorString(err));
            }
        }
    }
    return 0;
}
