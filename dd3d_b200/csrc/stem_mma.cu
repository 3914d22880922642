// VoVNet stem_1 (3x3 stride 2, 3 -> 64, FrozenBN + ReLU; reference vovnet.py:302 via conv3x3 / stem, forward vovnet.py:357-359)
// as a register-fragment kernel.
//
// The layer is a pure streaming problem: 0.39 GB of normalised input in, 1.57 GB of 64-channel output out per 32-image batch
// (0.30 ms at the HBM roof) around 75 GFLOP.  The tcgen05 form (stem_tc.cu) builds a K-major im2col tile in shared memory
// per 128 pixels (thread-gathered, 9 predicated 8-byte loads + swizzled stores per pixel, then UMMA, then a TMEM round trip)
// and ran at 1.09 ms = 1.8 TB/s.  Here a CTA copies the 17 x 66 input patch of an 8 x 32 output tile with cp.async (double
// buffered), and every warp feeds mma.sync.m16n8k16 straight from it: one K step per kernel row, its 16 k slots = 4
// consecutive input pixels x 4 channels (4th pixel / 4th channel carry zero weights), so lane t's fragment registers are ONE
// 8-byte shared-memory load of input pixel 2x - 1 + t.  The 64 x 48 weight fragments stay in registers for the whole kernel;
// results are rounded in registers, staged per warp and leave as full 128-byte rows.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "act16.cuh"
#include "device_once.cuh"
#include "small_kernels.cuh"

namespace dd3d {

namespace {

constexpr int TH = 8, TW = 32;                   // output tile
constexpr int IH = 2 * TH + 1, IW = 2 * TW + 2;  // input patch 17 x 66 (one spare column for the zero-weight k slot)
constexpr int kInBytes = IH * IW * 8;
constexpr int kThreads = 256, kWarps = 8;
constexpr int kStageBytes = kWarps * 16 * 128;   // one 16-pixel x 64-channel tile per warp
constexpr int kSmemBytes = 2 * kInBytes + kStageBytes;
static_assert(TH * TW == kWarps * 2 * 16, "two 16-pixel M tiles per warp");

struct StemParams {
    const __nv_bfloat16* in;  // [B][H][W][4]
    const __nv_bfloat16* w;   // [64][3][4][4]: cout, ky, kx (kx = 3 zero), c (c = 3 zero)
    const float* sb;          // scale[64] | bias[64]
    __nv_bfloat16* out;       // [B][Ho][Wo][out_pitch]
    int B, H, W, Ho, Wo, out_pitch, tiles_x, tiles_y;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <bool FP16>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    if (FP16) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    } else {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void tile_coords(const StemParams& p, int tile, int* b, int* oy0, int* ox0) {
    const int per = p.tiles_x * p.tiles_y;
    *b = tile / per;
    const int r = tile - *b * per;
    const int ty = r / p.tiles_x;
    *oy0 = ty * TH;
    *ox0 = (r - ty * p.tiles_x) * TW;
}

__device__ __forceinline__ void load_input(const StemParams& p, int tile, uint32_t dst) {
    int b, oy0, ox0;
    tile_coords(p, tile, &b, &oy0, &ox0);
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    const __nv_bfloat16* img = p.in + static_cast<size_t>(b) * p.H * p.W * 4;
    for (int i = threadIdx.x; i < IH * IW; i += kThreads) {
        const int y = i / IW, x = i - y * IW;
        const int gy = iy0 + y, gx = ix0 + x;
        const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
        cp_async8(dst + i * 8, img + (ok ? (static_cast<size_t>(gy) * p.W + gx) * 4 : 0), ok ? 8u : 0u);
    }
}

template <bool FP16>
__global__ void __launch_bounds__(kThreads, 2) stem_s2_mma_kernel(const StemParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t s_in = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = lane >> 2, t = lane & 3;
    const uint32_t s_stage = s_in + 2 * kInBytes + warp * (16 * 128);
    const int total = p.B * p.tiles_x * p.tiles_y;

    // weight fragments of all 8 n-tiles x 3 kernel rows, resident for the whole kernel: lane (g, t) holds cout nt*8 + g,
    // input pixel kx = t, channels (0, 1) in b0 and (2, 3) in b1 -- the same k-slot mapping as the A loads below
    uint2 wb[3][8];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) wb[ky][nt] = __ldg(reinterpret_cast<const uint2*>(p.w) + ((nt * 8 + g) * 3 + ky) * 4 + t);

    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    int tile = blockIdx.x;
    if (tile < total) load_input(p, tile, s_in);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int buf = 0;
    for (; tile < total; tile += gridDim.x, buf ^= 1) {
        const int next = tile + gridDim.x;
        if (next < total) load_input(p, next, s_in + (buf ^ 1) * kInBytes);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();  // this tile's patch is visible; every warp finished reading the other buffer one iteration ago
        int b, oy0, ox0;
        tile_coords(p, tile, &b, &oy0, &ox0);
        const uint32_t s_cur = s_in + buf * kInBytes;
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
            // M tile = 16 consecutive output pixels of one tile row: row = warp, columns 16 m .. 16 m + 15
            const int oy = warp, oxl = 16 * m;
            const uint32_t a_lo = s_cur + ((2 * oy) * IW + 2 * (oxl + g) + t) * 8;
            const uint32_t a_hi = a_lo + 16 * 8;  // output pixel + 8 -> input pixel + 16
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const uint2 lo = lds64(a_lo + ky * IW * 8), hi = lds64(a_hi + ky * IW * 8);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) mma16816<FP16>(acc[nt], lo.x, hi.x, lo.y, hi.y, wb[ky][nt].x, wb[ky][nt].y);
            }
            // BN + ReLU + rounding in registers; stage 16 pixels x 128 B (16-byte chunks XOR-swizzled by the pixel index)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float s0 = __ldg(p.sb + nt * 8 + 2 * t), s1 = __ldg(p.sb + nt * 8 + 2 * t + 1);
                const float b0 = __ldg(p.sb + 64 + nt * 8 + 2 * t), b1 = __ldg(p.sb + 64 + nt * 8 + 2 * t + 1);
                const uint32_t v_lo = pack2_act(fmaxf(fmaf(acc[nt][0], s0, b0), 0.f), fmaxf(fmaf(acc[nt][1], s1, b1), 0.f), FP16);
                const uint32_t v_hi = pack2_act(fmaxf(fmaf(acc[nt][2], s0, b0), 0.f), fmaxf(fmaf(acc[nt][3], s1, b1), 0.f), FP16);
                sts32(s_stage + g * 128 + ((nt ^ g) << 4) + t * 4, v_lo);
                sts32(s_stage + (g + 8) * 128 + ((nt ^ g) << 4) + t * 4, v_hi);
            }
            __syncwarp();
            const int gy = oy0 + oy;
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 16 pixels x 8 chunks = 128 chunks: 4 pixels per instruction, full 128-byte rows
                const int c = j * 32 + lane, x = c >> 3, ch = c & 7;
                const uint4 v = lds128(s_stage + x * 128 + ((ch ^ (x & 7)) << 4));
                const int gx = ox0 + oxl + x;
                if (gy < p.Ho && gx < p.Wo)
                    *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b * p.Ho + gy) * p.Wo + gx) * p.out_pitch + ch * 8) = v;
            }
            __syncwarp();  // the staging tile is reused by the next M tile
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

// in4: [B][H][W][4]; w: 16-bit [64][3][4][4] (cout, ky, kx, c; kx = 3 and c = 3 zero); sb: fp32 scale[64] | bias[64];
// out: [B][ceil(H/2)][ceil(W/2)][out_pitch].
cudaError_t launch_stem_s2_mma(const __nv_bfloat16* in4, const __nv_bfloat16* w, const float* sb, __nv_bfloat16* out,
                               int out_pitch, int B, int H, int W, int num_sms, cudaStream_t stream, int fp16) {
    if (out_pitch % 8 || B < 1 || H < 1 || W < 1) return cudaErrorInvalidValue;
    StemParams p;
    p.in = in4; p.w = w; p.sb = sb; p.out = out;
    p.B = B; p.H = H; p.W = W; p.Ho = (H + 1) / 2; p.Wo = (W + 1) / 2;
    p.out_pitch = out_pitch;
    p.tiles_x = (p.Wo + TW - 1) / TW;
    p.tiles_y = (p.Ho + TH - 1) / TH;
    static uint64_t attr_devices[2] = {0, 0};
    if (first_use_on_device(&attr_devices[fp16 ? 1 : 0])) {
        cudaError_t e = fp16 ? cudaFuncSetAttribute(stem_s2_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)
                             : cudaFuncSetAttribute(stem_s2_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
    }
    const int total = B * p.tiles_x * p.tiles_y;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(std::min(total, 2 * num_sms));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return fp16 ? cudaLaunchKernelEx(&cfg, stem_s2_mma_kernel<true>, p) : cudaLaunchKernelEx(&cfg, stem_s2_mma_kernel<false>, p);
}

}  // namespace dd3d
