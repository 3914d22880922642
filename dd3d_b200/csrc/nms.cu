// Per-image class-aware greedy NMS + post-NMS top-k + rescale/clip, one CTA per image, no host sync.
//
// Replaces Instances.cat over levels + FCOS2DInference.nms_and_top_k (reference core.py:130-135, fcos2d.py:346-367
// -> detectron2 batched_nms -> torchvision nms) and detectron2 detector_postprocess (core.py:153-160):
//   1. gather the <= L*topk decoded candidates of the image, sort by (score_3d desc, level asc, index asc)
//      with an in-smem bitonic sort (deterministic regardless of the atomics order upstream);
//   2. greedy NMS in sorted order, 64 boxes per step: resolve the 64x64 diagonal block serially, then let all
//      threads test the remaining boxes against the step's survivors (IoU > thr, same class => suppressed;
//      IoU = inter / (a + b - inter) exactly as torchvision);
//   3. if more than POST_NMS_TOPK remain keep those whose 2-D score >= the k-th largest 2-D score (fcos2d.py:359-365);
//   4. scale boxes to the requested output size, clip, drop empty boxes (detector_postprocess).
#include "detect.cuh"
#include "device_once.cuh"

namespace dd3d {

namespace {

constexpr int kNmsThreads = 1024;
constexpr int kMaxCand = 8192;  // bitonic sort capacity (>= L * topk)

__device__ __forceinline__ float iou_tv(const float4 a, const float4 b) {
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// ascending bitonic sort of n2 (power of two) 64-bit keys with a 16-bit payload, in shared memory
__device__ void bitonic_sort(uint64_t* key, uint16_t* val, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = key[i], b = key[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) {
                        key[i] = b;
                        key[ixj] = a;
                        const uint16_t t = val[i];
                        val[i] = val[ixj];
                        val[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// order-preserving compaction: dst gets the indices i (0..n) with flag[i] != 0; returns the count
__device__ int compact_indices(const uint8_t* flag, int n, int* dst, int* s_warp_sums, int* s_base) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const int f = (i < n && flag[i]) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        const int prefix = __popc(m & ((1u << lane) - 1));
        if (lane == 0) s_warp_sums[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            int v = s_warp_sums[lane];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_warp_sums[lane] = incl - v;  // exclusive
            if (lane == 31) s_warp_sums[32] = incl;
        }
        __syncthreads();
        const int base = *s_base;
        if (f) dst[base + s_warp_sums[warp] + prefix] = i;
        __syncthreads();
        if (threadIdx.x == 0) *s_base = base + s_warp_sums[32];
        __syncthreads();
    }
    return *s_base;
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const __grid_constant__ NmsParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int b = blockIdx.x;
    const int cap = kLevels * p.topk;
    // smem carve-up
    uint64_t* key = reinterpret_cast<uint64_t*>(smem_raw);                  // [kMaxCand]
    float4* boxes = reinterpret_cast<float4*>(key + kMaxCand);              // [cap]
    int* list = reinterpret_cast<int*>(boxes + cap);                        // [cap]
    int* list2 = list + cap;                                                // [cap]
    uint16_t* val = reinterpret_cast<uint16_t*>(list2 + cap);               // [kMaxCand]
    uint8_t* cls = reinterpret_cast<uint8_t*>(val + kMaxCand);              // [cap]
    uint8_t* flag = cls + cap;                                              // [cap]  (removed / keep flags)
    __shared__ int s_warp_sums[33];
    __shared__ int s_base;
    __shared__ int s_lvl_off[kLevels + 1];
    __shared__ unsigned long long s_kept_mask;
    __shared__ unsigned long long s_diag[64];

    const Det* cand = p.cand + static_cast<size_t>(b) * cap;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int l = 0; l < kLevels; ++l) {
            s_lvl_off[l] = off;
            off += min(p.cand_count[b * kLevels + l], p.topk);
        }
        s_lvl_off[kLevels] = off;
    }
    __syncthreads();
    const int n = s_lvl_off[kLevels];
    int n2 = 1;
    while (n2 < n) n2 <<= 1;

    // ---- 1. keys: (score3d desc, level asc, index asc); payload = slot in the candidate array
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        uint64_t k = ~0ull;
        uint16_t v = 0;
        if (i < n) {
            int l = 0;
#pragma unroll
            for (int t = 1; t < kLevels; ++t)
                if (i >= s_lvl_off[t]) l = t;
            const int slot = l * p.topk + (i - s_lvl_off[l]);
            const Det& d = cand[slot];
            const uint32_t sb = ~__float_as_uint(d.score3d);  // positive floats: larger score -> smaller key
            k = (static_cast<uint64_t>(sb) << 32) | (static_cast<uint64_t>(l) << 28) | static_cast<uint32_t>(d.index);
            v = static_cast<uint16_t>(slot);
        }
        key[i] = k;
        val[i] = v;
    }
    __syncthreads();
    if (p.do_nms) bitonic_sort(key, val, n2);

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const Det& d = cand[val[i]];
        boxes[i] = make_float4(d.box[0], d.box[1], d.box[2], d.box[3]);
        cls[i] = static_cast<uint8_t>(d.cls);
        flag[i] = 0;  // removed flag
    }
    __syncthreads();

    int nkeep = n;
    if (p.do_nms && p.nms_thresh > 0.f) {
        // ---- 2. greedy NMS, 64 sorted boxes per step
        for (int c0 = 0; c0 < n; c0 += 64) {
            const int cn = min(64, n - c0);
            // (A) diagonal block: 16 threads per row, 4 columns each; OR-reduce the 4-bit pieces with shuffles
            {
                const int i = threadIdx.x >> 4, part = threadIdx.x & 15;
                unsigned long long m = 0ull;
                if (i < cn) {
                    const float4 bi = boxes[c0 + i];
                    const uint8_t ci = cls[c0 + i];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = part * 4 + u;
                        if (j > i && j < cn && cls[c0 + j] == ci && iou_tv(bi, boxes[c0 + j]) > p.nms_thresh) m |= 1ull << j;
                    }
                }
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
                if (part == 0) s_diag[i] = m;
            }
            __syncthreads();
            // (B) warp 0 resolves the 64 boxes serially in registers (diag rows via shuffles)
            if (threadIdx.x < 32) {
                const int l = threadIdx.x;
                const unsigned long long d_lo = s_diag[l], d_hi = s_diag[l + 32];
                const unsigned r_lo = __ballot_sync(0xffffffffu, l < cn && flag[c0 + l]);
                const unsigned r_hi = __ballot_sync(0xffffffffu, l + 32 < cn && flag[c0 + l + 32]);
                unsigned long long removed = (static_cast<unsigned long long>(r_hi) << 32) | r_lo;
                if (cn < 64) removed |= ~0ull << cn;
                unsigned long long kept = 0ull;
#pragma unroll 8
                for (int i = 0; i < 64; ++i) {
                    const unsigned long long di = __shfl_sync(0xffffffffu, i < 32 ? d_lo : d_hi, i & 31);
                    if (!((removed >> i) & 1ull)) {
                        kept |= 1ull << i;
                        removed |= di;
                    }
                }
                if (l == 0) s_kept_mask = kept;
                if (l < cn) flag[c0 + l] = ((kept >> l) & 1ull) ? 0 : 1;
                if (l + 32 < cn) flag[c0 + l + 32] = ((kept >> (l + 32)) & 1ull) ? 0 : 1;
            }
            __syncthreads();
            // (C) every later box is tested against the step's survivors
            const unsigned long long kept = s_kept_mask;
            for (int k = c0 + cn + threadIdx.x; k < n; k += blockDim.x) {
                if (flag[k]) continue;
                const float4 bk = boxes[k];
                const uint8_t ck = cls[k];
                unsigned long long m = kept;
                bool rem = false;
                while (m) {
                    const int i = __ffsll(static_cast<long long>(m)) - 1;
                    m &= m - 1;
                    if (cls[c0 + i] == ck && iou_tv(boxes[c0 + i], bk) > p.nms_thresh) {
                        rem = true;
                        break;
                    }
                }
                if (rem) flag[k] = 1;
            }
            __syncthreads();
        }
        // keep flags = !removed
        for (int i = threadIdx.x; i < n; i += blockDim.x) flag[i] = flag[i] ? 0 : 1;
        __syncthreads();
        nkeep = compact_indices(flag, n, list, s_warp_sums, &s_base);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) list[i] = i;
        __syncthreads();
    }

    // ---- 3. post-NMS top-k on the 2-D score (>= k-th value keeps ties)
    if (p.do_nms && p.post_topk > 0 && nkeep > p.post_topk) {
        int m2 = 1;
        while (m2 < nkeep) m2 <<= 1;
        uint16_t* dummy = reinterpret_cast<uint16_t*>(list2);  // payload not needed; reuse list2 as scratch
        for (int i = threadIdx.x; i < m2; i += blockDim.x) {
            uint64_t k = ~0ull;
            if (i < nkeep) k = static_cast<uint64_t>(~__float_as_uint(cand[val[list[i]]].score));
            key[i] = k;
            dummy[i] = 0;
        }
        __syncthreads();
        // NOTE: sorting `key` destroys the sort keys of step 1 (no longer needed); val[] must stay intact,
        // so the payload array handed to the sort is the scratch one.
        bitonic_sort(key, dummy, m2);
        const uint32_t thr_bits = ~static_cast<uint32_t>(key[p.post_topk - 1]);
        const float thr = __uint_as_float(thr_bits);
        for (int i = threadIdx.x; i < nkeep; i += blockDim.x) flag[i] = (cand[val[list[i]]].score >= thr) ? 1 : 0;
        __syncthreads();
        const int m = compact_indices(flag, nkeep, list2, s_warp_sums, &s_base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) list2[i] = list[list2[i]];
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) list[i] = list2[i];
        __syncthreads();
        nkeep = m;
    }

    // ---- 4. detector_postprocess: scale, clip, drop empty
    const int img_h = p.sizes[b * 4 + 0], img_w = p.sizes[b * 4 + 1];
    const int out_h = p.sizes[b * 4 + 2], out_w = p.sizes[b * 4 + 3];
    const float sx = static_cast<float>(out_w) / static_cast<float>(img_w);
    const float sy = static_cast<float>(out_h) / static_cast<float>(img_h);
    if (p.do_postprocess) {
        for (int i = threadIdx.x; i < nkeep; i += blockDim.x) {
            float4 bx = boxes[list[i]];
            bx.x = fminf(fmaxf(bx.x * sx, 0.f), static_cast<float>(out_w));
            bx.z = fminf(fmaxf(bx.z * sx, 0.f), static_cast<float>(out_w));
            bx.y = fminf(fmaxf(bx.y * sy, 0.f), static_cast<float>(out_h));
            bx.w = fminf(fmaxf(bx.w * sy, 0.f), static_cast<float>(out_h));
            flag[i] = ((bx.z - bx.x) > 0.f && (bx.w - bx.y) > 0.f) ? 1 : 0;
        }
        __syncthreads();
        const int m = compact_indices(flag, nkeep, list2, s_warp_sums, &s_base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) list2[i] = list[list2[i]];
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) list[i] = list2[i];
        __syncthreads();
        nkeep = m;
    }
    if (threadIdx.x == 0) {
        if (nkeep > p.out_cap) atomicOr(p.flags, 2);
        p.out_count[b] = min(nkeep, p.out_cap);
    }
    const int nout = min(nkeep, p.out_cap);
    Det* out = p.out + static_cast<size_t>(b) * p.out_cap;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        Det d = cand[val[list[i]]];
        if (p.do_postprocess) {
            d.box[0] = fminf(fmaxf(d.box[0] * sx, 0.f), static_cast<float>(out_w));
            d.box[2] = fminf(fmaxf(d.box[2] * sx, 0.f), static_cast<float>(out_w));
            d.box[1] = fminf(fmaxf(d.box[1] * sy, 0.f), static_cast<float>(out_h));
            d.box[3] = fminf(fmaxf(d.box[3] * sy, 0.f), static_cast<float>(out_h));
        }
        out[i] = d;
    }
}

}  // namespace

cudaError_t launch_nms(const NmsParams& p, cudaStream_t stream) {
    const int cap = kLevels * p.topk;
    if (cap > kMaxCand || p.B <= 0) return cudaErrorInvalidValue;
    const size_t smem = static_cast<size_t>(kMaxCand) * 8 + static_cast<size_t>(cap) * 16 +
                        static_cast<size_t>(cap) * 8 + static_cast<size_t>(kMaxCand) * 2 + static_cast<size_t>(cap) * 2;
    static size_t attr_smem_dev[64] = {};  // per device; the limit is 227 KiB minus the static shared memory: ask for what we use
    size_t& attr_smem = attr_smem_dev[current_device_or_zero()];
    if (smem > attr_smem) {
        cudaError_t e =
            cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return e;
        }
        attr_smem = smem;
    }
    nms_kernel<<<p.B, kNmsThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dd3d
