// Test-time input pipeline on the GPU: raw uint8 HWC images -> ResizeShortestEdge -> model
// preprocess, one fused kernel writing the engine's NHWC bf16 input tensor.
//
// Replaces, per image, the CPU work of DefaultDatasetMapper.__call__ (tridet/data/dataset_mappers/dataset_mapper.py:
// 100-153) at test time: detectron2 ResizeShortestEdge (shape) + ResizeTransform.apply_image = PIL.Image.resize(BILINEAR)
// (pixels) + the HWC->CHW copy, and DD3D.preprocess_image / ImageList padding (core.py:61-72).  The resampling is
// Pillow's ImagingResample restated exactly (src/libImaging/Resample.c): separable antialiased triangle filter,
// coefficients computed in double on the host and rounded to 22-bit fixed point, horizontal pass rounded to uint8 before
// the vertical pass -- so the pixels are BIT-IDENTICAL to what the reference's dataloader produces.
#include "resize.cuh"

#include "act16.cuh"

#include <math.h>

#include <vector>

namespace dd3d {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c PRECISION_BITS
constexpr int kTileW = 64;
constexpr int kTileH = 16;
constexpr int kResizeThreads = 256;

__device__ __forceinline__ int clip8(int acc) { return min(max(acc >> kPrecisionBits, 0), 255); }


struct ResizeParams {
    const uint8_t* raw;      // [B][raw_h][raw_w][3]
    const ResizeImage* img;  // [B]
    __nv_bfloat16* dst;      // [B][Hp][Wp][4]
    int raw_h, raw_w, Hp, Wp, rows_cap;
    float m0, m1, m2, s0, s1, s2;
    int fp16;  // 16-bit output format (act16.cuh)
};

__global__ void __launch_bounds__(kResizeThreads) resize_preprocess_kernel(const ResizeParams p) {
    extern __shared__ uint8_t hbuf[];  // [rows_cap][kTileW][3] horizontally resampled rows of this tile
    const int b = blockIdx.z;
    const ResizeImage I = p.img[b];
    const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
    __nv_bfloat16* dst = p.dst + static_cast<size_t>(b) * p.Hp * p.Wp * 4;
    const bool inside = y0 < I.nh && x0 < I.nw;
    int r_lo = 0, nrows = 0;
    if (inside) {
        const int y_last = min(y0 + kTileH, I.nh) - 1;
        r_lo = I.ymin[y0];
        nrows = min(I.ymin[y_last] + I.ksy, I.h0) - r_lo;
        const uint8_t* src = p.raw + static_cast<size_t>(b) * p.raw_h * p.raw_w * 3;
        // ---- horizontal pass (ImagingResampleHorizontal_8bpc) on the rows the vertical pass of this tile reads
        for (int idx = threadIdx.x; idx < nrows * kTileW; idx += kResizeThreads) {
            const int rr = idx / kTileW, xx = idx - rr * kTileW;
            const int x = x0 + xx;
            if (x >= I.nw) continue;
            const int xs = I.flip ? I.nw - 1 - x : x;  // flipped output column x shows resized column nw-1-x
            const int xm = I.xmin[xs];
            const int32_t* k = I.kx + static_cast<size_t>(xs) * I.ksx;
            const uint8_t* row = src + static_cast<size_t>(r_lo + rr) * p.raw_w * 3;
            int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
            for (int t = 0; t < I.ksx; ++t) {
                const int c = __ldg(k + t);
                const uint8_t* px = row + min(xm + t, I.w0 - 1) * 3;
                a0 += px[0] * c;
                a1 += px[1] * c;
                a2 += px[2] * c;
            }
            uint8_t* o = hbuf + idx * 3;
            o[0] = static_cast<uint8_t>(clip8(a0));
            o[1] = static_cast<uint8_t>(clip8(a1));
            o[2] = static_cast<uint8_t>(clip8(a2));
        }
    }
    __syncthreads();
    // ---- vertical pass (ImagingResampleVertical_8bpc) + (x - mean) / std + zero padding, NHWC4 bf16
    for (int idx = threadIdx.x; idx < kTileH * kTileW; idx += kResizeThreads) {
        const int yy = idx / kTileW, xx = idx - yy * kTileW;
        const int y = y0 + yy, x = x0 + xx;
        if (y >= p.Hp || x >= p.Wp) continue;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (y < I.nh && x < I.nw) {
            const int ym = I.ymin[y];
            const int32_t* k = I.ky + static_cast<size_t>(y) * I.ksy;
            int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
            for (int t = 0; t < I.ksy; ++t) {
                const int c = __ldg(k + t);
                const uint8_t* px = hbuf + ((min(ym + t, I.h0 - 1) - r_lo) * kTileW + xx) * 3;
                a0 += px[0] * c;
                a1 += px[1] * c;
                a2 += px[2] * c;
            }
            v0 = (static_cast<float>(clip8(a0)) - p.m0) / p.s0;
            v1 = (static_cast<float>(clip8(a1)) - p.m1) / p.s1;
            v2 = (static_cast<float>(clip8(a2)) - p.m2) / p.s2;
        }
        uint2 o;
        o.x = pack2_act(v0, v1, p.fp16);
        o.y = pack2_act(v2, 0.f, p.fp16);
        *reinterpret_cast<uint2*>(dst + (static_cast<size_t>(y) * p.Wp + x) * 4) = o;
    }
}

}  // namespace

void resize_shortest_edge_shape(int h, int w, int min_size, int max_size, int* new_h, int* new_w) {
    if (min_size <= 0) {
        *new_h = h;
        *new_w = w;
        return;
    }
    double scale = min_size * 1.0 / (h < w ? h : w);
    double newh, neww;
    if (h < w) {
        newh = min_size;
        neww = scale * w;
    } else {
        newh = scale * h;
        neww = min_size;
    }
    const double big = newh > neww ? newh : neww;
    if (big > max_size) {
        scale = max_size * 1.0 / big;
        newh = newh * scale;
        neww = neww * scale;
    }
    *new_h = static_cast<int>(newh + 0.5);
    *new_w = static_cast<int>(neww + 0.5);
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc, bilinear filter (support 1), box = [0, in_size)
static void bilinear_coeffs(int in_size, int out_size, std::vector<int32_t>* xmins, std::vector<int32_t>* kk, int* ksize_out) {
    const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
    const double ss = 1.0 / filterscale;
    xmins->assign(out_size, 0);
    kk->assign(static_cast<size_t>(out_size) * ksize, 0);
    std::vector<double> w(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < ksize; ++x) w[x] = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w[x] = a < 1.0 ? 1.0 - a : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) w[x] /= ww;
        for (int x = 0; x < ksize; ++x) {
            const double v = w[x] * (1 << kPrecisionBits);
            (*kk)[static_cast<size_t>(xx) * ksize + x] = w[x] < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
        }
        (*xmins)[xx] = xmin;
    }
    *ksize_out = ksize;
}

ResizeTables::~ResizeTables() {
    for (auto& kv : cache) {
        cudaFree(kv.second.d_k);
        cudaFree(kv.second.d_min);
    }
    if (d_img) cudaFree(d_img);
}

const ResizeTables::Axis* ResizeTables::axis(int in_size, int out_size, cudaError_t* err) {
    auto it = cache.find({in_size, out_size});
    if (it != cache.end()) return &it->second;
    std::vector<int32_t> mins, kk;
    Axis a;
    bilinear_coeffs(in_size, out_size, &mins, &kk, &a.ksize);
    a.h_min = mins;
    if ((*err = cudaMalloc(&a.d_k, kk.size() * 4)) != cudaSuccess) return nullptr;
    if ((*err = cudaMalloc(&a.d_min, mins.size() * 4)) != cudaSuccess) return nullptr;
    if ((*err = cudaMemcpy(a.d_k, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return nullptr;
    if ((*err = cudaMemcpy(a.d_min, mins.data(), mins.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return nullptr;
    return &cache.emplace(std::make_pair(in_size, out_size), a).first->second;
}

cudaError_t ResizeTables::launch(const uint8_t* d_raw, int raw_h, int raw_w, const int32_t* h_raw_sizes,
                                 const int32_t* h_new_sizes, const int32_t* h_flip, __nv_bfloat16* d_out4, int B, int Hp,
                                 int Wp, const float mean[3], const float std[3], cudaStream_t stream, int fp16) {
    cudaError_t err = cudaSuccess;
    h_img.resize(B);
    int rows_cap = 1;
    for (int b = 0; b < B; ++b) {
        const int h0 = h_raw_sizes[2 * b], w0 = h_raw_sizes[2 * b + 1], nh = h_new_sizes[2 * b], nw = h_new_sizes[2 * b + 1];
        if (h0 < 1 || w0 < 1 || h0 > raw_h || w0 > raw_w || nh < 1 || nw < 1 || nh > Hp || nw > Wp) return cudaErrorInvalidValue;
        const Axis* ax = axis(w0, nw, &err);
        if (!ax) return err;
        const Axis* ay = axis(h0, nh, &err);
        if (!ay) return err;
        ResizeImage& I = h_img[b];
        I.h0 = h0; I.w0 = w0; I.nh = nh; I.nw = nw;
        I.kx = ax->d_k; I.xmin = ax->d_min; I.ksx = ax->ksize;
        I.ky = ay->d_k; I.ymin = ay->d_min; I.ksy = ay->ksize;
        I.flip = h_flip ? (h_flip[b] != 0) : 0;
        for (int y0 = 0; y0 < nh; y0 += kTileH) {  // input rows one output tile needs
            const int y_last = (y0 + kTileH < nh ? y0 + kTileH : nh) - 1;
            int hi = ay->h_min[y_last] + ay->ksize;
            if (hi > h0) hi = h0;
            if (hi - ay->h_min[y0] > rows_cap) rows_cap = hi - ay->h_min[y0];
        }
    }
    const size_t smem = static_cast<size_t>(rows_cap) * kTileW * 3;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;  // > ~60x vertical shrink: not a DD3D configuration
    if (B > img_cap) {
        if (d_img) cudaFree(d_img);
        if ((err = cudaMalloc(&d_img, sizeof(ResizeImage) * B)) != cudaSuccess) return err;
        img_cap = B;
    }
    if ((err = cudaMemcpyAsync(d_img, h_img.data(), sizeof(ResizeImage) * B, cudaMemcpyHostToDevice, stream)) != cudaSuccess)
        return err;
    int dev_now = 0;
    cudaGetDevice(&dev_now);
    if (dev_now != smem_device) {  // tables (and the opt-in) belong to the device they were created on
        smem_configured = 0;
        smem_device = dev_now;
    }
    if (smem > 48 * 1024 && smem > smem_configured) {
        if ((err = cudaFuncSetAttribute(resize_preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem))) != cudaSuccess)
            return err;
        smem_configured = smem;
    }
    ResizeParams p;
    p.raw = d_raw;
    p.img = d_img;
    p.dst = d_out4;
    p.raw_h = raw_h; p.raw_w = raw_w; p.Hp = Hp; p.Wp = Wp; p.rows_cap = rows_cap;
    p.m0 = mean[0]; p.m1 = mean[1]; p.m2 = mean[2];
    p.s0 = std[0]; p.s1 = std[1]; p.s2 = std[2];
    p.fp16 = fp16;
    dim3 grid((Wp + kTileW - 1) / kTileW, (Hp + kTileH - 1) / kTileH, B);
    resize_preprocess_kernel<<<grid, kResizeThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dd3d
