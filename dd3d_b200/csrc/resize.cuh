// GPU input pipeline: ResizeShortestEdge (PIL-exact antialiased bilinear) fused with the model preprocess (resize.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <utility>
#include <vector>

namespace dd3d {

// detectron2 ResizeShortestEdge.get_transform (sample_style "choice", one size), as used by
// tridet/data/augmentations/build.py:35-44 at test time.
void resize_shortest_edge_shape(int h, int w, int min_size, int max_size, int* new_h, int* new_w);

struct ResizeImage {  // per-image kernel operands
    int h0, w0, nh, nw;
    const int32_t* kx;    // [nw][ksx] 22-bit fixed-point horizontal coefficients
    const int32_t* xmin;  // [nw] first source column of each output column
    const int32_t* ky;    // [nh][ksy]
    const int32_t* ymin;  // [nh]
    int ksx, ksy;
    int flip;  // horizontal flip AFTER the resize (HFlipTransform.apply_image = np.flip(axis=1)); test-time augmentation
};

// Coefficient tables per (input size, output size) axis pair, cached on the device; not thread safe (one per engine).
class ResizeTables {
  public:
    ~ResizeTables();
    // d_raw: [B][raw_h][raw_w][3] uint8 (image b occupies the top-left h0 x w0 of its slot); h_raw_sizes / h_new_sizes:
    // [B][2] (h, w) on the host; d_out4: [B][Hp][Wp][4] bf16 = ((resized - mean) / std, zero padded).
    // h_flip: [B] flags or nullptr.
    cudaError_t launch(const uint8_t* d_raw, int raw_h, int raw_w, const int32_t* h_raw_sizes, const int32_t* h_new_sizes,
                       const int32_t* h_flip, __nv_bfloat16* d_out4, int B, int Hp, int Wp, const float mean[3],
                       const float std[3], cudaStream_t stream, int fp16 = 0);

  private:
    struct Axis {
        int32_t* d_k = nullptr;
        int32_t* d_min = nullptr;
        int ksize = 0;
        std::vector<int32_t> h_min;
    };
    const Axis* axis(int in_size, int out_size, cudaError_t* err);
    std::map<std::pair<int, int>, Axis> cache;
    std::vector<ResizeImage> h_img;
    ResizeImage* d_img = nullptr;
    int img_cap = 0;
    size_t smem_configured = 0;
    int smem_device = -1;
};

}  // namespace dd3d
