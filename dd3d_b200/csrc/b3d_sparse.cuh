// Sparse FCOS3D predictor (b3d_sparse.cu): the fused box3d 3x3 conv evaluated at the final 2-D candidates only.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "detect.cuh"

namespace dd3d {

struct B3dSparseLevel {
    const __nv_bfloat16* in;  // box3d tower output of the level, NHWC 16-bit [B][H][W][pitch], 256 channels
    const __nv_bfloat16* w;   // fused predictor weights [n_pad][9][256] (engine conv_layer layout; per level when PER_LEVEL_PREDICTORS)
    const float* scale;       // [n_pad] per-level epilogue (Scale folded)
    const float* bias;        // [n_pad] conv bias * scale (+ depth Offset)
    int H, W, pitch;
};

struct B3dSparseParams {
    B3dSparseLevel lvl[kLevels];
    const uint2* fin;           // [B][L][topk] (score bits, pixel * C + class): DecodeParams::fin
    const int32_t* cand_count;  // [B][L]
    float* rows;                // [B][L][topk][out_pitch] fp32, channel layout of a dense map pixel
    int B, C, topk, n_pad, out_pitch, fp16;
};

// widest fused predictor the kernel is instantiated for (14 n-tiles of 8: the 10 nuScenes classes); the engine keeps the dense
// predictor for models with more 3-D output channels
constexpr int kB3dSparseMaxN = 112;
cudaError_t launch_b3d_sparse(const B3dSparseParams& p, cudaStream_t stream);

}  // namespace dd3d
