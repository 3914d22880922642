// Bird's-eye-view rotated NMS, one CTA per image.
//
// Replaces the reference's `DO_BEV_NMS` branch of DD3D.forward (core.py:137-151): nuscenes_sample_aggregate with one
// dummy group per image (postprocessing.py:58-108) -> sample_bev_nms (:22-55: boxes to the global frame through the
// image's pose) -> bev_nms (tridet/layers/bev_nms.py:99-133: top-surface rectangle in BEV, :51-96) -> detectron2
// batched_nms_rotated (polygon-clipping rotated IoU, greedy, class aware, score = scores_3d).
// Runs after the 2-D NMS / top-k kernel on its <= out_cap survivors (already sorted by scores_3d) and, like the
// reference, BEFORE detector_postprocess -- which this kernel then applies itself (scale, clip, drop empty).
#include "detect.cuh"
#include "device_once.cuh"

#include <math.h>

namespace dd3d {

namespace {

constexpr int kBevMax = 256;
constexpr int kBevThreads = 128;

struct P2 {
    float x, y;
};
__device__ __forceinline__ float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float dot2(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ P2 sub2(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }

// detectron2 box_iou_rotated_utils.h: get_rotated_vertices
__device__ void rect_vertices(float x, float y, float w, float h, float a, P2* p) {
    const float th = a * 0.01745329251994329577f;
    const float c = cosf(th) * 0.5f, s = sinf(th) * 0.5f;
    p[0] = {x + s * h + c * w, y + c * h - s * w};
    p[1] = {x - s * h + c * w, y - c * h - s * w};
    p[2] = {2.f * x - p[0].x, 2.f * y - p[0].y};
    p[3] = {2.f * x - p[1].x, 2.f * y - p[1].y};
}

// get_intersection_points: edge x edge intersections + vertices of one rectangle inside the other (<= 24 points)
__device__ int intersection_points(const P2* p1, const P2* p2, P2* out) {
    P2 v1[4], v2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v1[i] = sub2(p1[(i + 1) & 3], p1[i]);
        v2[i] = sub2(p2[(i + 1) & 3], p2[i]);
    }
    int n = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const float det = cross2(v2[j], v1[i]);
            if (fabsf(det) <= 1e-14f) continue;
            const P2 v12 = sub2(p2[j], p1[i]);
            const float t1 = cross2(v2[j], v12) / det, t2 = cross2(v1[i], v12) / det;
            if (t1 >= 0.f && t1 <= 1.f && t2 >= 0.f && t2 <= 1.f) out[n++] = {p1[i].x + v1[i].x * t1, p1[i].y + v1[i].y * t1};
        }
    {
        const P2 AB = v2[0], DA = v2[3];
        const float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
        for (int i = 0; i < 4; ++i) {
            const P2 AP = sub2(p1[i], p2[0]);
            const float APdotAB = dot2(AP, AB), APdotAD = -dot2(AP, DA);
            if (APdotAB >= 0.f && APdotAD >= 0.f && APdotAB <= ABdotAB && APdotAD <= ADdotAD) out[n++] = p1[i];
        }
    }
    {
        const P2 AB = v1[0], DA = v1[3];
        const float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
        for (int i = 0; i < 4; ++i) {
            const P2 AP = sub2(p2[i], p1[0]);
            const float APdotAB = dot2(AP, AB), APdotAD = -dot2(AP, DA);
            if (APdotAB >= 0.f && APdotAD >= 0.f && APdotAB <= ABdotAB && APdotAD <= ADdotAD) out[n++] = p2[i];
        }
    }
    return n;
}

// convex_hull_graham (shift_to_zero = true) followed by polygon_area
__device__ float hull_area(P2* q, int n) {
    int t = 0;
    for (int i = 1; i < n; ++i)
        if (q[i].y < q[t].y || (q[i].y == q[t].y && q[i].x < q[t].x)) t = i;
    const P2 start = q[t];
    for (int i = 0; i < n; ++i) q[i] = sub2(q[i], start);
    {
        const P2 tmp = q[0];
        q[0] = q[t];
        q[t] = tmp;
    }
    float dist[24];
    for (int i = 0; i < n; ++i) dist[i] = dot2(q[i], q[i]);
    // insertion sort of q[1..n) by polar angle around q[0] (ties: nearer first), as detectron2's CUDA path
    for (int i = 2; i < n; ++i) {
        const P2 qi = q[i];
        const float di = dist[i];
        int j = i - 1;
        while (j >= 1) {
            const float tmp = cross2(qi, q[j]);  // qi before q[j] ?
            const bool before = (fabsf(tmp) < 1e-6f) ? (di < dist[j]) : (tmp > 0.f);
            if (!before) break;
            q[j + 1] = q[j];
            dist[j + 1] = dist[j];
            --j;
        }
        q[j + 1] = qi;
        dist[j + 1] = di;
    }
    int k = 1;
    while (k < n && dist[k] <= 1e-8f) ++k;
    if (k == n) return 0.f;
    P2 hull[24];
    hull[0] = q[0];
    hull[1] = q[k];
    int m = 2;
    for (int i = k + 1; i < n; ++i) {
        while (m > 1 && cross2(sub2(q[i], hull[m - 2]), sub2(hull[m - 1], hull[m - 2])) >= 0.f) --m;
        hull[m++] = q[i];
    }
    if (m <= 2) return 0.f;
    float area = 0.f;
    for (int i = 1; i < m - 1; ++i) area += fabsf(cross2(sub2(hull[i], hull[0]), sub2(hull[i + 1], hull[0])));
    return area * 0.5f;
}

// single_box_iou_rotated; boxes are (cx, cy, w, h, angle_deg)
__device__ float rotated_iou(const float* b1, const float* b2) {
    const float a1 = b1[2] * b1[3], a2 = b2[2] * b2[3];
    if (a1 < 1e-14f || a2 < 1e-14f) return 0.f;
    const float sx = (b1[0] + b2[0]) * 0.5f, sy = (b1[1] + b2[1]) * 0.5f;  // shift centres for precision
    P2 p1[4], p2[4], pts[24];
    rect_vertices(b1[0] - sx, b1[1] - sy, b1[2], b1[3], b1[4], p1);
    rect_vertices(b2[0] - sx, b2[1] - sy, b2[2], b2[3], b2[4], p2);
    const int n = intersection_points(p1, p2, pts);
    if (n <= 2) return 0.f;
    const float inter = hull_area(pts, n);
    return inter / (a1 + a2 - inter);
}

__device__ void quat_to_mat3(const float* q, float* R) {  // pytorch3d quaternion_to_matrix, real-first
    const float r = q[0], i = q[1], j = q[2], k = q[3];
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    R[0] = 1.f - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r); R[4] = 1.f - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = 1.f - two_s * (i * i + j * j);
}

// pytorch3d matrix_to_quaternion (rotation_conversions.py): best-conditioned of the four candidates, real part first,
// no sign standardisation -- the quaternion the reference stores in pred_boxes3d_global (postprocessing.py:43-46).
__device__ void mat3_to_quat(const float* m, float* q) {
    const float arg[4] = {1.f + m[0] + m[4] + m[8], 1.f + m[0] - m[4] - m[8], 1.f - m[0] + m[4] - m[8],
                          1.f - m[0] - m[4] + m[8]};
    float qa[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        qa[i] = arg[i] > 0.f ? sqrtf(arg[i]) : 0.f;
        if (qa[i] > qa[best]) best = i;
    }
    const float c[4][4] = {{qa[0] * qa[0], m[7] - m[5], m[2] - m[6], m[3] - m[1]},
                           {m[7] - m[5], qa[1] * qa[1], m[3] + m[1], m[2] + m[6]},
                           {m[2] - m[6], m[3] + m[1], qa[2] * qa[2], m[5] + m[7]},
                           {m[3] - m[1], m[6] + m[2], m[7] + m[5], qa[3] * qa[3]}};
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = c[best][i] / den;
}

__device__ void invert_K(const float* K, float* iK) {  // adjugate in double, like decode.cu
    const double a = K[0], bb = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i9 = K[8];
    const double A = e * i9 - f * h, Bc = -(d * i9 - f * g), Cc = d * h - e * g;
    const double rdet = 1.0 / (a * A + bb * Bc + c * Cc);
    iK[0] = static_cast<float>(A * rdet);
    iK[1] = static_cast<float>(-(bb * i9 - c * h) * rdet);
    iK[2] = static_cast<float>((bb * f - c * e) * rdet);
    iK[3] = static_cast<float>(Bc * rdet);
    iK[4] = static_cast<float>((a * i9 - c * g) * rdet);
    iK[5] = static_cast<float>(-(a * f - c * d) * rdet);
    iK[6] = static_cast<float>(Cc * rdet);
    iK[7] = static_cast<float>(-(a * h - bb * g) * rdet);
    iK[8] = static_cast<float>((a * e - bb * d) * rdet);
}

// One decoded box -> global rotation / translation (postprocessing.py:25-46) and its BEV top-surface rectangle
// (boxes3d.py:47-64, bev_nms.py:71-96).
__device__ void box_to_global(const Det& D, const float* iK, const float* Rw, const float* pose_t, float* R, float* t,
                              float* rect) {
    const float u = D.proj_ctr[0], v = D.proj_ctr[1];
    const float tv[3] = {(iK[0] * u + iK[1] * v + iK[2]) * D.depth, (iK[3] * u + iK[4] * v + iK[5]) * D.depth,
                         (iK[6] * u + iK[7] * v + iK[8]) * D.depth};
    float Rs[9];
    quat_to_mat3(D.quat, Rs);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) R[r * 3 + cc] = Rw[r * 3] * Rs[cc] + Rw[r * 3 + 1] * Rs[3 + cc] + Rw[r * 3 + 2] * Rs[6 + cc];
        t[r] = Rw[r * 3] * tv[0] + Rw[r * 3 + 1] * tv[1] + Rw[r * 3 + 2] * tv[2] + pose_t[r];
    }
    const float hl = 0.5f * D.size[1], hw = 0.5f * D.size[0], hh = 0.5f * D.size[2];  // (l, w, h) = size[1, 0, 2]
    // corners 0, 1, 5, 4 of the template: (+l,+w,+h), (+l,-w,+h), (-l,-w,+h), (-l,+w,+h)
    const float sx[4] = {hl, hl, -hl, -hl}, sy[4] = {hw, -hw, -hw, hw};
    float bx[4], by[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float X = R[0] * sx[k] + R[1] * sy[k] + R[2] * hh + t[0];
        const float Y = R[3] * sx[k] + R[4] * sy[k] + R[5] * hh + t[1];
        bx[k] = -Y;  // VEHICLE_TO_BEV_ROTATION: (x, y)_bev = (-Y, -X)
        by[k] = -X;
    }
    const float fx = bx[0] - bx[3], fy = by[0] - by[3];
    rect[0] = 0.5f * (bx[0] + bx[2]);
    rect[1] = 0.5f * (by[0] + by[2]);
    rect[2] = sqrtf((bx[0] - bx[1]) * (bx[0] - bx[1]) + (by[0] - by[1]) * (by[0] - by[1]));  // width
    rect[3] = sqrtf(fx * fx + fy * fy);                                                        // length
    rect[4] = atan2f(fx, fy) * 57.29577951308232f;
}

// ------------------------------------------------------------------------------------------------------------------
// NuscenesDD3D sample aggregation (nuscenes_dd3d.py:449-463 -> postprocessing.py:58-108): rotated NMS jointly over the
// cameras of one sample.  Kernel 1: one CTA per sample group (sort by scores_3d, pairwise IoU bit matrix, greedy scan);
// kernel 2: one CTA per image (cap on the survivors of the whole call, order-preserving compaction).
constexpr int kGrpMax = 768;      // boxes per sample group (6 cameras x out_cap 128)
constexpr int kGrpSort = 1024;
constexpr int kGrpThreads = 256;
constexpr int kGrpImages = 16;    // cameras per sample the kernel can hold
constexpr int kGrpWords = kGrpMax / 64;

struct AggParams {
    Det* dets;             // [B][cap]
    int32_t* counts;       // [B]
    const float* K;        // [B][9]
    const float* poses;    // [B][7]
    const int32_t* group;  // [B] sample group of each image
    float* global;         // [B][cap][10] : global quat (w,x,y,z), global tvec, size (pred_boxes3d_global)
    float* keep_score;     // [B][cap] scratch: scores_3d of the NMS survivors, -1 elsewhere
    int32_t* flags;        // bit 3: a group had more than kGrpMax boxes / kGrpImages images
    int B, cap, max_dets;
    float thr;
};

struct GrpSmem {
    unsigned long long mask[kGrpMax][kGrpWords];
    unsigned long long key[kGrpSort];
    float rect[kGrpMax][5];
    int cls[kGrpMax];
    uint32_t src[kGrpMax];  // image << 16 | slot
    int imgs[kGrpImages];
    int offs[kGrpImages + 1];
    int nimg;
};

__global__ void __launch_bounds__(kGrpThreads) sample_nms_kernel(const AggParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GrpSmem& S = *reinterpret_cast<GrpSmem*>(smem_raw);
    const int g = blockIdx.x;
    if (threadIdx.x == 0) {
        int m = 0, total = 0;
        S.offs[0] = 0;
        for (int b = 0; b < p.B; ++b) {
            if (p.group[b] != g) continue;
            if (m == kGrpImages) {
                atomicOr(p.flags, 8);
                break;
            }
            int c = min(p.counts[b], p.cap);
            if (total + c > kGrpMax) {
                atomicOr(p.flags, 8);
                c = kGrpMax - total;
            }
            S.imgs[m] = b;
            total += c;
            S.offs[++m] = total;
        }
        S.nimg = m;
    }
    __syncthreads();
    const int nimg = S.nimg, n = S.offs[nimg];
    // ---- 1. boxes to the global frame, BEV rectangles, sort keys; keep_score = -1 on every slot of the group's images
    for (int m = 0; m < nimg; ++m)
        for (int s = threadIdx.x; s < p.cap; s += blockDim.x) p.keep_score[static_cast<size_t>(S.imgs[m]) * p.cap + s] = -1.0f;
    for (int i = threadIdx.x; i < kGrpSort; i += blockDim.x) S.key[i] = ~0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int m = 0;
        while (i >= S.offs[m + 1]) ++m;
        const int b = S.imgs[m], slot = i - S.offs[m];
        const Det& D = p.dets[static_cast<size_t>(b) * p.cap + slot];
        float iK[9], Rw[9], R[9], t[3], q[4];
        invert_K(p.K + b * 9, iK);
        quat_to_mat3(p.poses + b * 7, Rw);
        box_to_global(D, iK, Rw, p.poses + b * 7 + 4, R, t, S.rect[i]);
        mat3_to_quat(R, q);
        float* go = p.global + (static_cast<size_t>(b) * p.cap + slot) * 10;
        go[0] = q[0]; go[1] = q[1]; go[2] = q[2]; go[3] = q[3];
        go[4] = t[0]; go[5] = t[1]; go[6] = t[2];
        go[7] = D.size[0]; go[8] = D.size[1]; go[9] = D.size[2];
        S.cls[i] = D.cls;
        S.src[i] = (static_cast<uint32_t>(b) << 16) | static_cast<uint32_t>(slot);
        // descending scores_3d (non-negative floats order like their bit patterns), ties by concatenation index
        S.key[i] = (static_cast<unsigned long long>(~__float_as_uint(D.score3d)) << 32) | static_cast<unsigned>(i);
    }
    for (int i = threadIdx.x; i < kGrpMax * kGrpWords; i += blockDim.x) (&S.mask[0][0])[i] = 0ull;
    __syncthreads();
    // ---- 2. bitonic sort of the keys
    for (int k = 2; k <= kGrpSort; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < kGrpSort; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = S.key[i], b2 = S.key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b2) == up) {
                        S.key[i] = b2;
                        S.key[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    // ---- 3. pairwise rotated IoU in sorted order (same class; same sample by construction)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int ri = warp; ri < n; ri += kGrpThreads / 32) {
        const int i = static_cast<int>(S.key[ri] & 0xffffffffull);
        for (int rj = ri + 1 + lane; rj < n; rj += 32) {
            const int j = static_cast<int>(S.key[rj] & 0xffffffffull);
            if (S.cls[i] == S.cls[j] && rotated_iou(S.rect[i], S.rect[j]) > p.thr)
                atomicOr(&S.mask[ri][rj >> 6], 1ull << (rj & 63));
        }
    }
    __syncthreads();
    // ---- 4. greedy scan; survivors publish their score
    if (threadIdx.x == 0) {
        unsigned long long removed[kGrpWords];
#pragma unroll
        for (int w = 0; w < kGrpWords; ++w) removed[w] = 0ull;
        for (int r = 0; r < n; ++r) {
            if ((removed[r >> 6] >> (r & 63)) & 1ull) continue;
#pragma unroll
            for (int w = 0; w < kGrpWords; ++w) removed[w] |= S.mask[r][w];
            const int i = static_cast<int>(S.key[r] & 0xffffffffull);
            const uint32_t src = S.src[i];
            const size_t at = static_cast<size_t>(src >> 16) * p.cap + (src & 0xffffu);
            p.keep_score[at] = p.dets[at].score3d;
        }
    }
}

__global__ void __launch_bounds__(kBevThreads) sample_compact_kernel(const AggParams p) {
    __shared__ unsigned char keep[kBevMax];
    __shared__ int new_pos[kBevMax];
    __shared__ int s_total;
    const int b = blockIdx.x;
    const int n = min(min(p.counts[b], p.cap), kBevMax);
    const size_t all = static_cast<size_t>(p.B) * p.cap;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    // survivors of the whole call (keep = keep[:max_dets] is applied to the concatenation of ALL images of the call,
    // postprocessing.py:87-93)
    int part = 0;
    for (size_t i = threadIdx.x; i < all; i += blockDim.x) part += p.keep_score[i] >= 0.f;
    atomicAdd(&s_total, part);
    __syncthreads();
    const bool capped = p.max_dets > 0 && s_total > p.max_dets;
    __syncthreads();
    for (int s = threadIdx.x; s < kBevMax; s += blockDim.x) {
        bool k = false;
        if (s < n) {
            const size_t me = static_cast<size_t>(b) * p.cap + s;
            const float sc = p.keep_score[me];
            k = sc >= 0.f;
            if (k && capped) {  // rank among the survivors in the NMS output order: score desc, concatenation index asc
                int rank = 0;
                for (size_t i = 0; i < all; ++i) {
                    const float o = p.keep_score[i];
                    rank += (o > sc) || (o == sc && i < me);
                }
                k = rank < p.max_dets;
            }
        }
        keep[s] = k ? 1 : 0;
    }
    __syncthreads();
    Det mine[kBevMax / kBevThreads];
    float gl[kBevMax / kBevThreads][10];
    for (int s = 0; s < kBevMax / kBevThreads; ++s) {
        const int i = threadIdx.x + s * kBevThreads;
        if (i < n && keep[i]) {
            mine[s] = p.dets[static_cast<size_t>(b) * p.cap + i];
            const float* go = p.global + (static_cast<size_t>(b) * p.cap + i) * 10;
#pragma unroll
            for (int t = 0; t < 10; ++t) gl[s][t] = go[t];
        }
    }
    if (threadIdx.x == 0) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            new_pos[i] = m;
            m += keep[i];
        }
        s_total = m;
    }
    __syncthreads();
    for (int s = 0; s < kBevMax / kBevThreads; ++s) {
        const int i = threadIdx.x + s * kBevThreads;
        if (i < n && keep[i]) {
            p.dets[static_cast<size_t>(b) * p.cap + new_pos[i]] = mine[s];
            float* go = p.global + (static_cast<size_t>(b) * p.cap + new_pos[i]) * 10;
#pragma unroll
            for (int t = 0; t < 10; ++t) go[t] = gl[s][t];
        }
    }
    if (threadIdx.x == 0) p.counts[b] = s_total;
}

struct BevParams {
    Det* dets;             // [B][cap], compacted in place
    int32_t* counts;       // [B]
    const float* K;        // [B][9]
    const float* poses;    // [B][7] : pose quaternion (w, x, y, z) and translation (sensor -> global)
    const int32_t* sizes;  // [B][4] : h, w, out_h, out_w
    int32_t* flags;        // bit 2: more than kBevMax boxes
    int cap, do_postprocess;
    float thr;
};

__global__ void __launch_bounds__(kBevThreads) bev_nms_kernel(const BevParams p) {
    __shared__ float rect[kBevMax][5];
    __shared__ int cls[kBevMax];
    __shared__ unsigned long long mask[kBevMax][kBevMax / 64];
    __shared__ unsigned char keep[kBevMax];
    __shared__ int new_pos[kBevMax];
    __shared__ int s_total;
    const int b = blockIdx.x;
    int n = p.counts[b];
    if (n > kBevMax) {
        if (threadIdx.x == 0) atomicOr(p.flags, 4);
        n = kBevMax;
    }
    Det* dets = p.dets + static_cast<size_t>(b) * p.cap;
    // ---- 1. global-frame top-surface rectangles (postprocessing.py:25-46, boxes3d.py:47-64, bev_nms.py:71-96)
    const float* K = p.K + b * 9;
    const float* pose = p.poses + b * 7;
    float Rw[9];
    quat_to_mat3(pose, Rw);
    float iK[9];
    invert_K(K, iK);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float R[9], t[3];
        box_to_global(dets[i], iK, Rw, pose + 4, R, t, rect[i]);
        cls[i] = dets[i].cls;
    }
    for (int i = threadIdx.x; i < kBevMax * (kBevMax / 64); i += blockDim.x) (&mask[0][0])[i] = 0ull;
    __syncthreads();
    // ---- 2. pairwise rotated IoU (same class, j > i)
    for (int pidx = threadIdx.x; pidx < n * n; pidx += blockDim.x) {
        const int i = pidx / n, j = pidx - i * n;
        if (j > i && cls[i] == cls[j] && rotated_iou(rect[i], rect[j]) > p.thr)
            atomicOr(&mask[i][j >> 6], 1ull << (j & 63));
    }
    __syncthreads();
    // ---- 3. greedy scan in score order (the 2-D NMS kernel left the detections sorted by scores_3d)
    if (threadIdx.x == 0) {
        unsigned long long removed[kBevMax / 64] = {0ull, 0ull, 0ull, 0ull};
        for (int i = 0; i < n; ++i) {
            const bool r = (removed[i >> 6] >> (i & 63)) & 1ull;
            keep[i] = r ? 0 : 1;
            if (!r)
                for (int w = 0; w < kBevMax / 64; ++w) removed[w] |= mask[i][w];
        }
    }
    __syncthreads();
    // ---- 4. detector_postprocess on the survivors + order-preserving compaction (in place via registers)
    const int img_h = p.sizes[b * 4 + 0], img_w = p.sizes[b * 4 + 1], out_h = p.sizes[b * 4 + 2], out_w = p.sizes[b * 4 + 3];
    const float sxs = static_cast<float>(out_w) / static_cast<float>(img_w), sys = static_cast<float>(out_h) / static_cast<float>(img_h);
    Det mine[kBevMax / kBevThreads];
    for (int s = 0; s < kBevMax / kBevThreads; ++s) {
        const int i = threadIdx.x + s * kBevThreads;
        if (i < n) {
            mine[s] = dets[i];
            if (p.do_postprocess) {
                Det& D = mine[s];
                D.box[0] = fminf(fmaxf(D.box[0] * sxs, 0.f), static_cast<float>(out_w));
                D.box[2] = fminf(fmaxf(D.box[2] * sxs, 0.f), static_cast<float>(out_w));
                D.box[1] = fminf(fmaxf(D.box[1] * sys, 0.f), static_cast<float>(out_h));
                D.box[3] = fminf(fmaxf(D.box[3] * sys, 0.f), static_cast<float>(out_h));
                if (!((D.box[2] - D.box[0]) > 0.f && (D.box[3] - D.box[1]) > 0.f)) keep[i] = 0;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            new_pos[i] = m;
            m += keep[i];
        }
        s_total = m;
    }
    __syncthreads();
    for (int s = 0; s < kBevMax / kBevThreads; ++s) {
        const int i = threadIdx.x + s * kBevThreads;
        if (i < n && keep[i]) dets[new_pos[i]] = mine[s];
    }
    if (threadIdx.x == 0) p.counts[b] = s_total;
}

}  // namespace

cudaError_t launch_bev_nms(Det* dets, int32_t* counts, const float* K, const float* poses, const int32_t* sizes,
                           int32_t* flags, int B, int cap, float thr, int do_postprocess, cudaStream_t stream) {
    BevParams p;
    p.dets = dets;
    p.counts = counts;
    p.K = K;
    p.poses = poses;
    p.sizes = sizes;
    p.flags = flags;
    p.cap = cap;
    p.do_postprocess = do_postprocess;
    p.thr = thr;
    bev_nms_kernel<<<B, kBevThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

size_t sample_aggregate_scratch_bytes(int B, int cap) { return static_cast<size_t>(B) * cap * sizeof(float); }

cudaError_t launch_sample_aggregate(Det* dets, int32_t* counts, const float* K, const float* poses, const int32_t* group,
                                    int num_groups, float* global, void* scratch, int32_t* flags, int B, int cap,
                                    float thr, int max_dets, cudaStream_t stream) {
    if (cap > kBevMax || B >= 65536) return cudaErrorInvalidValue;
    AggParams p;
    p.dets = dets;
    p.counts = counts;
    p.K = K;
    p.poses = poses;
    p.group = group;
    p.global = global;
    p.keep_score = static_cast<float*>(scratch);
    p.flags = flags;
    p.B = B;
    p.cap = cap;
    p.max_dets = max_dets;
    p.thr = thr;
    static uint64_t configured_devices = 0;
    if (first_use_on_device(&configured_devices)) {
        const cudaError_t e = cudaFuncSetAttribute(sample_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(sizeof(GrpSmem)));
        if (e != cudaSuccess) return e;
    }
    sample_nms_kernel<<<num_groups, kGrpThreads, sizeof(GrpSmem), stream>>>(p);
    sample_compact_kernel<<<B, kBevThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dd3d
