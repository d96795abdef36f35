// HexPlane multi-resolution feature sampling, forward and backward.
//
// Replaces, per call, the 6*levels F.grid_sample launches + 5*levels multiplies + cat of
// scene/hexplane.py:73-106 (interpolate_ms_features) and the normalisation of :19-20,
// :153-164 with one kernel each way.  Semantics restated from there and from ATen's
// grid_sampler_2d (bilinear, align_corners=True, padding_mode='border'):
//   n      = (xyz - aabb[0]) * (2 / (aabb[1] - aabb[0])) - 1       aabb[0] = max, aabb[1] = min
//   c      = (n.x, n.y, n.z, t)                                      t used raw
//   plane k of a level pairs coordinates (a,b) in (0,1)(0,2)(0,3)(1,2)(1,3)(2,3); x = c[a]
//            indexes the plane's width (resolution of a), y = c[b] its height
//   ix     = clamp(((x + 1) / 2) * (W - 1), 0, W - 1)               same for iy
//   value  = nw*w_nw + ne*w_ne + sw*w_sw + se*w_se                   corner order of ATen
//   feature[level] = 1 * v0 * v1 * v2 * v3 * v4 * v5                 (left to right), levels concatenated
// Planes are read CHANNELS-LAST ([H][W][32] in memory; the nn.Parameter keeps its
// [1,32,H,W] shape with torch.channels_last strides): one warp handles one point with one
// lane per channel, so every texel fetch is a single coalesced 128-byte line.
#include "common.cuh"
#include "../../include/b200gs.h"

namespace b200gs {

namespace {

constexpr int HP_C = 32;                 // channels per plane (output_coordinate_dim)
constexpr int HP_MAXL = B200GS_HEXPLANE_MAX_LEVELS;

__constant__ int kPairA[6] = {0, 0, 0, 1, 1, 2};
__constant__ int kPairB[6] = {1, 2, 3, 2, 3, 3};

struct Bilinear {
    int o_nw, o_ne, o_sw, o_se;          // texel offsets (in texels), -1 if out of bounds
    float w_nw, w_ne, w_sw, w_se;
    float ix, iy;                        // clamped pixel coordinates
    float gx_mult, gy_mult;              // d(ix)/d(coord), zero where the border clamp is active
    int ix_nw, iy_nw;
};

__device__ __forceinline__ float unnormalize_clip(float coord, int size, float& mult)
{
    // ATen grid_sampler_compute_source_index_set_grad, align_corners = true, border padding
    float v = ((coord + 1.f) / 2) * (size - 1);
    mult = (float)(size - 1) / 2;
    const float hi = (float)(size - 1);
    if (v <= 0.f) { v = 0.f; mult = 0.f; }          // clip_coordinates_set_grad: gradient 0 at/below 0
    else if (v >= hi) { v = hi; mult = 0.f; }       // and at/above size-1
    return v;
}

__device__ __forceinline__ Bilinear bilinear_setup(float x, float y, int W, int H)
{
    Bilinear b;
    b.ix = unnormalize_clip(x, W, b.gx_mult);
    b.iy = unnormalize_clip(y, H, b.gy_mult);
    const float fx = floorf(b.ix), fy = floorf(b.iy);
    b.ix_nw = (int)fx; b.iy_nw = (int)fy;
    const float ix_ne = fx + 1.f, iy_sw = fy + 1.f;
    b.w_nw = (ix_ne - b.ix) * (iy_sw - b.iy);
    b.w_ne = (b.ix - fx) * (iy_sw - b.iy);
    b.w_sw = (ix_ne - b.ix) * (b.iy - fy);
    b.w_se = (b.ix - fx) * (b.iy - fy);
    const bool x0 = b.ix_nw >= 0 && b.ix_nw < W, x1 = b.ix_nw + 1 >= 0 && b.ix_nw + 1 < W;
    const bool y0 = b.iy_nw >= 0 && b.iy_nw < H, y1 = b.iy_nw + 1 >= 0 && b.iy_nw + 1 < H;
    b.o_nw = (x0 && y0) ? b.iy_nw * W + b.ix_nw : -1;
    b.o_ne = (x1 && y0) ? b.iy_nw * W + b.ix_nw + 1 : -1;
    b.o_sw = (x0 && y1) ? (b.iy_nw + 1) * W + b.ix_nw : -1;
    b.o_se = (x1 && y1) ? (b.iy_nw + 1) * W + b.ix_nw + 1 : -1;
    return b;
}

__device__ __forceinline__ float4 load_corners(const float* __restrict__ plane, const Bilinear& b, int lane)
{
    float4 v;
    v.x = b.o_nw >= 0 ? __ldg(plane + (size_t)b.o_nw * HP_C + lane) : 0.f;
    v.y = b.o_ne >= 0 ? __ldg(plane + (size_t)b.o_ne * HP_C + lane) : 0.f;
    v.z = b.o_sw >= 0 ? __ldg(plane + (size_t)b.o_sw * HP_C + lane) : 0.f;
    v.w = b.o_se >= 0 ? __ldg(plane + (size_t)b.o_se * HP_C + lane) : 0.f;
    return v;
}

__device__ __forceinline__ float interp(const float4 v, const Bilinear& b)
{
    float acc = __fmul_rn(v.x, b.w_nw);
    acc = __fmaf_rn(v.y, b.w_ne, acc);
    acc = __fmaf_rn(v.z, b.w_sw, acc);
    acc = __fmaf_rn(v.w, b.w_se, acc);
    return acc;
}

__device__ __forceinline__ void normalized_coords(const float* __restrict__ pts, const float* __restrict__ times,
                                                  float time_scalar, const float* __restrict__ aabb, size_t g,
                                                  float c[4], float scale[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float a0 = __ldg(aabb + a), a1 = __ldg(aabb + 3 + a);
        scale[a] = __fdiv_rn(2.0f, __fsub_rn(a1, a0));
        c[a] = __fsub_rn(__fmul_rn(__fsub_rn(__ldg(pts + 3 * g + a), a0), scale[a]), 1.0f);
    }
    c[3] = times ? __ldg(times + g) : time_scalar;
}

__global__ void __launch_bounds__(256)
hexplane_fwd_kernel(const __grid_constant__ b200gs_hexplane_desc d, long long P, const float* __restrict__ pts,
                    const float* __restrict__ times, float time_scalar, float* __restrict__ feat)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int F = d.levels * HP_C;
    for (long long g = warp0; g < P; g += nwarps) {
        float c[4], scale[3];
        normalized_coords(pts, times, time_scalar, d.aabb, (size_t)g, c, scale);
        for (int l = 0; l < d.levels; ++l) {
            float4 v[6];
            Bilinear bl[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int a = kPairA[k], b = kPairB[k];
                bl[k] = bilinear_setup(c[a], c[b], d.res[l][a], d.res[l][b]);
                v[k] = load_corners(d.plane[l][k], bl[k], lane);
            }
            float f = 1.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) f = __fmul_rn(f, interp(v[k], bl[k]));
            feat[(size_t)g * F + l * HP_C + lane] = f;
        }
    }
}

// Backward: plane gradients (atomic accumulation of coalesced 128-byte lines) and the
// gradient w.r.t. the query points (through grid_sample's grid input; the time coordinate
// is a constant and receives none).
__global__ void __launch_bounds__(256)
hexplane_bwd_kernel(const __grid_constant__ b200gs_hexplane_desc d, long long P, const float* __restrict__ pts,
                    const float* __restrict__ times, float time_scalar, const float* __restrict__ dfeat,
                    float* __restrict__ dpts /* [P,3], written */)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int F = d.levels * HP_C;
    for (long long g = warp0; g < P; g += nwarps) {
        float c[4], scale[3];
        normalized_coords(pts, times, time_scalar, d.aabb, (size_t)g, c, scale);
        float gc[3] = {0.f, 0.f, 0.f};          // per-lane partial d loss / d normalised coord
        for (int l = 0; l < d.levels; ++l) {
            float4 v[6];
            Bilinear bl[6];
            float val[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int a = kPairA[k], b = kPairB[k];
                bl[k] = bilinear_setup(c[a], c[b], d.res[l][a], d.res[l][b]);
                v[k] = load_corners(d.plane[l][k], bl[k], lane);
                val[k] = interp(v[k], bl[k]);
            }
            const float go = __ldg(dfeat + (size_t)g * F + l * HP_C + lane);
            // prefix / suffix products give d feature / d val[k] without divisions
            float pre[6], suf[6];
            pre[0] = 1.f;
#pragma unroll
            for (int k = 1; k < 6; ++k) pre[k] = pre[k - 1] * val[k - 1];
            suf[5] = 1.f;
#pragma unroll
            for (int k = 4; k >= 0; --k) suf[k] = suf[k + 1] * val[k + 1];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float gv = go * pre[k] * suf[k];
                const Bilinear& b = bl[k];
                float* gp = d.grad_plane[l][k];
                if (gp != nullptr && gv != 0.f) {
                    if (b.o_nw >= 0) atomicAdd(gp + (size_t)b.o_nw * HP_C + lane, gv * b.w_nw);
                    if (b.o_ne >= 0) atomicAdd(gp + (size_t)b.o_ne * HP_C + lane, gv * b.w_ne);
                    if (b.o_sw >= 0) atomicAdd(gp + (size_t)b.o_sw * HP_C + lane, gv * b.w_sw);
                    if (b.o_se >= 0) atomicAdd(gp + (size_t)b.o_se * HP_C + lane, gv * b.w_se);
                }
                // ATen grid_sampler_2d_backward: gix, giy
                const float fx = (float)b.ix_nw, fy = (float)b.iy_nw;
                const float ix_e = fx + 1.f, iy_s = fy + 1.f;
                float gix = -v[k].x * (iy_s - b.iy) * gv + v[k].y * (iy_s - b.iy) * gv
                            - v[k].z * (b.iy - fy) * gv + v[k].w * (b.iy - fy) * gv;
                float giy = -v[k].x * (ix_e - b.ix) * gv - v[k].y * (b.ix - fx) * gv
                            + v[k].z * (ix_e - b.ix) * gv + v[k].w * (b.ix - fx) * gv;
                const int a = kPairA[k], bb = kPairB[k];
                if (a < 3) gc[a] += b.gx_mult * gix;
                if (bb < 3) gc[bb] += b.gy_mult * giy;
            }
        }
        if (dpts != nullptr) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float s = gc[a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                gc[a] = s * scale[a];
            }
            if (lane < 3) dpts[3 * (size_t)g + lane] = lane == 0 ? gc[0] : (lane == 1 ? gc[1] : gc[2]);
        }
    }
}

int validate(const b200gs_hexplane_desc* d)
{
    if (!d) { set_error("hexplane: null descriptor"); return -1; }
    if (d->levels < 1 || d->levels > HP_MAXL) { set_error("hexplane: levels=%d unsupported (1..%d)", d->levels, HP_MAXL); return -1; }
    if (d->channels != HP_C) { set_error("hexplane: %d channels per plane unsupported (need %d)", d->channels, HP_C); return -1; }
    if (!d->aabb) { set_error("hexplane: aabb is null"); return -1; }
    for (int l = 0; l < d->levels; ++l) {
        for (int a = 0; a < 4; ++a) if (d->res[l][a] < 1) { set_error("hexplane: bad resolution"); return -1; }
        for (int k = 0; k < 6; ++k) if (!d->plane[l][k]) { set_error("hexplane: plane pointer is null"); return -1; }
    }
    return 0;
}

int grid_for(long long P)
{
    long long blocks = (P + 7) / 8;                  // 8 warps (= points in flight) per block
    const long long cap = (long long)NUM_SMS * 8;    // persistent-style cap: 8 resident blocks per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
}  // namespace b200gs

using namespace b200gs;

extern "C" {

int b200gs_hexplane_forward(const b200gs_hexplane_desc* desc, long long P, const float* pts, const float* times,
                            float time_scalar, float* features, b200gs_stream_t stream)
{
    if (validate(desc)) return -1;
    if (P <= 0) return 0;
    hexplane_fwd_kernel<<<grid_for(P), 256, 0, (cudaStream_t)stream>>>(*desc, P, pts, times, time_scalar, features);
    return check_launch("hexplane_forward");
}

int b200gs_hexplane_backward(const b200gs_hexplane_desc* desc, long long P, const float* pts, const float* times,
                             float time_scalar, const float* d_features, float* d_pts, b200gs_stream_t stream)
{
    if (validate(desc)) return -1;
    if (P <= 0) return 0;
    hexplane_bwd_kernel<<<grid_for(P), 256, 0, (cudaStream_t)stream>>>(*desc, P, pts, times, time_scalar, d_features, d_pts);
    return check_launch("hexplane_backward");
}

}  // extern "C"
