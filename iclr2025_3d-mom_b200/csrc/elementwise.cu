// Small fused elementwise kernels of the training loop that PyTorch spreads over ~25 launches per view:
//   * activations  gaussian_renderer/__init__.py:130-132  scales = exp(s), rotations = F.normalize(r),
//                  opacity = sigmoid(o)  (scene/gaussian_model.py:37-47) -- forward and backward;
//   * L1 loss      utils/loss_utils.py:23-24 (mean |render - gt|) with its gradient written in the same pass
//                  (train_4DGS.py:210: the loss the fine stage optimises).
// All HBM-bound: 32 B read + 32 B written per Gaussian forward, 64 + 32 backward; 8 B read + 4 B written per
// image element.
#include "common.cuh"
#include "../../include/b200gs.h"

namespace b200gs {
namespace {

__global__ void __launch_bounds__(256) activations_fwd_kernel(long long P, const float* __restrict__ s, const float* __restrict__ r,
                                                              const float* __restrict__ o, float* __restrict__ so,
                                                              float* __restrict__ ro, float* __restrict__ oo)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) so[3 * i + c] = expf(__ldg(s + 3 * i + c));
    const float4 q = __ldg(reinterpret_cast<const float4*>(r) + i);
    // F.normalize: x / max(||x||_2, 1e-12)
    const float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    reinterpret_cast<float4*>(ro)[i] = make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
    oo[i] = 1.0f / (1.0f + expf(-__ldg(o + i)));
}

__global__ void __launch_bounds__(256) activations_bwd_kernel(long long P, const float* __restrict__ so, const float* __restrict__ r,
                                                              const float* __restrict__ oo, const float* __restrict__ gs,
                                                              const float* __restrict__ gr, const float* __restrict__ go,
                                                              float* __restrict__ ds, float* __restrict__ dr, float* __restrict__ dopa)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    if (ds) {
#pragma unroll
        for (int c = 0; c < 3; ++c) ds[3 * i + c] = gs ? __ldg(gs + 3 * i + c) * __ldg(so + 3 * i + c) : 0.f;     // d exp = g * exp
    }
    if (dr) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(r) + i);
        const float4 g = gr ? __ldg(reinterpret_cast<const float4*>(gr) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float nn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float4 d;
        if (nn > 1e-12f) {          // y = x / n:  dx = (g - y (y . g)) / n
            const float inv = 1.0f / nn;
            const float4 y = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
            const float yg = y.x * g.x + y.y * g.y + y.z * g.z + y.w * g.w;
            d = make_float4((g.x - y.x * yg) * inv, (g.y - y.y * yg) * inv, (g.z - y.z * yg) * inv, (g.w - y.w * yg) * inv);
        } else {                    // clamp active: y = x / 1e-12
            d = make_float4(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f, g.w * 1e12f);
        }
        reinterpret_cast<float4*>(dr)[i] = d;
    }
    if (dopa) {
        const float y = __ldg(oo + i);
        dopa[i] = go ? __ldg(go + i) * (1.0f - y) * y : 0.f;                                                        // sigmoid_backward
    }
}

// loss[0] += scale * sum |a - b| ; d[i] = scale * sign(a - b)   (sign(0) = 0 like torch.sign); sse[0] += sum (a - b)^2 when asked
// for (the numerator of utils/image_utils.py:psnr, which train_4DGS.py:212 evaluates on the same two images every iteration).
__device__ __forceinline__ void l1_block_finish(float acc, float sq, float scale, float* __restrict__ loss, float* __restrict__ sse)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    __shared__ float part[8], part_sq[8];
    if ((threadIdx.x & 31) == 0) { part[threadIdx.x >> 5] = acc; part_sq[threadIdx.x >> 5] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { s += part[w]; q += part_sq[w]; }
        atomicAdd(loss, s * scale);
        if (sse) atomicAdd(sse, q);
    }
}

__global__ void __launch_bounds__(256) l1_fwd_bwd_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                                                         float scale, float* __restrict__ loss, float* __restrict__ sse,
                                                         float* __restrict__ d)
{
    float acc = 0.f, sq = 0.f;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
        const float e[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc += fabsf(e[k]); sq = fmaf(e[k], e[k], sq); g[k] = e[k] > 0.f ? scale : (e[k] < 0.f ? -scale : 0.f); }
        if (d) reinterpret_cast<float4*>(d)[i] = make_float4(g[0], g[1], g[2], g[3]);
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float e = __ldg(a + i) - __ldg(b + i);
        acc += fabsf(e); sq = fmaf(e, e, sq);
        if (d) d[i] = e > 0.f ? scale : (e < 0.f ? -scale : 0.f);
    }
    l1_block_finish(acc, sq, scale, loss, sse);
}

// The same against the ground truth as the dataset holds it: uint8 HWC (a PIL image, scene/dataset_readers.py:1041), converted
// on the device exactly like utils/general_utils.py:PILtoTorch does on the host (float(u8) / 255.0f, IEEE division): the view's
// ground truth crosses PCIe as 3 B/pixel instead of 12.  render / d are CHW.
__global__ void __launch_bounds__(256) l1_fwd_bwd_u8_kernel(int H, int W, const float* __restrict__ a, const unsigned char* __restrict__ gt,
                                                            float scale, float* __restrict__ loss, float* __restrict__ sse,
                                                            float* __restrict__ d)
{
    float acc = 0.f, sq = 0.f;
    const long long n = (long long)H * W;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t = __fdiv_rn((float)__ldg(gt + 3 * i + c), 255.0f);
            const float e = __ldg(a + c * n + i) - t;
            acc += fabsf(e); sq = fmaf(e, e, sq);
            if (d) d[c * n + i] = e > 0.f ? scale : (e < 0.f ? -scale : 0.f);
        }
    }
    l1_block_finish(acc, sq, scale, loss, sse);
}

// Is every element of x bit-identical to x[0]?  out[0] = 1 / 0, out[1] = bits of x[0].  (gaussian_renderer/__init__.py:56 repeats
// the camera's one timestamp into a [P,1] tensor; knowing that lets the field serve the time planes from shared memory.)
__global__ void __launch_bounds__(256) uniform_check_kernel(long long n, const unsigned int* __restrict__ x, unsigned int* __restrict__ out)
{
    const unsigned int first = __ldg(x);
    bool same = true;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) same &= __ldg(x + i) == first;
    if (!__all_sync(0xffffffffu, same) && (threadIdx.x & 31) == 0) out[0] = 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = first;
}

// to8b of render_4DGS.py:49 / train_4DGS.py:335 on the device: out[y][x][c] = (uint8)(255 * clip(img[c][y][x], 0, 1))
// (truncation, like numpy's astype), CHW float -> HWC bytes; 12 B read + 3 B written per pixel.
__global__ void __launch_bounds__(256) to8b_hwc_kernel(int H, int W, const float* __restrict__ img, unsigned char* __restrict__ out)
{
    const long long n = (long long)H * W;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        unsigned char v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = fminf(fmaxf(__ldg(img + c * n + i), 0.f), 1.f);
            v[c] = (unsigned char)(255.f * x);
        }
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}

}  // namespace
}  // namespace b200gs

using namespace b200gs;

extern "C" {

int b200gs_activations_forward(long long P, const float* scales_raw, const float* rot_raw, const float* opacity_raw,
                               float* scales_out, float* rot_out, float* opacity_out, b200gs_stream_t stream)
{
    if (P <= 0) return 0;
    if (!scales_raw || !rot_raw || !opacity_raw || !scales_out || !rot_out || !opacity_out) { set_error("activations_forward: null pointer"); return -1; }
    activations_fwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, scales_raw, rot_raw, opacity_raw, scales_out, rot_out, opacity_out);
    return check_launch("activations_forward");
}

int b200gs_activations_backward(long long P, const float* scales_out, const float* rot_raw, const float* opacity_out,
                                const float* d_scales_out, const float* d_rot_out, const float* d_opacity_out,
                                float* d_scales_raw, float* d_rot_raw, float* d_opacity_raw, b200gs_stream_t stream)
{
    if (P <= 0) return 0;
    if (!scales_out || !rot_raw || !opacity_out) { set_error("activations_backward: null pointer"); return -1; }
    activations_bwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, scales_out, rot_raw, opacity_out, d_scales_out, d_rot_out,
                                                                                           d_opacity_out, d_scales_raw, d_rot_raw, d_opacity_raw);
    return check_launch("activations_backward");
}

int b200gs_l1_loss_fwd_bwd(long long n, const float* render, const float* target, float scale, float* loss_accum, float* sse_accum,
                           float* d_render, b200gs_stream_t stream)
{
    if (n <= 0) return 0;
    if (!render || !target || !loss_accum) { set_error("l1_loss: null pointer"); return -1; }
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > (long long)NUM_SMS * 8) blocks = (long long)NUM_SMS * 8;
    if (blocks < 1) blocks = 1;
    l1_fwd_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, render, target, scale, loss_accum, sse_accum, d_render);
    return check_launch("l1_loss");
}

int b200gs_l1_loss_fwd_bwd_u8(int H, int W, const float* render_chw, const unsigned char* target_hwc, float scale, float* loss_accum,
                              float* sse_accum, float* d_render_chw, b200gs_stream_t stream)
{
    if (H <= 0 || W <= 0) return 0;
    if (!render_chw || !target_hwc || !loss_accum) { set_error("l1_loss_u8: null pointer"); return -1; }
    long long blocks = ((long long)H * W + 255) / 256;
    if (blocks > (long long)NUM_SMS * 8) blocks = (long long)NUM_SMS * 8;
    l1_fwd_bwd_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(H, W, render_chw, target_hwc, scale, loss_accum, sse_accum, d_render_chw);
    return check_launch("l1_loss_u8");
}

int b200gs_uniform_value(long long n, const float* x, unsigned int* scratch_dev2, unsigned int* host_pinned2, b200gs_stream_t stream)
{
    if (n <= 0 || !x || !scratch_dev2 || !host_pinned2) { set_error("uniform_value: empty input or null pointer"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned int init[2] = {1u, 0u};
    host_pinned2[0] = init[0]; host_pinned2[1] = init[1];
    if (cudaMemcpyAsync(scratch_dev2, host_pinned2, 8, cudaMemcpyHostToDevice, st) != cudaSuccess) return check_launch("uniform_value") ? -1 : -1;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)NUM_SMS * 8) blocks = (long long)NUM_SMS * 8;
    uniform_check_kernel<<<(unsigned)blocks, 256, 0, st>>>(n, reinterpret_cast<const unsigned int*>(x), scratch_dev2);
    if (cudaMemcpyAsync(host_pinned2, scratch_dev2, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        check_launch("uniform_value");
        return -1;
    }
    return check_launch("uniform_value");
}

int b200gs_to8b_hwc(int H, int W, const float* image_chw, unsigned char* out_hwc, b200gs_stream_t stream)
{
    if (H <= 0 || W <= 0) return 0;
    if (!image_chw || !out_hwc) { set_error("to8b_hwc: null pointer"); return -1; }
    long long blocks = ((long long)H * W + 255) / 256;
    if (blocks > (long long)NUM_SMS * 8) blocks = (long long)NUM_SMS * 8;
    to8b_hwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(H, W, image_chw, out_hwc);
    return check_launch("to8b_hwc");
}

}  // extern "C"
