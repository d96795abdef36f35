// Fused multi-tensor Adam step and multi-tensor row gather (densify / prune bookkeeping).
//
// Adam follows the arithmetic of torch.optim.Adam's foreach path that the reference runs
// (scene/gaussian_model.py:209 -> torch/optim/adam.py::_multi_tensor_adam), element for
// element and rounding for rounding:
//     m = m + (1-b1) * (g - m)                      (_foreach_lerp_)
//     v = v * b2 ; v = v + (1-b2) * (g * g)         (_foreach_mul_, _foreach_addcmul_)
//     d = sqrt(v) / sqrt(1 - b2^t) + eps            (_foreach_sqrt, _foreach_div_, _foreach_add_)
//     p = p + (-lr / (1 - b1^t)) * (m / d)          (_foreach_addcdiv_)
// but as ONE launch over every parameter tensor (28 B of HBM traffic per parameter: p, m, v
// read + written, g read) instead of ~8 launches and ~80 B per parameter.
#include "common.cuh"
#include "../../include/b200gs.h"

namespace b200gs {

namespace {

constexpr int AD_THREADS = 256;
constexpr int AD_VEC = 4;
constexpr int AD_CHUNK = AD_THREADS * AD_VEC * 4;    // 4096 elements per block

struct AdamTable {
    b200gs_adam_tensor t[B200GS_ADAM_MAX_TENSORS];
    unsigned int first_block[B200GS_ADAM_MAX_TENSORS + 1];
    int n;
    float beta1_c;     // 1 - beta1, as float (the foreach kernels cast their double scalar to float)
    float beta2;
    float beta2_c;     // 1 - beta2
    float eps;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w1, float b2, float w2,
                                         float eps, float neg_step, float bc2_sqrt)
{
    m = __fmaf_rn(w1, __fsub_rn(g, m), m);
    v = __fmul_rn(v, b2);
    v = __fmaf_rn(w2, __fmul_rn(g, g), v);
    const float d = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = __fmaf_rn(neg_step, __fdiv_rn(m, d), p);
}

__global__ void __launch_bounds__(AD_THREADS) adam_multi_kernel(const __grid_constant__ AdamTable tab)
{
    // which tensor does this block belong to? (<= 64 entries: linear scan in registers is fine)
    const unsigned int b = blockIdx.x;
    int ti = 0;
    while (ti + 1 < tab.n && tab.first_block[ti + 1] <= b) ++ti;
    const b200gs_adam_tensor& T = tab.t[ti];
    const size_t base = (size_t)(b - tab.first_block[ti]) * AD_CHUNK;
    const size_t n = (size_t)T.numel;
    float* __restrict__ p = T.param; const float* __restrict__ g = T.grad;
    float* __restrict__ m = T.exp_avg; float* __restrict__ v = T.exp_avg_sq;
    const float neg_step = T.neg_step_size, bcs = T.bias_correction2_sqrt;
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t i = base + ((size_t)k * AD_THREADS + threadIdx.x) * AD_VEC;
        if (i >= n) break;
        if (aligned && i + AD_VEC <= n) {
            float4 P4 = *reinterpret_cast<float4*>(p + i);
            const float4 G4 = __ldg(reinterpret_cast<const float4*>(g + i));
            float4 M4 = *reinterpret_cast<float4*>(m + i);
            float4 V4 = *reinterpret_cast<float4*>(v + i);
            adam_one(P4.x, G4.x, M4.x, V4.x, tab.beta1_c, tab.beta2, tab.beta2_c, tab.eps, neg_step, bcs);
            adam_one(P4.y, G4.y, M4.y, V4.y, tab.beta1_c, tab.beta2, tab.beta2_c, tab.eps, neg_step, bcs);
            adam_one(P4.z, G4.z, M4.z, V4.z, tab.beta1_c, tab.beta2, tab.beta2_c, tab.eps, neg_step, bcs);
            adam_one(P4.w, G4.w, M4.w, V4.w, tab.beta1_c, tab.beta2, tab.beta2_c, tab.eps, neg_step, bcs);
            *reinterpret_cast<float4*>(p + i) = P4;
            *reinterpret_cast<float4*>(m + i) = M4;
            *reinterpret_cast<float4*>(v + i) = V4;
        } else {
            for (size_t j = i; j < n && j < i + AD_VEC; ++j) {
                float pp = p[j], mm = m[j], vv = v[j];
                adam_one(pp, g[j], mm, vv, tab.beta1_c, tab.beta2, tab.beta2_c, tab.eps, neg_step, bcs);
                p[j] = pp; m[j] = mm; v[j] = vv;
            }
        }
    }
}

// Adam over the two SH parameter tensors of scene/gaussian_model.py:136-140, _features_dc [P,1,3] and _features_rest
// [P,M-1,3], whose gradient arrives as ONE [P,M,3] buffer (what the rasterizer backward writes, backward.cu:20-139): the
// split / copy of the gradient disappears and the two tensors share one launch.  One thread per gradient element.
struct AdamShArgs {
    float* p_dc; float* m_dc; float* v_dc; float* p_rest; float* m_rest; float* v_rest;
    const float* grad; long long P; int M;
    float neg_step_dc, bcs_dc, neg_step_rest, bcs_rest, beta1_c, beta2, beta2_c, eps;
};

__global__ void __launch_bounds__(256) adam_sh_kernel(const __grid_constant__ AdamShArgs a)
{
    const long long row = 3LL * a.M, total = a.P * row;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long pt = e / row;
        const int c = (int)(e - pt * row);
        const float g = __ldg(a.grad + e);
        const bool dc = c < 3;
        const long long i = dc ? pt * 3 + c : pt * (row - 3) + (c - 3);
        float* p = (dc ? a.p_dc : a.p_rest) + i; float* m = (dc ? a.m_dc : a.m_rest) + i; float* v = (dc ? a.v_dc : a.v_rest) + i;
        float pp = *p, mm = *m, vv = *v;
        adam_one(pp, g, mm, vv, a.beta1_c, a.beta2, a.beta2_c, a.eps, dc ? a.neg_step_dc : a.neg_step_rest, dc ? a.bcs_dc : a.bcs_rest);
        *p = pp; *m = mm; *v = vv;
    }
}

// ---- multi-tensor row gather: dst_t[i, :] = src_t[index[i], :] for every tensor t --------
struct GatherTable {
    b200gs_gather_tensor t[B200GS_GATHER_MAX_TENSORS];
    unsigned long long first_elem[B200GS_GATHER_MAX_TENSORS + 1];   // prefix of n_out * row_floats
    int n;
    long long n_out;
    const long long* index;    // int64 row indices (what torch.nonzero gives), or null = identity
};

__global__ void __launch_bounds__(256) gather_rows_kernel(const __grid_constant__ GatherTable tab)
{
    const unsigned long long total = tab.first_elem[tab.n];
    for (unsigned long long e = (unsigned long long)blockIdx.x * 256 + threadIdx.x; e < total;
         e += (unsigned long long)gridDim.x * 256) {
        int ti = 0;
        while (ti + 1 < tab.n && tab.first_elem[ti + 1] <= e) ++ti;
        const b200gs_gather_tensor& T = tab.t[ti];
        const unsigned long long local = e - tab.first_elem[ti];
        const unsigned long long row = local / (unsigned)T.row_floats;
        const unsigned int col = (unsigned int)(local - row * (unsigned)T.row_floats);
        if (T.zero_tail_rows > 0 && (long long)row >= tab.n_out - T.zero_tail_rows) { T.dst[local] = 0.f; continue; }     // new rows of an Adam moment
        const long long src_row = tab.index ? tab.index[row] : (long long)row;
        T.dst[local] = __ldg(T.src + (size_t)src_row * T.row_floats + col);
    }
}

// ---- densify / prune decisions (scene/gaussian_model.py:681-698, :511-523, :541-565, :713-715, :362-365) ----------------------
// One pass over the Gaussians each; the arithmetic follows the torch expressions of the reference operator for operator
// (accurate expf / logf / division, no fast-math), so the same Gaussians are selected.
__global__ void __launch_bounds__(256) densify_select_kernel(long long N, const float* __restrict__ accum, const float* __restrict__ denom,
                                                             const float* __restrict__ scaling, float thr, float dense_extent,
                                                             unsigned char* __restrict__ clone_flag, unsigned char* __restrict__ split_flag)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long long)gridDim.x * 256) {
        float g = __fdiv_rn(__ldg(accum + i), __ldg(denom + i));          // grads = xyz_gradient_accum / denom
        if (g != g) g = 0.f;                                              // grads[grads.isnan()] = 0.0
        const float smax = fmaxf(fmaxf(__ldg(scaling + 3 * i), __ldg(scaling + 3 * i + 1)), __ldg(scaling + 3 * i + 2));
        const float world = expf(smax);                                   // max(get_scaling): exp is monotone, max commutes with it
        clone_flag[i] = (fabsf(g) >= thr) && (world <= dense_extent);     // torch.norm(grads, dim=-1) >= thr  &  max scale <= percent_dense * extent
        split_flag[i] = (g >= thr) && (world > dense_extent);             // padded_grad >= thr  &  max scale > percent_dense * extent
    }
}

__global__ void __launch_bounds__(256) prune_select_kernel(long long N, const float* __restrict__ opacity, const float* __restrict__ scaling,
                                                           const float* __restrict__ max_radii2D, float min_opacity, float max_screen_size,
                                                           float max_world, unsigned char* __restrict__ prune_flag)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long long)gridDim.x * 256) {
        const float o = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__ldg(opacity + i))));      // torch.sigmoid
        bool p = o < min_opacity;
        if (max_screen_size > 0.f) {
            const float smax = fmaxf(fmaxf(__ldg(scaling + 3 * i), __ldg(scaling + 3 * i + 1)), __ldg(scaling + 3 * i + 2));
            p = p || (__ldg(max_radii2D + i) > max_screen_size) || (expf(smax) > max_world);
        }
        prune_flag[i] = p;
    }
}

__global__ void __launch_bounds__(256) densification_stats_kernel(long long N, const float* __restrict__ vgrad, const unsigned char* __restrict__ filter,
                                                                  float* __restrict__ accum, float* __restrict__ denom)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long long)gridDim.x * 256) {
        if (!filter[i]) continue;
        const float x = __ldg(vgrad + 3 * i), y = __ldg(vgrad + 3 * i + 1);
        accum[i] = __fadd_rn(accum[i], __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));     // += norm(grad[:, :2])
        denom[i] = __fadd_rn(denom[i], 1.f);
    }
}

__global__ void __launch_bounds__(256) reset_opacity_kernel(long long N, const float* __restrict__ opacity, float* __restrict__ out)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < N; i += (long long)gridDim.x * 256) {
        const float o = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__ldg(opacity + i))));
        const float c = fminf(o, 0.01f);                                  // torch.min(get_opacity, ones * 0.01)
        out[i] = logf(__fdiv_rn(c, __fsub_rn(1.f, c)));                   // inverse_sigmoid
    }
}

}  // namespace
}  // namespace b200gs

using namespace b200gs;

extern "C" {

int b200gs_adam_multi(int n_tensors, const b200gs_adam_tensor* tensors, double beta1, double beta2, double eps,
                      b200gs_stream_t stream)
{
    if (n_tensors < 0) { set_error("adam_multi: negative tensor count"); return -1; }
    for (int done = 0; done < n_tensors;) {
        AdamTable tab;
        const int n = (n_tensors - done) < B200GS_ADAM_MAX_TENSORS ? (n_tensors - done) : B200GS_ADAM_MAX_TENSORS;
        unsigned long long blocks = 0;
        int k = 0;
        for (int i = 0; i < n; ++i) {
            const b200gs_adam_tensor& T = tensors[done + i];
            if (T.numel < 0 || (T.numel > 0 && (!T.param || !T.grad || !T.exp_avg || !T.exp_avg_sq))) {
                set_error("adam_multi: tensor %d has null pointers", done + i); return -1;
            }
            if (T.numel == 0) continue;
            tab.t[k] = T;
            tab.first_block[k] = (unsigned int)blocks;
            blocks += ((unsigned long long)T.numel + AD_CHUNK - 1) / AD_CHUNK;
            ++k;
        }
        if (blocks > 0x7fffffffull) { set_error("adam_multi: too many elements for one launch"); return -1; }
        tab.first_block[k] = (unsigned int)blocks;
        tab.n = k;
        tab.beta1_c = (float)(1.0 - beta1);
        tab.beta2 = (float)beta2;
        tab.beta2_c = (float)(1.0 - beta2);
        tab.eps = (float)eps;
        if (k > 0) adam_multi_kernel<<<(unsigned)blocks, AD_THREADS, 0, (cudaStream_t)stream>>>(tab);
        done += n;
    }
    return check_launch("adam_multi");
}

int b200gs_adam_sh(long long P, int M, const b200gs_adam_tensor* dc, const b200gs_adam_tensor* rest, const float* grad_pm3,
                   double beta1, double beta2, double eps, b200gs_stream_t stream)
{
    if (P <= 0) return 0;
    if (M < 2 || !dc || !rest || !grad_pm3 || !dc->param || !dc->exp_avg || !dc->exp_avg_sq || !rest->param || !rest->exp_avg || !rest->exp_avg_sq) {
        set_error("adam_sh: null pointer or M < 2"); return -1;
    }
    if (dc->numel != 3 * P || rest->numel != 3LL * (M - 1) * P) { set_error("adam_sh: tensor sizes do not match [P,1,3] / [P,M-1,3]"); return -1; }
    AdamShArgs a;
    a.p_dc = dc->param; a.m_dc = dc->exp_avg; a.v_dc = dc->exp_avg_sq; a.p_rest = rest->param; a.m_rest = rest->exp_avg; a.v_rest = rest->exp_avg_sq;
    a.grad = grad_pm3; a.P = P; a.M = M;
    a.neg_step_dc = dc->neg_step_size; a.bcs_dc = dc->bias_correction2_sqrt; a.neg_step_rest = rest->neg_step_size; a.bcs_rest = rest->bias_correction2_sqrt;
    a.beta1_c = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.beta2_c = (float)(1.0 - beta2); a.eps = (float)eps;
    long long blocks = (3LL * M * P + 255) / 256;
    if (blocks > (long long)NUM_SMS * 16) blocks = (long long)NUM_SMS * 16;
    adam_sh_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("adam_sh");
}

int b200gs_gather_rows_multi(int n_tensors, const b200gs_gather_tensor* tensors, const long long* index,
                             long long n_out, b200gs_stream_t stream)
{
    if (n_out <= 0 || n_tensors <= 0) return 0;
    for (int done = 0; done < n_tensors;) {
        GatherTable tab;
        const int n = (n_tensors - done) < B200GS_GATHER_MAX_TENSORS ? (n_tensors - done) : B200GS_GATHER_MAX_TENSORS;
        unsigned long long total = 0;
        int k = 0;
        for (int i = 0; i < n; ++i) {
            const b200gs_gather_tensor& T = tensors[done + i];
            if (T.row_floats <= 0) continue;
            if (!T.src || !T.dst) { set_error("gather_rows_multi: tensor %d has null pointers", done + i); return -1; }
            tab.t[k] = T;
            tab.first_elem[k] = total;
            total += (unsigned long long)n_out * (unsigned long long)T.row_floats;
            ++k;
        }
        tab.first_elem[k] = total;
        tab.n = k; tab.n_out = n_out; tab.index = index;
        if (k > 0) {
            unsigned long long blocks = (total + 255) / 256;
            if (blocks > (unsigned long long)NUM_SMS * 32) blocks = (unsigned long long)NUM_SMS * 32;
            gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(tab);
        }
        done += n;
    }
    return check_launch("gather_rows_multi");
}

static unsigned blocks_for(long long n)
{
    long long b = (n + 255) / 256;
    if (b > (long long)NUM_SMS * 16) b = (long long)NUM_SMS * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

int b200gs_densify_select(long long N, const float* grad_accum, const float* denom, const float* scaling_raw, float grad_threshold,
                          float dense_extent, unsigned char* clone_flag, unsigned char* split_flag, b200gs_stream_t stream)
{
    if (N <= 0) return 0;
    if (!grad_accum || !denom || !scaling_raw || !clone_flag || !split_flag) { set_error("densify_select: null pointer"); return -1; }
    densify_select_kernel<<<blocks_for(N), 256, 0, (cudaStream_t)stream>>>(N, grad_accum, denom, scaling_raw, grad_threshold, dense_extent, clone_flag, split_flag);
    return check_launch("densify_select");
}

int b200gs_prune_select(long long N, const float* opacity_raw, const float* scaling_raw, const float* max_radii2D, float min_opacity,
                        float max_screen_size, float max_world_extent, unsigned char* prune_flag, b200gs_stream_t stream)
{
    if (N <= 0) return 0;
    if (!opacity_raw || !prune_flag || (max_screen_size > 0.f && (!scaling_raw || !max_radii2D))) { set_error("prune_select: null pointer"); return -1; }
    prune_select_kernel<<<blocks_for(N), 256, 0, (cudaStream_t)stream>>>(N, opacity_raw, scaling_raw, max_radii2D, min_opacity, max_screen_size, max_world_extent, prune_flag);
    return check_launch("prune_select");
}

int b200gs_densification_stats(long long N, const float* viewspace_grad, const unsigned char* update_filter, float* grad_accum, float* denom,
                               b200gs_stream_t stream)
{
    if (N <= 0) return 0;
    if (!viewspace_grad || !update_filter || !grad_accum || !denom) { set_error("densification_stats: null pointer"); return -1; }
    densification_stats_kernel<<<blocks_for(N), 256, 0, (cudaStream_t)stream>>>(N, viewspace_grad, update_filter, grad_accum, denom);
    return check_launch("densification_stats");
}

int b200gs_reset_opacity(long long N, const float* opacity_raw, float* out, b200gs_stream_t stream)
{
    if (N <= 0) return 0;
    if (!opacity_raw || !out) { set_error("reset_opacity: null pointer"); return -1; }
    reset_opacity_kernel<<<blocks_for(N), 256, 0, (cudaStream_t)stream>>>(N, opacity_raw, out);
    return check_launch("reset_opacity");
}

}  // extern "C"
