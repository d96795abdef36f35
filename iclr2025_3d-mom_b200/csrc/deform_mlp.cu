// Deformation MLP (feature decoder + position / scale / rotation heads), forward and backward.
//
// Restates scene/deformation.py:55-65 (create_net), :68-85 (query_time) and :97-153
// (forward_dynamic) for the configuration the reference trains with:
//     hidden = Linear(F, 64)(feature)                                   feature_out, defor_depth <= 1
//     d{x,s,r} = Linear(64, k)(ReLU(Linear(64, 64)(ReLU(hidden))))      k = 3, 3, 4
//     pts   = xyz * 1 + (dx + delta_scale * (frame_num * scene_flow))
//     scale = scales * 1 + ds ;  rot = rotations + dr
// The reference runs seven cuBLAS SGEMMs with every [P,64] intermediate round-tripping HBM
// plus ~20 elementwise launches; here a persistent CTA keeps all weights in shared memory,
// walks 64-point tiles through register-tiled FP32 GEMMs, and only the ReLU'd activations
// needed by the backward pass (4 x 256 B per point) leave the SM.  The backward pass
// accumulates every weight gradient in registers across all tiles of a CTA and flushes them
// with one atomic per element per CTA.
#include "common.cuh"
#include "../../include/b200gs.h"

namespace b200gs {

namespace {

constexpr int TM = 64;        // points per tile
constexpr int MW = 64;        // net_width
constexpr int MT = 256;       // threads per CTA
constexpr int LDR = MW + 4;   // row-major tile stride (floats)

struct FwdArgs {
    b200gs_mlp_weights w;
    long long P;
    const float* feat; const float* xyz; const float* scales; const float* rot; const float* scene_flow;
    float frame_num, delta_scale;
    const float* frame_num_dev;   // optional device scalar overriding frame_num (avoids a host sync)
    float* pts_out; float* scales_out; float* rot_out;
    float* saved;            // [4][tiles][64][TM]: relu(hidden), relu(z_pos), relu(z_scale), relu(z_rot)
};

// acc[i][j] += sum_k At[k][r0+i] * B[k][c0+j]
template <int K, int LDB>
__device__ __forceinline__ void gemm_AtB(const float* __restrict__ At, const float* __restrict__ B,
                                         float acc[4][4], int r0, int c0)
{
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(At + k * TM + r0);
        const float4 b = *reinterpret_cast<const float4*>(B + k * LDB + c0);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// transposed feature tile: At[k][row] <- feat[row0 + row][k]
template <int F>
__device__ __forceinline__ void load_feat_tile(float* __restrict__ At, const float* __restrict__ feat,
                                               long long row0, long long P)
{
    const int row = threadIdx.x & 63, kq = threadIdx.x >> 6;
    const long long g = row0 + row;
    for (int kk = kq; kk < F / 4; kk += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < P) v = __ldg(reinterpret_cast<const float4*>(feat + (size_t)g * F) + kk);
        At[(kk * 4 + 0) * TM + row] = v.x;
        At[(kk * 4 + 1) * TM + row] = v.y;
        At[(kk * 4 + 2) * TM + row] = v.z;
        At[(kk * 4 + 3) * TM + row] = v.w;
    }
}

template <int F>
__global__ void __launch_bounds__(MT, 1) deform_mlp_fwd_kernel(const __grid_constant__ FwdArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float* W1t = smem;                          // [F][64]
    float* W2t = W1t + F * MW;                  // [3][64][64]
    float* W3t = W2t + 3 * MW * MW;             // [3][64][4]
    float* B1 = W3t + 3 * MW * 4;               // [64]
    float* B2 = B1 + MW;                        // [3][64]
    float* B3 = B2 + 3 * MW;                    // [3][4]
    float* At = B3 + 16;                        // [F][TM]
    float* Ht = At + F * TM;                    // [64][TM]
    float* Zt = Ht + MW * TM;                   // [64][TM]
    const int tid = threadIdx.x;
    const int kdim[3] = {3, 3, 4};

    for (int i = tid; i < MW * F; i += MT) { int o = i / F, in = i - o * F; W1t[in * MW + o] = __ldg(a.w.w1 + i); }
    for (int i = tid; i < MW; i += MT) B1[i] = __ldg(a.w.b1 + i);
    for (int h = 0; h < 3; ++h) {
        if (!a.w.w2[h]) continue;
        for (int i = tid; i < MW * MW; i += MT) { int o = i >> 6, in = i & 63; W2t[h * MW * MW + in * MW + o] = __ldg(a.w.w2[h] + i); }
        for (int i = tid; i < MW; i += MT) B2[h * MW + i] = __ldg(a.w.b2[h] + i);
        for (int i = tid; i < MW * 4; i += MT) {
            int in = i >> 2, o = i & 3;
            W3t[h * MW * 4 + i] = o < kdim[h] ? __ldg(a.w.w3[h] + o * MW + in) : 0.f;
        }
        if (tid < 4) B3[h * 4 + tid] = tid < kdim[h] ? __ldg(a.w.b3[h] + tid) : 0.f;
    }
    __syncthreads();

    const int r0 = (tid & 15) * 4, c0 = (tid >> 4) * 4;
    const long long ntiles = (a.P + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        load_feat_tile<F>(At, a.feat, row0, a.P);
        __syncthreads();
        {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = B1[c0 + j];
            gemm_AtB<F, MW>(At, W1t, acc, r0, c0);
            float* sv = a.saved + ((size_t)0 * ntiles + tile) * MW * TM;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
                *reinterpret_cast<float4*>(Ht + (c0 + j) * TM + r0) = v;
                *reinterpret_cast<float4*>(sv + (c0 + j) * TM + r0) = v;
            }
        }
        __syncthreads();
        for (int h = 0; h < 3; ++h) {
            const int row = tid & 63, j = tid >> 6;
            const long long g = row0 + row;
            if (a.w.w2[h]) {
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = B2[h * MW + c0 + jj];
                gemm_AtB<MW, MW>(Ht, W2t + h * MW * MW, acc, r0, c0);
                float* sv = a.saved + ((size_t)(1 + h) * ntiles + tile) * MW * TM;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 v = make_float4(fmaxf(acc[0][jj], 0.f), fmaxf(acc[1][jj], 0.f), fmaxf(acc[2][jj], 0.f), fmaxf(acc[3][jj], 0.f));
                    *reinterpret_cast<float4*>(Zt + (c0 + jj) * TM + r0) = v;
                    *reinterpret_cast<float4*>(sv + (c0 + jj) * TM + r0) = v;
                }
                __syncthreads();
                if (j < kdim[h] && g < a.P) {
                    float o = B3[h * 4 + j];
                    const float* w3 = W3t + h * MW * 4 + j;
#pragma unroll 8
                    for (int k = 0; k < MW; ++k) o = fmaf(Zt[k * TM + row], w3[k * 4], o);
                    if (h == 0) {
                        const float fn = a.frame_num_dev ? __ldg(a.frame_num_dev) : a.frame_num;
                        const float flow = __fmul_rn(a.delta_scale, __fmul_rn(fn, __ldg(a.scene_flow + 3 * g + j)));
                        a.pts_out[3 * g + j] = __fadd_rn(__fmul_rn(__ldg(a.xyz + 3 * g + j), 1.0f), __fadd_rn(o, flow));
                    } else if (h == 1) {
                        a.scales_out[3 * g + j] = __fadd_rn(__fmul_rn(__ldg(a.scales + 3 * g + j), 1.0f), o);
                    } else {
                        a.rot_out[4 * g + j] = __fadd_rn(__ldg(a.rot + 4 * g + j), o);
                    }
                }
                __syncthreads();
            } else if (g < a.P) {       // head disabled (no_dx / no_ds / no_dr): pass through
                if (h == 0 && j < 3) a.pts_out[3 * g + j] = __ldg(a.xyz + 3 * g + j);
                if (h == 1 && j < 3) a.scales_out[3 * g + j] = __ldg(a.scales + 3 * g + j);
                if (h == 2 && j < 4) a.rot_out[4 * g + j] = __ldg(a.rot + 4 * g + j);
            }
        }
    }
}

struct BwdArgs {
    b200gs_mlp_weights w;
    b200gs_mlp_grads gw;
    long long P;
    const float* feat; const float* saved;
    const float* d_pts; const float* d_scales; const float* d_rot;
    float* d_feat;
};

// acc[i][j] += sum_r dYrm[r][o0+i] * Xt[i0+j][r]     (o: rows of dW = out features, i: in features)
__device__ __forceinline__ void outer_acc(const float* __restrict__ dYrm, const float* __restrict__ Xt,
                                          float acc[4][4], int o0, int i0)
{
#pragma unroll 2
    for (int r = 0; r < TM; r += 4) {
        float x[4][4];     // x[j][rr] = Xt[i0+j][r+rr]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(Xt + (i0 + j) * TM + r);
            x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const float4 dy = *reinterpret_cast<const float4*>(dYrm + (r + rr) * LDR + o0);
            const float d[4] = {dy.x, dy.y, dy.z, dy.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], x[j][rr], acc[i][j]);
        }
    }
}

template <int F>
__global__ void __launch_bounds__(MT, 1) deform_mlp_bwd_kernel(const __grid_constant__ BwdArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float* W1 = smem;                           // [64][F]   (out, in) as stored by torch
    float* W2 = W1 + MW * F;                    // [3][64][64]
    float* W3 = W2 + 3 * MW * MW;               // [3][4][64]
    float* At = W3 + 3 * 4 * MW;                // [F][TM]   feature tile, transposed
    float* Ht = At + F * TM;                    // [64][TM]  relu(hidden), transposed
    float* Zt = Ht + MW * TM;                   // [64][TM]  relu(z) of the current head
    float* Dt = Zt + MW * TM;                   // [64][TM]  dz (then d hidden), transposed
    float* Drm = Dt + MW * TM;                  // [TM][LDR] same, row-major
    float* Dout = Drm + TM * LDR;               // [TM][4]   upstream gradient of the current head
    const int tid = threadIdx.x;
    const int kdim[3] = {3, 3, 4};

    for (int i = tid; i < MW * F; i += MT) W1[i] = __ldg(a.w.w1 + i);
    for (int h = 0; h < 3; ++h) {
        if (!a.w.w2[h]) continue;
        for (int i = tid; i < MW * MW; i += MT) W2[h * MW * MW + i] = __ldg(a.w.w2[h] + i);
        for (int i = tid; i < 4 * MW; i += MT) { int o = i >> 6; W3[h * 4 * MW + i] = o < kdim[h] ? __ldg(a.w.w3[h] + i) : 0.f; }
    }
    __syncthreads();

    const int r0 = (tid & 15) * 4, c0 = (tid >> 4) * 4;      // C-tile of the dX GEMMs (rows, cols)
    const int o0 = (tid & 15) * 4, i0 = (tid >> 4) * 4;      // block of the dW accumulators (out, in)
    float gW2[3][4][4], gW1[F / MW][4][4];
    float gW3[3][4][4];      // [head][out j][in c0+i], partial over this thread's rows
    float gB3[3] = {0.f, 0.f, 0.f}, gB2[3] = {0.f, 0.f, 0.f}, gB1 = 0.f;
#pragma unroll
    for (int h = 0; h < 3; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) gW3[h][i][j] = 0.f;
#pragma unroll
    for (int h = 0; h < 3; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) gW2[h][i][j] = 0.f;
#pragma unroll
    for (int q = 0; q < F / MW; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) gW1[q][i][j] = 0.f;

    const long long ntiles = (a.P + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        load_feat_tile<F>(At, a.feat, row0, a.P);
        {
            const float4* sv = reinterpret_cast<const float4*>(a.saved + ((size_t)0 * ntiles + tile) * MW * TM);
            for (int i = tid; i < MW * TM / 4; i += MT) reinterpret_cast<float4*>(Ht)[i] = __ldg(sv + i);
        }
        float dRh[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dRh[i][j] = 0.f;

#pragma unroll
        for (int h = 0; h < 3; ++h) {
            if (!a.w.w2[h]) continue;
            {
                const float4* sv = reinterpret_cast<const float4*>(a.saved + ((size_t)(1 + h) * ntiles + tile) * MW * TM);
                for (int i = tid; i < MW * TM / 4; i += MT) reinterpret_cast<float4*>(Zt)[i] = __ldg(sv + i);
                const float* dsrc = h == 0 ? a.d_pts : (h == 1 ? a.d_scales : a.d_rot);
                const int kd = kdim[h];
                const int row = tid >> 2, j = tid & 3;
                const long long g = row0 + row;
                Dout[tid] = (j < kd && g < a.P && dsrc) ? __ldg(dsrc + (size_t)g * kd + j) : 0.f;
            }
            __syncthreads();
            // dz = (d_out . W3) masked by relu(z) > 0, in both layouts
            {
                float dA[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 dv = *reinterpret_cast<const float4*>(Dout + (r0 + i) * 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float* w3 = W3 + h * 4 * MW + c0 + j;
                        dA[i][j] = dv.x * w3[0] + dv.y * w3[MW] + dv.z * w3[2 * MW] + dv.w * w3[3 * MW];
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 z = *reinterpret_cast<const float4*>(Zt + (c0 + j) * TM + r0);
                    dA[0][j] = z.x > 0.f ? dA[0][j] : 0.f;
                    dA[1][j] = z.y > 0.f ? dA[1][j] : 0.f;
                    dA[2][j] = z.z > 0.f ? dA[2][j] : 0.f;
                    dA[3][j] = z.w > 0.f ? dA[3][j] : 0.f;
                    *reinterpret_cast<float4*>(Dt + (c0 + j) * TM + r0) = make_float4(dA[0][j], dA[1][j], dA[2][j], dA[3][j]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(Drm + (r0 + i) * LDR + c0) = make_float4(dA[i][0], dA[i][1], dA[i][2], dA[i][3]);
            }
            // dW3[j][in] partial over this thread's 4 rows (kept per thread across tiles), db3[j]
            {
                float zv[4][4];    // zv[i][rr] = relu(z)[r0+rr][c0+i]
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(Zt + (c0 + i) * TM + r0);
                    zv[i][0] = v.x; zv[i][1] = v.y; zv[i][2] = v.z; zv[i][3] = v.w;
                }
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const float4 dv = *reinterpret_cast<const float4*>(Dout + (r0 + rr) * 4);
                    const float d[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) gW3[h][j][i] = fmaf(d[j], zv[i][rr], gW3[h][j][i]);
                }
                if (tid < kdim[h]) {
                    float sacc = 0.f;
                    for (int r = 0; r < TM; ++r) sacc += Dout[r * 4 + tid];
                    gB3[h] += sacc;
                }
            }
            __syncthreads();
            outer_acc(Drm, Ht, gW2[h], o0, i0);                       // dW2 += dz^T relu(h)
            if (tid < MW) {
                float s = 0.f;
#pragma unroll 8
                for (int r = 0; r < TM; ++r) s += Drm[r * LDR + tid];
                gB2[h] += s;
            }
            gemm_AtB<MW, MW>(Dt, W2 + h * MW * MW, dRh, r0, c0);       // d relu(h) += dz W2
            __syncthreads();
        }
        // d hidden = d relu(h) masked by relu(h) > 0
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 hv = *reinterpret_cast<const float4*>(Ht + (c0 + j) * TM + r0);
            dRh[0][j] = hv.x > 0.f ? dRh[0][j] : 0.f;
            dRh[1][j] = hv.y > 0.f ? dRh[1][j] : 0.f;
            dRh[2][j] = hv.z > 0.f ? dRh[2][j] : 0.f;
            dRh[3][j] = hv.w > 0.f ? dRh[3][j] : 0.f;
            *reinterpret_cast<float4*>(Dt + (c0 + j) * TM + r0) = make_float4(dRh[0][j], dRh[1][j], dRh[2][j], dRh[3][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(Drm + (r0 + i) * LDR + c0) = make_float4(dRh[i][0], dRh[i][1], dRh[i][2], dRh[i][3]);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < F / MW; ++q) outer_acc(Drm, At + q * MW * TM, gW1[q], o0, i0);     // dW1 += dh^T feat
        if (tid < MW) {
            float s = 0.f;
#pragma unroll 8
            for (int r = 0; r < TM; ++r) s += Drm[r * LDR + tid];
            gB1 += s;
        }
        // d feature = dh W1   ([TM x 64] x [64 x F]), written row-major
#pragma unroll
        for (int q = 0; q < F / MW; ++q) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            gemm_AtB<MW, F>(Dt, W1 + q * MW, acc, r0, c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long g = row0 + r0 + i;
                if (g < a.P)
                    *reinterpret_cast<float4*>(a.d_feat + (size_t)g * F + q * MW + c0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            }
        }
        __syncthreads();
    }

    // flush the per-CTA weight gradients
#pragma unroll
    for (int q = 0; q < F / MW; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(a.gw.w1 + (size_t)(o0 + i) * F + q * MW + i0 + j, gW1[q][i][j]);
    if (tid < MW) atomicAdd(a.gw.b1 + tid, gB1);
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        if (!a.w.w2[h]) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(a.gw.w2[h] + (o0 + i) * MW + i0 + j, gW2[h][i][j]);
        if (tid < MW) atomicAdd(a.gw.b2[h] + tid, gB2[h]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < kdim[h]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) atomicAdd(a.gw.w3[h] + j * MW + c0 + i, gW3[h][j][i]);
            }
        if (tid < kdim[h]) atomicAdd(a.gw.b3[h] + tid, gB3[h]);
    }
}

size_t fwd_smem(int F) { return (size_t)(F * MW + 3 * MW * MW + 3 * MW * 4 + MW + 3 * MW + 16 + F * TM + 2 * MW * TM) * sizeof(float); }
size_t bwd_smem(int F) { return (size_t)(MW * F + 3 * MW * MW + 3 * 4 * MW + F * TM + 3 * MW * TM + TM * LDR + TM * 4) * sizeof(float); }

int check_weights(const b200gs_mlp_weights* w)
{
    if (!w) { set_error("deform_mlp: null weights"); return -1; }
    if (w->width != MW) { set_error("deform_mlp: net_width=%d unsupported (need %d)", w->width, MW); return -1; }
    if (w->feat_dim != 32 && w->feat_dim != 64 && w->feat_dim != 96 && w->feat_dim != 128) { set_error("deform_mlp: feature dim %d unsupported", w->feat_dim); return -1; }
    if (w->feat_dim == 32 || w->feat_dim == 96) { set_error("deform_mlp: feature dim %d (1 or 3 levels) not compiled in", w->feat_dim); return -1; }
    if (!w->w1 || !w->b1) { set_error("deform_mlp: feature_out weights missing"); return -1; }
    for (int h = 0; h < 3; ++h)
        if (w->w2[h] && (!w->b2[h] || !w->w3[h] || !w->b3[h])) { set_error("deform_mlp: head %d incomplete", h); return -1; }
    return 0;
}

}  // namespace
}  // namespace b200gs

using namespace b200gs;

extern "C" {

size_t b200gs_deform_mlp_saved_floats(long long P)
{
    const long long ntiles = (P + TM - 1) / TM;
    return (size_t)4 * ntiles * MW * TM;
}

int b200gs_deform_mlp_forward(const b200gs_mlp_weights* w, long long P, const float* feat, const float* xyz,
                              const float* scales, const float* rot, const float* scene_flow, float frame_num,
                              const float* frame_num_dev, float delta_scale, float* pts_out, float* scales_out, float* rot_out, float* saved,
                              b200gs_stream_t stream)
{
    if (check_weights(w)) return -1;
    if (P <= 0) return 0;
    FwdArgs a;
    a.w = *w; a.P = P; a.feat = feat; a.xyz = xyz; a.scales = scales; a.rot = rot; a.scene_flow = scene_flow;
    a.frame_num = frame_num; a.frame_num_dev = frame_num_dev; a.delta_scale = delta_scale; a.pts_out = pts_out; a.scales_out = scales_out;
    a.rot_out = rot_out; a.saved = saved;
    const long long ntiles = (P + TM - 1) / TM;
    const int grid = (int)(ntiles < NUM_SMS ? ntiles : NUM_SMS);
    const size_t smem = fwd_smem(w->feat_dim);
    if (w->feat_dim == 64) {
        cudaFuncSetAttribute(deform_mlp_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        deform_mlp_fwd_kernel<64><<<grid, MT, smem, (cudaStream_t)stream>>>(a);
    } else {
        cudaFuncSetAttribute(deform_mlp_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        deform_mlp_fwd_kernel<128><<<grid, MT, smem, (cudaStream_t)stream>>>(a);
    }
    return check_launch("deform_mlp_forward");
}

int b200gs_deform_mlp_backward(const b200gs_mlp_weights* w, const b200gs_mlp_grads* gw, long long P, const float* feat,
                               const float* saved, const float* d_pts, const float* d_scales, const float* d_rot,
                               float* d_feat, b200gs_stream_t stream)
{
    if (check_weights(w)) return -1;
    if (!gw) { set_error("deform_mlp_backward: null gradient table"); return -1; }
    if (P <= 0) return 0;
    BwdArgs a;
    a.w = *w; a.gw = *gw; a.P = P; a.feat = feat; a.saved = saved; a.d_pts = d_pts; a.d_scales = d_scales;
    a.d_rot = d_rot; a.d_feat = d_feat;
    const long long ntiles = (P + TM - 1) / TM;
    const int grid = (int)(ntiles < NUM_SMS ? ntiles : NUM_SMS);
    const size_t smem = bwd_smem(w->feat_dim);
    if (w->feat_dim == 64) {
        cudaFuncSetAttribute(deform_mlp_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        deform_mlp_bwd_kernel<64><<<grid, MT, smem, (cudaStream_t)stream>>>(a);
    } else {
        cudaFuncSetAttribute(deform_mlp_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        deform_mlp_bwd_kernel<128><<<grid, MT, smem, (cudaStream_t)stream>>>(a);
    }
    return check_launch("deform_mlp_backward");
}

}  // extern "C"
