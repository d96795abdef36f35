// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// C-ABI shim around the *unmodified* reference rasterizer
// (/root/reference/submodules/depth-diff-gaussian-rasterization/cuda_rasterizer/*),
// compiled from the sources where they lie by oracle/build_ref.sh into
// oracle/_ref/libref_rast.so.  It plays the role of the reference's torch glue
// (rasterize_points.cu:35-202) without torch: it owns the three growable byte
// buffers (geometry / binning / image) and exposes their decoded fields so the
// parity tests can compare radii, keys, sorted lists and tile ranges bit-for-bit.
// Only tests/, __graft_entry__.smoke() and bench.py's reference leg may load it.
#include <cstdint>
#include <cfloat>
#include <cstring>
#include <functional>
#include <string>
#include <cuda_runtime.h>
#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"
#include "cuda_rasterizer/rasterizer_impl.h"

namespace {
struct Buf { char* p = nullptr; size_t cap = 0; size_t used = 0; };
Buf g_geom, g_bin, g_img;
int g_P = 0, g_R = 0, g_W = 0, g_H = 0;
std::string g_err;

std::function<char*(size_t)> grower(Buf& b) {
    return [&b](size_t n) -> char* {
        if (n > b.cap) {
            if (b.p) cudaFree(b.p);
            size_t cap = n + n / 4 + 1024;
            if (cudaMalloc(&b.p, cap) != cudaSuccess) { b.p = nullptr; b.cap = 0; return nullptr; }
            b.cap = cap;
        }
        b.used = n;
        return b.p;
    };
}
int fail(const char* what) {
    cudaError_t e = cudaGetLastError();
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Mirrors RasterizeGaussiansCUDA (rasterize_points.cu:35-117); null pointers stand for the
// empty tensors the Python layer passes for "None".
int ref_rast_forward(int P, int D, int M, const float* bg, int W, int H,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* scales, float scale_modifier,
                     const float* rotations, const float* cov3D_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* campos,
                     float tan_fovx, float tan_fovy, int prefiltered,
                     float* out_color, float* out_depth, int* radii, int debug)
{
    g_P = P; g_W = W; g_H = H; g_R = 0;
    if (P == 0) return 0;
    try {
        g_R = CudaRasterizer::Rasterizer::forward(
            grower(g_geom), grower(g_bin), grower(g_img), P, D, M, bg, W, H, means3D, shs,
            colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp,
            viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, prefiltered != 0,
            out_color, out_depth, radii, debug != 0);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    if (cudaPeekAtLastError() != cudaSuccess) return fail("ref forward");
    return g_R;
}

// Mirrors RasterizeGaussiansBackwardCUDA (rasterize_points.cu:119-202). All ten gradient
// outputs must be zero-filled by the caller, exactly as torch::zeros does there (:154-163).
int ref_rast_backward(int P, int D, int M, int R, const float* bg, int W, int H,
                      const float* means3D, const float* shs, const float* colors_precomp,
                      const float* scales, float scale_modifier, const float* rotations,
                      const float* cov3D_precomp, const float* viewmatrix,
                      const float* projmatrix, const float* campos, float tan_fovx,
                      float tan_fovy, const int* radii, const float* dL_dpix,
                      const float* dL_dpix_depth, float* dL_dmean2D, float* dL_dconic,
                      float* dL_dopacity, float* dL_dcolor, float* dL_ddepth,
                      float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                      float* dL_drot, int debug)
{
    if (P == 0) return 0;
    try {
        CudaRasterizer::Rasterizer::backward(
            P, D, M, R, bg, W, H, means3D, shs, colors_precomp, scales, scale_modifier,
            rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
            radii, g_geom.p, g_bin.p, g_img.p, dL_dpix, dL_dpix_depth, dL_dmean2D, dL_dconic,
            dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale,
            dL_drot, debug != 0);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    if (cudaPeekAtLastError() != cudaSuccess) return fail("ref backward");
    return 0;
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    if (P == 0) return 0;
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    if (cudaPeekAtLastError() != cudaSuccess) return fail("ref markVisible");
    return 0;
}

// Copies one decoded field of the reference's internal state (rasterizer_impl.h:21-73,
// carved as in rasterizer_impl.cu:155-194) of the LAST forward call to host memory.
// Returns the number of bytes the field holds, or -1. dst may be null to query the size.
long long ref_rast_get(const char* name, void* dst, long long dst_bytes)
{
    if (!g_geom.p) { g_err = "no forward call yet"; return -1; }
    char* c = g_geom.p;
    auto geom = CudaRasterizer::GeometryState::fromChunk(c, g_P);
    c = g_img.p;
    auto img = CudaRasterizer::ImageState::fromChunk(c, (size_t)g_W * g_H);
    CudaRasterizer::BinningState bin{};
    if (g_bin.p) { c = g_bin.p; bin = CudaRasterizer::BinningState::fromChunk(c, g_R); }
    const size_t P = g_P, R = g_R, N = (size_t)g_W * g_H;
    const size_t tiles = (size_t)((g_W + BLOCK_X - 1) / BLOCK_X) * ((g_H + BLOCK_Y - 1) / BLOCK_Y);
    const void* src = nullptr; size_t n = 0;
    std::string s(name);
    if      (s == "depths")              { src = geom.depths;        n = P * 4; }
    else if (s == "clamped")             { src = geom.clamped;       n = P * 3; }
    else if (s == "means2D")             { src = geom.means2D;       n = P * 8; }
    else if (s == "cov3D")               { src = geom.cov3D;         n = P * 24; }
    else if (s == "conic_opacity")       { src = geom.conic_opacity; n = P * 16; }
    else if (s == "rgb")                 { src = geom.rgb;           n = P * 12; }
    else if (s == "tiles_touched")       { src = geom.tiles_touched; n = P * 4; }
    else if (s == "point_offsets")       { src = geom.point_offsets; n = P * 4; }
    else if (s == "point_list")          { src = bin.point_list;     n = R * 4; }
    else if (s == "point_list_unsorted") { src = bin.point_list_unsorted; n = R * 4; }
    else if (s == "keys")                { src = bin.point_list_keys; n = R * 8; }
    else if (s == "keys_unsorted")       { src = bin.point_list_keys_unsorted; n = R * 8; }
    else if (s == "accum_alpha")         { src = img.accum_alpha;    n = N * 4; }
    else if (s == "n_contrib")           { src = img.n_contrib;      n = N * 4; }
    else if (s == "ranges")              { src = img.ranges;         n = tiles * 8; }
    else { g_err = "unknown field " + s; return -1; }
    if (dst) {
        if ((size_t)dst_bytes < n) { g_err = "dst too small"; return -1; }
        if (n && cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("ref_rast_get");
    }
    return (long long)n;
}

}  // extern "C"
