// Standalone rasterizer run through the C ABI (no Python, no torch): a seeded synthetic scene with the distributions of
// b200gs/synthetic.py (uniform cube, log-normal scales, random rotations / opacities / SH, camera 4.5 units in
// front of the cube), forward stage 1 + 2 and backward, device time per entry point (CUDA events on the launching stream) and
// checksums of every output.  Bit hashes for the deterministic outputs (radii, sorted instance list, tile ranges, colour,
// depth) and double-precision sums for the gradients (float atomics: order dependent) let two builds / two kernel variants
// be compared by diffing the printed lines:
//     tools/native/rast_check 1000000 1280 720 0.01 10
//     tools/native/rast_check 1000000 1280 720 0.01 10 some_option=1
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/native/rast_check tools/native/rast_check.cu \
//             -Iinclude -Liclr2025_3d-mom_b200/b200gs/lib -lb200gs -Xlinker -rpath -Xlinker '$ORIGIN/../../iclr2025_3d-mom_b200/b200gs/lib'
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "b200gs.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define BK(x) do { if ((x) != 0) { printf("b200gs error: %s (%s:%d)\n", b200gs_last_error(), __FILE__, __LINE__); exit(3); } } while (0)

static unsigned long long rng_state = 6666ull * 0x9E3779B97F4A7C15ull + 1;
static double urand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (double)(rng_state >> 11) * (1.0 / 9007199254740992.0); }
static double nrand() { double u = urand(), v = urand(); if (u < 1e-300) u = 1e-300; return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }
template <typename T> static T* to_dev(const std::vector<T>& h) { T* d; CK(cudaMalloc(&d, h.size() * sizeof(T))); CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }
template <typename T> static T* dev_alloc(size_t n) { T* d; CK(cudaMalloc(&d, (n ? n : 1) * sizeof(T))); CK(cudaMemset(d, 0, (n ? n : 1) * sizeof(T))); return d; }
static uint64_t hash_dev(const void* d, size_t bytes)
{
    std::vector<unsigned char> h(bytes);
    CK(cudaMemcpy(h.data(), d, bytes, cudaMemcpyDeviceToHost));
    uint64_t x = 1469598103934665603ull;
    const uint64_t* w = (const uint64_t*)h.data();
    for (size_t i = 0; i < bytes / 8; ++i) { x ^= w[i]; x *= 1099511628211ull; }
    for (size_t i = bytes / 8 * 8; i < bytes; ++i) { x ^= h[i]; x *= 1099511628211ull; }
    return x;
}
static void sums_dev(const char* name, const float* d, size_t n)
{
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost));
    double s = 0, a = 0; size_t bad = 0;
    for (float v : h) { if (!(v == v)) ++bad; else { s += v; a += fabs(v); } }
    printf("  %-12s sum % .9e  sum|.| %.9e  nan %zu\n", name, s, a, bad);
}
static void mat_mul4(const float* A, const float* B, float* C) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += (double)A[4 * i + k] * B[4 * k + j]; C[4 * i + j] = (float)s; } }

int main(int argc, char** argv)
{
    const int P = argc > 1 ? atoi(argv[1]) : 1000000, W = argc > 2 ? atoi(argv[2]) : 1280, H = argc > 3 ? atoi(argv[3]) : 720;
    const double mu = argc > 4 ? atof(argv[4]) : 0.01;
    const int reps = argc > 5 ? atoi(argv[5]) : 10;
    for (int i = 6; i < argc; ++i) {                       // name=value kernel options
        char* eq = strchr(argv[i], '=');
        if (!eq) continue;
        *eq = 0;
        BK(b200gs_set_option(argv[i], atoi(eq + 1)));
        printf("option %s = %d\n", argv[i], b200gs_get_option(argv[i]));
    }
    const int D = 3, M = 16;
    printf("b200gs %d: P = %d, %d x %d, scale_mu = %g, %d timed repetitions\n", b200gs_version(), P, W, H, mu, reps);
    std::vector<float> xyz((size_t)P * 3), scales((size_t)P * 3), rot((size_t)P * 4), opac(P), shs((size_t)P * M * 3);
    for (auto& v : xyz) v = (float)(urand() * 3.0 - 1.5);
    for (auto& v : scales) v = (float)exp(nrand() * 0.6 + log(mu));
    for (int i = 0; i < P; ++i) {
        double q[4], n = 0;
        for (int k = 0; k < 4; ++k) { q[k] = nrand(); n += q[k] * q[k]; }
        n = sqrt(n) > 1e-12 ? sqrt(n) : 1e-12;
        for (int k = 0; k < 4; ++k) rot[(size_t)4 * i + k] = (float)(q[k] / n);
        opac[i] = (float)(1.0 / (1.0 + exp(-nrand() * 1.5)));
        for (int c = 0; c < 3; ++c) shs[(size_t)i * M * 3 + c] = (float)(urand() * 3.0 - 1.5);
        for (int k = 3; k < M * 3; ++k) shs[(size_t)i * M * 3 + k] = (float)(nrand() * 0.05);
    }
    // camera of synthetic.make_camera: R = I, t = (0, 0, 4.5), focal 582.69 (H / 512) px, znear 0.01, zfar 100
    const double focal = 582.69 * (H / 512.0), tanx = W / (2 * focal), tany = H / (2 * focal), zn = 0.01, zf = 100.0;
    float view[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 4.5f, 1};            // world-to-view, transposed (row vectors)
    float projT[16] = {0};                                                           // projection, transposed
    projT[0] = (float)(1.0 / tanx); projT[5] = (float)(1.0 / tany); projT[10] = (float)(zf / (zf - zn)); projT[14] = (float)(-(zf * zn) / (zf - zn)); projT[11] = 1.f;
    float full[16]; mat_mul4(view, projT, full);
    float campos[3] = {0.f, 0.f, -4.5f}, bg[3] = {0.f, 0.f, 0.f};
    std::vector<float> dpix((size_t)3 * H * W);
    for (auto& v : dpix) v = (float)(nrand() / (3.0 * H * W));

    float *d_xyz = to_dev(xyz), *d_scales = to_dev(scales), *d_rot = to_dev(rot), *d_opac = to_dev(opac), *d_shs = to_dev(shs);
    float *d_view = to_dev(std::vector<float>(view, view + 16)), *d_full = to_dev(std::vector<float>(full, full + 16));
    float *d_campos = to_dev(std::vector<float>(campos, campos + 3)), *d_bg = to_dev(std::vector<float>(bg, bg + 3)), *d_dpix = to_dev(dpix);
    int* d_radii = dev_alloc<int>(P);
    float *d_color = dev_alloc<float>((size_t)3 * H * W), *d_depth = dev_alloc<float>((size_t)H * W);
    float *g_arena = dev_alloc<float>((size_t)12 * P), *g_m2d = dev_alloc<float>((size_t)3 * P), *g_col = dev_alloc<float>((size_t)3 * P),
          *g_op = dev_alloc<float>(P), *g_m3d = dev_alloc<float>((size_t)3 * P), *g_cov = dev_alloc<float>((size_t)6 * P),
          *g_sh = dev_alloc<float>((size_t)P * M * 3), *g_sc = dev_alloc<float>((size_t)3 * P), *g_rot = dev_alloc<float>((size_t)4 * P);
    size_t sz[3];
    BK(b200gs_rast_buffer_sizes(P, 0, W, H, sz));
    void *geom = nullptr, *img = nullptr, *bin = nullptr; size_t bin_cap = 0;
    CK(cudaMalloc(&geom, sz[0])); CK(cudaMalloc(&img, sz[2]));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t ev[4]; for (auto& e : ev) CK(cudaEventCreate(&e));
    double t1 = 0, t2 = 0, t3 = 0;
    unsigned long long cnt[2] = {0, 0};
    for (int rep = -2; rep < reps; ++rep) {
        CK(cudaEventRecord(ev[0], st));
        BK(b200gs_rast_forward_stage1(P, D, M, W, H, d_xyz, d_shs, nullptr, d_opac, d_scales, 1.f, d_rot, nullptr, d_view, d_full, d_campos,
                                      (float)tanx, (float)tany, 0, d_radii, geom, sz[0], cnt, st));
        BK(b200gs_rast_buffer_sizes(P, (long long)cnt[0], W, H, sz));
        if (sz[1] > bin_cap) { if (bin) CK(cudaFree(bin)); bin_cap = sz[1] + sz[1] / 8; CK(cudaMalloc(&bin, bin_cap)); }
        CK(cudaEventRecord(ev[1], st));
        BK(b200gs_rast_forward_stage2(P, (long long)cnt[0], (long long)cnt[1], W, H, d_bg, geom, bin, bin_cap, img, sz[2], d_color, d_depth, st));
        CK(cudaEventRecord(ev[2], st));
        BK(b200gs_rast_backward(P, D, M, (long long)cnt[0], W, H, d_bg, d_xyz, d_shs, nullptr, d_scales, 1.f, d_rot, nullptr, d_view, d_full, d_campos,
                                (float)tanx, (float)tany, d_radii, geom, bin, img, d_dpix, nullptr, g_arena, g_m2d, g_col, g_op, g_m3d, g_cov, g_sh, g_sc, g_rot, st));
        CK(cudaEventRecord(ev[3], st));
        CK(cudaStreamSynchronize(st));
        if (rep >= 0) {
            float a, b, c;
            CK(cudaEventElapsedTime(&a, ev[0], ev[1])); CK(cudaEventElapsedTime(&b, ev[1], ev[2])); CK(cudaEventElapsedTime(&c, ev[2], ev[3]));
            t1 += a; t2 += b; t3 += c;
        }
    }
    printf("instances R = %llu, visible Gaussians = %llu\n", cnt[0], cnt[1]);
    printf("device time per call: forward stage 1 (incl. the host read-back) %.3f ms, stage 2 %.3f ms, backward %.3f ms, total %.3f ms\n",
           t1 / reps, t2 / reps, t3 / reps, (t1 + t2 + t3) / reps);
    printf("deterministic outputs (bit hashes):\n");
    printf("  radii        %016llx\n  colour       %016llx\n  depth        %016llx\n", (unsigned long long)hash_dev(d_radii, (size_t)P * 4),
           (unsigned long long)hash_dev(d_color, (size_t)3 * H * W * 4), (unsigned long long)hash_dev(d_depth, (size_t)H * W * 4));
    {
        const size_t R = (size_t)cnt[0], tiles = (size_t)((W + 15) / 16) * ((H + 15) / 16);
        void* tmp = dev_alloc<unsigned char>(R * 8 > (size_t)H * W * 4 ? R * 8 : (size_t)H * W * 4);
        const char* fields[] = {"point_list", "keys", "ranges", "n_contrib"};
        const size_t bytes[] = {R * 4, R * 8, tiles * 8, (size_t)H * W * 4};
        for (int f = 0; f < 4; ++f) {
            const long long n = b200gs_rast_export(fields[f], P, (long long)R, W, H, geom, bin, img, tmp, (long long)bytes[f], st);
            CK(cudaStreamSynchronize(st));
            if (n < 0) printf("  %-12s export failed: %s\n", fields[f], b200gs_last_error());
            else printf("  %-12s %016llx\n", fields[f], (unsigned long long)hash_dev(tmp, (size_t)n));
        }
    }
    printf("gradients (float atomics, order dependent: compare to ~1e-6 relative):\n");
    sums_dev("dL_dmean2D", g_m2d, (size_t)3 * P); sums_dev("dL_dopacity", g_op, P); sums_dev("dL_dmean3D", g_m3d, (size_t)3 * P);
    sums_dev("dL_dcov3D", g_cov, (size_t)6 * P); sums_dev("dL_dsh", g_sh, (size_t)P * M * 3); sums_dev("dL_dscale", g_sc, (size_t)3 * P);
    sums_dev("dL_drot", g_rot, (size_t)4 * P);
    return 0;
}
