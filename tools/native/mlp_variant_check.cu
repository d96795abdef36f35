// Standalone check of the opt-in deformation-MLP kernel variants against the default kernels through the C ABI (no Python,
// no torch: starts in about a second, so it fits a very short GPU slot).
//   forward:  default vs "mlp_fwd_elect"  -> outputs and the activation stash must be BIT-IDENTICAL (same MMAs, same order)
//   backward: default vs "mlp_bwd_v2"     -> d_features bit-identical; weight / bias gradients are float atomics in both
//                                             kernels (order-dependent), so they are compared to 1e-5 of the tensor's scale
// and device time per launch (CUDA events on the launching stream, 5 launches each after one warm-up).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/native/mlp_variant_check tools/native/mlp_variant_check.cu \
//             -Iinclude -Liclr2025_3d-mom_b200/b200gs/lib -lb200gs -Xlinker -rpath -Xlinker '$ORIGIN/../../iclr2025_3d-mom_b200/b200gs/lib'
// run:   tools/native/mlp_variant_check [P]       (exit code 0 = all comparisons passed)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "b200gs.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define BK(x) do { if ((x) != 0) { printf("b200gs error: %s (%s:%d)\n", b200gs_last_error(), __FILE__, __LINE__); return 3; } } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static float frand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (float)((rng_state >> 40) * (1.0 / 16777216.0)); }
static float* dev_random(size_t n, float lo, float hi)
{
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = lo + (hi - lo) * frand();
    float* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(float)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    return d;
}
static float* dev_zero(size_t n) { float* d = nullptr; if (cudaMalloc(&d, n * sizeof(float)) != cudaSuccess) return nullptr; cudaMemset(d, 0, n * sizeof(float)); return d; }
static std::vector<float> to_host(const float* d, size_t n) { std::vector<float> h(n); cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost); return h; }
static bool same_bits(const char* what, const float* a, const float* b, size_t n)
{
    std::vector<float> x = to_host(a, n), y = to_host(b, n);
    size_t bad = 0; double worst = 0;
    for (size_t i = 0; i < n; ++i) if (memcmp(&x[i], &y[i], 4)) { ++bad; worst = fmax(worst, fabs((double)x[i] - y[i])); }
    printf("  %-28s %s (%zu of %zu differ, max |diff| %.3e)\n", what, bad ? "DIFFERENT" : "bit-identical", bad, n, worst);
    return bad == 0;
}
static bool close_rel(const char* what, const float* a, const float* b, size_t n, double tol)
{
    std::vector<float> x = to_host(a, n), y = to_host(b, n);
    double scale = 0, worst = 0;
    for (size_t i = 0; i < n; ++i) { scale = fmax(scale, fabs((double)x[i])); worst = fmax(worst, fabs((double)x[i] - y[i])); }
    const bool nan = !(worst == worst) || !(scale == scale);
    const bool ok = !nan && worst <= tol * fmax(scale, 1e-30);
    printf("  %-28s %s (max |diff| %.3e, scale %.3e, rel %.2e)\n", what, ok ? "ok" : "MISMATCH", worst, scale, worst / fmax(scale, 1e-30));
    return ok;
}

int main(int argc, char** argv)
{
    const long long P = argc > 1 ? atoll(argv[1]) : 1000000;
    const int W = 64, F = 64, kd[3] = {3, 3, 4};
    printf("b200gs %d, P = %lld\n", b200gs_version(), P);
    b200gs_mlp_weights w; memset(&w, 0, sizeof(w));
    w.feat_dim = F; w.width = W; w.feat_tiled = 0;
    w.w1 = dev_random((size_t)W * F, -0.2f, 0.2f); w.b1 = dev_random(W, -0.1f, 0.1f);
    for (int h = 0; h < 3; ++h) {
        w.w2[h] = dev_random((size_t)W * W, -0.2f, 0.2f); w.b2[h] = dev_random(W, -0.1f, 0.1f);
        w.w3[h] = dev_random((size_t)kd[h] * W, -0.2f, 0.2f); w.b3[h] = dev_random(kd[h], -0.1f, 0.1f);
    }
    float* feat = dev_random((size_t)P * F, 0.f, 1.f);
    float* xyz = dev_random((size_t)P * 3, -1.5f, 1.5f), *scales = dev_random((size_t)P * 3, -6.f, -4.f), *rot = dev_random((size_t)P * 4, -1.f, 1.f);
    float* flow = dev_random((size_t)P * 3, -1e-3f, 1e-3f);
    float* dp = dev_random((size_t)P * 3, -1.f, 1.f), *ds = dev_random((size_t)P * 3, -1.f, 1.f), *dr = dev_random((size_t)P * 4, -1.f, 1.f);
    const size_t nsaved = b200gs_deform_mlp_saved_floats(P);
    float* out[2][3], *saved[2], *dfeat[2];
    for (int v = 0; v < 2; ++v) {
        out[v][0] = dev_zero((size_t)P * 3); out[v][1] = dev_zero((size_t)P * 3); out[v][2] = dev_zero((size_t)P * 4);
        saved[v] = dev_zero(nsaved); dfeat[v] = dev_zero((size_t)P * F);
        if (!out[v][2] || !saved[v] || !dfeat[v]) { printf("out of device memory\n"); return 2; }
    }
    // gradient tables: one flat buffer per variant (16-byte aligned rows, as the trainer's arena provides)
    const size_t goff_w1 = 0, goff_b1 = goff_w1 + W * F, goff_w2 = goff_b1 + W, goff_b2 = goff_w2 + 3 * W * W, goff_w3 = goff_b2 + 3 * W,
                 goff_b3 = goff_w3 + 3 * 4 * W, gtotal = goff_b3 + 16;
    float* gbuf[2]; b200gs_mlp_grads g[2];
    for (int v = 0; v < 2; ++v) {
        gbuf[v] = dev_zero(gtotal);
        g[v].w1 = gbuf[v] + goff_w1; g[v].b1 = gbuf[v] + goff_b1;
        for (int h = 0; h < 3; ++h) { g[v].w2[h] = gbuf[v] + goff_w2 + h * W * W; g[v].b2[h] = gbuf[v] + goff_b2 + h * W; g[v].w3[h] = gbuf[v] + goff_w3 + h * 4 * W; g[v].b3[h] = gbuf[v] + goff_b3 + 4 * h; }
    }
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int REPS = 5;
    float ms_f[2], ms_b[2];
    bool ok = true;
    for (int v = 0; v < 2; ++v) {
        BK(b200gs_set_option("mlp_fwd_elect", v)); BK(b200gs_set_option("mlp_bwd_v2", v));
        for (int rep = -1; rep < REPS; ++rep) {            // rep -1 = warm-up
            if (rep == 0) CK(cudaEventRecord(e0, st));
            BK(b200gs_deform_mlp_forward(&w, P, feat, xyz, scales, rot, flow, 22.f, nullptr, 1.f, out[v][0], out[v][1], out[v][2], saved[v], st));
        }
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms_f[v], e0, e1)); ms_f[v] /= REPS;
        for (int rep = -1; rep < REPS; ++rep) {
            if (rep == 0) CK(cudaEventRecord(e0, st));
            if (rep == REPS - 1) CK(cudaMemsetAsync(gbuf[v], 0, gtotal * sizeof(float), st));      // compare one launch's gradients
            BK(b200gs_deform_mlp_backward(&w, &g[v], P, feat, saved[v], dp, ds, dr, dfeat[v], st));
        }
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms_b[v], e0, e1)); ms_b[v] /= REPS;
        printf("%s: forward %.3f ms, backward %.3f ms per launch\n", v ? "variants (mlp_fwd_elect, mlp_bwd_v2)" : "default kernels", ms_f[v], ms_b[v]);
    }
    printf("forward, default vs mlp_fwd_elect:\n");
    ok &= same_bits("pts_out", out[0][0], out[1][0], (size_t)P * 3);
    ok &= same_bits("scales_out", out[0][1], out[1][1], (size_t)P * 3);
    ok &= same_bits("rot_out", out[0][2], out[1][2], (size_t)P * 4);
    ok &= same_bits("activation stash + images", saved[0], saved[1], nsaved);
    printf("backward, default vs mlp_bwd_v2:\n");
    ok &= same_bits("d_features", dfeat[0], dfeat[1], (size_t)P * F);
    ok &= close_rel("dW1", g[0].w1, g[1].w1, (size_t)W * F, 1e-5);
    ok &= close_rel("db1", g[0].b1, g[1].b1, W, 1e-5);
    for (int h = 0; h < 3; ++h) {
        char name[32];
        snprintf(name, sizeof(name), "dW2[%d]", h); ok &= close_rel(name, g[0].w2[h], g[1].w2[h], (size_t)W * W, 1e-5);
        snprintf(name, sizeof(name), "db2[%d]", h); ok &= close_rel(name, g[0].b2[h], g[1].b2[h], W, 1e-5);
        snprintf(name, sizeof(name), "dW3[%d]", h); ok &= close_rel(name, g[0].w3[h], g[1].w3[h], (size_t)kd[h] * W, 1e-5);
        snprintf(name, sizeof(name), "db3[%d]", h); ok &= close_rel(name, g[0].b3[h], g[1].b3[h], kd[h], 1e-5);
    }
    printf("RESULT: %s; forward %.3f -> %.3f ms, backward %.3f -> %.3f ms\n", ok ? "PASS" : "FAIL", ms_f[0], ms_f[1], ms_b[0], ms_b[1]);
    return ok ? 0 : 1;
}
