// Standalone check of the opt-in deformation-MLP kernel variants against the default kernels through the C ABI (no Python,
// no torch: starts in about a second, so it fits a very short GPU slot).
//   forward:  default vs "mlp_fwd_elect"   -> outputs and the activation stash must be BIT-IDENTICAL (same MMAs, same order)
//   backward: default vs "mlp_bwd_v2" = v   -> d_features bit-identical; weight / bias gradients are float atomics in every
//                                              kernel (order-dependent), so they are compared to 1e-5 of the tensor's scale
// over several point counts (tails, fewer tiles than SMs), both feature layouts and disabled heads; at the first (large) case
// every variant is also timed: CUDA events on the launching stream, 5 launches after one warm-up, L2 flushed by the working
// set itself (1.4 / 1.6 GB per launch at 1M points).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/native/mlp_variant_check tools/native/mlp_variant_check.cu \
//             -Iinclude -Liclr2025_3d-mom_b200/b200gs/lib -lb200gs -Xlinker -rpath -Xlinker '$ORIGIN/../../iclr2025_3d-mom_b200/b200gs/lib'
// run:   tools/native/mlp_variant_check [P of the timed case] [comma-separated mlp_bwd_v2 values]     (exit code 0 = all passed)
//        tools/native/mlp_variant_check 1000000 ablate      (times the backward kernel with one part removed at a time)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "b200gs.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define BK(x) do { if ((x) != 0) { printf("b200gs error: %s (%s:%d)\n", b200gs_last_error(), __FILE__, __LINE__); exit(3); } } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static float frand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (float)((rng_state >> 40) * (1.0 / 16777216.0)); }
static std::vector<void*> g_allocs;
static float* dev_random(size_t n, float lo, float hi)
{
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = lo + (hi - lo) * frand();
    float* d = nullptr;
    CK(cudaMalloc(&d, (n ? n : 1) * sizeof(float)));
    CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    g_allocs.push_back(d);
    return d;
}
static float* dev_zero(size_t n) { float* d = nullptr; CK(cudaMalloc(&d, (n ? n : 1) * sizeof(float))); CK(cudaMemset(d, 0, n * sizeof(float))); g_allocs.push_back(d); return d; }
static void free_all() { for (void* p : g_allocs) cudaFree(p); g_allocs.clear(); }
static std::vector<float> to_host(const float* d, size_t n) { std::vector<float> h(n); CK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost)); return h; }
static bool same_bits(const char* what, const std::vector<float>& x, const float* b, bool quiet)
{
    std::vector<float> y = to_host(b, x.size());
    size_t bad = 0; double worst = 0;
    for (size_t i = 0; i < x.size(); ++i) if (memcmp(&x[i], &y[i], 4)) { ++bad; worst = fmax(worst, fabs((double)x[i] - y[i])); }
    if (bad || !quiet) printf("    %-26s %s (%zu of %zu differ, max |diff| %.3e)\n", what, bad ? "DIFFERENT" : "bit-identical", bad, x.size(), worst);
    return bad == 0;
}
static bool close_rel(const char* what, const std::vector<float>& x, const float* b, double tol, bool quiet)
{
    std::vector<float> y = to_host(b, x.size());
    double scale = 0, worst = 0; bool nan = false;
    for (size_t i = 0; i < x.size(); ++i) { scale = fmax(scale, fabs((double)x[i])); worst = fmax(worst, fabs((double)x[i] - y[i])); nan |= !(y[i] == y[i]); }
    const bool ok = !nan && worst <= tol * fmax(scale, 1e-30);
    if (!ok || !quiet) printf("    %-26s %s (max |diff| %.3e, scale %.3e, rel %.2e)\n", what, ok ? "ok" : "MISMATCH", worst, scale, worst / fmax(scale, 1e-30));
    return ok;
}

static const int W = 64, F = 64, KD[3] = {3, 3, 4};
static const size_t G_W1 = 0, G_B1 = G_W1 + W * F, G_W2 = G_B1 + W, G_B2 = G_W2 + 3 * W * W, G_W3 = G_B2 + 3 * W, G_B3 = G_W3 + 3 * 4 * W, G_TOTAL = G_B3 + 16;
static b200gs_mlp_grads grads_in(float* buf)
{
    b200gs_mlp_grads g;
    g.w1 = buf + G_W1; g.b1 = buf + G_B1;
    for (int h = 0; h < 3; ++h) { g.w2[h] = buf + G_W2 + h * W * W; g.b2[h] = buf + G_B2 + h * W; g.w3[h] = buf + G_W3 + h * 4 * W; g.b3[h] = buf + G_B3 + 4 * h; }
    return g;
}

// one configuration: default kernels first, then every variant against them
static bool run_case(long long P, int tiled, int heads, const std::vector<int>& bwd_variants, bool timed, cudaStream_t st)
{
    printf("case P = %lld, features %s, heads %d%d%d%s\n", P, tiled ? "tiled" : "row-major", heads & 1, (heads >> 1) & 1, (heads >> 2) & 1, timed ? " (timed)" : "");
    b200gs_mlp_weights w; memset(&w, 0, sizeof(w));
    w.feat_dim = F; w.width = W; w.feat_tiled = tiled;
    w.w1 = dev_random((size_t)W * F, -0.2f, 0.2f); w.b1 = dev_random(W, -0.1f, 0.1f);
    for (int h = 0; h < 3; ++h) if ((heads >> h) & 1) {
        w.w2[h] = dev_random((size_t)W * W, -0.2f, 0.2f); w.b2[h] = dev_random(W, -0.1f, 0.1f);
        w.w3[h] = dev_random((size_t)KD[h] * W, -0.2f, 0.2f); w.b3[h] = dev_random(KD[h], -0.1f, 0.1f);
    }
    const size_t rowsP = (size_t)((P + 127) / 128) * 128;              // the tiled layout is padded to whole 128-point tiles
    float* feat = dev_random(rowsP * F, 0.f, 1.f);
    float* xyz = dev_random((size_t)P * 3, -1.5f, 1.5f), *scales = dev_random((size_t)P * 3, -6.f, -4.f), *rot = dev_random((size_t)P * 4, -1.f, 1.f);
    float* flow = dev_random((size_t)P * 3, -1e-3f, 1e-3f);
    const float* dout[3] = {nullptr, nullptr, nullptr};
    for (int h = 0; h < 3; ++h) if ((heads >> h) & 1) dout[h] = dev_random((size_t)P * KD[h], -1.f, 1.f);
    const size_t nsaved = b200gs_deform_mlp_saved_floats(P);
    float* out[3] = {dev_zero((size_t)P * 3), dev_zero((size_t)P * 3), dev_zero((size_t)P * 4)};
    float* saved = dev_zero(nsaved), *dfeat = dev_zero(rowsP * F), *gbuf = dev_zero(G_TOTAL);
    b200gs_mlp_grads g = grads_in(gbuf);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int REPS = timed ? 5 : 1;
    auto forward = [&](int elect) -> float {
        BK(b200gs_set_option("mlp_fwd_elect", elect));
        CK(cudaMemsetAsync(saved, 0, nsaved * sizeof(float), st));
        for (int rep = timed ? -1 : 0; rep < REPS; ++rep) {
            if (rep == 0) CK(cudaEventRecord(e0, st));
            BK(b200gs_deform_mlp_forward(&w, P, feat, xyz, scales, rot, flow, 22.f, nullptr, 1.f, out[0], out[1], out[2], saved, st));
        }
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / REPS;
    };
    auto backward = [&](int variant) -> float {
        BK(b200gs_set_option("mlp_bwd_v2", variant));
        CK(cudaMemsetAsync(dfeat, 0, rowsP * F * sizeof(float), st));
        for (int rep = timed ? -1 : 0; rep < REPS; ++rep) {
            if (rep == 0) CK(cudaEventRecord(e0, st));
            if (rep == REPS - 1) CK(cudaMemsetAsync(gbuf, 0, G_TOTAL * sizeof(float), st));      // compare one launch's gradients
            BK(b200gs_deform_mlp_backward(&w, &g, P, feat, saved, dout[0], dout[1], dout[2], dfeat, st));
        }
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / REPS;
    };
    bool ok = true;
    const float f0 = forward(0);
    std::vector<float> r_out[3] = {to_host(out[0], (size_t)P * 3), to_host(out[1], (size_t)P * 3), to_host(out[2], (size_t)P * 4)};
    std::vector<float> r_saved = to_host(saved, nsaved);
    if (timed) printf("  forward: default %.3f ms\n", f0);
    for (int fv = 1; fv <= 2; ++fv) {
        const float f1 = forward(fv);
        const bool fok = same_bits("pts_out", r_out[0], out[0], true) & same_bits("scales_out", r_out[1], out[1], true) &
                         same_bits("rot_out", r_out[2], out[2], true) & same_bits("activation stash + images", r_saved, saved, true);
        if (timed || !fok) printf("  forward mlp_fwd_elect = %d: %.3f ms  %s\n", fv, f1, fok ? "bit-identical to the default kernel" : "MISMATCH");
        ok &= fok;
    }
    r_saved.clear(); r_saved.shrink_to_fit();
    const float b0 = backward(0);
    std::vector<float> r_dfeat = to_host(dfeat, rowsP * F), r_g = to_host(gbuf, G_TOTAL);
    if (timed) printf("  backward: default %.3f ms\n", b0);
    for (int v : bwd_variants) {
        const float bv = backward(v);
        bool vok = same_bits("d_features", r_dfeat, dfeat, true) & close_rel("weight / bias gradients", r_g, gbuf, 1e-5, true);
        if (timed || !vok) printf("  backward mlp_bwd_v2 = %2d: %.3f ms  %s\n", v, bv, vok ? "matches the default kernel" : "MISMATCH");
        if (!vok) { same_bits("d_features", r_dfeat, dfeat, false); close_rel("weight / bias gradients", r_g, gbuf, 1e-5, false); }
        ok &= vok;
    }
    BK(b200gs_set_option("mlp_fwd_elect", 0)); BK(b200gs_set_option("mlp_bwd_v2", 0));
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    free_all();
    printf("  -> %s\n", ok ? "pass" : "FAIL");
    return ok;
}

// what the default backward kernel costs without one of its parts (option "mlp_bwd_ablate": wrong results, timing only)
static void run_ablations(long long P, cudaStream_t st)
{
    printf("backward ablations at P = %lld, tiled features (results are wrong by construction; only the times mean something):\n", P);
    b200gs_mlp_weights w; memset(&w, 0, sizeof(w));
    w.feat_dim = F; w.width = W; w.feat_tiled = 1;
    w.w1 = dev_random((size_t)W * F, -0.2f, 0.2f); w.b1 = dev_random(W, -0.1f, 0.1f);
    for (int h = 0; h < 3; ++h) {
        w.w2[h] = dev_random((size_t)W * W, -0.2f, 0.2f); w.b2[h] = dev_random(W, -0.1f, 0.1f);
        w.w3[h] = dev_random((size_t)KD[h] * W, -0.2f, 0.2f); w.b3[h] = dev_random(KD[h], -0.1f, 0.1f);
    }
    const size_t rowsP = (size_t)((P + 127) / 128) * 128;
    float* feat = dev_random(rowsP * F, 0.f, 1.f);
    float* xyz = dev_random((size_t)P * 3, -1.5f, 1.5f), *scales = dev_random((size_t)P * 3, -6.f, -4.f), *rot = dev_random((size_t)P * 4, -1.f, 1.f);
    float* flow = dev_random((size_t)P * 3, -1e-3f, 1e-3f);
    const float* dout[3] = {dev_random((size_t)P * 3, -1.f, 1.f), dev_random((size_t)P * 3, -1.f, 1.f), dev_random((size_t)P * 4, -1.f, 1.f)};
    float* out[3] = {dev_zero((size_t)P * 3), dev_zero((size_t)P * 3), dev_zero((size_t)P * 4)};
    float* saved = dev_zero(b200gs_deform_mlp_saved_floats(P)), *dfeat = dev_zero(rowsP * F), *gbuf = dev_zero(G_TOTAL);
    b200gs_mlp_grads g = grads_in(gbuf);
    BK(b200gs_deform_mlp_forward(&w, P, feat, xyz, scales, rot, flow, 22.f, nullptr, 1.f, out[0], out[1], out[2], saved, st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int abl[] = {0, 1, 2, 4, 8, 16, 32, 10};
    const char* what[] = {"complete kernel", "d_out from constants (no global loads of d_out)", "no operand stores to shared memory", "no MMAs, no waits for them",
                          "no gradient math (dz, dW3, bias sums)", "no d_feature stores", "stash / feature rows from constants (no global loads)",
                          "neither operand stores nor gradient math"};
    for (int k = 0; k < 8; ++k) {
        BK(b200gs_set_option("mlp_bwd_ablate", abl[k]));
        for (int rep = -1; rep < 5; ++rep) {
            if (rep == 0) CK(cudaEventRecord(e0, st));
            BK(b200gs_deform_mlp_backward(&w, &g, P, feat, saved, dout[0], dout[1], dout[2], dfeat, st));
        }
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  mlp_bwd_ablate = %2d: %.3f ms  %s\n", abl[k], ms / 5, what[k]);
    }
    BK(b200gs_set_option("mlp_bwd_ablate", 0));
    free_all();
}

int main(int argc, char** argv)
{
    const long long P = argc > 1 ? atoll(argv[1]) : 1000000;
    setenv("B200GS_PROFILING", "1", 1);              // the library refuses experimental variants and ablation builds otherwise
    if (argc > 2 && !strcmp(argv[2], "ablate")) {
        cudaStream_t st; CK(cudaStreamCreate(&st));
        run_ablations(P, st);
        return 0;
    }
    std::vector<int> variants;
    for (char* t = strtok(argc > 2 ? argv[2] : (char*)"", ","); t; t = strtok(nullptr, ",")) variants.push_back(atoi(t));
    if (variants.empty()) variants = {1, 3, 5, 7, 9, 13, 15};
    printf("b200gs %d\n", b200gs_version());
    cudaStream_t st; CK(cudaStreamCreate(&st));
    bool ok = run_case(P, 0, 7, variants, true, st);
    ok &= run_case(P, 1, 7, variants, true, st);
    const long long small[] = {1, 64, 129, 3001, 40003};
    for (long long p : small)
        for (int tiled = 0; tiled < 2; ++tiled) ok &= run_case(p, tiled, 7, variants, false, st);
    const int masks[] = {5, 6, 1, 0};                     // disabled heads: different numbers of MMA groups per tile
    for (int m : masks) ok &= run_case(20011, 1, m, variants, false, st);
    printf("RESULT: %s\n", ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
