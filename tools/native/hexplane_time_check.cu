// Standalone check of the time-row HexPlane kernel variants through the C ABI ("hexplane_time_fwd" = 0 / 1: features bit-identical;
// "hexplane_time_bwd" = 0 / 1 / 2) on the
// reference's field shape (2 levels, 64 / 128 spatial resolution, 50 time steps, 32 channels), one timestamp per launch:
// d_factor_accum and d_pts must be bit-identical (same arithmetic per element), the time-plane gradients (vector REDs into
// replicated rows: order dependent) equal to 1e-5 of their scale; device time per launch (CUDA events, 10 launches after a warm-up).
// run: tools/native/hexplane_time_check [P = 1000000] [tiled = 1]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "b200gs.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define BK(x) do { if ((x) != 0) { printf("b200gs error: %s (%s:%d)\n", b200gs_last_error(), __FILE__, __LINE__); exit(3); } } while (0)

static unsigned long long rng_state = 0x2545F4914F6CDD1Dull;
static float frand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (float)((rng_state >> 40) * (1.0 / 16777216.0)); }
static float* dev_random(size_t n, float lo, float hi)
{
    std::vector<float> h(n);
    for (auto& v : h) v = lo + (hi - lo) * frand();
    float* d; CK(cudaMalloc(&d, n * sizeof(float))); CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}
static float* dev_zero(size_t n) { float* d; CK(cudaMalloc(&d, (n ? n : 1) * sizeof(float))); CK(cudaMemset(d, 0, n * sizeof(float))); return d; }
static std::vector<float> to_host(const float* d, size_t n) { std::vector<float> h(n); CK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost)); return h; }

int main(int argc, char** argv)
{
    const long long P = argc > 1 ? atoll(argv[1]) : 1000000;
    const int tiled = argc > 2 ? atoi(argv[2]) : 1;
    const int F = 64;
    const size_t rowsP = (size_t)((P + 127) / 128) * 128;
    b200gs_hexplane_desc d; memset(&d, 0, sizeof(d));
    d.levels = 2; d.channels = 32;
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    size_t plane_floats[2][6];
    for (int l = 0; l < 2; ++l) {
        const int r = l == 0 ? 64 : 128;
        d.res[l][0] = d.res[l][1] = d.res[l][2] = r; d.res[l][3] = 50;
        for (int k = 0; k < 6; ++k) {
            plane_floats[l][k] = (size_t)d.res[l][pb[k]] * d.res[l][pa[k]] * 32;
            d.plane[l][k] = dev_random(plane_floats[l][k], 0.1f, 1.0f);
            d.grad_plane[l][k] = dev_zero(plane_floats[l][k]);
        }
    }
    std::vector<float> aabb = {1.5f, 1.5f, 1.5f, -1.5f, -1.5f, -1.5f};
    float* d_aabb; CK(cudaMalloc(&d_aabb, 24)); CK(cudaMemcpy(d_aabb, aabb.data(), 24, cudaMemcpyHostToDevice));
    d.aabb = d_aabb;
    if (b200gs_hexplane_time_supported(&d) != 1) { printf("time-row kernels do not support this descriptor: %s\n", b200gs_last_error()); return 3; }
    float* pts = dev_random((size_t)P * 3, -1.6f, 1.6f);                 // a few percent outside the box: border clamp
    float* factor = dev_random((size_t)P * F, 0.05f, 0.6f);
    float* dfeat = dev_random(rowsP * F, -1.f, 1.f);
    float* dfac = dev_zero((size_t)P * F), *dpts = dev_zero((size_t)P * 3), *feat = dev_zero(rowsP * F);
    const size_t sb = b200gs_hexplane_time_row_scratch_bytes(&d, 64);
    void* scratch; CK(cudaMalloc(&scratch, sb));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const float t = 0.37f;
    printf("b200gs %d: P = %lld, features / d_features %s, scratch %zu bytes\n", b200gs_version(), P, tiled ? "tiled" : "row-major", sb);
    bool fwd_ok = true;
    {
        std::vector<float> r_feat;
        for (int variant = 0; variant < 3; ++variant) {
            BK(b200gs_set_option("hexplane_time_fwd", variant));
            float total = 0;
            for (int rep = -1; rep < 10; ++rep) {
                CK(cudaEventRecord(e0, st));
                BK(b200gs_hexplane_time_forward(&d, P, pts, nullptr, t, factor, feat, tiled, st));
                CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep >= 0) total += ms;
            }
            std::vector<float> a = to_host(feat, rowsP * F);
            size_t bad = 0;
            if (variant == 0) r_feat = a; else for (size_t i = 0; i < a.size(); ++i) bad += memcmp(&a[i], &r_feat[i], 4) != 0;
            printf("hexplane_time_fwd = %d: %.3f ms per launch%s\n", variant, total / 10, variant == 0 ? "" : (bad ? "  features DIFFER" : "  features bit-identical"));
            fwd_ok &= bad == 0;
        }
        BK(b200gs_set_option("hexplane_time_fwd", 0));
    }
    std::vector<float> r_dfac, r_dpts, r_gp[2][3];
    const int tk[3] = {2, 4, 5};
    bool ok = fwd_ok;
    for (int variant = 0; variant < 3; ++variant) {
        BK(b200gs_set_option("hexplane_time_bwd", variant));
        float total = 0;
        for (int rep = -1; rep < 10; ++rep) {
            if (rep == 9) {                                               // the compared launch starts from zeroed accumulators
                CK(cudaMemsetAsync(dfac, 0, (size_t)P * F * 4, st));
                for (int l = 0; l < 2; ++l) for (int k = 0; k < 3; ++k) CK(cudaMemsetAsync(d.grad_plane[l][tk[k]], 0, plane_floats[l][tk[k]] * 4, st));
            }
            CK(cudaEventRecord(e0, st));
            BK(b200gs_hexplane_time_backward(&d, P, pts, nullptr, t, factor, dfac, dfeat, dpts, scratch, sb, tiled, st));
            CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep >= 0) total += ms;
        }
        printf("hexplane_time_bwd = %d: %.3f ms per launch (backward + row flush)", variant, total / 10);
        std::vector<float> a = to_host(dfac, (size_t)P * F), b = to_host(dpts, (size_t)P * 3);
        if (variant == 0) { r_dfac = a; r_dpts = b; for (int l = 0; l < 2; ++l) for (int k = 0; k < 3; ++k) r_gp[l][k] = to_host(d.grad_plane[l][tk[k]], plane_floats[l][tk[k]]); printf("\n"); continue; }
        size_t bad = 0;
        for (size_t i = 0; i < a.size(); ++i) bad += memcmp(&a[i], &r_dfac[i], 4) != 0;
        for (size_t i = 0; i < b.size(); ++i) bad += memcmp(&b[i], &r_dpts[i], 4) != 0;
        double worst = 0;
        for (int l = 0; l < 2; ++l) for (int k = 0; k < 3; ++k) {
            std::vector<float> gp = to_host(d.grad_plane[l][tk[k]], plane_floats[l][tk[k]]);
            double scale = 0, diff = 0;
            for (size_t i = 0; i < gp.size(); ++i) { scale = fmax(scale, fabs((double)r_gp[l][k][i])); diff = fmax(diff, fabs((double)gp[i] - r_gp[l][k][i])); if (!(gp[i] == gp[i])) diff = 1e30; }
            worst = fmax(worst, diff / fmax(scale, 1e-30));
        }
        const bool vok = bad == 0 && worst <= 1e-5;
        printf("  d_factor / d_pts: %zu elements differ; time-plane gradients: max relative difference %.2e  %s\n", bad, worst, vok ? "ok" : "MISMATCH");
        ok &= vok;
    }
    printf("RESULT: %s\n", ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
