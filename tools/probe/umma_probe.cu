// Scratch probe: which un-swizzled shared-memory layouts does tcgen05.mma.kind::tf32 accept for
// MN-major operands?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../../iclr2025_3d-mom_b200/csrc -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "tc5_common.cuh"
using namespace b200gs; using namespace b200gs::tc5;

struct Args { u64 a_flags, b_flags; u32 a_lbo, a_sbo, b_lbo, b_sbo, idesc, nk, a_step, b_step, N; const float* A; const float* B; float* D; u32 a_bytes, b_bytes; };

__global__ void __launch_bounds__(128, 1) probe(Args g)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    float* As = (float*)sm; float* Bs = (float*)(sm + 65536);
    u64* bar = (u64*)(sm + 131072); u32* slot = (u32*)(bar + 1);
    for (u32 i = threadIdx.x; i < g.a_bytes / 4; i += 128) As[i] = g.A[i];
    for (u32 i = threadIdx.x; i < g.b_bytes / 4; i += 128) Bs[i] = g.B[i];
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const u32 tb = *slot;
    if (threadIdx.x == 0) {
        for (u32 j = 0; j < g.nk; ++j)
            mma_ss(tb, smem_desc(smem_u32(As) + j * g.a_step, g.a_lbo, g.a_sbo) | g.a_flags, smem_desc(smem_u32(Bs) + j * g.b_step, g.b_lbo, g.b_sbo) | g.b_flags, g.idesc, j > 0);
        tc_commit(bar);
    }
    mbar_wait(bar, 0); tc_fence_after();
    const u32 la = tb + ((u32)((threadIdx.x >> 5) * 32) << 16);
    for (u32 c0 = 0; c0 < g.N; c0 += 16) {
        u32 v[16]; tmem_ld16(la + c0, v); tmem_wait_ld();
        for (int e = 0; e < 16; ++e) g.D[threadIdx.x * g.N + c0 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512u) : "memory");
}

static float tf(float x) { u32 u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main()
{
    const int M = 128, N = 64, K = 128;      // D[m][n] = sum_k A[m][k] B[n][k]
    std::vector<float> A(M * K), B(N * K), Dref(M * N);
    srand(1);
    for (auto& v : A) v = tf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = tf((rand() % 2001 - 1000) / 1000.f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; Dref[m * N + n] = (float)s; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 64);
    float *dA, *dB, *dD; cudaMalloc(&dA, 65536); cudaMalloc(&dB, 65536); cudaMalloc(&dD, M * 256 * 4);
    std::vector<float> D(M * N);
    auto run = [&](const char* name, std::vector<float>& Ai, std::vector<float>& Bi, Args g) {
        cudaMemcpy(dA, Ai.data(), Ai.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), Bi.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, M * N * 4);
        g.A = dA; g.B = dB; g.D = dD; g.a_bytes = Ai.size() * 4; g.b_bytes = Bi.size() * 4; g.N = N;
        probe<<<1, 128, 131072 + 64>>>(g);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-40s CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0; int nz = 0;
        for (int i = 0; i < M * N; ++i) { err = fmax(err, fabs(D[i] - Dref[i])); mx = fmax(mx, fabs(Dref[i])); nz += D[i] != 0.f; }
        printf("%-40s max|err| %.3e (ref max %.2f) nonzero %d  D[0..3] %.4f %.4f %.4f %.4f  ref %.4f %.4f %.4f %.4f\n", name, err, mx, nz, D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3]);
    };
    // images
    std::vector<float> Ak(M * K), Bk(N * K), Amn(M * K), Bmn(N * K);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        Ak[(k / 4) * (M * 4) + m * 4 + (k & 3)] = A[m * K + k];                  // K-major: LBO = M*16, SBO = 128
        Amn[(m / 4) * (K * 4) + k * 4 + (m & 3)] = A[m * K + k];                 // MN-major: chunk(m/4) stride K*16, 16 B per k
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        Bk[(k / 4) * (N * 4) + n * 4 + (k & 3)] = B[n * K + k];
        Bmn[(n / 4) * (K * 4) + k * 4 + (n & 3)] = B[n * K + k];
    }
    const u32 id = (1u << 4) | (2u << 7) | (2u << 10) | ((u32)(N >> 3) << 17) | ((u32)(128 >> 4) << 24);
    Args g{};
    g.nk = K / 8;
    // 1. both K-major (known-good form)
    g.a_lbo = M * 16; g.a_sbo = 128; g.a_step = 2 * M * 16; g.b_lbo = N * 16; g.b_sbo = 128; g.b_step = 2 * N * 16; g.idesc = id;
    run("A K-major, B K-major", Ak, Bk, g);
    // Does the tensor core truncate or round FP32 containers to TF32? Full-precision inputs, two CPU references.
    {
        std::vector<float> Af(M * K), Bf(N * K), Akf(M * K), Bkf(N * K);
        for (auto& v : Af) v = (rand() % 2000001 - 1000000) / 1000000.f * 1.2345678f;
        for (auto& v : Bf) v = (rand() % 2000001 - 1000000) / 1000000.f * 0.7654321f;
        auto rna = [](float x) { u32 u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Akf[(k / 4) * (M * 4) + m * 4 + (k & 3)] = Af[m * K + k];
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bkf[(k / 4) * (N * 4) + n * 4 + (k & 3)] = Bf[n * K + k];
        g.a_lbo = M * 16; g.a_sbo = 128; g.a_step = 2 * M * 16; g.b_lbo = N * 16; g.b_sbo = 128; g.b_step = 2 * N * 16; g.idesc = id; g.a_flags = 0; g.b_flags = 0;
        for (int mode = 0; mode < 2; ++mode) {
            for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
                double acc = 0;
                for (int k = 0; k < K; ++k) acc += (double)(mode ? rna(Af[m * K + k]) : tf(Af[m * K + k])) * (mode ? rna(Bf[n * K + k]) : tf(Bf[n * K + k]));
                Dref[m * N + n] = (float)acc;
            }
            run(mode ? "full-precision inputs vs ROUNDED reference" : "full-precision inputs vs TRUNCATED reference", Akf, Bkf, g);
        }
    }
    return 0;
}
