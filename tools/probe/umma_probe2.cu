// Scratch probe 2: can ONE shared-memory image of dY serve both tcgen05 MMAs of the deformation-MLP backward?
// Today dY is stored twice: K-major without swizzle (A operand of the dX chain, D[p][in] += dY[p][out] W[out][in]) and MN-major
// SWIZZLE_128B_BASE32B (A operand of the weight gradient, D[out][in] += dY[p][out] X[p][in]).  Physically the second image is
// [half = out / 32][p][128 bytes = 32 out], 32-byte granule index XORed with p & 3.  Read as a K-major operand (rows = points,
// 128-byte rows of K = out) that is a 128-byte-swizzled K-major tile with a 32-byte swizzle base -- if the descriptor layout
// type 1 is accepted for K-major operands, the first image (66 KB, 40 % of the kernel's shared-memory stores) can go.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -Iiclr2025_3d-mom_b200/csrc -o tools/probe/umma_probe2 tools/probe/umma_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <sys/wait.h>
#include <unistd.h>
#include "tc5_common.cuh"
using namespace b200gs; using namespace b200gs::tc5;

struct Args { u64 a_flags; u32 a_lbo, a_sbo, a_off[8], idesc; const float* A; const float* B; float* D; };

__global__ void __launch_bounds__(128, 1) probe(Args g)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    float* As = (float*)sm; float* Bs = (float*)(sm + 32768);
    u64* bar = (u64*)(sm + 49152); u32* slot = (u32*)(bar + 1);
    for (u32 i = threadIdx.x; i < 8192; i += 128) As[i] = g.A[i];          // 32 KB: [2 halves][128 p][32 out]
    for (u32 i = threadIdx.x; i < 4096; i += 128) Bs[i] = g.B[i];          // 16 KB: K-major [16 k-chunks][64 n][4]
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const u32 tb = *slot;
    if (threadIdx.x == 0) {
        for (u32 j = 0; j < 8; ++j)
            mma_ss(tb, smem_desc(smem_u32(As) + g.a_off[j], g.a_lbo, g.a_sbo) | g.a_flags,
                   smem_desc(smem_u32(Bs) + j * 2 * (64 * 16), 64 * 16, 128), g.idesc, j > 0);
        tc_commit(bar);
    }
    mbar_wait(bar, 0); tc_fence_after();
    const u32 la = tb + ((u32)((threadIdx.x >> 5) * 32) << 16);
    for (u32 c0 = 0; c0 < 64; c0 += 16) {
        u32 v[16]; tmem_ld16(la + c0, v); tmem_wait_ld();
        for (int e = 0; e < 16; ++e) g.D[threadIdx.x * 64 + c0 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(64u) : "memory");
}

static float tf(float x) { u32 u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

struct Config { const char* name; u64 flags; u32 lbo, sbo; int image; };    // image 0: granule ^ (p & 3) (the kernel's DYM); 1: 16-byte chunk ^ (p & 7) (standard SWIZZLE_128B)

static int run(const Config& c)
{
    const int M = 128, N = 64, K = 64;      // D[p][in] = sum_out dY[p][out] W[out][in]  (B[n][k] = W[k][n])
    std::vector<float> A(M * K), B(N * K), Dref(M * N), Ai(8192, 0.f), Bi(4096, 0.f), D(M * N);
    srand(1);
    for (auto& v : A) v = tf((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = tf((rand() % 2001 - 1000) / 1000.f);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; Dref[m * N + n] = (float)s; }
    for (int p = 0; p < M; ++p) for (int k = 0; k < K; ++k) {
        const int half = k >> 5, kk = k & 31;
        int byte;
        if (c.image == 0) byte = half * 16384 + p * 128 + ((((kk >> 3) ^ (p & 3)) << 5) | ((kk & 7) << 2));
        else byte = half * 16384 + p * 128 + ((((kk >> 2) ^ (p & 7)) << 4) | ((kk & 3) << 2));
        Ai[byte / 4] = A[p * K + k];
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bi[(k / 4) * (N * 4) + n * 4 + (k & 3)] = B[n * K + k];
    float *dA, *dB, *dD;
    if (cudaMalloc(&dA, 32768) != cudaSuccess) { printf("%-58s no device\n", c.name); return 3; }
    cudaMalloc(&dB, 16384); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, Ai.data(), 32768, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, M * N * 4);
    Args g{};
    g.a_flags = c.flags; g.a_lbo = c.lbo; g.a_sbo = c.sbo; g.A = dA; g.B = dB; g.D = dD;
    g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((u32)(N >> 3) << 17) | ((u32)(128 >> 4) << 24);        // K-major A and B
    for (int j = 0; j < 8; ++j) g.a_off[j] = (j >> 2) * 16384 + (j & 3) * 32;      // 8 out-features = one 32-byte step inside the 128-byte row
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 64);
    probe<<<1, 128, 49152 + 64>>>(g);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-58s CUDA error %s\n", c.name, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0;
    for (int i = 0; i < M * N; ++i) { err = fmax(err, fabs(D[i] - Dref[i])); mx = fmax(mx, fabs(Dref[i])); }
    printf("%-58s max|err| %.3e (ref max %.2f) %s\n", c.name, err, mx, err < 1e-3 ? "<== EXACT" : "");
    return err < 1e-3 ? 0 : 1;
}

int main()
{
    const u64 T1 = 1ull << 61, T2 = 2ull << 61;
    const Config configs[] = {
        {"K-major, type 1 (SW128_BASE32B), SBO 1024, LBO 16, image ^p&3", T1, 16, 1024, 0},
        {"K-major, type 1 (SW128_BASE32B), SBO 1024, LBO 16384, image ^p&3", T1, 16384, 1024, 0},
        {"K-major, type 1 (SW128_BASE32B), SBO 512, LBO 16, image ^p&3", T1, 16, 512, 0},
        {"K-major, type 1 (SW128_BASE32B), SBO 512, LBO 16384, image ^p&3", T1, 16384, 512, 0},
        {"K-major, type 2 (SWIZZLE_128B), SBO 1024, LBO 16, image ^p&7 (control)", T2, 16, 1024, 1},
        {"K-major, type 2 (SWIZZLE_128B), SBO 1024, LBO 16, image ^p&3", T2, 16, 1024, 0},
        {"K-major, type 1 (SW128_BASE32B), SBO 1024, LBO 16, image ^p&7", T1, 16, 1024, 1},
    };
    int exact = 0;
    for (const Config& c : configs) {           // one process per configuration: an illegal descriptor kills the CUDA context
        fflush(stdout);
        const pid_t pid = fork();
        if (pid == 0) { const int rc = run(c); fflush(stdout); _exit(rc); }
        int status = 0;
        waitpid(pid, &status, 0);
        if (WIFEXITED(status) && WEXITSTATUS(status) == 0) ++exact;
        else if (!WIFEXITED(status)) printf("%-58s child died (signal %d)\n", c.name, WTERMSIG(status));
    }
    printf("%d configuration(s) exact\n", exact);
    return 0;
}
