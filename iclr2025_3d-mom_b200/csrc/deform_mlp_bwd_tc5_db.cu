// EXPERIMENTAL, UNVALIDATED (option "mlp_bwd_v2" = 151 / 183): deformation-MLP backward with ONE shared-memory image of dY per MMA
// group, double buffered.  Same maths, thread mappings, TMEM map and MMAs as deform_mlp_bwd_tc5.cu (read that file first); what
// differs is the schedule:
//   * dY is stored once per plane (hi, lo) in the MN-major SWIZZLE_128B_BASE32B image; the dX chain reads the same image as a
//     K-major operand (rows = points, 128-byte rows of 32 out-features, LBO 16 KB, 8-row group stride BwdArgs::dy_sbo) -- this
//     depends on the outcome of tools/probe/umma_probe2.cu;
//   * the 66 KB the K-major image used to take hold a SECOND dY image: group n stores into image n & 1 while the MMAs of group
//     n - 1 still read the other one, so no phase waits for the previous phase's MMAs before storing its operands.  Group n
//     waits for group n - 2 at its start (long finished), streams its own weight image into slot n & 1 under the cover of its
//     gradient math, and commits to mbarrier n & 1;
//   * the X operand (relu(hidden), later tf32(feature)) stays single: its two writers per tile wait for the group that still
//     reads it (the first head phase for the previous tile's feature group -- after its math, so the wait is short --, the
//     feature phase for the last head group, which it needs for D_RH anyway).
// Shared memory: W slot 0 | W slot 1 (32 KB each) | dY image 0 | dY image 1 (64 KB each) | X 32 KB | 2 mbarriers = 224 KB.
#include "tc5_common.cuh"
#include "../../include/b200gs.h"

namespace b200gs {
namespace tc5 {

namespace {
constexpr int BT = 256;
constexpr u32 KCH = 2064;                   // bounce scratch: bytes between 4-column chunks (2048 + 16 pad)
constexpr u64 DESC_SW128_32B = 1ull << 61;

__device__ __forceinline__ float4 tf32x4(float4 v)
{
    return make_float4(__uint_as_float(to_tf32(v.x)), __uint_as_float(to_tf32(v.y)), __uint_as_float(to_tf32(v.z)), __uint_as_float(to_tf32(v.w)));
}
__device__ __forceinline__ void split4(const float* x, float4& hi, float4& lo)
{
    hi = make_float4(x[0], x[1], x[2], x[3]);
    lo = make_float4(tf32_lo(x[0]), tf32_lo(x[1]), tf32_lo(x[2]), tf32_lo(x[3]));
}
__device__ __forceinline__ void sts128(u32 addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(u32 addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
}  // namespace

struct BwdDbArgs {
    b200gs_mlp_weights w;
    b200gs_mlp_grads gw;
    long long P;
    const float* feat; const float* saved;
    const float* d_pts; const float* d_scales; const float* d_rot;
    float* d_feat;
    u32 dy_sbo;
};

__global__ void __launch_bounds__(BT, 1) deform_mlp_bwd_tc5_db_kernel(const __grid_constant__ BwdDbArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* WS[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + 2 * MW * MW};
    unsigned char* DY0 = smem_raw + 65536;
    unsigned char* XH = DY0 + 2 * 65536;
    u64* bars = reinterpret_cast<u64*>(XH + 32768);
    u32* tmem_slot = reinterpret_cast<u32*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = lane & 7, sub = lane >> 3, c = warp & 1, pg = warp >> 1;          // SIMT mapping: columns 32 c + 4 q + e, points 32 pg + 4 i + sub
    const int col0 = 32 * c + 4 * q;
    const int pT = (warp & 3) * 32 + lane, cT = warp >> 2;                         // TMEM mapping: lane = point, 32 consecutive columns
    const int kdim[3] = {3, 3, 4};

    const float* images = reinterpret_cast<const float*>(a.saved) + 4 * stash_plane_floats(a.P);
    auto copy_image_pair = [&](float* dst, int m) {                   // 32 KB = 2048 16-byte pieces, 8 per thread
        const float4* src = reinterpret_cast<const float4*>(images + (size_t)(2 * m) * MW * MW);
#pragma unroll
        for (int i = 0; i < 8; ++i) cp_async16(reinterpret_cast<float4*>(dst) + tid + BT * i, src + tid + BT * i);
        cp_async_commit();
    };
    for (int i = tid; i < (2 * 65536 + 32768) / 16; i += BT) reinterpret_cast<float4*>(DY0)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        if (smem_u32(DY0) & 1023u) __trap();
        mbar_init(bars, 1); mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tbase = *tmem_slot;
    const u32 lane_addr = tbase + ((u32)((warp & 3) * 32) << 16);
    constexpr u32 C_RH = 0, C_FE = 64, C_W2 = 128, C_W1 = 320;
    const u32 id_kk = make_idesc(128, 64);
    const u32 id_mn = make_idesc(128, 64) | IDESC_A_MN | IDESC_B_MN;
    const u32 sDY[2] = {smem_u32(DY0), smem_u32(DY0) + 65536u}, sXH = smem_u32(XH), sWS[2] = {smem_u32(WS[0]), smem_u32(WS[1])};
    const u32 mn_off = (u32)c * 16384u + (u32)((((q >> 1) ^ sub) << 5) | ((q & 1) << 4));
    const u32 k_off = (u32)(8 * c + q) * KCH;
    const int p0 = 32 * pg + sub;
    bool first_tile = true;

    // ---- MMA groups: group n uses dY image / weight slot / mbarrier n & 1 ----
    u32 ngroup = 0, par[2] = {0u, 0u};
    bool pend[2] = {false, false};
    long long prev_row = -1;         // row (TMEM mapping) whose d_feature is still in D_FE ...
    u32 fe_bar = 0;                  // ... and the mbarrier of the group that produces it
    auto wait_bar = [&](u32 b) {
        if (pend[b]) { mbar_wait(bars + b, par[b]); par[b] ^= 1u; pend[b] = false; tc_fence_after(); }
    };
    auto store_dfeat_if_ready = [&]() {
        if (prev_row >= 0 && !pend[fe_bar]) {
            u32 v[32];
            tmem_ld32(lane_addr + C_FE + 32 * cT, v);
            tmem_wait_ld();
            if (prev_row < a.P) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(a.d_feat + (a.w.feat_tiled ? stash_off(prev_row, 32 * cT) + 16 * j : (size_t)prev_row * MW + 32 * cT + 4 * j)) =
                        make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
            prev_row = -1;
        }
    };
    auto issue_group = [&](u32 b, u32 d_col, bool d_accumulate, u32 w_col) {       // after the block barrier: warp 0 is converged
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                const u64 dB = smem_desc(sWS[b], MW * 16, 128);
                const u64 dA = smem_desc(sDY[b], 16384, a.dy_sbo) | DESC_SW128_32B;           // K-major view of the MN-major image
                const u64 dM = smem_desc(sDY[b], 16384, 512) | DESC_SW128_32B, dX = smem_desc(sXH, 16384, 512) | DESC_SW128_32B;
#pragma unroll
                for (int j = 0; j < 8; ++j) {          // K = 64 out features, 8 per instruction; lo*hi + hi*lo + hi*hi
                    const u64 bh = dB + (u64)((j * 2 * (MW * 16)) >> 4), bl = bh + (u64)((MW * MW * 4) >> 4);
                    const u64 ah = dA + (u64)(((j >> 2) * 16384 + (j & 3) * 32) >> 4), al = ah + (u64)(32768 >> 4);
                    mma_ss(tbase + d_col, al, bh, id_kk, (d_accumulate || j > 0) ? 1u : 0u);
                    mma_ss(tbase + d_col, ah, bl, id_kk, 1u);
                    mma_ss(tbase + d_col, ah, bh, id_kk, 1u);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j)           // K = 128 points, 8 per instruction
                    mma_ss(tbase + w_col, dM + (u64)((j * 1024) >> 4), dX + (u64)((j * 1024) >> 4), id_mn, (first_tile && j == 0) ? 0u : 1u);
                tc_commit(bars + b);
            }
            __syncwarp();
        }
        pend[b] = true;
        ++ngroup;
    };

    float gW3[3][4][4], gB2[3][4], gB3[3][4], gB1[4];
#pragma unroll
    for (int h = 0; h < 3; ++h)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gB2[h][k] = 0.f; gB3[h][k] = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) gW3[h][k][e] = 0.f;
        }
#pragma unroll
    for (int e = 0; e < 4; ++e) gB1[e] = 0.f;

    const bool en0 = a.w.w2[0] != nullptr, en1 = a.w.w2[1] != nullptr, en2 = a.w.w2[2] != nullptr;
    auto next_phase = [&](int ph) { return (ph < 0 && en0) ? 0 : (ph < 1 && en1) ? 1 : (ph < 2 && en2) ? 2 : 3; };
    auto load_rows = [&](float4* x, const float* src, long long r0) {
        const float* base = src + (a.w.feat_tiled ? stash_off(r0, col0) : (size_t)r0 * MW + col0);
        const size_t step = a.w.feat_tiled ? 256 : 4 * MW;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            x[i] = r0 + 4 * i < a.P ? __ldg(reinterpret_cast<const float4*>(base + step * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto load_stash = [&](float4* x, int plane, long long r0) {
        const float* src = a.saved + (size_t)plane * stash_plane_floats(a.P) + stash_off(r0, col0);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            x[i] = r0 + 4 * i < a.P ? __ldg(reinterpret_cast<const float4*>(src + 256 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto load_phase_rows = [&](int ph, float4* x, long long r0) {
        if (ph >= 3) load_rows(x, a.feat, r0); else load_stash(x, 1 + ph, r0);
    };
    auto load_phase_dout = [&](int ph, float (*d)[4], long long r0) {
        if (ph >= 3) return;
        const float* dsrc = ph == 0 ? a.d_pts : (ph == 1 ? a.d_scales : a.d_rot);
        const int kd = ph == 2 ? 4 : 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long r = r0 + 4 * i;
#pragma unroll
            for (int k = 0; k < 4; ++k) d[i][k] = (r < a.P && dsrc && k < kd) ? __ldg(dsrc + (size_t)r * kd + k) : 0.f;
        }
    };
    auto prefetch_phase_dout = [&](int ph, long long tile) {          // next phase's d_out rows into L2
        if (ph >= 3) return;
        const float* dsrc = ph == 0 ? a.d_pts : (ph == 1 ? a.d_scales : a.d_rot);
        const int kd = ph == 2 ? 4 : 3;
        if (!dsrc || tid > 4 * kd) return;
        const char* p = reinterpret_cast<const char*>(dsrc + (size_t)tile * ROWS * kd) + 128 * tid;
        if (p >= reinterpret_cast<const char*>(dsrc + (size_t)a.P * kd)) return;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
    };

    const long long nblocks = (a.P + ROWS - 1) / ROWS;
    float4 hrow[8], xin[8];
    if ((long long)blockIdx.x < nblocks) {
        load_stash(hrow, 0, (long long)blockIdx.x * ROWS + p0);
        load_phase_rows(next_phase(-1), xin, (long long)blockIdx.x * ROWS + p0);
    }
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const long long row0 = blk * ROWS + p0;
        u32 hmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            hmask |= (hrow[i].x > 0.f ? 1u : 0u) << (4 * i) | (hrow[i].y > 0.f ? 1u : 0u) << (4 * i + 1) |
                     (hrow[i].z > 0.f ? 1u : 0u) << (4 * i + 2) | (hrow[i].w > 0.f ? 1u : 0u) << (4 * i + 3);
        bool h_staged = false, rh_started = false;
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            if (!a.w.w2[h]) continue;
            const u32 b = ngroup & 1u;
            wait_bar(b);                         // group n - 2: its dY image and weight slot are free again
            store_dfeat_if_ready();
            copy_image_pair(WS[b], h);           // this group's weights, covered by the math below
            const int kd = kdim[h];
            float4 w3[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w3[k] = k < kd ? __ldg(reinterpret_cast<const float4*>(a.w.w3[h] + k * MW + col0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float din[8][4];
            load_phase_dout(h, din, row0);
            float dz[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float zz[4] = {xin[i].x, xin[i].y, xin[i].z, xin[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float wk[4] = {e == 0 ? w3[0].x : e == 1 ? w3[0].y : e == 2 ? w3[0].z : w3[0].w,
                                         e == 0 ? w3[1].x : e == 1 ? w3[1].y : e == 2 ? w3[1].z : w3[1].w,
                                         e == 0 ? w3[2].x : e == 1 ? w3[2].y : e == 2 ? w3[2].z : w3[2].w,
                                         e == 0 ? w3[3].x : e == 1 ? w3[3].y : e == 2 ? w3[3].z : w3[3].w};
                    float s = din[i][0] * wk[0];
                    s = fmaf(din[i][1], wk[1], s);
                    s = fmaf(din[i][2], wk[2], s);
                    if (kd > 3) s = fmaf(din[i][3], wk[3], s);
                    dz[i][e] = zz[e] > 0.f ? s : 0.f;
                    gB2[h][e] += dz[i][e];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < kd) gW3[h][k][e] = fmaf(din[i][k], zz[e], gW3[h][k][e]);
                }
                if (q == 0 && c == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) gB3[h][k] += din[i][k];
                }
            }
            load_phase_rows(next_phase(h), xin, row0);
            prefetch_phase_dout(next_phase(h), blk);
            if (!h_staged) {
                wait_bar(b ^ 1u);                // the previous tile's feature group still reads X
                store_dfeat_if_ready();
#pragma unroll
                for (int i = 0; i < 8; ++i) sts128(sXH + mn_off + (u32)(p0 + 4 * i) * 128u, tf32x4(hrow[i]));
                h_staged = true;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u32 pp = (u32)(p0 + 4 * i);
                float4 hi, lo;
                split4(dz[i], hi, lo);
                sts128(sDY[b] + mn_off + pp * 128u, hi);
                sts128(sDY[b] + 32768u + mn_off + pp * 128u, lo);
            }
            cp_async_wait<0>();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncthreads();
            issue_group(b, C_RH, rh_started, C_W2 + 64 * h);
            rh_started = true;
        }
        // ---- dh = d relu(hidden) masked ; d feature = dh W1 ; dW1 += dh^T feature ----
        if (next_phase(-1) == 3) load_rows(xin, a.feat, row0);        // every head disabled: nothing was prefetched
        float4 frow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) frow[i] = xin[i];
        if (blk + gridDim.x < nblocks) {
            load_stash(hrow, 0, (blk + gridDim.x) * ROWS + p0);
            load_phase_rows(next_phase(-1), xin, (blk + gridDim.x) * ROWS + p0);
            prefetch_phase_dout(next_phase(-1), blk + gridDim.x);
        }
        const u32 b = ngroup & 1u;
        wait_bar(b); wait_bar(b ^ 1u);           // D_RH is complete, both dY images and X are idle
        store_dfeat_if_ready();                  // before this tile's feature MMAs overwrite D_FE
        copy_image_pair(WS[b], 3);
        {   // D_RH: TMEM (lane = point) -> scratch inside dY image b -> the SIMT mapping
            u32 v[32];
            if (rh_started) {
                tmem_ld32(lane_addr + C_RH + 32 * cT, v);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sts128(sDY[b] + (u32)(8 * cT + j) * KCH + (u32)pT * 16u,
                       make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
        }
        tc_fence_before();
        __syncthreads();
        float4 rh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rh[i] = lds128(sDY[b] + k_off + (u32)(p0 + 4 * i) * 16u);
        __syncthreads();                         // everybody has read the scratch before anybody writes dh over it
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u32 pp = (u32)(p0 + 4 * i);
            float x[4] = {rh[i].x, rh[i].y, rh[i].z, rh[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                x[e] = (hmask >> (4 * i + e)) & 1u ? x[e] : 0.f;
                gB1[e] += x[e];
            }
            float4 hi, lo;
            split4(x, hi, lo);
            sts128(sDY[b] + mn_off + pp * 128u, hi);
            sts128(sDY[b] + 32768u + mn_off + pp * 128u, lo);
            sts128(sXH + mn_off + pp * 128u, tf32x4(frow[i]));
        }
        cp_async_wait<0>();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        fe_bar = b;
        issue_group(b, C_FE, false, C_W1);
        prev_row = blk * ROWS + pT;
        first_tile = false;
    }
    wait_bar(0u); wait_bar(1u);
    store_dfeat_if_ready();

    // ---- flush (as deform_mlp_bwd_tc5.cu, V2) ----
    if (!first_tile) {
        constexpr u32 SP = 68 * 4;
        for (int mi = 0; mi < 4; ++mi) {
            const int m = (mi + (int)blockIdx.x) & 3;
            if (m < 3 && !a.w.w2[m]) continue;
            float* dst = m < 3 ? a.gw.w2[m] : a.gw.w1;
            u32 v[32];
            tmem_ld32(lane_addr + (m < 3 ? C_W2 + 64 * m : C_W1) + 32 * cT, v);
            tmem_wait_ld();
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sts128(sDY[0] + (u32)pT * SP + (u32)(32 * cT + 4 * j) * 4u,
                       make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = tid + BT * ((k + (int)(blockIdx.x >> 2)) & 3);
                const int r = idx >> 4, c4 = idx & 15;
                const float4 x = lds128(sDY[0] + (u32)r * SP + (u32)c4 * 16u), y = lds128(sDY[0] + (u32)(r + 64) * SP + (u32)c4 * 16u);
                red_add_v4(dst + r * MW + 4 * c4, x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float s = gB1[e];
            s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 16);
            if (sub == 0) atomicAdd(a.gw.b1 + col0 + e, s);
        }
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            if (!a.w.w2[h]) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float s = gB2[h][e];
                s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (sub == 0) atomicAdd(a.gw.b2[h] + col0 + e, s);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float t = gW3[h][k][e];
                    t += __shfl_xor_sync(0xffffffffu, t, 8); t += __shfl_xor_sync(0xffffffffu, t, 16);
                    if (sub == 0 && k < kdim[h]) atomicAdd(a.gw.w3[h] + k * MW + col0 + e, t);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float s = gB3[h][k];
                s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (lane == 0 && c == 0 && k < kdim[h]) atomicAdd(a.gw.b3[h] + k, s);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512u) : "memory");
    }
}

}  // namespace tc5

int deform_mlp_backward_tc5_db(const b200gs_mlp_weights* w, const b200gs_mlp_grads* gw, long long P, const float* feat,
                               const float* saved, const float* d_pts, const float* d_scales, const float* d_rot,
                               float* d_feat, unsigned dy_sbo, cudaStream_t stream)
{
    tc5::BwdDbArgs a;
    a.w = *w; a.gw = *gw; a.P = P; a.feat = feat; a.saved = saved; a.d_pts = d_pts; a.d_scales = d_scales; a.d_rot = d_rot;
    a.d_feat = d_feat; a.dy_sbo = dy_sbo;
    const long long nblocks = (P + tc5::ROWS - 1) / tc5::ROWS;
    const int grid = (int)(nblocks < NUM_SMS ? nblocks : NUM_SMS);
    const size_t smem = 65536 + 2 * 65536 + 32768 + 64;
    cudaFuncSetAttribute(tc5::deform_mlp_bwd_tc5_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc5::deform_mlp_bwd_tc5_db_kernel<<<grid, tc5::BT, smem, stream>>>(a);
    return check_launch("deform_mlp_backward(tcgen05, double-buffered dY)");
}

}  // namespace b200gs
