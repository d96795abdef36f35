// distCUDA2: mean squared distance to the 3 nearest neighbours of every point.
//
// Same result as the reference's SimpleKNN::knn (KNN/simple_knn.cu:185-220): the three
// smallest values of d.x*d.x + d.y*d.y + d.z*d.z over all OTHER indices, averaged as
// (b0 + b1 + b2) / 3.  The reference prunes with one flat list of 1024-point boxes that
// every point scans end to end (O(P * P/1024)); here the Morton-sorted points carry a
// three-level box hierarchy (32 / 1024 / 32768 points), so a point only descends into
// boxes that can still beat its current third-best.  Pruning is conservative in float
// arithmetic (box distance <= distance to any point inside, by monotonic rounding), so the
// result is bit-identical whatever the traversal order.
#include <cfloat>
#include "common.cuh"
#include "radix_sort.cuh"

namespace b200gs {

namespace {

struct Box { float3 lo, hi; };

__device__ __forceinline__ u32 spread3(u32 x)
{
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

// scene bounds, seeded with the origin like the reference's reduction (simple_knn.cu:191)
__global__ void __launch_bounds__(256) knn_bounds_kernel(int P, const float* __restrict__ pts, float* __restrict__ bounds)
{
    __shared__ float s[6][8];
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { float v = pts[3 * (size_t)i + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) for (int c = 0; c < 3; ++c) { s[c][warp] = lo[c]; s[3 + c][warp] = hi[c]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][w]) : fmaxf(v, s[threadIdx.x][w]);
        // float min/max via the integer trick is sign-dependent; bounds straddle 0 so use CAS-free ordered ints
        int* dst = reinterpret_cast<int*>(bounds) + threadIdx.x;
        if (threadIdx.x < 3) { if (v < 0.f) atomicMax(reinterpret_cast<unsigned int*>(dst), __float_as_uint(v)); }   // v <= 0: larger bits = more negative
        else { if (v > 0.f) atomicMax(dst, __float_as_int(v)); }                                                      // v >= 0
    }
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ bounds,
                                                         u32* __restrict__ codes, u32* __restrict__ ids)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    u32 c = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = bounds[a], hi = bounds[3 + a];
        const float ext = hi - lo;
        float f = ext > 0.f ? ((pts[3 * (size_t)i + a] - lo) / ext) * 1023.f : 0.f;
        f = fminf(fmaxf(f, 0.f), 1023.f);
        c |= spread3((u32)f) << a;
    }
    codes[i] = c;
    ids[i] = (u32)i;
}

// gather into Morton order and build the 32-point leaf boxes (one warp per leaf)
__global__ void __launch_bounds__(256) knn_gather_kernel(int P, const float* __restrict__ pts, const u32* __restrict__ order,
                                                         float4* __restrict__ sorted, Box* __restrict__ leaf)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (i < P) {
        const u32 id = order[i];
        const float3 p = make_float3(pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2]);
        sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float(id));
        lo = p; hi = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && (i < P)) { Box b; b.lo = lo; b.hi = hi; leaf[i >> 5] = b; }
}

// one level up: box k = union of child boxes [32k, 32k+32)
__global__ void __launch_bounds__(256) knn_merge_kernel(int n_child, const Box* __restrict__ child, Box* __restrict__ parent)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (i < n_child) { lo = child[i].lo; hi = child[i].hi; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && i < n_child) { Box b; b.lo = lo; b.hi = hi; parent[i >> 5] = b; }
}

// simple_knn.cu:119-130
__device__ __forceinline__ float box_dist(const Box& b, const float3 p)
{
    float3 d = make_float3(0.f, 0.f, 0.f);
    if (p.x < b.lo.x || p.x > b.hi.x) d.x = fminf(fabsf(p.x - b.lo.x), fabsf(p.x - b.hi.x));
    if (p.y < b.lo.y || p.y > b.hi.y) d.y = fminf(fabsf(p.y - b.lo.y), fabsf(p.y - b.hi.y));
    if (p.z < b.lo.z || p.z > b.hi.z) d.z = fminf(fabsf(p.z - b.lo.z), fabsf(p.z - b.hi.z));
    return d.x * d.x + d.y * d.y + d.z * d.z;
}

// simple_knn.cu:132-145 (K = 3)
__device__ __forceinline__ void push3(const float3 ref, const float4 q, float* best)
{
    const float3 d = make_float3(q.x - ref.x, q.y - ref.y, q.z - ref.z);
    float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (best[j] > dist) { float t = best[j]; best[j] = dist; dist = t; }
}

__global__ void __launch_bounds__(128)
knn_search_kernel(int P, const float4* __restrict__ sorted, const Box* __restrict__ L1, int n1,
                  const Box* __restrict__ L2, int n2, const Box* __restrict__ L3, int n3, float* __restrict__ out)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= P) return;
    const float4 me = sorted[i];
    const float3 p = make_float3(me.x, me.y, me.z);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int j = max(0, i - 3); j <= min(P - 1, i + 3); ++j)
        if (j != i) push3(p, sorted[j], best);
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;
    for (int b3 = 0; b3 < n3; ++b3) {
        float d3 = box_dist(L3[b3], p);
        if (d3 > reject || d3 > best[2]) continue;
        const int e2 = min(n2, (b3 + 1) * 32);
        for (int b2 = b3 * 32; b2 < e2; ++b2) {
            float d2 = box_dist(L2[b2], p);
            if (d2 > reject || d2 > best[2]) continue;
            const int e1 = min(n1, (b2 + 1) * 32);
            for (int b1 = b2 * 32; b1 < e1; ++b1) {
                float d1 = box_dist(L1[b1], p);
                if (d1 > reject || d1 > best[2]) continue;
                const int e0 = min(P, (b1 + 1) * 32);
                for (int j = b1 * 32; j < e0; ++j)
                    if (j != i) push3(p, sorted[j], best);
            }
        }
    }
    out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct KnnPlan { size_t n1, n2, n3; size_t bytes; size_t sort_temp; };
KnnPlan knn_plan(size_t P, void* base, float** bounds, u32** ca, u32** cb, u32** ia, u32** ib, float4** sorted,
                 Box** L1, Box** L2, Box** L3, void** sort_temp)
{
    KnnPlan pl;
    pl.n1 = (P + 31) / 32; pl.n2 = (pl.n1 + 31) / 32; pl.n3 = (pl.n2 + 31) / 32;
    pl.sort_temp = radix_plan(P, 0, 30).temp_bytes;
    Carver c(base);
    float* b_ = c.take<float>(8);
    u32* ca_ = c.take<u32>(P); u32* cb_ = c.take<u32>(P); u32* ia_ = c.take<u32>(P); u32* ib_ = c.take<u32>(P);
    float4* s_ = c.take<float4>(P);
    Box* l1 = c.take<Box>(pl.n1); Box* l2 = c.take<Box>(pl.n2); Box* l3 = c.take<Box>(pl.n3);
    void* st = c.take<char>(pl.sort_temp);
    pl.bytes = c.used();
    if (bounds) { *bounds = b_; *ca = ca_; *cb = cb_; *ia = ia_; *ib = ib_; *sorted = s_; *L1 = l1; *L2 = l2; *L3 = l3; *sort_temp = st; }
    return pl;
}

}  // namespace

size_t dist2_scratch_bytes(size_t P)
{
    return knn_plan(P, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr).bytes;
}

int dist2(int P, const float* points, float* mean_dists, void* scratch, size_t scratch_bytes, cudaStream_t stream)
{
    if (P <= 0) return 0;
    float* bounds; u32 *ca, *cb, *ia, *ib; float4* sorted; Box *L1, *L2, *L3; void* st;
    KnnPlan pl = knn_plan((size_t)P, scratch, &bounds, &ca, &cb, &ia, &ib, &sorted, &L1, &L2, &L3, &st);
    if (scratch_bytes < pl.bytes) { set_error("dist2: scratch too small (%zu < %zu)", scratch_bytes, pl.bytes); return -1; }
    // bounds start at {0,0,0 | 0,0,0}: lows are accumulated as the bit pattern of a non-positive float
    // (0x80000000.. grows with magnitude), highs as a non-negative float's bits
    cudaMemsetAsync(bounds, 0, 8 * sizeof(float), stream);
    int grid = (P + 255) / 256; if (grid > NUM_SMS * 8) grid = NUM_SMS * 8;
    knn_bounds_kernel<<<grid, 256, 0, stream>>>(P, points, bounds);
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, bounds, ca, ia);
    int side = radix_sort_pairs(ca, ia, cb, ib, (size_t)P, 0, 30, st, pl.sort_temp, stream);
    if (side < 0) return -1;
    const u32* order = side ? ib : ia;
    knn_gather_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, order, sorted, L1);
    knn_merge_kernel<<<(unsigned)((pl.n1 + 255) / 256), 256, 0, stream>>>((int)pl.n1, L1, L2);
    knn_merge_kernel<<<(unsigned)((pl.n2 + 255) / 256), 256, 0, stream>>>((int)pl.n2, L2, L3);
    knn_search_kernel<<<(P + 127) / 128, 128, 0, stream>>>(P, sorted, L1, (int)pl.n1, L2, (int)pl.n2, L3, (int)pl.n3, mean_dists);
    return check_launch("dist2");
}

}  // namespace b200gs
