// Layout of the three caller-owned scratch buffers of the rasterizer.
//
// They take the place of the reference's GeometryState / BinningState / ImageState
// (RAST/cuda_rasterizer/rasterizer_impl.h:21-73) but are laid out for this engine:
// per-Gaussian render attributes are packed into three float4 records so compositing
// gathers them with 128-bit loads, and the binning buffer holds a depth-sorted Gaussian
// list plus 32-bit (tile, gaussian) instance pairs instead of 64-bit (tile|depth) keys.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace b200gs {

struct GeomState {
    u32*    depth_key;     // [P] float bits of view-space depth; 0xFFFFFFFF if culled
    u32*    order_iota;    // [P] 0..P-1 (payload of the depth sort)
    u32*    tiles_touched; // [P]
    uint2*  rect;          // [P] (xmin | xmax<<16, ymin | ymax<<16) in tiles
    float4* recA;          // [P] (mean2D.x, mean2D.y, conic.a, conic.b)
    float4* recB;          // [P] (conic.c, opacity, depth, reject_power)
    float4* recC;          // [P] (r, g, b, 0)
    float*  cov3D;         // [6P]
    unsigned char* clamped;// [P] bit c set <=> channel c was clamped at 0
    unsigned long long* counters; // [4]: num_rendered, num_visible, -, -
    u32*    gkeys_b;       // [P] ping-pong of the depth sort (stage 1 queues the sort behind the counter read-back)
    u32*    gvals_b;       // [P]
    void*   sort_temp; size_t sort_temp_bytes;     // histograms / look-back state of the depth sort
    static GeomState carve(void* buf, size_t P, size_t* bytes) {
        Carver c(buf);
        GeomState g;
        g.depth_key = c.take<u32>(P);
        g.order_iota = c.take<u32>(P);
        g.tiles_touched = c.take<u32>(P);
        g.rect = c.take<uint2>(P);
        g.recA = c.take<float4>(P);
        g.recB = c.take<float4>(P);
        g.recC = c.take<float4>(P);
        g.cov3D = c.take<float>(6 * P);
        g.clamped = c.take<unsigned char>(P);
        g.counters = c.take<unsigned long long>(4);
        g.gkeys_b = c.take<u32>(P);
        g.gvals_b = c.take<u32>(P);
        g.sort_temp_bytes = radix_plan(P, 0, 32).temp_bytes;
        g.sort_temp = c.take<char>(g.sort_temp_bytes);
        if (bytes) *bytes = c.used();
        return g;
    }
};

struct BinState {
    u32* ikeys_a;          // [R] tile id per instance (unsorted, depth order)
    u32* ivals_a;          // [R] gaussian id per instance
    u32* ikeys_b;          // [R]
    u32* ivals_b;          // [R]
    u64* emit_status;      // [ceil(P/256)+1] look-back state of the emission scan
    u32* emit_ticket;      // [64]
    void* sort_temp; size_t sort_temp_bytes;
    // filled in by stage 2 (host side only; recomputed by carve from R and P):
    static BinState carve(void* buf, size_t P, size_t R, int tile_bits, size_t* bytes) {
        Carver c(buf);
        BinState b;
        b.ikeys_a = c.take<u32>(R);
        b.ivals_a = c.take<u32>(R);
        b.ikeys_b = c.take<u32>(R);
        b.ivals_b = c.take<u32>(R);
        b.emit_status = c.take<u64>((P + 255) / 256 + 1);
        b.emit_ticket = c.take<u32>(64);
        b.sort_temp_bytes = radix_plan(R, 0, tile_bits).temp_bytes;
        b.sort_temp = c.take<char>(b.sort_temp_bytes);
        if (bytes) *bytes = c.used();
        return b;
    }
};

struct ImgState {
    float* final_T;        // [W*H]
    u32*   n_contrib;      // [W*H]
    uint2* ranges;         // [tiles]
    u32*   which;          // [4]: [0] = which ping-pong side holds the sorted instance list
    static ImgState carve(void* buf, size_t npix, size_t tiles, size_t* bytes) {
        Carver c(buf);
        ImgState s;
        s.final_T = c.take<float>(npix);
        s.n_contrib = c.take<u32>(npix);
        s.ranges = c.take<uint2>(tiles);
        s.which = c.take<u32>(4);
        if (bytes) *bytes = c.used();
        return s;
    }
};


// Conservative per-warp culling for the compositing kernels. A warp covers the pixel-centre box
// [bx, bx+7] x [by, by+3]; a splat (centre x,y; conic a,b,c; exponent bound `reject`) can only
// contribute to one of those pixels if max over the box of power = -0.5 q(d) reaches `reject`,
// q(d) = a dx^2 + 2 b dx dy + c dy^2, d = centre - pixel.  The minimum of the convex q over the box is
// attained at the box point closest to the centre along one of the two facing edges, so two
// candidates suffice.  Anything doubtful (non-PD conic, NaNs) is kept; survivors still run the exact
// per-pixel test, so results are unchanged.
__device__ __forceinline__ bool splat_may_touch_patch(const float4 A, const float4 B, float bx, float by, float patch_h = 3.f)
{
    const float a = A.z, b = A.w, c = B.x;
    if (!(a > 0.f) || !(c > 0.f)) return true;
    const float dx_hi = A.x - bx, dx_lo = dx_hi - 7.f;          // d.x range over the patch
    const float dy_hi = A.y - by, dy_lo = dy_hi - patch_h;          // patch_h = rows - 1 (8x4 patch: 3, 8x8 patch: 7)
    const float cx = fminf(fmaxf(0.f, dx_lo), dx_hi), cy = fminf(fmaxf(0.f, dy_lo), dy_hi);
    const float y1 = fminf(fmaxf(__fdividef(-b * cx, c), dy_lo), dy_hi);
    const float x2 = fminf(fmaxf(__fdividef(-b * cy, a), dx_lo), dx_hi);
    const float q1 = a * cx * cx + 2.f * b * cx * y1 + c * y1 * y1;
    const float q2 = a * x2 * x2 + 2.f * b * x2 * cy + c * cy * cy;
    const float pmax = -0.5f * fminf(q1, q2);
    return !(pmax < B.w - 1e-3f);
}

inline int tile_bits_for(size_t tiles) {
    int b = 1;
    while (((size_t)1 << b) < tiles) ++b;
    return b;
}

}  // namespace b200gs
