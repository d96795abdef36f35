// In-house onesweep LSD radix sort (stable) over 32-bit keys with 32-bit payloads.
//
// Replaces the reference's cub::DeviceRadixSort::SortPairs call sites
// (RAST/cuda_rasterizer/rasterizer_impl.cu:304-309, KNN/simple_knn.cu:208-213).
// One histogram kernel reads the keys once and builds the digit histograms of every
// pass; each pass is then ONE kernel: tiles of 4096 keys are ranked with warp-level
// match_any ballots, the per-tile digit counts are chained between tiles with a
// decoupled look-back, and the keys leave through shared memory so the global stores
// are contiguous per digit run.  All kernels are HBM-bound: per pass 8 B read +
// 8 B written per (key, value) pair.
#pragma once
#include "common.cuh"

namespace b200gs {

constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_RADIX_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 16;                        // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;      // 4096 keys per tile
constexpr int RS_MAX_PASSES = 4;

struct RadixPlan {
    int passes;
    size_t tiles;
    size_t temp_bytes;      // histograms + look-back state + tickets
};

inline RadixPlan radix_plan(size_t n, int begin_bit, int end_bit) {
    RadixPlan p;
    int bits = end_bit - begin_bit;
    p.passes = (bits + RS_RADIX_BITS - 1) / RS_RADIX_BITS;
    if (p.passes < 1) p.passes = 1;
    p.tiles = (n + RS_TILE - 1) / RS_TILE;
    if (p.tiles == 0) p.tiles = 1;
    // [passes][256] global digit bases, [passes] tickets (padded to 256 u32), [passes][tiles][256] look-back
    p.temp_bytes = align_up(((size_t)p.passes * RS_RADIX + 256 + (size_t)p.passes * p.tiles * RS_RADIX) * sizeof(u32), 256);
    return p;
}

// Sorts n (key, value) pairs on key bits [begin_bit, end_bit). keys_a/vals_a hold the input;
// keys_b/vals_b are same-sized ping-pong buffers. Returns 0 if the sorted result ends up in
// the *_a buffers, 1 if in the *_b buffers, -1 on error. vals may be null (keys only).
int radix_sort_pairs(u32* keys_a, u32* vals_a, u32* keys_b, u32* vals_b, size_t n,
                     int begin_bit, int end_bit, void* temp, size_t temp_bytes,
                     cudaStream_t stream);

}  // namespace b200gs
