// Packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE single-precision operations per issued instruction).
// The compositing kernels are issue-slot bound (75 % of issue slots busy with the FMA pipe at 34 %), so a lane that owns TWO
// pixels and carries their per-pixel state as pairs halves the issued FP32 instructions.  A scalar that is the same for both
// pixels (a splat constant) is written pk(s, s); ptxas folds that into the instruction's scalar-broadcast operand form
// (`FFMA2 R8, R8.F32x2.HI_LO, R19.F32, R12.F32x2.HI_LO`), no register moves.
#pragma once
#include <cuda_runtime.h>

namespace b200gs {

typedef unsigned long long f2;   // {lo, hi} = two floats

__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 pk1(float s) { return pk(s, s); }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

}  // namespace b200gs
