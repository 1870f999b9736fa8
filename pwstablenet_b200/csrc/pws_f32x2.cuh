// pws_f32x2.cuh -- packed fp32 pair arithmetic of sm_100 (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2).
//
// One issue slot does two IEEE fp32 operations, each half rounded exactly like its scalar counterpart (.rn, no .ftz),
// so results stay bit-identical to the scalar statement order of ATen's kernels.  SASS takes a scalar register or an
// immediate as a broadcast operand (FMUL2 R12, R22.F32, -R8.F32x2.HI_LO), which ptxas finds when a half pair is built
// from one value: make_float2(s, s) costs nothing.  The issue-bound backward kernel uses these for the coordinate
// pipeline, the grad_grid accumulation ((gy, gx) as one pair) and the scatter's weight products.
#pragma once
#include <cuda_runtime.h>

namespace pws {
namespace x2 {

__device__ __forceinline__ unsigned long long pack(float2 a)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 unpack(unsigned long long r)
{
    float2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
    return a;
}
__device__ __forceinline__ float2 add(float2 a, float2 b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack(a)), "l"(pack(b)));
    return unpack(r);
}
__device__ __forceinline__ float2 add_rm(float2 a, float2 b)  // both halves rounded toward -inf
{
    unsigned long long r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack(a)), "l"(pack(b)));
    return unpack(r);
}
__device__ __forceinline__ float2 sub(float2 a, float2 b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack(a)), "l"(pack(b)));
    return unpack(r);
}
__device__ __forceinline__ float2 mul(float2 a, float2 b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack(a)), "l"(pack(b)));
    return unpack(r);
}
__device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack(a)), "l"(pack(b)), "l"(pack(c)));
    return unpack(r);
}
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }  // broadcast operand

}  // namespace x2
}  // namespace pws
