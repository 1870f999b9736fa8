// pws_tile.cuh -- tile machinery shared by the staged forward and backward kernels.
//
// A CTA owns a TW x TH tile of OUTPUT pixels (lane <-> x, so shared-memory gathers
// of near-identity maps are bank-conflict free and global map reads / output
// writes are 128-byte coalesced per warp instruction).
//   1. the warp map of the tile is read (all loads of a thread issued back to back)
//      and turned into source indices;
//   2. the bounding box of the source taps the tile needs -- its halo under the
//      map -- is reduced across the CTA, and that box of the frame is staged into
//      shared memory with 128-bit loads;
//   3. the gather runs out of shared memory with compile-time pitches: the four
//      taps of a channel are one base register plus immediates.  The direct-gather
//      kernels spend most of their issue slots on 64-bit address arithmetic and
//      per-tap predicates (profiles/r01_fwd_direct.txt); this path does not.
// Tiles whose box does not fit the shared-memory budget (violent maps) gather from
// global memory inside the same kernel; results are identical either way.
#pragma once
#include "pws_common.cuh"

namespace pws {

// ---- specialised coordinate pipeline (padding / align_corners are template flags) --------
template <bool kBorder, bool kAlign>
__device__ __forceinline__ float src_index(float coord, int size, float size_f, float size_m1_f)
{
    const float t = __fadd_rn(coord, 1.0f);
    float c = kAlign ? __fmul_rn(__fmul_rn(t, 0.5f), size_m1_f) : __fmul_rn(__fmaf_rn(t, size_f, -1.0f), 0.5f);
    if (kBorder) c = fminf(size_m1_f, fmaxf(c, 0.0f));
    // safe_downgrade_to_int_range: (float)(INT_MAX-1) == 2^31, (float)INT_MIN == -2^31
    if (!(c <= 2147483648.0f && c >= -2147483648.0f)) c = -100.0f;  // also catches NaN; inf fails the range
    (void)size;
    return c;
}

// backward flavour: the multiplier d(index)/d(coord), zero where border padding clips
template <bool kBorder, bool kAlign>
__device__ __forceinline__ float src_index_grad(float coord, float size_f, float size_m1_f, float *gm)
{
    const float t = __fadd_rn(coord, 1.0f);
    float c = kAlign ? __fmul_rn(__fmul_rn(t, 0.5f), size_m1_f) : __fmul_rn(__fmaf_rn(t, size_f, -1.0f), 0.5f);
    float m = kAlign ? size_m1_f * 0.5f : size_f * 0.5f;
    if (kBorder) {
        if (c <= 0.0f) { c = 0.0f; m = 0.0f; }
        else if (c >= size_m1_f) { c = size_m1_f; m = 0.0f; }
    }
    if (!(c <= 2147483648.0f && c >= -2147483648.0f)) c = -100.0f;
    *gm = m;
    return c;
}

struct Box {
    int x0, y0;      // frame coordinates of shared-memory element (0,0); x0 is vector-aligned
    int w, h;        // staged extent (w a multiple of the vector width)
    bool fits;       // box fits the shared-memory budget
    bool all_valid;  // every tap of every pixel of the tile lies inside the frame
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// CTA-wide reduction of the per-thread floor(source index) ranges.  One barrier.
//   flo/fhi: min / max over the thread's pixels of floorf(ix) (resp. iy); a thread
//   without pixels passes +inf / -inf.
template <int kWarps, int VE, int BW, int BH>
__device__ __forceinline__ Box block_box(float fxlo, float fxhi, float fylo, float fyhi, int W, int H, int *s_part)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    // float -> int saturates; the +-inf of pixel-less threads become INT_MAX / INT_MIN
    int lox = __float2int_rd(fxlo), hix = __float2int_rd(fxhi), loy = __float2int_rd(fylo), hiy = __float2int_rd(fyhi);
    lox = __reduce_min_sync(0xffffffffu, lox);
    hix = __reduce_max_sync(0xffffffffu, hix);
    loy = __reduce_min_sync(0xffffffffu, loy);
    hiy = __reduce_max_sync(0xffffffffu, hiy);
    if (lane == 0) reinterpret_cast<int4 *>(s_part)[wrp] = make_int4(lox, hix, loy, hiy);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kWarps; ++k) {
        const int4 p = reinterpret_cast<const int4 *>(s_part)[k];
        lox = min(lox, p.x); hix = max(hix, p.y); loy = min(loy, p.z); hiy = max(hiy, p.w);
    }
    Box b;
    b.all_valid = lox >= 0 && hix <= W - 2 && loy >= 0 && hiy <= H - 2;
    if (hix < lox) { b.x0 = 0; b.y0 = 0; b.w = 0; b.h = 0; b.fits = true; b.all_valid = false; return b; }
    // taps are (x0, x0+1) x (y0, y0+1); out-of-frame taps are never read, so clamp
    const int cxlo = clampi(lox, 0, W - 1), cxhi = clampi(hix < INT_MAX ? hix + 1 : hix, 0, W - 1);
    const int cylo = clampi(loy, 0, H - 1), cyhi = clampi(hiy < INT_MAX ? hiy + 1 : hiy, 0, H - 1);
    b.x0 = cxlo & ~(VE - 1);
    b.y0 = cylo;
    b.w = ((cxhi - b.x0) / VE + 1) * VE;
    b.h = cyhi - cylo + 1;
    b.fits = (b.w <= BW) && (b.h <= BH);
    return b;
}

template <int VE> struct VecOf;
template <> struct VecOf<4> { using type = float4; };  // 4 x f32
template <> struct VecOf<8> { using type = uint4; };   // 8 x 16-bit

// Stage rows [b.y0, b.y0+b.h) x columns [b.x0, b.x0+b.w) of CS channel planes
// (W-contiguous) into smem[CS][BH][BW] with 16-byte copies; the host guarantees the
// alignment of every row start.  BW/VE <= 32: one lane per vector of a row, a warp
// per row, the loads of a channel issued together.
template <typename T, int CS, int VE, int BW, int BH, int kWarps>
__device__ __forceinline__ void stage_planar(const T *__restrict__ ip, int64_t sC, int sH, const Box &b, T *__restrict__ smem)
{
    static_assert(BW / VE <= 32, "one lane per row vector");
    using V = typename VecOf<VE>::type;
    constexpr int kIter = (BH + kWarps - 1) / kWarps;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const bool lane_on = lane < b.w / VE;
    const V *src = reinterpret_cast<const V *>(ip + (int64_t)(b.y0 + wrp) * sH + b.x0) + lane;
    V *dst = reinterpret_cast<V *>(smem + wrp * BW) + lane;
    const int64_t row_step = (int64_t)kWarps * sH / VE;  // in vectors
#pragma unroll
    for (int c = 0; c < CS; ++c) {
        V v[kIter];
#pragma unroll
        for (int k = 0; k < kIter; ++k)
            if (lane_on && wrp + k * kWarps < b.h) v[k] = __ldg(src + k * row_step);
#pragma unroll
        for (int k = 0; k < kIter; ++k)
            if (lane_on && wrp + k * kWarps < b.h) dst[k * (kWarps * BW / VE)] = v[k];
        src += sC / VE;
        dst += BH * BW / VE;
    }
}

// Can the planes of this view be copied with 16-byte vectors of VE elements?
inline bool rows_vectorizable(const View &v, int VE, int W)
{
    return v.s3 == 1 && (reinterpret_cast<uintptr_t>(v.p) % 16 == 0) && (v.sN % VE == 0) && (v.s1 % VE == 0) &&
           (v.s2 % VE == 0) && (W % VE == 0);
}

}  // namespace pws
