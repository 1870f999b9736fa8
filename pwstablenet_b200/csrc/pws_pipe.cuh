// pws_pipe.cuh -- pieces shared by the TMA-pipelined persistent kernels (warp_fwd_tma.cu, warp_bwd_tma.cu).
//
// Tile geometry, the shapes of the frame box a tile may request, and the "scout": the warp that turns
// a map tile (already in shared memory) into the bounding box of the source taps it implies and picks
// the smallest box shape that holds it.  The box IS the tile's halo under the map (north_star: "TMA
// staging of the source-tile halo implied by each output tile's warp-map bounding box").
#pragma once
#include "pws_tile.cuh"
#include "pws_tma.cuh"

namespace pws {
namespace pipe {

constexpr int kTW = 64, kTH = 16;                      // output pixels per tile
constexpr int kMapTileFloats = 2 * kTW * kTH;
constexpr int kMapTileBytes = kMapTileFloats * 4;
constexpr int kNumShapes = 3;
// Box shapes, smallest first (pixels x rows; all channels).  27 rows, not 28: four such slots, six map tiles, four
// grad_output tiles and the straggler queues of the backward then fit 227 KB.
// (Measured and dropped: 96-wide boxes, whose row pitch is a multiple of the 32 banks.  They remove the 2-way bank
// conflicts between lanes that sample different box rows -- 43 % of the backward's tap reads take a second wavefront --
// but the extra box bytes and the lower height limit cost more than the conflicts: 0.480 vs 0.467 ms.)
#ifndef PWS_BW0   // (overridable for A/B builds: tools/build_variant.py)
#define PWS_BW0 72
#define PWS_BH0 20
#define PWS_BW1 80
#define PWS_BH1 24
#define PWS_BW2 88
#define PWS_BH2 27
#endif
constexpr int kBW0 = PWS_BW0, kBH0 = PWS_BH0, kBW1 = PWS_BW1, kBH1 = PWS_BH1, kBW2 = PWS_BW2, kBH2 = PWS_BH2;
constexpr int kMaxBW = kBW2, kMaxBH = kBH2;
__host__ __device__ constexpr int box_w(int s) { return s == 0 ? kBW0 : s == 1 ? kBW1 : kBW2; }
__host__ __device__ constexpr int box_h(int s) { return s == 0 ? kBH0 : s == 1 ? kBH1 : kBH2; }
// channels-last RGB frames: the box is (3 * width) elements wide and a TMA box dimension holds at most 256
template <bool kCL> __host__ __device__ constexpr int box_w_of(int s) { return kCL && s == 2 ? 84 : box_w(s); }

// info.z of a tile: box shape in the low byte plus
enum : int {
    kInfoInterior = 1 << 8,   // every tap of every pixel is inside the frame and the tile is full: mask-free body
    kInfoFallback = 1 << 9,   // no box was loaded (does not fit / NaN / inf in the map): gather from global memory
    kInfoEmpty = 1 << 10      // no tap of the tile is inside the frame: nothing was loaded, nothing is read
};

__device__ __forceinline__ float fmin_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmax_nan(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin3_nan(float a, float b, float c) { float r; asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) { float r; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmin4(const float4 v) { return fmin_nan(fmin3_nan(v.x, v.y, v.z), v.w); }
__device__ __forceinline__ float fmax4(const float4 v) { return fmax_nan(fmax3_nan(v.x, v.y, v.z), v.w); }

// raw unnormalised coordinate (no clip, no guard): monotone in `coord`
template <bool kAlign>
__device__ __forceinline__ float unnorm(float coord, float size_f, float size_m1_f)
{
    const float t = __fadd_rn(coord, 1.0f);
    return kAlign ? __fmul_rn(__fmul_rn(t, 0.5f), size_m1_f) : __fmul_rn(__fmaf_rn(t, size_f, -1.0f), 0.5f);
}

// floor for 0 <= v < 2^22 without the conversion pipe: one round-down add against 1.5 * 2^23
__device__ __forceinline__ void floor_small(float v, float &vf, int &vi)
{
    const float t = __fadd_rd(v, 12582912.0f);
    vf = __fsub_rn(t, 12582912.0f);
    vi = __float_as_int(t) - 0x4B400000;
}

struct TileCoord { int n, h0, w0; };
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_x, int tiles_xy)
{
    TileCoord c;
    c.n = t / tiles_xy;
    const int r = t - c.n * tiles_xy;
    const int ty = r / tiles_x;
    c.h0 = ty * kTH;
    c.w0 = (r - ty * tiles_x) * kTW;
    return c;
}

// Walks tiles t0, t0 + step, t0 + 2*step, ... without a division per tile.
struct TileWalk {
    int n, ty, tx;        // current tile
    int dn, dy, dx;       // decomposition of `step`
    __device__ __forceinline__ void init(int t0, int step, int tiles_x, int tiles_y)
    {
        const int tiles_xy = tiles_x * tiles_y;
        n = t0 / tiles_xy; int r = t0 - n * tiles_xy; ty = r / tiles_x; tx = r - ty * tiles_x;
        dn = step / tiles_xy; r = step - dn * tiles_xy; dy = r / tiles_x; dx = r - dy * tiles_x;
    }
    __device__ __forceinline__ void next(int tiles_x, int tiles_y)
    {
        tx += dx; if (tx >= tiles_x) { tx -= tiles_x; ty += 1; }
        ty += dy; if (ty >= tiles_y) { ty -= tiles_y; n += 1; }
        n += dn;
    }
    __device__ __forceinline__ TileCoord coord() const { TileCoord c; c.n = n; c.h0 = ty * kTH; c.w0 = tx * kTW; return c; }
};

// Extremes of the x and y map values of a tile, over its valid rows x cols, NaN-propagating.
// Planar tile: [2][kTH][kTW]; interleaved: [kTH][kTW][2].  All 32 lanes of a warp call this.
template <bool kInter>
__device__ __forceinline__ void map_tile_range(const float *__restrict__ mp, int rows, int cols, int lane,
                                               float &xlo, float &xhi, float &ylo, float &yhi)
{
    xlo = INFINITY; xhi = -INFINITY; ylo = INFINITY; yhi = -INFINITY;
    if (rows == kTH && cols == kTW) {
        // full tile: 16 independent 128-bit reads per lane, then a tree
        float4 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = *reinterpret_cast<const float4 *>(mp + (k * 32 + lane) * 4);
        if (kInter) {
            // (x0,y0,x1,y1) quads: four independent chains per extreme
            float a[4], b[4], c[4], d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { a[k] = INFINITY; b[k] = -INFINITY; c[k] = INFINITY; d[k] = -INFINITY; }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                a[k & 3] = fmin3_nan(a[k & 3], v[k].x, v[k].z); b[k & 3] = fmax3_nan(b[k & 3], v[k].x, v[k].z);
                c[k & 3] = fmin3_nan(c[k & 3], v[k].y, v[k].w); d[k & 3] = fmax3_nan(d[k & 3], v[k].y, v[k].w);
            }
            xlo = fmin_nan(fmin_nan(a[0], a[1]), fmin_nan(a[2], a[3])); xhi = fmax_nan(fmax_nan(b[0], b[1]), fmax_nan(b[2], b[3]));
            ylo = fmin_nan(fmin_nan(c[0], c[1]), fmin_nan(c[2], c[3])); yhi = fmax_nan(fmax_nan(d[0], d[1]), fmax_nan(d[2], d[3]));
        } else {
            // v[0..7] lie in the x plane, v[8..15] in the y plane
            float lo[16], hi[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { lo[k] = fmin4(v[k]); hi[k] = fmax4(v[k]); }
            xlo = fmin_nan(fmin3_nan(lo[0], lo[1], lo[2]), fmin_nan(fmin3_nan(lo[3], lo[4], lo[5]), fmin_nan(lo[6], lo[7])));
            xhi = fmax_nan(fmax3_nan(hi[0], hi[1], hi[2]), fmax_nan(fmax3_nan(hi[3], hi[4], hi[5]), fmax_nan(hi[6], hi[7])));
            ylo = fmin_nan(fmin3_nan(lo[8], lo[9], lo[10]), fmin_nan(fmin3_nan(lo[11], lo[12], lo[13]), fmin_nan(lo[14], lo[15])));
            yhi = fmax_nan(fmax3_nan(hi[8], hi[9], hi[10]), fmax_nan(fmax3_nan(hi[11], hi[12], hi[13]), fmax_nan(hi[14], hi[15])));
        }
    } else if (kInter) {
        const int c2 = lane * 2;  // rows of 64 (x,y) pairs: a lane covers 2 pairs of every row
        if (c2 < cols)
            for (int r = 0; r < rows; ++r) {
                const float4 v = *reinterpret_cast<const float4 *>(mp + r * (2 * kTW) + 2 * c2);
                xlo = fmin3_nan(xlo, v.x, v.z); xhi = fmax3_nan(xhi, v.x, v.z);
                ylo = fmin3_nan(ylo, v.y, v.w); yhi = fmax3_nan(yhi, v.y, v.w);
            }
    } else {
        const int c4 = (lane & 15) * 4;
        if (c4 < cols)
            for (int r = lane >> 4; r < rows; r += 2) {
                const float4 vx = *reinterpret_cast<const float4 *>(mp + r * kTW + c4);
                const float4 vy = *reinterpret_cast<const float4 *>(mp + kTW * kTH + r * kTW + c4);
                xlo = fmin_nan(xlo, fmin4(vx)); xhi = fmax_nan(xhi, fmax4(vx));
                ylo = fmin_nan(ylo, fmin4(vy)); yhi = fmax_nan(yhi, fmax4(vy));
            }
    }
    // one warp-wide reduction instruction per extreme (sm_100a: redux.sync on f32 -> CREDUX)
    asm volatile("redux.sync.min.NaN.f32 %0, %1, 0xffffffff;" : "=f"(xlo) : "f"(xlo));
    asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(xhi) : "f"(xhi));
    asm volatile("redux.sync.min.NaN.f32 %0, %1, 0xffffffff;" : "=f"(ylo) : "f"(ylo));
    asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(yhi) : "f"(yhi));
}

// The same extremes read straight from global memory, in two halves so that a scout can keep the loads of its NEXT
// tile in flight while it deals with the current one (a global load takes ~3 us under this kernel's traffic; the
// scout would otherwise spend most of a tile time waiting for it).
// `tp` points at the tile's first x coordinate; s1 = row stride, s3 = stride between the x and y coordinates
// (planar maps) -- interleaved maps hold (x,y) pairs.  Alignment is the TMA eligibility of encode_map_tma.
// volatile: the load must be ISSUED where it is written (the compiler otherwise sinks it to its first use to save
// registers, which puts the whole memory latency back on the scout's path)
__device__ __forceinline__ float4 ldg_v4(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
struct MapRegs { float4 v[kTH]; };
// loads of a FULL tile (kTH rows x kTW columns): 16 independent 128-bit loads per lane
template <bool kInter>
__device__ __forceinline__ void map_tile_fetch(const float *__restrict__ tp, int s1, int s3, int lane, MapRegs &m)
{
    if (kInter) {
#pragma unroll
        for (int k = 0; k < kTH; ++k) m.v[k] = ldg_v4(tp + k * s1 + 4 * lane);  // a row is 128 floats: one per lane and row
    } else {
        const int c4 = (lane & 15) * 4, rr = lane >> 4;  // 16 lanes per row of 64, two rows per instruction
#pragma unroll
        for (int k = 0; k < kTH / 2; ++k) { m.v[k] = ldg_v4(tp + (2 * k + rr) * s1 + c4); m.v[kTH / 2 + k] = ldg_v4(tp + s3 + (2 * k + rr) * s1 + c4); }
    }
}
__device__ __forceinline__ void warp_range(float &xlo, float &xhi, float &ylo, float &yhi)
{
    // one warp-wide reduction instruction per extreme (sm_100a: redux.sync on f32 -> CREDUX)
    asm volatile("redux.sync.min.NaN.f32 %0, %1, 0xffffffff;" : "=f"(xlo) : "f"(xlo));
    asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(xhi) : "f"(xhi));
    asm volatile("redux.sync.min.NaN.f32 %0, %1, 0xffffffff;" : "=f"(ylo) : "f"(ylo));
    asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(yhi) : "f"(yhi));
}
template <bool kInter>
__device__ __forceinline__ void map_tile_reduce(const MapRegs &m, float &xlo, float &xhi, float &ylo, float &yhi)
{
    xlo = INFINITY; xhi = -INFINITY; ylo = INFINITY; yhi = -INFINITY;
    if (kInter) {
#pragma unroll
        for (int k = 0; k < kTH; ++k) {
            xlo = fmin3_nan(xlo, m.v[k].x, m.v[k].z); xhi = fmax3_nan(xhi, m.v[k].x, m.v[k].z);
            ylo = fmin3_nan(ylo, m.v[k].y, m.v[k].w); yhi = fmax3_nan(yhi, m.v[k].y, m.v[k].w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kTH / 2; ++k) {
            xlo = fmin_nan(xlo, fmin4(m.v[k])); xhi = fmax_nan(xhi, fmax4(m.v[k]));
            ylo = fmin_nan(ylo, fmin4(m.v[kTH / 2 + k])); yhi = fmax_nan(yhi, fmax4(m.v[kTH / 2 + k]));
        }
    }
    warp_range(xlo, xhi, ylo, yhi);
}
// partial tiles (the last tile row / column of a frame): load and reduce in one go
template <bool kInter>
__device__ __forceinline__ void map_tile_range_global(const float *__restrict__ tp, int s1, int s3, int rows, int cols, int lane,
                                                      float &xlo, float &xhi, float &ylo, float &yhi)
{
    xlo = INFINITY; xhi = -INFINITY; ylo = INFINITY; yhi = -INFINITY;
    if (kInter) {
        if (2 * lane < cols)  // a lane covers 2 pixels of every row (cols is a multiple of 4)
            for (int r = 0; r < rows; ++r) {
                const float4 v = ldg_v4(tp + r * s1 + 4 * lane);
                xlo = fmin3_nan(xlo, v.x, v.z); xhi = fmax3_nan(xhi, v.x, v.z);
                ylo = fmin3_nan(ylo, v.y, v.w); yhi = fmax3_nan(yhi, v.y, v.w);
            }
    } else {
        const int c4 = (lane & 15) * 4;
        if (c4 < cols)
            for (int r = lane >> 4; r < rows; r += 2) {
                const float4 vx = ldg_v4(tp + r * s1 + c4), vy = ldg_v4(tp + s3 + r * s1 + c4);
                xlo = fmin_nan(xlo, fmin4(vx)); xhi = fmax_nan(xhi, fmax4(vx));
                ylo = fmin_nan(ylo, fmin4(vy)); yhi = fmax_nan(yhi, fmax4(vy));
            }
    }
    warp_range(xlo, xhi, ylo, yhi);
}

// From the extremes of a tile's map to the box of the frame its taps need: (bx, by, shape | flags).
// unnormalise is monotone, so the extremes of the map give the extremes of the taps.
// kXAlign: pixels per 16 bytes of a frame row (4 for fp32, 8 for 16-bit frames) -- TMA wants the box start aligned
template <bool kBorder, bool kAlign, bool kCL = false, int kXAlign = 4>
__device__ __forceinline__ int4 box_of_range(float xlo, float xhi, float ylo, float yhi, int W, int H, bool full_tile)
{
    const float Wf = (float)W, Hf = (float)H, Wm1 = (float)(W - 1), Hm1 = (float)(H - 1);
    const float fxlo = unnorm<kAlign>(xlo, Wf, Wm1), fxhi = unnorm<kAlign>(xhi, Wf, Wm1);
    const float fylo = unnorm<kAlign>(ylo, Hf, Hm1), fyhi = unnorm<kAlign>(yhi, Hf, Hm1);
    // NaN / inf / beyond +-2^30: ATen's -100 guard breaks the monotonicity -> generic path
    const bool sane = fxlo >= -1073741824.0f && fxhi <= 1073741824.0f && fylo >= -1073741824.0f && fyhi <= 1073741824.0f;
    if (!sane) return make_int4(0, 0, kInfoFallback, 0);
    int x0lo = __float2int_rd(fxlo), x0hi = __float2int_rd(fxhi);
    int y0lo = __float2int_rd(fylo), y0hi = __float2int_rd(fyhi);
    // (border padding zeroes the map gradient of a coordinate sitting exactly on 0: keep such tiles masked)
    const bool interior = x0lo >= 0 && x0hi + 1 <= W - 1 && y0lo >= 0 && y0hi + 1 <= H - 1 && full_tile &&
                          (!kBorder || (fxlo > 0.0f && fylo > 0.0f));
    if (kBorder) {  // border padding clips the coordinate into [0, size-1] before the floor
        x0lo = clampi(x0lo, 0, W - 1); x0hi = clampi(x0hi, 0, W - 1);
        y0lo = clampi(y0lo, 0, H - 1); y0hi = clampi(y0hi, 0, H - 1);
    }
    // taps outside the frame are never read: clamp the box to the frame.
    // TMA wants the box start 16-byte aligned: x rounds down to a multiple of 4 elements.
    const int bx = max(x0lo, 0) & ~(kXAlign - 1), by = max(y0lo, 0);
    const int bw = min(x0hi + 1, W - 1) - bx + 1, bh = min(y0hi + 1, H - 1) - by + 1;
    if (bw <= 0 || bh <= 0) return make_int4(0, 0, kInfoEmpty, 0);
    const int shape = bw <= box_w_of<kCL>(0) && bh <= kBH0 ? 0 : bw <= box_w_of<kCL>(1) && bh <= kBH1 ? 1 : bw <= box_w_of<kCL>(2) && bh <= kBH2 ? 2 : -1;
    if (shape < 0) return make_int4(0, 0, kInfoFallback, 0);
    return make_int4(bx, by, shape | (interior ? kInfoInterior : 0), 0);
}

}  // namespace pipe

// host side (warp_fwd_tma.cu)
bool encode_map_tma(const View &grid, const Geometry &g, CUtensorMap *tm, bool *inter);
bool encode_frame_tma(const View &v, int W, int H, int C, int N, int bw, int bh, int bc, CUtensorMap *tm, int dtype = PWS_F32);
bool tma_disabled();
int sm_count();

}  // namespace pws
