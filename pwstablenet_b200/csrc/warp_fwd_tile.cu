// warp_fwd_tile.cu -- lean direct-gather forward warp for the reference's hot layouts (fallback of warp_fwd_tma.cu).
//
// Call sites: R/main_new.py:106,116 (NCHW frames, planar-stored maps), :197 (NCHW
// frames, interleaved affine_grid maps), :716 (same at native video resolution).
// Host-checked requirements: frame and output W-contiguous and 16-byte row aligned,
// fp32 map, C in {1,3}, f32 / f16 / bf16 frames.  Everything else takes the direct
// kernel (warp_fwd.cu).
#include "pws_tile.cuh"

#include <cstdlib>

namespace pws {

namespace {

constexpr int kTW = 64, kTH = 32, kBW = 80, kBH = 40, kThreads = 256, kWarps = 8;
constexpr int kPX = kTW / 32, kPY = kTH / kWarps;

// Layout-specialised direct gather: taps through L1 (no staging, no barrier).  4 pixels per thread, 64x16 tile.
template <typename T, int CS, bool kBorder, bool kAlign>
#ifndef PWS_FWD_MINB
#define PWS_FWD_MINB 5   // 48 registers, 40 warps/SM: measured ~7 % faster than 4 CTAs at 56 registers
#endif
__global__ void __launch_bounds__(kThreads, PWS_FWD_MINB)
fwd_lean_kernel(const View in, const View grid, const View out, const Geometry g)
{
    constexpr int PY = 2;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int n = blockIdx.z;
    const int w_base = blockIdx.x * kTW + lane, h_base = blockIdx.y * (kWarps * PY) + wrp;
    const float *__restrict__ gp = (const float *)grid.p + (int64_t)n * grid.sN;
    const T *__restrict__ ip = (const T *)in.p + (int64_t)n * in.sN;
    T *__restrict__ op = (T *)out.p + (int64_t)n * out.sN;
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
    const bool full = (blockIdx.x + 1) * kTW <= g.Wo && (blockIdx.y + 1) * (kWarps * PY) <= g.Ho;

    float sx[PY][kPX], sy[PY][kPX];
    {
        const bool inter = grid.s3 == 1;
        const float *__restrict__ q = gp + (int64_t)h_base * grid.s1 + (int64_t)w_base * grid.s2;
        const int64_t row = (int64_t)kWarps * grid.s1;
        const int col = 32 * grid.s2;
#pragma unroll
        for (int j = 0; j < PY; ++j)
#pragma unroll
            for (int i = 0; i < kPX; ++i) {
                sx[j][i] = -4.0f; sy[j][i] = -4.0f;
                if (full || (w_base + 32 * i < g.Wo && h_base + kWarps * j < g.Ho)) {
                    const float *__restrict__ a = q + j * row + i * col;
                    if (inter) {
                        const float2 v = __ldg(reinterpret_cast<const float2 *>(a));
                        sx[j][i] = v.x; sy[j][i] = v.y;
                    } else {
                        sx[j][i] = __ldg(a);
                        sy[j][i] = __ldg(a + grid.s3);
                    }
                }
            }
    }
    const int o_row = kWarps * out.s2, o_ch = out.s1, i_ch = in.s1;  // 32-bit in-frame offsets
    const int sH = in.s2;
    T *__restrict__ o0 = op + (int64_t)h_base * out.s2 + w_base;
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int i = 0; i < kPX; ++i) {
            const bool ok = full || (w_base + 32 * i < g.Wo && h_base + kWarps * j < g.Ho);
            const float ix = src_index<kBorder, kAlign>(sx[j][i], g.W, Wf, Wm1);
            const float iy = src_index<kBorder, kAlign>(sy[j][i], g.H, Hf, Hm1);
            const float x0f = floorf(ix), y0f = floorf(iy);
            const bool inside = x0f >= 0.0f && x0f <= Wm1 - 1.0f && y0f >= 0.0f && y0f <= Hm1 - 1.0f;
            T *__restrict__ o = o0 + (j * o_row + 32 * i);
            if (__all_sync(0xffffffffu, inside && ok)) {
                const float wx1 = fsub(ix, x0f), wx0 = fsub(x0f + 1.0f, ix);
                const float wy1 = fsub(iy, y0f), wy0 = fsub(y0f + 1.0f, iy);
                const float nw = fmul(wx0, wy0), ne = fmul(wx1, wy0), sw = fmul(wx0, wy1), se = fmul(wx1, wy1);
                const T *__restrict__ p0 = ip + ((int)y0f * sH + (int)x0f);
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    const T *__restrict__ pc = p0 + c * i_ch;
                    const float v0 = to_acc(ldg(pc)), v1 = to_acc(ldg(pc + 1));
                    const float v2 = to_acc(ldg(pc + sH)), v3 = to_acc(ldg(pc + sH + 1));
                    float acc = ffma(v0, nw, 0.f);
                    acc = ffma(v1, ne, acc);
                    acc = ffma(v2, sw, acc);
                    acc = ffma(v3, se, acc);
                    o[c * o_ch] = from_acc<T, float>(acc);
                }
            } else if (ok) {
                Taps<float> t;
                make_taps(ix, iy, g.H, g.W, t);
                const T *__restrict__ p0 = ip + (t.y0 * sH + t.x0);
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    const T *__restrict__ pc = p0 + c * i_ch;
                    float acc = 0.f;
                    if (t.mask & 1u) acc = ffma(to_acc(ldg(pc)), t.nw, acc);
                    if (t.mask & 2u) acc = ffma(to_acc(ldg(pc + 1)), t.ne, acc);
                    if (t.mask & 4u) acc = ffma(to_acc(ldg(pc + sH)), t.sw, acc);
                    if (t.mask & 8u) acc = ffma(to_acc(ldg(pc + sH + 1)), t.se, acc);
                    o[c * o_ch] = from_acc<T, float>(acc);
                }
            }
        }
}

template <typename T, int CS>
void launch_cs(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    const bool border = g.padding == PWS_PAD_BORDER, align = g.align != 0;
    {
        dim3 lb((g.Wo + kTW - 1) / kTW, (g.Ho + 15) / 16, g.N);
        if (border && align) { fwd_lean_kernel<T, CS, true, true><<<lb, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
        else if (border) { fwd_lean_kernel<T, CS, true, false><<<lb, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
        else if (align) { fwd_lean_kernel<T, CS, false, true><<<lb, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
        else { fwd_lean_kernel<T, CS, false, false><<<lb, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    }
}

template <typename T>
bool launch_t(const Problem &pb, cudaStream_t st)
{
    if (!rows_vectorizable(pb.in, 16 / (int)sizeof(T), pb.g.W)) return false;
    if (pb.g.C == 3) { launch_cs<T, 3>(pb, st); return true; }
    if (pb.g.C == 1) { launch_cs<T, 1>(pb, st); return true; }
    return false;
}

}  // namespace

// Returns true when the lean kernel took the call.
bool launch_forward_tile(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    if (pb.in.s3 != 1 || pb.out.s3 != 1 || pb.grid_dtype != PWS_F32) return false;
    if (g.N > 65535 || (g.Ho + kTH - 1) / kTH > 65535) return false;
    // interleaved maps are read as float2: needs 8-byte alignment of every pixel's pair
    if (pb.grid.s3 == 1 && ((reinterpret_cast<uintptr_t>(pb.grid.p) & 7) || (pb.grid.sN & 1) || (pb.grid.s1 & 1) || (pb.grid.s2 & 1)))
        return false;
    if (pb.in_dtype == PWS_F32) return launch_t<float>(pb, st);
    if (pb.in_dtype == PWS_F16) return launch_t<__half>(pb, st);
    if (pb.in_dtype == PWS_BF16) return launch_t<__nv_bfloat16>(pb, st);
    return false;
}

}  // namespace pws
