// warp_fwd_march.cu -- forward warp, marching kernel with register reuse of the taps.
//
// The lean / batched forward kernels are bound by L1 data-pipe wavefronts: 4*C unaligned
// tap loads per pixel (profiles/r01_fwd_lean.txt: l1tex wavefronts, issue slots and DRAM
// all sit at ~50 %).  For the maps this path exists for -- identity plus a smooth drift --
// almost all of those loads fetch a value some neighbour already holds:
//   * the east taps (x0+1, .) of lane i are the west taps of lane i+1 when the map
//     advances one source pixel per output pixel  -> one shuffle instead of a load;
//   * the north taps (., y0) of output row r+1 are the south taps of row r when the map
//     advances one source row per output row       -> kept in registers while a warp
//     marches down its strip.
// What is left in the regular case is C loads per pixel (the south-west tap of each
// channel) instead of 4*C; irregular lanes reload what they miss with predicated loads,
// so any map is handled and the values -- hence the results -- are identical to the
// direct kernels' (bit-exact vs ATen).  Tap loads of row r+1 are issued before row r is
// finished (software pipeline), the map is prefetched two rows ahead.
//
// Layout requirements as for the other fast kernels: W-contiguous frames/outputs,
// fp32 map (planar or interleaved), C in {1,3}.  R/main_new.py:106,116,109,118,197,716.
#include "pws_tile.cuh"

namespace pws {

namespace {

constexpr int kRows = 16, kWarpsX = 4, kWarpsY = 2;
constexpr int kThreads = 32 * kWarpsX * kWarpsY;
constexpr int kTW = 32 * kWarpsX, kTH = kRows * kWarpsY;

template <int CS>
struct RowTaps {          // everything one output row of a lane needs
    float ix, iy;
    int x0, y0;
    bool ok;              // lane has a pixel in this row
    bool chain;           // north taps = previous row's south taps
    float sw[CS];         // (x0, y0+1)  always loaded
    float nw[CS], ne[CS], se[CS];  // fresh loads, only meaningful where needed
};

template <typename T, int CS, bool kBorder, bool kAlign>
__global__ void __launch_bounds__(kThreads, 3)
fwd_march_kernel(const View in, const View grid, const View out, const Geometry g)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int n = blockIdx.z;
    const int w = blockIdx.x * kTW + (wrp % kWarpsX) * 32 + lane;
    const int h0 = blockIdx.y * kTH + (wrp / kWarpsX) * kRows;
    if (h0 >= g.Ho) return;  // warp-uniform
    const bool col_ok = w < g.Wo;
    const int rows = min(kRows, g.Ho - h0);
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
    const int Wl = g.W - 1, Hl = g.H - 1;

    const T *__restrict__ ip = (const T *)in.p + (int64_t)n * in.sN;
    const int sH = in.s2;
    const int i_ch = in.s1, o_ch = out.s1;  // 32-bit in-frame offsets
    const bool inter = grid.s3 == 1;
    const float *__restrict__ gq = (const float *)grid.p + (int64_t)n * grid.sN + (int64_t)h0 * grid.s1 + (int64_t)w * grid.s2;
    T *__restrict__ oq = (T *)out.p + (int64_t)n * out.sN + (int64_t)h0 * out.s2 + w;

    auto load_map = [&](const float *q, float &gx, float &gy) {
        gx = -4.f; gy = -4.f;
        if (col_ok) {
            if (inter) { const float2 v = __ldg(reinterpret_cast<const float2 *>(q)); gx = v.x; gy = v.y; }
            else { gx = __ldg(q); gy = __ldg(q + grid.s3); }
        }
    };

    // coordinates of a row + issue of its tap loads (needs the previous row's tap origin)
    auto issue = [&](RowTaps<CS> &t, float gx, float gy, int px0, int py0, bool prev_ok) {
        t.ix = src_index<kBorder, kAlign>(gx, g.W, Wf, Wm1);
        t.iy = src_index<kBorder, kAlign>(gy, g.H, Hf, Hm1);
        t.x0 = (int)floorf(t.ix); t.y0 = (int)floorf(t.iy);
        t.ok = col_ok;
        t.chain = prev_ok && t.x0 == px0 && t.y0 == py0 + 1;
        const int nx0 = __shfl_down_sync(0xffffffffu, t.x0, 1), ny0 = __shfl_down_sync(0xffffffffu, t.y0, 1);
        const bool take_e = lane < 31 && nx0 == t.x0 + 1 && ny0 == t.y0;   // east taps come from lane+1
        const int xc = clampi(t.x0, 0, Wl), xe = clampi(t.x0 + 1, 0, Wl);
        const int on = clampi(t.y0, 0, Hl) * sH, os = clampi(t.y0 + 1, 0, Hl) * sH;
#pragma unroll
        for (int c = 0; c < CS; ++c) t.sw[c] = to_acc(ldg(ip + (c * i_ch + os + xc)));
        if (__any_sync(0xffffffffu, !t.chain)) {
#pragma unroll
            for (int c = 0; c < CS; ++c) if (!t.chain) t.nw[c] = to_acc(ldg(ip + (c * i_ch + on + xc)));
        }
        if (__any_sync(0xffffffffu, !take_e)) {
#pragma unroll
            for (int c = 0; c < CS; ++c) {
                if (!take_e) t.se[c] = to_acc(ldg(ip + (c * i_ch + os + xe)));
                if (!take_e && !t.chain) t.ne[c] = to_acc(ldg(ip + (c * i_ch + on + xe)));
            }
        }
    };

    float gx_a, gy_a, gx_b, gy_b;       // map of the next row / the row after
    load_map(gq, gx_a, gy_a);
    gx_b = -4.f; gy_b = -4.f;
    if (rows > 1) load_map(gq + grid.s1, gx_b, gy_b);
    RowTaps<CS> cur, nxt;
    issue(cur, gx_a, gy_a, 0, 0, false);
    float c_sw[CS], c_se[CS];            // south taps of the previous row (carried)
#pragma unroll
    for (int c = 0; c < CS; ++c) { c_sw[c] = 0.f; c_se[c] = 0.f; }

    for (int r = 0; r < rows; ++r) {
        // ---- stage 1: row r+1 -- coordinates and tap loads; row r+2 -- map
        if (r + 1 < rows) {
            const float gx = gx_b, gy = gy_b;
            if (r + 2 < rows) load_map(gq + (int64_t)(r + 2) * grid.s1, gx_b, gy_b);
            issue(nxt, gx, gy, cur.x0, cur.y0, cur.ok);
        }
        // ---- stage 2: finish row r
        {
            const int nx0 = __shfl_down_sync(0xffffffffu, cur.x0, 1), ny0 = __shfl_down_sync(0xffffffffu, cur.y0, 1);
            const bool take_e = lane < 31 && nx0 == cur.x0 + 1 && ny0 == cur.y0;
            const float x0f = (float)cur.x0, y0f = (float)cur.y0;
            const float wx1 = fsub(cur.ix, x0f), wx0 = fsub(x0f + 1.0f, cur.ix);
            const float wy1 = fsub(cur.iy, y0f), wy0 = fsub(y0f + 1.0f, cur.iy);
            const float wnw = fmul(wx0, wy0), wne = fmul(wx1, wy0), wsw = fmul(wx0, wy1), wse = fmul(wx1, wy1);
            const bool xw = (unsigned)cur.x0 < (unsigned)g.W, xe = (unsigned)(cur.x0 + 1) < (unsigned)g.W;
            const bool yn = (unsigned)cur.y0 < (unsigned)g.H, ys = (unsigned)(cur.y0 + 1) < (unsigned)g.H;
            const bool all4 = xw && xe && yn && ys;
            const bool fast = __all_sync(0xffffffffu, all4 && cur.ok);
            T *__restrict__ o = oq + r * out.s2;
#pragma unroll
            for (int c = 0; c < CS; ++c) {
                const float v_nw = cur.chain ? c_sw[c] : cur.nw[c];
                const float v_sw = cur.sw[c];
                // lane+1's west taps are my east taps when the rows line up
                const float e_n = __shfl_down_sync(0xffffffffu, v_nw, 1);
                const float e_s = __shfl_down_sync(0xffffffffu, v_sw, 1);
                const float v_se = take_e ? e_s : cur.se[c];
                const float v_ne = take_e ? e_n : (cur.chain ? c_se[c] : cur.ne[c]);
                float acc;
                if (fast) {
                    acc = ffma(v_nw, wnw, 0.f);   // fma with +0 keeps ATen's sign of zero
                    acc = ffma(v_ne, wne, acc);
                    acc = ffma(v_sw, wsw, acc);
                    acc = ffma(v_se, wse, acc);
                } else {
                    acc = 0.f;                    // invalid taps are skipped, exactly as ATen skips them
                    if (xw && yn) acc = ffma(v_nw, wnw, acc);
                    if (xe && yn) acc = ffma(v_ne, wne, acc);
                    if (xw && ys) acc = ffma(v_sw, wsw, acc);
                    if (xe && ys) acc = ffma(v_se, wse, acc);
                }
                if (cur.ok) o[c * o_ch] = from_acc<T, float>(acc);
                c_sw[c] = v_sw; c_se[c] = v_se;
            }
        }
        cur = nxt;
    }
}

template <typename T, int CS>
void launch_cs(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    dim3 blocks((g.Wo + kTW - 1) / kTW, (g.Ho + kTH - 1) / kTH, g.N);
    const bool border = g.padding == PWS_PAD_BORDER, align = g.align != 0;
    if (border && align) { fwd_march_kernel<T, CS, true, true><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    else if (border) { fwd_march_kernel<T, CS, true, false><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    else if (align) { fwd_march_kernel<T, CS, false, true><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    else { fwd_march_kernel<T, CS, false, false><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
}

template <typename T>
bool launch_t(const Problem &pb, cudaStream_t st)
{
    if (pb.g.C == 3) { launch_cs<T, 3>(pb, st); return true; }
    if (pb.g.C == 1) { launch_cs<T, 1>(pb, st); return true; }
    return false;
}

}  // namespace

// Returns true when the marching kernel took the call.
bool launch_forward_march(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    if (pb.in.s3 != 1 || pb.out.s3 != 1 || pb.grid_dtype != PWS_F32) return false;
    if (g.N > 65535 || (g.Ho + kTH - 1) / kTH > 65535) return false;
    if (pb.grid.s3 == 1 && ((reinterpret_cast<uintptr_t>(pb.grid.p) & 7) || (pb.grid.sN & 1) || (pb.grid.s1 & 1) || (pb.grid.s2 & 1)))
        return false;
    if (pb.in_dtype == PWS_F32) return launch_t<float>(pb, st);
    if (pb.in_dtype == PWS_F16) return launch_t<__half>(pb, st);
    if (pb.in_dtype == PWS_BF16) return launch_t<__nv_bfloat16>(pb, st);
    return false;
}

}  // namespace pws
