// warp_fwd_batch.cu -- forward warp, batched-gather kernel (the default forward path).
//
// Same layout specialisations as fwd_lean_kernel (W-contiguous frames and output, fp32 map,
// C in {1,3}, padding / align_corners as template flags), restructured around what the first
// profiles showed: the lean kernel is bound by DRAM latency, four dependent rounds of 12 tap
// loads per thread.  Here a thread owns 8 pixels (64x32 tile per CTA); it issues all 16 map
// loads first, then for a batch of B pixels computes every tap address, issues all 4*C*B tap
// loads back to back (unconditional: tap coordinates are clamped into the frame, validity is
// applied as a select afterwards) and only then runs the fma chains and the stores.  That puts
// 2-4x more bytes in flight per SM at equal or lower occupancy.
//
// Call sites served: R/main_new.py:106,116,109,118 (planar maps), :197 (interleaved affine_grid
// maps), :716 when the frame is W-contiguous.
#include "pws_tile.cuh"

namespace pws {

namespace {

constexpr int kTW = 64, kTH = 32, kThreads = 256, kWarps = 8;
constexpr int kPX = kTW / 32, kPY = kTH / kWarps;  // 2 x 4 pixels per thread

template <typename T, int CS, bool kBorder, bool kAlign, int B>
__global__ void __launch_bounds__(kThreads, (B >= 4) ? 2 : 3)
fwd_batch_kernel(const View in, const View grid, const View out, const Geometry g)
{
    static_assert(B == 2 || B == 4, "batch of 2 (one row) or 4 (two rows) pixels");
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int n = blockIdx.z;
    const int w_base = blockIdx.x * kTW + lane, h_base = blockIdx.y * kTH + wrp;
    const float *__restrict__ gp = (const float *)grid.p + (int64_t)n * grid.sN;
    const T *__restrict__ ip = (const T *)in.p + (int64_t)n * in.sN;
    T *__restrict__ op = (T *)out.p + (int64_t)n * out.sN;
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
    const int Wl = g.W - 1, Hl = g.H - 1;
    const bool full = (blockIdx.x + 1) * kTW <= g.Wo && (blockIdx.y + 1) * kTH <= g.Ho;  // CTA-uniform

    // ---- all map loads of the thread, back to back
    float sx[kPY][kPX], sy[kPY][kPX];
    {
        const bool inter = grid.s3 == 1;
        const float *__restrict__ q = gp + (int64_t)h_base * grid.s1 + (int64_t)w_base * grid.s2;
        const int64_t row = (int64_t)kWarps * grid.s1;
        const int col = 32 * grid.s2;
#pragma unroll
        for (int j = 0; j < kPY; ++j)
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

    const int sH = in.s2;
    const int i_ch = in.s1, o_ch = out.s1, o_row = kWarps * out.s2;  // 32-bit in-frame offsets
    T *__restrict__ o0 = op + (int64_t)h_base * out.s2 + w_base;
    constexpr int JB = B / kPX;  // rows per batch

#pragma unroll
    for (int jb = 0; jb < kPY; jb += JB) {
        float wgt[B][4];
        int off[B][2];     // offsets of the two tap rows (west tap), clamped into the frame
        int dxe[B];        // 0/1: step to the east tap (0 when clamped onto the same column)
        unsigned msk[B];
        bool allin = true;
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int j = jb + b / kPX, i = b % kPX;
            const bool ok = full || (w_base + 32 * i < g.Wo && h_base + kWarps * j < g.Ho);
            const float ix = src_index<kBorder, kAlign>(sx[j][i], g.W, Wf, Wm1);
            const float iy = src_index<kBorder, kAlign>(sy[j][i], g.H, Hf, Hm1);
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float wx1 = fsub(ix, x0f), wx0 = fsub(x0f + 1.0f, ix);
            const float wy1 = fsub(iy, y0f), wy0 = fsub(y0f + 1.0f, iy);
            wgt[b][0] = fmul(wx0, wy0); wgt[b][1] = fmul(wx1, wy0);
            wgt[b][2] = fmul(wx0, wy1); wgt[b][3] = fmul(wx1, wy1);
            const int x0 = (int)x0f, y0 = (int)y0f;
            const bool xw = (unsigned)x0 < (unsigned)g.W, xe = (unsigned)(x0 + 1) < (unsigned)g.W;
            const bool yn = (unsigned)y0 < (unsigned)g.H, ys = (unsigned)(y0 + 1) < (unsigned)g.H;
            msk[b] = ok ? (((xw && yn) ? 1u : 0u) | ((xe && yn) ? 2u : 0u) | ((xw && ys) ? 4u : 0u) | ((xe && ys) ? 8u : 0u)) : 0u;
            allin = allin && (msk[b] == 15u);
            const int xc = clampi(x0, 0, Wl), yc = clampi(y0, 0, Hl);
            dxe[b] = clampi(x0 + 1, 0, Wl) - xc;
            off[b][0] = yc * sH + xc;
            off[b][1] = clampi(y0 + 1, 0, Hl) * sH + xc;
        }
        // every tap load of the batch, unconditional and independent
        float v[B][CS][4];
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int c = 0; c < CS; ++c) {
                v[b][c][0] = to_acc(ldg(ip + (c * i_ch + off[b][0])));
                v[b][c][1] = to_acc(ldg(ip + (c * i_ch + off[b][0] + dxe[b])));
                v[b][c][2] = to_acc(ldg(ip + (c * i_ch + off[b][1])));
                v[b][c][3] = to_acc(ldg(ip + (c * i_ch + off[b][1] + dxe[b])));
            }
        if (__all_sync(0xffffffffu, allin)) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                T *__restrict__ o = o0 + ((jb + b / kPX) * o_row + 32 * (b % kPX));
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    float acc = ffma(v[b][c][0], wgt[b][0], 0.f);  // fma with +0 keeps ATen's sign of zero
                    acc = ffma(v[b][c][1], wgt[b][1], acc);
                    acc = ffma(v[b][c][2], wgt[b][2], acc);
                    acc = ffma(v[b][c][3], wgt[b][3], acc);
                    o[c * o_ch] = from_acc<T, float>(acc);
                }
            }
        } else {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int j = jb + b / kPX, i = b % kPX;
                if (!(full || (w_base + 32 * i < g.Wo && h_base + kWarps * j < g.Ho))) continue;
                T *__restrict__ o = o0 + (j * o_row + 32 * i);
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    float acc = 0.f;  // invalid taps are skipped, exactly as ATen skips them
                    if (msk[b] & 1u) acc = ffma(v[b][c][0], wgt[b][0], acc);
                    if (msk[b] & 2u) acc = ffma(v[b][c][1], wgt[b][1], acc);
                    if (msk[b] & 4u) acc = ffma(v[b][c][2], wgt[b][2], acc);
                    if (msk[b] & 8u) acc = ffma(v[b][c][3], wgt[b][3], acc);
                    o[c * o_ch] = from_acc<T, float>(acc);
                }
            }
        }
    }
}

template <typename T, int CS, int B>
void launch_b(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    dim3 blocks((g.Wo + kTW - 1) / kTW, (g.Ho + kTH - 1) / kTH, g.N);
    const bool border = g.padding == PWS_PAD_BORDER, align = g.align != 0;
    if (border && align) { fwd_batch_kernel<T, CS, true, true, B><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    else if (border) { fwd_batch_kernel<T, CS, true, false, B><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    else if (align) { fwd_batch_kernel<T, CS, false, true, B><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
    else { fwd_batch_kernel<T, CS, false, false, B><<<blocks, kThreads, 0, st>>>(pb.in, pb.grid, pb.out, g); note_launch(); }
}

template <typename T>
bool launch_t(const Problem &pb, int batch, cudaStream_t st)
{
    if (pb.g.C == 3) { if (batch == 4) launch_b<T, 3, 4>(pb, st); else launch_b<T, 3, 2>(pb, st); return true; }
    if (pb.g.C == 1) { if (batch == 4) launch_b<T, 1, 4>(pb, st); else launch_b<T, 1, 2>(pb, st); return true; }
    return false;
}

}  // namespace

// Returns true when the batched kernel took the call.  batch: 2 or 4 pixels per gather round.
bool launch_forward_batch(const Problem &pb, int batch, cudaStream_t st)
{
    const Geometry &g = pb.g;
    if (pb.in.s3 != 1 || pb.out.s3 != 1 || pb.grid_dtype != PWS_F32) return false;
    if (g.N > 65535 || (g.Ho + kTH - 1) / kTH > 65535) return false;
    if (pb.grid.s3 == 1 && ((reinterpret_cast<uintptr_t>(pb.grid.p) & 7) || (pb.grid.sN & 1) || (pb.grid.s1 & 1) || (pb.grid.s2 & 1)))
        return false;
    if (pb.in_dtype == PWS_F32) return launch_t<float>(pb, batch, st);
    return false;
}

}  // namespace pws
