// warp_fwd.cu -- forward pixel-wise warp, direct-gather kernel + taps debug kernel.
//
// Replaces launch_grid_sampler_2d_forward_kernel ($TORCH/include/ATen/native/cuda/GridSampler.h:12-14)
// for the reference call sites R/main_new.py:106,109,116,118,197,716.
//
// Direct kernel: any strides, any C, frames f32/f16/bf16/f64.  A CTA owns a
// 64x16 output tile; lane <-> x so map reads, frame gathers (near-identity
// maps) and output writes are all 128-byte coalesced per warp instruction, and
// the 2-D tile keeps the y0 / y0+1 source rows of neighbouring output rows in
// the same SM's L1.  Each thread issues the map loads of its 4 pixels before
// any dependent gather (8 independent loads in flight per thread).
#include "pws_common.cuh"

#include <cstdlib>

namespace pws {

namespace {

constexpr int kTileW = 64, kTileH = 16, kPX = 2, kPY = 2, kThreads = 256;

template <typename T, typename G, int CS>
__global__ void __launch_bounds__(kThreads)
fwd_direct_kernel(const View in, const View grid, const View out, const Geometry g,
                  const int tiles_x, const int tiles_y)
{
    using A = typename Acc<T>::type;
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x;
    const int rest = tile / tiles_x;
    const int ty = rest % tiles_y;
    const int n = rest / tiles_y;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;

    const G *__restrict__ gp = (const G *)grid.p + (int64_t)n * grid.sN;
    const T *__restrict__ ip = (const T *)in.p + (int64_t)n * in.sN;
    T *__restrict__ op = (T *)out.p + (int64_t)n * out.sN;
    const bool align = g.align != 0;
    const int C = CS > 0 ? CS : g.C;

    A gx[kPY][kPX], gy[kPY][kPX];
    bool ok[kPY][kPX];
#pragma unroll
    for (int j = 0; j < kPY; ++j)
#pragma unroll
        for (int i = 0; i < kPX; ++i) {
            const int w = tx * kTileW + lane + 32 * i;
            const int h = ty * kTileH + wrp + 8 * j;
            ok[j][i] = (w < g.Wo) && (h < g.Ho);
            gx[j][i] = (A)0; gy[j][i] = (A)0;
            if (ok[j][i]) {
                const int off = h * grid.s1 + w * grid.s2;
                gx[j][i] = to_acc(ldg(gp + off));
                gy[j][i] = to_acc(ldg(gp + off + grid.s3));
            }
        }

#pragma unroll
    for (int j = 0; j < kPY; ++j)
#pragma unroll
        for (int i = 0; i < kPX; ++i) {
            if (!ok[j][i]) continue;
            const int w = tx * kTileW + lane + 32 * i;
            const int h = ty * kTileH + wrp + 8 * j;
            Taps<A> t;
            make_taps(source_index(gx[j][i], g.W, g.padding, align),
                      source_index(gy[j][i], g.H, g.padding, align), g.H, g.W, t);
            const int o_nw = t.y0 * in.s2 + t.x0 * in.s3;
            const int o_out = h * out.s2 + w * out.s3;
#pragma unroll 3
            for (int c = 0; c < C; ++c) {
                const T *__restrict__ pc = ip + c * in.s1;
                // loads first (independent), then the fma chain in ATen's order nw, ne, sw, se
                A v0 = (A)0, v1 = (A)0, v2 = (A)0, v3 = (A)0;
                if (t.mask & 1u) v0 = to_acc(ldg(pc + o_nw));
                if (t.mask & 2u) v1 = to_acc(ldg(pc + o_nw + in.s3));
                if (t.mask & 4u) v2 = to_acc(ldg(pc + o_nw + in.s2));
                if (t.mask & 8u) v3 = to_acc(ldg(pc + o_nw + in.s2 + in.s3));
                A acc = (A)0;
                if (t.mask & 1u) acc = ffma(v0, t.nw, acc);
                if (t.mask & 2u) acc = ffma(v1, t.ne, acc);
                if (t.mask & 4u) acc = ffma(v2, t.sw, acc);
                if (t.mask & 8u) acc = ffma(v3, t.se, acc);
                op[o_out + c * out.s1] = from_acc<T, A>(acc);
            }
        }
}

__global__ void __launch_bounds__(256)
taps_kernel(const View grid, const Geometry g, int32_t *__restrict__ x0, int32_t *__restrict__ y0,
            uint8_t *__restrict__ mask, float *__restrict__ wts, const int64_t total)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int w = (int)(idx % g.Wo);
    const int64_t r = idx / g.Wo;
    const int h = (int)(r % g.Ho);
    const int n = (int)(r / g.Ho);
    const float *gp = (const float *)grid.p + (int64_t)n * grid.sN + h * grid.s1 + w * grid.s2;
    Taps<float> t;
    make_taps(source_index(__ldg(gp), g.W, g.padding, g.align != 0),
              source_index(__ldg(gp + grid.s3), g.H, g.padding, g.align != 0), g.H, g.W, t);
    x0[idx] = t.x0; y0[idx] = t.y0; mask[idx] = (uint8_t)t.mask;
    if (wts) reinterpret_cast<float4 *>(wts)[idx] = make_float4(t.nw, t.ne, t.sw, t.se);
}

template <typename T, typename G>
int launch_direct(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    const int tiles_x = (g.Wo + kTileW - 1) / kTileW, tiles_y = (g.Ho + kTileH - 1) / kTileH;
    const int64_t tiles = (int64_t)tiles_x * tiles_y * g.N;
    if (tiles > INT_MAX) { set_error("forward: too many tiles (%lld)", (long long)tiles); return PWS_EUNSUPPORTED; }
    dim3 grid_dim((unsigned)tiles), block(kThreads);
    if (g.C == 3) { fwd_direct_kernel<T, G, 3><<<grid_dim, block, 0, st>>>(pb.in, pb.grid, pb.out, g, tiles_x, tiles_y); note_launch(); }
    else if (g.C == 1) { fwd_direct_kernel<T, G, 1><<<grid_dim, block, 0, st>>>(pb.in, pb.grid, pb.out, g, tiles_x, tiles_y); note_launch(); }
    else { fwd_direct_kernel<T, G, 0><<<grid_dim, block, 0, st>>>(pb.in, pb.grid, pb.out, g, tiles_x, tiles_y); note_launch(); }
    return PWS_OK;
}

}  // namespace

bool launch_forward_tma(const Problem &pb, cudaStream_t st);               // warp_fwd_tma.cu (TMA-pipelined persistent)
bool launch_forward_tile(const Problem &pb, cudaStream_t st);              // warp_fwd_tile.cu (lean direct gather)

// Development builds (-DPWS_DEV_HOOKS, tools/build_variant.py) can force the fallback kernels through the environment;
// the product library has no run-time switches.
static bool force_direct()
{
#ifdef PWS_DEV_HOOKS
    static const bool v = [] { const char *e = std::getenv("PWS_FORCE_DIRECT"); return e && e[0] == '1'; }();
    return v;
#else
    return false;
#endif
}

int launch_forward(const Problem &pb, cudaStream_t st)
{
    if (!force_direct() && small_problem(pb) && launch_forward_tile(pb, st)) { note_kernel("fwd_lean"); return PWS_OK; }
    if (!force_direct() && launch_forward_tma(pb, st)) return PWS_OK;
    if (!force_direct() && launch_forward_tile(pb, st)) { note_kernel("fwd_lean"); return PWS_OK; }
    note_kernel("fwd_direct");
    const int it = pb.in_dtype, gt = pb.grid_dtype;
    if (it == PWS_F32 && gt == PWS_F32) return launch_direct<float, float>(pb, st);
    if (it == PWS_F64 && gt == PWS_F64) return launch_direct<double, double>(pb, st);
    if (it == PWS_F16 && gt == PWS_F32) return launch_direct<__half, float>(pb, st);
    if (it == PWS_F16 && gt == PWS_F16) return launch_direct<__half, __half>(pb, st);
    if (it == PWS_BF16 && gt == PWS_F32) return launch_direct<__nv_bfloat16, float>(pb, st);
    if (it == PWS_BF16 && gt == PWS_BF16) return launch_direct<__nv_bfloat16, __nv_bfloat16>(pb, st);
    set_error("forward: unsupported dtype pair (frame %d, map %d)", it, gt);
    return PWS_EUNSUPPORTED;
}

int launch_taps(const View &grid, const Geometry &g, int32_t *x0, int32_t *y0, uint8_t *mask, float *w, cudaStream_t st)
{
    const int64_t total = (int64_t)g.N * g.Ho * g.Wo;
    if (total == 0) return PWS_OK;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > INT_MAX) { set_error("taps: too many pixels"); return PWS_EUNSUPPORTED; }
    { taps_kernel<<<(unsigned)blocks, 256, 0, st>>>(grid, g, x0, y0, mask, w, total); note_launch(); }
    return PWS_OK;
}

}  // namespace pws
