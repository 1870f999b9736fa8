// warp_stages.cu -- K maps applied to the SAME frame in one launch, pre / post scale folded in, forward and backward.
//
// R/main_new.py:103-110 (and :112-119 for the second clip): the three cascade stages of netG each emit a complete map,
// and every map samples the SAME unstable frame:
//     for nl in range(3):  fake = grid_sample((rgb + 1) * 127.5, grid1[nl]) / 127.5 - 1
// The reference runs that as 3 elementwise passes + 3 sampler launches + 6 more elementwise passes per clip, each one a
// launch-bound kernel on a 12.6 MB working set.  Here one launch reads the raw frame, applies (x + a) * b to the taps,
// samples with every map and writes acc * m + e for each: the scaled frame and the unscaled samples never exist.  The
// maps of the cascade are refinements of each other, so the taps of stage k + 1 are the taps of stage k for most pixels
// and stay in registers (the frame is read once, not K times).
// Backward (autograd of the above, R/main_new.py:214): grad_grid of every stage from the stage's grad_output, with the
// chain rule through the post scale (grad * m, as autograd computes it for `x / 127.5`: a multiplication by the rounded
// reciprocal) and the taps pre-scaled as in the forward; optional grad_input (direct REDs: the training sites never ask).
// Arithmetic is statement for statement that of the unfused torch pipeline: (x + a) * b rounded per tap, ATen's sampler
// order (pws_common.cuh), `acc * m` then `+ e` -- results are bit-identical to the reference's sequence of torch calls.
#include "pws_common.cuh"

namespace pws {

namespace {

constexpr int kTileW = 64, kTileH = 16, kThreads = 256;

template <bool kBorder, bool kAlign>
__device__ __forceinline__ float src_index_b(float coord, float size_f, float size_m1_f, float *gm)
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

// forward flavour: border padding clips with min / max, which maps a NaN coordinate to 0 (ATen's clip_coordinates),
// where the backward's comparison chain above leaves it NaN (-> the -100 guard)
template <bool kBorder, bool kAlign>
__device__ __forceinline__ float src_index_f(float coord, float size_f, float size_m1_f)
{
    const float t = __fadd_rn(coord, 1.0f);
    float c = kAlign ? __fmul_rn(__fmul_rn(t, 0.5f), size_m1_f) : __fmul_rn(__fmaf_rn(t, size_f, -1.0f), 0.5f);
    if (kBorder) c = fminf(size_m1_f, fmaxf(c, 0.0f));
    if (!(c <= 2147483648.0f && c >= -2147483648.0f)) c = -100.0f;
    return c;
}

// the four taps of every channel of one source pixel position, pre-scaled; reloaded only when the position changes
template <int CS>
struct TapCache {
    int x0, y0;
    float v[CS][4];
};

template <int CS>
__device__ __forceinline__ void load_taps(TapCache<CS> &tc, const float *__restrict__ ip, const View &in, const Taps<float> &t,
                                          const StageScale &sc)
{
    if (tc.x0 == t.x0 && tc.y0 == t.y0) return;
    tc.x0 = t.x0; tc.y0 = t.y0;
    const int o_nw = t.y0 * in.s2 + t.x0 * in.s3;
#pragma unroll
    for (int c = 0; c < CS; ++c) {
        const float *__restrict__ pc = ip + c * in.s1 + o_nw;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (t.mask & 1u) v0 = __ldg(pc);
        if (t.mask & 2u) v1 = __ldg(pc + in.s3);
        if (t.mask & 4u) v2 = __ldg(pc + in.s2);
        if (t.mask & 8u) v3 = __ldg(pc + in.s2 + in.s3);
        if (sc.has_pre) {   // (x + a) * b, elementwise as the reference does before sampling
            v0 = __fmul_rn(__fadd_rn(v0, sc.pre_add), sc.pre_mul); v1 = __fmul_rn(__fadd_rn(v1, sc.pre_add), sc.pre_mul);
            v2 = __fmul_rn(__fadd_rn(v2, sc.pre_add), sc.pre_mul); v3 = __fmul_rn(__fadd_rn(v3, sc.pre_add), sc.pre_mul);
        }
        tc.v[c][0] = v0; tc.v[c][1] = v1; tc.v[c][2] = v2; tc.v[c][3] = v3;
    }
}

template <int CS, bool kBorder, bool kAlign>
__global__ void __launch_bounds__(kThreads)
stages_fwd_kernel(const View in, const StageViews sv, const int K, const Geometry g, const StageScale sc)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, n = blockIdx.z;
    const float *__restrict__ ip = (const float *)in.p + (int64_t)n * in.sN;
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int w = blockIdx.x * kTileW + lane + 32 * i, h = blockIdx.y * kTileH + wrp + 8 * j;
            if (!(w < g.Wo && h < g.Ho)) continue;
            TapCache<CS> tc;
            tc.x0 = INT_MIN; tc.y0 = INT_MIN;
            for (int k = 0; k < K; ++k) {
                const View &mv = sv.map[k];
                const float *mp = (const float *)mv.p + (int64_t)n * mv.sN + h * mv.s1 + w * mv.s2;
                const float ix = src_index_f<kBorder, kAlign>(__ldg(mp), Wf, Wm1);
                const float iy = src_index_f<kBorder, kAlign>(__ldg(mp + mv.s3), Hf, Hm1);
                Taps<float> t;
                make_taps(ix, iy, g.H, g.W, t);
                // the cache key is the tap position; the mask is a function of it
                load_taps<CS>(tc, ip, in, t, sc);
                const View &ov = sv.io[k];
                float *op = (float *)ov.p + (int64_t)n * ov.sN + h * ov.s2 + w * ov.s3;
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    float acc = 0.f;
                    if (t.mask & 1u) acc = ffma(tc.v[c][0], t.nw, acc);
                    if (t.mask & 2u) acc = ffma(tc.v[c][1], t.ne, acc);
                    if (t.mask & 4u) acc = ffma(tc.v[c][2], t.sw, acc);
                    if (t.mask & 8u) acc = ffma(tc.v[c][3], t.se, acc);
                    if (sc.has_post) acc = __fadd_rn(__fmul_rn(acc, sc.post_mul), sc.post_add);
                    op[c * ov.s1] = acc;
                }
            }
        }
}

template <int CS, bool kBorder, bool kAlign, bool kGin>
__global__ void __launch_bounds__(kThreads)
stages_bwd_kernel(const View in, const StageViews sv, const int K, const View gin, const Geometry g, const StageScale sc)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, n = blockIdx.z;
    const float *__restrict__ ip = (const float *)in.p + (int64_t)n * in.sN;
    float *__restrict__ gip = kGin ? (float *)gin.p + (int64_t)n * gin.sN : nullptr;
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int w = blockIdx.x * kTileW + lane + 32 * i, h = blockIdx.y * kTileH + wrp + 8 * j;
            if (!(w < g.Wo && h < g.Ho)) continue;
            TapCache<CS> tc;
            tc.x0 = INT_MIN; tc.y0 = INT_MIN;
            for (int k = 0; k < K; ++k) {
                const View &mv = sv.map[k];
                const float *mp = (const float *)mv.p + (int64_t)n * mv.sN + h * mv.s1 + w * mv.s2;
                float gxm, gym;
                const float ix = src_index_b<kBorder, kAlign>(__ldg(mp), Wf, Wm1, &gxm);
                const float iy = src_index_b<kBorder, kAlign>(__ldg(mp + mv.s3), Hf, Hm1, &gym);
                Taps<float> t;
                make_taps(ix, iy, g.H, g.W, t);
                load_taps<CS>(tc, ip, in, t, sc);
                const float x0f = (float)t.x0, y0f = (float)t.y0;
                const float dw = fsub(x0f + 1.0f, ix), de = fsub(ix, x0f), dn = fsub(y0f + 1.0f, iy), ds = fsub(iy, y0f);
                const View &ov = sv.io[k];
                const float *gop = (const float *)ov.p + (int64_t)n * ov.sN + h * ov.s2 + w * ov.s3;
                float gix = 0.f, giy = 0.f;
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    float go = __ldg(gop + c * ov.s1);
                    if (sc.has_post) go = __fmul_rn(go, sc.post_mul);   // autograd of `x / d`: grad * (1/d)
                    const float v0 = tc.v[c][0], v1 = tc.v[c][1], v2 = tc.v[c][2], v3 = tc.v[c][3];
                    // ATen's statement order: t = v*d rounded, then one fma with gOut
                    if (t.mask & 1u) { gix = ffma(-fmul(v0, dn), go, gix); giy = ffma(-fmul(v0, dw), go, giy); }
                    if (t.mask & 2u) { gix = ffma(fmul(v1, dn), go, gix);  giy = ffma(-fmul(v1, de), go, giy); }
                    if (t.mask & 4u) { gix = ffma(-fmul(v2, ds), go, gix); giy = ffma(fmul(v2, dw), go, giy); }
                    if (t.mask & 8u) { gix = ffma(fmul(v3, ds), go, gix);  giy = ffma(fmul(v3, de), go, giy); }
                    if (kGin) {
                        // d out / d frame = w_tap, through the pre scale (x + a) * b: times b
                        const float gs = sc.has_pre ? __fmul_rn(go, sc.pre_mul) : go;
                        float *pc = gip + c * gin.s1 + t.y0 * gin.s2 + t.x0;
                        if (t.mask & 1u) atomicAdd(pc, fmul(t.nw, gs));
                        if (t.mask & 2u) atomicAdd(pc + 1, fmul(t.ne, gs));
                        if (t.mask & 4u) atomicAdd(pc + gin.s2, fmul(t.sw, gs));
                        if (t.mask & 8u) atomicAdd(pc + gin.s2 + 1, fmul(t.se, gs));
                    }
                }
                const View &gv = sv.gg[k];
                if (gv.p) {
                    float *ggp = (float *)gv.p + (int64_t)n * gv.sN + h * gv.s1 + w * gv.s2;
                    ggp[0] = fmul(gxm, gix);
                    ggp[gv.s3] = fmul(gym, giy);
                }
            }
        }
}

}  // namespace

int launch_stages_forward(const View &in, const StageViews &sv, int K, const Geometry &g, const StageScale &sc, cudaStream_t st)
{
    if (g.N > 65535 || (g.Ho + kTileH - 1) / kTileH > 65535) { set_error("stages forward: batch or height too large"); return PWS_EUNSUPPORTED; }
    dim3 blocks((g.Wo + kTileW - 1) / kTileW, (g.Ho + kTileH - 1) / kTileH, g.N);
    const bool border = g.padding == PWS_PAD_BORDER, align = g.align != 0;
#define PWS_F(CS) do { \
        if (border && align) stages_fwd_kernel<CS, true, true><<<blocks, kThreads, 0, st>>>(in, sv, K, g, sc); \
        else if (border) stages_fwd_kernel<CS, true, false><<<blocks, kThreads, 0, st>>>(in, sv, K, g, sc); \
        else if (align) stages_fwd_kernel<CS, false, true><<<blocks, kThreads, 0, st>>>(in, sv, K, g, sc); \
        else stages_fwd_kernel<CS, false, false><<<blocks, kThreads, 0, st>>>(in, sv, K, g, sc); } while (0)
    if (g.C == 3) PWS_F(3); else if (g.C == 1) PWS_F(1);
    else { set_error("stages forward: C must be 1 or 3 (got %d)", g.C); return PWS_EUNSUPPORTED; }
#undef PWS_F
    note_launch();
    note_kernel("stages_fwd");
    return PWS_OK;
}

int launch_stages_backward(const View &in, const StageViews &sv, int K, const View &gin, bool want_gin, const Geometry &g,
                           const StageScale &sc, cudaStream_t st)
{
    if (g.N > 65535 || (g.Ho + kTileH - 1) / kTileH > 65535) { set_error("stages backward: batch or height too large"); return PWS_EUNSUPPORTED; }
    dim3 blocks((g.Wo + kTileW - 1) / kTileW, (g.Ho + kTileH - 1) / kTileH, g.N);
    const bool border = g.padding == PWS_PAD_BORDER, align = g.align != 0;
    if (want_gin) {
        cudaError_t e = cudaMemsetAsync(gin.p, 0, (size_t)g.N * g.C * g.H * g.W * sizeof(float), st);
        if (e != cudaSuccess) { set_error("stages backward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return PWS_ECUDA; }
    }
#define PWS_B(CS, GIN) do { \
        if (border && align) stages_bwd_kernel<CS, true, true, GIN><<<blocks, kThreads, 0, st>>>(in, sv, K, gin, g, sc); \
        else if (border) stages_bwd_kernel<CS, true, false, GIN><<<blocks, kThreads, 0, st>>>(in, sv, K, gin, g, sc); \
        else if (align) stages_bwd_kernel<CS, false, true, GIN><<<blocks, kThreads, 0, st>>>(in, sv, K, gin, g, sc); \
        else stages_bwd_kernel<CS, false, false, GIN><<<blocks, kThreads, 0, st>>>(in, sv, K, gin, g, sc); } while (0)
    if (g.C == 3) { if (want_gin) PWS_B(3, true); else PWS_B(3, false); }
    else if (g.C == 1) { if (want_gin) PWS_B(1, true); else PWS_B(1, false); }
    else { set_error("stages backward: C must be 1 or 3 (got %d)", g.C); return PWS_EUNSUPPORTED; }
#undef PWS_B
    note_launch();
    note_kernel("stages_bwd");
    return PWS_OK;
}

}  // namespace pws
