// warp_fused.cu -- forward warp with the map composition fused into the sample.
//
// In the reference the map handed to grid_sample is itself the product of a few
// elementwise / resampling passes over HBM:
//   a8  map = tanh-drift.permute(0,2,3,1) + affine_grid(theta)     R/lib/networks_cascading.py:164,235
//   a9  map = drift + identity meshgrid (generate_maps)             R/lib/utils.py:386-403
//   a10 map = bilinear upsample of the 256x256 map to the frame     R/main_new.py:706-710, R/main.py:639-641
//   a11 frame = (x+1)*127.5 before, out/127.5-1 after               R/main_new.py:106-107
//   a4  frame arrives as uint8 HWC from cv2 and leaves as uint8     R/main_new.py:679-684,717-721
// Here all of that happens in registers of the sampling kernel: the composed and
// upsampled map, the scaled frame and the fp32 output of the unfused pipeline are
// never written to HBM.  pws_compose_map emits the map the kernel uses, so tests can
// hold it to the torch-composed map and hold the fused sample to the plain sample of
// that very map (bit-exact).
//
// Generic strides (uint8 HWC frames and planar low-resolution drifts are the norm here),
// taps through L1, 4 pixels per thread as in the direct kernel.
#include "pws_common.cuh"

namespace pws {

namespace {

__device__ __forceinline__ float lattice_base(int i, int S, int base, int align)
{
    if (base == PWS_BASE_IDENTITY)  // generate_maps: X*2/(S-1)-1 on float tensors
        return __fsub_rn(__fdiv_rn(__fmul_rn((float)i, 2.0f), (float)(S - 1)), 1.0f);
    // affine_grid's linspace(-1,1,S) (times (S-1)/S when not align_corners)
    if (S <= 1) return 0.0f;
    const float step = __fdiv_rn(2.0f, (float)(S - 1));
    float v = (i < S / 2) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(S - 1 - i), 1.0f);  // as nvcc contracts ATen's linspace
    if (!align) v = __fdiv_rn(__fmul_rn(v, (float)(S - 1)), (float)S);
    return v;
}

// map value at lattice point (i, j) of item n:  drift + base
__device__ __forceinline__ void lattice_map(const MapSpec &m, int n, int i, int j, float &mx, float &my)
{
    float dx = 0.f, dy = 0.f;
    if (m.drift.p) {
        const float *d = (const float *)m.drift.p + (int64_t)n * m.drift.sN + i * m.drift.s1 + j * m.drift.s2;
        // interleaved (x,y) pairs -- the lattice the host side composes -- come as one 8-byte load
        if (m.drift.s3 == 1 && !((reinterpret_cast<uintptr_t>(m.drift.p) & 7) | ((m.drift.sN | m.drift.s1 | m.drift.s2) & 1))) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(d));
            dx = v.x; dy = v.y;
        } else {
            dx = __ldg(d); dy = __ldg(d + m.drift.s3);
        }
    }
    if (m.base == PWS_BASE_NONE) { mx = dx; my = dy; return; }
    const float bx = lattice_base(j, m.mw, m.base, m.base_align), by = lattice_base(i, m.mh, m.base, m.base_align);
    if (m.base == PWS_BASE_IDENTITY) { mx = __fadd_rn(dx, bx); my = __fadd_rn(dy, by); return; }
    const float *t = m.theta + n * 6;
    const float ax = __fmaf_rn(bx, __ldg(t + 0), __fmaf_rn(by, __ldg(t + 1), __ldg(t + 2)));
    const float ay = __fmaf_rn(bx, __ldg(t + 3), __fmaf_rn(by, __ldg(t + 4), __ldg(t + 5)));
    mx = __fadd_rn(dx, ax); my = __fadd_rn(dy, ay);
}

struct UpCoef { int i0, ip; float l0, l1; };

// upsample_bilinear2d source index (UpSample.cuh:96-130)
__device__ __forceinline__ UpCoef up_coef(int dst, int src_size, float scale, bool align)
{
    float s = align ? __fmul_rn(scale, (float)dst) : __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
    if (!align && s < 0.f) s = 0.f;
    UpCoef c;
    c.i0 = (int)s;
    c.ip = (c.i0 < src_size - 1) ? 1 : 0;
    c.l1 = __fsub_rn(s, (float)c.i0);
    c.l0 = __fsub_rn(1.0f, c.l1);
    return c;
}

// the upsampled map from the row / column coefficients of an output pixel
__device__ __forceinline__ void map_up(const MapSpec &m, int n, const UpCoef &cy, const UpCoef &cx, float &gx, float &gy)
{
    float x00, y00, x01, y01, x10, y10, x11, y11;
    lattice_map(m, n, cy.i0, cx.i0, x00, y00);
    lattice_map(m, n, cy.i0, cx.i0 + cx.ip, x01, y01);
    lattice_map(m, n, cy.i0 + cy.ip, cx.i0, x10, y10);
    lattice_map(m, n, cy.i0 + cy.ip, cx.i0 + cx.ip, x11, y11);
    // h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11), contracted as nvcc contracts ATen's expression
    gx = __fmaf_rn(cy.l0, __fmaf_rn(cx.l0, x00, __fmul_rn(cx.l1, x01)), __fmul_rn(cy.l1, __fmaf_rn(cx.l0, x10, __fmul_rn(cx.l1, x11))));
    gy = __fmaf_rn(cy.l0, __fmaf_rn(cx.l0, y00, __fmul_rn(cx.l1, y01)), __fmul_rn(cy.l1, __fmaf_rn(cx.l0, y10, __fmul_rn(cx.l1, y11))));
}

// the map the sampler sees at output pixel (h, w)
__device__ __forceinline__ void map_at(const MapSpec &m, int n, int h, int w, float rh, float rw, float &gx, float &gy)
{
    if (m.upsample == PWS_UP_NONE) { lattice_map(m, n, h, w, gx, gy); return; }
    const bool al = m.upsample == PWS_UP_ALIGNED;
    map_up(m, n, up_coef(h, m.mh, rh, al), up_coef(w, m.mw, rw, al), gx, gy);
}

__device__ __forceinline__ void up_scales(const MapSpec &m, int Ho, int Wo, float &rh, float &rw)
{
    rh = rw = 0.f;
    if (m.upsample == PWS_UP_ALIGNED) {
        rh = Ho > 1 ? __fdiv_rn((float)(m.mh - 1), (float)(Ho - 1)) : 0.f;
        rw = Wo > 1 ? __fdiv_rn((float)(m.mw - 1), (float)(Wo - 1)) : 0.f;
    } else if (m.upsample == PWS_UP_HALF_PIXEL) {
        rh = __fdiv_rn((float)m.mh, (float)Ho);
        rw = __fdiv_rn((float)m.mw, (float)Wo);
    }
}

// host twin (float division is correctly rounded on both sides)
inline void up_scales_host(const MapSpec &m, int Ho, int Wo, float &rh, float &rw)
{
    rh = rw = 0.f;
    if (m.upsample == PWS_UP_ALIGNED) {
        rh = Ho > 1 ? (float)(m.mh - 1) / (float)(Ho - 1) : 0.f;
        rw = Wo > 1 ? (float)(m.mw - 1) / (float)(Wo - 1) : 0.f;
    } else if (m.upsample == PWS_UP_HALF_PIXEL) {
        rh = (float)m.mh / (float)Ho;
        rw = (float)m.mw / (float)Wo;
    }
}

template <typename T> __device__ __forceinline__ float load_px(const T *p) { return to_acc(ldg(p)); }
template <> __device__ __forceinline__ float load_px<uint8_t>(const uint8_t *p) { return (float)__ldg(p); }
template <typename T> __device__ __forceinline__ void store_px(T *p, float v) { *p = from_acc<T, float>(v); }
// uint8 egress: truncation, what `.astype(np.uint8)` does to the in-range values a warp of 0..255 produces
template <> __device__ __forceinline__ void store_px<uint8_t>(uint8_t *p, float v) { *p = (uint8_t)(int)fminf(fmaxf(v, 0.f), 255.f); }

constexpr int kTileW = 64, kTileH = 16, kThreads = 256;

template <typename TI, typename TO, int CS>
__global__ void __launch_bounds__(kThreads)
fwd_fused_kernel(const View in, const MapSpec m, const View out, const Geometry g)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int n = blockIdx.z;
    const TI *__restrict__ ip = (const TI *)in.p + (int64_t)n * in.sN;
    TO *__restrict__ op = (TO *)out.p + (int64_t)n * out.sN;
    const bool align = g.align != 0;
    const int C = CS > 0 ? CS : g.C;
    float rh, rw;
    up_scales(m, g.Ho, g.Wo, rh, rw);

    float gx[2][2], gy[2][2];
    if (m.upsample == PWS_UP_NONE) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int w = blockIdx.x * kTileW + lane + 32 * i, h = blockIdx.y * kTileH + wrp + 8 * j;
                gx[j][i] = 0.f; gy[j][i] = 0.f;
                if (w < g.Wo && h < g.Ho) lattice_map(m, n, h, w, gx[j][i], gy[j][i]);
            }
    } else {
        // a thread's four pixels share two rows and two columns: two row and two column coefficient sets
        const bool al = m.upsample == PWS_UP_ALIGNED;
        UpCoef cy[2], cx[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) cy[j] = up_coef(min(blockIdx.y * kTileH + wrp + 8 * j, g.Ho - 1), m.mh, rh, al);
#pragma unroll
        for (int i = 0; i < 2; ++i) cx[i] = up_coef(min(blockIdx.x * kTileW + lane + 32 * i, g.Wo - 1), m.mw, rw, al);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) map_up(m, n, cy[j], cx[i], gx[j][i], gy[j][i]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int w = blockIdx.x * kTileW + lane + 32 * i, h = blockIdx.y * kTileH + wrp + 8 * j;
            if (!(w < g.Wo && h < g.Ho)) continue;
            Taps<float> t;
            make_taps(source_index(gx[j][i], g.W, g.padding, align), source_index(gy[j][i], g.H, g.padding, align), g.H, g.W, t);
            const int o_nw = t.y0 * in.s2 + t.x0 * in.s3;
            const int o_out = h * out.s2 + w * out.s3;
#pragma unroll 3
            for (int c = 0; c < C; ++c) {
                const TI *__restrict__ pc = ip + c * in.s1;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                if (t.mask & 1u) v0 = load_px(pc + o_nw);
                if (t.mask & 2u) v1 = load_px(pc + o_nw + in.s3);
                if (t.mask & 4u) v2 = load_px(pc + o_nw + in.s2);
                if (t.mask & 8u) v3 = load_px(pc + o_nw + in.s2 + in.s3);
                if (m.has_pre) {  // (x + a) * b, elementwise as the reference does before sampling
                    v0 = __fmul_rn(__fadd_rn(v0, m.pre_add), m.pre_mul); v1 = __fmul_rn(__fadd_rn(v1, m.pre_add), m.pre_mul);
                    v2 = __fmul_rn(__fadd_rn(v2, m.pre_add), m.pre_mul); v3 = __fmul_rn(__fadd_rn(v3, m.pre_add), m.pre_mul);
                }
                float acc = 0.f;
                if (t.mask & 1u) acc = ffma(v0, t.nw, acc);
                if (t.mask & 2u) acc = ffma(v1, t.ne, acc);
                if (t.mask & 4u) acc = ffma(v2, t.sw, acc);
                if (t.mask & 8u) acc = ffma(v3, t.se, acc);
                if (m.has_post) acc = __fadd_rn(__fdiv_rn(acc, m.post_div), m.post_add);
                store_px(op + o_out + c * out.s1, acc);
            }
        }
}

// The inference site, specialised (R/main_new.py:679-684,697-721): uint8 HWC frame in, uint8 HWC frame out, the map
// is the bilinear upsample of a dense interleaved low-resolution lattice (what the host side composes once per call).
// Same arithmetic as the generic kernel above, statement for statement -- only the addressing is compile-time: the
// generic kernel spends 390 instructions per pixel, most of them on run-time strides and layout checks, and is
// issue-bound (79 % of the issue slots).
template <bool kBorder, bool kAlignCorners, bool kUpAligned>
__global__ void __launch_bounds__(kThreads)
fwd_fused_u8_kernel(const uint8_t *__restrict__ in, const float2 *__restrict__ lattice, uint8_t *__restrict__ out,
                    const int H, const int W, const int Ho, const int Wo, const int mh, const int mw, const float rh, const float rw)
{
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int n = blockIdx.z;
    const uint8_t *__restrict__ ip = in + (int64_t)n * H * W * 3;
    uint8_t *__restrict__ op = out + (int64_t)n * Ho * Wo * 3;
    const float2 *__restrict__ L = lattice + (int64_t)n * mh * mw;
    const int padding = kBorder ? PWS_PAD_BORDER : PWS_PAD_ZEROS;

    UpCoef cy[2], cx[2];
    int hh[2], ww[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) { hh[j] = blockIdx.y * kTileH + wrp + 8 * j; cy[j] = up_coef(min(hh[j], Ho - 1), mh, rh, kUpAligned); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { ww[i] = blockIdx.x * kTileW + lane + 32 * i; cx[i] = up_coef(min(ww[i], Wo - 1), mw, rw, kUpAligned); }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (!(ww[i] < Wo && hh[j] < Ho)) continue;
            const float2 *r0 = L + cy[j].i0 * mw + cx[i].i0, *r1 = r0 + cy[j].ip * mw;
            const float2 v00 = __ldg(r0), v01 = __ldg(r0 + cx[i].ip), v10 = __ldg(r1), v11 = __ldg(r1 + cx[i].ip);
            // h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11), contracted as in map_up()
            const float gx = __fmaf_rn(cy[j].l0, __fmaf_rn(cx[i].l0, v00.x, __fmul_rn(cx[i].l1, v01.x)), __fmul_rn(cy[j].l1, __fmaf_rn(cx[i].l0, v10.x, __fmul_rn(cx[i].l1, v11.x))));
            const float gy = __fmaf_rn(cy[j].l0, __fmaf_rn(cx[i].l0, v00.y, __fmul_rn(cx[i].l1, v01.y)), __fmul_rn(cy[j].l1, __fmaf_rn(cx[i].l0, v10.y, __fmul_rn(cx[i].l1, v11.y))));
            Taps<float> t;
            make_taps(source_index(gx, W, padding, kAlignCorners), source_index(gy, H, padding, kAlignCorners), H, W, t);
            const uint8_t *__restrict__ p0 = ip + ((int64_t)t.y0 * W + t.x0) * 3;
            const uint8_t *__restrict__ p1 = p0 + W * 3;
            uint8_t *__restrict__ o = op + ((int64_t)hh[j] * Wo + ww[i]) * 3;
            if (t.mask == 15u) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float acc = ffma((float)__ldg(p0 + c), t.nw, 0.f);
                    acc = ffma((float)__ldg(p0 + 3 + c), t.ne, acc);
                    acc = ffma((float)__ldg(p1 + c), t.sw, acc);
                    acc = ffma((float)__ldg(p1 + 3 + c), t.se, acc);
                    store_px(o + c, acc);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float acc = 0.f;
                    if (t.mask & 1u) acc = ffma((float)__ldg(p0 + c), t.nw, acc);
                    if (t.mask & 2u) acc = ffma((float)__ldg(p0 + 3 + c), t.ne, acc);
                    if (t.mask & 4u) acc = ffma((float)__ldg(p1 + c), t.sw, acc);
                    if (t.mask & 8u) acc = ffma((float)__ldg(p1 + 3 + c), t.se, acc);
                    store_px(o + c, acc);
                }
            }
        }
}

__global__ void __launch_bounds__(256)
compose_map_kernel(const MapSpec m, const View out, const int N, const int Ho, const int Wo)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * Ho * Wo) return;
    const int w = (int)(idx % Wo);
    const int64_t r = idx / Wo;
    const int h = (int)(r % Ho), n = (int)(r / Ho);
    float rh, rw, gx, gy;
    up_scales(m, Ho, Wo, rh, rw);
    map_at(m, n, h, w, rh, rw, gx, gy);
    float *o = (float *)out.p + (int64_t)n * out.sN + h * out.s1 + w * out.s2;
    o[0] = gx; o[out.s3] = gy;
}

template <typename TI, typename TO>
void launch_io(const View &in, const MapSpec &m, const View &out, const Geometry &g, cudaStream_t st)
{
    dim3 blocks((g.Wo + kTileW - 1) / kTileW, (g.Ho + kTileH - 1) / kTileH, g.N);
    if (g.C == 3) { fwd_fused_kernel<TI, TO, 3><<<blocks, kThreads, 0, st>>>(in, m, out, g); note_launch(); }
    else if (g.C == 1) { fwd_fused_kernel<TI, TO, 1><<<blocks, kThreads, 0, st>>>(in, m, out, g); note_launch(); }
    else { fwd_fused_kernel<TI, TO, 0><<<blocks, kThreads, 0, st>>>(in, m, out, g); note_launch(); }
    note_kernel("fwd_fused");
}

}  // namespace

int launch_forward_fused(const View &in, int in_dtype, const MapSpec &m, const View &out, int out_dtype,
                         const Geometry &g, cudaStream_t st)
{
    if (g.N > 65535 || (g.Ho + kTileH - 1) / kTileH > 65535) { set_error("fused forward: batch or height too large"); return PWS_EUNSUPPORTED; }
    // the inference site: dense uint8 HWC in and out, dense interleaved lattice, upsampling, nothing else
    const bool u8_fast = in_dtype == PWS_U8 && out_dtype == PWS_U8 && g.C == 3 && m.base == PWS_BASE_NONE && m.drift.p &&
        m.upsample != PWS_UP_NONE && !m.has_pre && !m.has_post &&
        in.s1 == 1 && in.s3 == 3 && in.s2 == 3 * g.W && (g.N == 1 || in.sN == (int64_t)3 * g.W * g.H) &&
        out.s1 == 1 && out.s3 == 3 && out.s2 == 3 * g.Wo && (g.N == 1 || out.sN == (int64_t)3 * g.Wo * g.Ho) &&
        m.drift.s3 == 1 && m.drift.s2 == 2 && m.drift.s1 == 2 * m.mw && (g.N == 1 || m.drift.sN == (int64_t)2 * m.mw * m.mh) &&
        !(reinterpret_cast<uintptr_t>(m.drift.p) & 7);
    if (u8_fast) {
        dim3 blocks((g.Wo + kTileW - 1) / kTileW, (g.Ho + kTileH - 1) / kTileH, g.N);
        float rh, rw;
        up_scales_host(m, g.Ho, g.Wo, rh, rw);
        const uint8_t *ip = (const uint8_t *)in.p;
        const float2 *lp = (const float2 *)m.drift.p;
        uint8_t *op = (uint8_t *)out.p;
        const bool border = g.padding == PWS_PAD_BORDER, al = g.align != 0, upa = m.upsample == PWS_UP_ALIGNED;
#define PWS_U8(B, A, U) fwd_fused_u8_kernel<B, A, U><<<blocks, kThreads, 0, st>>>(ip, lp, op, g.H, g.W, g.Ho, g.Wo, m.mh, m.mw, rh, rw)
        if (border) { if (al) { if (upa) PWS_U8(true, true, true); else PWS_U8(true, true, false); } else { if (upa) PWS_U8(true, false, true); else PWS_U8(true, false, false); } }
        else        { if (al) { if (upa) PWS_U8(false, true, true); else PWS_U8(false, true, false); } else { if (upa) PWS_U8(false, false, true); else PWS_U8(false, false, false); } }
#undef PWS_U8
        note_launch();
        note_kernel("fwd_fused_u8");
        return PWS_OK;
    }
    if (in_dtype == PWS_F32 && out_dtype == PWS_F32) launch_io<float, float>(in, m, out, g, st);
    else if (in_dtype == PWS_U8 && out_dtype == PWS_F32) launch_io<uint8_t, float>(in, m, out, g, st);
    else if (in_dtype == PWS_U8 && out_dtype == PWS_U8) launch_io<uint8_t, uint8_t>(in, m, out, g, st);
    else if (in_dtype == PWS_F32 && out_dtype == PWS_U8) launch_io<float, uint8_t>(in, m, out, g, st);
    else if (in_dtype == PWS_BF16 && out_dtype == PWS_BF16) launch_io<__nv_bfloat16, __nv_bfloat16>(in, m, out, g, st);
    else if (in_dtype == PWS_F16 && out_dtype == PWS_F16) launch_io<__half, __half>(in, m, out, g, st);
    else { set_error("fused forward: unsupported frame/output dtype pair (%d -> %d)", in_dtype, out_dtype); return PWS_EUNSUPPORTED; }
    return PWS_OK;
}

int launch_compose_map(const MapSpec &m, const View &out, int N, int Ho, int Wo, cudaStream_t st)
{
    const int64_t total = (int64_t)N * Ho * Wo;
    if (total == 0) return PWS_OK;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > INT_MAX) { set_error("compose_map: too many pixels"); return PWS_EUNSUPPORTED; }
    compose_map_kernel<<<(unsigned)blocks, 256, 0, st>>>(m, out, N, Ho, Wo);
    note_launch();
    note_kernel("compose_map");
    return PWS_OK;
}

}  // namespace pws
