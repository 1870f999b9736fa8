// warp_fwd_tma.cu -- TMA-pipelined persistent forward warp (sm_100a).
//
// Call sites: R/main_new.py:106,116 (NCHW fp32 frames, planar-stored maps), :197 (interleaved
// affine_grid maps), :716 (the same at native video resolution).
//
// One persistent CTA per SM works on 64x16 tiles of OUTPUT pixels handed out by a per-launch counter.  Three roles,
// mbarrier rings between them:
//   warp 0  producer  fetches tile indices and streams the tiles' warp maps into shared memory (TMA, kMapStages ahead);
//   warps 1-2 scouts  (alternate tiles) reduce the map tile to the bounding box of the source taps it implies --
//                     the tile's halo under the map -- picks the smallest of three box shapes that
//                     holds it and issues the TMA load of that box of the frame (all channels);
//   warps 3-10 consumers gather the four taps of every pixel out of the shared-memory box and
//                     store the result with 128-byte coalesced writes.
// Neither the map nor the frame is touched by a generic load, so no warp ever waits on DRAM with
// its own registers: HBM latency is covered by the depth of the two rings instead of by occupancy.
// Tiles whose every tap is inside the frame run a mask-free body; tiles whose box fits no shape
// (violent maps) or whose map holds NaN/inf gather from global memory inside the same kernel.
// Arithmetic is the ATen order of pws_common.cuh / pws_tile.cuh: results are bit-identical to the
// other forward kernels.
// kCL: the frame is a channels-last view (R/main_new.py:679-684,716: permute(0,3,1,2) of an HWC buffer).  The
// box is then a (C*BW) x BH slab of the (C*W, H, N) array -- a pixel's channels sit in consecutive words, lanes
// on consecutive pixels are 3 words apart (no bank conflict) -- and the output stays dense NCHW like ATen's.
// T = __half / __nv_bfloat16: 16-bit frames with fp32 maps (BASELINE config 5); the box holds 16-bit elements, taps are
// upcast, the arithmetic is fp32 and the result is rounded once.
#include "pws_pipe.cuh"
#include "pws_launch.cuh"

#include <atomic>
#include <cstdlib>

namespace pws {

using namespace pipe;

namespace {

constexpr int kScouts = 2, kGroups = 2, kGroupWarps = 8, kConsumers = kGroups * kGroupWarps;
constexpr int kThreads = (1 + kScouts + kConsumers) * 32;
#ifndef PWS_FWD_MAP_STAGES
#define PWS_FWD_MAP_STAGES 8
#endif
#ifndef PWS_FWD_BOX_STAGES
#define PWS_FWD_BOX_STAGES 4
#endif
constexpr int kMapStages = PWS_FWD_MAP_STAGES, kBoxStages = PWS_FWD_BOX_STAGES;
constexpr int kInfoStop = 1 << 11;  // info.z: no more tiles for this consumer group
static_assert(kScouts == kGroups, "scout w feeds consumer group w (and tells it when the tiles have run out)");

// Tiles are handed out dynamically, as in the backward (warp_bwd_tma.cu): the producer takes the next tile index from a
// per-launch counter and passes it on through shared memory.  A launch owns one slot (pws_launch.cuh: the slot is
// leased per device and reused only after its previous launch has finished); the last CTA to leave resets it.
constexpr int kSlots = kLaunchSlots;
__device__ unsigned int g_tile_next[kSlots];
__device__ unsigned int g_exit_count[kSlots];

template <int CS, int kElem = 4> struct Smem {
    static constexpr int kBoxBytes = (kMaxBW * kMaxBH * CS * kElem + 127) / 128 * 128;
    static constexpr int kMapOff = 0;
    static constexpr int kBoxOff = kMapStages * kMapTileBytes;
    static constexpr int kInfoOff = kBoxOff + kBoxStages * kBoxBytes;
    static constexpr int kBarOff = kInfoOff + kBoxStages * 32;
    static constexpr int kTileOff = kBarOff + (2 * kMapStages + 2 * kBoxStages) * 8;  // tile index of each map stage
    static constexpr int kTotal = kTileOff + kMapStages * 4;
    static_assert(kBoxBytes % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

struct TmaParams {
    CUtensorMap map;                // planar: (Wo, Ho, 2, N) box (64,16,2,1); interleaved: (2Wo, Ho, N) box (128,16,1)
    CUtensorMap box[kNumShapes];    // frame (W, H, C, N), box (BW, BH, CS, 1); channels-last: (C*W, H, N), box (CS*BW, BH, 1)
};

// T: frame / output element (float; __half or __nv_bfloat16 with fp32 maps: taps are upcast, the arithmetic is fp32,
// the result is rounded once -- the upcast-sample-round semantics of the other 16-bit kernels)
template <typename T, int CS, bool kBorder, bool kAlign, bool kInter, bool kCL>
__global__ void __launch_bounds__(kThreads, 1)
fwd_tma_kernel(const __grid_constant__ TmaParams tp, const View in, const View out, const Geometry g,
               const int tiles_x, const int tiles_y, const int total_tiles, const int slot)
{
    using S = Smem<CS, (int)sizeof(T)>;
    constexpr int kXAlign = 16 / (int)sizeof(T);
    extern __shared__ __align__(1024) unsigned char smem[];
    float *const s_map = reinterpret_cast<float *>(smem + S::kMapOff);
    unsigned char *const s_box = smem + S::kBoxOff;
    int4 *const s_info = reinterpret_cast<int4 *>(smem + S::kInfoOff);
    uint64_t *const map_full = reinterpret_cast<uint64_t *>(smem + S::kBarOff);
    uint64_t *const map_empty = map_full + kMapStages;
    uint64_t *const box_full = map_empty + kMapStages;
    uint64_t *const box_empty = box_full + kBoxStages;
    volatile int *const s_tile = reinterpret_cast<volatile int *>(smem + S::kTileOff);

    // Roles are numbered from the TOP warp of the CTA down: the scheduler favours the higher warp ids when several
    // warps are ready, and the producer and the scouts -- a handful of instructions per tile, but every consumer
    // waits on them -- must not queue behind sixteen busy consumer warps.
    const int warp = (kThreads / 32 - 1) - (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int tiles_xy = tiles_x * tiles_y;
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMapStages; ++s) { tma::mbar_init(map_full + s, 1); tma::mbar_init(map_empty + s, kGroupWarps); }
        for (int s = 0; s < kBoxStages; ++s) { tma::mbar_init(box_full + s, 1); tma::mbar_init(box_empty + s, kGroupWarps); }
        tma::fence_barrier_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ===== producer: fetches tile indices (the first per CTA is static, the rest come from the launch's counter, one
        // fetch ahead) and streams the tiles' warp maps; a negative index tells a scout that the tiles have run out =====
        if (lane == 0) {
            tma::prefetch_desc(&tp.map);
            int t = blockIdx.x;
            int t_next = (int)atomicAdd(&g_tile_next[slot], 1u) + (int)gridDim.x;
            int stops = 0;
            for (int it = 0; stops < kScouts; ++it) {
                const int s = it % kMapStages, ph = (it / kMapStages) & 1;
                tma::mbar_wait_relaxed(map_empty + s, ph ^ 1);
                if (t >= total_tiles) {
                    s_tile[s] = -1;
                    tma::mbar_arrive(map_full + s);
                    ++stops;
                    continue;
                }
                const int t_new = (int)atomicAdd(&g_tile_next[slot], 1u) + (int)gridDim.x;
                const TileCoord tc = tile_coord(t, tiles_x, tiles_xy);
                s_tile[s] = t;
                tma::mbar_arrive_expect_tx(map_full + s, kMapTileBytes);
                if (kInter) tma::load_3d(s_map + s * kMapTileFloats, &tp.map, map_full + s, 2 * tc.w0, tc.h0, tc.n);
                else tma::load_4d(s_map + s * kMapTileFloats, &tp.map, map_full + s, tc.w0, tc.h0, 0, tc.n);
                t = t_next; t_next = t_new;
            }
        }
    } else if (warp <= kScouts) {
        // ===== scouts (alternate tiles): map tile -> tap bounding box -> frame box load =====
        if (lane == 0) { tma::prefetch_desc(&tp.box[0]); tma::prefetch_desc(&tp.box[1]); tma::prefetch_desc(&tp.box[2]); }
        for (int it = warp - 1;; it += kScouts) {
            const int ms = it % kMapStages, mph = (it / kMapStages) & 1;
            const int bs = it % kBoxStages, bph = (it / kBoxStages) & 1;
            tma::mbar_wait_relaxed(map_full + ms, mph);
            const int t = s_tile[ms];
            if (t < 0) {
                tma::mbar_wait_relaxed(box_empty + bs, bph ^ 1);
                if (lane == 0) { s_info[2 * bs] = make_int4(0, 0, kInfoStop, 0); tma::mbar_arrive(box_full + bs); }
                break;
            }
            const TileCoord tc = tile_coord(t, tiles_x, tiles_xy);
            const int cols = min(kTW, g.Wo - tc.w0), rows = min(kTH, g.Ho - tc.h0);
            float xlo, xhi, ylo, yhi;
            map_tile_range<kInter>(s_map + ms * kMapTileFloats, rows, cols, lane, xlo, xhi, ylo, yhi);
            tma::mbar_wait_relaxed(box_empty + bs, bph ^ 1);
            if (lane == 0) {
                int4 info = box_of_range<kBorder, kAlign, kCL, kXAlign>(xlo, xhi, ylo, yhi, g.W, g.H, cols == kTW && rows == kTH);
                info.w = tc.n;
                s_info[2 * bs] = info;
                s_info[2 * bs + 1] = make_int4(tc.h0, tc.w0, 0, 0);
                if (info.z & (kInfoFallback | kInfoEmpty)) tma::mbar_arrive(box_full + bs);
                else {
                    const int shape = info.z & 0xff;
                    tma::mbar_arrive_expect_tx(box_full + bs, box_w_of<kCL>(shape) * box_h(shape) * CS * (int)sizeof(T));
                    if (kCL) tma::load_3d(s_box + (size_t)bs * S::kBoxBytes, &tp.box[shape], box_full + bs, CS * info.x, info.y, tc.n);
                    else tma::load_4d(s_box + (size_t)bs * S::kBoxBytes, &tp.box[shape], box_full + bs, info.x, info.y, 0, tc.n);
                }
            }
            __syncwarp();
        }
    } else {
        // ===== consumers =====
        const int cw = warp - 1 - kScouts, grp = cw / kGroupWarps, wg = cw % kGroupWarps;
        for (int it = grp;; it += kGroups) {
            const int ms = it % kMapStages;
            const int bs = it % kBoxStages, bph = (it / kBoxStages) & 1;
            // one warp of the group polls the mbarrier, the others park on a hardware barrier (no spin); the scout has
            // seen the map tile before it loaded the box, so the box barrier covers both
            if (wg == 0) tma::mbar_wait(box_full + bs, bph);
            tma::named_bar_sync(1 + grp, kGroupWarps * 32);
            const float *mp = s_map + ms * kMapTileFloats;
            const int4 info = s_info[2 * bs], where = s_info[2 * bs + 1];
            if (info.z & kInfoStop) break;
            TileCoord tc;
            tc.n = info.w; tc.h0 = where.x; tc.w0 = where.y;
            const int shape = info.z & 0xff;
            // element strides inside the box: next pixel, next row, next channel
            constexpr int px = kCL ? CS : 1;
            const int pitch = px * box_w_of<kCL>(shape);
            const int plane = kCL ? 1 : box_w(shape) * box_h(shape);
            const T *bp = reinterpret_cast<const T *>(s_box + (size_t)bs * S::kBoxBytes);
            T *__restrict__ op = (T *)out.p + (int64_t)tc.n * out.sN + (int64_t)tc.h0 * out.s2 + tc.w0;
            const int o_ch = out.s1, o_row = out.s2;

            if (info.z & kInfoInterior) {
                const int base = -(info.y * pitch + info.x * px);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = wg * 2 + j;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int x = lane + 32 * i;
                        float gx, gy;
                        if (kInter) { const float2 v = *reinterpret_cast<const float2 *>(mp + r * (2 * kTW) + 2 * x); gx = v.x; gy = v.y; }
                        else { gx = mp[r * kTW + x]; gy = mp[kTW * kTH + r * kTW + x]; }
                        const float ix = unnorm<kAlign>(gx, Wf, Wm1), iy = unnorm<kAlign>(gy, Hf, Hm1);
                        float x0f, y0f; int x0, y0;
                        floor_small(ix, x0f, x0); floor_small(iy, y0f, y0);
                        const float wx1 = fsub(ix, x0f), wx0 = fsub(x0f + 1.0f, ix);
                        const float wy1 = fsub(iy, y0f), wy0 = fsub(y0f + 1.0f, iy);
                        const float nw = fmul(wx0, wy0), ne = fmul(wx1, wy0), sw = fmul(wx0, wy1), se = fmul(wx1, wy1);
                        const T *__restrict__ p0 = bp + (y0 * pitch + x0 * px + base);
                        const T *__restrict__ p1 = p0 + pitch;
                        T *__restrict__ o = op + (r * o_row + x);
#pragma unroll
                        for (int c = 0; c < CS; ++c) {
                            float acc = ffma(to_acc(p0[c * plane]), nw, 0.f);
                            acc = ffma(to_acc(p0[c * plane + px]), ne, acc);
                            acc = ffma(to_acc(p1[c * plane]), sw, acc);
                            acc = ffma(to_acc(p1[c * plane + px]), se, acc);
                            o[c * o_ch] = from_acc<T, float>(acc);
                        }
                    }
                }
            } else {
                const bool fallback = (info.z & kInfoFallback) != 0;
                const T *__restrict__ ip = (const T *)in.p + (int64_t)tc.n * in.sN;
                const int sH = in.s2, i_ch = in.s1, sW = kCL ? in.s3 : 1;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = wg * 2 + j;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int x = lane + 32 * i;
                        if (tc.h0 + r >= g.Ho || tc.w0 + x >= g.Wo) continue;
                        float gx, gy;
                        if (kInter) { const float2 v = *reinterpret_cast<const float2 *>(mp + r * (2 * kTW) + 2 * x); gx = v.x; gy = v.y; }
                        else { gx = mp[r * kTW + x]; gy = mp[kTW * kTH + r * kTW + x]; }
                        const float ix = src_index<kBorder, kAlign>(gx, g.W, Wf, Wm1);
                        const float iy = src_index<kBorder, kAlign>(gy, g.H, Hf, Hm1);
                        Taps<float> tp4;
                        make_taps(ix, iy, g.H, g.W, tp4);
                        T *__restrict__ o = op + (r * o_row + x);
                        if (!fallback) {
                            const T *__restrict__ p0 = bp + ((tp4.y0 - info.y) * pitch + (tp4.x0 - info.x) * px);
#pragma unroll
                            for (int c = 0; c < CS; ++c) {
                                const T *__restrict__ pc = p0 + c * plane;
                                float acc = 0.f;
                                if (tp4.mask & 1u) acc = ffma(to_acc(pc[0]), tp4.nw, acc);
                                if (tp4.mask & 2u) acc = ffma(to_acc(pc[px]), tp4.ne, acc);
                                if (tp4.mask & 4u) acc = ffma(to_acc(pc[pitch]), tp4.sw, acc);
                                if (tp4.mask & 8u) acc = ffma(to_acc(pc[pitch + px]), tp4.se, acc);
                                o[c * o_ch] = from_acc<T, float>(acc);
                            }
                        } else {
                            const T *__restrict__ p0 = ip + (tp4.y0 * sH + tp4.x0 * sW);
#pragma unroll
                            for (int c = 0; c < CS; ++c) {
                                const T *__restrict__ pc = p0 + c * i_ch;
                                float acc = 0.f;
                                if (tp4.mask & 1u) acc = ffma(to_acc(ldg(pc)), tp4.nw, acc);
                                if (tp4.mask & 2u) acc = ffma(to_acc(ldg(pc + sW)), tp4.ne, acc);
                                if (tp4.mask & 4u) acc = ffma(to_acc(ldg(pc + sH)), tp4.sw, acc);
                                if (tp4.mask & 8u) acc = ffma(to_acc(ldg(pc + sH + sW)), tp4.se, acc);
                                o[c * o_ch] = from_acc<T, float>(acc);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { tma::mbar_arrive(map_empty + ms); tma::mbar_arrive(box_empty + bs); }
        }
    }
    __syncthreads();
    // the last CTA to leave hands the counter slot back clean
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&g_exit_count[slot], 1u) == gridDim.x - 1) {
            g_tile_next[slot] = 0u;
            g_exit_count[slot] = 0u;
            __threadfence();
        }
    }
}

template <typename T, int CS, bool kBorder, bool kAlign, bool kInter, bool kCL>
bool launch_k(const TmaParams &tp, const Problem &pb, int tiles_x, int tiles_y, int total, cudaStream_t st)
{
    auto kern = fwd_tma_kernel<T, CS, kBorder, kAlign, kInter, kCL>;
    constexpr int kSmem = Smem<CS, (int)sizeof(T)>::kTotal;
    static std::atomic<uint64_t> attr_done{0};  // per instantiation, one bit per device
    if (!ensure_dynamic_smem(reinterpret_cast<const void *>(kern), kSmem, attr_done)) return false;
    const int grid = total < sm_count() ? total : sm_count();
    SlotLease lease(kRingForward, st);
    if (!lease.ok()) return false;  // stream capture: the caller takes the non-persistent kernel
    kern<<<grid, kThreads, kSmem, st>>>(tp, pb.in, pb.out, pb.g, tiles_x, tiles_y, total, lease.slot());
    note_launch();
    note_kernel(kCL ? "fwd_tma_cl" : sizeof(T) == 2 ? "fwd_tma_16" : "fwd_tma");
    return true;
}

template <typename T, int CS, bool kInter, bool kCL>
bool launch_ba(const TmaParams &tp, const Problem &pb, int tx, int ty, int total, cudaStream_t st)
{
    const bool border = pb.g.padding == PWS_PAD_BORDER, align = pb.g.align != 0;
    if (border && align) return launch_k<T, CS, true, true, kInter, kCL>(tp, pb, tx, ty, total, st);
    if (border) return launch_k<T, CS, true, false, kInter, kCL>(tp, pb, tx, ty, total, st);
    if (align) return launch_k<T, CS, false, true, kInter, kCL>(tp, pb, tx, ty, total, st);
    return launch_k<T, CS, false, false, kInter, kCL>(tp, pb, tx, ty, total, st);
}

template <typename T>
bool launch_c(const TmaParams &tp, const Problem &pb, bool inter, int tx, int ty, int total, cudaStream_t st)
{
    if (pb.g.C == 3) return inter ? launch_ba<T, 3, true, false>(tp, pb, tx, ty, total, st) : launch_ba<T, 3, false, false>(tp, pb, tx, ty, total, st);
    return inter ? launch_ba<T, 1, true, false>(tp, pb, tx, ty, total, st) : launch_ba<T, 1, false, false>(tp, pb, tx, ty, total, st);
}

}  // namespace

bool tma_disabled()
{
#ifdef PWS_DEV_HOOKS   // development builds only: the product library has no run-time switches
    static const bool v = [] { const char *e = std::getenv("PWS_NO_TMA"); return e && e[0] == '1'; }();
    return v;
#else
    return false;
#endif
}

int sm_count()
{
    static int n[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

// Encodes the map tensor map of a TMA kernel (shared with the backward): planar maps as (Wo,Ho,2,N),
// interleaved ones as (2Wo,Ho,N).  Returns false when the layout is neither or misaligned.
bool encode_map_tma(const View &grid, const Geometry &g, CUtensorMap *tm, bool *inter)
{
    if (reinterpret_cast<uintptr_t>(grid.p) & 15) return false;
    if (g.Wo % 4) return false;
    if (grid.s2 == 1 && grid.s3 >= 1) {  // planar rows
        if ((grid.s1 % 4) || (grid.s3 % 4) || (grid.sN % 4)) return false;
        *inter = false;
        const uint64_t dims[4] = {(uint64_t)g.Wo, (uint64_t)g.Ho, 2, (uint64_t)g.N};
        const uint64_t sN = g.N > 1 ? (uint64_t)grid.sN : (uint64_t)grid.s3 * 2;
        const uint64_t strides[3] = {(uint64_t)grid.s1, (uint64_t)grid.s3, sN};
        const uint32_t box[4] = {kTW, kTH, 2, 1};
        return tma::encode_f32(tm, grid.p, 4, dims, strides, box);
    }
    if (grid.s3 == 1 && grid.s2 == 2) {  // interleaved pairs
        if ((grid.s1 % 4) || (grid.sN % 4)) return false;
        *inter = true;
        const uint64_t dims[3] = {(uint64_t)2 * g.Wo, (uint64_t)g.Ho, (uint64_t)g.N};
        const uint64_t sN = g.N > 1 ? (uint64_t)grid.sN : (uint64_t)grid.s1 * g.Ho;
        const uint64_t strides[2] = {(uint64_t)grid.s1, sN};
        const uint32_t box[3] = {2 * kTW, kTH, 1};
        return tma::encode_f32(tm, grid.p, 3, dims, strides, box);
    }
    return false;
}

// frame-shaped (W,H,C,N) tensor map with the given box; fp32 or 16-bit float elements
bool encode_frame_tma(const View &v, int W, int H, int C, int N, int bw, int bh, int bc, CUtensorMap *tm, int dtype)
{
    const int eb = dtype == PWS_F32 ? 4 : 2, al = 16 / eb;  // element bytes, elements per 16 bytes
    if (dtype != PWS_F32 && dtype != PWS_F16 && dtype != PWS_BF16) return false;
    if (reinterpret_cast<uintptr_t>(v.p) & 15) return false;
    if (v.s3 != 1 || (v.s2 % al) || (v.s1 % al) || (v.sN % al) || (bw % al)) return false;
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N};
    const uint64_t s1 = C > 1 ? (uint64_t)v.s1 : (uint64_t)v.s2 * H;
    const uint64_t sN = N > 1 ? (uint64_t)v.sN : s1 * C;
    const uint64_t strides[3] = {(uint64_t)v.s2, s1, sN};
    const uint32_t box[4] = {(uint32_t)bw, (uint32_t)bh, (uint32_t)bc, 1};
    const CUtensorMapDataType dt = dtype == PWS_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == PWS_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    return tma::encode_elems(tm, dt, eb, v.p, 4, dims, strides, box);
}

// channels-last frame (element strides s1 == 1, s3 == C): (C*W, H, N) fp32 tensor map, box (C*bw, bh, 1)
bool encode_frame_cl_tma(const View &v, int W, int H, int C, int N, int bw, int bh, CUtensorMap *tm)
{
    if (reinterpret_cast<uintptr_t>(v.p) & 15) return false;
    if (v.s1 != 1 || v.s3 != C || (v.s2 % 4) || (v.sN % 4) || C * bw > 256) return false;
    const uint64_t dims[3] = {(uint64_t)C * W, (uint64_t)H, (uint64_t)N};
    const uint64_t sN = N > 1 ? (uint64_t)v.sN : (uint64_t)v.s2 * H;
    const uint64_t strides[2] = {(uint64_t)v.s2, sN};
    const uint32_t box[3] = {(uint32_t)(C * bw), (uint32_t)bh, 1};
    return tma::encode_f32(tm, v.p, 3, dims, strides, box);
}

// Returns true when the TMA kernel took the call.
bool launch_forward_tma(const Problem &pb, cudaStream_t st)
{
    const Geometry &g = pb.g;
    if (tma_disabled()) return false;
    if (pb.grid_dtype != PWS_F32) return false;
    if (pb.in_dtype != PWS_F32 && pb.in_dtype != PWS_F16 && pb.in_dtype != PWS_BF16) return false;
    if (g.C != 1 && g.C != 3) return false;
    if (pb.out.s3 != 1) return false;
    if (g.W > (1 << 22) || g.H > (1 << 22)) return false;  // floor_small
    const int tiles_x = (g.Wo + kTW - 1) / kTW, tiles_y = (g.Ho + kTH - 1) / kTH;
    const int64_t total = (int64_t)tiles_x * tiles_y * g.N;
    if (total <= 0 || total > INT_MAX) return false;
    TmaParams tp;
    bool inter = false;
    if (!encode_map_tma(pb.grid, g, &tp.map, &inter)) return false;
    if (pb.in_dtype != PWS_F32) {  // 16-bit frames (BASELINE config 5: 4K bf16 frames, fp32 maps)
        for (int s = 0; s < kNumShapes; ++s)
            if (!encode_frame_tma(pb.in, g.W, g.H, g.C, g.N, box_w(s), box_h(s), g.C, &tp.box[s], pb.in_dtype)) return false;
        return pb.in_dtype == PWS_F16 ? launch_c<__half>(tp, pb, inter, tiles_x, tiles_y, (int)total, st)
                                      : launch_c<__nv_bfloat16>(tp, pb, inter, tiles_x, tiles_y, (int)total, st);
    }
    if (g.C == 3 && pb.in.s1 == 1 && pb.in.s3 == 3) {  // channels-last RGB (the inference site)
        for (int s = 0; s < kNumShapes; ++s)
            if (!encode_frame_cl_tma(pb.in, g.W, g.H, g.C, g.N, box_w_of<true>(s), box_h(s), &tp.box[s])) return false;
        return inter ? launch_ba<float, 3, true, true>(tp, pb, tiles_x, tiles_y, (int)total, st) : launch_ba<float, 3, false, true>(tp, pb, tiles_x, tiles_y, (int)total, st);
    }
    for (int s = 0; s < kNumShapes; ++s)
        if (!encode_frame_tma(pb.in, g.W, g.H, g.C, g.N, box_w(s), box_h(s), g.C, &tp.box[s])) return false;
    return launch_c<float>(tp, pb, inter, tiles_x, tiles_y, (int)total, st);
}

}  // namespace pws
