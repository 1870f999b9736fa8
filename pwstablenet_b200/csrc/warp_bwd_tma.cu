// warp_bwd_tma.cu -- TMA-pipelined persistent backward warp (sm_100a).
//
// Autograd of R/main_new.py:106,116 (grad -> map) and :197 (grad -> frame), triggered at :214.
//
// Same pipeline as warp_fwd_tma.cu: one persistent CTA per SM walks 64x16 tiles of OUTPUT pixels;
//   warp 0      producer   streams the tile's warp map and grad_output into shared memory (TMA);
//   warps 1-2   scouts     (alternate tiles) reduce the map tile to the bounding box of its source taps and
//                          load that box of the frame -- only when grad_grid is wanted, it is its one use;
//   warps 3-10  consumers  two groups of warps, each group owning every other tile; inside a tile a
//                          warp owns a 32-pixel-wide strip of 2*16/warps-per-group rows and marches down it.
// The scatter into grad_input is the marching scheme of warp_bwd_lean.cu: a lane whose right neighbour
// samples the next source pixel hands its east taps over by shuffle, the south taps ride down the strip
// in registers, and what is left is ~1 RED.ADD.F32 per source pixel and channel on consecutive addresses.
// The stragglers (east taps nobody takes over, parked sums whose chain broke) are compacted through a
// per-warp queue in shared memory and leave as dense 32-lane REDs (the L2 accepts ~25 G RED
// instructions/s whether 1 or 32 lanes are active: tools/exp/exp_red.cu).
// grad_grid is accumulated in ATen's exact statement order from taps read out of the shared-memory box:
// bit-identical to the other backward kernels; grad_input differs only by atomic order.
#include "pws_pipe.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>

namespace pws {

using namespace pipe;

namespace {

#ifdef PWS_EXP_NORED   // experiment: drop the scatter's atomics (results are wrong) to time everything else
#define PWS_RED(p, v) do { if ((v) == 1.2345e-30f) tma::red_add_f32((p), (v)); } while (0)
#else
// explicit fire-and-forget reduction: with a fence elsewhere in the kernel nvcc turns atomicAdd into the
// returning ATOMG form, whose round trip to L2 the scatter would then wait for
#define PWS_RED(p, v) tma::red_add_f32((p), (v))
#endif

#ifndef PWS_BWD_UNROLL
#define PWS_BWD_UNROLL 2
#endif
constexpr int kRowUnroll = PWS_BWD_UNROLL;
constexpr int kScouts = 2, kGroups = 2, kGroupWarps = 8, kConsumers = kGroups * kGroupWarps;
constexpr int kThreads = (kScouts + kConsumers + 1) * 32;  // scouts, consumers, the warp that zero-fills grad_input
constexpr int kZeroWarp = kScouts + kConsumers;
static_assert(kScouts == kGroups, "scout w feeds consumer group w (and tells it when the tiles have run out)");
constexpr int kInfoStop = 1 << 11;  // info.z: no more tiles for this consumer group

// grad_input is zero-filled INSIDE the kernel, one frame at a time, two frames ahead of the scatter: the
// zeroed lines are still in L2 when the REDs land and no separate memset pass runs ahead of the kernel.
// A launch owns one slot of per-frame completion counters; the last CTA to leave resets the slot.
constexpr int kSyncSlots = 64, kSyncFrames = 256;
// The zero-fill is tracked in bands of 1/8 frame (rows [ceil(b*H/8), ceil((b+1)*H/8)) of every channel plane): a tile
// only needs the bands its taps can reach.  (Measured: with band tracking the best look-ahead is still 3-4 bands --
// 0.593 / 0.482 / 0.469 / 0.468 ms at 1 / 2 / 3 / 4 -- so the finer grain buys robustness, not speed.)
constexpr int kBands = 8;
__device__ unsigned int g_zero_done[kSyncSlots][kSyncFrames * kBands];
__device__ unsigned int g_exit_count[kSyncSlots];
// Tiles are handed out dynamically: the SMs of a B200 do not run this kernel at the same pace (the spread is several
// per cent: distance to the L2 slices, neighbours on the same TPC), and with a static round-robin every frame ended
// with the fast CTAs waiting at the zero-fill counter of the next frame for the slow ones.
__device__ unsigned int g_tile_next[kSyncSlots];
// one slot sequence for every instantiation of the kernel: they all share the counters above
std::atomic<unsigned> g_next_slot{0};

// the scouts' progress words: a flag polled by the zero-fill warp, not data
#ifdef PWS_BWD_ATOMIC_PROGRESS   // shared-memory atomics: silences compute-sanitizer's racecheck (used for the sanitizer runs)
__device__ __forceinline__ void progress_store(int *p, int v) { atomicExch(p, v); }
__device__ __forceinline__ int progress_load(int *p) { return atomicOr(p, 0); }
#else
__device__ __forceinline__ void progress_store(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
__device__ __forceinline__ int progress_load(int *p) { return *reinterpret_cast<volatile int *>(p); }
#endif
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr int kStripRows = 2 * kTH / kGroupWarps;  // a warp owns a 32-column strip of this many rows
constexpr int kQueueCap = 64;  // entries per consumer warp: 31 left over + one class of stragglers from 32 lanes

// ring depth and the largest box shape a ring slot holds (tiles that need a larger box gather from global memory)
#ifndef PWS_BWD_STAGES
#define PWS_BWD_STAGES 4
#endif
#ifndef PWS_BWD_SLOT_SHAPE
#define PWS_BWD_SLOT_SHAPE 2
#endif
constexpr int kSlotShape = PWS_BWD_SLOT_SHAPE;

// Two rings.  A stage of the main ring holds a tile's grad_output, its frame box and its info words behind one `full`
// and one `empty` mbarrier.  The map tiles live in a ring of their own that is two slots deeper: a scout loads the
// map of its NEXT tile while it dispatches the current one, so that the bounding box of a tile is known before the
// tile's stage frees up and the grad_output and box loads leave the moment it does.
template <int CS, bool kGgrid> struct Smem {
    static constexpr int kStages = kGgrid ? PWS_BWD_STAGES : 6;
    static constexpr int kMapStages = kStages + kScouts;
    static constexpr int kGoutBytes = CS * kTW * kTH * 4;
    static constexpr int kBoxBytes = kGgrid ? (box_w(kSlotShape) * box_h(kSlotShape) * CS * 4 + 127) / 128 * 128 : 0;
    static constexpr int kMapOff = 0;
    static constexpr int kGoutOff = kMapStages * kMapTileBytes;
    static constexpr int kBoxOff = kGoutOff + kStages * kGoutBytes;
    static constexpr int kQueueOff = kBoxOff + kStages * kBoxBytes;
    static constexpr int kQueueEntry = CS == 3 ? 16 : 8;
    static constexpr int kInfoOff = kQueueOff + kConsumers * kQueueCap * kQueueEntry;
    static constexpr int kBarOff = kInfoOff + kStages * 32;
    static constexpr int kProgressOff = kBarOff + 2 * (kStages + kMapStages) * 8;
    static constexpr int kTotal = kProgressOff + 16;
    static_assert(kGoutBytes % 128 == 0 && kBoxBytes % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

struct TmaParams {
    CUtensorMap map;              // as in warp_fwd_tma.cu
    CUtensorMap gout;             // (Wo, Ho, C, N), box (64,16,CS,1)
    CUtensorMap box[kNumShapes];  // frame (W, H, C, N), box (BW, BH, CS, 1)
};

template <int CS>
struct Carry {
    int x, y;  // target of the parked south-west sums
    bool live;
    float v[CS];
};

// Straggler queue of a consumer warp: one entry = a source pixel offset plus its CS channel values.
template <int CS> struct QEntry;
template <> struct QEntry<1> { int off; float v[1]; };
template <> struct QEntry<3> { int off; float v[3]; };
template <int CS>
struct Queue {
    QEntry<CS> *buf;  // shared memory, kQueueCap entries
    int count;        // warp-uniform
};

template <int CS>
__device__ __forceinline__ void queue_put(Queue<CS> &q, int pos, int off, const float (&v)[CS])
{
    if (CS == 3) *reinterpret_cast<int4 *>(q.buf + pos) = make_int4(off, __float_as_int(v[0]), __float_as_int(v[CS > 1 ? 1 : 0]), __float_as_int(v[CS > 2 ? 2 : 0]));
    else *reinterpret_cast<int2 *>(q.buf + pos) = make_int2(off, __float_as_int(v[0]));
}
template <int CS>
__device__ __forceinline__ void queue_pop_red(const Queue<CS> &q, int pos, float *const gip0, const int64_t gs1, const uint64_t pol_last)
{
    if (CS == 3) {
        const int4 e = *reinterpret_cast<const int4 *>(q.buf + pos);
        float *const p = gip0 + e.x;  // the three channel planes of one source pixel
        PWS_RED(p, __int_as_float(e.y));
        PWS_RED(p + gs1, __int_as_float(e.z));
        PWS_RED(p + 2 * gs1, __int_as_float(e.w));
    } else {
        const int2 e = *reinterpret_cast<const int2 *>(q.buf + pos);
        PWS_RED(gip0 + e.x, __int_as_float(e.y));
    }
}
// dense 32-lane REDs while at least a warp's worth of entries is queued
template <int CS>
__device__ __forceinline__ void queue_drain(Queue<CS> &q, float *const gip0, const int64_t gs1, int lane, const uint64_t pol_last)
{
    __syncwarp();
#pragma unroll 1
    while (q.count >= 32) {
        queue_pop_red<CS>(q, q.count - 32 + lane, gip0, gs1, pol_last);
        q.count -= 32;
    }
    __syncwarp();
}
template <int CS>
__device__ __forceinline__ void queue_flush(Queue<CS> &q, float *const gip0, const int64_t gs1, int lane, const uint64_t pol_last)
{
    queue_drain<CS>(q, gip0, gs1, lane, pol_last);
    if (lane < q.count) queue_pop_red<CS>(q, lane, gip0, gs1, pol_last);
    q.count = 0;
    __syncwarp();
}

// One output row (32 pixels) of a warp.  kMasked=false: every lane is a real pixel with 4 valid taps.
// kBoxTaps: the taps of grad_grid come from the shared-memory box (tap = box[(y - by) * pitch + (x - bx)] per plane).
template <int CS, bool kGin, bool kGgrid, bool kMasked, bool kBoxTaps>
__device__ __forceinline__ void bwd_row(
    const int lane, const bool px_ok, const unsigned live,
    const float ix, const float iy, const float x0f, const float y0f, const int x0, const int y0,
    const float gxm, const float gym, const float (&go)[CS],
    const float *__restrict__ box, const int pitch, const int plane,
    const float *__restrict__ ip, const int sH, const int i_ch, const int H, const int W,
    float *const gip0, const int64_t gs1,
    float *__restrict__ ggq, const int gg_s3, Carry<CS> &cy, Queue<CS> &q, const uint64_t pol_last, const uint64_t pol_first)
{
    const float dw = fsub(x0f + 1.0f, ix), de = fsub(ix, x0f), dn = fsub(y0f + 1.0f, iy), ds = fsub(iy, y0f);
    unsigned mask = 15u;
    if (kMasked) {
        const bool xw = (unsigned)x0 < (unsigned)W, xe = (unsigned)(x0 + 1) < (unsigned)W;
        const bool yn = (unsigned)y0 < (unsigned)H, ys = (unsigned)(y0 + 1) < (unsigned)H;
        mask = ((xw && yn) ? 1u : 0u) | ((xe && yn) ? 2u : 0u) | ((xw && ys) ? 4u : 0u) | ((xe && ys) ? 8u : 0u);
        if (!px_ok) mask = 0u;
    }

    if (kGgrid && (!kMasked || px_ok)) {
        float gix = 0.f, giy = 0.f;
        const float *__restrict__ p0 = kBoxTaps ? box + (y0 * pitch + x0) : ip + (y0 * sH + x0);
        const int row = kBoxTaps ? pitch : sH, ch = kBoxTaps ? plane : i_ch;
#pragma unroll
        for (int k = 0; k < CS; ++k) {
            const float *__restrict__ pc = p0 + k * ch;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            if (kBoxTaps) {
                if (!kMasked || (mask & 1u)) v0 = pc[0];
                if (!kMasked || (mask & 2u)) v1 = pc[1];
                if (!kMasked || (mask & 4u)) v2 = pc[row];
                if (!kMasked || (mask & 8u)) v3 = pc[row + 1];
            } else {
                if (!kMasked || (mask & 1u)) v0 = __ldg(pc);
                if (!kMasked || (mask & 2u)) v1 = __ldg(pc + 1);
                if (!kMasked || (mask & 4u)) v2 = __ldg(pc + row);
                if (!kMasked || (mask & 8u)) v3 = __ldg(pc + row + 1);
            }
            // ATen's statement order: t = v*d rounded, then one fma with gOut
            if (!kMasked || (mask & 1u)) { gix = ffma(-fmul(v0, dn), go[k], gix); giy = ffma(-fmul(v0, dw), go[k], giy); }
            if (!kMasked || (mask & 2u)) { gix = ffma(fmul(v1, dn), go[k], gix);  giy = ffma(-fmul(v1, de), go[k], giy); }
            if (!kMasked || (mask & 4u)) { gix = ffma(-fmul(v2, ds), go[k], gix); giy = ffma(fmul(v2, dw), go[k], giy); }
            if (!kMasked || (mask & 8u)) { gix = ffma(fmul(v3, ds), go[k], gix);  giy = ffma(fmul(v3, de), go[k], giy); }
        }
        gix = fmul(gxm, gix); giy = fmul(gym, giy);
        tma::st_f32_hint(ggq, gix, pol_first); tma::st_f32_hint(ggq + gg_s3, giy, pol_first);
    }

    if (kGin) {
        const float nw = fmul(dw, dn), ne = fmul(de, dn), sw = fmul(dw, ds), se = fmul(de, ds);
        const int o_nw = y0 * W + x0;  // grad_input is dense NCHW
        bool take, given;
#ifndef PWS_BWD_SHFL4
        if (!kMasked) {
            // every tap is inside the frame, so x0 <= W - 2 and "left neighbour's offset + 1 == mine" can only mean the
            // same row: one shuffle and one vote instead of four shuffles (lane l gives iff lane l + 1 takes)
            const int o_left = __shfl_up_sync(0xffffffffu, o_nw, 1);
            take = lane > 0 && o_left + 1 == o_nw;
            given = ((__ballot_sync(0xffffffffu, take) >> 1) >> lane) & 1u;
        } else
#endif
        {
            const int px0 = __shfl_up_sync(0xffffffffu, x0, 1), py0 = __shfl_up_sync(0xffffffffu, y0, 1);
            const int nx0 = __shfl_down_sync(0xffffffffu, x0, 1), ny0 = __shfl_down_sync(0xffffffffu, y0, 1);
            take = lane > 0 && px0 + 1 == x0 && py0 == y0;
            given = lane < 31 && nx0 == x0 + 1 && ny0 == y0;
            if (kMasked) {
                take = take && px_ok && ((live >> (lane - 1)) & 1u);
                given = given && px_ok && ((live >> (lane + 1)) & 1u);
            }
        }
        const bool chain = cy.live && cy.x == x0 && cy.y == y0;
        const int o_cy = cy.y * W + cy.x;
        // stragglers of this row -> queue: east taps nobody takes over (top and bottom), parked sums whose chain broke
        const bool p_e1 = !given && (mask & 2u), p_e2 = !given && (mask & 8u), p_f = cy.live && !chain;
        const unsigned lt = (1u << lane) - 1u;
        float etop[CS], ebot[CS], brk[CS];
#pragma unroll
        for (int k = 0; k < CS; ++k) {
            float top = fmul(nw, go[k]), bot = fmul(sw, go[k]);
            etop[k] = fmul(ne, go[k]); ebot[k] = fmul(se, go[k]);
            const float ptop = __shfl_up_sync(0xffffffffu, etop[k], 1), pbot = __shfl_up_sync(0xffffffffu, ebot[k], 1);
            if (take) { top += ptop; bot += pbot; }
            brk[k] = cy.v[k];
            if (chain) top += cy.v[k];
            if (mask & 1u) PWS_RED(gip0 + o_nw + k * gs1, top);
            cy.v[k] = bot;
        }
#ifndef PWS_EXP_NOQUEUE   // experiment: drop the stragglers (results are wrong) to time the queue's share of a row
        if (!kMasked) {
            // all taps valid: the two east classes share their predicate
            const unsigned b = __ballot_sync(0xffffffffu, p_e1);
            if (b) {
                const int n = __popc(b), rank = __popc(b & lt);
                if (p_e1) queue_put<CS>(q, q.count + rank, o_nw + 1, etop);
                q.count += n;
                if (q.count >= 32) queue_drain<CS>(q, gip0, gs1, lane, pol_last);
                if (p_e1) queue_put<CS>(q, q.count + rank, o_nw + W + 1, ebot);
                q.count += n;
                if (q.count >= 32) queue_drain<CS>(q, gip0, gs1, lane, pol_last);
            }
        } else {
            const unsigned b1 = __ballot_sync(0xffffffffu, p_e1), b2 = __ballot_sync(0xffffffffu, p_e2);
            if (b1) {
                if (p_e1) queue_put<CS>(q, q.count + __popc(b1 & lt), o_nw + 1, etop);
                q.count += __popc(b1);
                if (q.count >= 32) queue_drain<CS>(q, gip0, gs1, lane, pol_last);
            }
            if (b2) {
                if (p_e2) queue_put<CS>(q, q.count + __popc(b2 & lt), o_nw + W + 1, ebot);
                q.count += __popc(b2);
                if (q.count >= 32) queue_drain<CS>(q, gip0, gs1, lane, pol_last);
            }
        }
        {
            const unsigned b = __ballot_sync(0xffffffffu, p_f);
            if (b) {
                const int pos = q.count + __popc(b & lt);
                if (p_f) queue_put<CS>(q, pos, o_cy, brk);
                q.count += __popc(b);
                if (q.count >= 32) queue_drain<CS>(q, gip0, gs1, lane, pol_last);
            }
        }
#endif
        cy.x = x0; cy.y = y0 + 1;
        cy.live = (mask & 4u) != 0u;
    }
}

// A strip of a tile that is not "interior" (frame border, partial tile, fallback): the masked bodies.
template <int CS, bool kBorder, bool kAlign, bool kInter, bool kGin, bool kGgrid>
__device__ __forceinline__ void masked_strip(
    const int lane, const int4 info, const int h0, const int w0, const int row0, const int col0,
    const float *__restrict__ mp, const float *__restrict__ gop, const float *__restrict__ bp, const int pitch, const int plane,
    const float *__restrict__ ip, const int sH, const int i_ch, const Geometry g,
    float *const gip0, const int64_t gs1, float *__restrict__ ggq, const int gg_s1, const int gg_s3, Carry<CS> &cy, Queue<CS> &q,
    const uint64_t pol_last, const uint64_t pol_first)
{
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
    const bool col_ok = w0 + lane < g.Wo;
    const unsigned live = __ballot_sync(0xffffffffu, col_ok);
    const bool box_taps = kGgrid && !(info.z & (kInfoFallback | kInfoEmpty));
    const int rows = min(kStripRows, g.Ho - h0);  // may be <= 0 for the strips below the last row
    for (int r = 0; r < rows; ++r) {
        float gx, gy, go[CS];
        if (kInter) { const float2 v = *reinterpret_cast<const float2 *>(mp + (row0 + r) * (2 * kTW) + 2 * (col0 + lane)); gx = v.x; gy = v.y; }
        else { gx = mp[(row0 + r) * kTW + col0 + lane]; gy = mp[kTW * kTH + (row0 + r) * kTW + col0 + lane]; }
#pragma unroll
        for (int k = 0; k < CS; ++k) go[k] = gop[k * (kTW * kTH) + r * kTW];
        float gxm, gym;
        const float ix = src_index_grad<kBorder, kAlign>(gx, Wf, Wm1, &gxm);
        const float iy = src_index_grad<kBorder, kAlign>(gy, Hf, Hm1, &gym);
        const float x0f = floorf(ix), y0f = floorf(iy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        if (box_taps)
            bwd_row<CS, kGin, kGgrid, true, true>(lane, col_ok, live, ix, iy, x0f, y0f, x0, y0, gxm, gym, go,
                                                   bp, pitch, plane, ip, sH, i_ch, g.H, g.W, gip0, gs1, ggq, gg_s3, cy, q, pol_last, pol_first);
        else
            bwd_row<CS, kGin, kGgrid, true, false>(lane, col_ok, live, ix, iy, x0f, y0f, x0, y0, gxm, gym, go,
                                                    bp, pitch, plane, ip, sH, i_ch, g.H, g.W, gip0, gs1, ggq, gg_s3, cy, q, pol_last, pol_first);
        if (kGgrid) ggq += gg_s1;
    }
}

template <int CS, bool kBorder, bool kAlign, bool kInter, bool kGin, bool kGgrid>
__global__ void __launch_bounds__(kThreads, 1)
bwd_tma_kernel(const __grid_constant__ TmaParams tp, const View in, const View grid, const View gin, const View ggrid, const Geometry g,
               const int tiles_x, const int tiles_y, const int total_tiles, const int n_begin, const int n_frames, const int slot, const int zero_ahead)
{
    using S = Smem<CS, kGgrid>;
    constexpr int kStages = S::kStages, kMapStages = S::kMapStages;
    extern __shared__ __align__(1024) unsigned char smem[];
    float *const s_map = reinterpret_cast<float *>(smem + S::kMapOff);
    unsigned char *const s_gout = smem + S::kGoutOff;
    unsigned char *const s_box = smem + S::kBoxOff;
    QEntry<CS> *const s_queue = reinterpret_cast<QEntry<CS> *>(smem + S::kQueueOff);
    int4 *const s_info = reinterpret_cast<int4 *>(smem + S::kInfoOff);
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + S::kBarOff);
    uint64_t *const empty = full + kStages;
    uint64_t *const map_full = empty + kStages;
    uint64_t *const map_empty = map_full + kMapStages;
    // [0], [1] per scout: how far it has dispatched, in eighths of a frame; [2]: the last band (8 * frame + band) known
    // to be zero-filled by every CTA, published by the zero-fill warp.  Written by the scouts and polled by the zero-fill
    // warp (volatile; compute-sanitizer's racecheck flags exactly this pair and nothing else -- build with
    // -DPWS_BWD_ATOMIC_PROGRESS to run it clean)
    int *const s_progress = reinterpret_cast<int *>(smem + S::kProgressOff);

    // Roles are numbered from the TOP warp of the CTA down (the scheduler favours the higher warp ids when several
    // warps are ready; the scouts are a handful of instructions per tile, but every consumer waits on them).
    const int warp = (kThreads / 32 - 1) - (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int tiles_xy = tiles_x * tiles_y;
    const float Wf = (float)g.W, Hf = (float)g.H, Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
    // L2 priorities: everything that streams through once (map, grad_output, frame boxes, grad_grid) is evict_first,
    // so the replacement order favours grad_input, which is revisited between its zero-fill and its last RED.
    // (evict_last for grad_input was tried: its lines then outlive the kernel and crowd the next one out of L2.)
    const uint64_t pol_first = tma::policy_evict_first();
    const uint64_t pol_last = tma::policy_evict_normal();
    const uint64_t pol_box = pol_first, pol_gg = pol_first, pol_zero = pol_last;

    if (threadIdx.x == 0) {
        s_progress[0] = 0; s_progress[1] = 0;
#ifdef PWS_BWD_PUBLISH
        s_progress[2] = -1;
#endif
        // full: the scout arrives twice -- once with the byte count of the TMA loads, once when the tile's frame is known
        // to be zero-filled by every CTA; empty: one arrival per consumer warp of the group
        for (int s = 0; s < kStages; ++s) { tma::mbar_init(full + s, 2); tma::mbar_init(empty + s, kGroupWarps); }
        for (int s = 0; s < kMapStages; ++s) { tma::mbar_init(map_full + s, 1); tma::mbar_init(map_empty + s, kGroupWarps); }
        tma::fence_barrier_init();
    }
    __syncthreads();

    if (warp < kScouts) {
        // ===== scouts: scout w serves the iterations it = w, w + 2, ... = consumer group w =====
        // Per tile: fetch the tile index, start the map load of the NEXT tile, reduce this tile's map (already in the map
        // ring) to the bounding box of its taps, wait for the ring stage, issue the grad_output and frame box loads on the
        // stage's barrier, and make sure the tile's frame has been zero-filled by every CTA before giving the stage its
        // second arrival.  The consumers do none of this: between two tiles they wait on
        // one mbarrier.
        if (lane == 0) {
            tma::prefetch_desc(&tp.map); tma::prefetch_desc(&tp.gout);
            if (kGgrid) { tma::prefetch_desc(&tp.box[0]); tma::prefetch_desc(&tp.box[1]); tma::prefetch_desc(&tp.box[2]); }
        }
        // The first 2 * kScouts tiles of a CTA are static, the rest comes from the launch's counter.  The scout works
        // on tile `t` while the map of its next tile `t1` is being loaded into the map ring and the index of the one
        // after is being fetched.
        const int stride = kScouts * (int)gridDim.x;
        int t = blockIdx.x + warp * gridDim.x, t1 = t + stride;
        // map load of iteration `i` (tile coordinates c) into its slot of the map ring; lane 0 only
        auto load_map = [&](int i, const TileCoord &c) {
            const int ms = i % kMapStages;
            tma::mbar_arrive_expect_tx(map_full + ms, kMapTileBytes);
            if (kInter) tma::load_3d_hint(s_map + ms * kMapTileFloats, &tp.map, map_full + ms, 2 * c.w0, c.h0, n_begin + c.n, pol_first);
            else tma::load_4d_hint(s_map + ms * kMapTileFloats, &tp.map, map_full + ms, c.w0, c.h0, 0, n_begin + c.n, pol_first);
        };
        TileCoord tc = tile_coord(min(t, total_tiles - 1), tiles_x, tiles_xy);
        if (lane == 0 && t < total_tiles) load_map(warp, tc);  // the ring starts out empty
        int zero_seen = -1;  // bands [0, zero_seen] (index = 8 * frame + band) are known to be zero-filled by every CTA
#ifdef PWS_EXP_CLOCKS
        long long s_range = 0, s_wait = 0, s_issue = 0, s_zero = 0; int s_nowait = 0, s_n = 0;
#endif
        for (int it = warp;; it += kScouts) {
            const int st = it % kStages, ph = (it / kStages) & 1;
            const int ms = it % kMapStages, mph = (it / kMapStages) & 1;
            const int ms1 = (it + kScouts) % kMapStages, mph1 = ((it + kScouts) / kMapStages) & 1;
#ifdef PWS_EXP_CLOCKS
            const long long q0 = clock64();
#endif
            if (t >= total_tiles) {
                tma::mbar_wait_relaxed(empty + st, ph ^ 1);
                if (lane == 0) {
                    s_info[2 * st] = make_int4(0, 0, kInfoStop, 0);
                    tma::mbar_arrive(full + st); tma::mbar_arrive(full + st);
                    progress_store(&s_progress[warp], INT_MAX - 4);  // out of tiles: let the zero-fill warp run to the end
                }
                break;
            }
            int t2 = 0;
            if (lane == 0) t2 = (int)atomicAdd(&g_tile_next[slot], 1u) + 2 * stride;  // the tile after the next
            const TileCoord tc1 = tile_coord(min(t1, total_tiles - 1), tiles_x, tiles_xy);
            // next tile's map: its slot was last used kMapStages iterations ago and is normally free by now
            bool next_loaded = t1 >= total_tiles;
            if (!next_loaded && __shfl_sync(0xffffffffu, (int)tma::mbar_test(map_empty + ms1, mph1 ^ 1), 0)) {
                if (lane == 0) load_map(it + kScouts, tc1);
                next_loaded = true;
            }
            // this tile: bounding box of its taps from the map tile in shared memory
            const int cols = min(kTW, g.Wo - tc.w0), rows = min(kTH, g.Ho - tc.h0);
            tma::mbar_wait_relaxed(map_full + ms, mph);
            float xlo, xhi, ylo, yhi;
            map_tile_range<kInter>(s_map + ms * kMapTileFloats, rows, cols, lane, xlo, xhi, ylo, yhi);
            int4 info = box_of_range<kBorder, kAlign>(xlo, xhi, ylo, yhi, g.W, g.H, cols == kTW && rows == kTH);
            if (kGgrid && kSlotShape < kNumShapes - 1 && !(info.z & (kInfoFallback | kInfoEmpty)) && (info.z & 0xff) > kSlotShape)
                info = make_int4(0, 0, kInfoFallback, 0);
            info.w = tc.n;
            const int progress = 8 * tc.n + (8 * (t - tc.n * tiles_xy)) / tiles_xy;  // in eighths of a frame
            const bool want_box = kGgrid && !(info.z & (kInfoFallback | kInfoEmpty));
            const int shape = info.z & 0xff;
#ifdef PWS_EXP_CLOCKS
            const long long q1 = clock64();
#endif
            tma::mbar_wait_relaxed(empty + st, ph ^ 1);
#ifdef PWS_EXP_CLOCKS
            const long long q2 = clock64();
            long long q3 = q2;
#endif
            if (lane == 0) {
                s_info[2 * st] = info;
#ifdef PWS_EXP_CLOCKS
                s_info[2 * st + 1] = make_int4(tc.h0, tc.w0, (int)clock64(), 0);
#else
                s_info[2 * st + 1] = make_int4(tc.h0, tc.w0, 0, 0);
#endif
                tma::mbar_arrive_expect_tx(full + st, S::kGoutBytes + (want_box ? box_w(shape) * box_h(shape) * CS * 4 : 0));
                tma::load_4d_hint(s_gout + (size_t)st * S::kGoutBytes, &tp.gout, full + st, tc.w0, tc.h0, 0, n_begin + tc.n, pol_first);
                if (want_box) tma::load_4d_hint(s_box + (size_t)st * S::kBoxBytes, &tp.box[shape], full + st, info.x, info.y, 0, n_begin + tc.n, pol_box);
#ifdef PWS_EXP_CLOCKS
                q3 = clock64();
#endif
                if (kGin && !(info.z & kInfoEmpty)) {
                    // the last band this tile's REDs can land in: the box rows when the taps were bounded, else the whole frame
                    const int y_hi = (info.z & kInfoFallback) ? g.H - 1 : min(info.y + box_h(shape), g.H) - 1;
                    const int need = kBands * tc.n + (kBands * y_hi) / g.H;
                    // the zero-fill warp keeps `zero_ahead` bands ahead of what the scouts publish: the output position of
                    // the tile, or the band it needs if that is further on (a map that samples far away must not starve)
                    progress_store(&s_progress[warp], max(progress, need));
#ifdef PWS_BWD_PUBLISH   // experiment (tools/exp/README.md): never run on a GPU in this form
                    if (need > zero_seen) {
                        // the zero-fill warp of this CTA watches the launch's counters and publishes how far every CTA
                        // has got: the scouts never poll global memory themselves
                        while ((zero_seen = progress_load(&s_progress[2])) < need) __nanosleep(64);
                        __threadfence_block();
                    }
#else
                    if (need > zero_seen) {  // bands complete in order: every CTA fills them in order
                        while (ld_acquire(&g_zero_done[slot][need]) < gridDim.x) __nanosleep(64);
                        zero_seen = need;
                    }
#endif
                } else if (kGin) {
                    progress_store(&s_progress[warp], progress);
                }
                // second arrival: passes on (cta scope) the acquire above and this scout's view of the map tile
                tma::mbar_arrive(full + st);
            }
            if (!next_loaded) {
                tma::mbar_wait_relaxed(map_empty + ms1, mph1 ^ 1);
                if (lane == 0) load_map(it + kScouts, tc1);
            }
#ifdef PWS_EXP_CLOCKS
            if (lane == 0) { s_range += q1 - q0; s_wait += q2 - q1; s_issue += q3 - q2; s_zero += clock64() - q3; s_nowait += (q2 - q1 < 200); ++s_n; }
#endif
            t = t1; tc = tc1;
            t1 = __shfl_sync(0xffffffffu, t2, 0);
        }
#ifdef PWS_EXP_CLOCKS
        if (lane == 0 && (blockIdx.x % 49) == 0)
            printf("scout cta %3d w %d: tiles %3d range %8lld wait_empty %8lld (no wait: %3d) issue %8lld zero %8lld\n", blockIdx.x, warp, s_n, s_range, s_wait, s_nowait, s_issue, s_zero);
#endif
    } else if (warp == kZeroWarp) {
        // ===== zero-fill of grad_input, band by band, this CTA's 1/gridDim share of every band, `zero_ahead` bands ahead =====
        if (kGin) {
            const int64_t plane = (int64_t)g.H * g.W;  // dense NCHW frame (host-checked), W % 4 == 0, base 16-byte aligned
#ifdef PWS_EXP_CLOCKS
            long long z_trig = 0, z_st = 0, z_fence = 0;
#endif
            const int last = n_frames * kBands - 1;
#ifdef PWS_BWD_PUBLISH
            // bands complete in order (every CTA fills them in order), so one acquire load per band tells how far all CTAs
            // have got; the acquire, the block-scope fences around the shared-memory word and the scouts' mbarrier
            // arrivals carry the ordering on to the consumers' REDs
            int pub = -1;
            auto advance = [&]() {
                if (lane == 0) {
                    const int before = pub;
                    while (pub < last && ld_acquire(&g_zero_done[slot][pub + 1]) >= gridDim.x) ++pub;
                    if (pub != before) { __threadfence_block(); progress_store(&s_progress[2], pub); }
                }
                __syncwarp();
            };
#endif
            for (int idx = 0; idx <= last; ++idx) {
                const int f = idx / kBands, b = idx % kBands;
#ifdef PWS_EXP_CLOCKS
                const long long z0 = clock64();
#endif
#ifdef PWS_BWD_PUBLISH
                // lane 0 decides, the warp follows: the condition reads volatile words, and lanes that left the loop at
                // different iterations would meet the __syncwarp() inside advance() from different program points
                for (;;) {
                    int go_on = 0;
                    if (lane == 0) go_on = max(progress_load(&s_progress[0]), progress_load(&s_progress[1])) + zero_ahead < idx;
                    if (!__shfl_sync(0xffffffffu, go_on, 0)) break;
                    advance();
                    __nanosleep(256);
                }
#else
                while (max(progress_load(&s_progress[0]), progress_load(&s_progress[1])) + zero_ahead < idx) __nanosleep(256);
#endif
#ifdef PWS_EXP_CLOCKS
                const long long z1 = clock64();
#endif
                const int r0 = (b * g.H + kBands - 1) / kBands, r1 = ((b + 1) * g.H + kBands - 1) / kBands;
                const int band_vec = (r1 - r0) * (g.W / 4);  // float4 per channel plane
                const int share = (band_vec + gridDim.x - 1) / gridDim.x;
                const int v0 = blockIdx.x * share, v1 = min(band_vec, v0 + share);
                float *const fp = (float *)gin.p + (int64_t)(n_begin + f) * gin.sN + (int64_t)r0 * g.W;
                for (int c = 0; c < CS; ++c) {
                    float4 *__restrict__ dst = reinterpret_cast<float4 *>(fp + c * plane);
                    for (int v = v0 + lane; v < v1; v += 32) tma::st_zero_v4_hint(dst + v, pol_zero);
                }
#ifdef PWS_EXP_CLOCKS
                const long long z2 = clock64();
#endif
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(&g_zero_done[slot][idx], 1u);
#ifdef PWS_BWD_PUBLISH
                advance();
#endif
#ifdef PWS_EXP_CLOCKS
                z_trig += z1 - z0; z_st += z2 - z1; z_fence += clock64() - z2;
#endif
            }
#ifdef PWS_BWD_PUBLISH
            while (__shfl_sync(0xffffffffu, pub, 0) < last) { advance(); __nanosleep(128); }
#endif
#ifdef PWS_EXP_CLOCKS
            if (lane == 0 && (blockIdx.x % 49) == 0) printf("zclk cta %3d: trig %8lld store %8lld fence %8lld\n", blockIdx.x, z_trig, z_st, z_fence);
#endif
        }
    } else {
        // ===== consumers: group `grp` owns the iterations it = grp, grp + 2, ...; a warp owns a 32 x 4 strip of the tile =====
        const int cw = warp - kScouts, grp = cw / kGroupWarps, wg = cw % kGroupWarps;
        const int col0 = (wg & 1) * 32, row0 = (wg >> 1) * kStripRows;
        Queue<CS> q;
        q.buf = s_queue + cw * kQueueCap;
        q.count = 0;
        const float gxm_in = (kAlign ? Wm1 : Wf) * 0.5f, gym_in = (kAlign ? Hm1 : Hf) * 0.5f;
#ifdef PWS_EXP_CLOCKS
        long long c_wait = 0, c_bar = 0, c_body = 0, lat_late = 0, slack_late = 0, slack_ok = 0;
        int n_late = 0, n_ok = 0;
#endif
        for (int it = grp;; it += kGroups) {
            const int is = it % kStages, bs = is, ph = (it / kStages) & 1;
#ifdef PWS_EXP_CLOCKS
            const long long k2 = clock64();
#endif
            // one warp of the group waits on the stage's mbarrier, the others park on a hardware barrier and spin on nothing
#ifdef PWS_BWD_NOBAR
            tma::mbar_wait(full + is, ph);
#else
            if (wg == 0) tma::mbar_wait(full + is, ph);
#endif
#ifdef PWS_EXP_CLOCKS
            const long long k3 = clock64();
#endif
#ifndef PWS_BWD_NOBAR
            tma::named_bar_sync(1 + grp, kGroupWarps * 32);
#endif
#ifdef PWS_EXP_CLOCKS
            const long long k4 = clock64();
            c_wait += k3 - k2; c_bar += k4 - k3;
            if (wg == 0 && !(s_info[2 * bs].z & kInfoStop)) {
                const int issue = s_info[2 * bs + 1].z;
                if (k3 - k2 > 300) { ++n_late; lat_late += (int)k3 - issue; slack_late += (int)k2 - issue; }
                else { ++n_ok; slack_ok += (int)k2 - issue; }
            }
#endif
            const int4 info = s_info[2 * bs], where = s_info[2 * bs + 1];
            if (info.z & kInfoStop) break;
            const int n = n_begin + info.w, h0 = where.x + row0, w0 = where.y + col0;
            const int shape = info.z & 0xff;
            const int pitch = box_w(shape), plane = box_w(shape) * box_h(shape);
            const int ms = it % kMapStages;
            const float *mp = s_map + ms * kMapTileFloats;
            const float *gop = reinterpret_cast<const float *>(s_gout + (size_t)is * S::kGoutBytes) + row0 * kTW + col0 + lane;
            const float *bp = reinterpret_cast<const float *>(s_box + (size_t)bs * S::kBoxBytes) - (info.y * pitch + info.x);
            const float *__restrict__ ip = (const float *)in.p + (int64_t)n * in.sN;
            float *const gip0 = kGin ? (float *)gin.p + (int64_t)n * gin.sN : nullptr;  // this frame's grad_input (dense NCHW)
            const int64_t gs1 = gin.s1;                                               // its channel stride
            float *__restrict__ ggq = kGgrid ? (float *)ggrid.p + (int64_t)n * ggrid.sN + (int64_t)h0 * ggrid.s1 + (int64_t)(w0 + lane) * ggrid.s2 : nullptr;

            Carry<CS> cy;
            cy.x = 0; cy.y = 0; cy.live = false;
#pragma unroll
            for (int k = 0; k < CS; ++k) cy.v[k] = 0.f;

            if (info.z & kInfoInterior) {
#if defined(PWS_BWD_SWP)
                // software-pipelined rows (experiment): the grad_grid part of row r + 1 -- loads, coordinates, taps, no
                // branches -- is written ahead of the scatter part of row r, so that the two share a basic block and
                // the scheduler can fill the scatter's shuffle / vote latencies with it
                float c_ix, c_iy, c_x0f, c_y0f, c_go[CS]; int c_x0, c_y0;
                auto load_row = [&](int r, float &ix, float &iy, float &x0f, float &y0f, int &x0, int &y0, float (&go)[CS]) {
                    float gx, gy;
                    if (kInter) { const float2 v = *reinterpret_cast<const float2 *>(mp + (row0 + r) * (2 * kTW) + 2 * (col0 + lane)); gx = v.x; gy = v.y; }
                    else { gx = mp[(row0 + r) * kTW + col0 + lane]; gy = mp[kTW * kTH + (row0 + r) * kTW + col0 + lane]; }
#pragma unroll
                    for (int k = 0; k < CS; ++k) go[k] = gop[k * (kTW * kTH) + r * kTW];
                    ix = unnorm<kAlign>(gx, Wf, Wm1); iy = unnorm<kAlign>(gy, Hf, Hm1);
                    floor_small(ix, x0f, x0); floor_small(iy, y0f, y0);
                };
                load_row(0, c_ix, c_iy, c_x0f, c_y0f, c_x0, c_y0, c_go);
                if (kGgrid) {
                    bwd_row<CS, false, true, false, true>(lane, true, 0xffffffffu, c_ix, c_iy, c_x0f, c_y0f, c_x0, c_y0, gxm_in, gym_in, c_go,
                                                          bp, pitch, plane, ip, in.s2, in.s1, g.H, g.W, gip0, gs1, ggq, ggrid.s3, cy, q, pol_last, pol_gg);
                    ggq += ggrid.s1;
                }
#pragma unroll
                for (int r = 0; r < kStripRows; ++r) {
                    float n_ix = 0.f, n_iy = 0.f, n_x0f = 0.f, n_y0f = 0.f, n_go[CS]; int n_x0 = 0, n_y0 = 0;
#pragma unroll
                    for (int k = 0; k < CS; ++k) n_go[k] = 0.f;
                    if (r + 1 < kStripRows) {
                        load_row(r + 1, n_ix, n_iy, n_x0f, n_y0f, n_x0, n_y0, n_go);
                        if (kGgrid) {
                            bwd_row<CS, false, true, false, true>(lane, true, 0xffffffffu, n_ix, n_iy, n_x0f, n_y0f, n_x0, n_y0, gxm_in, gym_in, n_go,
                                                                  bp, pitch, plane, ip, in.s2, in.s1, g.H, g.W, gip0, gs1, ggq, ggrid.s3, cy, q, pol_last, pol_gg);
                            ggq += ggrid.s1;
                        }
                    }
                    if (kGin)
                        bwd_row<CS, true, false, false, true>(lane, true, 0xffffffffu, c_ix, c_iy, c_x0f, c_y0f, c_x0, c_y0, gxm_in, gym_in, c_go,
                                                              bp, pitch, plane, ip, in.s2, in.s1, g.H, g.W, gip0, gs1, ggq, ggrid.s3, cy, q, pol_last, pol_gg);
                    c_ix = n_ix; c_iy = n_iy; c_x0f = n_x0f; c_y0f = n_y0f; c_x0 = n_x0; c_y0 = n_y0;
#pragma unroll
                    for (int k = 0; k < CS; ++k) c_go[k] = n_go[k];
                }
#else
#pragma unroll kRowUnroll
                for (int r = 0; r < kStripRows; ++r) {
                    float gx, gy, go[CS];
                    if (kInter) { const float2 v = *reinterpret_cast<const float2 *>(mp + (row0 + r) * (2 * kTW) + 2 * (col0 + lane)); gx = v.x; gy = v.y; }
                    else { gx = mp[(row0 + r) * kTW + col0 + lane]; gy = mp[kTW * kTH + (row0 + r) * kTW + col0 + lane]; }
#pragma unroll
                    for (int k = 0; k < CS; ++k) go[k] = gop[k * (kTW * kTH) + r * kTW];
                    const float ix = unnorm<kAlign>(gx, Wf, Wm1), iy = unnorm<kAlign>(gy, Hf, Hm1);
                    float x0f, y0f; int x0, y0;
                    floor_small(ix, x0f, x0); floor_small(iy, y0f, y0);
                    bwd_row<CS, kGin, kGgrid, false, true>(lane, true, 0xffffffffu, ix, iy, x0f, y0f, x0, y0, gxm_in, gym_in, go,
                                                            bp, pitch, plane, ip, in.s2, in.s1, g.H, g.W, gip0, gs1, ggq, ggrid.s3, cy, q, pol_last, pol_gg);
                    if (kGgrid) ggq += ggrid.s1;
                }
#endif
            } else {
                masked_strip<CS, kBorder, kAlign, kInter, kGin, kGgrid>(lane, info, h0, w0, row0, col0, mp, gop, bp, pitch, plane, ip, in.s2, in.s1,
                                                                        g, gip0, gs1, ggq, ggrid.s1, ggrid.s3, cy, q, pol_last, pol_gg);
            }
            if (kGin) {
                if (cy.live) {
                    const int o_cy = cy.y * g.W + cy.x;
#pragma unroll
                    for (int k = 0; k < CS; ++k) PWS_RED(gip0 + o_cy + k * gs1, cy.v[k]);
                }
                queue_flush<CS>(q, gip0, gs1, lane, pol_last);
            }
            __syncwarp();
            if (lane == 0) { tma::mbar_arrive(empty + is); tma::mbar_arrive(map_empty + ms); }
#ifdef PWS_EXP_CLOCKS
            c_body += clock64() - k4;
#endif
        }
#ifdef PWS_EXP_CLOCKS
        if (lane == 0 && (blockIdx.x % 49) == 0 && (wg == 0 || wg == 5))
            printf("clk cta %3d grp %d wg %d: wait %8lld bar %8lld body %8lld | late tiles %3d: latency %6lld slack %6lld | on-time tiles %3d: slack %6lld\n", blockIdx.x, grp, wg, c_wait, c_bar, c_body,
                   n_late, n_late ? lat_late / n_late : 0, n_late ? slack_late / n_late : 0, n_ok, n_ok ? slack_ok / n_ok : 0);
#endif
    }
    __syncthreads();
    // the last CTA to leave hands the counter slot back clean
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&g_exit_count[slot], 1u) == gridDim.x - 1) {
            if (kGin) for (int i = 0; i < n_frames * kBands; ++i) g_zero_done[slot][i] = 0u;
            g_tile_next[slot] = 0u;
            g_exit_count[slot] = 0u;
            __threadfence();
        }
    }
}

int zero_ahead()
{
    // how far the zero-fill runs ahead of the scatter, in eighths of a frame
    static const int v = [] { const char *e = std::getenv("PWS_BWD_ZERO_AHEAD"); int a = e ? std::atoi(e) : 4; return a < 1 ? 1 : a; }();
    return v;
}

template <int CS, bool kBorder, bool kAlign, bool kInter, bool kGin, bool kGgrid>
bool launch_k(const TmaParams &tp, const Problem &pb, int tiles_x, int tiles_y, int total, int n0, cudaStream_t st)
{
    auto kern = bwd_tma_kernel<CS, kBorder, kAlign, kInter, kGin, kGgrid>;
    using S = Smem<CS, kGgrid>;
    static bool attr_done = false;  // per instantiation
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr_done = true;
    }
    const int grid = total < sm_count() ? total : sm_count();
    const int n_frames = total / (tiles_x * tiles_y);
    const int slot = (int)(g_next_slot.fetch_add(1u, std::memory_order_relaxed) % kSyncSlots);
    kern<<<grid, kThreads, S::kTotal, st>>>(tp, pb.in, pb.grid, pb.gin, pb.ggrid, pb.g, tiles_x, tiles_y, total, n0, n_frames, slot, zero_ahead());
    note_launch();
    note_kernel("bwd_tma");
    return true;
}

template <int CS, bool kBorder, bool kAlign, bool kInter>
bool launch_mask(const TmaParams &tp, const Problem &pb, int tx, int ty, int total, int n0, cudaStream_t st)
{
    if (pb.want_gin && pb.want_ggrid) return launch_k<CS, kBorder, kAlign, kInter, true, true>(tp, pb, tx, ty, total, n0, st);
    if (pb.want_gin) return launch_k<CS, kBorder, kAlign, kInter, true, false>(tp, pb, tx, ty, total, n0, st);
    return launch_k<CS, kBorder, kAlign, kInter, false, true>(tp, pb, tx, ty, total, n0, st);
}

template <int CS, bool kInter>
bool launch_ba(const TmaParams &tp, const Problem &pb, int tx, int ty, int total, int n0, cudaStream_t st)
{
    const bool border = pb.g.padding == PWS_PAD_BORDER, align = pb.g.align != 0;
    if (border && align) return launch_mask<CS, true, true, kInter>(tp, pb, tx, ty, total, n0, st);
    if (border) return launch_mask<CS, true, false, kInter>(tp, pb, tx, ty, total, n0, st);
    if (align) return launch_mask<CS, false, true, kInter>(tp, pb, tx, ty, total, n0, st);
    return launch_mask<CS, false, false, kInter>(tp, pb, tx, ty, total, n0, st);
}

}  // namespace

// Prepared launch state of an eligible problem (tensor maps are encoded once per call, not per chunk).
struct BwdTmaPlan {
    TmaParams tp;
    bool inter;
    int tiles_x, tiles_y;
};

// Returns a heap plan when the TMA backward can take the problem, nullptr otherwise.
BwdTmaPlan *backward_tma_plan(const Problem &pb)
{
    const Geometry &g = pb.g;
    if (tma_disabled()) return nullptr;
    if (pb.in_dtype != PWS_F32 || pb.grid_dtype != PWS_F32) return nullptr;
    if (g.C != 1 && g.C != 3) return nullptr;
    if (g.W > (1 << 22) || g.H > (1 << 22)) return nullptr;
    if (pb.want_gin && !(pb.gin.s3 == 1 && pb.gin.s2 == g.W && pb.gin.s1 == g.W * g.H)) return nullptr;
    // in-kernel zero-fill writes float4: frame base and frame stride 16-byte aligned
    if (pb.want_gin && ((reinterpret_cast<uintptr_t>(pb.gin.p) & 15) || (pb.gin.sN % 4) || (g.W % 4))) return nullptr;
    if (pb.want_ggrid) {
        // written with plain stores: planar (two scalars) or interleaved (one float2)
        if (pb.ggrid.s3 == 1 && ((reinterpret_cast<uintptr_t>(pb.ggrid.p) & 7) || (pb.ggrid.sN & 1) || (pb.ggrid.s1 & 1) || (pb.ggrid.s2 & 1)))
            return nullptr;
    }
    BwdTmaPlan *pl = new BwdTmaPlan;
    pl->tiles_x = (g.Wo + kTW - 1) / kTW;
    pl->tiles_y = (g.Ho + kTH - 1) / kTH;
    bool ok = encode_map_tma(pb.grid, g, &pl->tp.map, &pl->inter);
    ok = ok && encode_frame_tma(pb.gout, g.Wo, g.Ho, g.C, g.N, kTW, kTH, g.C, &pl->tp.gout);
    for (int s = 0; ok && s < kNumShapes; ++s) ok = encode_frame_tma(pb.in, g.W, g.H, g.C, g.N, box_w(s), box_h(s), g.C, &pl->tp.box[s]);
    if ((int64_t)pl->tiles_x * pl->tiles_y * g.N > INT_MAX) ok = false;
    if (!ok) { delete pl; return nullptr; }
    return pl;
}

void backward_tma_free(BwdTmaPlan *pl) { delete pl; }

int backward_tma_max_frames() { return kSyncFrames; }

// Launch frames [n0, n0+nn), nn <= backward_tma_max_frames().  Zero-fills grad_input of those frames itself.
bool launch_backward_tma(const BwdTmaPlan *pl, const Problem &pb, int n0, int nn, cudaStream_t st)
{
    const int total = pl->tiles_x * pl->tiles_y * nn;
    if (total <= 0) return true;
    static const bool alias = [] { const char *e = std::getenv("PWS_EXP_GIN_ALIAS"); return e && e[0] == '1'; }();
    if (alias) {  // experiment: every frame scatters into frame 0's grad_input (L2-resident footprint; results are wrong)
        Problem p2 = pb;
        p2.gin.sN = 0;
        if (pb.g.C == 3)
            return pl->inter ? launch_ba<3, true>(pl->tp, p2, pl->tiles_x, pl->tiles_y, total, n0, st)
                             : launch_ba<3, false>(pl->tp, p2, pl->tiles_x, pl->tiles_y, total, n0, st);
    }
    if (pb.g.C == 3)
        return pl->inter ? launch_ba<3, true>(pl->tp, pb, pl->tiles_x, pl->tiles_y, total, n0, st)
                         : launch_ba<3, false>(pl->tp, pb, pl->tiles_x, pl->tiles_y, total, n0, st);
    return pl->inter ? launch_ba<1, true>(pl->tp, pb, pl->tiles_x, pl->tiles_y, total, n0, st)
                     : launch_ba<1, false>(pl->tp, pb, pl->tiles_x, pl->tiles_y, total, n0, st);
}

}  // namespace pws
