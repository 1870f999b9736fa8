// Microbenchmark: TMA box load / store / f32 add-reduce throughput on B200 for the tile shapes the
// pipelined warp kernels use (boxes of a (W,H,C*N) fp32 tensor at a 64x16 tile pitch).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp_tma exp_tma.cu && ./exp_tma
#include <cstdio>
#include <vector>
#include "../../pwstablenet_b200/csrc/pws_tma.cuh"
using namespace pws::tma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int TW = 64, TH = 16;

// mode 0: load boxes (producer warp + 4 consumer warps, STAGES-deep ring)
template <int BW, int BH, int STAGES>
__global__ void __launch_bounds__(160) load_kernel(const __grid_constant__ CUtensorMap tm, int tx_n, int ty_n, int nimg, float *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BOX = BW * BH * 3;
    float *buf = reinterpret_cast<float *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * BOX * 4);
    uint64_t *empty = full + STAGES;
    const int total = tx_n * ty_n * nimg;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        fence_barrier_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 4) {
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n), ty = r / tx_n, tx = r % tx_n;
                mbar_arrive_expect_tx(full + s, BOX * 4);
                load_3d(buf + (size_t)s * BOX, &tm, full + s, tx * TW - (BW - TW) / 2, ty * TH - (BH - TH) / 2, img * 3);
            }
        }
    } else {
        float acc = 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(full + s, ph);
            const float *b = buf + (size_t)s * BOX;
            for (int i = threadIdx.x; i < BOX; i += 128 * 8) acc += b[i];
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        if (acc == 123.456f) sink[0] = acc;
    }
}

// mode 1: store, mode 2: reduce-add.  All 128 threads fill the box, thread 0 issues the bulk op.
template <int BW, int BH, int STAGES, bool RED>
__global__ void __launch_bounds__(128) store_kernel(const __grid_constant__ CUtensorMap tm, int tx_n, int ty_n, int nimg, int fill_div, int clampneg)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BOX = BW * BH * 3;
    float *buf = reinterpret_cast<float *>(smem);
    const int total = tx_n * ty_n * nimg;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) {
            if (threadIdx.x == 0) wait_group_read<STAGES - 1>();
            __syncthreads();
        }
        float *b = buf + (size_t)s * BOX;
        for (int i = threadIdx.x * 4; i < BOX / fill_div; i += 128 * 4) *reinterpret_cast<float4 *>(b + i) = make_float4(1.f, 1.f, 1.f, 1.f);
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int img = t / (tx_n * ty_n), r = t % (tx_n * ty_n), ty = r / tx_n, tx = r % tx_n;
            int c0 = tx * TW - (BW - TW) / 2, c1 = ty * TH - (BH - TH) / 2;
            if (clampneg) { c0 = c0 < 0 ? 0 : c0; c1 = c1 < 0 ? 0 : c1; }
            if (RED) reduce_add_3d(&tm, b, c0, c1, img * 3);
            else store_3d(&tm, b, c0, c1, img * 3);
            commit_group();
        }
    }
    if (threadIdx.x == 0) wait_group<0>();
}

__global__ void sum_kernel(const float *p, size_t n, double *out)
{
    double a = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += p[i];
    for (int o = 16; o; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, a);
}

template <int BW, int BH>
int run(float *p, double *dsum, int W, int H, int nimg, int ctas_per_sm, int clampneg = 1)
{
    CUtensorMap tm;
    if (!encode_3d_f32(&tm, p, W, H, (uint64_t)3 * nimg, W, (uint64_t)W * H, BW, BH, 3)) { printf("encode failed\n"); return 1; }
    constexpr int STAGES = 4, BOX = BW * BH * 3;
    const int tx_n = (W + TW - 1) / TW, ty_n = (H + TH - 1) / TH;
    const size_t n = (size_t)W * H * 3 * nimg;
    const int smem_l = STAGES * BOX * 4 + 2 * STAGES * 8, smem_s = STAGES * BOX * 4;
    CK(cudaFuncSetAttribute(load_kernel<BW, BH, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_l));
    CK(cudaFuncSetAttribute(store_kernel<BW, BH, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
    CK(cudaFuncSetAttribute(store_kernel<BW, BH, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_s));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * ctas_per_sm;
    float *sink; CK(cudaMalloc(&sink, 4));
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int it = 0; it < 4; ++it) {
            if (mode >= 2) CK(cudaMemsetAsync(p, 0, n * 4));
            cudaEventRecord(a);
            if (mode == 0) load_kernel<BW, BH, STAGES><<<grid, 160, smem_l>>>(tm, tx_n, ty_n, nimg, sink);
            else if (mode == 1) store_kernel<BW, BH, STAGES, false><<<grid, 128, smem_s>>>(tm, tx_n, ty_n, nimg, 1, clampneg);
            else if (mode == 2) store_kernel<BW, BH, STAGES, true><<<grid, 128, smem_s>>>(tm, tx_n, ty_n, nimg, 1, clampneg);
            else store_kernel<BW, BH, STAGES, true><<<grid, 128, smem_s>>>(tm, tx_n, ty_n, nimg, 1 << 20, clampneg);  // no fill: issue cost only
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it > 0 && ms < best) best = ms;
        }
        const char *names[4] = {"load", "store", "reduce-add", "reduce-add(no fill)"};
        const double boxes = (double)tx_n * ty_n * nimg;
        printf("  box %3dx%2dx3 ctas/SM %d %-20s %8.3f ms  box-bytes %7.1f GB/s  unique-bytes %7.1f GB/s  %6.2f M boxes/s\n", BW, BH, ctas_per_sm,
               names[mode], best, boxes * BOX * 4 / best / 1e6, n * 4.0 / best / 1e6, boxes / best / 1e3);
        if (mode == 2) {
            CK(cudaMemset(dsum, 0, 8));
            sum_kernel<<<1024, 256>>>(p, n, dsum);
            double h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
            // expected: every in-range element of every box counted once
            double expect = 0;
            for (int ty = 0; ty < ty_n; ++ty)
                for (int tx = 0; tx < tx_n; ++tx) {
                    int x0 = tx * TW - (BW - TW) / 2, y0 = ty * TH - (BH - TH) / 2;
                    if (clampneg) { x0 = x0 < 0 ? 0 : x0; y0 = y0 < 0 ? 0 : y0; }
                    int xa = x0 < 0 ? 0 : x0, xb = x0 + BW > W ? W : x0 + BW, ya = y0 < 0 ? 0 : y0, yb = y0 + BH > H ? H : y0 + BH;
                    expect += (double)(xb - xa) * (yb - ya);
                }
            expect *= 3.0 * nimg;
            printf("     reduce check: sum %.0f expected %.0f %s\n", h, expect, h == expect ? "OK" : "MISMATCH");
        }
    }
    cudaFree(sink);
    return 0;
}

int main()
{
    const int W = 1920, H = 1080;
    float *p; double *dsum;
    CK(cudaMalloc(&p, (size_t)W * H * 3 * 16 * 4));
    CK(cudaMalloc(&dsum, 8));
    for (int nimg : {2, 16}) {
        printf("frames %d (%.0f MB)\n", nimg, W * H * 3.0 * nimg * 4 / 1e6);
        for (int c : {1, 2}) {
            if (run<64, 16>(p, dsum, W, H, nimg, c)) return 1;
            if (run<80, 24>(p, dsum, W, H, nimg, c)) return 1;
        }
        if (run<128, 16>(p, dsum, W, H, nimg, 2)) return 1;
    }
    return 0;
}
