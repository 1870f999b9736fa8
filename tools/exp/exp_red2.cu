// Microbenchmark: float RED throughput by ADDRESS PATTERN (what the straggler queue's scattered REDs cost next to the
// dense "tops").  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_red2 exp_red2.cu && ./exp_red2
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void red(float *p, float v) { asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// MODE 0: 32 consecutive floats, 128-byte aligned (1 line, 4 sectors)      1: the same, +3 floats (2 lines, 5 sectors)
//      2: 32 lanes in 32 different lines (32 sectors)                       3: two half rows in different lines (2 lines, 4 sectors)
//      4: 8 active lanes in 8 different lines                               5: 32 lanes, 8 lines, one sector per lane pair... (16 sectors)
//      6: 3 lanes x different lines (a partial flush)                        7: 32 lanes in 32 different lines of ONE 4 KB page-ish neighbourhood
template <int MODE>
__global__ void k(float *p, unsigned n_lines, int reps)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = 0; r < reps; ++r) {
        const unsigned h = hash(warp + (unsigned)r * nwarps);
        const unsigned line = h % (n_lines - 64);
        float *base = p + (size_t)line * 32;
        if (MODE == 0) red(base + lane, 1.f);
        if (MODE == 1) red(base + 3 + lane, 1.f);
        if (MODE == 2) red(p + (size_t)(hash(h + lane * 7919u) % n_lines) * 32 + (lane & 7) * 4, 1.f);
        if (MODE == 3) red(p + (size_t)((lane < 16 ? line : hash(h + 1) % n_lines)) * 32 + 8 + (lane & 15), 1.f);
        if (MODE == 4) { if (lane < 8) red(p + (size_t)(hash(h + lane * 7919u) % n_lines) * 32 + lane, 1.f); }
        if (MODE == 5) red(p + (size_t)(hash(h + (lane >> 2) * 7919u) % n_lines) * 32 + (lane & 3) * 8, 1.f);
        if (MODE == 6) { if (lane < 3) red(p + (size_t)(hash(h + lane * 7919u) % n_lines) * 32 + lane, 1.f); }
        if (MODE == 7) red(base + (size_t)(hash(h + lane) % 60) * 32 + lane, 1.f);
        // patterns of the marching scatter's "tops": a 32-pixel output row on a stretched, slanted map
        if (MODE == 8) red(base + 5 + lane - (lane > 13), 1.f);                       // consecutive, one duplicate address
        if (MODE == 9) red(base + 5 + lane + (lane > 13), 1.f);                       // consecutive, one gap
        if (MODE == 10) red(base + 5 + lane + (lane > 17 ? 1920 : 0), 1.f);           // slanted: the row continues one frame row (1920 floats) down
        if (MODE == 11) red(base + 5 + lane + (lane > 9 ? 1920 : 0) + (lane > 21 ? 1921 : 0) - (lane > 15), 1.f);   // two slants, a dup
        if (MODE == 12) { if (lane < 16) red(base + 5 + lane, 1.f); }                 // half a warp, consecutive
        if (MODE == 13) { red(base + 5 + lane, 1.f); red(base + 1920 + 5 + lane, 1.f); }   // two consecutive REDs, adjacent frame rows
        if (MODE == 14) { if (lane <= 17) red(base + 5 + lane, 1.f); if (lane > 17) red(base + 1920 + 5 + lane, 1.f); }   // the slanted row as two predicated consecutive REDs
    }
}

template <int MODE>
int run(const char *name, double sectors, float *p, size_t n_floats)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 148 * 4, threads = 512, reps = 256;
    for (int it = 0; it < 3; ++it) {
        CK(cudaMemsetAsync(p, 0, n_floats * 4));
        cudaEventRecord(a);
        k<MODE><<<blocks, threads>>>(p, (unsigned)(n_floats / 32), reps);
        cudaEventRecord(b);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double instr = (double)blocks * threads / 32 * reps;
        if (it == 2)
            printf("%-58s footprint %6.1f MB: %7.3f ms  %6.2f G RED-instr/s  %7.1f G sectors/s  (%.2f clk/instr/SM)\n", name, n_floats * 4 / 1e6, ms,
                   instr / ms / 1e6, instr * sectors / ms / 1e6, ms * 1e-3 * 1.965e9 * 148 / instr);
    }
    return 0;
}

int main()
{
    float *p;
    const size_t big = (size_t)128 << 20;
    CK(cudaMalloc(&p, big * 4));
    for (size_t n : {(size_t)6 << 20, (size_t)128 << 20}) {
        run<0>("32 consecutive floats, aligned (1 line, 4 sectors)", 4, p, n);
        run<1>("32 consecutive floats, +3 (2 lines, 5 sectors)", 5, p, n);
        run<3>("two half rows in two lines (2 lines, 4 sectors)", 4, p, n);
        run<7>("32 lanes, 32 lines within 8 KB", 32, p, n);
        run<2>("32 lanes, 32 random lines (32 sectors)", 32, p, n);
        run<5>("32 lanes, 8 random lines, 4 sectors each (32 sectors)", 32, p, n);
        run<4>("8 lanes, 8 random lines (8 sectors)", 8, p, n);
        run<6>("3 lanes, 3 random lines (3 sectors)", 3, p, n);
        run<8>("consecutive with one duplicate address", 4, p, n);
        run<9>("consecutive with one gap", 5, p, n);
        run<10>("slanted row: 18 + 14 lanes, 1920 floats apart", 6, p, n);
        run<11>("two slants and a duplicate", 8, p, n);
        run<12>("16 of 32 lanes, consecutive", 3, p, n);
        run<13>("two full consecutive REDs (per pair)", 10, p, n);
        run<14>("slanted row as two predicated consecutive REDs (per pair)", 6, p, n);
    }
    return 0;
}
