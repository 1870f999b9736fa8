// Microbenchmark: L2 float atomic (RED) throughput on B200 for the shapes the backward scatter uses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_red exp_red.cu && ./exp_red
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// mode 0: full-warp RED.32 aligned rows; 1: unaligned (+1 float); 2: RED.128 (v4) aligned; 3: single-lane RED.32
template <int MODE>
__global__ void red_kernel(float *p, size_t n_floats, int reps)
{
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t nwarps = (gridDim.x * (size_t)blockDim.x) >> 5;
    for (int r = 0; r < reps; ++r) {
        if (MODE == 2) {
            size_t base = ((warp + (size_t)r * nwarps) * 128) % (n_floats - 256);
            float *a = p + base + lane * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(a), "f"(1.0f) : "memory");
        } else {
            size_t base = ((warp + (size_t)r * nwarps) * 32) % (n_floats - 256);
            if (MODE == 1) base += 1;
            if (MODE == 3) { if (lane == 0) atomicAdd(p + base, 1.0f); }
            else atomicAdd(p + base + lane, 1.0f);
        }
    }
}

template <int MODE>
int run(const char *name, float *p, size_t n_floats, int reps, bool prezero)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 148 * 8, threads = 512;
    for (int it = 0; it < 3; ++it) {
        if (prezero) CK(cudaMemsetAsync(p, 0, n_floats * 4));
        cudaEventRecord(a);
        red_kernel<MODE><<<blocks, threads>>>(p, n_floats, reps);
        cudaEventRecord(b);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double warps = (double)blocks * threads / 32, instr = warps * reps;
        const double lanes = (MODE == 3) ? 1 : 32, fl = (MODE == 2) ? 4 : 1;
        const double bytes = instr * lanes * fl * 4;
        if (it == 2)
            printf("%-34s footprint %6.1f MB prezero %d: %7.3f ms  %7.2f G RED-instr/s  %7.1f GB/s of floats\n", name,
                   n_floats * 4 / 1e6, (int)prezero, ms, instr / ms / 1e6, bytes / ms / 1e6);
    }
    return 0;
}

int main()
{
    float *p;
    const size_t big = (size_t)256 << 20;  // 1 GiB of floats
    CK(cudaMalloc(&p, big * 4));
    for (size_t n : {(size_t)6 << 20, (size_t)256 << 20}) {   // 24 MB (L2 resident) and 1 GB
        const int reps = 512;
        run<0>("RED.32 full warp, aligned", p, n, reps, true);
        run<1>("RED.32 full warp, +1 float", p, n, reps, true);
        run<2>("RED.128 (v4) full warp", p, n, reps, true);
        run<3>("RED.32 single lane", p, n, reps, true);
    }
    return 0;
}
