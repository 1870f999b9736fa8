// Microbenchmark: issue / pipe throughput of the instructions the backward kernel is made of (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_pipes exp_pipes.cu && ./exp_pipes
// Prints warp-instructions per clock per SM for independent chains of each instruction (8 chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int kIters = 2048, kChains = 8;

template <int OP>
__global__ void k(float *out, const float *in, long long *cycles)
{
    __shared__ float sm[4096];
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += blockDim.x) sm[i] = in[i & 255];
    __syncthreads();
    float a[kChains], b = in[tid & 255], c = in[(tid + 7) & 255];
    unsigned long long p[kChains], pb, pc;
    int ia[kChains], ib = (int)in[3] + tid, ic = (int)in[5] | 1;
    for (int j = 0; j < kChains; ++j) { a[j] = in[(tid + j) & 255]; ia[j] = tid * j + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(p[j]) : "f"(a[j]), "f"(b)); }
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(c));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int j = 0; j < kChains; ++j) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(pb), "l"(pc));
            if (OP == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[j]) : "l"(pc));
            if (OP == 3) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ia[j]) : "r"(ib), "r"(ic));
            if (OP == 4) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ib));
            if (OP == 5) asm volatile("xor.b32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ib));
            if (OP == 6) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(ia[j]));
            if (OP == 7) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[j]) : "r"((unsigned)__cvta_generic_to_shared(sm + ((tid + j * 32 + it) & 4095))) : "memory");
            if (OP == 8) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c)); asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ib)); }
            if (OP == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(pb), "l"(pc)); asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ib)); }
            if (OP == 10) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b));
            if (OP == 11) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b));
            if (OP == 12) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ia[j]) : "r"(ib), "r"(ic)); }
            if (OP == 13) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(a[j]) : "f"(b));
        }
    }
    const long long t1 = clock64();
    float s = 0.f; int si = 0;
    for (int j = 0; j < kChains; ++j) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[j])); s += a[j] + lo + hi; si += ia[j]; }
    out[blockIdx.x * blockDim.x + tid] = s + (float)si;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char *name, int per_iter, float *out, float *in, long long *cyc)
{
    for (int warps : {4, 8, 16, 32}) {
        k<OP><<<148, warps * 32>>>(out, in, cyc);
        CK(cudaDeviceSynchronize());
        long long h[148];
        CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
        const double instr = (double)kIters * kChains * per_iter * warps;
        printf("%-44s %2d warps/SM: %6.3f warp-instr/clk/SM\n", name, warps, instr / avg);
    }
    return 0;
}

int main()
{
    float *out, *in; long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&in, 4096 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0f + (i % 13) * 1e-3f;
    CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    run<0>("FFMA (3 registers)", 1, out, in, cyc);
    run<13>("FFMA (immediate addend)", 1, out, in, cyc);
    run<10>("FMUL", 1, out, in, cyc);
    run<11>("FADD", 1, out, in, cyc);
    run<1>("FFMA2 (3 register pairs)", 1, out, in, cyc);
    run<2>("FMUL2 (pair x broadcast)", 1, out, in, cyc);
    run<3>("IMAD", 1, out, in, cyc);
    run<4>("IADD", 1, out, in, cyc);
    run<5>("XOR (LOP3)", 1, out, in, cyc);
    run<6>("SHFL.UP", 1, out, in, cyc);
    run<7>("LDS.32 conflict-free", 1, out, in, cyc);
    run<8>("FFMA + IADD interleaved", 2, out, in, cyc);
    run<12>("FFMA + IMAD interleaved", 2, out, in, cyc);
    run<9>("FFMA2 + IADD interleaved", 2, out, in, cyc);
    return 0;
}
