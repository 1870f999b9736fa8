// Reproducer: which (rank, box) combinations of frame-box TMA loads execute.
#include <cstdio>
#include "../../pwstablenet_b200/csrc/pws_tma.cuh"
using namespace pws::tma;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int c0, int c1, float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 65536);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 32) {
        mbar_arrive_expect_tx(bar, bytes);
        if (RANK == 4) load_4d(smem, &tm, bar, c0, c1, 0, 1);
        else load_3d(smem, &tm, bar, c0, c1, 3);
    }
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) out[0] = reinterpret_cast<float *>(smem)[0];
}
int main()
{
    const int W = 1920, H = 1080, C = 3, N = 2;
    float *d, *o; CK(cudaMalloc(&d, (size_t)W * H * C * N * 4)); CK(cudaMalloc(&o, 4));
    CK(cudaMemset(d, 0, (size_t)W * H * C * N * 4));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000));
    CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000));
    const uint32_t bws[] = {64, 72, 80, 88, 96}, bhs[] = {16, 20, 24, 28};
    for (int rank = 3; rank <= 4; ++rank)
        for (uint32_t bw : bws)
            for (uint32_t bh : bhs) {
                CUtensorMap tm;
                bool ok;
                if (rank == 4) {
                    const uint64_t dims[4] = {W, H, C, N}, st[3] = {W, (uint64_t)W * H, (uint64_t)W * H * C};
                    const uint32_t box[4] = {bw, bh, 3, 1};
                    ok = encode_f32(&tm, d, 4, dims, st, box);
                } else {
                    const uint64_t dims[3] = {W, H, (uint64_t)C * N}, st[2] = {W, (uint64_t)W * H};
                    const uint32_t box[3] = {bw, bh, 3};
                    ok = encode_f32(&tm, d, 3, dims, st, box);
                }
                if (!ok) { printf("rank %d box %ux%u: encode failed\n", rank, bw, bh); continue; }
                const int bytes = bw * bh * 3 * 4;
                if (bytes > 65536) { printf("rank %d box %ux%u: skipped (too big for the test buffer)\n", rank, bw, bh); continue; }
                for (int c0 : {0, 5}) {
                    if (rank == 4) k<4><<<1, 64, 66000>>>(tm, bytes, c0, 7, o); else k<3><<<1, 64, 66000>>>(tm, bytes, c0, 7, o);
                    cudaError_t e = cudaDeviceSynchronize();
                    printf("rank %d box %2ux%2u (%5d B) c0=%d: %s\n", rank, bw, bh, bytes, c0, cudaGetErrorString(e));
                    if (e != cudaSuccess) { printf("context dead, stopping\n"); return 0; }
                }
            }
    return 0;
}
