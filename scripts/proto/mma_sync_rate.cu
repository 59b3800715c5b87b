// mma_sync_rate.cu - issue rate of the legacy warp-level tensor path (mma.sync, SASS HMMA) on sm_100a, register
// operands only: m16n8k8 TF32 and m16n8k16 BF16, fp32 accumulate.  Context for the go / no-go on moving the
// 16 -> 16 conv layers off the FFMA pipe (profiles/r02_summary.md).
#include <cuda_runtime.h>
#include <cstdio>
template <int KIND, int NACC>
__global__ void __launch_bounds__(256) k_rate(float* out, int iters) {
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void run(const char* name, double flop_per_mma, int nsm) {
    float* d; cudaMalloc(&d, 148 * 4 * 256 * 4 * 4);
    const int iters = 4096, NACC = 8, blocks = nsm * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k_rate<KIND, NACC><<<blocks, 256>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double mmas = (double)blocks * 8 * iters * NACC;
    printf("%s: %.3f ms, %.1f TFLOP/s, %.2f cycles per warp-MMA per SM sub-partition at 1.965 GHz (%s)\n", name, best,
           mmas * flop_per_mma / (best * 1e-3) * 1e-12, best * 1e-3 * 1.965e9 / (mmas / (nsm * 4)), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    run<0>("mma.sync m16n8k8 tf32", 2.0 * 16 * 8 * 8, p.multiProcessorCount);
    run<1>("mma.sync m16n8k16 bf16", 2.0 * 16 * 8 * 16, p.multiProcessorCount);
    return 0;
}
