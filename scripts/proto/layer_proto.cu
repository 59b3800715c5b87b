// Prototype: one hidden layer (k=3, 16->16) of the incremental window update as a
// stand-alone batched kernel: warp per item, tile gathered from global memory, weights
// read from constant memory (uniform datapath), FFMA2 accumulation.  Measures the
// achieved fraction of the FP32 peak for several region sizes / sites-per-lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../qmcnn_b200/csrc layer_proto.cu -o layer_proto
#include <cstdio>
#include <vector>
#include "qmc_device.cuh"
using namespace qmc;

template <int P, int COUT>
__global__ void __launch_bounds__(512, 1)
k_layer(const float* __restrict__ in_tiles, float* __restrict__ out, int n_items, int rh, int rw, int wbase, int bbase) {
    extern __shared__ float4 smem4[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int th = rh + 2, tw = rw + 2, tarea = th * tw;
    float* tile = reinterpret_cast<float*>(smem4) + (size_t)warp * tarea * 16;
    const int rarea = rh * rw;
    for (int item = blockIdx.x * nwarps + warp; item < n_items; item += gridDim.x * nwarps) {
        const float* src = in_tiles + (size_t)item * tarea * 16;
        for (int i = lane; i < tarea * 4; i += 32) cp_async16(reinterpret_cast<float4*>(tile) + i, src + i * 4);
        cp_async_wait_all();
        __syncwarp();
        float4* o4 = reinterpret_cast<float4*>(out + (size_t)item * rarea * COUT);
        conv_region_tiled<3, 16, COUT, P, true, 1>(wbase, bbase, nullptr, tile, 0, tw, tarea, rh, rw, lane,
            [&](int, int pos, int, int, int cog, float4 a) {
                a.x = tanhf(a.x); a.y = tanhf(a.y); a.z = tanhf(a.z); a.w = tanhf(a.w);
                o4[cog * rarea + pos] = a;
            });
        __syncwarp();
    }
}

template <int P, int COUT>
void run(int side, int n_items, const float* d_in, float* d_out, int nsm) {
    const int rh = side, rw = side, tarea = (side + 2) * (side + 2);
    const size_t per_warp = (size_t)tarea * 16 * 4;
    int warps = (int)((220 * 1024) / per_warp);
    if (warps > 16) warps = 16;
    const size_t smem = per_warp * warps;
    cudaFuncSetAttribute(k_layer<P, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_layer<P, COUT><<<nsm, warps * 32, smem>>>(d_in, d_out, n_items, rh, rw, 0, 9 * 16 * COUT);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = 2.0 * n_items * side * side * 9.0 * 16 * COUT;
    printf("side %2d  P=%d COUT=%2d warps/SM=%2d  %.3f ms  %.1f TFLOP/s  err=%s\n", side, P, COUT, warps, best,
           flop / (best * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int n_items = 148 * 16 * 8;
    const size_t in_floats = (size_t)n_items * 15 * 15 * 16, out_floats = (size_t)n_items * 13 * 13 * 16;
    float *d_in, *d_out;
    cudaMalloc(&d_in, in_floats * 4); cudaMalloc(&d_out, out_floats * 4);
    std::vector<float> h(in_floats);
    for (size_t i = 0; i < in_floats; ++i) h[i] = 0.001f * (float)((i * 2654435761u) % 1000) - 0.5f;
    cudaMemcpy(d_in, h.data(), in_floats * 4, cudaMemcpyHostToDevice);
    std::vector<float> w(kConstFloats);
    for (int i = 0; i < kConstFloats; ++i) w[i] = 0.01f * (float)((i * 40503u) % 200) - 1.0f;
    cudaMemcpyToSymbol(c_params, w.data(), kConstFloats * 4);
    run<2, 16>(5, n_items, d_in, d_out, prop.multiProcessorCount);
    run<2, 16>(7, n_items, d_in, d_out, prop.multiProcessorCount);
    run<3, 16>(9, n_items, d_in, d_out, prop.multiProcessorCount);
    run<4, 16>(11, n_items, d_in, d_out, prop.multiProcessorCount);
    run<3, 16>(11, n_items, d_in, d_out, prop.multiProcessorCount);
    run<6, 8>(13, n_items, d_in, d_out, prop.multiProcessorCount);
    run<4, 8>(13, n_items, d_in, d_out, prop.multiProcessorCount);
    return 0;
}
