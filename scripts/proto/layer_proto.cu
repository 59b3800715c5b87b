// Prototype: one hidden layer (k=3, 16->16) of the incremental window update as a
// stand-alone batched kernel: warp per item, tile gathered from global memory, weights
// read from constant memory (uniform datapath), FFMA2 accumulation.  Measures the
// achieved fraction of the FP32 peak for several region sizes / sites-per-lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../qmcnn_b200/csrc layer_proto.cu -o layer_proto
#include <cstdio>
#include <vector>
#include "qmc_device.cuh"
using namespace qmc;

// Parameter block in constant memory: weight reads with warp-uniform addresses compile to LDCU into uniform
// registers and FFMA2 takes the weight pair as a uniform operand (FFMA2 R, R.F32, UR.F32x2, R): no LDS, no vector
// registers, no shared memory for weights.  ptxas does this only in a kernel that contains nothing but the conv
// (here); inside the persistent product kernels it falls back to per-lane LDC (profiles/r01_summary.md), which is
// why the product's conv_region_tiled reads its weights from shared memory.
constexpr int kConstFloats = 15360;
static __constant__ float c_params[kConstFloats];

// conv_region_tiled (qmc_device.cuh) with the weights taken from c_params
template <int K, int CIN, int COUT, int P, typename OutF>
__device__ __forceinline__ void conv_region_const(int wbase, int bbase, const float* tin, int tw, int tarea, int rh,
                                                  int rw, int lane, OutF out) {
    constexpr int NCG = CIN / 4;
    const int npos = rh * rw, G = (npos + P - 1) / P;
    const float4* tin4 = reinterpret_cast<const float4*>(tin);
    const FastDiv drw(rw);
    for (int g0 = 0; g0 < G; g0 += kWarp) {
        const bool lane_on = g0 + lane < G;
        const int g = lane_on ? g0 + lane : 0;
        int toff[P], ys[P], xs[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            int pos = g + j * G;
            if (pos >= npos) pos = g;
            ys[j] = drw.div(pos);
            xs[j] = pos - ys[j] * rw;
            toff[j] = ys[j] * tw + xs[j];
        }
        float2 acc[P][COUT / 2];
#pragma unroll
        for (int q2 = 0; q2 < COUT / 2; ++q2) {
            const float2 b = make_float2(c_params[bbase + 2 * q2], c_params[bbase + 2 * q2 + 1]);
#pragma unroll
            for (int j = 0; j < P; ++j) acc[j][q2] = b;
        }
#pragma unroll 1
        for (int d = 0; d < K * K; ++d) {
            const int dy = d / K, dx = d - dy * K;
            const float4* tp = tin4 + dy * tw + dx;
            const int wrow = wbase + d * CIN * COUT;
#pragma unroll 1
            for (int cg = 0; cg < NCG; ++cg) {
                float4 in[P];
#pragma unroll
                for (int j = 0; j < P; ++j) in[j] = tp[cg * tarea + toff[j]];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float2 w[COUT / 2];
#pragma unroll
                    for (int q2 = 0; q2 < COUT / 2; ++q2) {
                        const int wi = wrow + (cg * 4 + c4) * COUT + q2 * 2;   // warp-uniform -> LDCU
                        w[q2] = make_float2(c_params[wi], c_params[wi + 1]);
                    }
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const float v = c4 == 0 ? in[j].x : c4 == 1 ? in[j].y : c4 == 2 ? in[j].z : in[j].w;
                        const float2 v2 = make_float2(v, v);
#pragma unroll
                        for (int q2 = 0; q2 < COUT / 2; ++q2) acc[j][q2] = __ffma2_rn(v2, w[q2], acc[j][q2]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int pos = g + j * G;
            if (!lane_on || pos >= npos) continue;
#pragma unroll
            for (int q4 = 0; q4 < COUT / 4; ++q4)
                out(pos, ys[j], xs[j], q4, make_float4(acc[j][q4 * 2].x, acc[j][q4 * 2].y, acc[j][q4 * 2 + 1].x, acc[j][q4 * 2 + 1].y));
        }
    }
}

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
        conv_region_const<3, 16, COUT, P>(wbase, bbase, tile, tw, tarea, rh, rw, lane,
            [&](int pos, int, int, int cog, float4 a) {
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
