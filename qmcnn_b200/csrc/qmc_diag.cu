// qmc_diag.cu - roofline denominators the driver's MEASURED_PEAKS.json does not
// carry: FP32 FMA and MUFU (ex2) issue peaks, measured on the device with
// unrolled register-only loops.  Diagnostics only; not on the hot path.
#include <cuda_runtime.h>
#include "qmcnn_b200.h"
#include "qmc_ip.cuh"

namespace {

template <int ILP>
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_fma2_peak(float* out, int iters, float a, float b) {
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_mufu_peak(float* out, int iters) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.01f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
}

} // namespace

static int diag_peaks_impl(int device, double* fp32_tflops, double* mufu_gops, double* ffma2_tflops) {
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return QMC_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    float* d = nullptr;
    cudaMalloc(&d, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    constexpr int ILP = 8;
    const int grid = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double best_f = 0, best_m = 0, best_f2 = 0;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0;
        cudaEventRecord(e0);
        k_fma_peak<ILP><<<grid, threads>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * grid * threads * (double)iters * 8 * ILP;
        if (ms > 0 && flop / (ms * 1e-3) * 1e-12 > best_f) best_f = flop / (ms * 1e-3) * 1e-12;
        cudaEventRecord(e0);
        k_fma2_peak<ILP><<<grid, threads>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms > 0 && 2.0 * flop / (ms * 1e-3) * 1e-12 > best_f2) best_f2 = 2.0 * flop / (ms * 1e-3) * 1e-12;
        cudaEventRecord(e0);
        k_mufu_peak<ILP><<<grid, threads>>>(d, iters / 4);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * threads * (iters / 4) * 8.0 * ILP;
        if (ms > 0 && ops / (ms * 1e-3) * 1e-9 > best_m) best_m = ops / (ms * 1e-3) * 1e-9;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    cudaSetDevice(prev);
    if (fp32_tflops) *fp32_tflops = best_f;
    if (mufu_gops) *mufu_gops = best_m;
    if (ffma2_tflops) *ffma2_tflops = best_f2;
    return e == cudaSuccess ? QMC_OK : QMC_ERR_CUDA;
}

extern "C" int qmc_diag_peaks(int device, double* fp32_tflops, double* mufu_gops) {
    return diag_peaks_impl(device, fp32_tflops, mufu_gops, nullptr);
}

extern "C" int qmc_diag_peaks2(int device, double* fp32_tflops, double* ffma2_tflops, double* mufu_gops) {
    return diag_peaks_impl(device, fp32_tflops, mufu_gops, ffma2_tflops);
}


// every float: ip_tanh4 (the small-argument fast path of the in-place evaluator's epilogue) against tanhf, bit for bit
__global__ void k_tanh_check(unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)i);
        const float4 r = qmc::ip_tanh4(make_float4(x, x * 0.5f, -x, 0.f));
        const float w0 = tanhf(x), w1 = tanhf(x * 0.5f), w2 = tanhf(-x);
        const bool ok0 = __float_as_uint(r.x) == __float_as_uint(w0) || (r.x != r.x && w0 != w0);
        const bool ok1 = __float_as_uint(r.y) == __float_as_uint(w1) || (r.y != r.y && w1 != w1);
        const bool ok2 = __float_as_uint(r.z) == __float_as_uint(w2) || (r.z != r.z && w2 != w2);
        if (!(ok0 && ok1 && ok2)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" int qmc_diag_tanh_check(int device, unsigned long long* mismatches /*host*/) {
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return QMC_ERR_BAD_ARGUMENT;
    unsigned long long* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 8);
    if (e == cudaSuccess) e = cudaMemset(d, 0, 8);
    if (e == cudaSuccess) {
        k_tanh_check<<<148 * 8, 256>>>(d);
        e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    cudaSetDevice(prev);
    return e == cudaSuccess ? QMC_OK : QMC_ERR_CUDA;
}
