// probe: which L2 cache-hint forms run on sm_100a (st.global / ld.global / cp.async with createpolicy)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pol_last() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long pol_first() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__global__ void k_policy(unsigned long long* out) { out[0] = pol_last(); out[1] = pol_first(); }
__global__ void k_st(float4* g) {
    const unsigned long long p = pol_last();
    float4* a = g + threadIdx.x;
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f), "l"(p) : "memory");
}
__global__ void k_ld(const float4* g, float4* o) {
    const unsigned long long p = pol_first();
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g + threadIdx.x), "l"(p));
    o[threadIdx.x] = v;
}
__global__ void k_cpasync(const float4* g, float4* o, int first) {
    __shared__ float4 s[32];
    const unsigned long long p = first ? pol_first() : pol_last();
    const unsigned d = (unsigned)__cvta_generic_to_shared(s + threadIdx.x);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(g + threadIdx.x), "l"(p) : "memory");
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    o[threadIdx.x] = s[threadIdx.x];
}
#define CK(name) { cudaError_t e = cudaDeviceSynchronize(); printf("%-12s %s\n", name, cudaGetErrorString(e)); if (e != cudaSuccess) return 1; }
int main() {
    float4 *g, *o; unsigned long long* pp;
    cudaMalloc(&g, 32 * 16); cudaMalloc(&o, 32 * 16); cudaMalloc(&pp, 16);
    k_policy<<<1, 1>>>(pp); CK("createpolicy");
    k_st<<<1, 32>>>(g); CK("st hint");
    k_ld<<<1, 32>>>(g, o); CK("ld hint");
    k_cpasync<<<1, 32>>>(g, o, 0); CK("cp.async last");
    k_cpasync<<<1, 32>>>(g, o, 1); CK("cp.async first");
    float4 h[32]; cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
    printf("value %g %g %g %g\n", h[5].x, h[5].y, h[5].z, h[5].w);
    return 0;
}
