// Dependent-chain latency microbenchmark for the ops on the lattice recursion's critical path (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_latency ubench_latency.cu && ./ubench_latency
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ float ex2a(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float lg2a(float x){float y; asm volatile("lg2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int OP> __global__ void k(float* out, long long* cyc, float a, float b){
    float x = a + threadIdx.x * 1e-9f, y = b;
    __shared__ float sm[64];
    sm[threadIdx.x] = a; __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __fadd_rn(x, y);
        if (OP == 1) x = ex2a(x);
        if (OP == 2) x = lg2a(x);
        if (OP == 3) x = __shfl_up_sync(0xffffffffu, x, 1);
        if (OP == 4) x = fmaxf(x, y) ;
        if (OP == 5) { x = ex2a(x); x = __fadd_rn(x, y); }
        if (OP == 6) { volatile float* p = sm; x = p[(int)x & 31]; }
        if (OP == 7) { x = ex2a(x); y = ex2a(y); }      // two independent MUFU chains: issue rate
        if (OP == 8) { x = __fadd_rn(x, a); y = __fadd_rn(y, a); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main(){
    float* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
    const char* names[] = {"FADD dep", "MUFU.EX2 dep", "MUFU.LG2 dep", "SHFL.UP dep", "FMNMX dep", "EX2+FADD dep", "LDS dep", "2x EX2 indep", "2x FADD indep"};
    long long h;
#define RUN(OP) k<OP><<<1,32>>>(out,cyc,0.5f,0.25f); k<OP><<<1,32>>>(out,cyc,0.5f,0.25f); cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost); printf("%-16s %.2f cycles/iter\n", names[OP], (double)h/N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    // same with 2 warps on the same SMSP? (block of 160 threads: warps 0 and 4 share a sub-partition)
    k<1><<<1,160>>>(out,cyc,0.5f,0.25f); cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost); printf("EX2 dep, 5 warps   %.2f cycles/iter\n", (double)h/N);
    k<7><<<1,160>>>(out,cyc,0.5f,0.25f); cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost); printf("2xEX2, 5 warps     %.2f cycles/iter\n", (double)h/N);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
