// Microbenchmark of one CTC lattice step in the scaled-linear format (mantissa, integer exponent), single warp,
// K = 2 nodes per lane.  Compares against tools/ubench_step.cu (split-log2 arithmetic, ~225 cycles/step).
//   VAR 0: scales 2^(e_i - E) by MUFU.EX2 on integer-valued floats
//   VAR 1: scales built on the FMA/ALU pipes (magic-number float->int, shift into the exponent field)
//   +2   : with the per-frame float4 store
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_step_linear ubench_step_linear.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#define STEPS 2048
constexpr float DEAD = -1.0e30f;
constexpr float MAGIC = 12582912.0f;
__device__ __forceinline__ float ex2a(float x){float y; asm("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}

template <int VAR>
__device__ __forceinline__ float pow2i(float d) {            // d integer-valued, <= 0
    if (VAR & 1) {
        const float c = fmaxf(d, -127.f);
        const int bits = __float_as_int(c + MAGIC);
        return __int_as_float((bits << 23) + 0x3f800000);
    } else {
        return ex2a(d);
    }
}

template <int VAR>
__global__ void k(float4* out, long long* cyc, const float2* __restrict__ lpg, int W, unsigned flag) {
    __shared__ float2 lp[16 * 128];
    for (int i = threadIdx.x; i < 16 * 128; i += blockDim.x) lp[i] = lpg[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    out += (size_t)(threadIdx.x >> 5) * STEPS * 32;
    float m0 = (lane == 0) ? 1.f : 0.f, e0 = (lane == 0) ? 0.f : DEAD, m1 = 0.f, e1 = DEAD;
    const int c1 = 1 + lane;
    const float openoff = ((flag >> lane) & 1u) ? 0.f : DEAD;
    long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < STEPS; t += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2* row = lp + ((t + i) & 15) * W;
        float mp = __shfl_up_sync(0xffffffffu, m1, 1);
        float ep = __shfl_up_sync(0xffffffffu, e1, 1);
        if (lane == 0) { mp = 0.f; ep = DEAD; }
        const float2 em0 = row[0], em1 = row[c1];
        const float E0 = fmaxf(e0, ep);
        const float epo = ep + openoff;
        const float E1 = fmaxf(fmaxf(e1, e0), epo);
        const float pre0 = m0 * pow2i<VAR>(e0 - E0) + mp * pow2i<VAR>(ep - E0);
        const float pre1 = m1 * pow2i<VAR>(e1 - E1) + m0 * pow2i<VAR>(e0 - E1) + mp * pow2i<VAR>(epo - E1);
        m0 = pre0 * em0.x; e0 = E0 + em0.y;
        m1 = pre1 * em1.x; e1 = E1 + em1.y;
        if (VAR & 2) out[(size_t)(t + i) * 32 + lane] = make_float4(m0, e0, m1, e1);
      }
      // renormalise once per chunk: pull the mantissa's exponent into e
      {
          int b0 = __float_as_int(m0), b1 = __float_as_int(m1);
          const int x0 = ((b0 >> 23) & 0xff) - 127, x1 = ((b1 >> 23) & 0xff) - 127;
          if (m0 != 0.f) { e0 += (float)x0; m0 = __int_as_float((b0 & 0x807fffff) | 0x3f800000); }
          if (m1 != 0.f) { e1 += (float)x1; m1 = __int_as_float((b1 & 0x807fffff) | 0x3f800000); }
      }
    }
    long long t1 = clock64();
    out[lane] = make_float4(m0, e0, m1, e1);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ex2_exact(int* bad) {
    const int n = -140 + (int)threadIdx.x;                // -140 .. 115
    const float y = ex2a((float)n);
    const float want = (n < -126) ? 0.f : ldexpf(1.f, n);
    if (y != want) atomicAdd(bad, 1);
}

int main() {
    float4* out; float2* lpg; long long* cyc; int* bad;
    cudaMalloc(&out, sizeof(float4) * (size_t)STEPS * 32 * 8 + 1024); cudaMalloc(&cyc, 64); cudaMalloc(&lpg, sizeof(float2) * 16 * 128);
    cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    float2* hl = new float2[16 * 128];
    for (int i = 0; i < 16 * 128; ++i) hl[i] = make_float2(0.71f + 0.05f * (i % 13), -(float)(8 + i % 7));
    cudaMemcpy(lpg, hl, sizeof(float2) * 16 * 128, cudaMemcpyHostToDevice);
    long long hc;
#define RUN(VAR, NW, name) k<VAR><<<1, 32 * NW>>>(out, cyc, lpg, 100, 0xaaaaaaaau); k<VAR><<<1, 32 * NW>>>(out, cyc, lpg, 100, 0xaaaaaaaau); \
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %d warps %.1f cycles/step\n", name, NW, (double)hc / STEPS);
    RUN(0, 1, "mufu scale") RUN(1, 1, "alu scale") RUN(2, 1, "mufu scale + store") RUN(3, 1, "alu scale + store")
    RUN(2, 4, "mufu scale + store") RUN(3, 4, "alu scale + store") RUN(2, 8, "mufu scale + store") RUN(3, 8, "alu scale + store")
    ex2_exact<<<1, 256>>>(bad);
    int hb; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("ex2.approx.ftz on integers -140..115: %d mismatches vs ldexp\n", hb);
    float4 r[32]; cudaMemcpy(r, out, sizeof(r), cudaMemcpyDeviceToHost);
    printf("lane 3: m0 %g e0 %g m1 %g e1 %g\n", r[3].x, r[3].y, r[3].z, r[3].w);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
