// Microbenchmark of one CTC lattice step (K nodes per lane, split-log2 arithmetic), single warp.
// Variants isolate what the recursion's ~450 cycles/step are made of.
#include <cstdio>
#include <cuda_runtime.h>
#define STEPS 2048
constexpr float SENT = -1.0e30f;
constexpr float MAGIC = 12582912.0f;
__device__ __forceinline__ float ex2a(float x){float y; asm("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ float lg2a(float x){float y; asm("lg2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}

template<int K, int VAR>
__global__ void k(float2* out, long long* cyc, const float2* __restrict__ lpg, int W, unsigned flag) {
    __shared__ float2 lp[16 * 128];
    for (int i = threadIdx.x; i < 16 * 128; i += blockDim.x) lp[i] = lpg[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    out += (size_t)(threadIdx.x >> 5) * STEPS * 32 * 8;
    float h[K], l[K], uh[K], ul[K]; int ci[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { h[r] = (lane == 0 && r == 0) ? 0.f : SENT; l[r] = 0.f; uh[r] = h[r]; ul[r] = 0.f; ci[r] = (r & 1) ? 1 + (K * lane + r) / 2 : 0; }
    long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < STEPS; t += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2* row = lp + ((t + i) & 15) * W;
        float eh[K + 2], el[K + 2];
        if (VAR == 3 || VAR == 4) {
            eh[0] = __shfl_up_sync(0xffffffffu, uh[K - 2], 1); el[0] = __shfl_up_sync(0xffffffffu, ul[K - 2], 1);
            eh[1] = __shfl_up_sync(0xffffffffu, uh[K - 1], 1); el[1] = __shfl_up_sync(0xffffffffu, ul[K - 1], 1);
        } else {
            eh[0] = __shfl_up_sync(0xffffffffu, h[K - 2], 1); el[0] = __shfl_up_sync(0xffffffffu, l[K - 2], 1);
            eh[1] = __shfl_up_sync(0xffffffffu, h[K - 1], 1); el[1] = __shfl_up_sync(0xffffffffu, l[K - 1], 1);
        }
        if (lane == 0) { eh[0] = SENT; eh[1] = SENT; el[0] = 0.f; el[1] = 0.f; }
#pragma unroll
        for (int r = 0; r < K; ++r) { eh[2 + r] = h[r]; el[2 + r] = l[r]; }
#pragma unroll
        for (int r = 0; r < K; ++r) {
            const int I = 2 + r;
            float ph, pl;
            if ((r & 1) == 0) {
                const float hm = fmaxf(eh[I], eh[I - 1]);
                ph = hm; pl = lg2a(ex2a((eh[I] - hm) + el[I]) + ex2a((eh[I - 1] - hm) + el[I - 1]));
            } else {
                const float h2 = ((flag >> r) & 1u) ? eh[I - 2] : SENT;
                const float hm = fmaxf(fmaxf(eh[I], eh[I - 1]), h2);
                ph = hm; pl = lg2a(ex2a((eh[I] - hm) + el[I]) + ex2a((eh[I - 1] - hm) + el[I - 1]) + ex2a((h2 - hm) + el[I - 2]));
            }
            const float2 e = row[ci[r]];
            float nl = pl + e.y, nh = ph + e.x;
            if (VAR == 3 || VAR == 4) {           // VAR 3/4: publish un-normalised, renormalise own copy lazily
                uh[r] = nh; ul[r] = nl;
            }
            if (VAR != 1) {                       // VAR 1: no renormalisation (just to see its cost)
                const float rr = __fsub_rn(__fadd_rn(nl, MAGIC), MAGIC);
                nh += rr; nl -= rr;
            }
            h[r] = nh; l[r] = nl;
            if (VAR == 2 || VAR == 4) out[(size_t)(t + i) * 32 * K + K * lane + r] = make_float2(nh, nl);   // VAR 2: with fv stores
        }
      }
    }
    long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) acc += h[r] + l[r];
    out[lane] = make_float2(acc, 0.f);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float2 *out, *lpg; long long* cyc;
    cudaMalloc(&out, sizeof(float2) * (size_t)STEPS * 32 * 8 * 8 + 1024); cudaMalloc(&cyc, 64); cudaMalloc(&lpg, sizeof(float2) * 16 * 128);
    float2* hl = new float2[16 * 128];
    for (int i = 0; i < 16 * 128; ++i) hl[i] = make_float2(-(float)(8 + i % 7), 0.01f * (i % 13) - 0.05f);
    cudaMemcpy(lpg, hl, sizeof(float2) * 16 * 128, cudaMemcpyHostToDevice);
    long long hc;
#define RUN(K, VAR, name) k<K, VAR><<<1, 32>>>(out, cyc, lpg, 100, 0xaaaaaaaau); k<K, VAR><<<1, 32>>>(out, cyc, lpg, 100, 0xaaaaaaaau); \
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); printf("K=%d %-22s %.1f cycles/step\n", K, name, (double)hc / STEPS);
    RUN(2, 0, "base") RUN(2, 1, "no renorm") RUN(2, 2, "with stores") RUN(2, 3, "lazy renorm") RUN(2, 4, "lazy renorm+stores") RUN(4, 3, "lazy renorm") RUN(4, 4, "lazy renorm+stores")
    RUN(4, 0, "base") RUN(4, 2, "with stores")
    RUN(6, 0, "base") RUN(6, 2, "with stores")
    RUN(8, 0, "base")
#define RUNW(K, VAR, NW) k<K, VAR><<<1, 32 * NW>>>(out, cyc, lpg, 100, 0xaaaaaaaau); k<K, VAR><<<1, 32 * NW>>>(out, cyc, lpg, 100, 0xaaaaaaaau); \
    cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); printf("K=%d var %d, %d warps/CTA: %.1f cycles/step\n", K, VAR, NW, (double)hc / STEPS);
    RUNW(2, 2, 2) RUNW(2, 2, 4) RUNW(2, 2, 5) RUNW(2, 2, 6) RUNW(2, 2, 8) RUNW(6, 2, 2) RUNW(6,2,5) RUNW(2, 0, 6) RUNW(2, 0, 8)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
