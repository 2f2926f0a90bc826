// ubench_read.cu -- how fast can one B200 READ HBM through each path the kernels use?  (development tool)
//   mode 0: 1-D bulk async copies (cp.async.bulk, UBLKCP) of ROW-byte rows into a ring of S slots per CTA, 1 CTA per SM,
//           a slot is re-armed as soon as its row has landed (no compute)
//   mode 1: 3-D tensor tile loads (cp.async.bulk.tensor, UTMALDG), box = 240 rows x 8 floats out of a (B, V, T) tensor
//   mode 2: LDG.128 streaming reads, W warps per CTA, 4 loads in flight per thread
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_read tools/ubench_read.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok;
}

__global__ void __launch_bounds__(32, 1) bulk_rows(const float *src, size_t nrows, int row_bytes, int S, unsigned *ticket) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
    unsigned char *ring = sm + 1024;
    const size_t slot = ((size_t)row_bytes + 127) / 128 * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        unsigned phase[32] = {0};
        bool busy[32] = {false};
        bool more = true;
        int inflight = 0;
        while (more || inflight) {
            for (int s = 0; s < S; ++s) {
                if (busy[s]) {
                    if (mbar_try(&bars[s], phase[s] & 1)) { busy[s] = false; ++phase[s]; --inflight; } else continue;
                }
                if (more) {
                    const unsigned r = atomicAdd(ticket, 1u);
                    if (r >= nrows) { more = false; continue; }
                    mbar_expect(&bars[s], row_bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + s * slot)),
                                 "l"(reinterpret_cast<const char *>(src) + (size_t)r * row_bytes), "r"(row_bytes), "r"(smem_u32(&bars[s])) : "memory");
                    busy[s] = true; ++inflight;
                }
            }
        }
    }
}

__device__ __forceinline__ bool mbar_test(uint64_t *b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok;
}
// mode 0b: one lane per slot, every lane re-arms its own slot the moment its row has landed; tickets prefetched
__global__ void __launch_bounds__(32, 1) bulk_rows_par(const float *src, size_t nrows, int row_bytes, int S, unsigned *ticket) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
    unsigned char *ring = sm + 1024;
    const size_t slot = ((size_t)row_bytes + 127) / 128 * 128;
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane >= S) return;
    unsigned phase = 0;
    unsigned next = atomicAdd(ticket, 1u);
    while (next < nrows) {
        const unsigned r = next;
        next = atomicAdd(ticket, 1u);                 // in flight while the row loads
        mbar_expect(&bars[lane], row_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + lane * slot)),
                     "l"(reinterpret_cast<const char *>(src) + (size_t)r * row_bytes), "r"(row_bytes), "r"(smem_u32(&bars[lane])) : "memory");
        while (!mbar_test(&bars[lane], phase & 1)) { }
        ++phase;
    }
}

__global__ void __launch_bounds__(32, 1) tensor_tiles(const __grid_constant__ CUtensorMap map, int B, int K, int nTB, int S, unsigned *ticket, int box_bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm);
    unsigned char *ring = sm + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        unsigned phase[32] = {0};
        bool busy[32] = {false};
        bool more = true;
        int inflight = 0, k = 0, b = 0, tb = 0;
        bool have = false;
        while (more || inflight) {
            for (int s = 0; s < S; ++s) {
                if (busy[s]) {
                    if (mbar_try(&bars[s], phase[s] & 1)) { busy[s] = false; ++phase[s]; --inflight; } else continue;
                }
                if (more) {
                    if (!have || k == K) {
                        const unsigned f = atomicAdd(ticket, 1u);
                        if (f >= (unsigned)(B * nTB)) { more = false; continue; }
                        b = f / nTB; tb = f % nTB; k = 0; have = true;       // consecutive tickets = adjacent frame blocks
                    }
                    mbar_expect(&bars[s], box_bytes);
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(ring + (size_t)s * box_bytes)),
                                 "l"(&map), "r"(tb * 8), "r"(k * 240), "r"(b), "r"(smem_u32(&bars[s])) : "memory");
                    ++k; busy[s] = true; ++inflight;
                }
            }
        }
    }
}

__global__ void ldg_stream(const float4 *src, size_t n4, float *sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(src + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(src + i + 3 * stride));
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) *sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int B = 64, T = 800, V = 3500;
    const size_t n = (size_t)B * T * V;
    float *buf, *sink;
    unsigned *ticket;
    cudaMalloc(&buf, n * 4); cudaMalloc(&sink, 4); cudaMalloc(&ticket, 4);
    cudaMemset(buf, 0, n * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char *what, int param, float ms) { printf("%-28s %4d  %8.3f ms  %7.1f GB/s\n", what, param, ms, n * 4 / ms / 1e6); };
    cudaFuncSetAttribute(bulk_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tensor_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int rb : {14000, 7008, 32000}) {
        for (int S : {2, 4, 6, 8, 12, 15}) {
            const size_t slot = ((size_t)rb + 127) / 128 * 128;
            if (1024 + S * slot > 227 * 1024) continue;
            const size_t nrows = n * 4 / rb;
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaMemset(ticket, 0, 4);
                cudaEventRecord(e0);
                bulk_rows<<<sms, 32, 1024 + S * slot>>>(buf, nrows, rb, S, ticket);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            char nm[64]; snprintf(nm, 64, "bulk rows %d B, slots", rb);
            report(nm, S, best);
        }
    }
    cudaFuncSetAttribute(bulk_rows_par, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int rb : {14000, 7008, 3504, 32000}) {
        for (int S : {2, 4, 6, 8, 12, 15, 24, 32}) {
            const size_t slot = ((size_t)rb + 127) / 128 * 128;
            if (1024 + S * slot > 227 * 1024) continue;
            const size_t nrows = n * 4 / rb;
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaMemset(ticket, 0, 4);
                cudaEventRecord(e0);
                bulk_rows_par<<<sms, 32, 1024 + S * slot>>>(buf, nrows, rb, S, ticket);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            char nm[64]; snprintf(nm, 64, "par bulk rows %d B, slots", rb);
            report(nm, S, best);
        }
    }
    // tensor tiles
    EncodeTiledFn fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&fn, cudaEnableDefault, &q);
    for (int promo = 0; promo < 2 && fn; ++promo) {
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)V, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)T * 4, (cuuint64_t)T * V * 4};
        const cuuint32_t box[3] = {8, 240, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); break; }
        for (int S : {4, 8, 12, 16, 24, 28}) {
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaMemset(ticket, 0, 4);
                cudaEventRecord(e0);
                tensor_tiles<<<sms, 32, 1024 + S * 7680>>>(map, B, 15, T / 8, S, ticket, 7680);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            report(promo ? "tensor 240x8 L2promo128, S" : "tensor 240x8 no promo, S", S, best);
        }
    }
    for (int wpc : {8, 16, 32}) {
        for (int cps : {1, 2}) {
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                ldg_stream<<<sms * cps, 32 * wpc>>>(reinterpret_cast<const float4 *>(buf), n / 4, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            char nm[64]; snprintf(nm, 64, "LDG.128 x4, %d CTA/SM, warps", cps);
            report(nm, wpc, best);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
