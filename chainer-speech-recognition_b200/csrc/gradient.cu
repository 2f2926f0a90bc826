// gradient.cu -- kernel 3: fused gradient, plus the deterministic loss sum.
//
// Replaces GramCTC.backward (asr/loss/gram_ctc.py:284-297) and _compute_label_probability (:180-217):
//   grad[t,b,k] = ( softmax[t,b,k] - exp( LSE_{j: symbol_j = k} (alpha_t[j]+beta_t[j]) - log P_b ) ) * gy * scale
//   and exactly 0 for t >= input_length[b] (:296).
// The reference keeps the softmax tensor alive from forward (717 MB) and materialises a second
// (T,B,V) array of per-unit posteriors; here the activations are re-read once, the softmax is
// recomputed from the saved per-frame normaliser, and the <= L+1 non-zero posteriors of a frame are
// merged from gamma on the fly.  One read + one write of (T,B,V), nothing else of that size.
//
// One warp per frame; 128-bit streaming loads/stores; the few label columns are patched after the
// row has been written (same warp, ordered by __syncwarp).
#include "common.cuh"
#include "kernels.h"

namespace b200ctc {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kUnroll = 4;

__device__ __forceinline__ void zero_row(float *__restrict__ g, int V, int lane) {
    const int mis = (int)((reinterpret_cast<uintptr_t>(g) >> 2) & 3);
    const int head = mis ? min(V, 4 - mis) : 0;
    if (lane < head) g[lane] = 0.f;
    float4 *g4 = reinterpret_cast<float4 *>(g + head);
    const int n4 = (V - head) >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < n4; i += 32) stg_stream4(g4 + i, z);
    const int tail0 = head + 4 * n4;
    if (tail0 + lane < V) g[tail0 + lane] = 0.f;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) gradient_kernel(GradParams gp, WsLayout w,
                                                                     const unsigned char *ws, int b_major,
                                                                     int sm_floats_per_warp) {
    extern __shared__ float sm_all[];
    const ProblemDesc &d = gp.d;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *e_sm = sm_all + (size_t)warp * sm_floats_per_warp;        // [Np] exp2(gamma) of this frame
    const int warp_global = blockIdx.x * kWarpsPerCta + warp;
    const int warps_total = gridDim.x * kWarpsPerCta;
    const UttInfo *utt = reinterpret_cast<const UttInfo *>(ws + w.off_utt);
    const float *lse_all = reinterpret_cast<const float *>(ws + w.off_lse);
    const float *gam_all = reinterpret_cast<const float *>(ws + w.off_gam);
    const int *usym_all = reinterpret_cast<const int *>(ws + w.off_usym);
    const int *uoff_all = reinterpret_cast<const int *>(ws + w.off_uoff);
    const int *unode_all = reinterpret_cast<const int *>(ws + w.off_unode);
    const int per = d.kind == 0 ? 2 : 3;
    const long long frames = (long long)d.B * d.T;

    for (long long f = warp_global; f < frames; f += warps_total) {
        int b, t;
        if (b_major) { b = (int)(f / d.T); t = (int)(f % d.T); }
        else { t = (int)(f / d.B); b = (int)(f % d.B); }
        float *grow = gp.grad_out + (int64_t)t * gp.gstride_t + (int64_t)b * gp.gstride_b;
        const UttInfo ui = utt[b];
        if (t >= ui.Tb) { zero_row(grow, d.V, lane); continue; }             // :296

        const float *row = d.acts + (int64_t)t * d.stride_t + (int64_t)b * d.stride_b;
        const float lse2 = lse_all[(size_t)b * d.T + t];
        const float gy = gp.per_utterance ? gp.grad_loss[b] : gp.grad_loss[0];
        const float sc = gy * gp.scale;                                      // :291-294
        const float c = -lse2;

        // ---- posteriors of this frame: e[j] = 2^gamma[t][j] ----
        const float *gam = gam_all + ((size_t)b * d.T + t) * w.Np;
        float blank_part = 0.f;
        for (int j = lane; j < ui.Nb; j += 32) {
            const float e = ex2_approx(__ldg(gam + j));
            e_sm[j] = e;
            if (j % per == 0) blank_part += e;
        }
        blank_part = warp_sum(blank_part);
        __syncwarp();

        // ---- stream the row: grad = softmax * sc ----
        const int mis = (int)((reinterpret_cast<uintptr_t>(row) >> 2) & 3);
        const int gmis = (int)((reinterpret_cast<uintptr_t>(grow) >> 2) & 3);
        if (mis == gmis) {
            const int head = mis ? min(d.V, 4 - mis) : 0;
            if (lane < head) grow[lane] = ex2_approx(fmaf(row[lane], LOG2E_HI, c)) * sc;
            const float4 *row4 = reinterpret_cast<const float4 *>(row + head);
            float4 *g4 = reinterpret_cast<float4 *>(grow + head);
            const int n4 = (d.V - head) >> 2;
            int i = lane;
            for (; i + 32 * (kUnroll - 1) < n4; i += 32 * kUnroll) {
                float4 v[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) v[u] = ldg_stream4(row4 + i + 32 * u);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    float4 o;
                    o.x = ex2_approx(fmaf(v[u].x, LOG2E_HI, c)) * sc;
                    o.y = ex2_approx(fmaf(v[u].y, LOG2E_HI, c)) * sc;
                    o.z = ex2_approx(fmaf(v[u].z, LOG2E_HI, c)) * sc;
                    o.w = ex2_approx(fmaf(v[u].w, LOG2E_HI, c)) * sc;
                    stg_stream4(g4 + i + 32 * u, o);
                }
            }
            for (; i < n4; i += 32) {
                const float4 v = ldg_stream4(row4 + i);
                float4 o;
                o.x = ex2_approx(fmaf(v.x, LOG2E_HI, c)) * sc;
                o.y = ex2_approx(fmaf(v.y, LOG2E_HI, c)) * sc;
                o.z = ex2_approx(fmaf(v.z, LOG2E_HI, c)) * sc;
                o.w = ex2_approx(fmaf(v.w, LOG2E_HI, c)) * sc;
                stg_stream4(g4 + i, o);
            }
            const int tail0 = head + 4 * n4;
            if (tail0 + lane < d.V) grow[tail0 + lane] = ex2_approx(fmaf(row[tail0 + lane], LOG2E_HI, c)) * sc;
        } else {
            for (int k = lane; k < d.V; k += 32) grow[k] = ex2_approx(fmaf(row[k], LOG2E_HI, c)) * sc;
        }
        __syncwarp();

        // ---- patch the label columns: subtract the merged posterior (:180-217, :290) ----
        if (lane == 0) {
            const float p = ex2_approx(fmaf(__ldg(row + d.blank), LOG2E_HI, c));
            grow[d.blank] = (p - blank_part) * sc;
        }
        __syncwarp();
        const int *usym = usym_all + (size_t)b * w.Nmax;
        const int *uoff = uoff_all + (size_t)b * (w.Nmax + 1);
        const int *unode = unode_all + (size_t)b * w.Nmax;
        for (int u = lane; u < ui.Ub; u += 32) {
            const int sym = usym[u];
            const int n0 = uoff[u], n1 = uoff[u + 1];
            float post = 0.f;
            for (int n = n0; n < n1; ++n) {
                const int j = unode[n];
                if (j < ui.Nb) post += e_sm[j];
            }
            if (sym == d.blank) post += blank_part;               // a label equal to the blank id
            const float p = ex2_approx(fmaf(__ldg(row + sym), LOG2E_HI, c));
            grow[sym] = (p - post) * sc;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) loss_sum_kernel(const float *loss, int B, float *out) {
    __shared__ double part[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) acc += (double)loss[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)part[0];
}

}  // namespace

cudaError_t launch_loss_sum(const float *loss_per_utt, int B, float *loss_sum, cudaStream_t stream) {
    loss_sum_kernel<<<1, 256, 0, stream>>>(loss_per_utt, B, loss_sum);
    return cudaGetLastError();
}

cudaError_t launch_gradient(const GradParams &g, const WsLayout &w, const void *ws, cudaStream_t stream) {
    const long long frames = (long long)g.d.B * g.d.T;
    if (frames == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_warp = w.Np;
    const size_t smem = sizeof(float) * (size_t)per_warp * kWarpsPerCta;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gradient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long ctas = (frames + kWarpsPerCta - 1) / kWarpsPerCta;
    const long long cap = (long long)sms * 8;
    if (ctas > cap) ctas = cap;
    const int b_major = g.gstride_b > g.gstride_t ? 1 : 0;
    gradient_kernel<<<(int)ctas, kWarpsPerCta * 32, smem, stream>>>(g, w, static_cast<const unsigned char *>(ws),
                                                                    b_major, per_warp);
    return cudaGetLastError();
}

}  // namespace b200ctc
