// lattice.cu -- alpha/beta recursion over the blank-extended label lattice (kernel 2).
//
// Replaces the two serial Python loops of the reference, _compute_transition_probability
// (asr/loss/gram_ctc.py:142-178): loop 1 (:153-156) = alpha, loop 2 (:171-175) = beta, each a dense
// (B,N,N) log-matmul per frame there.  Here the adjacency is walked as the banded structure it is
// (SURVEY.md section 8a "banded lattice spec"):
//   CTC      node q: self, q-1, and q-2 iff q is a label node whose label differs from the previous one
//   Gram-CTC forward  type 0 (blank): q, q-1, q-2        type 1 (unigram): q, q-1, q-2, q-3*
//                     type 2 (bigram): q, q-5, q-7, q-6* (dead when the bigram id is -1)
//            reversed type 0 (blank): q, q-1, q-5        type 1 (bigram): q, q-1, q-2, q-6*
//                     type 2 (unigram): q, q-2, q-3*, q-7          (* = only if the two ids differ)
//
// Mapping: one CTA per utterance, two warps.  Warp 0 runs alpha forward in time, warp 1 runs beta
// backward in time on the reversed lattice; each lane owns K consecutive nodes in registers, so a
// step needs no barrier and no shared-memory round trip for the state: only the 2 (CTC) or 7
// (Gram-CTC) boundary values come from the neighbouring lane by warp shuffle.  The per-frame rows
// of gathered label log-probs are staged in shared memory by 1-D bulk async copies (TMA engine)
// several frames ahead, completion tracked by mbarriers.
//
// Meet in the middle: alpha visits frames 0..mid first, beta visits frames T-1..mid first, each
// writing its values ("first visitor", array fv).  After one CTA barrier log P is known
// (sum over nodes of alpha_mid * beta_mid) and each warp continues over the other half as "second
// visitor": it reads the first visitor's row (again staged by bulk copy) and writes
// gamma[t][j] = alpha_t[j] + beta_t[j] - log P directly.  Every (frame, node) cell is therefore
// written once by each direction and the gradient kernel reads one float per cell.
//
// beta convention as in the reference (:171-175): beta_t EXCLUDES the emission at t, so
// alpha_t + beta_t sums (in the log-sum-exp sense) to log P at every valid frame.
#include "common.cuh"
#include "kernels.h"

namespace b200ctc {

namespace {

constexpr int kStages = 3;

template <int K, bool GRAM>
struct LaneState {
    float h[K], l[K];        // split-log2 value of the K nodes this lane owns
    int ci[K];               // column of each node's symbol in the emission row (0 = blank)
    uint64_t flag;           // bit r: the "ids differ" edge into slot r is open
    uint64_t dead;           // bit r: dead node (bigram id -1, gram_ctc.py:94-98)
    uint64_t valid;          // bit r: node index < Nb
};

struct Pipe {
    uint64_t *bar;           // [kStages]
    float2 *lp;              // [kStages][C][W]
    float2 *fv;              // [kStages][C][Np]
    uint32_t parity;         // bit s: parity to wait for on stage s
};

struct UttCtx {
    const float2 *lp_g;      // this utterance's emission rows   [T][W]
    float2 *fv_g;            // first-visitor rows               [T][Np]
    float *gam_g;            // gamma rows                       [T][Np]
    int W, Np, C, Nb;
    float Ph, Pl;
};

// ---- one lattice step: new (pre-emission) value of slot r from the extended old array ----
// eh/el hold nodes [K*lane - PAD, K*lane + K); index PAD + r is slot r itself.
template <int K, bool GRAM, bool REV, int R>
__device__ __forceinline__ void slot_update(const float (&eh)[K + (GRAM ? 7 : 2)], const float (&el)[K + (GRAM ? 7 : 2)],
                                            uint64_t flag, uint64_t dead, float &pre_h, float &pre_l) {
    constexpr int PAD = GRAM ? 7 : 2;
    constexpr int I = PAD + R;
    const bool open = (flag >> R) & 1ull;
    if constexpr (!GRAM) {
        if constexpr ((R & 1) == 0) {                         // blank node: self, q-1
            const float hm = fmaxf(eh[I], eh[I - 1]);
            const float d0 = (eh[I] - hm) + el[I];
            const float d1 = (eh[I - 1] - hm) + el[I - 1];
            pre_h = hm;
            pre_l = lg2_approx(ex2_approx(d0) + ex2_approx(d1));
        } else {                                              // label node: self, q-1, q-2*
            const float h2 = open ? eh[I - 2] : SENT;
            const float hm = fmaxf(fmaxf(eh[I], eh[I - 1]), h2);
            const float d0 = (eh[I] - hm) + el[I];
            const float d1 = (eh[I - 1] - hm) + el[I - 1];
            const float d2 = (h2 - hm) + el[I - 2];
            pre_h = hm;
            pre_l = lg2_approx(ex2_approx(d0) + ex2_approx(d1) + ex2_approx(d2));
        }
    } else {
        constexpr int TYPE = R % 3;
        // predecessor offsets per (direction, type); the starred one is gated by `open`
        constexpr int A = (TYPE == 0) ? 1 : (!REV ? (TYPE == 1 ? 1 : 5) : (TYPE == 1 ? 1 : 2));
        constexpr int Bk = (TYPE == 0) ? (!REV ? 2 : 5) : (!REV ? (TYPE == 1 ? 2 : 7) : (TYPE == 1 ? 2 : 7));
        constexpr int Cs = (TYPE == 0) ? 0 : (!REV ? (TYPE == 1 ? 3 : 6) : (TYPE == 1 ? 6 : 3));
        float hm = fmaxf(fmaxf(eh[I], eh[I - A]), eh[I - Bk]);
        float hc = SENT, lc = 0.f;
        if constexpr (Cs != 0) {
            hc = open ? eh[I - Cs] : SENT;
            lc = el[I - Cs];
            hm = fmaxf(hm, hc);
        }
        const float d0 = (eh[I] - hm) + el[I];
        const float d1 = (eh[I - A] - hm) + el[I - A];
        const float d2 = (eh[I - Bk] - hm) + el[I - Bk];
        float s = ex2_approx(d0) + ex2_approx(d1) + ex2_approx(d2);
        if constexpr (Cs != 0) s += ex2_approx((hc - hm) + lc);
        pre_h = hm;
        pre_l = lg2_approx(s);
        // a bigram node sits at type 2 forward / type 1 reversed
        constexpr bool CAN_BE_DEAD = (!REV && TYPE == 2) || (REV && TYPE == 1);
        if constexpr (CAN_BE_DEAD) {
            if ((dead >> R) & 1ull) { pre_h = SENT; pre_l = 0.f; }
        }
    }
}

template <int K, bool GRAM, bool REV, bool SECOND, int R>
struct SlotLoop {
    __device__ __forceinline__ static void run(LaneState<K, GRAM> &st, const float (&eh)[K + (GRAM ? 7 : 2)],
                                               const float (&el)[K + (GRAM ? 7 : 2)], const float2 *lprow,
                                               const float2 *fvrow, float2 *fv_out, float *gam_out, int jbase,
                                               float Ph, float Pl, bool store) {
        float ph, pl;
        slot_update<K, GRAM, REV, R>(eh, el, st.flag, st.dead, ph, pl);
        const float2 e = lprow[st.ci[R]];
        // post-emission value, renormalised so that hi stays integer-valued and |lo| <= 0.5
        float nl = pl + e.y;
        float nh = ph + e.x;
        const float rr = rint_small(nl);
        nh += rr;
        nl -= rr;
        st.h[R] = nh;
        st.l[R] = nl;
        // node index in forward coordinates
        const int j = REV ? (jbase - R) : (jbase + R);
        const bool ok = (st.valid >> R) & 1ull;
        // alpha keeps the emission at t, beta excludes it (gram_ctc.py:171-175)
        const float vh = REV ? ph : nh;
        const float vl = REV ? pl : nl;
        if constexpr (!SECOND) {
            if (ok && store) fv_out[j] = make_float2(vh, vl);
        } else {
            if (ok) {
                const float2 o = fvrow[j];
                gam_out[j] = ((vh + o.x) - Ph) + ((vl + o.y) - Pl);
            }
        }
        if constexpr (R + 1 < K)
            SlotLoop<K, GRAM, REV, SECOND, R + 1>::run(st, eh, el, lprow, fvrow, fv_out, gam_out, jbase, Ph, Pl, store);
    }
};

template <int K, bool GRAM>
__device__ __forceinline__ void gather_neighbours(const LaneState<K, GRAM> &st, float (&eh)[K + (GRAM ? 7 : 2)],
                                                  float (&el)[K + (GRAM ? 7 : 2)], int lane) {
    constexpr int PAD = GRAM ? 7 : 2;
#pragma unroll
    for (int k = 0; k < PAD; ++k) {
        // node K*lane - PAD + k lives in the previous lane's slot K - PAD + k
        float nh = __shfl_up_sync(0xffffffffu, st.h[K - PAD + k], 1);
        float nl = __shfl_up_sync(0xffffffffu, st.l[K - PAD + k], 1);
        eh[k] = (lane == 0) ? SENT : nh;
        el[k] = (lane == 0) ? 0.f : nl;
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
        eh[PAD + r] = st.h[r];
        el[PAD + r] = st.l[r];
    }
}

// Visit n frames starting at f0 (ascending for alpha, descending for beta).
template <int K, bool GRAM, bool REV, bool SECOND>
__device__ __forceinline__ void run_phase(LaneState<K, GRAM> &st, Pipe &pipe, const UttCtx &c, int f0, int n,
                                       bool store_last, int lane) {
    if (n <= 0) return;
    const int C = c.C;
    const int nchunks = (n + C - 1) / C;
    const uint32_t row_lp_bytes = (uint32_t)c.W * 8u;
    const uint32_t row_fv_bytes = (uint32_t)c.Np * 8u;

    auto issue = [&](int chunk) {
        const int s = chunk % kStages;
        const int i0 = chunk * C;
        const int cnt = min(C, n - i0);
        const int flo = REV ? (f0 - i0 - cnt + 1) : (f0 + i0);
        if (lane == 0) {
            const uint32_t bl = row_lp_bytes * cnt;
            const uint32_t bf = SECOND ? row_fv_bytes * cnt : 0u;
            mbar_arrive_expect_tx(&pipe.bar[s], bl + bf);
            bulk_g2s(pipe.lp + (size_t)s * C * c.W, c.lp_g + (size_t)flo * c.W, bl, &pipe.bar[s]);
            if (SECOND) bulk_g2s(pipe.fv + (size_t)s * C * c.Np, c.fv_g + (size_t)flo * c.Np, bf, &pipe.bar[s]);
        }
    };

    for (int ch = 0; ch < min(kStages, nchunks); ++ch) issue(ch);

    const int jbase = REV ? (c.Nb - 1 - K * lane) : (K * lane);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int s = ch % kStages;
        const int i0 = ch * C;
        const int cnt = min(C, n - i0);
        mbar_wait(&pipe.bar[s], (pipe.parity >> s) & 1u);
        pipe.parity ^= (1u << s);
        const float2 *lp_s = pipe.lp + (size_t)s * C * c.W;
        const float2 *fv_s = pipe.fv + (size_t)s * C * c.Np;
        for (int i = 0; i < cnt; ++i) {
            const int row = REV ? (cnt - 1 - i) : i;
            const int f = REV ? (f0 - i0 - i) : (f0 + i0 + i);
            float eh[K + (GRAM ? 7 : 2)], el[K + (GRAM ? 7 : 2)];
            gather_neighbours<K, GRAM>(st, eh, el, lane);
            const bool store = store_last || (i0 + i + 1 < n);
            SlotLoop<K, GRAM, REV, SECOND, 0>::run(st, eh, el, lp_s + (size_t)row * c.W, fv_s + (size_t)row * c.Np,
                                                   c.fv_g + (size_t)f * c.Np, c.gam_g + (size_t)f * c.Np, jbase, c.Ph,
                                                   c.Pl, store);
        }
        __syncwarp();
        if (ch + kStages < nchunks) issue(ch + kStages);
    }
}

template <int K, bool GRAM>
__device__ __forceinline__ void init_lane(LaneState<K, GRAM> &st, const LatticeParams &p, int b, int Lb, int Nb,
                                          bool rev, int lane) {
    const int32_t *lab = p.labels + (size_t)b * p.Lmax;
    const int32_t *big = GRAM ? p.bigrams + (size_t)b * p.Lmax : nullptr;
    st.flag = 0ull;
    st.dead = 0ull;
    st.valid = 0ull;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const int q = K * lane + r;
        const int j = rev ? (Nb - 1 - q) : q;            // forward node index
        const bool ok = (q < Nb);
        st.h[r] = (q == 0) ? 0.f : SENT;                 // virtual state: all mass on node 0 (gram_ctc.py:144)
        st.l[r] = 0.f;
        if (ok) st.valid |= (1ull << r);
        int ci = 0;
        if (ok) {
            if constexpr (!GRAM) {
                if (j & 1) {
                    const int i = j >> 1;
                    ci = 1 + i;
                    bool open;
                    if (!rev) open = (i >= 1) && (lab[i] != lab[i - 1]);
                    else open = (i + 1 < Lb) && (lab[i + 1] != lab[i]);
                    if (open) st.flag |= (1ull << r);
                }
            } else {
                const int i = j / 3, type = j % 3;
                if (type == 1) {
                    ci = 1 + i;
                    bool open;
                    if (!rev) open = (i >= 1) && (lab[i] != lab[i - 1]);
                    else open = (i + 1 < Lb) && (lab[i + 1] != lab[i]);
                    if (open) st.flag |= (1ull << r);
                } else if (type == 2) {
                    ci = 1 + p.Lmax + i;
                    bool open;
                    if (!rev) open = (i >= 2) && (big[i] != big[i - 2]);
                    else open = (i + 2 < Lb) && (big[i + 2] != big[i]);
                    if (open) st.flag |= (1ull << r);
                    if (big[i] == -1) st.dead |= (1ull << r);
                }
            }
        }
        st.ci[r] = ci;
    }
}

template <int K, bool GRAM>
__global__ void __launch_bounds__(64, 1) lattice_kernel(LatticeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    UttInfo *ui = p.utt + b;
    const int Tb = ui->Tb, Lb = ui->Lb, Nb = ui->Nb;

    // shared memory carve-up: per direction [kStages][C][W] + [kStages][C][Np] float2, then barriers, then P
    const size_t lp_elems = (size_t)kStages * p.C * p.W;
    const size_t fv_elems = (size_t)kStages * p.C * p.Np;
    float2 *base = reinterpret_cast<float2 *>(smem_raw);
    Pipe pipe;
    pipe.lp = base + (size_t)warp * (lp_elems + fv_elems);
    pipe.fv = pipe.lp + lp_elems;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + 2 * (lp_elems + fv_elems));
    pipe.bar = bars + warp * kStages;
    pipe.parity = 0u;
    float *pshare = reinterpret_cast<float *>(bars + 2 * kStages);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * kStages; ++i) mbar_init(&bars[i], 1);
        mbar_init_fence();
    }
    __syncthreads();

    if (Tb <= 0) {                                       // no frames: P = 1 iff there is nothing to emit
        if (threadIdx.x == 0) {
            const bool feas = (Lb == 0);
            ui->Ph = feas ? 0.f : -SENT;
            ui->Pl = 0.f;
            ui->loss = feas ? 0.f : 1e10f;
            ui->flags |= feas ? 0 : 2;
            p.loss_per_utt[b] = ui->loss;
        }
        return;
    }

    UttCtx c;
    c.lp_g = p.lp + (size_t)b * p.T * p.W;
    c.fv_g = p.fv + (size_t)b * p.T * p.Np;
    c.gam_g = p.gam + (size_t)b * p.T * p.Np;
    c.W = p.W; c.Np = p.Np; c.C = p.C; c.Nb = Nb;
    c.Ph = 0.f; c.Pl = 0.f;

    const int mid = (Tb - 1) >> 1;
    LaneState<K, GRAM> st;
    init_lane<K, GRAM>(st, p, b, Lb, Nb, warp == 1, lane);

    // ---- phase A: first visitors ----
    if (warp == 0) run_phase<K, GRAM, false, false>(st, pipe, c, 0, mid + 1, /*store_last=*/false, lane);
    else           run_phase<K, GRAM, true, false>(st, pipe, c, Tb - 1, Tb - mid, /*store_last=*/true, lane);
    fence_proxy_async();
    __threadfence_block();
    __syncthreads();

    // ---- log P at the meeting frame: LSE_j(alpha_mid[j] + beta_mid[j]) ----
    if (warp == 0) {
        const float2 *brow = c.fv_g + (size_t)mid * p.Np;
        float bh[K], bl[K];
        float pm = SENT;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            const int j = K * lane + r;
            float2 o = make_float2(SENT, 0.f);
            if (j < Nb) o = __ldcg(brow + j);
            bh[r] = o.x; bl[r] = o.y;
            pm = fmaxf(pm, st.h[r] + o.x);
        }
        pm = warp_max(pm);
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < K; ++r) s += ex2_approx(((st.h[r] + bh[r]) - pm) + (st.l[r] + bl[r]));
        s = warp_sum(s);
        const bool feas = (pm > SENT_TEST) && (s > 0.f);
        float Ph, Pl;
        if (feas) {
            const float lg = log2f(s);
            const float rr = rintf(lg);
            Ph = pm + rr;
            Pl = lg - rr;
        } else {
            Ph = -SENT;                                  // +1e30: every gamma becomes log 0
            Pl = 0.f;
        }
        // gamma at the meeting frame
#pragma unroll
        for (int r = 0; r < K; ++r) {
            const int j = K * lane + r;
            if (j < Nb) c.gam_g[(size_t)mid * p.Np + j] = ((st.h[r] + bh[r]) - Ph) + ((st.l[r] + bl[r]) - Pl);
        }
        if (lane == 0) {
            pshare[0] = Ph;
            pshare[1] = Pl;
            const float loss = feas ? (float)(-((double)Ph + (double)Pl) * LN2_D) : 1e10f;
            ui->Ph = Ph;
            ui->Pl = Pl;
            ui->loss = loss;
            if (!feas) ui->flags |= 2;
            p.loss_per_utt[b] = loss;
        }
    }
    __syncthreads();
    c.Ph = pshare[0];
    c.Pl = pshare[1];
    fence_proxy_async();

    // ---- phase B: second visitors ----
    if (warp == 0) run_phase<K, GRAM, false, true>(st, pipe, c, mid + 1, Tb - 1 - mid, true, lane);
    else           run_phase<K, GRAM, true, true>(st, pipe, c, mid - 1, mid, true, lane);
}

template <int K, bool GRAM>
cudaError_t launch_one(const LatticeParams &p, size_t smem, cudaStream_t stream) {
    auto kern = lattice_kernel<K, GRAM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.B, 64, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

int lattice_max_nodes(int kind) { return kind == 0 ? 32 * 24 : 32 * 36; }

cudaError_t launch_lattice(int kind, LatticeParams p, int Nmax, cudaStream_t stream, int *status) {
    *status = 0;
    // chunk length: as many frames per bulk copy as fit ~96 KB per direction, at most 16
    const size_t per_frame = (size_t)(p.W + p.Np) * sizeof(float2);
    int C = (int)((size_t)(100 * 1024) / ((size_t)kStages * per_frame));
    if (C > 16) C = 16;
    if (C < 1) { *status = 2; return cudaSuccess; }
    p.C = C;
    const size_t smem = 2 * (size_t)kStages * C * per_frame + 2 * kStages * sizeof(uint64_t) + 16;
    const int need = (Nmax + 31) / 32;
    if (kind == 0) {
        if (need <= 2) return launch_one<2, false>(p, smem, stream);
        if (need <= 4) return launch_one<4, false>(p, smem, stream);
        if (need <= 6) return launch_one<6, false>(p, smem, stream);
        if (need <= 8) return launch_one<8, false>(p, smem, stream);
        if (need <= 12) return launch_one<12, false>(p, smem, stream);
        if (need <= 16) return launch_one<16, false>(p, smem, stream);
        if (need <= 24) return launch_one<24, false>(p, smem, stream);
    } else {
        if (need <= 9) return launch_one<9, true>(p, smem, stream);
        if (need <= 12) return launch_one<12, true>(p, smem, stream);
        if (need <= 18) return launch_one<18, true>(p, smem, stream);
        if (need <= 24) return launch_one<24, true>(p, smem, stream);
        if (need <= 36) return launch_one<36, true>(p, smem, stream);
    }
    *status = 2;
    return cudaSuccess;
}

}  // namespace b200ctc
