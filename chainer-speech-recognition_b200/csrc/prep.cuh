// prep.cuh -- symbol tables for the gradient kernel (device code shared with the lattice launch).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace b200ctc {

// ---------------------------------------------------------------------------------------------
// per-utterance bookkeeping for the gradient kernel (runs as extra CTAs of the lattice launch) -- which vocabulary ids can the lattice
// emit, and which nodes share an id (the merge of _compute_label_probability, gram_ctc.py:180-217).
// One CTA per utterance.  Output, per utterance:
//   usym[u]            distinct emitted ids, sorted ascending, blank included (Ub of them)
//   uoff[u]..uoff[u+1] range in unode[] listing the non-blank-type nodes that carry id usym[u]
//                      (the blank-type nodes -- every 2nd / 3rd node -- are summed directly)
//   urec[u]            {usym[u], uoff[u], node count, first node or -1}: one 16-byte load gives the gradient kernel all
//                      it needs for an id carried by a single node (the common case) -- no chain of dependent loads
//   bm[], pc[]         V-bit bitmap of emitted ids and, per 32-bit word, the number of set bits before it,
//                      so that "posterior slot of column k" = pc[k/32] + popc(bm[k/32] & low bits).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prep_utterance(const ProblemDesc &d, const WsLayout &w, unsigned char *ws, int b, int *sm) {
    __shared__ int s_U;
    UttInfo *ui = reinterpret_cast<UttInfo *>(ws + w.off_utt) + b;
    int *usym = reinterpret_cast<int *>(ws + w.off_usym) + (size_t)b * w.Nmax;
    int *uoff = reinterpret_cast<int *>(ws + w.off_uoff) + (size_t)b * (w.Nmax + 1);
    int *unode = reinterpret_cast<int *>(ws + w.off_unode) + (size_t)b * w.Nmax;
    unsigned *bm_g = reinterpret_cast<unsigned *>(ws + w.off_bm) + (size_t)b * w.nwords;
    int *pc_g = reinterpret_cast<int *>(ws + w.off_pc) + (size_t)b * w.nwords;

    int Tb = d.input_lengths ? d.input_lengths[b] : d.T;
    int Lb = d.label_lengths ? d.label_lengths[b] : d.Lmax;
    int flags = 0;
    if (Tb < 0 || Tb > d.T) { flags |= 1; Tb = max(0, min(Tb, d.T)); }
    if (Lb < 0 || Lb > d.Lmax) { flags |= 1; Lb = max(0, min(Lb, d.Lmax)); }
    const int per = d.kind == 0 ? 2 : 3;
    const int Nb = per * Lb + 1;
    // entry 0 = the blank id itself; entries 1..M = the non-blank-type nodes in node order
    // (CTC: label i -> node 2i+1; Gram: unigram i -> node 3i+1, bigram i -> node 3i+2).
    const int M = (per - 1) * Lb;
    const int E = M + 1;
    int *esym = sm;               // [E] symbol (or -1)
    int *efirst = sm + E;         // [E] first entry with the same symbol
    int *epos = sm + 2 * E;       // [E] number of earlier node entries with the same symbol
    int *erank = sm + 3 * E;      // [E] sorted position of a first entry's symbol
    int *cnt = sm + 4 * E;        // [E] node-list length per distinct symbol
    unsigned *bm = reinterpret_cast<unsigned *>(sm + 5 * E);   // [nwords]
    const int32_t *lab = d.labels + (size_t)b * d.Lmax;
    const int32_t *big = d.kind == 1 ? d.bigrams + (size_t)b * d.Lmax : nullptr;
    if (threadIdx.x == 0) s_U = 0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        int s;
        if (e == 0) s = d.blank;
        else if (d.kind == 0) s = lab[e - 1];
        else s = ((e - 1) & 1) ? big[(e - 1) >> 1] : lab[(e - 1) >> 1];
        if (s < 0 || s >= d.V) s = -1;                     // dead bigram (or an id outside the vocabulary)
        esym[e] = s;
        cnt[e] = 0;
    }
    for (int i = threadIdx.x; i < w.nwords; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int s = esym[e];
        int first = e, pos = 0;
        bool found = false;
        if (s >= 0) {
            for (int e2 = 0; e2 < e; ++e2)
                if (esym[e2] == s) {
                    if (!found) { first = e2; found = true; }
                    if (e2 >= 1) ++pos;
                }
        }
        efirst[e] = first;
        epos[e] = pos;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int s = esym[e];
        if (s < 0 || efirst[e] != e) continue;
        int r = 0;
        for (int e2 = 0; e2 < E; ++e2)
            if (esym[e2] >= 0 && efirst[e2] == e2 && esym[e2] < s) ++r;
        erank[e] = r;
        usym[r] = s;
        atomicAdd(&s_U, 1);
        atomicOr(&bm[s >> 5], 1u << (s & 31));
    }
    __syncthreads();
    for (int e = 1 + threadIdx.x; e < E; e += blockDim.x)
        if (esym[e] >= 0) atomicAdd(&cnt[erank[efirst[e]]], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int U = s_U;
        int acc = 0;
        for (int u = 0; u < U; ++u) { uoff[u] = acc; acc += cnt[u]; }
        uoff[U] = acc;
        ui->Tb = Tb; ui->Lb = Lb; ui->Nb = Nb; ui->Ub = U;
        ui->flags = flags; ui->ublank = erank[0];
    }
    if (threadIdx.x == 32) {
        int acc = 0;
        for (int i = 0; i < w.nwords; ++i) { pc_g[i] = acc; bm_g[i] = bm[i]; acc += __popc(bm[i]); }
    }
    __syncthreads();
    for (int e = 1 + threadIdx.x; e < E; e += blockDim.x) {
        if (esym[e] < 0) continue;
        const int en = e - 1;
        const int node = d.kind == 0 ? (2 * en + 1) : (3 * (en >> 1) + 1 + (en & 1));
        unode[uoff[erank[efirst[e]]] + epos[e]] = node;
    }
    __syncthreads();
    int4 *urec = reinterpret_cast<int4 *>(ws + w.off_urec) + (size_t)b * w.Nmax;
    for (int u = threadIdx.x; u < s_U; u += blockDim.x)
        urec[u] = make_int4(usym[u], uoff[u], cnt[u], cnt[u] > 0 ? unode[uoff[u]] : -1);
}


inline size_t prep_smem_bytes(int kind, int Lmax, int nwords) {
    const int per = kind == 0 ? 1 : 2;
    return sizeof(int) * (5 * ((size_t)per * Lmax + 1) + (size_t)nwords);
}

}  // namespace b200ctc
