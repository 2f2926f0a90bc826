/*
 * oracle/ctc_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C, float64, CPU restatement of the reference CTC / Gram-CTC loss forward+backward
 * (musyoku/chainer-speech-recognition, asr/loss/gram_ctc.py).  It is the checker for the CUDA
 * path and the "port" CPU baseline that bench.py times on the GPU box's host cores.  It is never
 * linked into, loaded by, or called from the product library.
 *
 * Follows, in the reference:
 *   node symbols          _label_to_path                        asr/loss/gram_ctc.py:24-32
 *   forward edges         _create_forward_connection_matrix     :66-99
 *   backward edges        _create_backward_connection_matrix    :103-140 (= transpose of the forward set)
 *   alpha/beta            _compute_transition_probability       :142-178 (beta_t excludes the emission at t)
 *   loss                  GramCTC.forward                       :279-281
 *   gradient/scale/mask   GramCTC.backward                      :284-297
 *   per-symbol merge      _compute_label_probability            :180-217
 *   softmax, log          _softmax :18-21, _log_matrix :48-57
 * plain CTC = the same lattice with every bigram node dead (:95-98), i.e. the 2L+1 lattice.
 *
 * The reference evaluates each frame as a dense (B,N,N) log-matmul; its adjacency has <= 4
 * non-zeros per row, so this file walks the same edges directly (O(T*N) instead of O(T*N^2)).
 * That makes this port a *faster* CPU baseline than the reference's own NumPy path.
 * log(0) is -inf here (-1e10 in the reference, :222); an infeasible alignment is mapped back to the
 * reference's observable value, loss = 1e10, with a zero posterior.
 *
 * Parallelism: OpenMP over utterances (the reference is single-threaded NumPy).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NEG_INF (-INFINITY)
#define MAX_EDGES 4

typedef struct {
    int N;
    int *sym;           /* vocabulary id per node, -1 for a dead node            */
    int *npred;         /* number of predecessors                                 */
    int *pred;          /* [N][MAX_EDGES] predecessor node index                  */
    int nfinal;
    int final_nodes[3];
} lattice_t;

static inline double lse2(double a, double b) {
    if (a == NEG_INF) return b;
    if (b == NEG_INF) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

static void lattice_free(lattice_t *lat) {
    free(lat->sym); free(lat->npred); free(lat->pred);
}

/* classic blank-interleaved lattice: self, j-1, j-2 iff label node and label differs */
static void build_ctc(lattice_t *lat, const int *lab, int L, int blank) {
    int N = 2 * L + 1;
    lat->N = N;
    lat->sym = (int *)malloc(sizeof(int) * N);
    lat->npred = (int *)calloc(N, sizeof(int));
    lat->pred = (int *)malloc(sizeof(int) * N * MAX_EDGES);
    for (int j = 0; j < N; ++j) {
        int *p = lat->pred + j * MAX_EDGES, n = 0;
        lat->sym[j] = (j & 1) ? lab[j >> 1] : blank;
        p[n++] = j;
        if (j >= 1) p[n++] = j - 1;
        if ((j & 1) && j >= 3 && lab[j >> 1] != lab[(j >> 1) - 1]) p[n++] = j - 2;
        lat->npred[j] = n;
    }
    lat->nfinal = 0;
    lat->final_nodes[lat->nfinal++] = N - 1;
    if (N >= 2) lat->final_nodes[lat->nfinal++] = N - 2;
}

/* gram lattice, gram_ctc.py:66-99: node j, i=j/3, type=j%3 (0 blank, 1 unigram_i, 2 bigram_i) */
static void build_gram(lattice_t *lat, const int *uni, const int *bi, int L, int blank) {
    int N = 3 * L + 1;
    lat->N = N;
    lat->sym = (int *)malloc(sizeof(int) * N);
    lat->npred = (int *)calloc(N, sizeof(int));
    lat->pred = (int *)malloc(sizeof(int) * N * MAX_EDGES);
    for (int j = 0; j < N; ++j) {
        int i = j / 3, type = j % 3;
        lat->sym[j] = type == 0 ? blank : (type == 1 ? uni[i] : bi[i]);   /* bigram id -1 => dead */
    }
    for (int j = 0; j < N; ++j) {
        int i = j / 3, type = j % 3, n = 0;
        int *p = lat->pred + j * MAX_EDGES;
        int cand[4], nc = 0;
        if (lat->sym[j] < 0) { lat->npred[j] = 0; continue; }             /* :95-98 */
        cand[nc++] = j;                                                   /* :82 */
        if (type != 2) { cand[nc++] = j - 1; cand[nc++] = j - 2; }        /* :83 */
        if (type == 1 && i >= 1 && uni[i] != uni[i - 1]) cand[nc++] = j - 3;   /* :84 */
        if (type == 2) {
            cand[nc++] = j - 5;                                           /* :86 */
            if (i >= 2 && bi[i] != bi[i - 2]) cand[nc++] = j - 6;         /* :85 */
            cand[nc++] = j - 7;                                           /* :86 */
        }
        for (int c = 0; c < nc; ++c)
            if (cand[c] >= 0 && lat->sym[cand[c]] >= 0) p[n++] = cand[c];
        lat->npred[j] = n;
    }
    lat->nfinal = 0;
    for (int n = N - 1; n >= 0 && n >= N - 3; --n)
        if (lat->sym[n] >= 0) lat->final_nodes[lat->nfinal++] = n;
}

/*
 * kind: 0 = CTC, 1 = Gram-CTC.  x: element (t,b,v) at x[t*stride_t + b*stride_b + v], float32.
 * labels/bigrams: (B, Lmax) int32 row-major (bigrams ignored for kind 0).
 * loss_out: (B) float64, per-utterance loss.  grad_out (may be NULL): dense (T,B,V) float32,
 * d loss_b / d x, multiplied by grad_scale[b] if grad_scale != NULL, zero for t >= in_len[b].
 * gamma_out (may be NULL): (B, T, Nmax) float64, alpha+beta-logP (natural log), -inf where unreachable.
 * argmax_out (may be NULL): (B,T) int64 greedy indices over raw activations (first max wins).
 * Returns 0 on success.
 */
int ctc_oracle_run(int kind, const float *x, int64_t stride_t, int64_t stride_b,
                   int B, int T, int V, const int32_t *labels, const int32_t *bigrams, int Lmax,
                   const int32_t *in_len, const int32_t *lab_len, int blank,
                   double *loss_out, float *grad_out, const double *grad_scale,
                   double *gamma_out, int64_t *argmax_out, int nthreads) {
    if (kind != 0 && kind != 1) return 1;
    if (B < 0 || T < 0 || V <= 0 || blank < 0 || blank >= V) return 1;
    const int Nmax = (kind == 0 ? 2 : 3) * Lmax + 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    int status = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const int Tb = in_len ? in_len[b] : T;
        const int L = lab_len ? lab_len[b] : Lmax;
        if (Tb < 0 || Tb > T || L < 0 || L > Lmax) { status = 1; continue; }
        lattice_t lat;
        if (kind == 0) build_ctc(&lat, labels + (size_t)b * Lmax, L, blank);
        else build_gram(&lat, labels + (size_t)b * Lmax, bigrams + (size_t)b * Lmax, L, blank);
        const int N = lat.N;
        double *lse = (double *)malloc(sizeof(double) * (Tb > 0 ? Tb : 1));
        double *emit = (double *)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1) * N);
        double *alpha = (double *)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1) * N);
        double *beta = (double *)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1) * N);
        double *carry = (double *)malloc(sizeof(double) * N);
        double *post = (double *)malloc(sizeof(double) * V);

        /* log-softmax normaliser per frame (:18-21, :274) and greedy argmax */
        for (int t = 0; t < T; ++t) {
            const float *row = x + t * stride_t + b * stride_b;
            if (argmax_out) {
                int best = 0; float bv = row[0];
                for (int v = 1; v < V; ++v) {
                    if (bv != bv) break;                          /* NaN is maximal, first one wins */
                    if (row[v] > bv || row[v] != row[v]) { bv = row[v]; best = v; }
                }
                argmax_out[(size_t)b * T + t] = best;
            }
            if (t >= Tb) continue;
            double m = row[0];
            for (int v = 1; v < V; ++v) if (row[v] > m) m = row[v];
            double s = 0.0;
            for (int v = 0; v < V; ++v) s += exp((double)row[v] - m);
            lse[t] = m + log(s);
            for (int j = 0; j < N; ++j)
                emit[(size_t)t * N + j] = lat.sym[j] >= 0 ? (double)row[lat.sym[j]] - lse[t] : NEG_INF;
        }

        /* alpha, :153-156 */
        for (int j = 0; j < N; ++j) carry[j] = NEG_INF;
        carry[0] = 0.0;                                           /* :144 */
        for (int t = 0; t < Tb; ++t) {
            double *a = alpha + (size_t)t * N;
            for (int j = 0; j < N; ++j) {
                double acc = NEG_INF;
                const int *p = lat.pred + j * MAX_EDGES;
                for (int e = 0; e < lat.npred[j]; ++e) acc = lse2(acc, carry[p[e]]);
                a[j] = acc + emit[(size_t)t * N + j];
                if (lat.sym[j] < 0) a[j] = NEG_INF;
            }
            memcpy(carry, a, sizeof(double) * N);
        }
        double logP = NEG_INF;
        if (Tb > 0)
            for (int f = 0; f < lat.nfinal; ++f) logP = lse2(logP, alpha[(size_t)(Tb - 1) * N + lat.final_nodes[f]]);
        else if (L == 0) logP = 0.0;

        /* beta (excludes emission at t), :171-175; carry holds beta_{t+1}+emit_{t+1}; virtual end = final blank */
        for (int j = 0; j < N; ++j) carry[j] = NEG_INF;
        carry[N - 1] = 0.0;
        for (int t = Tb - 1; t >= 0; --t) {
            double *bt = beta + (size_t)t * N;
            for (int j = 0; j < N; ++j) bt[j] = NEG_INF;
            for (int j = 0; j < N; ++j) {                         /* scatter along the transposed edges */
                const int *p = lat.pred + j * MAX_EDGES;
                if (carry[j] == NEG_INF) continue;
                for (int e = 0; e < lat.npred[j]; ++e) bt[p[e]] = lse2(bt[p[e]], carry[j]);
            }
            for (int j = 0; j < N; ++j) carry[j] = bt[j] + emit[(size_t)t * N + j];
        }

        const int feasible = isfinite(logP);
        loss_out[b] = feasible ? -logP : 1e10;                    /* :279 ; reference quirk: exactly 1e10 */

        if (gamma_out)
            for (int t = 0; t < T; ++t)
                for (int j = 0; j < Nmax; ++j) {
                    double g = NEG_INF;
                    if (feasible && t < Tb && j < N) {
                        double a = alpha[(size_t)t * N + j], bb = beta[(size_t)t * N + j];
                        if (a != NEG_INF && bb != NEG_INF) g = a + bb - logP;
                    }
                    gamma_out[((size_t)b * T + t) * Nmax + j] = g;
                }

        if (grad_out) {
            const double scale = grad_scale ? grad_scale[b] : 1.0;
            for (int t = 0; t < T; ++t) {
                float *g = grad_out + ((size_t)t * B + b) * V;
                if (t >= Tb) { memset(g, 0, sizeof(float) * V); continue; }   /* :296 */
                const float *row = x + t * stride_t + b * stride_b;
                memset(post, 0, sizeof(double) * V);
                if (feasible)
                    for (int j = 0; j < N; ++j) {
                        double a = alpha[(size_t)t * N + j], bb = beta[(size_t)t * N + j];
                        if (lat.sym[j] >= 0 && a != NEG_INF && bb != NEG_INF)
                            post[lat.sym[j]] += exp(a + bb - logP);           /* :180-217, :290 */
                    }
                for (int v = 0; v < V; ++v)
                    g[v] = (float)((exp((double)row[v] - lse[t]) - post[v]) * scale);   /* :290-294 */
            }
        }
        free(lse); free(emit); free(alpha); free(beta); free(carry); free(post);
        lattice_free(&lat);
    }
    return status;
}

int ctc_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
