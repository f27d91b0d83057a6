/* oracle/ldpc_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the one hot path of thadikari/ldpc_decoders:
 *   - flooding BP, min-sum and sum-product:  src/bpa.py:17-102, src/math_utils.py:5-94
 *   - BEC erasure message passing:           src/bec.py:70-122
 *   - channel LLR front ends:                src/bsc.py:21,25  src/biawgn.py:10,28  src/bec.py:76,85
 * It is the checker the CUDA path is compared with.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product package never does (see oracle/README.md).
 *
 * Parity status: PINNED — tests/test_oracle_golden.py checks this file against
 * fixtures generated from the unmodified reference (tests/golden/make_golden.py):
 * the six Test.sample KATs, seeded multi-frame runs and teacher-forced SPA sweeps.
 *
 * Graph tables (built by oracle/oracle.py from np.where(H), src/bpa.py:12):
 *   edge e = position in row-major (check-major, ascending variable) order
 *   chk_ptr[m+1], edge_var[E]    check-major CSR
 *   var_ptr[n+1], var_edges[E]   per variable, its edge ids ascending (= ascending check)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

#define ORACLE_MSA 0
#define ORACLE_SPA 1

typedef struct {
    int n, m, E;
    const int32_t *chk_ptr, *edge_var, *var_ptr, *var_edges;
} graph_t;

/* src/bpa.py:29  ((H @ x_hat) % 2 == 0).all()  on hard bits */
static int syndrome_is_zero(const graph_t *g, const uint8_t *x)
{
    for (int c = 0; c < g->m; ++c) {
        unsigned p = 0;
        for (int e = g->chk_ptr[c]; e < g->chk_ptr[c + 1]; ++e) p ^= x[g->edge_var[e]];
        if (p & 1u) return 0;
    }
    return 1;
}


/* Frames are independent: split [0,B) into chunks handed out through an atomic
 * counter to nthreads POSIX threads (the reference's own parallel model is N
 * independent processes, run_sims.sh:15). */
typedef int (*range_fn)(void *ctx, int b0, int b1);
typedef struct { range_fn fn; void *ctx; int B, chunk; atomic_int next; atomic_int err; } pf_t;

static void *pf_worker(void *p)
{
    pf_t *s = (pf_t *)p;
    for (;;) {
        const int b0 = atomic_fetch_add(&s->next, s->chunk);
        if (b0 >= s->B) break;
        const int b1 = b0 + s->chunk < s->B ? b0 + s->chunk : s->B;
        const int rc = s->fn(s->ctx, b0, b1);
        if (rc) atomic_store(&s->err, rc);
    }
    return NULL;
}

static int parallel_frames(int B, int nthreads, range_fn fn, void *ctx)
{
    if (B <= 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    if (nthreads == 1) return fn(ctx, 0, B);
    pf_t s;
    s.fn = fn; s.ctx = ctx; s.B = B;
    s.chunk = B / (nthreads * 8) > 0 ? B / (nthreads * 8) : 1;
    atomic_init(&s.next, 0);
    atomic_init(&s.err, 0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return -2;
    int started = 0;
    for (int i = 0; i < nthreads; ++i) {
        if (pthread_create(&th[i], NULL, pf_worker, &s) != 0) break;
        ++started;
    }
    if (started == 0) { free(th); return fn(ctx, 0, B); }
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
    return atomic_load(&s.err);
}

#define T double
#define SFX(name) name##_f64
#define TANH tanh
#define LOG log
#define EXP exp
#define ATANH atanh
#define FABS fabs
#include "bp_body.inc"
#undef T
#undef SFX
#undef TANH
#undef LOG
#undef EXP
#undef ATANH
#undef FABS

#define T float
#define SFX(name) name##_f32
#define TANH tanhf
#define LOG logf
#define EXP expf
#define ATANH atanhf
#define FABS fabsf
#include "bp_body.inc"
#undef T
#undef SFX
#undef TANH
#undef LOG
#undef EXP
#undef ATANH
#undef FABS

/* ---------------------------------------------------------------------------
 * BEC: literal integer message passing, src/bec.py:83-122.
 * symbols {0,1,2 = erasure}; messages [-1,+1,0][symbol] (bec.py:76,85).
 * reason: 0 decoded (bec.py:97), 1 maximum (bec.py:96), 2 stopping (bec.py:120), 4 cap.
 * ------------------------------------------------------------------------- */
static inline int isgn(int a) { return (a > 0) - (a < 0); }

static void bec_decode_one(const graph_t *g, const uint8_t *y, int max_iter, int iter_cap,
                           uint8_t *x_hat, int32_t *iters_out, uint8_t *reason_out, int *work /* [2E+2n] */)
{
    static const int msg_of_symbol[3] = { -1, 1, 0 };
    int *v2c = work, *c2v = work + g->E, *prior = work + 2 * (size_t)g->E, *marg = prior + g->n;
    for (int v = 0; v < g->n; ++v) prior[v] = msg_of_symbol[y[v] > 2 ? 2 : y[v]];
    for (int e = 0; e < g->E; ++e) { v2c[e] = prior[g->edge_var[e]]; c2v[e] = 0; }   /* bec.py:86 */
    memcpy(x_hat, y, (size_t)g->n);                                                /* bec.py:89 */
    int it = 0;
    uint8_t reason;
    for (;;) {
        if (0 < max_iter && max_iter <= it) { reason = 1; break; }                 /* bec.py:96 */
        if (iter_cap > 0 && it >= iter_cap) { reason = 4; break; }
        int erased = 0;
        for (int v = 0; v < g->n; ++v) erased += (x_hat[v] == 2);
        if (erased == 0) { reason = 0; break; }                                    /* bec.py:97 */

        for (int c = 0; c < g->m; ++c) {                                           /* bec.py:100-112 */
            const int e0 = g->chk_ptr[c], e1 = g->chk_ptr[c + 1];
            int sums = 0, incoming = 0;
            for (int e = e0; e < e1; ++e) { sums += 1 - abs(v2c[e]); incoming += (v2c[e] > 0); }
            for (int e = e0; e < e1; ++e) {
                if (sums == 0) c2v[e] = v2c[e];
                else if (sums > 1) c2v[e] = 0;
                else c2v[e] = (1 - abs(v2c[e])) * (2 * (incoming % 2) - 1);
            }
        }
        for (int v = 0; v < g->n; ++v) {                                           /* bec.py:115 */
            int s = 0;
            for (int k = g->var_ptr[v]; k < g->var_ptr[v + 1]; ++k) s += c2v[g->var_edges[k]];
            marg[v] = prior[v] + s;
        }
        for (int e = 0; e < g->E; ++e) v2c[e] = isgn(marg[g->edge_var[e]] - c2v[e]);   /* bec.py:116 */
        int same = 1;
        for (int v = 0; v < g->n; ++v) {                                           /* bec.py:119 symbols[sign] */
            const int s = isgn(marg[v]);
            const uint8_t xn = (s == 0) ? 2 : (s > 0 ? 1 : 0);
            marg[v] = xn;
            same &= (xn == x_hat[v]);
        }
        if (same) { reason = 2; break; }                                           /* bec.py:120 */
        for (int v = 0; v < g->n; ++v) x_hat[v] = (uint8_t)marg[v];                /* bec.py:121 */
        ++it;
    }
    *iters_out = it;
    if (reason_out) *reason_out = reason;
}

typedef struct {
    const graph_t *g; const uint8_t *y; int max_iter, iter_cap;
    uint8_t *x_hat; int32_t *iters; uint8_t *reason;
} bec_job;

static int bec_range(void *ctx, int b0, int b1)
{
    const bec_job *j = (const bec_job *)ctx;
    const graph_t *g = j->g;
    const size_t n = (size_t)g->n;
    int *work = (int *)malloc(sizeof(int) * (2 * (size_t)g->E + 2 * n));
    if (!work) return -2;
    for (int b = b0; b < b1; ++b)
        bec_decode_one(g, j->y + b * n, j->max_iter, j->iter_cap, j->x_hat + b * n,
                       j->iters + b, j->reason ? j->reason + b : NULL, work);
    free(work);
    return 0;
}

int oracle_bec(int n, int m, int E, const int32_t *chk_ptr, const int32_t *edge_var,
               const int32_t *var_ptr, const int32_t *var_edges,
               int B, const uint8_t *y /* [B,n] in {0,1,2} */, int max_iter, int iter_cap,
               uint8_t *x_hat /* [B,n] */, int32_t *iters, uint8_t *reason, int nthreads)
{
    graph_t g = { n, m, E, chk_ptr, edge_var, var_ptr, var_edges };
    bec_job job = { &g, y, max_iter, iter_cap, x_hat, iters, reason };
    return parallel_frames(B, nthreads, bec_range, &job);
}

/* ---------------------------------------------------------------------------
 * Channel LLR front ends (float64, exactly the reference's expressions).
 * ------------------------------------------------------------------------- */
/* src/bsc.py:21,25   llr = log(1-p) - log(p);  priors = llr * (1 - 2*y)
 * llr is passed in (computed by numpy's log in oracle.py, like the reference). */
void oracle_llr_bsc(double llr, size_t count, const uint8_t *y, double *priors)
{
    for (size_t i = 0; i < count; ++i) priors[i] = llr * (double)(1 - 2 * (int)y[i]);
}

/* src/biawgn.py:10,28   nv = 10 ** (-snr_db / 10);  priors = (-2 * y) / nv
 * noise_var is passed in (computed by numpy's pow in oracle.py so it is the
 * very same double the reference uses). */
void oracle_llr_biawgn(double noise_var, size_t count, const double *y, double *priors)
{
    for (size_t i = 0; i < count; ++i) priors[i] = (-2.0 * y[i]) / noise_var;
}

int oracle_abi_version(void) { return 1; }
