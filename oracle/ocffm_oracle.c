/* oracle/ocffm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see ocffm_oracle.h).
 *
 * A plain-C fp64 restatement of the reference algorithm.  Each function names the reference
 * lines it follows.  Storage differs from the reference on purpose (SoA CSR instead of
 * Node*; no thread-private scratch) -- only the arithmetic and its order of block updates
 * are the reference's.  Parity status: PINNED against dumps of the unmodified reference
 * (tests/test_oracle_vs_reference.py, tests/golden/).
 */
#include "ocffm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t rows, D, nnz;
    uint64_t *rowptr;
    uint32_t *idx;
    double *val;
    uint64_t *freq;
} csr_t;

typedef struct {
    uint64_t rows, nnz;
    uint64_t *rowptr;
    uint32_t *idx;
    double *val; /* y-tilde cache (Node::val of Y, ffm.cpp:393,400) */
} ycsr_t;

struct oc_problem {
    int fu, fv, f, k, self_side, freq;
    uint64_t m, n, mt, un; /* un = U->n = max label + 1 */
    double lambda, w, r;
    csr_t *XU, *XV, *XT;
    ycsr_t YU, YV, YT;
    double *popular;
    uint64_t *nnx_t;
    int nnx_given;
    int nr_blocks;
    double **W, **H, **P, **Q;
    double *a, *b, *sa, *sb;
    uint64_t cg_total;
};

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* index_vec, ffm.cpp:53-55 */
static int blk(const oc_problem *p, int f1, int f2) {
    return f2 + (p->f - 1) * f1 - f1 * (f1 - 1) / 2;
}
static int is_user(const oc_problem *p, int fg) { return fg < p->fu; }
static const csr_t *field_of(const oc_problem *p, int fg) {
    return is_user(p, fg) ? &p->XU[fg] : &p->XV[fg - p->fu];
}
static uint64_t rows_of(const oc_problem *p, int fg) { return is_user(p, fg) ? p->m : p->n; }
static int block_exists(const oc_problem *p, int f1, int f2) {
    /* ffm.cpp:502-503: with --ns only user-field x item-field blocks exist */
    if (p->self_side) return 1;
    return f1 < p->fu && f2 >= p->fu;
}
static int is_side_block(const oc_problem *p, int f1, int f2) {
    return (f1 < p->fu && f2 < p->fu) || (f1 >= p->fu && f2 >= p->fu);
}

oc_problem *oc_create(int fu, int fv, uint64_t m, uint64_t n, int k, double lambda, double omega,
                      double r, int self_side, int freq) {
    oc_problem *p = (oc_problem *)xcalloc(1, sizeof(*p));
    p->fu = fu; p->fv = fv; p->f = fu + fv; p->k = k;
    p->m = m; p->n = n; p->lambda = lambda; p->w = omega; p->r = r;
    p->self_side = self_side; p->freq = freq;
    p->XU = (csr_t *)xcalloc(fu, sizeof(csr_t));
    p->XV = (csr_t *)xcalloc(fv, sizeof(csr_t));
    p->XT = (csr_t *)xcalloc(fu, sizeof(csr_t));
    p->nr_blocks = p->f * (p->f + 1) / 2;
    p->W = (double **)xcalloc(p->nr_blocks, sizeof(double *));
    p->H = (double **)xcalloc(p->nr_blocks, sizeof(double *));
    p->P = (double **)xcalloc(p->nr_blocks, sizeof(double *));
    p->Q = (double **)xcalloc(p->nr_blocks, sizeof(double *));
    p->a = (double *)xcalloc(m, sizeof(double));
    p->b = (double *)xcalloc(n, sizeof(double));
    p->sa = (double *)xcalloc(m, sizeof(double));
    p->sb = (double *)xcalloc(n, sizeof(double));
    return p;
}

static void free_csr(csr_t *c) { free(c->rowptr); free(c->idx); free(c->val); free(c->freq); }
static void free_y(ycsr_t *y) { free(y->rowptr); free(y->idx); free(y->val); }

void oc_destroy(oc_problem *p) {
    if (!p) return;
    for (int i = 0; i < p->fu; i++) { free_csr(&p->XU[i]); free_csr(&p->XT[i]); }
    for (int i = 0; i < p->fv; i++) free_csr(&p->XV[i]);
    free(p->XU); free(p->XV); free(p->XT);
    free_y(&p->YU); free_y(&p->YV); free_y(&p->YT);
    for (int i = 0; i < p->nr_blocks; i++) { free(p->W[i]); free(p->H[i]); free(p->P[i]); free(p->Q[i]); }
    free(p->W); free(p->H); free(p->P); free(p->Q);
    free(p->a); free(p->b); free(p->sa); free(p->sb); free(p->popular); free(p->nnx_t);
    free(p);
}

void oc_set_field(oc_problem *p, int side, int fi, uint64_t rows, uint64_t D,
                  const uint64_t *rowptr, const uint32_t *idx, const double *val) {
    csr_t *c = side == OC_SIDE_U ? &p->XU[fi] : side == OC_SIDE_V ? &p->XV[fi] : &p->XT[fi];
    free_csr(c);
    c->rows = rows; c->D = D; c->nnz = rowptr[rows];
    c->rowptr = (uint64_t *)xcalloc(rows + 1, sizeof(uint64_t));
    c->idx = (uint32_t *)xcalloc(c->nnz, sizeof(uint32_t));
    c->val = (double *)xcalloc(c->nnz, sizeof(double));
    memcpy(c->rowptr, rowptr, (rows + 1) * sizeof(uint64_t));
    memcpy(c->idx, idx, c->nnz * sizeof(uint32_t));
    memcpy(c->val, val, c->nnz * sizeof(double));
    /* freq[idx] = number of occurrences of the feature (ffm.cpp:235-241) */
    c->freq = (uint64_t *)xcalloc(D, sizeof(uint64_t));
    for (uint64_t t = 0; t < c->nnz; t++) c->freq[idx[t]]++;
    if (side == OC_SIDE_T) p->mt = rows;
}

void oc_set_labels(oc_problem *p, int which, uint64_t rows, const uint64_t *rowptr,
                   const uint32_t *idx) {
    ycsr_t *y = which == OC_SIDE_T ? &p->YT : &p->YU;
    free_y(y);
    y->rows = rows; y->nnz = rowptr[rows];
    y->rowptr = (uint64_t *)xcalloc(rows + 1, sizeof(uint64_t));
    y->idx = (uint32_t *)xcalloc(y->nnz, sizeof(uint32_t));
    y->val = (double *)xcalloc(y->nnz, sizeof(double));
    memcpy(y->rowptr, rowptr, (rows + 1) * sizeof(uint64_t));
    memcpy(y->idx, idx, y->nnz * sizeof(uint32_t));
    if (which == OC_SIDE_T) { p->mt = rows; return; }

    /* U->n = max label + 1 (ffm.cpp:97); popular = label histogram / total (ffm.cpp:143,172-176) */
    uint64_t un = 0;
    for (uint64_t t = 0; t < y->nnz; t++) if ((uint64_t)idx[t] + 1 > un) un = (uint64_t)idx[t] + 1;
    p->un = un;
    free(p->popular);
    p->popular = (double *)xcalloc(un, sizeof(double));
    for (uint64_t t = 0; t < y->nnz; t++) p->popular[idx[t]] += 1;
    double tot = 0;
    for (uint64_t j = 0; j < un; j++) tot += p->popular[j];
    for (uint64_t j = 0; j < un; j++) p->popular[j] /= tot;

    /* transY (ffm.cpp:259-294): CSC sorted by (item, user); labels >= V->m are skipped.
     * A stable counting sort over rows visited in user order gives exactly that order. */
    ycsr_t *c = &p->YV;
    free_y(c);
    c->rows = p->n;
    c->rowptr = (uint64_t *)xcalloc(p->n + 1, sizeof(uint64_t));
    uint64_t kept = 0;
    for (uint64_t t = 0; t < y->nnz; t++)
        if (idx[t] < p->n) { c->rowptr[idx[t] + 1]++; kept++; }
    for (uint64_t j = 0; j < p->n; j++) c->rowptr[j + 1] += c->rowptr[j];
    c->nnz = kept;
    c->idx = (uint32_t *)xcalloc(kept, sizeof(uint32_t));
    c->val = (double *)xcalloc(kept, sizeof(double));
    uint64_t *cur = (uint64_t *)xcalloc(p->n + 1, sizeof(uint64_t));
    memcpy(cur, c->rowptr, (p->n + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < rows; i++)
        for (uint64_t t = rowptr[i]; t < rowptr[i + 1]; t++)
            if (idx[t] < p->n) c->idx[cur[idx[t]]++] = (uint32_t)i;
    free(cur);
}

void oc_set_test_nnx(oc_problem *p, const uint64_t *nnx) {
    free(p->nnx_t);
    p->nnx_t = (uint64_t *)xcalloc(p->mt, sizeof(uint64_t));
    memcpy(p->nnx_t, nnx, p->mt * sizeof(uint64_t));
    p->nnx_given = 1;
}

/* ---- RNG init ----------------------------------------------------------------------- */
/* qrsqrt, ffm.cpp:3-12: one Newton step on the 64-bit "fast inverse square root" guess */
static double approx_rsqrt(double x) {
    const double half = 0.5 * x;
    uint64_t bits;
    memcpy(&bits, &x, sizeof bits);
    bits = 0x5fe6eb50c7b537a9ULL - (bits >> 1);
    memcpy(&x, &bits, sizeof bits);
    return x * (1.5 - half * x * x);
}

/* init_mat, ffm.cpp:71-78.  default_random_engine is libstdc++'s minstd_rand0
 * (x <- 16807 x mod 2^31-1); uniform_real_distribution<double> draws generate_canonical<53>:
 * two engine outputs, (x1-1) + (x2-1)*R over R^2 with R = 2^31-2, clamped below 1. */
static void init_mat_rng(double *v, uint64_t rows, uint64_t cols) {
    uint64_t state = (uint64_t)rand() % 2147483647ULL;
    if (state == 0) state = 1;
    const double lo = -0.1 * approx_rsqrt((double)cols), hi = 0.1 * approx_rsqrt((double)cols);
    const double R = 2147483646.0;
    for (uint64_t t = 0; t < rows * cols; t++) {
        double sum = 0, scale = 1;
        for (int d = 0; d < 2; d++) {
            state = (state * 16807ULL) % 2147483647ULL;
            sum += (double)(state - 1) * scale;
            scale *= R;
        }
        double u = sum / scale;
        if (u >= 1.0) u = nextafter(1.0, 0.0);
        v[t] = u * (hi - lo) + lo;
    }
}

static void alloc_block(oc_problem *p, int f1, int f2) {
    const int b = blk(p, f1, f2);
    const uint64_t D1 = field_of(p, f1)->D, D2 = field_of(p, f2)->D;
    const uint64_t m1 = rows_of(p, f1), m2 = rows_of(p, f2);
    if (!p->W[b]) p->W[b] = (double *)xcalloc(D1 * p->k, sizeof(double));
    if (!p->H[b]) p->H[b] = (double *)xcalloc(D2 * p->k, sizeof(double));
    if (!p->P[b]) p->P[b] = (double *)xcalloc(m1 * p->k, sizeof(double));
    if (!p->Q[b]) p->Q[b] = (double *)xcalloc(m2 * p->k, sizeof(double));
}

void oc_init_model_rng(oc_problem *p) {
    /* block order f1 asc, f2 asc; W then H (ffm.cpp:495-506, 344-345) */
    for (int f1 = 0; f1 < p->f; f1++)
        for (int f2 = f1; f2 < p->f; f2++) {
            if (!block_exists(p, f1, f2)) continue;
            alloc_block(p, f1, f2);
            const int b = blk(p, f1, f2);
            init_mat_rng(p->W[b], field_of(p, f1)->D, p->k);
            init_mat_rng(p->H[b], field_of(p, f2)->D, p->k);
        }
}

uint64_t oc_block_rows(oc_problem *p, int f1, int f2, int which) {
    return which == 'W' ? field_of(p, f1)->D : field_of(p, f2)->D;
}

void oc_set_block(oc_problem *p, int f1, int f2, int which, const double *data) {
    alloc_block(p, f1, f2);
    const int b = blk(p, f1, f2);
    memcpy(which == 'W' ? p->W[b] : p->H[b], data,
           oc_block_rows(p, f1, f2, which) * p->k * sizeof(double));
}

void oc_get_block(oc_problem *p, int f1, int f2, int which, double *data) {
    const int b = blk(p, f1, f2);
    memcpy(data, which == 'W' ? p->W[b] : p->H[b],
           oc_block_rows(p, f1, f2, which) * p->k * sizeof(double));
}

/* ---- primitives ---------------------------------------------------------------------- */
static double dot(const double *x, const double *y, uint64_t n) {
    double s = 0;
    for (uint64_t i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}

/* UTX, ffm.cpp:314-331: C = X * A, X is CSR rows x D, A is D x k */
static void spmm(const csr_t *X, const double *A, double *C, int k) {
    memset(C, 0, X->rows * (uint64_t)k * sizeof(double));
    for (uint64_t i = 0; i < X->rows; i++) {
        double *c = C + i * k;
        for (uint64_t t = X->rowptr[i]; t < X->rowptr[i + 1]; t++) {
            const double *arow = A + (uint64_t)X->idx[t] * k;
            const double v = X->val[t];
            for (int d = 0; d < k; d++) c[d] += v * arow[d];
        }
    }
}

/* column sums weighted by x: out[d] = sum_i x[i] * A[i,d]  (mv(..., trans=true), ffm.cpp:47-51) */
static void col_weighted_sum(const double *A, const double *x, uint64_t rows, int k, double *out) {
    for (int d = 0; d < k; d++) out[d] = 0;
    for (uint64_t i = 0; i < rows; i++) {
        const double xi = x ? x[i] : 1.0;
        for (int d = 0; d < k; d++) out[d] += A[i * k + d] * xi;
    }
}

/* cache_sasb, ffm.cpp:514-535 */
static void cache_sasb(oc_problem *p) {
    const int k = p->k;
    memset(p->sa, 0, p->m * sizeof(double));
    memset(p->sb, 0, p->n * sizeof(double));
    double *tk = (double *)xcalloc(k, sizeof(double));
    for (int f1 = 0; f1 < p->fu; f1++)
        for (int f2 = p->fu; f2 < p->f; f2++) {
            const double *P1 = p->P[blk(p, f1, f2)], *Q1 = p->Q[blk(p, f1, f2)];
            col_weighted_sum(Q1, NULL, p->n, k, tk);
            for (uint64_t i = 0; i < p->m; i++) p->sa[i] += dot(P1 + i * k, tk, k);
            col_weighted_sum(P1, NULL, p->m, k, tk);
            for (uint64_t j = 0; j < p->n; j++) p->sb[j] += dot(Q1 + j * k, tk, k);
        }
    free(tk);
}

/* calc_side / add_side, ffm.cpp:352-373 */
static void calc_side(oc_problem *p) {
    const int k = p->k;
    for (int f1 = 0; f1 < p->fu; f1++)
        for (int f2 = f1; f2 < p->fu; f2++) {
            const int b = blk(p, f1, f2);
            for (uint64_t i = 0; i < p->m; i++) p->a[i] += dot(p->P[b] + i * k, p->Q[b] + i * k, k);
        }
    for (int f1 = p->fu; f1 < p->f; f1++)
        for (int f2 = f1; f2 < p->f; f2++) {
            const int b = blk(p, f1, f2);
            for (uint64_t j = 0; j < p->n; j++) p->b[j] += dot(p->P[b] + j * k, p->Q[b] + j * k, k);
        }
}

/* calc_cross, ffm.cpp:375-386 */
static double cross_value(const oc_problem *p, uint64_t i, uint64_t j) {
    double s = 0;
    for (int f1 = 0; f1 < p->fu; f1++)
        for (int f2 = p->fu; f2 < p->f; f2++) {
            const int b = blk(p, f1, f2);
            s += dot(p->P[b] + i * p->k, p->Q[b] + j * p->k, p->k);
        }
    return s;
}

/* init_y_tilde, ffm.cpp:388-403 */
static void init_y_tilde(oc_problem *p) {
    for (uint64_t i = 0; i < p->m; i++)
        for (uint64_t t = p->YU.rowptr[i]; t < p->YU.rowptr[i + 1]; t++) {
            const uint64_t j = p->YU.idx[t];
            p->YU.val[t] = p->a[i] + p->b[j] + cross_value(p, i, j) - 1;
        }
    for (uint64_t j = 0; j < p->n; j++)
        for (uint64_t t = p->YV.rowptr[j]; t < p->YV.rowptr[j + 1]; t++) {
            const uint64_t i = p->YV.idx[t];
            p->YV.val[t] = p->a[i] + p->b[j] + cross_value(p, i, j) - 1;
        }
}

void oc_init_state(oc_problem *p) {
    memset(p->a, 0, p->m * sizeof(double));
    memset(p->b, 0, p->n * sizeof(double));
    for (int f1 = 0; f1 < p->f; f1++)
        for (int f2 = f1; f2 < p->f; f2++) {
            if (!block_exists(p, f1, f2)) continue;
            alloc_block(p, f1, f2);
            const int b = blk(p, f1, f2);
            spmm(field_of(p, f1), p->W[b], p->P[b], p->k); /* init_pair, ffm.cpp:348-349 */
            spmm(field_of(p, f2), p->H[b], p->Q[b], p->k);
        }
    cache_sasb(p);                    /* ffm.cpp:508 */
    if (p->self_side) calc_side(p);   /* ffm.cpp:509-510 */
    init_y_tilde(p);                  /* ffm.cpp:511 */
}

/* ---- state readers --------------------------------------------------------------------- */
uint64_t oc_get_vec(oc_problem *p, const char *name, double *out) {
    const double *src = NULL;
    uint64_t n = 0;
    if (!strcmp(name, "a")) { src = p->a; n = p->m; }
    else if (!strcmp(name, "b")) { src = p->b; n = p->n; }
    else if (!strcmp(name, "sa")) { src = p->sa; n = p->m; }
    else if (!strcmp(name, "sb")) { src = p->sb; n = p->n; }
    else if (!strcmp(name, "ytilde_csr")) { src = p->YU.val; n = p->YU.nnz; }
    else if (!strcmp(name, "ytilde_csc")) { src = p->YV.val; n = p->YV.nnz; }
    else if (!strcmp(name, "popular")) { src = p->popular; n = p->un; }
    if (src && out) memcpy(out, src, n * sizeof(double));
    return n;
}

void oc_get_embed(oc_problem *p, int f1, int f2, int which, double *out) {
    const int b = blk(p, f1, f2);
    if (which == 'P') memcpy(out, p->P[b], rows_of(p, f1) * p->k * sizeof(double));
    else memcpy(out, p->Q[b], rows_of(p, f2) * p->k * sizeof(double));
}

void oc_get_csc(oc_problem *p, uint64_t *colptr, uint32_t *rowidx) {
    memcpy(colptr, p->YV.rowptr, (p->n + 1) * sizeof(uint64_t));
    memcpy(rowidx, p->YV.idx, p->YV.nnz * sizeof(uint32_t));
}

/* ---- one half block solve --------------------------------------------------------------- */
/* The reference calls every phase with (field being updated fa, the block's parameter matrix
 * W1, the companion embedding Q1, the embedding P1 that W1 generates); this gathers them. */
typedef struct {
    int fa, fb, b, side;
    double *W1, *Q1, *P1;
    const csr_t *X;
    const ycsr_t *Y;
    uint64_t m1, n1, D;
    const double *a1, *b1, *sa1;
} half_t;

static half_t half_of(oc_problem *p, int f1, int f2, int which) {
    half_t h;
    h.b = blk(p, f1, f2);
    h.side = is_side_block(p, f1, f2);
    if (which == 'W') { h.fa = f1; h.fb = f2; h.W1 = p->W[h.b]; h.Q1 = p->Q[h.b]; h.P1 = p->P[h.b]; }
    else { h.fa = f2; h.fb = f1; h.W1 = p->H[h.b]; h.Q1 = p->P[h.b]; h.P1 = p->Q[h.b]; }
    const int u = is_user(p, h.fa);
    h.X = field_of(p, h.fa);
    h.Y = u ? &p->YU : &p->YV;
    h.m1 = u ? p->m : p->n;
    h.n1 = u ? p->n : p->m;
    h.D = h.X->D;
    h.a1 = u ? p->a : p->b;
    h.b1 = u ? p->b : p->a;
    h.sa1 = u ? p->sa : p->sb;
    return h;
}

/* regulariser part shared by gradient and Hessian-vector product:
 * out += lambda * M, or lambda * freq[row] * M_row with --freq (ffm.cpp:561-570, 647-656, 786-794) */
static void add_reg(const oc_problem *p, const half_t *h, const double *M, double *out) {
    const int k = p->k;
    for (uint64_t row = 0; row < h->D; row++) {
        const double c = p->freq ? p->lambda * (double)h->X->freq[row] : p->lambda;
        for (int d = 0; d < k; d++) out[row * k + d] += c * M[row * k + d];
    }
}

/* gd_side, ffm.cpp:537-592 */
static void grad_side(oc_problem *p, const half_t *h, double *G) {
    const int k = p->k;
    const double w = p->w, r = p->r;
    double b_sum = 0;
    for (uint64_t j = 0; j < h->n1; j++) b_sum += h->b1[j];
    double *acc = (double *)xcalloc(h->D * k, sizeof(double));
    for (uint64_t i = 0; i < h->m1; i++) {
        const double *q = h->Q1 + i * k; /* the SAME row of the companion embedding */
        double z = w * ((double)h->n1 * (h->a1[i] - r) + b_sum + h->sa1[i]);
        for (uint64_t t = h->Y->rowptr[i]; t < h->Y->rowptr[i + 1]; t++)
            z += (1 - w) * h->Y->val[t] - w * (1 - r);
        for (uint64_t t = h->X->rowptr[i]; t < h->X->rowptr[i + 1]; t++) {
            double *g = acc + (uint64_t)h->X->idx[t] * k;
            for (int d = 0; d < k; d++) g[d] += q[d] * h->X->val[t] * z;
        }
    }
    for (uint64_t t = 0; t < h->D * k; t++) G[t] += acc[t];
    free(acc);
}

/* hs_side, ffm.cpp:594-628 */
static void hess_side(oc_problem *p, const half_t *h, const double *V, double *Hv) {
    const int k = p->k;
    const double w = p->w;
    double *acc = (double *)xcalloc(h->D * k, sizeof(double));
    for (uint64_t i = 0; i < h->m1; i++) {
        const double *q = h->Q1 + i * k;
        const double d1 = (1 - w) * (double)(uint32_t)(h->Y->rowptr[i + 1] - h->Y->rowptr[i]) + w * (double)h->n1;
        double z = 0;
        for (uint64_t t = h->X->rowptr[i]; t < h->X->rowptr[i + 1]; t++) {
            const double *v = V + (uint64_t)h->X->idx[t] * k;
            for (int d = 0; d < k; d++) z += q[d] * h->X->val[t] * v[d];
        }
        z *= d1;
        for (uint64_t t = h->X->rowptr[i]; t < h->X->rowptr[i + 1]; t++) {
            double *o = acc + (uint64_t)h->X->idx[t] * k;
            for (int d = 0; d < k; d++) o[d] += q[d] * h->X->val[t] * z;
        }
    }
    for (uint64_t t = 0; t < h->D * k; t++) Hv[t] += acc[t];
    free(acc);
}

/* C (k x k) = A^T B with A, B of `rows` x k  (mm(a,b,c,k,l), ffm.cpp:41-45) */
static void gram(const double *A, const double *B, uint64_t rows, int k, double *C) {
    memset(C, 0, (size_t)k * k * sizeof(double));
    for (uint64_t i = 0; i < rows; i++)
        for (int x = 0; x < k; x++) {
            const double ax = A[i * k + x];
            for (int y = 0; y < k; y++) C[x * k + y] += ax * B[i * k + y];
        }
}

/* gd_cross, ffm.cpp:630-703 */
static void grad_cross(oc_problem *p, const half_t *h, double *G) {
    const int k = p->k;
    const double w = p->w, r = p->r;
    const int u = is_user(p, h->fa);
    double *oQ = (double *)xcalloc(k, sizeof(double)), *bQ = (double *)xcalloc(k, sizeof(double));
    double *QTQ = (double *)xcalloc((size_t)k * k, sizeof(double));
    double *T = (double *)xcalloc(h->m1 * k, sizeof(double));
    col_weighted_sum(h->Q1, NULL, h->n1, k, oQ);   /* ffm.cpp:660 */
    col_weighted_sum(h->Q1, h->b1, h->n1, k, bQ);  /* ffm.cpp:661 */
    /* T = sum over ALL cross pairs (al, be) of Pa * (Qa^T Q1), "P" being this side (ffm.cpp:663-670) */
    for (int al = 0; al < p->fu; al++)
        for (int be = p->fu; be < p->f; be++) {
            const int ab = blk(p, al, be);
            const double *Pa = u ? p->P[ab] : p->Q[ab];
            const double *Qa = u ? p->Q[ab] : p->P[ab];
            gram(Qa, h->Q1, h->n1, k, QTQ);
            for (uint64_t i = 0; i < h->m1; i++)
                for (int x = 0; x < k; x++) {
                    const double pv = Pa[i * k + x];
                    for (int y = 0; y < k; y++) T[i * k + y] += pv * QTQ[x * k + y];
                }
        }
    double *acc = (double *)xcalloc(h->D * k, sizeof(double));
    double *pk = (double *)xcalloc(k, sizeof(double));
    for (uint64_t i = 0; i < h->m1; i++) {
        for (int d = 0; d < k; d++) pk[d] = 0;
        for (uint64_t t = h->Y->rowptr[i]; t < h->Y->rowptr[i + 1]; t++) {
            const double scale = (1 - w) * h->Y->val[t] - w * (1 - r);
            const double *q = h->Q1 + (uint64_t)h->Y->idx[t] * k;
            for (int d = 0; d < k; d++) pk[d] += scale * q[d];
        }
        const double z = h->a1[i] - r;
        const double *t1 = T + i * k;
        for (uint64_t t = h->X->rowptr[i]; t < h->X->rowptr[i + 1]; t++) {
            double *g = acc + (uint64_t)h->X->idx[t] * k;
            for (int d = 0; d < k; d++)
                g[d] += (pk[d] + w * (t1[d] + z * oQ[d] + bQ[d])) * h->X->val[t];
        }
    }
    for (uint64_t t = 0; t < h->D * k; t++) G[t] += acc[t];
    free(acc); free(pk); free(T); free(QTQ); free(oQ); free(bQ);
}

/* hs_cross + the V*QTQ product in front of it, ffm.cpp:706-742, 799 */
static void hess_cross(oc_problem *p, const half_t *h, const double *QTQ, const double *V, double *Hv) {
    const int k = p->k;
    const double w = p->w;
    double *VQ = (double *)xcalloc(h->D * k, sizeof(double));
    for (uint64_t row = 0; row < h->D; row++)
        for (int x = 0; x < k; x++) {
            const double v = V[row * k + x];
            for (int y = 0; y < k; y++) VQ[row * k + y] += v * QTQ[x * k + y];
        }
    double *acc = (double *)xcalloc(h->D * k, sizeof(double));
    double *phi = (double *)xcalloc(k, sizeof(double)), *tau = (double *)xcalloc(k, sizeof(double));
    double *ka = (double *)xcalloc(k, sizeof(double));
    for (uint64_t i = 0; i < h->m1; i++) {
        for (int d = 0; d < k; d++) phi[d] = tau[d] = ka[d] = 0;
        for (uint64_t t = h->X->rowptr[i]; t < h->X->rowptr[i + 1]; t++) {
            const uint64_t row = h->X->idx[t];
            for (int d = 0; d < k; d++) {
                phi[d] += h->X->val[t] * V[row * k + d];
                tau[d] += h->X->val[t] * VQ[row * k + d];
            }
        }
        for (uint64_t t = h->Y->rowptr[i]; t < h->Y->rowptr[i + 1]; t++) {
            const double *q = h->Q1 + (uint64_t)h->Y->idx[t] * k;
            const double s = dot(phi, q, k);
            for (int d = 0; d < k; d++) ka[d] += s * q[d];
        }
        for (uint64_t t = h->X->rowptr[i]; t < h->X->rowptr[i + 1]; t++) {
            double *o = acc + (uint64_t)h->X->idx[t] * k;
            for (int d = 0; d < k; d++) o[d] += ((1 - w) * ka[d] + w * tau[d]) * h->X->val[t];
        }
    }
    for (uint64_t t = 0; t < h->D * k; t++) Hv[t] += acc[t];
    free(acc); free(phi); free(tau); free(ka); free(VQ);
}

static void grad_half(oc_problem *p, const half_t *h, double *G) {
    memset(G, 0, h->D * p->k * sizeof(double));
    add_reg(p, h, h->W1, G);
    if (h->side) grad_side(p, h, G);
    else grad_cross(p, h, G);
}

static void hv_half(oc_problem *p, const half_t *h, const double *QTQ, const double *V, double *Hv) {
    memset(Hv, 0, h->D * p->k * sizeof(double));
    add_reg(p, h, V, Hv);                           /* ffm.cpp:786-794 */
    if (h->side) hess_side(p, h, V, Hv);            /* ffm.cpp:796-797 */
    else hess_cross(p, h, QTQ, V, Hv);              /* ffm.cpp:799-800 */
}

/* cg, ffm.cpp:744-813 */
static int cg_half(oc_problem *p, const half_t *h, const double *G, double *S) {
    const int k = p->k;
    const uint64_t len = h->D * k;
    double *V = (double *)xcalloc(len, sizeof(double)), *R = (double *)xcalloc(len, sizeof(double));
    double *Hv = (double *)xcalloc(len, sizeof(double));
    double *QTQ = NULL;
    if (!h->side) {
        QTQ = (double *)xcalloc((size_t)k * k, sizeof(double));
        gram(h->Q1, h->Q1, h->n1, k, QTQ);          /* ffm.cpp:767-771 */
    }
    double g2 = 0;
    for (uint64_t t = 0; t < len; t++) { R[t] = -G[t]; V[t] = R[t]; g2 += G[t] * G[t]; }
    double r2 = g2;
    int it = 0;
    const int max_cg = 20;
    const double eps = 9e-2;
    while (g2 * eps < r2 && it < max_cg) {          /* ffm.cpp:780 */
        it++;
        hv_half(p, h, QTQ, V, Hv);
        const double vHv = dot(V, Hv, len);
        const double gamma = r2, alpha = gamma / vHv;
        for (uint64_t t = 0; t < len; t++) S[t] += alpha * V[t];
        for (uint64_t t = 0; t < len; t++) R[t] -= alpha * Hv[t];
        r2 = dot(R, R, len);
        const double beta = r2 / gamma;
        for (uint64_t t = 0; t < len; t++) V[t] = R[t] + beta * V[t];
    }
    p->cg_total += (uint64_t)it;
    free(V); free(R); free(Hv); free(QTQ);
    return it;
}

/* update_side (ffm.cpp:405-437) and update_cross (ffm.cpp:439-465) */
static void update_half(oc_problem *p, const half_t *h, const double *S) {
    const int k = p->k;
    const int u = is_user(p, h->fa);
    for (uint64_t t = 0; t < h->D * k; t++) h->W1[t] += S[t];
    double *XS = (double *)xcalloc(h->m1 * k, sizeof(double));
    spmm(h->X, S, XS, k);
    for (uint64_t t = 0; t < h->m1 * k; t++) h->P1[t] += XS[t];
    ycsr_t *Yown = u ? &p->YU : &p->YV;     /* rows indexed like this side */
    ycsr_t *Yoth = u ? &p->YV : &p->YU;     /* rows indexed by the other side */
    if (h->side) {
        double *a1 = u ? p->a : p->b;
        for (uint64_t i = 0; i < h->m1; i++) {
            const double gap = dot(XS + i * k, h->Q1 + i * k, k);
            a1[i] += gap;
            for (uint64_t t = Yown->rowptr[i]; t < Yown->rowptr[i + 1]; t++) Yown->val[t] += gap;
            XS[i * k] = gap; /* reuse column 0 as the gap vector for the second copy */
        }
        for (uint64_t j = 0; j < Yoth->rows; j++)
            for (uint64_t t = Yoth->rowptr[j]; t < Yoth->rowptr[j + 1]; t++)
                Yoth->val[t] += XS[(uint64_t)Yoth->idx[t] * k];
    } else {
        for (uint64_t i = 0; i < h->m1; i++)
            for (uint64_t t = Yown->rowptr[i]; t < Yown->rowptr[i + 1]; t++)
                Yown->val[t] += dot(XS + i * k, h->Q1 + (uint64_t)Yown->idx[t] * k, k);
        for (uint64_t j = 0; j < Yoth->rows; j++)
            for (uint64_t t = Yoth->rowptr[j]; t < Yoth->rowptr[j + 1]; t++)
                Yoth->val[t] += dot(XS + (uint64_t)Yoth->idx[t] * k, h->Q1 + j * k, k);
    }
    free(XS);
}

void oc_grad(oc_problem *p, int f1, int f2, int which, double *G) {
    half_t h = half_of(p, f1, f2, which);
    grad_half(p, &h, G);
}

void oc_hess_vec(oc_problem *p, int f1, int f2, int which, const double *V, double *Hv) {
    half_t h = half_of(p, f1, f2, which);
    double *QTQ = NULL;
    if (!h.side) {
        QTQ = (double *)xcalloc((size_t)p->k * p->k, sizeof(double));
        gram(h.Q1, h.Q1, h.n1, p->k, QTQ);
    }
    hv_half(p, &h, QTQ, V, Hv);
    free(QTQ);
}

int oc_cg(oc_problem *p, int f1, int f2, int which, const double *G, double *S) {
    half_t h = half_of(p, f1, f2, which);
    memset(S, 0, h.D * p->k * sizeof(double));
    const int it = cg_half(p, &h, G, S);
    p->cg_total -= (uint64_t)it; /* observation only */
    return it;
}

static void solve_half(oc_problem *p, int f1, int f2, int which) {
    half_t h = half_of(p, f1, f2, which);
    const uint64_t len = h.D * p->k;
    double *G = (double *)xcalloc(len, sizeof(double)), *S = (double *)xcalloc(len, sizeof(double));
    grad_half(p, &h, G);
    cg_half(p, &h, G, S);
    update_half(p, &h, S);
    free(G); free(S);
}

void oc_solve_block(oc_problem *p, int f1, int f2) {
    solve_half(p, f1, f2, 'W');   /* W first ... */
    solve_half(p, f1, f2, 'H');   /* ... then H against the already-updated P1 (ffm.cpp:826-832, 843-849) */
}

void oc_one_epoch(oc_problem *p) {
    if (p->self_side) {
        for (int f1 = 0; f1 < p->fu; f1++)
            for (int f2 = f1; f2 < p->fu; f2++) oc_solve_block(p, f1, f2);
        for (int f1 = p->fu; f1 < p->f; f1++)
            for (int f2 = f1; f2 < p->f; f2++) oc_solve_block(p, f1, f2);
    }
    for (int f1 = 0; f1 < p->fu; f1++)
        for (int f2 = p->fu; f2 < p->f; f2++) oc_solve_block(p, f1, f2);
    if (p->self_side) cache_sasb(p);
}

uint64_t oc_cg_iters_total(oc_problem *p) { return p->cg_total; }

/* func / pq / norm_block, ffm.cpp:1303-1351.  The reference sums every block of the upper
 * triangle and therefore crashes under --ns; here absent blocks contribute nothing. */
double oc_func(oc_problem *p) {
    const int k = p->k;
    double res = 0;
    for (uint64_t i = 0; i < p->m; i++) {
        for (uint64_t j = 0; j < p->n; j++) {
            double yhat = 0;
            for (int f1 = 0; f1 < p->f; f1++)
                for (int f2 = f1; f2 < p->f; f2++) {
                    if (!block_exists(p, f1, f2)) continue;
                    const int b = blk(p, f1, f2);
                    const uint64_t pi = is_user(p, f1) ? i : j, qj = is_user(p, f2) ? i : j;
                    yhat += dot(p->Q[b] + qj * k, p->P[b] + pi * k, k);
                }
            int pos = 0;
            for (uint64_t t = p->YU.rowptr[i]; t < p->YU.rowptr[i + 1]; t++)
                if (p->YU.idx[t] == j) { pos = 1; break; }
            res += pos ? (1 - yhat) * (1 - yhat) : p->w * (p->r - yhat) * (p->r - yhat);
        }
    }
    for (int f1 = 0; f1 < p->f; f1++)
        for (int f2 = f1; f2 < p->f; f2++) {
            if (!block_exists(p, f1, f2)) continue;
            const int b = blk(p, f1, f2);
            const uint64_t lw = field_of(p, f1)->D * k, lh = field_of(p, f2)->D * k;
            res += p->lambda * (dot(p->W[b], p->W[b], lw) + dot(p->H[b], p->H[b], lh));
        }
    return 0.5 * res;
}

/* ---- evaluation -------------------------------------------------------------------------- */
#define OC_MIN_Z (-1000.0) /* ffm.h:40 */
static const uint32_t OC_TOPK[5] = {5, 10, 20, 40, 80}; /* init_va, ffm.cpp:899-909 */

static int is_label(const uint32_t *labels, uint32_t n_labels, uint32_t item) {
    for (uint32_t t = 0; t < n_labels; t++) if (labels[t] == item) return 1;
    return 0;
}

/* ndcg(), ffm.cpp:1059-1128, for one cut-off K: gain 1/log2(rank+2), IDCG over the first
 * min(|labels|, K) ranks */
double oc_ndcg_at(const uint32_t *ranking, uint32_t n_ranked, const uint32_t *labels,
                  uint32_t n_labels, uint32_t K) {
    double dcg = 0, idcg = 0;
    for (uint32_t rank = 0; rank < K && rank < n_ranked; rank++) {
        if (is_label(labels, n_labels, ranking[rank])) dcg += 1.0 / log2((double)rank + 2);
        if (n_labels > rank) idcg += 1.0 / log2((double)rank + 2);
    }
    return dcg / idcg;
}

/* validate(), ffm.cpp:925-1016 with pred_z (:915-923), prec_k (:1018-1057), ndcg (:1059-1128) */
void oc_validate(oc_problem *p, double *prec, double *ndcg, double *ploss_out, uint32_t *topk,
                 double *Zout) {
    const int k = p->k;
    const uint64_t mt = p->mt, n = p->n, un = p->un;
    double **Pva = (double **)xcalloc(p->nr_blocks, sizeof(double *));
    double **Qva = (double **)xcalloc(p->nr_blocks, sizeof(double *));
    for (int f1 = 0; f1 < p->f; f1++)
        for (int f2 = f1; f2 < p->f; f2++) {
            if (!block_exists(p, f1, f2)) continue;
            const int b = blk(p, f1, f2);
            const csr_t *X1 = is_user(p, f1) ? &p->XT[f1] : &p->XV[f1 - p->fu];
            const csr_t *X2 = is_user(p, f2) ? &p->XT[f2] : &p->XV[f2 - p->fu];
            Pva[b] = (double *)xcalloc(X1->rows * k, sizeof(double));
            Qva[b] = (double *)xcalloc(X2->rows * k, sizeof(double));
            spmm(X1, p->W[b], Pva[b], k);
            spmm(X2, p->H[b], Qva[b], k);
        }
    double *at = (double *)xcalloc(mt, sizeof(double)), *bt = (double *)xcalloc(n, sizeof(double));
    if (p->self_side) {
        for (int f1 = 0; f1 < p->fu; f1++)
            for (int f2 = f1; f2 < p->fu; f2++) {
                const int b = blk(p, f1, f2);
                for (uint64_t i = 0; i < mt; i++) at[i] += dot(Pva[b] + i * k, Qva[b] + i * k, k);
            }
        for (int f1 = p->fu; f1 < p->f; f1++)
            for (int f2 = f1; f2 < p->f; f2++) {
                const int b = blk(p, f1, f2);
                for (uint64_t j = 0; j < n; j++) bt[j] += dot(Pva[b] + j * k, Qva[b] + j * k, k);
            }
    }
    double hits_tot[5] = {0}, ndcg_tot[5] = {0}, ploss = 0;
    const uint64_t zcap = n > un ? n : un;
    double *z = (double *)xcalloc(zcap, sizeof(double));
    for (uint64_t i = 0; i < mt; i++) {
        uint64_t nnx = 0;
        if (p->nnx_given) nnx = p->nnx_t[i];
        else for (int fi = 0; fi < p->fu; fi++) nnx += p->XT[fi].rowptr[i + 1] - p->XT[fi].rowptr[i];
        uint64_t zlen;
        if (nnx == 0) {                                    /* cold row: rank by popularity */
            zlen = un;
            memcpy(z, p->popular, un * sizeof(double));
        } else {
            zlen = n;
            memcpy(z, bt, n * sizeof(double));
            for (int f1 = 0; f1 < p->fu; f1++)
                for (int f2 = p->fu; f2 < p->f; f2++) {
                    const int b = blk(p, f1, f2);
                    const double *pi = Pva[b] + i * k;
                    for (uint64_t j = 0; j < n; j++) z[j] += dot(Qva[b] + j * k, pi, k);
                }
        }
        if (Zout) {
            for (uint64_t j = 0; j < n; j++) Zout[i * n + j] = j < zlen ? z[j] : 0.0;
        }
        const uint32_t *labels = p->YT.idx + p->YT.rowptr[i];
        const uint32_t n_labels = (uint32_t)(p->YT.rowptr[i + 1] - p->YT.rowptr[i]);
        for (uint32_t t = 0; t < n_labels; t++)
            if (labels[t] < zlen) {
                const double e = 1 - z[labels[t]] - at[i];
                ploss += e * e;
            }
        /* top-80 by repeated first-argmax over [0, U->n), winners overwritten with MIN_Z */
        double hit[5] = {0}, dcg[5] = {0}, idcg[5] = {0};
        uint32_t count = 0;
        for (int s = 0; s < 5; s++) {
            while (count < OC_TOPK[s]) {
                if (count >= un) break;
                uint64_t arg = 0;
                for (uint64_t j = 1; j < un; j++) if (z[j] > z[arg]) arg = j;
                z[arg] = OC_MIN_Z;
                if (topk) topk[i * 80 + count] = (uint32_t)arg;
                if (is_label(labels, n_labels, (uint32_t)arg)) {
                    hit[s] += 1;
                    dcg[s] += 1.0 / log2((double)count + 2);
                }
                if (n_labels > count) idcg[s] += 1.0 / log2((double)count + 2);
                count++;
            }
        }
        if (topk) for (uint32_t c = count; c < 80; c++) topk[i * 80 + c] = UINT32_MAX;
        for (int s = 1; s < 5; s++) { hit[s] += hit[s - 1]; dcg[s] += dcg[s - 1]; idcg[s] += idcg[s - 1]; }
        for (int s = 0; s < 5; s++) { hits_tot[s] += hit[s]; ndcg_tot[s] += dcg[s] / idcg[s]; }
    }
    for (int s = 0; s < 5; s++) {
        prec[s] = hits_tot[s] / ((double)mt * OC_TOPK[s]);
        ndcg[s] = ndcg_tot[s] / (double)mt;
    }
    *ploss_out = sqrt(ploss / (double)mt);
    for (int b = 0; b < p->nr_blocks; b++) { free(Pva[b]); free(Qva[b]); }
    free(Pva); free(Qva); free(at); free(bt); free(z);
}
