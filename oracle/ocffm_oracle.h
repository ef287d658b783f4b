/* oracle/ocffm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, fp64, single-threaded restatement of the reference's one-class FFM solver and
 * full-ranking evaluator (the hot path of johncreed/one-class-ffm: ffm.cpp one_epoch() and
 * validate()).  It exists only to CHECK the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (libocffm_cuda.so) never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function below
 * against dumps of the unmodified reference (oracle/ref_harness.cpp -> tests/golden/) and the
 * reference's own nDCG known-answer fixture (script/nDCG_degub_tool).
 */
#ifndef OCFFM_ORACLE_H
#define OCFFM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oc_problem oc_problem;

enum { OC_SIDE_U = 0, OC_SIDE_V = 1, OC_SIDE_T = 2 };

oc_problem *oc_create(int fu, int fv, uint64_t m, uint64_t n, int k, double lambda, double omega,
                      double r, int self_side, int freq);
void oc_destroy(oc_problem *p);

/* ImpData::Xs[fi] after split_fields (ffm.cpp:185-257); freq is derived here (ffm.cpp:235-241) */
void oc_set_field(oc_problem *p, int side, int fi, uint64_t rows, uint64_t D,
                  const uint64_t *rowptr, const uint32_t *idx, const double *val);
/* U->Y (which=OC_SIDE_U) or Uva->Y (which=OC_SIDE_T) as CSR.  For U this also derives U->n,
 * `popular` (ffm.cpp:143,172-176) and the CSC copy of transY (ffm.cpp:259-294). */
void oc_set_labels(oc_problem *p, int which, uint64_t rows, const uint64_t *rowptr,
                   const uint32_t *idx);
/* nnx of the test rows (features kept by the reader, ffm.cpp:156,178-181); default: derived
 * from the fields given through oc_set_field(OC_SIDE_T, ...) */
void oc_set_test_nnx(oc_problem *p, const uint64_t *nnx);

/* init_mat for every block in reference order (ffm.cpp:71-78, 495-506) using libc rand() */
void oc_init_model_rng(oc_problem *p);
uint64_t oc_block_rows(oc_problem *p, int f1, int f2, int which /* 'W' or 'H' */);
void oc_set_block(oc_problem *p, int f1, int f2, int which, const double *data);
void oc_get_block(oc_problem *p, int f1, int f2, int which, double *data);

/* everything of ImpProblem::init() after the RNG (ffm.cpp:346-349, 508-511) */
void oc_init_state(oc_problem *p);

/* state readers; name in {"a","b","sa","sb","ytilde_csr","ytilde_csc","popular"} or
 * "P"/"Q" with the block given */
uint64_t oc_get_vec(oc_problem *p, const char *name, double *out);
void oc_get_embed(oc_problem *p, int f1, int f2, int which /* 'P' or 'Q' */, double *out);
void oc_get_csc(oc_problem *p, uint64_t *colptr, uint32_t *rowidx);

/* one half of a block solve, observed only: which = 'W' (update W[f12]) or 'H' */
void oc_grad(oc_problem *p, int f1, int f2, int which, double *G);               /* gd_side / gd_cross */
void oc_hess_vec(oc_problem *p, int f1, int f2, int which, const double *V, double *Hv); /* cg():783-801 */
int oc_cg(oc_problem *p, int f1, int f2, int which, const double *G, double *S); /* cg(); returns #iters */

void oc_solve_block(oc_problem *p, int f1, int f2);   /* solve_side / solve_cross (ffm.cpp:815-850) */
void oc_one_epoch(oc_problem *p);                     /* ffm.cpp:852-870 */
uint64_t oc_cg_iters_total(oc_problem *p);

/* brute-force objective, func() (ffm.cpp:1321-1351); with --ns only cross pairs are summed */
double oc_func(oc_problem *p);

/* validate() (ffm.cpp:925-1016): prec[5], ndcg[5] for K=5,10,20,40,80, ploss; optional
 * topk[m_t*80] item ids in rank order (UINT32_MAX where fewer than 80 items are ranked) and
 * optional Z[m_t*n] raw scores. */
void oc_validate(oc_problem *p, double *prec, double *ndcg, double *ploss, uint32_t *topk,
                 double *Z);

/* pure helper pinned by the reference's known-answer fixture: per-user nDCG@K of a given
 * ranking (ffm.cpp:1059-1128) */
double oc_ndcg_at(const uint32_t *ranking, uint32_t n_ranked, const uint32_t *labels,
                  uint32_t n_labels, uint32_t K);

#ifdef __cplusplus
}
#endif
#endif
