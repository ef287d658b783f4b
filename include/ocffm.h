/* include/ocffm.h -- C ABI of libocffm_cuda.so, the B200 (sm_100a) implementation of the
 * one-class FFM solver + full-ranking evaluator of johncreed/one-class-ffm.
 *
 * The reference has no FFI: its boundary is the C++ class API of ffm.h consumed by
 * train.cpp:177-199.  Each entry point below replaces one piece of that API (cited as
 * ffm.h / ffm.cpp file:line); the host-side C++ mirror (one-class-ffm_b200/host/ffm.h) keeps
 * the reference's class names and method signatures and is a thin caller of this header.
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every array is caller-owned HOST memory unless the name
 *     says otherwise; the context owns all device memory.
 *   - every function returns 0 on success, a negative OCFFM_E_* code on failure, never throws;
 *     ocffm_last_error() gives the message (thread local).
 *   - there is NO CPU fallback: without a CUDA device ocffm_create fails with OCFFM_E_NODEVICE.
 *   - model matrices cross the boundary as fp64 row-major [rows x k] exactly like the
 *     reference's Vec (ffm.h:34-38); on the device they are stored in the context's compute
 *     type (OCFFM_F32 or OCFFM_F64) with the latent dimension padded to kp = 2^ceil(log2 k) >= 4.
 *   - calls on one context are not thread-safe (the reference's ImpProblem is single-owner).
 */
#ifndef OCFFM_H
#define OCFFM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCFFM_ABI_VERSION 2

#if defined(__GNUC__)
#define OCFFM_API __attribute__((visibility("default")))
#else
#define OCFFM_API
#endif

enum {
    OCFFM_OK = 0,
    OCFFM_E_INVALID = -1,   /* bad argument / call order */
    OCFFM_E_NODEVICE = -2,  /* no CUDA device (there is no CPU fallback) */
    OCFFM_E_CUDA = -3,      /* CUDA runtime error */
    OCFFM_E_NOMEM = -4,
    OCFFM_E_COMM = -5,      /* NCCL error */
    OCFFM_E_UNSUPPORTED = -6
};

enum { OCFFM_F32 = 0, OCFFM_F64 = 1 };               /* device compute/storage type */
enum { OCFFM_SIDE_U = 0, OCFFM_SIDE_V = 1, OCFFM_SIDE_T = 2 }; /* train users, items, test users */

/* Parameter (ffm.h:42-49) + device selection.  nr_threads/model_path/predict_path stay on the host. */
typedef struct ocffm_params {
    double lambda;     /* -l, Parameter::lambda */
    double omega;      /* -w, Parameter::omega  */
    double r;          /* -r, Parameter::r      */
    uint32_t k;        /* -k, Parameter::k      */
    int32_t self_side; /* 0 with --ns           */
    int32_t freq;      /* 1 with --freq         */
    int32_t dtype;     /* OCFFM_F32 | OCFFM_F64 */
    int32_t device;    /* CUDA ordinal; -1 = current device */
} ocffm_params;

typedef struct ocffm_ctx ocffm_ctx;

OCFFM_API int ocffm_abi_version(void);
OCFFM_API const char *ocffm_last_error(void);
/* number of CUDA devices visible (0 on a CPU-only host; never fails) */
OCFFM_API int ocffm_device_count(void);

/* ImpProblem::ImpProblem (ffm.h:84-86): fu/fv = U->f / V->f, m = U->m, n = V->m */
OCFFM_API int ocffm_create(ocffm_ctx **out, const ocffm_params *prm, uint32_t fu, uint32_t fv, uint64_t m,
                 uint64_t n);
OCFFM_API int ocffm_destroy(ocffm_ctx *ctx);

/* ---- multi-GPU (one process per GPU; rows of U, V, Omega and test rows are sharded) ----
 * SURVEY.md 8(e).  Rank 0 makes an id, the caller ships the 128 bytes to the other ranks by
 * any means (bench.py uses torch.distributed), every rank calls ocffm_comm_init BEFORE any
 * data is set.  With nranks == 1 (or never called) no NCCL symbol is touched. */
OCFFM_API int ocffm_comm_unique_id(void *id128);
/* the contiguous row slice [lo, hi) rank `rank` of `nranks` owns out of `rows` rows (pure host
 * arithmetic, usable without a GPU; the same rule shards users, items and test rows) */
OCFFM_API int ocffm_shard_range(uint64_t rows, int nranks, int rank, uint64_t *lo, uint64_t *hi);
OCFFM_API int ocffm_comm_init(ocffm_ctx *ctx, int nranks, int rank, const void *id128);

/* ---- data: ImpData after read()/split_fields()/transY() (ffm.h:59-79, ffm.cpp:80-294) ----
 * One field's design matrix ImpData::Xs[field] as CSR over ALL rows of that side
 * (rows = m, n or m_t); D = Ds[field] of the TRAINING data also for OCFFM_SIDE_T
 * (features >= D are already dropped by the reader, ffm.cpp:104-105). */
OCFFM_API int ocffm_set_field(ocffm_ctx *ctx, int side, uint32_t field, uint64_t rows, uint64_t D,
                    const uint64_t *rowptr, const uint32_t *idx, const double *val);
/* U->Y as CSR (ffm.cpp:160-170) and, optionally, V->Y = transY(U->Y) as CSC sorted by
 * (item, user) (ffm.cpp:259-294); pass NULLs to have the library derive the CSC.
 * n_ranked = U->n = max label + 1 (ffm.cpp:97); popular = U->popular (ffm.cpp:143,172-176),
 * n_ranked doubles, may be NULL (derived). */
OCFFM_API int ocffm_set_labels(ocffm_ctx *ctx, uint64_t m, const uint64_t *rowptr, const uint32_t *idx,
                     const uint64_t *csc_colptr, const uint32_t *csc_rowidx, uint64_t n_ranked,
                     const double *popular);
/* Uva->Y as CSR and Uva->nnx (kept features per test row; 0 -> ranked by `popular`,
 * ffm.cpp:975-977).  nnx may be NULL (derived from the OCFFM_SIDE_T fields). */
OCFFM_API int ocffm_set_test_labels(ocffm_ctx *ctx, uint64_t m_t, const uint64_t *rowptr,
                          const uint32_t *idx, const uint64_t *nnx);

/* ---- model: ImpProblem::W / H (ffm.h:107), block (f1 <= f2) in GLOBAL field ids --------
 * which = 'W' (rows = Ds[f1]) or 'H' (rows = Ds[f2]); data is fp64 [rows x k]. */
OCFFM_API int ocffm_set_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, const double *data,
                    uint64_t rows);
OCFFM_API int ocffm_get_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *data,
                    uint64_t rows);

/* ImpProblem::init() after init_mat (ffm.cpp:346-349, 508-511): P = X W, Q = X H for every
 * block, cache_sasb, calc_side, init_y_tilde. */
OCFFM_API int ocffm_init_state(ocffm_ctx *ctx);

/* Host mirrors.  The reference keeps W / H in host memory (ffm.h:107) and one_epoch() updates them in
 * place; a caller that needs that after EVERY outer iteration registers a PINNED fp64 [rows x k]
 * buffer per block half.  While a mirror is registered, ocffm_one_epoch converts and copies the block
 * into it on a second stream right after the block's solve of the iteration -- overlapped with the
 * remaining block solves -- and returns once every mirror is current.  pinned == NULL unregisters.
 * (ocffm_get_block remains the one-off, serial download.) */
OCFFM_API int ocffm_mirror_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *pinned,
                                 uint64_t rows);

/* Device-side model init: every stored W / H block drawn on the GPU from
 * U(-s, s), s = 0.1 * qrsqrt(k) as init_mat does (ffm.cpp:3-12, 71-78), by a counter-based generator:
 * element (row, col) of a block is a pure function of (seed, block, W|H, row, col), so the model is
 * identical on every rank and for every launch shape, and no model bytes cross PCIe.  It is NOT the
 * libstdc++ minstd stream of the reference -- use ocffm_set_block with host-drawn values (what the
 * C++ host layer does by default) when the reference's exact initial model is wanted. */
OCFFM_API int ocffm_init_model(ocffm_ctx *ctx, uint64_t seed);

/* (lambda, omega, r) sweeps over one uploaded data set (script/grid.sh:186-240 runs a 12 x 3 grid of
 * solves on the same files): change the hyper-parameters of a live context; data, CSC, work lists
 * and hot-feature tables stay resident.  Set the model blocks and call ocffm_init_state again. */
OCFFM_API int ocffm_set_hyper(ocffm_ctx *ctx, double lambda, double omega, double r);

/* ---- solver (ffm.cpp:815-870) ---- */
OCFFM_API int ocffm_solve_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2); /* solve_side / solve_cross */
OCFFM_API int ocffm_one_epoch(ocffm_ctx *ctx);                             /* one_epoch */
/* observation of one half block solve WITHOUT applying it (parity tests):
 * gd_side/gd_cross (ffm.cpp:537-592, 630-703); Hv = lambda V + hs_* (ffm.cpp:783-801);
 * cg (ffm.cpp:744-813) -> S and the iteration count. */
OCFFM_API int ocffm_grad(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *G, uint64_t rows);
OCFFM_API int ocffm_hess_vec(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, const double *V,
                   double *Hv, uint64_t rows);
OCFFM_API int ocffm_cg(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, const double *G, double *S,
             uint64_t rows, int32_t *iters);

/* ImpProblem::func() (ffm.cpp:1321-1351) evaluated exactly through Gram identities from the
 * device state (O((m+n) K^2) instead of O(m n K)); under --ns only cross blocks are summed. */
OCFFM_API int ocffm_objective(ocffm_ctx *ctx, double *value);

/* ImpProblem::validate() (ffm.cpp:925-1016): prec[5], ndcg[5] for K = 5,10,20,40,80, ploss.
 * topk (optional) receives m_t x 80 item ids in rank order, UINT32_MAX past the ranked count. */
OCFFM_API int ocffm_validate(ocffm_ctx *ctx, double *prec, double *ndcg, double *ploss, uint32_t *topk);

/* ---- state readers for parity tests (fp64, unpadded) ----
 * name: "a" "b" "sa" "sb" "ytilde_csr" "ytilde_csc" "popular"; returns the element count in
 * *count when out == NULL. */
OCFFM_API int ocffm_get_vec(ocffm_ctx *ctx, const char *name, double *out, uint64_t *count);
/* P[f12] (which='P', rows of field f1's side) or Q[f12] (which='Q'), [rows x k] */
OCFFM_API int ocffm_get_embed(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *out,
                    uint64_t rows);
/* CSC of Omega as held on the device (bit-exact index parity) */
OCFFM_API int ocffm_get_csc(ocffm_ctx *ctx, uint64_t *colptr, uint32_t *rowidx);

/* ---- instrumentation ---- */
typedef struct ocffm_stats {
    uint64_t kernel_launches;   /* kernels this library launched on this context since reset */
    uint64_t cg_iters;          /* CG iterations since reset */
    uint64_t nnz_traversed;     /* SURVEY.md 8(d) N_trav since reset */
    uint64_t algo_bytes;        /* SURVEY.md 8(d) algorithmic bytes since reset */
    /* phase times (ms) since reset, only with OCFFM_PROFILE=2: gradient / CG / update of the
     * same-side half solves, then of the cross half solves */
    double ms_side_grad, ms_side_cg, ms_side_update, ms_cross_grad, ms_cross_cg, ms_cross_update;
    uint64_t hv_launches;       /* launches of the dominant kernel (hv_cross rows) */
    uint64_t hv_algo_bytes;     /* its algorithmic bytes since reset */
    double hv_ms;               /* its device time since reset (CUDA events, OCFFM_PROFILE=1) */
    /* ABI 2.  All counters above are THIS RANK's share (sum over ranks = whole job). */
    uint64_t omega_device_bytes;   /* HBM held by this rank's slices of Omega (both orientations: row pointers,
                                      column ids, y-tilde, work-item lists) */
    uint64_t row_gram_bytes;       /* per-row observed Gram buffer (0 when the path is off) */
    uint64_t row_gram_builds;      /* half solves that built it since reset */
    /* cross halves solved by the persistent CG kernel (one cooperative launch per half solve: direction +
     * V QTQ, hs_cross row pass and step of every iteration), timed with CUDA events under OCFFM_PROFILE=1 */
    double cg_kernel_ms;
    uint64_t cg_kernel_algo_bytes; /* algorithmic bytes of those solves (Hessian passes + CG vector passes) */
    uint64_t cg_kernel_launches, cg_kernel_iters;
} ocffm_stats;
OCFFM_API int ocffm_get_stats(ocffm_ctx *ctx, ocffm_stats *out);
OCFFM_API int ocffm_reset_stats(ocffm_ctx *ctx);
OCFFM_API int ocffm_synchronize(ocffm_ctx *ctx);
/* the context's CUDA stream (cudaStream_t as void*) so a caller can bracket calls with events */
OCFFM_API int ocffm_stream(ocffm_ctx *ctx, void **stream);

#ifdef __cplusplus
}
#endif
#endif
