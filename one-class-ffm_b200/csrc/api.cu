// api.cu -- the C ABI (include/ocffm.h) and the host-side schedule of the solver and evaluator.
//
// The schedule (which block, which half, W before H, CG stop test) is the reference's
// (ffm.cpp:744-870); everything numeric runs in the kernels of rows.cu / dense.cu / eval.cu on
// one CUDA stream.  There is no CPU code path for any of it: without a device, ocffm_create fails.
//
// Device layout (T = float or double, kp = padded k):
//   W[f12] [D_f1 x kp], H[f12] [D_f2 x kp]              parameter blocks, replicated on every rank
//   Pc [m x Kc], Qc [n x Kc], Kc = fu*fv*kp             cross-pair embeddings concatenated so that
//                                                       pair p = a*fv + (b-fu) is columns [p*kp,(p+1)*kp)
//   P[f12], Q[f12] [rows x kp]                          same-side embeddings (only with self_side)
//   a [m], b [n], sa [m], sb [n]
//   Omega twice: by user (CSR, U->Y) and by item (CSC, V->Y), each with its own y-tilde copy
//   (ffm.cpp:393,400) and a work-item list (kernels.h OmegaView).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/ocffm.h"
#include "common.cuh"
#include "kernels.h"

namespace ocffm {

thread_local uint64_t *g_launch_counter = nullptr;
static thread_local std::string g_last_error;

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time so that the library loads on hosts without NCCL and a single-GPU
// context never touches it.
// ---------------------------------------------------------------------------------------------
struct Nccl {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;

    static Nccl &get() {
        static Nccl n;
        if (!n.h) {
            const char *names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char *nm : names) {
                n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (n.h) break;
            }
            if (!n.h) throw Error(OCFFM_E_COMM, std::string("cannot load libnccl: ") + dlerror());
#define OC_SYM(field, name)                                                      \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, name));             \
    if (!n.field) throw Error(OCFFM_E_COMM, std::string("missing NCCL symbol ") + name)
            OC_SYM(GetUniqueId, "ncclGetUniqueId");
            OC_SYM(CommInitRank, "ncclCommInitRank");
            OC_SYM(CommDestroy, "ncclCommDestroy");
            OC_SYM(AllReduce, "ncclAllReduce");
            OC_SYM(Broadcast, "ncclBroadcast");
            OC_SYM(GroupStart, "ncclGroupStart");
            OC_SYM(GroupEnd, "ncclGroupEnd");
            OC_SYM(GetErrorString, "ncclGetErrorString");
#undef OC_SYM
        }
        return n;
    }
};

#define OC_NCCL(expr)                                                                       \
    do {                                                                                    \
        ncclResult_t r__ = (expr);                                                          \
        if (r__ != ncclSuccess)                                                             \
            throw ::ocffm::Error(OCFFM_E_COMM, std::string(#expr) + ": " + ::ocffm::Nccl::get().GetErrorString(r__)); \
    } while (0)

template <typename T>
ncclDataType_t nccl_type();
template <>
ncclDataType_t nccl_type<float>() { return ncclFloat; }
template <>
ncclDataType_t nccl_type<double>() { return ncclDouble; }

struct Comm {
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    // peer-memory path (peer.cu) for messages up to pv.cap bytes; NCCL carries the rest
    bool peer_ok = false;
    PeerView pv{};
    unsigned char *area_owned = nullptr;   // this rank's staging area (pv.base[rank] when rank < kPeerMaxRanks)
    int *peer_err_host = nullptr;
    std::vector<void *> opened;
    unsigned long long seq = 0;
    uint64_t peer_calls = 0, nccl_calls = 0;

    void close_peers() {
        for (void *p : opened) cudaIpcCloseMemHandle(p);
        opened.clear();
        peer_ok = false;
    }
    ~Comm() {
        close_peers();
        if (area_owned) cudaFree(area_owned);
        if (karea_owned) cudaFree(karea_owned);
        if (peer_err_host) cudaFreeHost(peer_err_host);
        if (comm) Nccl::get().CommDestroy(comm);
    }
    bool active() const { return nranks > 1; }

    // Maps every rank's staging area into this process (CUDA IPC; the ranks of one node).  Any
    // failure on any rank leaves the whole job on NCCL: the decision is all-reduced.
    void peer_setup(cudaStream_t s) {
        if (!active() || nranks > kPeerMaxRanks) return;
        if (const char *e = getenv("OCFFM_PEER")) if (atoi(e) == 0) return;
        size_t cap = size_t(512) << 10;
        if (const char *e = getenv("OCFFM_PEER_CAP_KB")) cap = size_t(std::max(1, atoi(e))) << 10;
        struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
        static_assert(sizeof(Msg) == 80, "handle message layout");
        Msg mine{};
        mine.ok = 1;
        unsigned char *area = nullptr;
        const size_t bytes = peer_area_bytes(nranks, cap);
        if (cudaMalloc(&area, bytes) != cudaSuccess) { mine.ok = 0; area = nullptr; }
        area_owned = area;
        if (mine.ok && cudaMemsetAsync(area, 0, bytes, s) != cudaSuccess) mine.ok = 0;
        if (mine.ok && cudaIpcGetMemHandle(&mine.h, area) != cudaSuccess) mine.ok = 0;
        if (mine.ok && cudaHostAlloc(&peer_err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) mine.ok = 0;
        cudaGetLastError();
        DevBuf<unsigned char> x;
        x.alloc(sizeof(Msg) * nranks);
        OC_CUDA(cudaMemcpyAsync(x.p + sizeof(Msg) * rank, &mine, sizeof(Msg), cudaMemcpyHostToDevice, s));
        Nccl &n = Nccl::get();
        OC_NCCL(n.GroupStart());
        for (int q = 0; q < nranks; ++q)
            OC_NCCL(n.Broadcast(x.p + sizeof(Msg) * q, x.p + sizeof(Msg) * q, sizeof(Msg), ncclChar, q, comm, s));
        OC_NCCL(n.GroupEnd());
        std::vector<Msg> all(nranks);
        OC_CUDA(cudaMemcpyAsync(all.data(), x.p, sizeof(Msg) * nranks, cudaMemcpyDeviceToHost, s));
        OC_CUDA(cudaStreamSynchronize(s));
        bool ok = true;
        for (int q = 0; q < nranks; ++q) ok = ok && all[q].ok;
        pv.nranks = nranks;
        pv.rank = rank;
        pv.cap = cap;
        pv.base[rank] = area;
        if (ok) {
            for (int q = 0; q < nranks && ok; ++q) {
                if (q == rank) continue;
                void *p = nullptr;
                if (cudaIpcOpenMemHandle(&p, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    ok = false;
                    cudaGetLastError();
                } else {
                    opened.push_back(p);
                    pv.base[q] = static_cast<unsigned char *>(p);
                }
            }
            if (ok) {
                *peer_err_host = 0;
                ok = cudaHostGetDevicePointer(reinterpret_cast<void **>(&pv.error), peer_err_host, 0) == cudaSuccess;
            }
        }
        // every rank must take the same path
        DevBuf<float> agree;
        agree.alloc(1);
        const float v = ok ? 1.0f : 0.0f;
        OC_CUDA(cudaMemcpyAsync(agree.p, &v, sizeof(float), cudaMemcpyHostToDevice, s));
        OC_NCCL(n.AllReduce(agree.p, agree.p, 1, ncclFloat, ncclSum, comm, s));
        float total = 0;
        OC_CUDA(cudaMemcpyAsync(&total, agree.p, sizeof(float), cudaMemcpyDeviceToHost, s));
        OC_CUDA(cudaStreamSynchronize(s));
        peer_ok = total == float(nranks);
    }
    // Collective: maps the same cudaMalloc'ed buffer of every rank into this process.  Returns false
    // (on every rank alike) when any rank could not export or import a handle.
    bool share(void *mine, void **out, cudaStream_t s) {
        if (!peer_ok) return false;
        struct Msg { cudaIpcMemHandle_t h; int ok; int pad[3]; };
        Msg m{};
        m.ok = cudaIpcGetMemHandle(&m.h, mine) == cudaSuccess;
        cudaGetLastError();
        DevBuf<unsigned char> x;
        x.alloc(sizeof(Msg) * nranks);
        OC_CUDA(cudaMemcpyAsync(x.p + sizeof(Msg) * rank, &m, sizeof(Msg), cudaMemcpyHostToDevice, s));
        Nccl &n = Nccl::get();
        OC_NCCL(n.GroupStart());
        for (int q = 0; q < nranks; ++q)
            OC_NCCL(n.Broadcast(x.p + sizeof(Msg) * q, x.p + sizeof(Msg) * q, sizeof(Msg), ncclChar, q, comm, s));
        OC_NCCL(n.GroupEnd());
        std::vector<Msg> all(nranks);
        OC_CUDA(cudaMemcpyAsync(all.data(), x.p, sizeof(Msg) * nranks, cudaMemcpyDeviceToHost, s));
        OC_CUDA(cudaStreamSynchronize(s));
        bool ok = true;
        for (int q = 0; q < nranks; ++q) ok = ok && all[q].ok;
        for (int q = 0; q < nranks && ok; ++q) {
            out[q] = mine;
            if (q == rank) continue;
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = false;
                cudaGetLastError();
            } else {
                opened.push_back(p);
                out[q] = p;
            }
        }
        DevBuf<float> agree;
        agree.alloc(1);
        const float v = ok ? 1.0f : 0.0f;
        OC_CUDA(cudaMemcpyAsync(agree.p, &v, sizeof(float), cudaMemcpyHostToDevice, s));
        OC_NCCL(n.AllReduce(agree.p, agree.p, 1, ncclFloat, ncclSum, comm, s));
        float total = 0;
        OC_CUDA(cudaMemcpyAsync(&total, agree.p, sizeof(float), cudaMemcpyDeviceToHost, s));
        OC_CUDA(cudaStreamSynchronize(s));
        return total == float(nranks);
    }
    // second, tiny area for the scalar reductions INSIDE the persistent CG kernels (kernels.h PeerK)
    PeerK pk{};
    bool peerk_ok = false;
    uint4 *karea_owned = nullptr;
    void peerk_setup(cudaStream_t s) {
        pk.nranks = 1;
        if (!peer_ok) return;
        if (const char *e = getenv("OCFFM_PEER_KERNEL")) if (atoi(e) == 0) return;
        const size_t bytes = size_t(2) * nranks * sizeof(uint4);
        if (cudaMalloc(&karea_owned, bytes + sizeof(unsigned)) != cudaSuccess) { karea_owned = nullptr; cudaGetLastError(); }
        if (karea_owned) OC_CUDA(cudaMemsetAsync(karea_owned, 0, bytes + sizeof(unsigned), s));
        void *mapped[kPeerMaxRanks] = {};
        void *mine = karea_owned;   // share() is collective: a rank whose allocation failed exports a null handle -> all fall back
        const bool ok = share(mine, mapped, s);
        if (!ok || !karea_owned) return;
        for (int q = 0; q < nranks; ++q) pk.area[q] = static_cast<uint4 *>(mapped[q]);
        pk.nranks = nranks;
        pk.rank = rank;
        pk.seq = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(karea_owned) + bytes);
        pk.error = pv.error;
        peerk_ok = true;
    }
    unsigned long long gather_seq = 0;
    template <typename T>
    void peer_gather_rows(const PeerBuffers &pb, uint64_t rows, uint32_t ld, cudaStream_t s) {
        const uint64_t lo = rows * rank / nranks, hi = rows * (rank + 1) / nranks;
        peer_allgather_rows<T>(pv, pb, lo, hi, ld, ++gather_seq, s);
    }
    void check_peer() {
        if (peer_ok && *static_cast<volatile int *>(peer_err_host))
            throw Error(OCFFM_E_COMM, "peer all-reduce timed out waiting for another rank");
    }

    template <typename T>
    void allreduce(T *buf, size_t n, cudaStream_t s) {
        if (!active() || !n) return;
        if (peer_ok && n * sizeof(T) <= pv.cap) {
            ++peer_calls;
            peer_allreduce<T>(pv, buf, n, ++seq, s);
            return;
        }
        ++nccl_calls;
        OC_NCCL(Nccl::get().AllReduce(buf, buf, n, nccl_type<T>(), ncclSum, comm, s));
    }
    // all-gather of row slices with unequal sizes: rank q owns rows [lo(q), lo(q+1)) of a
    // [rows x ld] matrix and broadcasts them to everyone (one grouped launch)
    template <typename T>
    void allgather_rows(T *buf, uint64_t rows, size_t ld, cudaStream_t s) {
        if (!active()) return;
        Nccl &n = Nccl::get();
        OC_NCCL(n.GroupStart());
        for (int q = 0; q < nranks; ++q) {
            const uint64_t lo = rows * q / nranks, hi = rows * (q + 1) / nranks;
            if (hi > lo)
                OC_NCCL(n.Broadcast(buf + lo * ld, buf + lo * ld, (hi - lo) * ld, nccl_type<T>(), q,
                                    comm, s));
        }
        OC_NCCL(n.GroupEnd());
    }
};

// ---------------------------------------------------------------------------------------------
struct CtxBase {
    virtual ~CtxBase() {}
    virtual void comm_init(int nranks, int rank, const void *id) = 0;
    virtual void set_field(int side, uint32_t field, uint64_t rows, uint64_t D, const uint64_t *rowptr,
                           const uint32_t *idx, const double *val) = 0;
    virtual void set_labels(uint64_t m, const uint64_t *rowptr, const uint32_t *idx,
                            const uint64_t *cp, const uint32_t *ri, uint64_t n_ranked,
                            const double *popular) = 0;
    virtual void set_test_labels(uint64_t mt, const uint64_t *rowptr, const uint32_t *idx,
                                 const uint64_t *nnx) = 0;
    virtual void set_block(uint32_t f1, uint32_t f2, int which, const double *data, uint64_t rows) = 0;
    virtual void get_block(uint32_t f1, uint32_t f2, int which, double *data, uint64_t rows) = 0;
    virtual void mirror_block(uint32_t f1, uint32_t f2, int which, double *data, uint64_t rows) = 0;
    virtual void set_hyper(double lambda, double omega, double r) = 0;
    virtual void init_model(uint64_t seed) = 0;
    virtual void init_state() = 0;
    virtual void solve_block(uint32_t f1, uint32_t f2) = 0;
    virtual void one_epoch() = 0;
    virtual void grad(uint32_t f1, uint32_t f2, int which, double *G, uint64_t rows) = 0;
    virtual void hess_vec(uint32_t f1, uint32_t f2, int which, const double *V, double *Hv,
                          uint64_t rows) = 0;
    virtual void cg(uint32_t f1, uint32_t f2, int which, const double *G, double *S, uint64_t rows,
                    int32_t *iters) = 0;
    virtual void objective(double *value) = 0;
    virtual void validate(double *prec, double *ndcg, double *ploss, uint32_t *topk) = 0;
    virtual void get_vec(const char *name, double *out, uint64_t *count) = 0;
    virtual void get_embed(uint32_t f1, uint32_t f2, int which, double *out, uint64_t rows) = 0;
    virtual void get_csc(uint64_t *colptr, uint32_t *rowidx) = 0;
    virtual void get_stats(ocffm_stats *out) = 0;
    virtual void reset_stats() = 0;
    virtual void synchronize() = 0;
    virtual void *stream() = 0;
};

static uint32_t pad_k(uint32_t k) {
    uint32_t kp = 4;
    while (kp < k) kp <<= 1;
    return kp;
}

template <typename T>
struct Problem final : CtxBase {
    ocffm_params prm;
    uint32_t fu, fv, f, k, kp, Fx, Kc;
    uint64_t m, n, mt = 0, n_ranked = 0;
    int device = 0;
    cudaStream_t st = nullptr;
    Comm comm;
    uint32_t chunk = 64;
    bool diag_fast = true;      // OCFFM_DIAG_FAST=0 disables the fused same-side CG pass
    // OCFFM_NOTAU=1: on unit identity fields w * (V QTQ) is written into Hv by the row GEMM and the Hessian
    // row pass skips tau and the rows without pairs.  Measured (round 2): C2 18.62 -> 18.44 ms per outer
    // iteration, but C4 226 -> 233 ms (the GEMM epilogue re-reads V for the V.Hv share) -- off by default.
    bool notau_allowed = false;
    bool slice_cg = true;       // OCFFM_SLICE_CG=0: always replicate CG vectors across ranks
    // The reference keeps two copies of y-tilde (by user and by item, ffm.cpp:393,400) and adds every
    // update to both with the same arithmetic, so they stay bit-identical.  On one GPU the update
    // is applied to the orientation being swept only and the other one is refreshed by a permuted
    // copy right before it is next read (4 bytes gathered per entry instead of a d-wide row).
    // OCFFM_MIRROR_YT=0 updates both copies as the reference does; multi-rank runs always do
    // (each rank owns different slices of the two orientations).
    // multi-rank: the CG step S of every rank mapped here, so a sliced half publishes its part of
    // the step by direct NVLink stores (peer.cu) instead of an NCCL all-gather; OCFFM_PEER_GATHER=0
    bool peer_gather = false, peer_gather_allowed = true;
    PeerBuffers S_peers{};
    void *S_shared = nullptr;
    bool mirror_yt = false, mirror_allowed = true;
    // OCFFM_FUSED_DOT=0: separate direction / regulariser+dot kernels per CG iteration (5 instead of 3)
    bool fused_dot = true;
    uint32_t hot_min = 16384;   // OCFFM_HOT_MIN: occurrences that make a feature "hot" (0 = off)
    // Per-row observed Gram for the Hessian passes of cross halves (rows.cu "Mrow").  Default: fp32
    // contexts with kp = 32 (at kp 64 building the Gram costs as much as the gathers it saves, at kp 16
    // it lost 10 ms per outer iteration on the Outbrain shape; fp64 contexts are the strict-parity mode
    // and keep the reference's summation structure).
    // Building the blocks costs k*k FMAs per pair (16 gather passes' worth of FLOPs at k = 32), so it
    // pays only for half solves with many CG iterations: by default (OCFFM_MROW=1, fp32) a half builds
    // them when its previous solve took >= OCFFM_MROW_ITERS (8) iterations.  OCFFM_MROW=0 never,
    // =2 always and also in fp64 contexts (tests); OCFFM_MROW_MIN: pairs per heavy row.
    int mrow_mode = 1;
    int mrow_iters = 8;
    std::vector<int> last_iters;   // CG iterations of the previous solve of every half (index 2*block + which)
    uint32_t mrow_min = 48;
    size_t mrow_cap_bytes = size_t(3) << 30;
    bool mrow_on = false, mrow_ready = false;
    DevBuf<T> mrow;
    uint64_t mrow_builds = 0;
    bool profile = false;

    struct Field {
        bool set = false;
        uint64_t rows = 0, D = 0, nnz = 0, nnz_local = 0;
        DevBuf<uint32_t> rowptr, idx, hot_feat;
        DevBuf<T> val, freq, shadow;
        DevBuf<int16_t> hot_slot;
        uint32_t n_hot = 0;
        bool diagonal = false;   // one feature per row, features form a permutation of 0..D-1
        bool identity = false;   // diagonal with idx[i] == i: a rank's rows touch only its own feature slice
        bool unit_identity = false;   // identity and every value is 1: X is the identity matrix
        uint32_t row0 = 0, row1 = 0;
        CsrView<T> view() const {
            return {rowptr.p, idx.p, val.p, row0, row1, n_hot ? hot_slot.p : nullptr, shadow.p, diagonal, identity};
        }
        CsrView<T> view_all() const {
            return {rowptr.p, idx.p, val.p, 0, uint32_t(rows), n_hot ? hot_slot.p : nullptr, shadow.p, diagonal,
                    identity};
        }
    };
    struct Omega {
        bool set = false;
        uint64_t rows = 0, nnz = 0, nnz_local = 0;
        uint32_t n_items = 0, n_items_nonempty = 0, row0 = 0, row1 = 0;
        DevBuf<uint32_t> rowptr, idx, wi_row, wi_beg, wi_cnt;
        DevBuf<T> yt;
        DevBuf<uint32_t> mirror_pos;      // position of each entry in the other orientation (mirror_yt)
        bool fresh = true;                // yt holds every update made so far
        std::vector<uint64_t> h_rowptr;   // kept for get_csc / stats
        std::vector<uint32_t> h_idx;
        // Sharded storage: only this rank's rows live in HBM -- rowptr[row0..row1], and idx / yt for
        // the nnz [base, base + nnz_local) of those rows.  Kernels keep using GLOBAL row numbers and
        // nnz positions, so they get the device pointers shifted back by row0 / base.
        uint64_t base = 0;
        const uint32_t *rowptr_v() const { return reinterpret_cast<const uint32_t *>(uintptr_t(rowptr.p) - size_t(row0) * 4); }
        const uint32_t *idx_v() const { return reinterpret_cast<const uint32_t *>(uintptr_t(idx.p) - size_t(base) * 4); }
        T *yt_v() const { return reinterpret_cast<T *>(uintptr_t(yt.p) - size_t(base) * sizeof(T)); }
        OmegaView<T> view() const {
            return {wi_row.p, wi_beg.p, wi_cnt.p, n_items, rowptr_v(), idx_v(), yt_v(), row0, row1, nnz_local};
        }
        // Per-row observed Gram (rows.cu "Mrow"): the local rows with >= mrow_min pairs are HEAVY --
        // slot s of the Gram buffer belongs to row heavy_rows[s]; hw_* are the build items (<= 1024
        // pairs each, rows split over several items come first and are the zeroed prefix
        // [0, n_multi)); lw_* is the ordinary work-item list restricted to the other (LIGHT) rows.
        DevBuf<uint32_t> heavy_rows, hw_slot, hw_beg, hw_cnt, lw_row, lw_beg, lw_cnt;
        uint32_t n_heavy = 0, n_multi = 0, n_hw = 0, n_lw = 0;
        uint64_t nnz_heavy = 0;
        uint32_t n_lw_nonempty = 0;
        OmegaView<T> light_view(bool nonempty_only = false) const {
            return {lw_row.p, lw_beg.p, lw_cnt.p, nonempty_only ? n_lw_nonempty : n_lw, rowptr_v(), idx_v(), yt_v(), row0,
                    row1, nnz_local};
        }
        OmegaView<T> view_nonempty() const {
            return {wi_row.p, wi_beg.p, wi_cnt.p, n_items_nonempty, rowptr_v(), idx_v(), yt_v(), row0, row1, nnz_local};
        }
    };
    struct Block {
        bool exists = false, side = false, has_w = false, has_h = false;
        uint32_t f1 = 0, f2 = 0;
        int pair = -1;
        DevBuf<T> W, H, P, Q;
        double *mirW = nullptr, *mirH = nullptr;   // registered pinned host mirrors (ocffm_mirror_block)
    };

    std::vector<Field> XU, XV, XT;
    Omega YU, YV;
    DevBuf<uint32_t> t_rowptr, t_idx, topk_ids, cold_ids, part_id;
    DevBuf<T> part_score, ev_Pva, ev_Qva, ev_at, ev_bt, ev_t1, ev_t2;
    DevBuf<float> tc_phi, tc_plo, tc_qhi, tc_qlo, tc_cand_score;
    DevBuf<uint32_t> tc_cand_id, tc_row_thr;
    bool cold_ready = false;
    bool eval_tc = true;   // OCFFM_EVAL_TC=0 forces the SIMT scorer
    DevBuf<uint8_t> t_cold;
    std::vector<uint8_t> h_cold;
    bool test_set = false;
    uint32_t t_row0 = 0, t_row1 = 0;
    DevBuf<T> popular;
    std::vector<double> h_popular;
    std::vector<Block> blocks;
    DevBuf<T> Pc, Qc, a, b, sa, sb;
    bool state_ready = false;

    // scratch
    DevBuf<T> G, S, R, V, VQ, Hv, Tm, XS, gap, ysum, GT, oQ, bQ, vecKc;
    DevBuf<double> gram64, acc64;
    SolveScalars *sc = nullptr;
    // pinned AND mapped: the CG kernels store g2 / r2 / iteration counts straight into host memory, so the
    // host never issues a small D2H memcpy inside a solve (it would queue behind the bulk DMA of the host
    // mirrors on the same copy engine)
    double *h_scal = nullptr, *d_hscal = nullptr;
    unsigned *d_hiters = nullptr;
    float *h_phase = nullptr, *d_hphase = nullptr;   // mapped pinned: row-phase time of every persistent cross launch
    double *h_stage = nullptr; // pinned staging for model blocks crossing the ABI as fp64
    size_t h_stage_n = 0;
    DevBuf<double> d_stage;
    cudaEvent_t cg_ev[24], g2_ev;
    // host mirrors: blocks are streamed out on a second stream while the rest of the iteration runs
    // (conversion to fp64 runs on the compute stream -- a tenth of a millisecond per block -- into one of
    // kMirSlots staging buffers; only the DMA runs on the copy stream, so no second-stream kernel has to
    // fight the solver's CTAs for SM slots)
    static constexpr int kMirSlots = 8;
    cudaStream_t copy_st = nullptr;
    cudaEvent_t mir_ready[kMirSlots] = {}, mir_free[kMirSlots] = {};
    DevBuf<double> mir_stage[kMirSlots];
    bool mir_used[kMirSlots] = {};
    int mir_next = 0;
    uint64_t mirrors = 0, mirrored_bytes = 0;

    // stats
    uint64_t launches = 0, cg_iters = 0, nnz_trav = 0, algo_bytes = 0, hv_launches = 0,
             hv_algo_bytes = 0;
    double ms[6] = {0, 0, 0, 0, 0, 0}, hv_ms = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> hv_events;
    size_t hv_events_used = 0;

    Problem(const ocffm_params &p, uint32_t fu_, uint32_t fv_, uint64_t m_, uint64_t n_)
        : prm(p), fu(fu_), fv(fv_), f(fu_ + fv_), k(p.k), m(m_), n(n_) {
        OC_REQUIRE(fu >= 1 && fv >= 1, "need at least one user field and one item field");
        OC_REQUIRE(k >= 1 && k <= 128, "k must be in 1..128");
        OC_REQUIRE(m < (1ull << 31) && n < (1ull << 31), "row counts must be below 2^31");
        kp = pad_k(k);
        Fx = fu * fv;
        Kc = Fx * kp;
        if (p.device >= 0) OC_CUDA(cudaSetDevice(p.device));
        OC_CUDA(cudaGetDevice(&device));
        OC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        OC_CUDA(cudaMalloc(&sc, sizeof(SolveScalars)));
        OC_CUDA(cudaHostAlloc(&h_scal, 64 * sizeof(double), cudaHostAllocMapped));
        OC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_hscal), h_scal, 0));
        for (auto &e : cg_ev) OC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        OC_CUDA(cudaEventCreateWithFlags(&g2_ev, cudaEventDisableTiming));
        XU.resize(fu);
        XV.resize(fv);
        XT.resize(fu);
        blocks.resize(size_t(f) * (f + 1) / 2);
        last_iters.assign(blocks.size() * 2, 0);
        for (uint32_t f1 = 0; f1 < f; ++f1)
            for (uint32_t f2 = f1; f2 < f; ++f2) {
                Block &bk = blocks[bidx(f1, f2)];
                bk.f1 = f1;
                bk.f2 = f2;
                bk.side = (f1 < fu && f2 < fu) || (f1 >= fu && f2 >= fu);
                bk.exists = prm.self_side || !bk.side;   // ffm.cpp:502-503
                if (!bk.side) bk.pair = int(f1 * fv + (f2 - fu));
            }
        if (const char *e = getenv("OCFFM_CHUNK")) chunk = std::max(1, atoi(e));
        if (const char *e = getenv("OCFFM_EVAL_TC")) eval_tc = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_MIRROR_YT")) mirror_allowed = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_PEER_GATHER")) peer_gather_allowed = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_FUSED_DOT")) fused_dot = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_HOT_MIN")) hot_min = uint32_t(std::max(0, atoi(e)));
        if (const char *e = getenv("OCFFM_MROW")) mrow_mode = atoi(e);
        if (const char *e = getenv("OCFFM_MROW_MIN")) mrow_min = uint32_t(std::max(1, atoi(e)));
        if (const char *e = getenv("OCFFM_MROW_ITERS")) mrow_iters = std::max(0, atoi(e));
        if (const char *e = getenv("OCFFM_MROW_CAP_MB")) mrow_cap_bytes = size_t(std::max(1, atoi(e))) << 20;
        mrow_on = mrow_mode != 0 && row_gram_supported(int(kp)) &&
                  ((std::is_same<T, float>::value && kp == 32) || mrow_mode >= 2);
        if (const char *e = getenv("OCFFM_DIAG_FAST")) diag_fast = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_NOTAU")) notau_allowed = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_PERSIST_CG")) { persist_mode = atoi(e); persist_on = persist_mode != 0; }
        if (const char *e = getenv("OCFFM_SLICE_CG")) slice_cg = atoi(e) != 0;
        if (const char *e = getenv("OCFFM_PROFILE")) { profile_level = atoi(e); profile = profile_level != 0; }
        a.alloc(m); b.alloc(n); sa.alloc(m); sb.alloc(n);
        a.zero(st); b.zero(st); sa.zero(st); sb.zero(st);
        Pc.alloc(m * Kc); Qc.alloc(n * Kc);
        const uint64_t mx = std::max(m, n);
        Tm.alloc(mx * kp); XS.alloc(mx * kp); gap.alloc(mx); ysum.alloc(mx);
        GT.alloc(size_t(Kc) * kp); oQ.alloc(kp); bQ.alloc(kp); vecKc.alloc(Kc);
        gram64.alloc(size_t(Kc) * kp + 2 * kp);
        acc64.alloc(64);
        sync();
    }
    ~Problem() override {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);
        comm.close_peers();   // imports first, then this rank's own buffers are freed
        for (auto &e : hv_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto &e : cgk_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto &e : cg_ev) cudaEventDestroy(e);
        cudaEventDestroy(g2_ev);
        if (sc) cudaFree(sc);
        if (h_scal) cudaFreeHost(h_scal);
        if (h_iters) cudaFreeHost(h_iters);
        if (h_phase) cudaFreeHost(h_phase);
        if (h_stage) cudaFreeHost(h_stage);
        if (copy_st) { cudaStreamSynchronize(copy_st); cudaStreamDestroy(copy_st); }
        for (int i = 0; i < kMirSlots; ++i) {
            if (mir_ready[i]) cudaEventDestroy(mir_ready[i]);
            if (mir_free[i]) cudaEventDestroy(mir_free[i]);
        }
        if (st) cudaStreamDestroy(st);
    }

    // index_vec, ffm.cpp:53-55
    size_t bidx(uint32_t f1, uint32_t f2) const { return f2 + size_t(f - 1) * f1 - size_t(f1) * (f1 - 1) / 2; }
    void sync() {
        OC_CUDA(cudaStreamSynchronize(st));
        comm.check_peer();
        drain_pending();
    }
    void bind() {
        OC_CUDA(cudaSetDevice(device));
        g_launch_counter = &launches;
    }
    Field &field_of(uint32_t fg) { return fg < fu ? XU[fg] : XV[fg - fu]; }
    uint64_t rows_of(uint32_t fg) const { return fg < fu ? m : n; }
    Block &block(uint32_t f1, uint32_t f2) {
        OC_REQUIRE(f1 <= f2 && f2 < f, "block indices must satisfy f1 <= f2 < f");
        Block &bk = blocks[bidx(f1, f2)];
        OC_REQUIRE(bk.exists, "block does not exist (same-side block under --ns)");
        return bk;
    }

    // ------------------------------------------------------------------------------------------
    void comm_init(int nranks, int rank, const void *id) override {
        OC_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
        OC_REQUIRE(!YU.set, "ocffm_comm_init must precede ocffm_set_labels");
        comm.nranks = nranks;
        comm.rank = rank;
        if (nranks > 1) {
            ncclUniqueId uid;
            memcpy(&uid, id, sizeof(uid));
            OC_NCCL(Nccl::get().CommInitRank(&comm.comm, nranks, uid, rank));
            comm.peer_setup(st);
            comm.peerk_setup(st);
        }
    }
    uint32_t lo(uint64_t rows) const { return uint32_t(rows * comm.rank / comm.nranks); }
    uint32_t hi(uint64_t rows) const { return uint32_t(rows * (comm.rank + 1) / comm.nranks); }

    void set_field(int side, uint32_t field, uint64_t rows, uint64_t D, const uint64_t *rowptr,
                   const uint32_t *idx, const double *val) override {
        OC_REQUIRE(side >= 0 && side <= 2, "side must be OCFFM_SIDE_U/V/T");
        std::vector<Field> &vec = side == OCFFM_SIDE_U ? XU : side == OCFFM_SIDE_V ? XV : XT;
        OC_REQUIRE(field < vec.size(), "field out of range");
        const uint64_t want = side == OCFFM_SIDE_U ? m : side == OCFFM_SIDE_V ? n : rows;
        OC_REQUIRE(rows == want, "row count does not match the context");
        OC_REQUIRE(rows < (1ull << 31) && D < (1ull << 31), "sizes must be below 2^31");
        const uint64_t nnz = rowptr[rows];
        OC_REQUIRE(nnz < (1ull << 32), "nnz must be below 2^32");
        Field &F = vec[field];
        F.rows = rows; F.D = D; F.nnz = nnz;
        std::vector<uint32_t> rp(rows + 1);
        for (uint64_t i = 0; i <= rows; ++i) rp[i] = uint32_t(rowptr[i]);
        // freq (ffm.cpp:235-241) is counted in integers like the reference's ImpLong: a float counter
        // stops growing at 2^24 occurrences
        std::vector<T> v(nnz), fr(D);
        std::vector<uint64_t> occ(D, 0);
        for (uint64_t t = 0; t < nnz; ++t) {
            OC_REQUIRE(idx[t] < D, "feature index >= D");
            v[t] = T(val[t]);
            ++occ[idx[t]];
        }
        for (uint64_t d = 0; d < D; ++d) fr[d] = T(occ[d]);
        F.diagonal = F.identity = F.unit_identity = false;
        if (nnz == rows && D == rows && rows > 0) {
            bool ok = true;
            for (uint64_t i = 0; i < rows && ok; ++i) ok = rowptr[i] == i && occ[idx[i]] == 1;
            F.diagonal = ok;
            bool id = ok;
            for (uint64_t i = 0; i < rows && id; ++i) id = idx[i] == i;
            F.identity = id;
            bool unit = id;
            for (uint64_t i = 0; i < rows && unit; ++i) unit = val[i] == 1.0;
            F.unit_identity = unit;
        }
        F.rowptr.upload(rp, st);
        F.idx.upload(idx, nnz, st);
        F.val.upload(v, st);
        F.freq.upload(fr, st);
        // hot features: those hit by >= hot_min scatter contributions per pass (at most 1024,
        // the most frequent first); they get kHotReplicas shadow rows each (kernels.h CsrView)
        F.n_hot = 0;
        if (side != OCFFM_SIDE_T && hot_min > 0) {
            std::vector<std::pair<uint64_t, uint32_t>> cand;
            for (uint64_t d = 0; d < D; ++d)
                if (occ[d] >= hot_min) cand.emplace_back(occ[d], uint32_t(d));
            std::sort(cand.begin(), cand.end(), [](const auto &x, const auto &y) { return x.first > y.first; });
            if (cand.size() > 1024) cand.resize(1024);
            if (!cand.empty()) {
                std::vector<int16_t> slot(D, int16_t(-1));
                std::vector<uint32_t> feat(cand.size());
                for (size_t i = 0; i < cand.size(); ++i) { slot[cand[i].second] = int16_t(i); feat[i] = cand[i].second; }
                F.hot_slot.upload(slot, st);
                F.hot_feat.upload(feat, st);
                F.shadow.alloc(cand.size() * size_t(kHotReplicas) * kp);
                F.n_hot = uint32_t(cand.size());
            }
        }
        F.row0 = lo(rows);
        F.row1 = hi(rows);
        F.nnz_local = rowptr[F.row1] - rowptr[F.row0];
        F.set = true;
        sync();
        if (side == OCFFM_SIDE_T) { mt = rows; t_row0 = lo(rows); t_row1 = hi(rows); }
        state_ready = false;
    }

    void build_omega(Omega &Y, uint64_t rows, const uint64_t *rowptr, const uint32_t *idx) {
        Y.rows = rows;
        Y.nnz = rowptr[rows];
        OC_REQUIRE(Y.nnz < (1ull << 31), "|Omega| must be below 2^31");
        Y.h_rowptr.assign(rowptr, rowptr + rows + 1);
        Y.h_idx.assign(idx, idx + Y.nnz);
        Y.row0 = lo(rows);
        Y.row1 = hi(rows);
        Y.base = rowptr[Y.row0];
        Y.nnz_local = rowptr[Y.row1] - rowptr[Y.row0];
        std::vector<uint32_t> rp(Y.row1 - Y.row0 + 1);
        for (uint64_t i = Y.row0; i <= Y.row1; ++i) rp[i - Y.row0] = uint32_t(rowptr[i]);
        // work items of the rows with pairs first, then one empty item per row without any: a pass that
        // has nothing to add for an empty row (hess_cross with notau) launches on the prefix only
        std::vector<uint32_t> wr, wb, wc;
        wr.reserve(Y.row1 - Y.row0 + Y.nnz_local / chunk);
        wb.reserve(wr.capacity());
        wc.reserve(wr.capacity());
        for (uint32_t i = Y.row0; i < Y.row1; ++i) {
            const uint64_t b0 = rowptr[i], e0 = rowptr[i + 1];
            for (uint64_t t = b0; t < e0; t += chunk) {
                wr.push_back(i);
                wb.push_back(uint32_t(t));
                wc.push_back(uint32_t(std::min<uint64_t>(chunk, e0 - t)) | (t == b0 ? 0x80000000u : 0u));
            }
        }
        Y.n_items_nonempty = uint32_t(wr.size());
        for (uint32_t i = Y.row0; i < Y.row1; ++i)
            if (rowptr[i] == rowptr[i + 1]) { wr.push_back(i); wb.push_back(uint32_t(rowptr[i])); wc.push_back(0x80000000u); }
        Y.n_items = uint32_t(wr.size());
        build_heavy_lists(Y, rowptr);
        Y.rowptr.upload(rp, st);
        Y.idx.upload(idx + Y.base, Y.nnz_local, st);
        Y.wi_row.upload(wr, st);
        Y.wi_beg.upload(wb, st);
        Y.wi_cnt.upload(wc, st);
        Y.yt.alloc(Y.nnz_local);
        Y.yt.zero(st);
        Y.set = true;
        sync();
    }

    // heavy / light split of the local rows for the per-row Gram path (see Omega)
    void build_heavy_lists(Omega &Y, const uint64_t *rowptr) {
        Y.n_heavy = Y.n_multi = Y.n_hw = Y.n_lw = 0;
        Y.nnz_heavy = 0;
        if (!mrow_on) return;
        const uint32_t kItem = 1024;
        // raise the threshold until the Gram buffer fits the cap
        uint64_t thr = mrow_min;
        for (;;) {
            uint64_t cntr = 0;
            for (uint32_t i = Y.row0; i < Y.row1; ++i) cntr += (rowptr[i + 1] - rowptr[i]) >= thr;
            if (cntr * kp * kp * sizeof(T) <= mrow_cap_bytes) break;
            thr *= 2;
        }
        std::vector<uint32_t> multi, single;
        for (uint32_t i = Y.row0; i < Y.row1; ++i) {
            const uint64_t c = rowptr[i + 1] - rowptr[i];
            if (c >= thr) (c > kItem ? multi : single).push_back(i);
        }
        // longest first: the build kernel's CTAs finish together
        auto by_len = [&](uint32_t x, uint32_t y) { return rowptr[x + 1] - rowptr[x] > rowptr[y + 1] - rowptr[y]; };
        std::stable_sort(multi.begin(), multi.end(), by_len);
        std::stable_sort(single.begin(), single.end(), by_len);
        std::vector<uint32_t> hr(multi);
        hr.insert(hr.end(), single.begin(), single.end());
        std::vector<uint32_t> hs, hb, hc, lr, lb, lc;
        for (uint32_t s = 0; s < hr.size(); ++s) {
            const uint64_t b0 = rowptr[hr[s]], e0 = rowptr[hr[s] + 1];
            Y.nnz_heavy += e0 - b0;
            const bool split = e0 - b0 > kItem;
            for (uint64_t t = b0; t < e0; t += kItem) {
                hs.push_back(s);
                hb.push_back(uint32_t(t));
                hc.push_back(uint32_t(std::min<uint64_t>(kItem, e0 - t)) | (split ? 0x80000000u : 0u));
            }
        }
        for (uint32_t i = Y.row0; i < Y.row1; ++i) {
            const uint64_t b0 = rowptr[i], e0 = rowptr[i + 1];
            if (e0 - b0 >= thr) continue;
            for (uint64_t t = b0; t < e0; t += chunk) {
                lr.push_back(i);
                lb.push_back(uint32_t(t));
                lc.push_back(uint32_t(std::min<uint64_t>(chunk, e0 - t)) | (t == b0 ? 0x80000000u : 0u));
            }
        }
        Y.n_lw_nonempty = uint32_t(lr.size());
        for (uint32_t i = Y.row0; i < Y.row1; ++i)
            if (rowptr[i] == rowptr[i + 1] && thr > 0) { lr.push_back(i); lb.push_back(uint32_t(rowptr[i])); lc.push_back(0x80000000u); }
        Y.n_heavy = uint32_t(hr.size());
        Y.n_multi = uint32_t(multi.size());
        Y.n_hw = uint32_t(hs.size());
        Y.n_lw = uint32_t(lr.size());
        if (!Y.n_heavy) return;
        Y.heavy_rows.upload(hr, st);
        Y.hw_slot.upload(hs, st); Y.hw_beg.upload(hb, st); Y.hw_cnt.upload(hc, st);
        Y.lw_row.upload(lr, st); Y.lw_beg.upload(lb, st); Y.lw_cnt.upload(lc, st);
    }

    void set_labels(uint64_t m_, const uint64_t *rowptr, const uint32_t *idx, const uint64_t *cp,
                    const uint32_t *ri, uint64_t n_ranked_, const double *pop) override {
        OC_REQUIRE(m_ == m, "label row count does not match the context");
        const uint64_t nnz = rowptr[m];
        build_omega(YU, m, rowptr, idx);
        // U->n and popular (ffm.cpp:97, 143, 172-176)
        uint64_t un = 0;
        for (uint64_t t = 0; t < nnz; ++t) un = std::max<uint64_t>(un, uint64_t(idx[t]) + 1);
        n_ranked = n_ranked_ ? n_ranked_ : un;
        OC_REQUIRE(n_ranked >= un, "n_ranked smaller than max label + 1");
        h_popular.assign(n_ranked, 0.0);
        if (pop) {
            std::copy(pop, pop + n_ranked, h_popular.begin());
        } else {
            for (uint64_t t = 0; t < nnz; ++t) h_popular[idx[t]] += 1.0;
            double tot = 0;
            for (double v : h_popular) tot += v;
            for (double &v : h_popular) v /= tot;
        }
        std::vector<T> pv(h_popular.begin(), h_popular.end());
        popular.upload(pv, st);
        cold_ready = false;
        // CSC by (item, user): transY (ffm.cpp:259-294); labels >= n are dropped there
        std::vector<uint64_t> colptr;
        std::vector<uint32_t> rowidx;
        if (!cp || !ri) {
            colptr.assign(n + 1, 0);
            uint64_t kept = 0;
            for (uint64_t t = 0; t < nnz; ++t)
                if (idx[t] < n) { colptr[idx[t] + 1]++; kept++; }
            for (uint64_t j = 0; j < n; ++j) colptr[j + 1] += colptr[j];
            rowidx.resize(kept);
            std::vector<uint64_t> cur(colptr.begin(), colptr.end() - 1);
            for (uint64_t i = 0; i < m; ++i)
                for (uint64_t t = rowptr[i]; t < rowptr[i + 1]; ++t)
                    if (idx[t] < n) rowidx[cur[idx[t]]++] = uint32_t(i);
            cp = colptr.data();
            ri = rowidx.data();
        }
        OC_REQUIRE(cp[n] == nnz, "labels >= number of items are not supported (the reference leaves "
                                 "its two copies of Y inconsistent, ffm.cpp:267-268)");
        build_omega(YV, n, cp, ri);
        mirror_yt = false;
        if (!comm.active() && mirror_allowed && nnz) {
            // entry t of the CSR <-> entry pos of the CSC; needs the columns listed by ascending user
            std::vector<uint32_t> u2v(nnz), v2u(nnz);
            std::vector<uint64_t> cur(cp, cp + n);
            bool ok = true;
            for (uint64_t i = 0; i < m && ok; ++i)
                for (uint64_t t = rowptr[i]; t < rowptr[i + 1]; ++t) {
                    const uint64_t pos = cur[idx[t]]++;
                    if (pos >= cp[idx[t] + 1] || ri[pos] != i) { ok = false; break; }
                    u2v[t] = uint32_t(pos);
                    v2u[pos] = uint32_t(t);
                }
            if (ok) {
                YU.mirror_pos.upload(u2v, st);
                YV.mirror_pos.upload(v2u, st);
                sync();
                mirror_yt = true;
            }
        }
        YU.fresh = YV.fresh = true;
        state_ready = false;
    }
    // make Y.yt current before it is read or updated in place
    void freshen(Omega &Y) {
        if (!mirror_yt || Y.fresh) return;
        Omega &other = &Y == &YU ? YV : YU;
        gather_copy<T>(Y.yt.p, other.yt.p, Y.mirror_pos.p, Y.nnz, st);
        algo_bytes += Y.nnz * (4 + 2 * sizeof(T));
        Y.fresh = true;
    }

    void set_test_labels(uint64_t mt_, const uint64_t *rowptr, const uint32_t *idx,
                         const uint64_t *nnx) override {
        for (uint32_t fi = 0; fi < fu; ++fi)
            OC_REQUIRE(XT[fi].set && XT[fi].rows == mt_, "set every OCFFM_SIDE_T field first");
        mt = mt_;
        const uint64_t nnz = rowptr[mt];
        OC_REQUIRE(nnz < (1ull << 31), "too many test labels");
        std::vector<uint32_t> rp(mt + 1);
        for (uint64_t i = 0; i <= mt; ++i) rp[i] = uint32_t(rowptr[i]);
        t_rowptr.upload(rp, st);
        t_idx.upload(idx, nnz, st);
        h_cold.assign(mt, 0);
        if (nnx) {
            for (uint64_t i = 0; i < mt; ++i) h_cold[i] = nnx[i] == 0;
        } else {
            std::vector<uint64_t> cnt(mt, 0);
            std::vector<uint32_t> frp;
            for (uint32_t fi = 0; fi < fu; ++fi) {
                frp.resize(mt + 1);
                XT[fi].rowptr.download(frp.data(), mt + 1, st);
                sync();
                for (uint64_t i = 0; i < mt; ++i) cnt[i] += frp[i + 1] - frp[i];
            }
            for (uint64_t i = 0; i < mt; ++i) h_cold[i] = cnt[i] == 0;
        }
        t_cold.upload(h_cold, st);
        topk_ids.alloc(mt * 80);
        cold_ids.alloc(80);
        t_row0 = lo(mt);
        t_row1 = hi(mt);
        sync();
        test_set = true;
    }

    // ------------------------------------------------------------------------------------------
    DevBuf<T> &block_mat(Block &bk, int which) { return which == 'W' ? bk.W : bk.H; }
    uint64_t block_rows(const Block &bk, int which) {
        return which == 'W' ? field_of(bk.f1).D : field_of(bk.f2).D;
    }
    void ensure_stage(size_t n_) {
        if (n_ <= h_stage_n) return;
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        OC_CUDA(cudaMallocHost(&h_stage, n_ * sizeof(double)));
        h_stage_n = n_;
    }
    static bool is_pinned(const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    }
    void ensure_dstage(size_t n_) {
        if (d_stage.n < n_) d_stage.alloc(n_);
    }
    // fp64 [rows x k] host -> device -> T [rows x kp] (conversion + zero padding on the device).
    // Pinned caller buffers are DMA'd directly; pageable ones go through a pinned staging copy.
    void upload_padded(DevBuf<T> &dst, const double *src, uint64_t rows) {
        const size_t cnt = rows * k;
        ensure_dstage(cnt);
        const double *from = src;
        if (!is_pinned(src)) {
            ensure_stage(cnt);
            memcpy(h_stage, src, cnt * sizeof(double));
            from = h_stage;
        }
        OC_CUDA(cudaMemcpyAsync(d_stage.p, from, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
        dst.ensure(rows * kp);
        pad_from_f64<T>(d_stage.p, dst.p, rows, k, kp, st);
        sync();
    }
    void download_unpadded(const T *src, uint64_t ld, double *dst, uint64_t rows) {
        const size_t cnt = rows * k;
        ensure_dstage(cnt);
        unpad_to_f64<T>(src, uint32_t(ld), d_stage.p, rows, k, st);
        if (is_pinned(dst)) {
            OC_CUDA(cudaMemcpyAsync(dst, d_stage.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
            sync();
            return;
        }
        ensure_stage(cnt);
        OC_CUDA(cudaMemcpyAsync(h_stage, d_stage.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
        sync();
        memcpy(dst, h_stage, cnt * sizeof(double));
    }
    void set_block(uint32_t f1, uint32_t f2, int which, const double *data, uint64_t rows) override {
        OC_REQUIRE(which == 'W' || which == 'H', "which must be 'W' or 'H'");
        Block &bk = block(f1, f2);
        OC_REQUIRE(field_of(which == 'W' ? f1 : f2).set, "set the block's field before its parameters");
        OC_REQUIRE(rows == block_rows(bk, which), "rows must equal Ds of the block's field");
        upload_padded(block_mat(bk, which), data, rows);
        (which == 'W' ? bk.has_w : bk.has_h) = true;
        state_ready = false;
    }
    void get_block(uint32_t f1, uint32_t f2, int which, double *data, uint64_t rows) override {
        OC_REQUIRE(which == 'W' || which == 'H', "which must be 'W' or 'H'");
        Block &bk = block(f1, f2);
        OC_REQUIRE(which == 'W' ? bk.has_w : bk.has_h, "block was never set");
        OC_REQUIRE(rows == block_rows(bk, which), "rows must equal Ds of the block's field");
        download_unpadded(block_mat(bk, which).p, kp, data, rows);
    }

    // The reference keeps W / H in host memory and one_epoch() updates them in place.  A caller that
    // needs the host copy current after every outer iteration registers pinned fp64 mirrors; one_epoch
    // then converts + copies each block on a second stream right after the block's solve, overlapped
    // with the remaining block solves, instead of a serial download of the whole model afterwards.
    void mirror_block(uint32_t f1, uint32_t f2, int which, double *data, uint64_t rows) override {
        OC_REQUIRE(which == 'W' || which == 'H', "which must be 'W' or 'H'");
        Block &bk = block(f1, f2);
        OC_REQUIRE(rows == block_rows(bk, which), "rows must equal Ds of the block's field");
        OC_REQUIRE(!data || is_pinned(data), "a host mirror must be pinned (page-locked) memory");
        double *&slot = which == 'W' ? bk.mirW : bk.mirH;
        if (!slot && data) ++mirrors;
        if (slot && !data) --mirrors;
        slot = data;
        if (data && !copy_st) {
            OC_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
            for (int i = 0; i < kMirSlots; ++i) {
                OC_CUDA(cudaEventCreateWithFlags(&mir_ready[i], cudaEventDisableTiming));
                OC_CUDA(cudaEventCreateWithFlags(&mir_free[i], cudaEventDisableTiming));
            }
        }
    }
    void stream_out(Block &bk) {
        for (int which : {'W', 'H'}) {
            double *dst = which == 'W' ? bk.mirW : bk.mirH;
            if (!dst) continue;
            const uint64_t rows = block_rows(bk, which);
            const int slot = mir_next;
            mir_next = (mir_next + 1) % kMirSlots;
            if (mir_used[slot]) OC_CUDA(cudaStreamWaitEvent(st, mir_free[slot], 0));   // its previous DMA has drained
            unpad_to_f64<T>(block_mat(bk, which).p, kp, mir_stage[slot].p, rows, k, st);
            OC_CUDA(cudaEventRecord(mir_ready[slot], st));
            OC_CUDA(cudaStreamWaitEvent(copy_st, mir_ready[slot], 0));
            OC_CUDA(cudaMemcpyAsync(dst, mir_stage[slot].p, rows * k * sizeof(double), cudaMemcpyDeviceToHost, copy_st));
            OC_CUDA(cudaEventRecord(mir_free[slot], copy_st));
            mir_used[slot] = true;
            mirrored_bytes += rows * k * sizeof(double);
        }
    }

    // ------------------------------------------------------------------------------------------
    // cache_sasb, ffm.cpp:514-535: sa = Pc colsum(Qc), sb = Qc colsum(Pc)
    void cache_sasb() {
        double *cs = gram64.p;   // >= Kc doubles
        OC_CUDA(cudaMemsetAsync(cs, 0, Kc * sizeof(double), st));
        col_sums<T>(Qc.p, Kc, Kc, 0, uint32_t(n), cs, st);
        convert_from_f64<T>(cs, vecKc.p, Kc, st);
        matvec_rows<T>(Pc.p, Kc, Kc, uint32_t(m), vecKc.p, sa.p, st);
        OC_CUDA(cudaMemsetAsync(cs, 0, Kc * sizeof(double), st));
        col_sums<T>(Pc.p, Kc, Kc, 0, uint32_t(m), cs, st);
        convert_from_f64<T>(cs, vecKc.p, Kc, st);
        matvec_rows<T>(Qc.p, Kc, Kc, uint32_t(n), vecKc.p, sb.p, st);
        algo_bytes += (uint64_t(Fx) * (m + n) * k * sizeof(T) + (m + n) * sizeof(T)) / uint64_t(comm.nranks);
    }

    // init_pair's init_mat for every stored block, drawn on the device (counter-based, see dense.cu)
    void init_model(uint64_t seed) override {
        // the reference's scale 0.1 * qrsqrt(k), qrsqrt being its fast inverse square root (ffm.cpp:3-12)
        auto qrsqrt = [](double x) {
            const double xhalf = 0.5 * x;
            int64_t i;
            memcpy(&i, &x, sizeof i);
            i = 0x5fe6eb50c7b537a9ll - (i >> 1);
            memcpy(&x, &i, sizeof x);
            x = x * (1.5 - xhalf * x * x);
            return x;
        };
        const double scale = 0.1 * qrsqrt(double(k));
        for (uint32_t f1 = 0; f1 < f; ++f1)
            for (uint32_t f2 = f1; f2 < f; ++f2) {
                Block &bk = blocks[bidx(f1, f2)];
                if (!bk.exists) continue;
                OC_REQUIRE(field_of(f1).set && field_of(f2).set, "set every field before ocffm_init_model");
                const uint64_t rw = field_of(f1).D, rh = field_of(f2).D;
                bk.W.ensure(rw * kp);
                bk.H.ensure(rh * kp);
                init_uniform<T>(bk.W.p, rw, k, kp, seed, 2 * bidx(f1, f2), scale, st);
                init_uniform<T>(bk.H.p, rh, k, kp, seed, 2 * bidx(f1, f2) + 1, scale, st);
                bk.has_w = bk.has_h = true;
            }
        sync();
        state_ready = false;
    }
    void set_hyper(double lambda, double omega, double r) override {
        prm.lambda = lambda;
        prm.omega = omega;
        prm.r = r;
        state_ready = false;   // y-tilde and the caches depend on nothing but the model; rebuild anyway
    }
    void init_state() override {
        OC_REQUIRE(YU.set && YV.set, "labels not set");
        for (auto &F : XU) OC_REQUIRE(F.set, "a user field is missing");
        for (auto &F : XV) OC_REQUIRE(F.set, "an item field is missing");
        for (auto &bk : blocks)
            if (bk.exists) OC_REQUIRE(bk.has_w && bk.has_h, "a parameter block was never set");
        if (comm.active() && comm.peer_ok && peer_gather_allowed && slice_cg) {
            // the CG step at its final size, shared once (collective: every rank is in init_state)
            uint64_t maxD = 0;
            for (auto &F : XU) maxD = std::max(maxD, F.D);
            for (auto &F : XV) maxD = std::max(maxD, F.D);
            S.ensure(maxD * kp);
            if (S_shared != S.p) {
                peer_gather = comm.share(S.p, S_peers.buf, st);
                S_shared = S.p;
            }
        }
        // init_pair, ffm.cpp:346-349
        for (auto &bk : blocks) {
            if (!bk.exists) continue;
            Field &X1 = field_of(bk.f1), &X2 = field_of(bk.f2);
            if (bk.side) {
                bk.P.alloc(X1.rows * kp);
                bk.Q.alloc(X2.rows * kp);
                spmm_rows<T>(X1.view_all(), bk.W.p, bk.P.p, kp, kp, st);
                spmm_rows<T>(X2.view_all(), bk.H.p, bk.Q.p, kp, kp, st);
            } else {
                spmm_rows<T>(X1.view_all(), bk.W.p, Pc.p + size_t(bk.pair) * kp, Kc, kp, st);
                spmm_rows<T>(X2.view_all(), bk.H.p, Qc.p + size_t(bk.pair) * kp, Kc, kp, st);
            }
        }
        cache_sasb();                       // ffm.cpp:508
        a.zero(st);
        b.zero(st);
        if (prm.self_side) {                // calc_side, ffm.cpp:360-373
            for (auto &bk : blocks) {
                if (!bk.exists || !bk.side) continue;
                if (bk.f1 < fu) rowwise_dot<T>(bk.P.p, bk.Q.p, uint32_t(m), kp, a.p, 1, st);
                else rowwise_dot<T>(bk.P.p, bk.Q.p, uint32_t(n), kp, b.p, 1, st);
            }
        }
        // init_y_tilde, ffm.cpp:388-403, both copies
        ytilde_base<T>(YU.view(), a.p, b.p, st);
        ytilde_base<T>(YV.view(), b.p, a.p, st);
        for (uint32_t p = 0; p < Fx; ++p) {
            sddmm_add<T>(YU.view(), Pc.p + size_t(p) * kp, Kc, Qc.p + size_t(p) * kp, Kc, kp, st);
            sddmm_add<T>(YV.view(), Qc.p + size_t(p) * kp, Kc, Pc.p + size_t(p) * kp, Kc, kp, st);
        }
        YU.fresh = YV.fresh = true;
        sync();
        state_ready = true;
    }

    // ------------------------------------------------------------------------------------------
    // one half of a block solve (the reference's (f1, W1, Q1, P1) argument tuples)
    struct Half {
        Block *bk;
        bool side, user;   // user: the rows being swept are users
        Field *X;
        Omega *Yown, *Yoth;
        uint64_t m1, n1, D;
        T *W1, *Q1, *P1;
        uint32_t ldq, ldp;
        T *a1, *b1, *sa1;
        const T *freq;
        // Multi-rank, identity field: features [s0, s1) are touched by this rank's rows only, so
        // G / Hv need no all-reduce and all CG vector work runs on the slice (scalars are reduced)
        bool sliced;
        uint64_t s0, s1;
        size_t soff() const { return size_t(s0); }
        // this rank's share of the half (statistics are per rank: the bench sums them over ranks)
        uint64_t m1l, nnzYl, nnzXl, Dl;
    };
    Half half_of(uint32_t f1, uint32_t f2, int which) {
        OC_REQUIRE(which == 'W' || which == 'H', "which must be 'W' or 'H'");
        OC_REQUIRE(state_ready, "call ocffm_init_state first");
        Block &bk = block(f1, f2);
        Half h;
        h.bk = &bk;
        h.side = bk.side;
        const uint32_t fa = which == 'W' ? f1 : f2;
        h.user = fa < fu;
        h.X = &field_of(fa);
        h.Yown = h.user ? &YU : &YV;
        h.Yoth = h.user ? &YV : &YU;
        h.m1 = h.user ? m : n;
        h.n1 = h.user ? n : m;
        h.D = h.X->D;
        h.a1 = h.user ? a.p : b.p;
        h.b1 = h.user ? b.p : a.p;
        h.sa1 = h.user ? sa.p : sb.p;
        h.freq = prm.freq ? h.X->freq.p : nullptr;
        if (bk.side) {
            h.W1 = which == 'W' ? bk.W.p : bk.H.p;
            h.Q1 = which == 'W' ? bk.Q.p : bk.P.p;
            h.P1 = which == 'W' ? bk.P.p : bk.Q.p;
            h.ldq = h.ldp = kp;
        } else {
            T *pslice = Pc.p + size_t(bk.pair) * kp, *qslice = Qc.p + size_t(bk.pair) * kp;
            h.W1 = which == 'W' ? bk.W.p : bk.H.p;
            h.Q1 = which == 'W' ? qslice : pslice;
            h.P1 = which == 'W' ? pslice : qslice;
            h.ldq = h.ldp = Kc;
        }
        const uint64_t len = h.D * kp;
        G.ensure(len); S.ensure(len); R.ensure(len); V.ensure(len); VQ.ensure(len); Hv.ensure(len);
        h.sliced = comm.active() && h.X->identity && slice_cg;
        h.s0 = h.sliced ? h.X->row0 : 0;
        h.s1 = h.sliced ? h.X->row1 : h.D;
        h.m1l = h.Yown->row1 - h.Yown->row0;
        h.nnzYl = h.Yown->nnz_local;
        h.nnzXl = h.X->nnz_local;
        h.Dl = h.s1 - h.s0;
        return h;
    }

    static uint64_t gather_bytes(uint64_t full, uint64_t rows_gathered, uint64_t row_bytes) {
        return full <= (64ull << 20) ? full : rows_gathered * row_bytes;   // SURVEY.md 8 gather rule
    }

    // Gram stack + oQ + bQ of the companion side for a cross half (ffm.cpp:658-670 without the
    // T accumulation; rows [pair*kp, (pair+1)*kp) of GT are QTQ of ffm.cpp:770)
    void prepare_cross(const Half &h) {
        const T *Qs = h.user ? Qc.p : Pc.p;   // the other side's concatenated embeddings
        gram64.zero(st);
        double *out = gram64.p, *cs = gram64.p + size_t(Kc) * kp, *wsum = cs + kp;
        const uint32_t r0 = uint32_t(h.n1 * comm.rank / comm.nranks),
                       r1 = uint32_t(h.n1 * (comm.rank + 1) / comm.nranks);
        gram_stack<T>(Qs, Kc, Kc, h.Q1, h.ldq, kp, r0, r1, h.b1, out, cs, wsum, 0, st);
        comm.allreduce(gram64.p, gram64.n, st);
        convert_from_f64<T>(out, GT.p, size_t(Kc) * kp, st);
        convert_from_f64<T>(cs, oQ.p, kp, st);
        convert_from_f64<T>(wsum, bQ.p, kp, st);
        algo_bytes += uint64_t(r1 - r0) * k * sizeof(T) * (Fx + 1);
    }
    const T *qtq_of(const Half &h) const { return GT.p + size_t(h.bk->pair) * kp * kp; }

    // scatter part of the gradient into G (zeroed here); lambda W is added by cg_init
    void grad_scatter(const Half &h) {
        const size_t s = sizeof(T);
        OC_CUDA(cudaMemsetAsync(G.p, 0, h.D * kp * sizeof(T), st));
        if (h.X->n_hot) h.X->shadow.zero(st);
        freshen(*h.Yown);
        const uint64_t nnzY = h.nnzYl, nnzX = h.nnzXl;
        if (h.side) {
            OC_CUDA(cudaMemsetAsync(ysum.p, 0, h.m1 * sizeof(T), st));
            ytilde_rowsum<T>(h.Yown->view(), ysum.p, kp, st);
            OC_CUDA(cudaMemsetAsync(&sc->bsum, 0, sizeof(double), st));
            reduce_sum<T>(h.b1, h.n1, 0, &sc->bsum, st);
            side_rows<T>(0, h.Yown->view(), h.X->view(), h.Q1, h.a1, h.sa1, ysum.p, &sc->bsum, nullptr,
                         T(prm.omega), T(prm.r), T(h.n1), G.p, kp, kNoGate, nullptr, st);
            algo_bytes += nnzY * s + h.m1l * (k + 3) * s + nnzX * (4 + s) + 2 * h.Dl * k * s;
        } else {
            prepare_cross(h);
            const T *Ps = h.user ? Pc.p : Qc.p;
            const uint32_t r0 = h.Yown->row0, r1 = h.Yown->row1;
            rowgemm<T>(Ps + size_t(r0) * Kc, Kc, Kc, GT.p, Tm.p + size_t(r0) * kp, r1 - r0, kp, kNoGate, st);
            grad_cross_rows<T>(h.Yown->view(), h.X->view(), h.Q1, h.ldq, Tm.p, h.a1, oQ.p, bQ.p,
                               T(prm.omega), T(prm.r), G.p, kp, st);
            algo_bytes += (h.m1l + 1) * 8 + nnzY * (4 + s) + gather_bytes(h.n1 * k * s, nnzY, k * s) +
                          uint64_t(Fx) * h.m1l * k * s + h.m1l * s + nnzX * (4 + s) + 2 * h.Dl * k * s;
        }
        if (h.X->n_hot) fold_hot<T>(h.X->shadow.p, h.X->hot_feat.p, h.X->n_hot, G.p, kp, st);
        if (!h.sliced) comm.allreduce(G.p, h.D * kp, st);
        nnz_trav += nnzY + nnzX;
    }

    // M_i = sum_{j in Omega_i} q_j q_j^T for the heavy rows of this half (Q1 is fixed during the half
    // solve, ffm.cpp:744-813), so that every CG iteration streams kp x kp numbers per heavy row
    // instead of gathering |Omega_i| rows of Q1
    void build_mrow(const Half &h, bool adaptive = false) {
        mrow_ready = false;
        if (!mrow_on || h.side) return;
        if (adaptive && mrow_mode < 2 && last_iters[half_id(h)] < mrow_iters) return;
        const Omega &Y = *h.Yown;
        if (!Y.n_heavy) return;
        mrow.ensure(size_t(Y.n_heavy) * kp * kp);
        if (Y.n_multi) OC_CUDA(cudaMemsetAsync(mrow.p, 0, size_t(Y.n_multi) * kp * kp * sizeof(T), st));
        row_gram<T>(Y.hw_slot.p, Y.hw_beg.p, Y.hw_cnt.p, Y.n_hw, Y.idx_v(), h.Q1, h.ldq, mrow.p, int(kp), st);
        algo_bytes += Y.nnz_heavy * 4 + gather_bytes(h.n1 * k * sizeof(T), Y.nnz_heavy, k * sizeof(T)) +
                      uint64_t(Y.n_heavy) * k * k * sizeof(T);
        mrow_ready = true;
        ++mrow_builds;
    }

    size_t half_id(const Half &h) const {
        return 2 * size_t(h.bk - blocks.data()) + (h.W1 == h.bk->W.p ? 0 : 1);
    }

    size_t hv_event_pair() {
        if (hv_events_used == hv_events.size()) {
            cudaEvent_t e0, e1;
            OC_CUDA(cudaEventCreate(&e0));
            OC_CUDA(cudaEventCreate(&e1));
            hv_events.emplace_back(e0, e1);
        }
        return hv_events_used++;
    }
    void drain_hv_events() {
        for (size_t i = 0; i < hv_events_used; ++i) {
            float t = 0;
            if (cudaEventElapsedTime(&t, hv_events[i].first, hv_events[i].second) == cudaSuccess) hv_ms += t;
        }
        hv_events_used = 0;
    }

    // Hv (without the regulariser) for the direction in V.  Returns nothing; the caller accounts
    // the statistics with account_hess() once it knows the iteration really ran.
    // fuse_it >= 0 (solver path): the pass also performs the direction update of CG iteration
    // fuse_it on the way (cross halves: inside the V * QTQ row GEMM) and adds V . Hv_data to
    // sc->vHv[fuse_it]; fuse_it < 0: plain Hv of the direction in V.
    void hess_scatter(const Half &h, Gate gate, int fuse_it = -1) {
        if (h.X->n_hot) h.X->shadow.zero(st);
        double *dot_out = fuse_it >= 0 ? sc->vpart[fuse_it] : nullptr;
        if (h.side) {
            side_rows<T>(1, h.Yown->view(), h.X->view(), h.Q1, nullptr, nullptr, nullptr, nullptr, V.p,
                         T(prm.omega), T(prm.r), T(h.n1), Hv.p, kp, gate, dot_out, st);
        } else {
            const size_t o = h.soff() * kp;
            // identity field with unit values: w * X^T X (V QTQ) = w * V QTQ goes into Hv inside the row
            // GEMM, the row pass then skips tau and the rows without pairs (every rank must own its rows'
            // features alone: one rank, or a sliced half)
            const bool notau = fuse_it >= 0 && notau_allowed && h.X->unit_identity && (!comm.active() || h.sliced);
            if (fuse_it >= 0) {
                uint64_t lo = 0, hi = h.s1 - h.s0;
                share_of(h, lo, hi);
                rowgemm_dir<T>(V.p + o, R.p + o, Hv.p + o, h.freq ? h.freq + h.soff() : nullptr, T(prm.lambda),
                               lo, hi, qtq_of(h), VQ.p + o, h.s1 - h.s0, kp, fuse_it, sc, notau ? T(prm.omega) : T(0), st);
            } else {
                rowgemm<T>(V.p + o, kp, kp, qtq_of(h), VQ.p + o, h.s1 - h.s0, kp, gate, st);
            }
            size_t ev = 0;
            if (profile) {
                if (hv_events_used >= 4096) { sync(); drain_hv_events(); }
                ev = hv_event_pair();
                OC_CUDA(cudaEventRecord(hv_events[ev].first, st));
            }
            if (mrow_ready) {
                // heavy rows from their Gram blocks, the light rows by gathers
                hess_heavy_rows<T>(h.Yown->heavy_rows.p, h.Yown->n_heavy, h.X->view(), mrow.p, V.p, VQ.p,
                                   T(prm.omega), Hv.p, int(kp), gate, dot_out, notau, st);
                hess_cross_rows<T>(h.Yown->light_view(notau), h.X->view(), h.Q1, h.ldq, V.p, VQ.p, T(prm.omega),
                                   Hv.p, kp, gate, dot_out, notau, st);
            } else {
                hess_cross_rows<T>(notau ? h.Yown->view_nonempty() : h.Yown->view(), h.X->view(), h.Q1, h.ldq, V.p,
                                   VQ.p, T(prm.omega), Hv.p, kp, gate, dot_out, notau, st);
            }
            if (profile) OC_CUDA(cudaEventRecord(hv_events[ev].second, st));
        }
        if (h.X->n_hot) fold_hot<T>(h.X->shadow.p, h.X->hot_feat.p, h.X->n_hot, Hv.p, kp, st);
        if (!h.sliced) comm.allreduce(Hv.p, h.D * kp, st);
    }
    // rows (relative to the half's slice) whose lambda c_f |V_f|^2 this rank accounts for: all of a
    // slice, or a 1/nranks share of a replicated vector (the partial sums are all-reduced)
    void share_of(const Half &h, uint64_t &lo, uint64_t &hi) const {
        const uint64_t rows = h.s1 - h.s0;
        if (comm.active() && !h.sliced) {
            lo = rows * comm.rank / comm.nranks;
            hi = rows * (comm.rank + 1) / comm.nranks;
        } else {
            lo = 0;
            hi = rows;
        }
    }
    void account_hess(const Half &h, uint64_t iters, bool separate_pass = true) {
        const size_t s = sizeof(T);
        const uint64_t nnzY = h.nnzYl, nnzX = h.nnzXl;
        if (h.side) {
            algo_bytes += iters * (nnzX * (4 + s) + gather_bytes(h.Dl * k * s, nnzX, k * s) + h.m1l * k * s +
                                   h.m1l * 4 + h.Dl * k * s);
            nnz_trav += iters * nnzX;
        } else {
            const uint64_t bytes = (h.m1l + 1) * 8 + nnzX * (4 + s) +
                                   gather_bytes(h.Dl * k * s, nnzX, k * s) + nnzY * 4 +
                                   gather_bytes(h.n1 * k * s, nnzY, k * s) + h.Dl * k * s;
            algo_bytes += iters * bytes;
            if (separate_pass) {   // (the persistent CG kernel has its own counters)
                hv_algo_bytes += iters * bytes;
                hv_launches += iters;
            }
            nnz_trav += iters * (nnzY + nnzX);
        }
        algo_bytes += iters * 7 * h.Dl * k * s;
    }

    // cg, ffm.cpp:744-813.  G holds the gradient WITHOUT lambda W when add_reg is set (solver
    // path, the regulariser is fused into cg_init), or the full gradient otherwise.
    // Iteration it+1 is enqueued before r2[it+1] has been read back; its kernels carry a device
    // side gate that re-evaluates the reference's stop test (g2 * 0.09 < r2, ffm.cpp:780), so a
    // speculative iteration past the stop is a no-op and the GPU never idles on the host.
    void enqueue_cg_iter(const Half &h, int it) {
        const size_t o = h.soff() * kp;
        const uint64_t Ds = h.s1 - h.s0, len = Ds * kp;
        const T *fq = h.freq ? h.freq + h.soff() : nullptr;
        T reg = T(prm.lambda);   // regulariser still to be added to Hv by cg_step
        int slotted = 0;         // V.Hv arrives as sc->vpart[it] partial sums
        const bool partial = h.sliced;  // vHv / r2 are sums over this rank's part only
        if (h.side && h.X->diagonal && diag_fast && (!comm.active() || h.sliced)) {
            // row-local Hessian: direction update, Hv, regulariser and V.Hv in one pass
            side_diag_iter<T>(h.Yown->view(), h.X->view(), h.Q1, V.p, R.p, Hv.p, h.freq, T(prm.lambda),
                              T(prm.omega), T(h.n1), kp, it, sc, st);
            reg = T(0);
        } else if (fused_dot) {
            // 3 kernels per cross iteration: [direction + V QTQ], [Hessian rows + V.Hv], [step]
            if (h.side) {
                uint64_t lo, hi;
                share_of(h, lo, hi);
                cg_dir<T>(V.p + o, R.p + o, Hv.p + o, len, it, sc, fq, T(prm.lambda), kp, lo, hi, 1, st);
            }
            hess_scatter(h, Gate{sc, it}, it);
            slotted = 1;
            comm.allreduce(sc->vpart[it], size_t(kDotSlots), st);   // every rank holds a share
        } else {
            cg_dir<T>(V.p + o, R.p + o, Hv.p + o, len, it, sc, fq, T(0), kp, 0, 0, 0, st);
            hess_scatter(h, Gate{sc, it});
            cg_reg_dot<T>(Hv.p + o, V.p + o, fq, T(prm.lambda), Ds, kp, it, sc, 1, st);
            reg = T(0);
        }
        if (partial && !slotted) comm.allreduce(&sc->vHv[it], 1, st);   // slice partial -> global V.Hv
        cg_step<T>(S.p + o, R.p + o, V.p + o, Hv.p + o, len, it, sc, fq, reg, kp, slotted,
                   partial ? nullptr : d_hscal + 1 + it, st);
        if (partial) {   // the slice's share is summed over the ranks first: the host copy comes by memcpy
            comm.allreduce(&sc->r2[it + 1], 1, st);
            OC_CUDA(cudaMemcpyAsync(h_scal + 1 + it, &sc->r2[it + 1], sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        OC_CUDA(cudaEventRecord(cg_ev[it], st));
    }
    // ---- persistent CG (cg_side_persist): iteration counts are read back at the next sync ----------
    bool persist_on = true;            // OCFFM_PERSIST_CG=0: per-iteration kernels + host stop test everywhere
    struct PendingCg { Half h; int slot; int ev; };
    // OCFFM_PROFILE: events around every persistent CG kernel of a cross half
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> cgk_events;
    size_t cgk_events_used = 0;
    double cgk_ms = 0;
    uint64_t cgk_bytes = 0, cgk_launches = 0, cgk_iters = 0;
    std::vector<PendingCg> pending;
    unsigned *h_iters = nullptr;       // pinned
    static constexpr int kIterSlots = 1024;
    int last_drained_iters = 0;
    void drain_pending() {             // the stream is idle: every slot has landed
        for (const PendingCg &pc : pending) {
            const int it = int(h_iters[pc.slot]);
            const uint64_t before = algo_bytes;
            // profiled cross solves: the row phases of the persistent kernel count as Hessian passes (their
            // time comes from the kernel's own %globaltimer stamps between the grid barriers)
            const bool timed_rows = pc.ev >= 0;
            account_hess(pc.h, uint64_t(it), timed_rows);
            if (timed_rows) hv_ms += double(h_phase[pc.slot]);
            if (pc.ev >= 0) {   // a profiled cross solve: the fused kernel's time and ALL its algorithmic bytes
                float t = 0;
                if (cudaEventElapsedTime(&t, cgk_events[pc.ev].first, cgk_events[pc.ev].second) == cudaSuccess) cgk_ms += t;
                cgk_bytes += algo_bytes - before;
                cgk_launches += 1;
                cgk_iters += uint64_t(it);
            }
            cg_iters += uint64_t(it);
            last_iters[half_id(pc.h)] = it;
            last_drained_iters = it;
        }
        pending.clear();
        cgk_events_used = 0;
    }
    // OCFFM_PERSIST_CG: 0 off, 1 same-side halves only, 2 (default) cross halves too when kp >= 32, 3 cross
    // halves at every supported kp.  Measured on the Outbrain shape (k = 16, 4-lane groups): the persistent
    // cross kernel costs 3 ms and the per-row Gram 10 ms per outer iteration there, so both wait for kp >= 32.
    int persist_mode = 2;
    bool persist_eligible(const Half &h) const {
        if (!persist_on || h.X->n_hot) return false;
        // several ranks: only halves whose CG vectors are per-rank slices (identity fields); the two scalars
        // of an iteration are summed over the ranks inside the kernel (PeerK)
        if (comm.active() && !(h.sliced && comm.peerk_ok)) return false;
        if (h.side) return true;
        return persist_mode >= 2 && (kp >= 32 || persist_mode >= 3) && cg_cross_persist_supported(int(kp), sizeof(T));
    }
    void run_cg_persist(const Half &h, bool add_reg) {
        const size_t o = h.soff() * kp;
        const uint64_t Ds = h.s1 - h.s0;
        OC_CUDA(cudaMemsetAsync(sc, 0, sizeof(SolveScalars), st));
        cg_init<T>(G.p + o, h.W1 + o, h.freq ? h.freq + h.soff() : nullptr, add_reg ? T(prm.lambda) : T(0), R.p + o,
                   V.p + o, S.p + o, Ds, kp, sc, nullptr, st);
        if (h.sliced) comm.allreduce(&sc->r2[0], 1, st);
        PeerK pk = comm.pk;
        if (!h.sliced) pk.nranks = 1;
        if (!h_iters) {
            OC_CUDA(cudaHostAlloc(&h_iters, kIterSlots * sizeof(unsigned), cudaHostAllocMapped));
            OC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_hiters), h_iters, 0));
            OC_CUDA(cudaHostAlloc(&h_phase, kIterSlots * sizeof(float), cudaHostAllocMapped));
            OC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_hphase), h_phase, 0));
        }
        if (int(pending.size()) >= kIterSlots) sync();   // (never inside one outer iteration: 2 x blocks << 1024)
        const int slot = int(pending.size());
        int ev = -1;
        if (profile && !h.side) {
            if (cgk_events_used == cgk_events.size()) {
                cudaEvent_t e0, e1;
                OC_CUDA(cudaEventCreate(&e0));
                OC_CUDA(cudaEventCreate(&e1));
                cgk_events.emplace_back(e0, e1);
            }
            ev = int(cgk_events_used++);
            OC_CUDA(cudaEventRecord(cgk_events[ev].first, st));
        }
        if (h.side)
            cg_side_persist<T>(h.Yown->view(), h.X->view(), h.Q1, V.p, R.p, S.p, Hv.p, h.freq, T(prm.lambda),
                               T(prm.omega), T(h.n1), Ds, int(kp), h.X->diagonal && (diag_fast || h.sliced), sc, 20, 9e-2,
                               d_hiters + slot, h.s0, pk, st);
        else
            cg_cross_persist<T>(mrow_ready ? h.Yown->light_view() : h.Yown->view(), h.X->view(), h.Q1, h.ldq, qtq_of(h),
                                V.p, R.p, S.p, Hv.p, VQ.p, h.freq, T(prm.lambda), T(prm.omega), Ds, int(kp), sc, 20,
                                9e-2, d_hiters + slot, mrow_ready ? h.Yown->heavy_rows.p : nullptr,
                                mrow_ready ? h.Yown->n_heavy : 0u, mrow_ready ? mrow.p : nullptr, h.s0, pk,
                                profile ? d_hphase + slot : nullptr, st);
        if (ev >= 0) OC_CUDA(cudaEventRecord(cgk_events[ev].second, st));
        pending.push_back(PendingCg{h, slot, ev});
    }

    int run_cg(const Half &h, bool add_reg) {
        if (!h.side) build_mrow(h, true);   // decides mrow_ready before the path is chosen
        if (persist_eligible(h)) {
            run_cg_persist(h, add_reg);
            return -1;   // the count arrives with the next sync (drain_pending)
        }
        const size_t o = h.soff() * kp;
        const T *fq = h.freq ? h.freq + h.soff() : nullptr;
        if (h.side) mrow_ready = false;
        OC_CUDA(cudaMemsetAsync(sc, 0, sizeof(SolveScalars), st));
        cg_init<T>(G.p + o, h.W1 + o, fq, add_reg ? T(prm.lambda) : T(0), R.p + o, V.p + o, S.p + o, h.s1 - h.s0,
                   kp, sc, h.sliced ? nullptr : d_hscal, st);
        if (h.sliced) comm.allreduce(&sc->r2[0], 1, st);
        // No hard stream sync in here: g2 is fetched asynchronously while iteration 0 (which the
        // device gate closes by itself when g2 == 0) is already enqueued, and after the loop the
        // stream order alone protects S / R / V -- the GPU never waits for the host.
        if (h.sliced) OC_CUDA(cudaMemcpyAsync(h_scal, &sc->r2[0], sizeof(double), cudaMemcpyDeviceToHost, st));
        OC_CUDA(cudaEventRecord(g2_ev, st));
        const int max_cg = 20;
        const double eps = 9e-2;
        int it = 0;
        enqueue_cg_iter(h, 0);                                     // speculative
        OC_CUDA(cudaEventSynchronize(g2_ev));
        const double g2 = h_scal[0];
        if (g2 * eps < g2) {
            for (;;) {
                if (it + 1 < max_cg) enqueue_cg_iter(h, it + 1);   // speculative
                OC_CUDA(cudaEventSynchronize(cg_ev[it]));
                const double r2 = h_scal[1 + it];
                ++it;
                if (!(g2 * eps < r2) || it >= max_cg) break;
            }
        }
        account_hess(h, uint64_t(it));
        cg_iters += uint64_t(it);
        last_iters[half_id(h)] = it;
        return it;
    }

    // update_side / update_cross, ffm.cpp:405-465
    void apply_update(const Half &h) {
        const size_t s = sizeof(T);
        const uint64_t len = h.D * kp, nnzY = h.nnzYl, nnzX = h.nnzXl;
        // sliced half: one all-gather of the step per half solve, then every rank applies the whole
        // step to its replicas (W, P, a) with the single-rank kernels
        if (h.sliced) {
            if (peer_gather) comm.template peer_gather_rows<T>(S_peers, h.D, kp, st);
            else comm.allgather_rows(S.p, h.D, kp, st);
        }
        axpy<T>(h.W1, S.p, T(1), len, st);
        const T *q_side = h.side ? h.Q1 : nullptr;
        if (!comm.active() || h.sliced) {
            spmm_update<T>(h.X->view_all(), S.p, XS.p, h.P1, h.ldp, q_side, gap.p, h.a1, kp, st);
        } else {
            spmm_rows<T>(h.X->view(), S.p, XS.p, kp, kp, st);
            comm.allgather_rows(XS.p, h.m1, kp, st);
            // P1 += XS (and the side terms) on every rank's replica
            spmm_update_from_xs(h, q_side);
        }
        freshen(*h.Yown);
        const uint64_t copies = mirror_yt ? 1 : 2;
        if (mirror_yt) h.Yoth->fresh = false;
        if (h.side) {
            ytilde_add_gap<T>(h.Yown->view(), gap.p, 1, st);
            if (!mirror_yt) ytilde_add_gap<T>(h.Yoth->view(), gap.p, 0, st);
            algo_bytes += 3 * h.Dl * k * s + nnzX * (4 + s) + gather_bytes(h.Dl * k * s, nnzX, k * s) +
                          4 * h.m1l * k * s + copies * (nnzY * (4 + 2 * s) + h.m1l * s);
        } else {
            sddmm_add<T>(h.Yown->view(), XS.p, kp, h.Q1, h.ldq, kp, st);
            if (!mirror_yt) sddmm_add<T>(h.Yoth->view(), h.Q1, h.ldq, XS.p, kp, kp, st);
            algo_bytes += 3 * h.Dl * k * s + nnzX * (4 + s) + gather_bytes(h.Dl * k * s, nnzX, k * s) +
                          4 * h.m1l * k * s +
                          copies * (nnzY * (4 + 2 * s)) + gather_bytes(h.n1 * k * s, nnzY, k * s) +
                          (copies - 1) * gather_bytes(h.m1l * k * s, nnzY, k * s);
        }
        // the reference's traversal count (both copies, ffm.cpp:423-436, 451-464)
        nnz_trav += 2 * nnzY + nnzX;
    }
    // multi-rank variant of the tail of spmm_update: XS is complete on every rank
    void spmm_update_from_xs(const Half &h, const T *q_side) {
        // P1[row, :] += XS[row, :] for all rows: expressed as spmm_update over an identity pattern
        // would cost an index array; use axpy for contiguous side blocks and a strided add otherwise
        if (h.ldp == kp) {
            axpy<T>(h.P1, XS.p, T(1), h.m1 * kp, st);
        } else {
            strided_add(h.P1, h.ldp, XS.p, h.m1);
        }
        if (q_side) {
            rowwise_dot<T>(XS.p, q_side, uint32_t(h.m1), kp, gap.p, 0, st);
            axpy_scalar_vec(h.a1, gap.p, h.m1);
        }
    }
    void strided_add(T *dst, uint32_t ld, const T *src, uint64_t rows);
    void axpy_scalar_vec(T *y, const T *x, uint64_t n_);

    // phase timers (OCFFM_PROFILE >= 2): events at the phase boundaries of every half solve
    std::vector<cudaEvent_t> ph_events;
    std::vector<int> ph_kind;   // 0 = side, 1 = cross
    int profile_level = 0;
    void phase_mark(int kind) {
        if (profile_level < 2) return;
        cudaEvent_t e;
        OC_CUDA(cudaEventCreate(&e));
        OC_CUDA(cudaEventRecord(e, st));
        ph_events.push_back(e);
        ph_kind.push_back(kind);
    }
    void drain_phases() {
        // events come in groups of 4 per half solve: start, after grad, after cg, after update
        for (size_t i = 0; i + 3 < ph_events.size(); i += 4) {
            float g = 0, c = 0, u = 0;
            cudaEventElapsedTime(&g, ph_events[i], ph_events[i + 1]);
            cudaEventElapsedTime(&c, ph_events[i + 1], ph_events[i + 2]);
            cudaEventElapsedTime(&u, ph_events[i + 2], ph_events[i + 3]);
            const int o = ph_kind[i] ? 3 : 0;
            ms[o + 0] += g; ms[o + 1] += c; ms[o + 2] += u;
        }
        for (auto e : ph_events) cudaEventDestroy(e);
        ph_events.clear();
        ph_kind.clear();
    }
    void solve_half(uint32_t f1, uint32_t f2, int which) {
        Half h = half_of(f1, f2, which);
        const int kind = h.side ? 0 : 1;
        phase_mark(kind);
        grad_scatter(h);
        phase_mark(kind);
        run_cg(h, true);
        phase_mark(kind);
        apply_update(h);
        phase_mark(kind);
    }
    void solve_block(uint32_t f1, uint32_t f2) override {
        solve_half(f1, f2, 'W');   // W first, then H against the updated P1 (ffm.cpp:826-832, 843-849)
        solve_half(f1, f2, 'H');
    }
    void one_epoch() override {
        OC_REQUIRE(state_ready, "call ocffm_init_state first");
        if (mirrors) {   // the staging buffer must not be re-allocated while the copy stream uses it
            uint64_t mx = 0;
            for (auto &bk : blocks)
                if (bk.exists) mx = std::max(mx, std::max(block_rows(bk, 'W'), block_rows(bk, 'H')));
            for (auto &b : mir_stage) b.ensure(mx * k);
        }
        auto solve = [&](uint32_t f1, uint32_t f2) {
            solve_block(f1, f2);
            if (mirrors) stream_out(blocks[bidx(f1, f2)]);
        };
        if (prm.self_side) {
            for (uint32_t f1 = 0; f1 < fu; ++f1)
                for (uint32_t f2 = f1; f2 < fu; ++f2) solve(f1, f2);
            for (uint32_t f1 = fu; f1 < f; ++f1)
                for (uint32_t f2 = f1; f2 < f; ++f2) solve(f1, f2);
        }
        for (uint32_t f1 = 0; f1 < fu; ++f1)
            for (uint32_t f2 = fu; f2 < f; ++f2) solve(f1, f2);
        if (prm.self_side) cache_sasb();
        sync();
        if (mirrors) OC_CUDA(cudaStreamSynchronize(copy_st));   // every registered mirror is current on return
        if (profile) { drain_hv_events(); drain_phases(); }
    }

    // ---- observation entry points (parity tests) -------------------------------------------------
    void add_reg_into(T *dst, const T *src, const Half &h) {
        // dst += lambda (freq) src, via cg_reg_dot's first half would also touch scalars; reuse cg_init
        // on a scratch is overkill -> small dedicated path through axpy when no freq
        if (!h.freq) {
            axpy<T>(dst, src, T(prm.lambda), h.D * kp, st);
        } else {
            OC_CUDA(cudaMemsetAsync(sc, 0, sizeof(SolveScalars), st));
            cg_reg_dot<T>(dst, src, h.freq, T(prm.lambda), h.D, kp, 0, sc, 0, st);
        }
    }
    void grad(uint32_t f1, uint32_t f2, int which, double *Gout, uint64_t rows) override {
        Half h = half_of(f1, f2, which);
        OC_REQUIRE(rows == h.D, "rows must equal Ds of the updated field");
        grad_scatter(h);
        if (h.sliced) comm.allgather_rows(G.p, h.D, kp, st);
        add_reg_into(G.p, h.W1, h);
        download_unpadded(G.p, kp, Gout, h.D);
    }
    void hess_vec(uint32_t f1, uint32_t f2, int which, const double *Vin, double *Hout,
                  uint64_t rows) override {
        Half h = half_of(f1, f2, which);
        OC_REQUIRE(rows == h.D, "rows must equal Ds of the updated field");
        upload_padded(V, Vin, h.D);
        if (!h.side) prepare_cross(h);
        build_mrow(h);
        OC_CUDA(cudaMemsetAsync(Hv.p, 0, h.D * kp * sizeof(T), st));
        hess_scatter(h, kNoGate);
        if (h.sliced) comm.allgather_rows(Hv.p, h.D, kp, st);
        add_reg_into(Hv.p, V.p, h);
        download_unpadded(Hv.p, kp, Hout, h.D);
    }
    void cg(uint32_t f1, uint32_t f2, int which, const double *Gin, double *Sout, uint64_t rows,
            int32_t *iters) override {
        Half h = half_of(f1, f2, which);
        OC_REQUIRE(rows == h.D, "rows must equal Ds of the updated field");
        upload_padded(G, Gin, h.D);
        if (!h.side) prepare_cross(h);
        const uint64_t before = cg_iters;
        int it = run_cg(h, false);
        if (it < 0) { sync(); it = last_drained_iters; }
        if (h.sliced) comm.allgather_rows(S.p, h.D, kp, st);
        cg_iters = before;
        if (iters) *iters = it;
        download_unpadded(S.p, kp, Sout, h.D);
    }

    // ------------------------------------------------------------------------------------------
    // func(), ffm.cpp:1321-1351, through Gram identities (see DESIGN.md "objective")
    void objective(double *value) override {
        OC_REQUIRE(state_ready, "call ocffm_init_state first");
        const size_t gsz = size_t(Kc) * kp;
        std::vector<double> PtP(size_t(Kc) * Kc), QtQ(size_t(Kc) * Kc), sP(Kc), sQ(Kc), Pa(Kc), Qb(Kc);
        std::vector<double> tmp(gsz + 2 * kp);
        auto side_grams = [&](const T *M, uint64_t rows, const T *wv, std::vector<double> &MtM,
                              std::vector<double> &sM, std::vector<double> &Mw) {
            for (uint32_t p = 0; p < Fx; ++p) {
                gram64.zero(st);
                gram_stack<T>(M, Kc, Kc, M + size_t(p) * kp, Kc, kp, 0, uint32_t(rows), wv, gram64.p,
                              gram64.p + gsz, gram64.p + gsz + kp, 1, st);
                gram64.download(tmp.data(), gsz + 2 * kp, st);
                sync();
                for (uint32_t c = 0; c < Kc; ++c)
                    for (uint32_t d = 0; d < kp; ++d) MtM[size_t(c) * Kc + p * kp + d] = tmp[size_t(c) * kp + d];
                for (uint32_t d = 0; d < kp; ++d) {
                    sM[p * kp + d] = tmp[gsz + d];
                    Mw[p * kp + d] = tmp[gsz + kp + d];
                }
            }
        };
        side_grams(Pc.p, m, a.p, PtP, sP, Pa);
        side_grams(Qc.p, n, b.p, QtQ, sQ, Qb);
        acc64.zero(st);
        reduce_sum<T>(a.p, m, 0, acc64.p + 0, st);
        reduce_sum<T>(a.p, m, 1, acc64.p + 1, st);
        reduce_sum<T>(b.p, n, 0, acc64.p + 2, st);
        reduce_sum<T>(b.p, n, 1, acc64.p + 3, st);
        // Omega part over the local rows of the user orientation
        {
            freshen(YU);
            const uint64_t b0 = YU.h_rowptr[YU.row0], e0 = YU.h_rowptr[YU.row1];
            omega_objective<T>(YU.yt_v() + b0, e0 - b0, T(prm.omega), T(prm.r), acc64.p + 4, st);
        }
        comm.allreduce(acc64.p + 4, 1, st);
        for (auto &bk : blocks) {
            if (!bk.exists) continue;
            reduce_sum<T>(bk.W.p, bk.W.n, 1, acc64.p + 5, st);
            reduce_sum<T>(bk.H.p, bk.H.n, 1, acc64.p + 5, st);
        }
        double h[8];
        acc64.download(h, 8, st);
        sync();
        const double r = prm.r, w = prm.omega;
        const double sa_ = h[0], saa = h[1], sb_ = h[2], sbb = h[3];
        const double sc_ = sa_ - double(m) * r, scc = saa - 2 * r * sa_ + double(m) * r * r;
        double fro = 0, cPsQ = 0, bQsP = 0;
        for (size_t i = 0; i < PtP.size(); ++i) fro += PtP[i] * QtQ[i];
        for (uint32_t c = 0; c < Kc; ++c) {
            cPsQ += (Pa[c] - r * sP[c]) * sQ[c];
            bQsP += Qb[c] * sP[c];
        }
        const double all = double(n) * scc + double(m) * sbb + 2 * sc_ * sb_ + fro + 2 * cPsQ + 2 * bQsP;
        *value = 0.5 * (h[4] + w * all + prm.lambda * h[5]);
    }

    // ------------------------------------------------------------------------------------------
    // validate(), ffm.cpp:925-1016
    void validate(double *prec, double *ndcg, double *ploss, uint32_t *topk) override {
        OC_REQUIRE(test_set, "test labels not set");
        for (auto &bk : blocks)
            if (bk.exists) OC_REQUIRE(bk.has_w && bk.has_h, "a parameter block was never set");
        // evaluation buffers persist across calls (no cudaMalloc/cudaFree on the hot path)
        DevBuf<T> &Pva = ev_Pva, &Qva = ev_Qva, &at = ev_at, &bt = ev_bt, &t1 = ev_t1, &t2 = ev_t2;
        Pva.ensure(mt * Kc);
        Qva.ensure(n * Kc);
        at.ensure(mt); bt.ensure(n);
        OC_CUDA(cudaMemsetAsync(at.p, 0, mt * sizeof(T), st));
        OC_CUDA(cudaMemsetAsync(bt.p, 0, n * sizeof(T), st));
        t1.ensure(std::max(mt, n) * kp);
        t2.ensure(std::max(mt, n) * kp);
        for (auto &bk : blocks) {           // ffm.cpp:932-963
            if (!bk.exists) continue;
            if (!bk.side) {
                spmm_rows<T>(XT[bk.f1].view_all(), bk.W.p, Pva.p + size_t(bk.pair) * kp, Kc, kp, st);
                spmm_rows<T>(XV[bk.f2 - fu].view_all(), bk.H.p, Qva.p + size_t(bk.pair) * kp, Kc, kp, st);
            } else if (bk.f1 < fu) {
                spmm_rows<T>(XT[bk.f1].view_all(), bk.W.p, t1.p, kp, kp, st);
                spmm_rows<T>(XT[bk.f2].view_all(), bk.H.p, t2.p, kp, kp, st);
                rowwise_dot<T>(t1.p, t2.p, uint32_t(mt), kp, at.p, 1, st);
            } else {
                spmm_rows<T>(XV[bk.f1 - fu].view_all(), bk.W.p, t1.p, kp, kp, st);
                spmm_rows<T>(XV[bk.f2 - fu].view_all(), bk.H.p, t2.p, kp, kp, st);
                rowwise_dot<T>(t1.p, t2.p, uint32_t(n), kp, bt.p, 1, st);
            }
        }
        if (!cold_ready) {   // the popularity ranking never changes after set_labels
            vector_topk<T>(popular.p, uint32_t(n_ranked), cold_ids.p, st);
            cold_ready = true;
        }
        bool used_tc = false;
        if constexpr (std::is_same<T, float>::value) {
            if (eval_tc && score_topk_tc_supported(Kc)) {
                // tcgen05 path: 3xTF32 operands (hi, lo) of both sides, then the fused scorer
                const uint32_t nsplit = score_topk_tc_splits(t_row1 - t_row0, uint32_t(n_ranked));
                tc_phi.ensure(mt * Kc); tc_plo.ensure(mt * Kc);
                tc_qhi.ensure(n * Kc);  tc_qlo.ensure(n * Kc);
                split_tf32(Pva.p, tc_phi.p, tc_plo.p, mt * Kc, st);
                split_tf32(Qva.p, tc_qhi.p, tc_qlo.p, n * Kc, st);
                const size_t slots = size_t(mt) * nsplit * score_topk_tc_cand_slots();
                tc_cand_score.ensure(slots);
                tc_cand_id.ensure(slots);
                tc_row_thr.ensure(mt);
                part_score.ensure(size_t(mt) * nsplit * 80);
                part_id.ensure(size_t(mt) * nsplit * 80);
                score_topk_tc(tc_phi.p, tc_plo.p, mt, tc_qhi.p, tc_qlo.p, n, Kc, bt.p, t_row0, t_row1,
                              uint32_t(n_ranked), t_cold.p, nsplit, tc_cand_score.p, tc_cand_id.p,
                              part_score.p, part_id.p, tc_row_thr.p, st);
                merge_topk<float>(part_score.p, part_id.p, nsplit, t_row0, t_row1, topk_ids.p, st);
                used_tc = true;
            }
        }
        if (!used_tc) {
            const uint32_t nsplit = score_topk_splits(t_row1 - t_row0, uint32_t(n_ranked));
            part_score.ensure(size_t(mt) * nsplit * 80);
            part_id.ensure(size_t(mt) * nsplit * 80);
            score_topk<T>(Pva.p, Qva.p, Kc, bt.p, t_row0, t_row1, uint32_t(n_ranked), t_cold.p, nsplit,
                          part_score.p, part_id.p, topk_ids.p, st);
        }
        acc64.zero(st);
        eval_metrics<T>(topk_ids.p, cold_ids.p, t_cold.p, t_rowptr.p, t_idx.p, t_row0, t_row1, Pva.p,
                        Qva.p, Kc, at.p, bt.p, popular.p, uint32_t(n), uint32_t(n_ranked), acc64.p, st);
        comm.allreduce(acc64.p, 16, st);
        double h[16];
        acc64.download(h, 16, st);
        sync();
        const int cut[5] = {5, 10, 20, 40, 80};
        for (int s = 0; s < 5; ++s) {
            prec[s] = h[s] / (double(mt) * cut[s]);      // ffm.cpp:1013
            ndcg[s] = h[5 + s] / double(mt);             // ffm.cpp:1014
        }
        *ploss = std::sqrt(h[10] / double(mt));          // ffm.cpp:1002
        if (topk) {
            OC_REQUIRE(!comm.active(), "top-k ids are only returned by single-rank contexts");
            topk_ids.download(topk, mt * 80, st);
            uint32_t cold80[80];
            cold_ids.download(cold80, 80, st);
            sync();
            for (uint64_t i = 0; i < mt; ++i)
                if (h_cold[i]) memcpy(topk + i * 80, cold80, sizeof(cold80));
        }
    }

    // ------------------------------------------------------------------------------------------
    void get_vec(const char *name, double *out, uint64_t *count) override {
        const std::string s(name);
        const T *src = nullptr;
        uint64_t cnt = 0;
        if (s == "a") { src = a.p; cnt = m; }
        else if (s == "b") { src = b.p; cnt = n; }
        else if (s == "sa") { src = sa.p; cnt = m; }
        else if (s == "sb") { src = sb.p; cnt = n; }
        else if (s == "ytilde_csr" || s == "ytilde_csc") {
            // a rank holds the cache of its own rows only: the other entries are returned as zeros
            Omega &Y = s == "ytilde_csr" ? YU : YV;
            freshen(Y);
            if (count) *count = Y.nnz;
            if (!out) return;
            std::vector<T> hloc(Y.nnz_local);
            Y.yt.download(hloc.data(), Y.nnz_local, st);
            sync();
            std::fill(out, out + Y.nnz, 0.0);
            for (uint64_t i = 0; i < Y.nnz_local; ++i) out[Y.base + i] = double(hloc[i]);
            return;
        }
        else if (s == "popular") { src = popular.p; cnt = n_ranked; }
        else throw Error(OCFFM_E_INVALID, "unknown vector name " + s);
        if (count) *count = cnt;
        if (!out) return;
        std::vector<T> h(cnt);
        OC_CUDA(cudaMemcpyAsync(h.data(), src, cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
        sync();
        for (uint64_t i = 0; i < cnt; ++i) out[i] = double(h[i]);
    }
    void get_embed(uint32_t f1, uint32_t f2, int which, double *out, uint64_t rows) override {
        OC_REQUIRE(which == 'P' || which == 'Q', "which must be 'P' or 'Q'");
        OC_REQUIRE(state_ready, "call ocffm_init_state first");
        Block &bk = block(f1, f2);
        const uint64_t want = which == 'P' ? rows_of(f1) : rows_of(f2);
        OC_REQUIRE(rows == want, "rows must equal the row count of the embedding's side");
        if (bk.side) download_unpadded(which == 'P' ? bk.P.p : bk.Q.p, kp, out, rows);
        else download_unpadded((which == 'P' ? Pc.p : Qc.p) + size_t(bk.pair) * kp, Kc, out, rows);
    }
    void get_csc(uint64_t *colptr, uint32_t *rowidx) override {
        OC_REQUIRE(YV.set, "labels not set");
        // the device holds this rank's rows only; the full CSC is kept on the host
        std::copy(YV.h_rowptr.begin(), YV.h_rowptr.end(), colptr);
        std::copy(YV.h_idx.begin(), YV.h_idx.end(), rowidx);
    }
    void get_stats(ocffm_stats *out) override {
        drain_hv_events();
        memset(out, 0, sizeof(*out));
        out->kernel_launches = launches;
        out->cg_iters = cg_iters;
        out->nnz_traversed = nnz_trav;
        out->algo_bytes = algo_bytes;
        out->hv_launches = hv_launches;
        out->hv_algo_bytes = hv_algo_bytes;
        out->hv_ms = hv_ms;
        auto omega_bytes = [](const Omega &Y) {
            return (Y.rowptr.n + Y.idx.n + Y.wi_row.n + Y.wi_beg.n + Y.wi_cnt.n + Y.mirror_pos.n + Y.heavy_rows.n +
                    Y.hw_slot.n + Y.hw_beg.n + Y.hw_cnt.n + Y.lw_row.n + Y.lw_beg.n + Y.lw_cnt.n) * sizeof(uint32_t) +
                   Y.yt.n * sizeof(T);
        };
        out->omega_device_bytes = omega_bytes(YU) + omega_bytes(YV);
        out->row_gram_bytes = mrow.n * sizeof(T);
        out->row_gram_builds = mrow_builds;
        out->cg_kernel_ms = cgk_ms;
        out->cg_kernel_algo_bytes = cgk_bytes;
        out->cg_kernel_launches = cgk_launches;
        out->cg_kernel_iters = cgk_iters;
        // side: grad / cg / update ; cross: grad / cg / update  (OCFFM_PROFILE >= 2)
        out->ms_side_grad = ms[0]; out->ms_side_cg = ms[1]; out->ms_side_update = ms[2];
        out->ms_cross_grad = ms[3]; out->ms_cross_cg = ms[4]; out->ms_cross_update = ms[5];
    }
    void reset_stats() override {
        sync();
        drain_hv_events();
        launches = cg_iters = nnz_trav = algo_bytes = hv_launches = hv_algo_bytes = mrow_builds = 0;
        cgk_ms = 0;
        cgk_bytes = cgk_launches = cgk_iters = 0;
        hv_ms = 0;
        drain_phases();
        for (double &v : ms) v = 0;
    }
    void synchronize() override { sync(); }
    void *stream() override { return st; }
};

// small element-wise helpers used only by the multi-rank update path
template <typename T>
__global__ void k_strided_add(T *dst, uint32_t ld, const T *src, uint64_t rows, uint32_t kp) {
    pdl_enter();
    const uint64_t nvec = rows * (kp / 4);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t row = i / (kp / 4), c = (i % (kp / 4)) * 4;
        V4<T> d = ld4(dst + row * ld + c);
        const V4<T> s = ld4(src + row * kp + c);
        d.x += s.x; d.y += s.y; d.z += s.z; d.w += s.w;
        st4(dst + row * ld + c, d);
    }
}
template <typename T>
__global__ void k_vec_add(T *y, const T *x, uint64_t n) {
    pdl_enter();
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x)
        y[i] += x[i];
}
template <typename T>
void Problem<T>::strided_add(T *dst, uint32_t ld, const T *src, uint64_t rows) {
    if (!rows) return;
    const uint64_t nvec = rows * (kp / 4);
    OC_LAUNCH((k_strided_add<T>), unsigned(std::min<uint64_t>((nvec + 255) / 256, kSMs * 8)), 256, 0, st,
              dst, ld, src, rows, kp);
}
template <typename T>
void Problem<T>::axpy_scalar_vec(T *y, const T *x, uint64_t n_) {
    if (!n_) return;
    OC_LAUNCH((k_vec_add<T>), unsigned(std::min<uint64_t>((n_ + 255) / 256, kSMs * 8)), 256, 0, st, y, x, n_);
}

}  // namespace ocffm

// ===============================================================================================
// C ABI
// ===============================================================================================
struct ocffm_ctx {
    std::unique_ptr<ocffm::CtxBase> impl;
};

namespace {
template <typename F>
int guarded(F &&fn) {
    try {
        fn();
        return OCFFM_OK;
    } catch (const ocffm::Error &e) {
        ocffm::g_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        ocffm::g_last_error = "host allocation failed";
        return OCFFM_E_NOMEM;
    } catch (const std::exception &e) {
        ocffm::g_last_error = e.what();
        return OCFFM_E_INVALID;
    } catch (...) {
        ocffm::g_last_error = "unknown error";
        return OCFFM_E_INVALID;
    }
}
template <typename T>
ocffm::Problem<T> *as(ocffm_ctx *c) { return static_cast<ocffm::Problem<T> *>(c->impl.get()); }
}  // namespace

#define OC_CTX(ctx)                                                                    \
    if (!(ctx) || !(ctx)->impl) {                                                      \
        ocffm::g_last_error = "null context";                                          \
        return OCFFM_E_INVALID;                                                        \
    }

template <typename F>
static int with_ctx(ocffm_ctx *ctx, F &&fn) {
    OC_CTX(ctx);
    return guarded([&] {
        // bind the device and the launch counter for this thread
        if (auto *p = dynamic_cast<ocffm::Problem<float> *>(ctx->impl.get())) p->bind();
        else if (auto *q = dynamic_cast<ocffm::Problem<double> *>(ctx->impl.get())) q->bind();
        fn(*ctx->impl);
    });
}

extern "C" {

int ocffm_abi_version(void) { return OCFFM_ABI_VERSION; }
const char *ocffm_last_error(void) { return ocffm::g_last_error.c_str(); }

int ocffm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ocffm_create(ocffm_ctx **out, const ocffm_params *prm, uint32_t fu, uint32_t fv, uint64_t m,
                 uint64_t n) {
    if (!out || !prm) {
        ocffm::g_last_error = "null argument";
        return OCFFM_E_INVALID;
    }
    *out = nullptr;
    return guarded([&] {
        if (ocffm_device_count() <= 0)
            throw ocffm::Error(OCFFM_E_NODEVICE,
                               "no CUDA device: libocffm_cuda has no CPU fallback (sm_100a only)");
        auto ctx = std::make_unique<ocffm_ctx>();
        if (prm->dtype == OCFFM_F32) ctx->impl.reset(new ocffm::Problem<float>(*prm, fu, fv, m, n));
        else if (prm->dtype == OCFFM_F64) ctx->impl.reset(new ocffm::Problem<double>(*prm, fu, fv, m, n));
        else throw ocffm::Error(OCFFM_E_INVALID, "dtype must be OCFFM_F32 or OCFFM_F64");
        *out = ctx.release();
    });
}

int ocffm_destroy(ocffm_ctx *ctx) {
    if (!ctx) return OCFFM_OK;
    return guarded([&] { delete ctx; });
}

int ocffm_comm_unique_id(void *id128) {
    return guarded([&] {
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclUniqueId uid;
        OC_NCCL(ocffm::Nccl::get().GetUniqueId(&uid));
        memcpy(id128, &uid, sizeof(uid));
    });
}
int ocffm_shard_range(uint64_t rows, int nranks, int rank, uint64_t *lo, uint64_t *hi) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi) {
        ocffm::g_last_error = "bad rank / nranks";
        return OCFFM_E_INVALID;
    }
    *lo = rows * uint64_t(rank) / uint64_t(nranks);
    *hi = rows * uint64_t(rank + 1) / uint64_t(nranks);
    return OCFFM_OK;
}
int ocffm_comm_init(ocffm_ctx *ctx, int nranks, int rank, const void *id128) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.comm_init(nranks, rank, id128); });
}
int ocffm_set_field(ocffm_ctx *ctx, int side, uint32_t field, uint64_t rows, uint64_t D,
                    const uint64_t *rowptr, const uint32_t *idx, const double *val) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(rowptr && (idx || rowptr[rows] == 0) && (val || rowptr[rows] == 0), "null array");
        c.set_field(side, field, rows, D, rowptr, idx, val);
    });
}
int ocffm_set_labels(ocffm_ctx *ctx, uint64_t m, const uint64_t *rowptr, const uint32_t *idx,
                     const uint64_t *csc_colptr, const uint32_t *csc_rowidx, uint64_t n_ranked,
                     const double *popular) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(rowptr, "null array");
        c.set_labels(m, rowptr, idx, csc_colptr, csc_rowidx, n_ranked, popular);
    });
}
int ocffm_set_test_labels(ocffm_ctx *ctx, uint64_t m_t, const uint64_t *rowptr, const uint32_t *idx,
                          const uint64_t *nnx) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(rowptr, "null array");
        c.set_test_labels(m_t, rowptr, idx, nnx);
    });
}
int ocffm_set_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, const double *data,
                    uint64_t rows) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(data, "null array");
        c.set_block(f1, f2, which, data, rows);
    });
}
int ocffm_get_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *data, uint64_t rows) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(data, "null array");
        c.get_block(f1, f2, which, data, rows);
    });
}
int ocffm_mirror_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *pinned, uint64_t rows) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.mirror_block(f1, f2, which, pinned, rows); });
}
int ocffm_set_hyper(ocffm_ctx *ctx, double lambda, double omega, double r) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.set_hyper(lambda, omega, r); });
}
int ocffm_init_model(ocffm_ctx *ctx, uint64_t seed) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.init_model(seed); });
}
int ocffm_init_state(ocffm_ctx *ctx) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.init_state(); });
}
int ocffm_solve_block(ocffm_ctx *ctx, uint32_t f1, uint32_t f2) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        c.solve_block(f1, f2);
        c.synchronize();
    });
}
int ocffm_one_epoch(ocffm_ctx *ctx) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.one_epoch(); });
}
int ocffm_grad(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *G, uint64_t rows) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.grad(f1, f2, which, G, rows); });
}
int ocffm_hess_vec(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, const double *V, double *Hv,
                   uint64_t rows) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.hess_vec(f1, f2, which, V, Hv, rows); });
}
int ocffm_cg(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, const double *G, double *S,
             uint64_t rows, int32_t *iters) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.cg(f1, f2, which, G, S, rows, iters); });
}
int ocffm_objective(ocffm_ctx *ctx, double *value) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.objective(value); });
}
int ocffm_validate(ocffm_ctx *ctx, double *prec, double *ndcg, double *ploss, uint32_t *topk) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(prec && ndcg && ploss, "null output");
        c.validate(prec, ndcg, ploss, topk);
    });
}
int ocffm_get_vec(ocffm_ctx *ctx, const char *name, double *out, uint64_t *count) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        OC_REQUIRE(name, "null name");
        c.get_vec(name, out, count);
    });
}
int ocffm_get_embed(ocffm_ctx *ctx, uint32_t f1, uint32_t f2, int which, double *out, uint64_t rows) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.get_embed(f1, f2, which, out, rows); });
}
int ocffm_get_csc(ocffm_ctx *ctx, uint64_t *colptr, uint32_t *rowidx) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.get_csc(colptr, rowidx); });
}
int ocffm_get_stats(ocffm_ctx *ctx, ocffm_stats *out) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) {
        c.synchronize();
        c.get_stats(out);
    });
}
int ocffm_reset_stats(ocffm_ctx *ctx) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.reset_stats(); });
}
int ocffm_synchronize(ocffm_ctx *ctx) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { c.synchronize(); });
}
int ocffm_stream(ocffm_ctx *ctx, void **stream) {
    return with_ctx(ctx, [&](ocffm::CtxBase &c) { *stream = c.stream(); });
}

}  // extern "C"
