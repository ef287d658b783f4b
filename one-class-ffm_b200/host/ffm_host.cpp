// host/ffm_host.cpp -- data layer, model I/O and the ImpProblem front-end of the B200 build.
//
// The data layer reproduces the reference reader's CONTRACT (ffm.cpp:80-312: text format, what
// m / n / f / Ds / nnx / popular / freq mean, which features are dropped, CSC order of transY)
// with a single-pass buffer parser; the solver front-end only schedules C-ABI calls.
#include "ffm.h"

#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <random>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace std;

namespace {

struct Cursor {
    const char *p, *end;
    void skip_blank() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; }
    bool at_eol() const { return p >= end || *p == '\n'; }
};

// unsigned integer in the sense of `istream >> unsigned long`
bool parse_ulong(Cursor &c, ImpLong &out) {
    c.skip_blank();
    if (c.at_eol() || *c.p < '0' || *c.p > '9') return false;
    ImpLong v = 0;
    while (c.p < c.end && *c.p >= '0' && *c.p <= '9') v = v * 10 + ImpLong(*c.p++ - '0');
    out = v;
    return true;
}

bool parse_sep(Cursor &c) {   // `istream >> char`: any non-blank character
    c.skip_blank();
    if (c.at_eol()) return false;
    ++c.p;
    return true;
}

bool parse_double(Cursor &c, double &out) {
    c.skip_blank();
    if (c.at_eol()) return false;
    char buf[64];
    size_t len = 0;
    const char *q = c.p;
    while (q < c.end && len + 1 < sizeof(buf) && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r') buf[len++] = *q++;
    buf[len] = 0;
    char *stop = nullptr;
    const double v = strtod(buf, &stop);
    if (stop == buf) return false;
    c.p += (stop - buf);
    out = v;
    return true;
}

// label list "j1,j2,...": every piece goes through stoi like the reference (ffm.cpp:95-96)
void parse_labels(const char *b, const char *e, vector<ImpInt> &out) {
    const char *p = b;
    while (p < e) {
        const char *q = p;
        while (q < e && *q != ',') ++q;
        // fast path for plain digits; anything else goes through stoi for identical semantics
        // (leading blanks / sign accepted, trailing junk ignored, invalid_argument("stoi") on junk)
        bool plain = q > p && q - p <= 9;
        ImpInt v = 0;
        for (const char *d = p; plain && d < q; ++d) {
            if (*d < '0' || *d > '9') plain = false;
            else v = v * 10 + ImpInt(*d - '0');
        }
        out.push_back(plain ? v : ImpInt(stoi(string(p, q))));
        p = q + 1;
    }
}

string slurp(const string &path) {
    string buf;
    FILE *fh = fopen(path.c_str(), "rb");
    if (!fh) return buf;   // like the reference: a missing file silently reads as empty (ffm.cpp:81-88)
    fseek(fh, 0, SEEK_END);
    const long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    buf.resize(sz > 0 ? size_t(sz) : 0);
    if (sz > 0 && fread(&buf[0], 1, size_t(sz), fh) != size_t(sz)) buf.clear();
    fclose(fh);
    return buf;
}

// The reference seeds its uniform init with a fast inverse square root (one Newton step on the
// 64-bit magic-constant guess), so 0.1/sqrt(k) is only approximate; parity needs the same value.
double approx_rsqrt(double x) {
    const double half = 0.5 * x;
    uint64_t bits;
    memcpy(&bits, &x, sizeof bits);
    bits = 0x5fe6eb50c7b537a9ULL - (bits >> 1);
    memcpy(&x, &bits, sizeof bits);
    return x * (1.5 - half * x * x);
}

void check(int rc, const char *what) {
    if (rc != OCFFM_OK) throw runtime_error(string(what) + ": " + ocffm_last_error());
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// ImpData
// ---------------------------------------------------------------------------------------------
// One contiguous range of lines parsed by one thread (the reference's two-pass
// istringstream reader is serial and dominates end-to-end time once an epoch takes milliseconds).
namespace {
struct ParsedChunk {
    vector<ImpInt> labels;          // concatenated label lists of the chunk's lines
    vector<ImpLong> label_end;      // per line: end offset into `labels`
    vector<Node> nodes;             // kept features of the chunk's lines
    vector<ImpLong> node_end;       // per line: end offset into `nodes`
    ImpLong max_label1 = 0, max_fid1 = 0;
    bool bad_label = false;
};

void parse_range(const char *b, const char *e, bool has_label, const ImpLong *ds, ParsedChunk &out) {
    Cursor c{b, e};
    vector<ImpInt> labels;
    while (c.p < c.end) {
        const char *eol = static_cast<const char *>(memchr(c.p, '\n', size_t(c.end - c.p)));
        const char *line_end = eol ? eol : c.end;
        Cursor ln{c.p, line_end};
        if (has_label) {
            ln.skip_blank();
            const char *lb = ln.p;
            while (ln.p < ln.end && *ln.p != ' ' && *ln.p != '\t' && *ln.p != '\r') ++ln.p;
            labels.clear();
            try {
                parse_labels(lb, ln.p, labels);
            } catch (const exception &) {
                out.bad_label = true;      // re-thrown as invalid_argument("stoi") by the caller
                return;
            }
            for (ImpInt j : labels) {
                out.max_label1 = max<ImpLong>(out.max_label1, ImpLong(j) + 1);
                out.labels.push_back(j);
            }
        }
        ImpLong fid, idx;
        double val;
        for (;;) {
            if (!parse_ulong(ln, fid) || !parse_sep(ln) || !parse_ulong(ln, idx) || !parse_sep(ln) ||
                !parse_double(ln, val))
                break;
            out.max_fid1 = max(out.max_fid1, fid + 1);
            if (ds != nullptr && ds[fid] <= idx) continue;   // out-of-vocabulary test feature
            Node nd;
            nd.fid = ImpInt(fid);
            nd.idx = idx;
            nd.val = val;
            out.nodes.push_back(nd);
        }
        out.label_end.push_back(out.labels.size());
        out.node_end.push_back(out.nodes.size());
        c.p = eol ? eol + 1 : c.end;
    }
}
}  // namespace

void ImpData::read(bool has_label, const ImpLong *ds) {
    const string buf = slurp(file_name);
    const char *base = buf.data(), *end = buf.data() + buf.size();
    // cut the buffer at line boundaries, one range per thread
    int nthreads = 1;
#ifdef _OPENMP
    const char *mc = getenv("OCFFM_READER_CHUNK");          // smallest range worth a thread (tests lower it)
    const size_t min_chunk = mc ? size_t(max(1, atoi(mc))) : size_t(1 << 16);
    nthreads = max(1, min(omp_get_max_threads(), int(buf.size() / min_chunk) + 1));
#endif
    vector<const char *> cut(nthreads + 1, end);
    cut[0] = base;
    for (int t = 1; t < nthreads; t++) {
        const char *p = base + buf.size() * size_t(t) / size_t(nthreads);
        if (p < cut[t - 1]) p = cut[t - 1];
        const char *nl = p < end ? static_cast<const char *>(memchr(p, '\n', size_t(end - p))) : nullptr;
        cut[t] = nl ? nl + 1 : end;
    }
    vector<ParsedChunk> chunks(nthreads);
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (int t = 0; t < nthreads; t++) parse_range(cut[t], cut[t + 1], has_label, ds, chunks[t]);
    for (const ParsedChunk &c : chunks)
        if (c.bad_label) throw invalid_argument("stoi");

    // stitch the ranges together in file order
    m = 0;
    ImpLong tot_lab = 0, tot_nodes = 0;
    for (const ParsedChunk &c : chunks) {
        m += c.node_end.size();
        tot_lab += c.labels.size();
        tot_nodes += c.nodes.size();
        n = max(n, c.max_label1);
        f = max(f, c.max_fid1);
    }
    M.clear();
    N.resize(tot_nodes);
    y_idx.resize(tot_lab);
    y_rowptr.assign(m + 1, 0);
    vector<ImpLong> xptr(m + 1, 0);
    ImpLong row = 0, lab0 = 0, nod0 = 0;
    for (const ParsedChunk &c : chunks) {
        copy(c.labels.begin(), c.labels.end(), y_idx.begin() + lab0);
        copy(c.nodes.begin(), c.nodes.end(), N.begin() + nod0);
        for (size_t i = 0; i < c.node_end.size(); i++, row++) {
            y_rowptr[row + 1] = lab0 + c.label_end[i];
            xptr[row + 1] = nod0 + c.node_end[i];
        }
        lab0 += c.labels.size();
        nod0 += c.nodes.size();
    }
    nnz_x = N.size();
    nnx.resize(m);
    nny.resize(m);
    for (ImpLong i = 0; i < m; i++) {
        nnx[i] = xptr[i + 1] - xptr[i];
        nny[i] = y_rowptr[i + 1] - y_rowptr[i];
    }
    X.resize(m + 1);
    Y.resize(m + 1);
    for (ImpLong i = 0; i <= m; i++) X[i] = N.data() + xptr[i];
    if (has_label) {
        nnz_y = y_idx.size();
        M.resize(nnz_y);
        popular.assign(n, 0);
        for (ImpLong t = 0; t < nnz_y; t++) {
            M[t].idx = y_idx[t];
            popular[y_idx[t]] += 1;
        }
        for (ImpLong i = 0; i <= m; i++) Y[i] = M.data() + y_rowptr[i];
        ImpDouble total = 0;
        for (ImpDouble v : popular) total += v;
        for (ImpDouble &v : popular) v /= total;
    }
}

// ---- binary cache of a parsed + split file --------------------------------------
// Layout: magic "OCFFMBIN2", source size, source mtime (ns), the Ds filter the file was read with
// (empty for a training / item file; the training set's Ds for a test file, whose features
// idx >= Ds[fid] were dropped and whose nnx depends on it, ffm.cpp:104-105), then m n f nnz_x nnz_y,
// Ds, nnx, nny, per field (rowptr, idx, val, freq), y_rowptr, y_idx, popular.  A cache is used only
// if the text file's size AND modification time are the recorded ones and the filter is the same.
namespace {
template <typename T>
void put_vec(ofstream &o, const vector<T> &v) {
    const uint64_t n = v.size();
    o.write(reinterpret_cast<const char *>(&n), sizeof n);
    if (n) o.write(reinterpret_cast<const char *>(v.data()), streamsize(n * sizeof(T)));
}
template <typename T>
bool get_vec(ifstream &i, vector<T> &v) {
    uint64_t n = 0;
    if (!i.read(reinterpret_cast<char *>(&n), sizeof n)) return false;
    v.resize(n);
    return n == 0 || bool(i.read(reinterpret_cast<char *>(v.data()), streamsize(n * sizeof(T))));
}
// {size, mtime in ns} of the source text file; {0, 0} when it cannot be stat'ed
void file_stamp_of(const string &path, uint64_t out[2]) {
    struct stat st;
    out[0] = out[1] = 0;
    if (stat(path.c_str(), &st) != 0) return;
    out[0] = uint64_t(st.st_size);
    out[1] = uint64_t(st.st_mtim.tv_sec) * 1000000000ull + uint64_t(st.st_mtim.tv_nsec);
}
}  // namespace

void ImpData::save_cache(const string &path, const vector<ImpLong> *filter) const {
    ofstream o(path, ios::binary | ios::trunc);
    o.write("OCFFMBIN2", 9);
    uint64_t stamp[2];
    file_stamp_of(file_name, stamp);
    o.write(reinterpret_cast<const char *>(stamp), sizeof stamp);
    put_vec(o, filter ? *filter : vector<ImpLong>());
    const uint64_t hdr[5] = {m, n, f, nnz_x, nnz_y};
    o.write(reinterpret_cast<const char *>(hdr), sizeof hdr);
    put_vec(o, Ds); put_vec(o, nnx); put_vec(o, nny);
    for (ImpLong fi = 0; fi < f; fi++) {
        put_vec(o, Xf[fi].rowptr); put_vec(o, Xf[fi].idx); put_vec(o, Xf[fi].val); put_vec(o, freq[fi]);
    }
    put_vec(o, y_rowptr); put_vec(o, y_idx); put_vec(o, popular);
}

bool ImpData::load_cache(const string &path, const vector<ImpLong> *filter) {
    ifstream i(path, ios::binary);
    char magic[9];
    if (!i || !i.read(magic, 9) || memcmp(magic, "OCFFMBIN2", 9) != 0) return false;
    uint64_t stamp[2], now[2], hdr[5];
    file_stamp_of(file_name, now);
    if (!i.read(reinterpret_cast<char *>(stamp), sizeof stamp) || stamp[0] != now[0] || stamp[1] != now[1] || !now[0])
        return false;
    vector<ImpLong> recorded;
    if (!get_vec(i, recorded) || recorded != (filter ? *filter : vector<ImpLong>())) return false;
    if (!i.read(reinterpret_cast<char *>(hdr), sizeof hdr)) return false;
    m = hdr[0]; n = hdr[1]; f = hdr[2]; nnz_x = hdr[3]; nnz_y = hdr[4];
    if (!get_vec(i, Ds) || !get_vec(i, nnx) || !get_vec(i, nny)) return false;
    Xf.assign(f, FieldCSR());
    freq.assign(f, vector<ImpLong>());
    for (ImpLong fi = 0; fi < f; fi++)
        if (!get_vec(i, Xf[fi].rowptr) || !get_vec(i, Xf[fi].idx) || !get_vec(i, Xf[fi].val) || !get_vec(i, freq[fi]))
            return false;
    if (!get_vec(i, y_rowptr) || !get_vec(i, y_idx) || !get_vec(i, popular)) return false;
    // rebuild the Node view of the labels (Y is what transY() consumes)
    M.assign(y_idx.size(), Node());
    for (size_t t = 0; t < y_idx.size(); t++) M[t].idx = y_idx[t];
    Y.resize(m + 1);
    if (y_rowptr.size() == m + 1)
        for (ImpLong r = 0; r <= m; r++) Y[r] = M.data() + y_rowptr[r];
    X.clear();
    N.clear();
    return true;
}

void ImpData::split_fields() {
    Xf.assign(f, FieldCSR());
    Ds.assign(f, 0);
    freq.assign(f, vector<ImpLong>());
    for (ImpLong fi = 0; fi < f; fi++) Xf[fi].rowptr.assign(m + 1, 0);
    for (ImpLong i = 0; i < m; i++)
        for (const Node *x = X[i]; x < X[i + 1]; x++) Xf[x->fid].rowptr[i + 1]++;
    for (ImpLong fi = 0; fi < f; fi++) {
        FieldCSR &F = Xf[fi];
        for (ImpLong i = 0; i < m; i++) F.rowptr[i + 1] += F.rowptr[i];
        F.idx.resize(F.rowptr[m]);
        F.val.resize(F.rowptr[m]);
    }
    vector<ImpLong> fill(f, 0);
    for (ImpLong i = 0; i < m; i++)
        for (const Node *x = X[i]; x < X[i + 1]; x++) {   // in-line order is kept inside a field
            FieldCSR &F = Xf[x->fid];
            const ImpLong t = fill[x->fid]++;
            F.idx[t] = ImpInt(x->idx);
            F.val[t] = x->val;
            Ds[x->fid] = max(Ds[x->fid], x->idx + 1);
        }
    for (ImpLong fi = 0; fi < f; fi++) {
        freq[fi].assign(Ds[fi], 0);
        for (ImpInt id : Xf[fi].idx) freq[fi][id]++;
    }
    X.clear(); X.shrink_to_fit();
    N.clear(); N.shrink_to_fit();
}

void ImpData::transY(const vector<Node *> &YT) {
    // CSC of the label matrix ordered by (item, user): a stable counting sort over users in
    // ascending order gives the order the reference obtains with std::sort (ffm.cpp:274-279)
    n = YT.size() - 1;
    vector<ImpLong> colptr(m + 1, 0);
    ImpLong kept = 0;
    for (ImpLong i = 0; i < n; i++)
        for (const Node *y = YT[i]; y < YT[i + 1]; y++) {
            if (y->idx >= m) continue;   // label beyond the item file
            colptr[y->idx + 1]++;
            kept++;
        }
    for (ImpLong j = 0; j < m; j++) colptr[j + 1] += colptr[j];
    M.assign(kept, Node());
    y_idx.assign(kept, 0);
    vector<ImpLong> cur(colptr.begin(), colptr.end() - 1);
    for (ImpLong i = 0; i < n; i++)
        for (const Node *y = YT[i]; y < YT[i + 1]; y++) {
            if (y->idx >= m) continue;
            const ImpLong t = cur[y->idx]++;
            M[t].idx = i;
            M[t].val = y->val;
            y_idx[t] = ImpInt(i);
        }
    nnz_y = kept;
    y_rowptr = colptr;
    Y.resize(m + 1);
    for (ImpLong j = 0; j <= m; j++) Y[j] = M.data() + colptr[j];
}

void ImpData::print_data_info() {
    cout << "File:" << file_name;
    cout << setw(12) << "m:" << m;
    cout << setw(12) << "n:" << n;
    cout << setw(12) << "f:" << f;
    cout << setw(12) << "d:" << Ds[0] << endl;
}

// ---------------------------------------------------------------------------------------------
// ImpProblem
// ---------------------------------------------------------------------------------------------
ImpProblem::~ImpProblem() {
    if (ctx) ocffm_destroy(ctx);
}

bool ImpProblem::block_exists(ImpInt f1, ImpInt f2) const {
    return param->self_side || (f1 < fu && f2 >= fu);
}

ImpLong ImpProblem::block_rows(ImpInt fg) const { return fg < fu ? U->Ds[fg] : V->Ds[fg - fu]; }

void ImpProblem::init_model_random() {
    // Same calls, same order as the reference (ffm.cpp:71-78, 495-506): every matrix gets a
    // minstd engine seeded with the next rand() and D*k draws from U(-s, s), s = 0.1*rsqrt~(k).
    const ImpInt nr_blocks = f * (f + 1) / 2;
    W.assign(nr_blocks, Vec());
    H.assign(nr_blocks, Vec());
    auto fill_uniform = [&](Vec &v, ImpLong rows) {
        default_random_engine engine(rand());
        const double s = 0.1 * approx_rsqrt(double(k));
        uniform_real_distribution<ImpDouble> dist(-s, s);
        v.resize(rows * k);
        for (ImpDouble &x : v) x = dist(engine);
    };
    for (ImpInt f1 = 0; f1 < f; f1++)
        for (ImpInt f2 = f1; f2 < f; f2++) {
            if (!block_exists(f1, f2)) continue;
            fill_uniform(W[index_of(f1, f2)], block_rows(f1));
            fill_uniform(H[index_of(f1, f2)], block_rows(f2));
        }
}

void ImpProblem::attach() {
    ocffm_params prm;
    prm.lambda = param->lambda;
    prm.omega = param->omega;
    prm.r = param->r;
    prm.k = param->k;
    prm.self_side = param->self_side;
    prm.freq = param->freq;
    prm.dtype = param->dtype;
    // An unmodified reference driver (train.cpp) has no flag for the device arithmetic: the
    // environment can pick it (OCFFM_DTYPE=f64 reproduces the reference's log byte for byte).
    if (const char *e = getenv("OCFFM_DTYPE")) {
        if (!strcmp(e, "f64")) prm.dtype = OCFFM_F64;
        else if (!strcmp(e, "f32")) prm.dtype = OCFFM_F32;
    }
    prm.device = param->device;
    check(ocffm_create(&ctx, &prm, fu, fv, m, n), "ocffm_create");
    for (ImpInt fi = 0; fi < fu; fi++) {
        const FieldCSR &F = U->Xf[fi];
        check(ocffm_set_field(ctx, OCFFM_SIDE_U, fi, m, U->Ds[fi], F.rowptr.data(), F.idx.data(), F.val.data()),
              "ocffm_set_field(U)");
    }
    for (ImpInt fi = 0; fi < fv; fi++) {
        const FieldCSR &F = V->Xf[fi];
        check(ocffm_set_field(ctx, OCFFM_SIDE_V, fi, n, V->Ds[fi], F.rowptr.data(), F.idx.data(), F.val.data()),
              "ocffm_set_field(V)");
    }
    if (V->nnz_y != U->nnz_y)
        throw invalid_argument("a label exceeds the number of lines of the item file");
    check(ocffm_set_labels(ctx, m, U->y_rowptr.data(), U->y_idx.data(), V->y_rowptr.data(), V->y_idx.data(),
                           U->n, U->popular.data()),
          "ocffm_set_labels");
    if (!Uva->file_name.empty()) {
        mt = Uva->m;
        const vector<ImpLong> empty_ptr(mt + 1, 0);
        for (ImpInt fi = 0; fi < fu; fi++) {
            if (fi < Uva->Xf.size()) {
                const FieldCSR &F = Uva->Xf[fi];
                check(ocffm_set_field(ctx, OCFFM_SIDE_T, fi, mt, U->Ds[fi], F.rowptr.data(), F.idx.data(),
                                      F.val.data()),
                      "ocffm_set_field(T)");
            } else {   // the test file never mentions this field
                check(ocffm_set_field(ctx, OCFFM_SIDE_T, fi, mt, U->Ds[fi], empty_ptr.data(), nullptr, nullptr),
                      "ocffm_set_field(T)");
            }
        }
        check(ocffm_set_test_labels(ctx, mt, Uva->y_rowptr.data(), Uva->y_idx.data(), Uva->nnx.data()),
              "ocffm_set_test_labels");
    }
}

void ImpProblem::push_model() {
    for (ImpInt f1 = 0; f1 < f; f1++)
        for (ImpInt f2 = f1; f2 < f; f2++) {
            if (!block_exists(f1, f2)) continue;
            const ImpInt b = index_of(f1, f2);
            check(ocffm_set_block(ctx, f1, f2, 'W', W[b].data(), block_rows(f1)), "ocffm_set_block(W)");
            check(ocffm_set_block(ctx, f1, f2, 'H', H[b].data(), block_rows(f2)), "ocffm_set_block(H)");
        }
    host_model_stale = false;
}

void ImpProblem::pull_model() const {
    if (!host_model_stale || !ctx) return;
    for (ImpInt f1 = 0; f1 < f; f1++)
        for (ImpInt f2 = f1; f2 < f; f2++) {
            if (!block_exists(f1, f2)) continue;
            const ImpInt b = index_of(f1, f2);
            check(ocffm_get_block(ctx, f1, f2, 'W', W[b].data(), block_rows(f1)), "ocffm_get_block(W)");
            check(ocffm_get_block(ctx, f1, f2, 'H', H[b].data(), block_rows(f2)), "ocffm_get_block(H)");
        }
    host_model_stale = false;
}

void ImpProblem::prepare_shapes() {
    lambda = param->lambda;
    w = param->omega;
    r = param->r;
    m = U->m;
    n = V->m;
    fu = ImpInt(U->f);
    fv = ImpInt(V->f);
    f = fu + fv;
    k = param->k;
}

void ImpProblem::init() {
    prepare_shapes();
    const bool device_init = param->gpu_init_seed != 0 && W.empty();
    if (W.empty() && !device_init) init_model_random();   // a model loaded by load_binary_model() is kept
    attach();
    if (device_init) {
        // counter-based init on the GPU: no host RNG pass, no model upload
        check(ocffm_init_model(ctx, param->gpu_init_seed), "ocffm_init_model");
        const ImpInt nr_blocks = f * (f + 1) / 2;
        W.assign(nr_blocks, Vec());
        H.assign(nr_blocks, Vec());
        for (ImpInt f1 = 0; f1 < f; f1++)
            for (ImpInt f2 = f1; f2 < f; f2++) {
                if (!block_exists(f1, f2)) continue;
                W[index_of(f1, f2)].assign(block_rows(f1) * k, 0);
                H[index_of(f1, f2)].assign(block_rows(f2) * k, 0);
            }
        host_model_stale = true;   // the writers pull the blocks from the device on demand
    } else {
        push_model();
    }
    check(ocffm_init_state(ctx), "ocffm_init_state");
}

void ImpProblem::restart(ImpDouble new_lambda, ImpDouble new_omega) {
    if (!ctx) throw runtime_error("ImpProblem::restart() needs init() first");
    param->lambda = lambda = new_lambda;
    param->omega = w = new_omega;
    srand(1);   // the state rand() has at the start of a process (ISO C), i.e. of a separate run
    init_model_random();
    check(ocffm_set_hyper(ctx, lambda, w, r), "ocffm_set_hyper");
    push_model();
    check(ocffm_init_state(ctx), "ocffm_init_state");
}

void ImpProblem::one_epoch() {
    check(ocffm_one_epoch(ctx), "ocffm_one_epoch");
    host_model_stale = true;
}

ImpDouble ImpProblem::func() {
    double v = 0;
    check(ocffm_objective(ctx, &v), "ocffm_objective");
    return v;
}

void ImpProblem::init_va(ImpInt size) {
    if (Uva->file_name.empty()) return;
    mt = Uva->m;
    va_loss_prec.assign(size, 0);
    va_loss_ndcg.assign(size, 0);
    top_k.resize(size);
    // header line of the log, byte-compatible with the reference (ffm.cpp:899-912) so that
    // script/logs.tools keep working
    cout << "iter";
    ImpInt cut = 5;
    for (ImpInt i = 0; i < size; i++, cut *= 2) {
        top_k[i] = cut;
        cout << setw(9) << "( p@ " << cut << ", " << setw(6) << "nDCG@" << cut << " )";
    }
    cout << setw(12) << "ploss" << endl;
}

void ImpProblem::validate() {
    double prec[5], ndcg[5], pl = 0;
    check(ocffm_validate(ctx, prec, ndcg, &pl, nullptr), "ocffm_validate");
    for (size_t i = 0; i < top_k.size() && i < 5; i++) {
        va_loss_prec[i] = prec[i];
        va_loss_ndcg[i] = ndcg[i];
    }
    loss = pl;
}

void ImpProblem::print_epoch_info(ImpInt t) {
    cout << setw(2) << t + 1;
    if (!Uva->file_name.empty()) {
        for (size_t i = 0; i < top_k.size(); i++) {
            cout << setw(9) << "( " << setprecision(3) << va_loss_prec[i] * 100 << " ,";
            cout << setw(6) << setprecision(3) << va_loss_ndcg[i] * 100 << " )";
        }
        cout << setw(13) << setprecision(3) << loss;
    }
    cout << endl;
}

void ImpProblem::solve() {
    init_va(5);
    for (ImpInt iter = 0; iter < param->nr_pass; iter++) {
        one_epoch();
        if (!Uva->file_name.empty() && iter % 10 == 9) {   // ffm.cpp:1155-1158
            validate();
            print_epoch_info(iter);
        }
    }
}

// ---- model output: layouts of ffm.cpp:1163-1267 -------------------------------------------------
void ImpProblem::write_header(ofstream &o_f) const {
    o_f << f << endl << fu << endl << fv << endl << k << endl;
    for (ImpInt fi = 0; fi < fu; fi++) o_f << U->Ds[fi] << endl;
    for (ImpInt fi = 0; fi < fv; fi++) o_f << V->Ds[fi] << endl;
}

void ImpProblem::write_W_and_H(ofstream &o_f) const {
    pull_model();
    auto rows_out = [&](const Vec &blk, ImpLong rows, char tag, ImpInt fi, ImpInt fj) {
        for (ImpLong row = 0; row < rows; row++) {
            o_f << tag << ',' << fi << ',' << fj << ',' << row;
            for (ImpInt d = 0; d < k; d++) o_f << " " << blk[row * k + d];
            o_f << endl;
        }
    };
    for (ImpInt fi = 0; fi < f; fi++)
        for (ImpInt fj = fi; fj < f; fj++) {
            if (!block_exists(fi, fj)) continue;
            rows_out(W[index_of(fi, fj)], block_rows(fi), 'W', fi, fj);
            rows_out(H[index_of(fi, fj)], block_rows(fj), 'H', fi, fj);
        }
}

void save_model(const ImpProblem &prob, string &model_path) {
    ofstream f_out(model_path, ios::out | ios::trunc);
    prob.write_header(f_out);
    prob.write_W_and_H(f_out);
}

void ImpProblem::save_binary_model(string &model_path) {
    pull_model();
    ofstream of(model_path, ios::binary | ios::trunc);
    auto put = [&](const void *p, size_t bytes) { of.write(static_cast<const char *>(p), streamsize(bytes)); };
    put(&f, sizeof(ImpInt));
    put(&fu, sizeof(ImpInt));
    put(&fv, sizeof(ImpInt));
    put(&k, sizeof(ImpInt));
    put(U->Ds.data(), sizeof(ImpLong) * fu);
    put(V->Ds.data(), sizeof(ImpLong) * fv);
    for (ImpInt fi = 0; fi < f; fi++)
        for (ImpInt fj = fi; fj < f; fj++) {
            if (!block_exists(fi, fj)) continue;
            ImpInt fij = index_of(fi, fj);
            ImpLong wn = W[fij].size(), hn = H[fij].size();
            put(&fij, sizeof(ImpInt));
            put(&wn, sizeof(ImpLong));
            put(&hn, sizeof(ImpLong));
            put(W[fij].data(), sizeof(ImpDouble) * wn);
            put(H[fij].data(), sizeof(ImpDouble) * hn);
        }
}

// A loader that loads.  Call after U/V are read and before init(): init() then
// keeps these blocks instead of drawing random ones, which is also how parity tests inject state.
void ImpProblem::load_binary_model(string &model_path) {
    ifstream in(model_path, ios::binary);
    if (!in) throw invalid_argument("cannot open model file " + model_path);
    auto get = [&](void *p, size_t bytes) {
        in.read(static_cast<char *>(p), streamsize(bytes));
        if (!in) throw invalid_argument("truncated model file " + model_path);
    };
    ImpInt mf, mfu, mfv, mk;
    get(&mf, sizeof mf); get(&mfu, sizeof mfu); get(&mfv, sizeof mfv); get(&mk, sizeof mk);
    if (mf != mfu + mfv) throw invalid_argument("corrupt model file " + model_path + " (f != fu + fv)");
    if (mfu != U->f || mfv != V->f || mk != param->k)
        throw invalid_argument("model file does not match the data (fields or k differ)");
    fu = mfu; fv = mfv; f = mf; k = mk;
    vector<ImpLong> du(fu), dv(fv);
    get(du.data(), sizeof(ImpLong) * fu);
    get(dv.data(), sizeof(ImpLong) * fv);
    for (ImpInt i = 0; i < fu; i++)
        if (du[i] != U->Ds[i]) throw invalid_argument("model file does not match the data (Ds differ)");
    for (ImpInt i = 0; i < fv; i++)
        if (dv[i] != V->Ds[i]) throw invalid_argument("model file does not match the data (Ds differ)");
    const ImpInt nr_blocks = f * (f + 1) / 2;
    W.assign(nr_blocks, Vec());
    H.assign(nr_blocks, Vec());
    for (ImpInt fi = 0; fi < f; fi++)
        for (ImpInt fj = fi; fj < f; fj++) {
            if (!block_exists(fi, fj)) continue;
            ImpInt fij;
            ImpLong wn, hn;
            get(&fij, sizeof fij); get(&wn, sizeof wn); get(&hn, sizeof hn);
            if (fij != index_of(fi, fj))   // the first stored block tells: (0,0) with same-side blocks, (0,fu) under --ns
                throw invalid_argument(string("model file block layout mismatch: the file was written ") +
                                       (param->self_side ? "with" : "without") + " --ns, this run is " +
                                       (param->self_side ? "without" : "with") + " it");
            if (wn != block_rows(fi) * k || hn != block_rows(fj) * k)
                throw invalid_argument("model file block layout mismatch");
            W[fij].resize(wn);
            H[fij].resize(hn);
            get(W[fij].data(), sizeof(ImpDouble) * wn);
            get(H[fij].data(), sizeof(ImpDouble) * hn);
        }
    if (ctx) {
        push_model();
        check(ocffm_init_state(ctx), "ocffm_init_state");
    }
}
