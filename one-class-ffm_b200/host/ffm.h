// host/ffm.h -- host-side mirror of the reference's public interface (ffm.h:34-154) for the
// B200 build.  Same type names, same public members and method signatures that train.cpp
// (train.cpp:177-199) and downstream tooling use, so the reference's driver compiles against
// this header unchanged; every numeric method of ImpProblem is a thin caller of the C ABI in
// include/ocffm.h (libocffm_cuda.so).  There is no CPU implementation of the solver here.
//
// Differences, all additive:
//   * ImpData keeps the per-field design matrices as SoA CSR (`Xf`) -- what the device wants --
//     instead of Node* rows; Y is still exposed as vector<Node*> over `M` because
//     train.cpp:187 passes U->Y to V->transY().  Node::val of Y is NOT the y-tilde cache any
//     more (that lives in HBM, twice, like the reference's two copies).
//   * Parameter gains `dtype` (OCFFM_F32 default / OCFFM_F64) and `device`.
//   * load_binary_model() really loads (the reference's version opens an ofstream and
//     truncates the file, ffm.cpp:1269-1301).
#pragma once
// The reference's header pulls these in for its users (ffm.h:1-31) and train.cpp relies on it
// (shared_ptr, string, cout, stoi, invalid_argument, omp_set_num_threads ... unqualified), so the
// mirror provides the same prelude, `using namespace std;` included.
#include <algorithm>
#include <cassert>
#include <climits>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>
#include <stdlib.h>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

#include <omp.h>

using namespace std;

#include "../../include/ocffm.h"

typedef double ImpFloat;
typedef double ImpDouble;
typedef unsigned int ImpInt;
typedef unsigned long int ImpLong;
typedef std::vector<ImpDouble> Vec;

const int MIN_Z = -1000;

class Parameter {
public:
    ImpFloat omega, lambda, r;
    ImpInt nr_pass, k, nr_threads;
    std::string model_path, predict_path;
    bool self_side, freq = false;
    int dtype = OCFFM_F32;   // device arithmetic; OCFFM_F64 reproduces the reference to ~1e-10
    int device = -1;         // CUDA ordinal, -1 = current
    unsigned long gpu_init_seed = 0;   // != 0: draw the initial model on the GPU (ocffm_init_model), not with init_mat's RNG
    Parameter() : omega(0.1), lambda(1e-5), r(-1), nr_pass(20), k(4), nr_threads(1), self_side(true) {}
};

class Node {
public:
    ImpInt fid;
    ImpLong idx;
    ImpDouble val;
    Node() : fid(0), idx(0), val(0) {}
};

// one field's design matrix, CSR over all rows of the file (ImpData::Xs[fi] of the reference)
struct FieldCSR {
    std::vector<ImpLong> rowptr;   // [m+1]
    std::vector<ImpInt> idx;       // [nnz]
    std::vector<ImpDouble> val;    // [nnz]
};

class ImpData {
public:
    std::string file_name;
    ImpLong m, n, f, nnz_x, nnz_y;
    std::vector<ImpLong> nnx, nny;
    std::vector<Node> M, N;          // labels (M) and, between read() and split_fields(), features (N)
    std::vector<Node *> X, Y;        // row pointers into N / M, m+1 entries each

    std::vector<FieldCSR> Xf;        // per-field CSR, filled by split_fields()
    std::vector<ImpLong> Ds;
    std::vector<std::vector<ImpLong>> freq;
    std::vector<ImpDouble> popular;

    // CSR / CSC of the labels as flat arrays (what crosses the C ABI)
    std::vector<ImpLong> y_rowptr;   // [m+1]
    std::vector<ImpInt> y_idx;       // [nnz_y]

    ImpData(std::string file_name) : file_name(file_name), m(0), n(0), f(0), nnz_x(0), nnz_y(0) {}
    void read(bool has_label, const ImpLong *ds = nullptr);
    void print_data_info();
    void split_fields();
    void transY(const std::vector<Node *> &YT);
    // additions: read() parses line ranges in parallel (OpenMP threads); a parsed
    // and split file can be cached in binary form and reloaded instead of re-parsing the text
    // `filter`: the Ds vector the file was read with (read(has_label, ds)), i.e. the TRAINING set's Ds
    // for a test file; it is recorded in the cache, and a cache written under another filter, or for
    // a source file whose size or modification time changed since, is rejected.
    void save_cache(const std::string &path, const std::vector<ImpLong> *filter = nullptr) const;   // after split_fields()
    bool load_cache(const std::string &path, const std::vector<ImpLong> *filter = nullptr);         // false: missing / stale / corrupt
};

class ImpProblem {
public:
    ImpProblem(std::shared_ptr<ImpData> &U, std::shared_ptr<ImpData> &Uva, std::shared_ptr<ImpData> &V,
               std::shared_ptr<Parameter> &param)
        : U(U), Uva(Uva), V(V), param(param) {}
    ~ImpProblem();

    void init();
    void solve();
    // Next point of a (lambda, omega) grid on the data already resident on the device
    // (script/grid.sh runs one process per point; each starts rand() from its default seed, so
    // does this): fresh random model, new hyper-parameters, caches rebuilt.  Needs init() first.
    void restart(ImpDouble new_lambda, ImpDouble new_omega);
    ImpDouble func();

    void write_header(std::ofstream &o_f) const;
    void write_W_and_H(std::ofstream &o_f) const;

    void save_binary_model(std::string &model_path);
    void load_binary_model(std::string &model_path);

    // additions ---------------------------------------------------------------------------------
    void prepare_shapes();             // fu, fv, f, k, m, n, lambda, w, r from U / V / param
    void init_model_random();          // init_mat for every block, reference order and RNG
    void attach();                     // create the device context and upload U, V, Uva
    void validate();                   // public: `predict`-only flows (ffm.cpp:925-1016)
    void print_epoch_info(ImpInt t);
    void init_va(ImpInt size);
    const Vec &block_W(ImpInt f1, ImpInt f2) const { return W[index_of(f1, f2)]; }
    const Vec &block_H(ImpInt f1, ImpInt f2) const { return H[index_of(f1, f2)]; }
    ImpDouble loss = 0;
    Vec va_loss_prec, va_loss_ndcg;
    std::vector<ImpInt> top_k;

private:
    ImpDouble lambda = 0, w = 0, r = 0;
    std::shared_ptr<ImpData> U, Uva, V;
    std::shared_ptr<Parameter> param;
    ImpInt k = 0, fu = 0, fv = 0, f = 0;
    ImpLong m = 0, n = 0, mt = 0;
    mutable std::vector<Vec> W, H;     // host copies, refreshed from the device on demand
    mutable bool host_model_stale = false;
    ocffm_ctx *ctx = nullptr;

    ImpInt index_of(ImpInt f1, ImpInt f2) const { return f2 + (f - 1) * f1 - f1 * (f1 - 1) / 2; }
    bool block_exists(ImpInt f1, ImpInt f2) const;
    ImpLong block_rows(ImpInt fg) const;
    void push_model();
    void pull_model() const;
    void one_epoch();
};

void save_model(const ImpProblem &prob, std::string &model_path);
