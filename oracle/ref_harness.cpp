// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Drives the UNMODIFIED reference (ffm.cpp / ffm.h compiled where they lie under
// /root/reference, see oracle/Makefile) phase by phase and dumps every intermediate
// the parity tests need.  Nothing here restates the algorithm: every number written
// comes out of a reference function.  The only trick is `#define private public`
// (after the std headers, otherwise libstdc++'s <sstream> breaks) so the harness can
// call ImpProblem's private phases (gd_*, hs_*, cg, update_*, validate, ...).
//
//   ref_harness dump  <item> <train> <test|-> <out.ocfd> [train-style flags]
//   ref_harness time  <item> <train> <test|-> <epochs>   [train-style flags]
//
// Dump container ("OCFD1"): repeated records  "<name> <dtype> <ndim> <dims...>\n" + raw bytes.
#include <iostream>
#include <random>
#include <fstream>
#include <sstream>
#include <memory>
#include <cstring>
#include <stdlib.h>
#include <unordered_set>
#include <algorithm>
#include <functional>
#include <iomanip>
#include <climits>
#include <utility>
#include <numeric>
#include <cassert>
#include <chrono>
#include <immintrin.h>
#include <omp.h>

#define private public
#include "ffm.h"
#undef private

extern "C" { long ocffm_shim_dscal_calls = 0; }

// reference free functions we call (ffm.cpp:21-60); declared here, defined in ffm.cpp
void axpy(const ImpDouble *x, ImpDouble *y, const ImpLong &l, const ImpDouble &lambda);
void mm(const ImpDouble *a, const ImpDouble *b, ImpDouble *c, const ImpLong l, const ImpLong n, const ImpInt k);
void mm(const ImpDouble *a, const ImpDouble *b, ImpDouble *c, const ImpLong k, const ImpLong l);
const ImpInt index_vec(const ImpInt f1, const ImpInt f2, const ImpInt f);

static FILE *g_out = nullptr;

template <typename T> static const char *dtype_of();
template <> const char *dtype_of<double>() { return "f8"; }
template <> const char *dtype_of<unsigned long>() { return "u8"; }
template <> const char *dtype_of<unsigned int>() { return "u4"; }
template <> const char *dtype_of<long>() { return "i8"; }

template <typename T>
static void put(const std::string &name, const T *data, std::vector<size_t> dims) {
    size_t n = 1;
    fprintf(g_out, "%s %s %zu", name.c_str(), dtype_of<T>(), dims.size());
    for (size_t d : dims) { fprintf(g_out, " %zu", d); n *= d; }
    fputc('\n', g_out);
    if (n) fwrite(data, sizeof(T), n, g_out);
}
template <typename T>
static void put(const std::string &name, const std::vector<T> &v) { put(name, v.data(), {v.size()}); }
static void put_scalar(const std::string &name, double v) { put(name, &v, {1}); }

static void dump_csr(const std::string &pfx, const vector<Node *> &X, ImpLong m, bool with_val) {
    vector<ImpLong> rowptr(m + 1), idx;
    vector<double> val;
    for (ImpLong i = 0; i <= m; i++) rowptr[i] = X[i] - X[0];
    for (const Node *x = X[0]; x < X[m]; x++) { idx.push_back(x->idx); val.push_back(x->val); }
    put(pfx + ".rowptr", rowptr);
    put(pfx + ".idx", idx);
    if (with_val) put(pfx + ".val", val);
}

static void dump_data(const std::string &pfx, ImpData &d, bool has_label) {
    vector<ImpLong> hdr = {d.m, d.n, d.f, d.nnz_x, d.nnz_y};
    put(pfx + ".hdr", hdr);
    put(pfx + ".Ds", d.Ds);
    put(pfx + ".nnx", d.nnx);
    for (ImpLong fi = 0; fi < d.f; fi++) {
        dump_csr(pfx + ".X" + to_string(fi), d.Xs[fi], d.m, true);
        put(pfx + ".freq" + to_string(fi), d.freq[fi]);
    }
    if (has_label) {
        dump_csr(pfx + ".Y", d.Y, d.m, false);
        put(pfx + ".popular", d.popular);
    }
}

static void dump_yvals(const std::string &name, const vector<Node *> &Y, ImpLong m) {
    vector<double> v;
    for (const Node *y = Y[0]; y < Y[m]; y++) v.push_back(y->val);
    put(name, v);
}

static void dump_state(const std::string &pfx, ImpProblem &p, bool with_embed = true) {
    const ImpInt f = p.f;
    for (ImpInt f1 = 0; f1 < f; f1++)
        for (ImpInt f2 = f1; f2 < f; f2++) {
            const ImpInt f12 = index_vec(f1, f2, f);
            if (p.W[f12].empty()) continue;
            const std::string b = pfx + "." + to_string(f1) + "_" + to_string(f2);
            put(b + ".W", p.W[f12]);
            put(b + ".H", p.H[f12]);
            if (with_embed) {
                put(b + ".P", p.P[f12]);
                put(b + ".Q", p.Q[f12]);
            }
        }
    put(pfx + ".a", p.a);
    put(pfx + ".b", p.b);
    put(pfx + ".sa", p.sa);
    put(pfx + ".sb", p.sb);
    dump_yvals(pfx + ".ytilde_csr", p.U->Y, p.U->m);
    dump_yvals(pfx + ".ytilde_csc", p.V->Y, p.V->m);
}

// One half block solve, observed but NOT applied: G from gd_*, Hv for V = -G the way
// cg() assembles it (ffm.cpp:783-801), S and the CG iteration count from cg() itself.
static void probe_half(const std::string &pfx, ImpProblem &p, bool side, ImpInt f1, ImpInt f2,
                       ImpInt f12, bool w_part) {
    Vec &W1 = w_part ? p.W[f12] : p.H[f12];
    Vec &Q1 = w_part ? p.Q[f12] : p.P[f12];
    Vec &P1 = w_part ? p.P[f12] : p.Q[f12];
    const ImpInt fa = w_part ? f1 : f2, fb = w_part ? f2 : f1;
    const ImpLong sz = W1.size();
    Vec G(sz, 0), S(sz, 0);
    if (side) p.gd_side(fa, W1, Q1, G);
    else p.gd_cross(fa, f12, Q1, W1, G);
    put(pfx + ".G", G);

    // Hv for the first CG direction V = -G
    const bool fa_user = fa < p.fu;
    shared_ptr<ImpData> U1 = fa_user ? p.U : p.V;
    const ImpInt fi = fa_user ? fa : fa - p.fu;
    const vector<Node *> &X = U1->Xs[fi];
    const vector<Node *> &Y = U1->Y;
    const ImpLong m1 = fa_user ? p.m : p.n, n1 = fa_user ? p.n : p.m;
    const ImpLong Df1 = U1->Ds[fi];
    const ImpInt k = p.k;
    Vec V(sz), Hv(sz, 0), Hv_(p.param->nr_threads * sz, 0);
    for (ImpLong i = 0; i < sz; i++) V[i] = -G[i];
    if (p.param->freq) {
        for (ImpLong i = 0; i < Df1; i++)
            axpy(V.data() + i * k, Hv.data() + i * k, k, p.lambda * ImpDouble(U1->freq[fi][i]));
    } else {
        axpy(V.data(), Hv.data(), sz, p.lambda);
    }
    if (side) {
        p.hs_side(m1, n1, V, Hv, Q1, X, Y, Hv_);
    } else {
        Vec QTQ(k * k, 0), VQTQ(sz, 0);
        mm(Q1.data(), Q1.data(), QTQ.data(), k, n1);
        mm(V.data(), QTQ.data(), VQTQ.data(), Df1, k, k);
        p.hs_cross(m1, n1, V, VQTQ, Hv, Q1, X, Y, Hv_);
        put(pfx + ".QTQ", QTQ);
    }
    put(pfx + ".Hv", Hv);

    const long c0 = ocffm_shim_dscal_calls;
    p.cg(fa, fb, S, Q1, G, P1);
    put(pfx + ".S", S);
    put_scalar(pfx + ".cg_iters", double(ocffm_shim_dscal_calls - c0));
}

struct Args {
    shared_ptr<Parameter> param = make_shared<Parameter>();
    string item, train, test;
};

static int parse_flags(int argc, char **argv, int i, Args &a) {
    for (; i < argc; i++) {
        string s = argv[i];
        if (s == "-l") a.param->lambda = atof(argv[++i]);
        else if (s == "-k") a.param->k = atoi(argv[++i]);
        else if (s == "-t") a.param->nr_pass = atoi(argv[++i]);
        else if (s == "-w") a.param->omega = atof(argv[++i]);
        else if (s == "-r") a.param->r = atof(argv[++i]);
        else if (s == "-c") a.param->nr_threads = atoi(argv[++i]);
        else if (s == "--ns") a.param->self_side = false;
        else if (s == "--freq") a.param->freq = true;
        else { fprintf(stderr, "unknown flag %s\n", s.c_str()); return 1; }
    }
    return 0;
}

static void load(Args &a, shared_ptr<ImpData> &U, shared_ptr<ImpData> &V, shared_ptr<ImpData> &Ut) {
    // same sequence as the reference's main (train.cpp:177-193)
    U = make_shared<ImpData>(a.train);
    V = make_shared<ImpData>(a.item);
    Ut = make_shared<ImpData>(a.test);
    U->read(true);
    U->split_fields();
    V->read(false);
    V->transY(U->Y);
    V->split_fields();
    if (!Ut->file_name.empty()) {
        Ut->read(true, U->Ds.data());
        Ut->split_fields();
    }
}

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: ref_harness dump|time <item> <train> <test|-> <out|epochs> [flags]\n");
        return 2;
    }
    const string mode = argv[1];
    Args a;
    a.item = argv[2];
    a.train = argv[3];
    a.test = (string(argv[4]) == "-") ? "" : argv[4];
    const string out_or_epochs = argv[5];
    if (parse_flags(argc, argv, 6, a)) return 2;
    omp_set_num_threads(a.param->nr_threads);

    shared_ptr<ImpData> U, V, Ut;
    auto t0 = chrono::steady_clock::now();
    load(a, U, V, Ut);
    auto t1 = chrono::steady_clock::now();
    ImpProblem prob(U, Ut, V, a.param);

    if (mode == "time") {
        const int epochs = atoi(out_or_epochs.c_str());
        prob.init();
        auto t2 = chrono::steady_clock::now();
        // silence init_va's header
        streambuf *old = cout.rdbuf();
        ostringstream sink;
        cout.rdbuf(sink.rdbuf());
        prob.init_va(5);
        cout.rdbuf(old);
        printf("{\"m\": %lu, \"n\": %lu, \"nnz_y\": %lu, \"threads\": %u, \"read_s\": %.6f, \"init_s\": %.6f, \"epochs\": [",
               U->m, V->m, U->nnz_y, a.param->nr_threads,
               chrono::duration<double>(t1 - t0).count(), chrono::duration<double>(t2 - t1).count());
        for (int e = 0; e < epochs; e++) {
            // one_epoch()'s own block order (ffm.cpp:852-870), block by block through the
            // reference's solve_side / solve_cross / cache_sasb so that per-block CG counts
            // (needed for N_trav) can be read off the shim counter
            const long c0 = ocffm_shim_dscal_calls;
            string blocks;
            auto s = chrono::steady_clock::now();
            auto run = [&](ImpInt f1, ImpInt f2, bool side) {
                const long b0 = ocffm_shim_dscal_calls;
                if (side) prob.solve_side(f1, f2);
                else prob.solve_cross(f1, f2);
                blocks += (blocks.empty() ? "" : ", ") + string("{\"f1\": ") + to_string(f1) + ", \"f2\": " +
                          to_string(f2) + ", \"cg\": " + to_string(ocffm_shim_dscal_calls - b0) + "}";
            };
            const ImpInt fu = prob.fu, f = prob.f;
            if (a.param->self_side) {
                for (ImpInt f1 = 0; f1 < fu; f1++)
                    for (ImpInt f2 = f1; f2 < fu; f2++) run(f1, f2, true);
                for (ImpInt f1 = fu; f1 < f; f1++)
                    for (ImpInt f2 = f1; f2 < f; f2++) run(f1, f2, true);
            }
            for (ImpInt f1 = 0; f1 < fu; f1++)
                for (ImpInt f2 = fu; f2 < f; f2++) run(f1, f2, false);
            if (a.param->self_side) prob.cache_sasb();
            auto t = chrono::steady_clock::now();
            printf("%s{\"sec\": %.6f, \"cg_iters\": %ld, \"blocks\": [%s]}", e ? ", " : "",
                   chrono::duration<double>(t - s).count(), ocffm_shim_dscal_calls - c0, blocks.c_str());
            fflush(stdout);
        }
        printf("]");
        if (!Ut->file_name.empty()) {
            auto s = chrono::steady_clock::now();
            prob.validate();
            auto t = chrono::steady_clock::now();
            printf(", \"validate_s\": %.6f, \"m_t\": %lu, \"p_at_10\": %.9g, \"ndcg_at_10\": %.9g, \"ploss\": %.9g",
                   chrono::duration<double>(t - s).count(), Ut->m,
                   prob.va_loss_prec[1], prob.va_loss_ndcg[1], prob.loss);
        }
        printf("}\n");
        return 0;
    }

    g_out = fopen(out_or_epochs.c_str(), "wb");
    if (!g_out) { perror("open"); return 1; }
    fputs("OCFD1\n", g_out);
    vector<double> prm = {a.param->lambda, a.param->omega, a.param->r, double(a.param->k),
                          double(a.param->nr_pass), double(a.param->self_side), double(a.param->freq)};
    put("params", prm);
    dump_data("U", *U, true);
    dump_data("V", *V, false);
    dump_csr("V.Y", V->Y, V->m, false);   // CSC of Omega built by transY (ffm.cpp:259-294)
    if (!Ut->file_name.empty()) dump_data("T", *Ut, true);

    prob.init();
    dump_state("init", prob);
    const bool small = double(U->m) * double(V->m) <= 4e6;
    const bool small_dump = double(U->m) * double(V->m) <= 1e4;  // P,Q after epochs only for tiny cases
    if (a.param->self_side && small) put_scalar("init.func", prob.func());

    const ImpInt fu = prob.fu, f = prob.f;
    if (a.param->self_side) {
        probe_half("probe.side_u.W", prob, true, 0, 0, index_vec(0, 0, f), true);
        probe_half("probe.side_u.H", prob, true, 0, 0, index_vec(0, 0, f), false);
        probe_half("probe.side_v.W", prob, true, fu, f - 1, index_vec(fu, f - 1, f), true);
        probe_half("probe.side_v.H", prob, true, fu, f - 1, index_vec(fu, f - 1, f), false);
    }
    probe_half("probe.cross.W", prob, false, 0, fu, index_vec(0, fu, f), true);
    probe_half("probe.cross.H", prob, false, 0, fu, index_vec(0, fu, f), false);
    probe_half("probe.cross_last.W", prob, false, fu - 1, f - 1, index_vec(fu - 1, f - 1, f), true);
    probe_half("probe.cross_last.H", prob, false, fu - 1, f - 1, index_vec(fu - 1, f - 1, f), false);

    streambuf *old = cout.rdbuf();
    ostringstream sink;
    cout.rdbuf(sink.rdbuf());
    prob.init_va(5);
    cout.rdbuf(old);

    vector<double> funcs, cgs;
    for (ImpInt e = 0; e < a.param->nr_pass; e++) {
        const long c0 = ocffm_shim_dscal_calls;
        prob.one_epoch();
        cgs.push_back(double(ocffm_shim_dscal_calls - c0));
        if (a.param->self_side && small) funcs.push_back(prob.func());
        if (e == 0 && small_dump) dump_state("epoch1", prob, true);
    }
    put("epochs.func", funcs);
    put("epochs.cg_iters", cgs);
    dump_state("final", prob, small_dump);

    if (!Ut->file_name.empty()) {
        prob.validate();
        put("va.prec", prob.va_loss_prec);
        put("va.ndcg", prob.va_loss_ndcg);
        put_scalar("va.ploss", prob.loss);
        // raw scores the reference ranks: bt + sum_cross Qva*Pva_i (ffm.cpp:948-981), or `popular`
        const ImpLong mt = Ut->m, n = V->m;
        Vec at(mt, 0), bt(n, 0);
        if (a.param->self_side) {
            for (ImpInt f1 = 0; f1 < fu; f1++)
                for (ImpInt f2 = f1; f2 < fu; f2++) {
                    const ImpInt f12 = index_vec(f1, f2, f);
                    prob.add_side(prob.Pva[f12], prob.Qva[f12], mt, at);
                }
            for (ImpInt f1 = fu; f1 < f; f1++)
                for (ImpInt f2 = f1; f2 < f; f2++) {
                    const ImpInt f12 = index_vec(f1, f2, f);
                    prob.add_side(prob.Pva[f12], prob.Qva[f12], n, bt);
                }
        }
        put("va.at", at);
        put("va.bt", bt);
        if (double(mt) * double(n) <= 4e6) {
            Vec Z(mt * n, 0);
            for (ImpLong i = 0; i < mt; i++) {
                copy(bt.begin(), bt.end(), Z.begin() + i * n);
                prob.pred_z(i, Z.data() + i * n);
            }
            put("va.Z", Z.data(), {mt, n});
        }
    }
    fclose(g_out);
    return 0;
}
