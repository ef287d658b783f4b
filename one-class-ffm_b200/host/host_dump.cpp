// host/host_dump.cpp -- test helper: runs ONLY the host data layer (reader, split_fields,
// transY) and the RNG model init on text files and writes the arrays in the OCFD1 container
// of oracle/ref_harness.cpp, so tests can compare them bit for bit with the reference's dumps.
// No GPU involved.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ffm.h"

using namespace std;

static FILE *g_out;
template <typename T> const char *dt();
template <> const char *dt<double>() { return "f8"; }
template <> const char *dt<unsigned long>() { return "u8"; }
template <> const char *dt<unsigned int>() { return "u4"; }

template <typename T>
static void put(const string &name, const vector<T> &v) {
    fprintf(g_out, "%s %s 1 %zu\n", name.c_str(), dt<T>(), v.size());
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), g_out);
}

static void dump(const string &pfx, ImpData &d, bool has_label) {
    vector<ImpLong> hdr = {d.m, d.n, d.f, d.nnz_x, d.nnz_y};
    put(pfx + ".hdr", hdr);
    put(pfx + ".Ds", d.Ds);
    put(pfx + ".nnx", d.nnx);
    for (ImpLong fi = 0; fi < d.f; fi++) {
        put(pfx + ".X" + to_string(fi) + ".rowptr", d.Xf[fi].rowptr);
        vector<ImpLong> idx(d.Xf[fi].idx.begin(), d.Xf[fi].idx.end());
        put(pfx + ".X" + to_string(fi) + ".idx", idx);
        put(pfx + ".X" + to_string(fi) + ".val", d.Xf[fi].val);
        put(pfx + ".freq" + to_string(fi), d.freq[fi]);
    }
    if (has_label) {
        put(pfx + ".Y.rowptr", d.y_rowptr);
        vector<ImpLong> idx(d.y_idx.begin(), d.y_idx.end());
        put(pfx + ".Y.idx", idx);
        put(pfx + ".popular", d.popular);
    }
}

int main(int argc, char **argv) {
    // host_dump --probe-cache <cache> <source text file> [d0,d1,...]: would load_cache accept it?
    if (argc >= 4 && !strcmp(argv[1], "--probe-cache")) {
        ImpData d(argv[3]);
        vector<ImpLong> filter;
        if (argc > 4)
            for (char *tok = strtok(argv[4], ","); tok; tok = strtok(nullptr, ",")) filter.push_back(strtoul(tok, nullptr, 10));
        puts(d.load_cache(argv[2], argc > 4 ? &filter : nullptr) ? "hit" : "miss");
        return 0;
    }
    if (argc < 6) {
        fprintf(stderr, "usage: host_dump <item> <train> <test|-> <out.ocfd> <k> [--ns]\n");
        return 2;
    }
    shared_ptr<ImpData> U = make_shared<ImpData>(argv[2]);
    shared_ptr<ImpData> V = make_shared<ImpData>(argv[1]);
    shared_ptr<ImpData> Ut = make_shared<ImpData>(strcmp(argv[3], "-") ? argv[3] : "");
    // optional last argument "--via-cache <dir>": every file goes text -> cache -> fresh object
    string cache_dir;
    for (int i = 6; i + 1 < argc; i++)
        if (!strcmp(argv[i], "--via-cache")) cache_dir = argv[i + 1];
    auto roundtrip = [&](shared_ptr<ImpData> &d, const char *tag, const vector<ImpLong> *filter = nullptr) {
        if (cache_dir.empty()) return;
        const string path = cache_dir + "/" + tag + ".bin";
        d->save_cache(path, filter);
        shared_ptr<ImpData> fresh = make_shared<ImpData>(d->file_name);
        if (!fresh->load_cache(path, filter)) { fprintf(stderr, "cache reload failed for %s\n", tag); exit(3); }
        d = fresh;
    };
    U->read(true);
    U->split_fields();
    roundtrip(U, "U");
    V->read(false);
    V->split_fields();
    roundtrip(V, "V");
    V->transY(U->Y);
    if (!Ut->file_name.empty()) {
        Ut->read(true, U->Ds.data());
        Ut->split_fields();
        roundtrip(Ut, "T", &U->Ds);
    }
    g_out = fopen(argv[4], "wb");
    fputs("OCFD1\n", g_out);
    dump("U", *U, true);
    dump("V", *V, false);
    put("V.Y.rowptr", V->y_rowptr);
    {
        vector<ImpLong> idx(V->y_idx.begin(), V->y_idx.end());
        put("V.Y.idx", idx);
    }
    if (!Ut->file_name.empty()) dump("T", *Ut, true);
    // RNG init exactly as ImpProblem::init() would draw it (no device needed)
    shared_ptr<Parameter> prm = make_shared<Parameter>();
    prm->k = atoi(argv[5]);
    prm->self_side = !(argc > 6 && !strcmp(argv[6], "--ns"));
    ImpProblem prob(U, Ut, V, prm);
    prob.prepare_shapes();
    prob.init_model_random();
    const ImpInt f = ImpInt(U->f + V->f), fu = ImpInt(U->f);
    for (ImpInt f1 = 0; f1 < f; f1++)
        for (ImpInt f2 = f1; f2 < f; f2++) {
            if (!prm->self_side && !(f1 < fu && f2 >= fu)) continue;
            put("init." + to_string(f1) + "_" + to_string(f2) + ".W", prob.block_W(f1, f2));
            put("init." + to_string(f1) + "_" + to_string(f2) + ".H", prob.block_H(f1, f2));
        }
    fclose(g_out);
    // optional "--binary-roundtrip <path>": save_binary_model (the reference's layout,
    // ffm.cpp:1239-1267) -> load_binary_model into a FRESH problem; every block must come back
    // bit for bit.  No device involved (the loader pushes to the device only when one is attached).
    for (int i = 6; i + 1 < argc; i++)
        if (!strcmp(argv[i], "--binary-roundtrip")) {
            string path = argv[i + 1];
            prob.save_binary_model(path);
            ImpProblem fresh(U, Ut, V, prm);
            fresh.prepare_shapes();
            fresh.load_binary_model(path);
            for (ImpInt f1 = 0; f1 < f; f1++)
                for (ImpInt f2 = f1; f2 < f; f2++) {
                    if (!prm->self_side && !(f1 < fu && f2 >= fu)) continue;
                    if (fresh.block_W(f1, f2) != prob.block_W(f1, f2) || fresh.block_H(f1, f2) != prob.block_H(f1, f2)) {
                        fprintf(stderr, "binary model round trip: block (%u,%u) differs\n", f1, f2);
                        return 4;
                    }
                }
            printf("binary-roundtrip ok\n");
        }
    return 0;
}
