// host/train.cpp -- the `train` command of the B200 build.  Same positional arguments
// (item file first, then the training file), same flags and same error behaviour as the
// reference driver (train.cpp:34-207); the work happens in libocffm_cuda.so.
//
//   train [options] item_feature_file train_file
//
// Flags of the reference: -l -k -t -w -r -c -p -o --ns --freq (parsing stops at the first token
// that is not a flag, numeric flags only need one digit somewhere in their value, exactly like
// is_numerical there).  Added: --f64 (fp64 on the device), --device N, --load <binary model>,
// --save-binary <path>, --predict-only (evaluate a loaded model once, no training), --cache, and
// --grid-l / --grid-w: a (lambda, omega) grid like script/grid.sh:186-240 solved point after point
// on the data uploaded once; every point prints "config -l <l> -w <w>" and then exactly the log a
// separate run with those flags prints; -o / --save-binary paths get ".l<l>.w<w>" appended.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <vector>
#include <algorithm>

#include "ffm.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace std;

namespace {

struct Option {
    shared_ptr<Parameter> param = make_shared<Parameter>();
    string item_path, train_path, test_path, model_path, load_path, binary_path;
    bool predict_only = false, cache = false;
    vector<double> grid_l, grid_w;
};

// "1,4,16" -> {1, 4, 16}; every item needs a digit, like the other numeric flags
vector<double> number_list(const string &text, const string &flag) {
    vector<double> out;
    size_t pos = 0;
    while (pos <= text.size()) {
        const size_t end = min(text.find(',', pos), text.size());
        const string item = text.substr(pos, end - pos);
        bool digit = false;
        for (char c : item) digit = digit || isdigit(static_cast<unsigned char>(c));
        if (!digit) throw invalid_argument(flag + " should be followed by a comma-separated list of numbers");
        out.push_back(atof(item.c_str()));
        pos = end + 1;
    }
    return out;
}

string number_tag(double v) {
    ostringstream os;
    os << v;
    return os.str();
}

bool has_digit(const char *s) {
    for (; *s; ++s)
        if (isdigit(static_cast<unsigned char>(*s))) return true;
    return false;
}

const char *kUsage =
    "usage: train [options] item_feature_file train_file\n"
    "\n"
    "options:\n"
    "-l <lambda_2>: set regularization coefficient on r regularizer (default 1e-5)\n"
    "-t <iter>: set number of iterations (default 20)\n"
    "-p <path>: set path to test set\n"
    "-o <path>: set path to save model file\n"
    "-w <omega>: set cost weight for the negatives\n"
    "-r <rating>: set rating for the negatives\n"
    "-c <threads>: set number of host threads (kept for compatibility; the solver runs on the GPU)\n"
    "-k <rank>: set number of rank\n"
    "--ns: drop the same-side blocks\n"
    "--freq: enable freq-aware lambda\n"
    "--f64: solve in fp64 on the device (default fp32 storage, fp64 scalars)\n"
    "--gpu-init <seed>: draw the initial model on the GPU (counter-based generator, same distribution as the\n"
    "    reference's init_mat but not its random stream; default: the reference's host RNG, bit for bit)\n"
    "--device <n>: CUDA device ordinal\n"
    "--load <path>: start from a binary model (save_binary_model layout)\n"
    "--save-binary <path>: also write the binary model\n"
    "--predict-only: evaluate the loaded model on the test set and exit\n"
    "--cache: keep / reuse <file>.ocffm.bin binary caches of the parsed text files\n"
    "--grid-l <l1,l2,..> / --grid-w <w1,w2,..>: solve every (lambda, omega) pair on the data uploaded once\n";

// value of a numeric flag; `miss` is thrown when the value is absent, `bad` when it has no digit
const char *numeric_value(int argc, char **argv, int &i, const char *miss, const char *bad) {
    if (i + 1 >= argc) throw invalid_argument(miss);
    ++i;
    if (!has_digit(argv[i])) throw invalid_argument(bad);
    return argv[i];
}

Option parse_option(int argc, char **argv) {
    if (argc == 1) throw invalid_argument(kUsage);
    Option opt;
    Parameter &p = *opt.param;
    int i = 1;
    for (; i < argc; i++) {
        const string a = argv[i];
        if (a == "-l") p.lambda = atof(numeric_value(argc, argv, i, "need to specify l regularization coefficient after -l", "-l should be followed by a number"));
        else if (a == "-k") p.k = atoi(numeric_value(argc, argv, i, "need to specify rank after -k", "-k should be followed by a number"));
        else if (a == "-t") p.nr_pass = atoi(numeric_value(argc, argv, i, "need to specify max number of iterations after -t", "-t should be followed by a number"));
        else if (a == "-w") p.omega = atof(numeric_value(argc, argv, i, "need to specify the negative weight after -w", "-w should be followed by a number"));
        else if (a == "-r") p.r = atof(numeric_value(argc, argv, i, "need to specify the negative rating after -r", "-r should be followed by a number"));
        else if (a == "-c") p.nr_threads = ImpInt(atof(numeric_value(argc, argv, i, "missing core numbers after -c", "-c should be followed by a number")));
        else if (a == "--device") p.device = atoi(numeric_value(argc, argv, i, "missing ordinal after --device", "--device should be followed by a number"));
        else if (a == "-p" || a == "-o" || a == "--load" || a == "--save-binary") {
            if (i == argc - 1) throw invalid_argument("need to specify path after " + a);
            const string v = argv[++i];
            (a == "-p" ? opt.test_path : a == "-o" ? opt.model_path : a == "--load" ? opt.load_path : opt.binary_path) = v;
        } else if (a == "--grid-l" || a == "--grid-w") {
            if (i == argc - 1) throw invalid_argument("need to specify a list after " + a);
            (a == "--grid-l" ? opt.grid_l : opt.grid_w) = number_list(argv[++i], a);
        } else if (a == "--ns") p.self_side = false;
        else if (a == "--freq") p.freq = true;
        else if (a == "--f64") p.dtype = OCFFM_F64;
        else if (a == "--gpu-init") {
            p.gpu_init_seed = strtoul(numeric_value(argc, argv, i, "need to specify a seed after --gpu-init", "--gpu-init should be followed by a number"), nullptr, 10);
            if (p.gpu_init_seed == 0) throw invalid_argument("--gpu-init needs a non-zero seed");
        }
        else if (a == "--predict-only") opt.predict_only = true;
        else if (a == "--cache") opt.cache = true;
        else break;   // first non-flag token ends option parsing
    }
    if (i >= argc) throw invalid_argument("training data not specified");
    opt.item_path = argv[i++];
    if (i < argc) opt.train_path = argv[i++];
    return opt;
}

}  // namespace

int main(int argc, char *argv[]) {
    try {
        Option opt = parse_option(argc, argv);
        shared_ptr<ImpData> U = make_shared<ImpData>(opt.train_path);
        shared_ptr<ImpData> V = make_shared<ImpData>(opt.item_path);
        shared_ptr<ImpData> Ut = make_shared<ImpData>(opt.test_path);

#ifdef _OPENMP
        omp_set_num_threads(int(opt.param->nr_threads));   // -c: host threads (reader), train.cpp:174
#endif
        // parse (in parallel) or reload the binary cache of an earlier parse
        auto load = [&](shared_ptr<ImpData> &d, bool has_label, const ImpLong *ds, const char *tag) {
            const string cpath = d->file_name + (ds ? string(".") + tag : string("")) + ".ocffm.bin";
            const vector<ImpLong> *filter = ds ? &U->Ds : nullptr;   // a test file is filtered by the training Ds
            if (opt.cache && d->load_cache(cpath, filter)) return;
            d->read(has_label, ds);
            d->split_fields();
            if (opt.cache) d->save_cache(cpath, filter);
        };
        load(U, true, nullptr, "tr");
        {   // the item file's labels come from transY, which must run before its fields are cached
            const string cpath = V->file_name + ".ocffm.bin";
            if (!(opt.cache && V->load_cache(cpath))) {
                V->read(false);
                V->split_fields();
                if (opt.cache) V->save_cache(cpath);
            }
            V->transY(U->Y);
        }
        if (!Ut->file_name.empty()) load(Ut, true, U->Ds.data(), "te");

        const bool grid = !opt.grid_l.empty() || !opt.grid_w.empty();
        if (grid && (opt.predict_only || !opt.load_path.empty()))
            throw invalid_argument("--grid-l / --grid-w start every point from a fresh random model: "
                                   "not with --load or --predict-only");
        if (opt.grid_l.empty()) opt.grid_l.push_back(opt.param->lambda);
        if (opt.grid_w.empty()) opt.grid_w.push_back(opt.param->omega);
        if (grid) {
            opt.param->lambda = opt.grid_l[0];
            opt.param->omega = opt.grid_w[0];
        }

        ImpProblem prob(U, Ut, V, opt.param);
        if (!opt.load_path.empty()) prob.load_binary_model(opt.load_path);
        prob.init();
        if (grid) {
            bool first = true;
            for (double l : opt.grid_l)
                for (double w : opt.grid_w) {
                    if (!first) prob.restart(l, w);
                    first = false;
                    cout << "config -l " << number_tag(l) << " -w " << number_tag(w) << endl;
                    prob.solve();
                    const string tag = ".l" + number_tag(l) + ".w" + number_tag(w);
                    string text_path = opt.model_path + tag, binary_path = opt.binary_path + tag;
                    if (!opt.model_path.empty()) save_model(prob, text_path);
                    if (!opt.binary_path.empty()) prob.save_binary_model(binary_path);
                }
            return 0;
        }
        if (opt.predict_only) {
            if (Ut->file_name.empty()) throw invalid_argument("--predict-only needs -p <test set>");
            prob.init_va(5);
            prob.validate();
            prob.print_epoch_info(0);
        } else {
            prob.solve();
        }
        if (!opt.model_path.empty()) save_model(prob, opt.model_path);
        if (!opt.binary_path.empty()) prob.save_binary_model(opt.binary_path);
    } catch (invalid_argument &e) {
        cerr << e.what() << endl;
        return 1;
    } catch (runtime_error &e) {
        cerr << "train: " << e.what() << endl;
        return 2;
    }
    return 0;
}
