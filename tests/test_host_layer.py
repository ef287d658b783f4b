"""Host side of the drop-in boundary (one-class-ffm_b200/host): the data layer and the RNG model
init against the reference's own reader / init_mat, bit for bit, and the CLI's error behaviour.
CPU only -- nothing here touches the GPU."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import pyoracle
from conftest import GOLDEN, ROOT, load_golden, params_of

PKG = os.path.join(ROOT, "one-class-ffm_b200")
TRAIN = os.path.join(PKG, "train")
HOST_DUMP = os.path.join(PKG, "host_dump")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "host")])


@pytest.mark.parametrize("case", ["tiny", "tiny_ns", "tiny_freq", "small"])
@pytest.mark.parametrize("mode", ["serial", "parallel", "via_cache"])
def test_reader_split_fields_transY_and_init_match_reference(case, mode):
    """serial: one thread; parallel: 5 OpenMP threads over ~64-byte line ranges; via_cache: every
    parsed file goes through save_cache/load_cache before it is used."""
    d = load_golden(case)
    prm, _ = params_of(d)
    base = os.path.join(GOLDEN, case, case)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "h.ocfd")
        cmd = [HOST_DUMP, base + ".item", base + ".tr", base + ".te", out, str(prm["k"])]
        cmd.append("--ns" if not prm["self_side"] else "--side")
        env = dict(os.environ, OMP_NUM_THREADS="1")
        if mode == "parallel":
            env.update(OMP_NUM_THREADS="5", OCFFM_READER_CHUNK="64")
        if mode == "via_cache":
            cmd += ["--via-cache", tmp]
        subprocess.check_call(cmd, env=env)
        h = pyoracle.load_ocfd(out)
    assert len(h) > 30
    for name, got in h.items():
        want = d[name]
        assert got.shape == want.shape and np.array_equal(got, want), name   # indices AND RNG draws: bit-exact


def run(args):
    return subprocess.run([TRAIN] + args, capture_output=True, text=True)


def test_cli_usage_and_argument_errors():
    r = run([])
    assert r.returncode == 1 and r.stderr.startswith("usage: train [options] item_feature_file train_file")
    r = run(["-k", "x", "a", "b"])
    assert r.returncode == 1 and "-k should be followed by a number" in r.stderr
    r = run(["-l"])
    assert r.returncode == 1 and "after -l" in r.stderr
    r = run(["-p"])
    assert r.returncode == 1 and "need to specify path after -p" in r.stderr
    r = run(["-k", "8", "--ns"])
    assert r.returncode == 1 and "training data not specified" in r.stderr
    # grid runner flags
    r = run(["--grid-l", "1,x", "a", "b"])
    assert r.returncode == 1 and "--grid-l should be followed by a comma-separated list of numbers" in r.stderr
    r = run(["--gpu-init"])
    assert r.returncode == 1 and "need to specify a seed after --gpu-init" in r.stderr
    r = run(["--gpu-init", "0", "a", "b"])
    assert r.returncode == 1 and "non-zero seed" in r.stderr
    r = run(["--grid-w"])
    assert r.returncode == 1 and "need to specify a list after --grid-w" in r.stderr
    base = os.path.join(GOLDEN, "tiny", "tiny")
    r = run(["--grid-l", "1,4", "--predict-only", base + ".item", base + ".tr"])
    assert r.returncode == 1 and "not with --load or --predict-only" in r.stderr


def test_cli_fails_loudly_without_gpu():
    import ocffm
    if ocffm.device_count() > 0:
        pytest.skip("only meaningful on a CPU-only host")
    base = os.path.join(GOLDEN, "tiny", "tiny")
    r = run(["-k", "8", "-t", "1", base + ".item", base + ".tr"])
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_cache_is_rejected_when_stale():
    """a test-set cache stores features already filtered by the TRAINING Ds; it must be
    rejected when that filter changes, and any cache when its source text changed (size or mtime)."""
    import shutil
    base = os.path.join(GOLDEN, "tiny", "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        src = {}
        for ext in ("item", "tr", "te"):
            src[ext] = shutil.copy(base + "." + ext, os.path.join(tmp, "tiny." + ext))
        subprocess.check_call([HOST_DUMP, src["item"], src["tr"], src["te"], os.path.join(tmp, "h.ocfd"), "8",
                               "--side", "--via-cache", tmp])
        d = load_golden("tiny")
        ds = ",".join(str(int(x)) for x in d["U.Ds"])

        def probe(cache, source, *filt):
            return subprocess.check_output([HOST_DUMP, "--probe-cache", os.path.join(tmp, cache), source, *filt],
                                           text=True).strip()
        assert probe("T.bin", src["te"], ds) == "hit"
        assert probe("U.bin", src["tr"]) == "hit"
        grown = ",".join(str(int(x) + 1) for x in d["U.Ds"])
        assert probe("T.bin", src["te"], grown) == "miss"        # training Ds changed -> old filtering is stale
        assert probe("T.bin", src["te"]) == "miss"               # read without a filter != read with one
        assert probe("U.bin", src["tr"], ds) == "miss"
        st = os.stat(src["tr"])
        os.utime(src["tr"], ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))   # same size, newer
        assert probe("U.bin", src["tr"]) == "miss"
        with open(os.path.join(tmp, "old.bin"), "wb") as fh:     # a cache of the previous format
            fh.write(b"OCFFMBIN1" + b"\0" * 64)
        assert probe("old.bin", src["item"]) == "miss"


@pytest.mark.parametrize("ns", [False, True])
def test_binary_model_round_trip_on_host(ns):
    """save_binary_model (reference layout, ffm.cpp:1239-1267) -> load_binary_model returns every block
    bit for bit; a file written with/without --ns is refused by the other mode with a clear message."""
    base = os.path.join(GOLDEN, "tiny", "tiny")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "m.bin")
        out = subprocess.check_output([HOST_DUMP, base + ".item", base + ".tr", base + ".te", os.path.join(tmp, "h.ocfd"),
                                       "8", "--ns" if ns else "--side", "--binary-roundtrip", path], text=True)
        assert "binary-roundtrip ok" in out
        # header of the reference layout: u32 f, fu, fv, k
        hdr = np.fromfile(path, dtype=np.uint32, count=4)
        assert list(hdr) == [4, 2, 2, 8]
        r = run(["-k", "8", "--load", path, "--predict-only", "-p", base + ".te"] + ([] if ns else ["--ns"]) +
                [base + ".item", base + ".tr"])
        assert r.returncode == 1 and "--ns" in r.stderr and "layout mismatch" in r.stderr
        bad = bytearray(open(path, "rb").read())
        bad[0] = 7                                               # f != fu + fv
        open(path, "wb").write(bytes(bad))
        r = run(["-k", "8", "--load", path, "--predict-only", "-p", base + ".te"] + (["--ns"] if ns else []) +
                [base + ".item", base + ".tr"])
        assert r.returncode == 1 and "corrupt model file" in r.stderr
