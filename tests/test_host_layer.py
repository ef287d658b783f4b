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
