"""CPU-side checks: the C-ABI library loads and exports every declared symbol, fails loudly
without a GPU, and the host mirror of the reference API behaves like the reference (checked
against the oracle and the golden vectors)."""
import os
import re
import sys

import numpy as np
import pytest

import bias_jl_b200 as b
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bias_cuda.h")).read()
    declared = set(re.findall(r"BIAS_API\s+(?:const\s+char\s*\*|int)\s*(bias_\w+)\s*\(", hdr))
    assert declared == set(b.SYMBOLS), declared ^ set(b.SYMBOLS)
    L = b.lib()
    for name in declared:
        assert hasattr(L, name), f"libbias_cuda.so does not export {name}"
    assert L.bias_version() >= 100


def test_no_cpu_fallback():
    if b.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(b.BiasError) as e:
        b.Context(0, 1)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bias.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle-backed", ""), (dirpath, f)


def test_host_conjugates_match_oracle(golden):
    g = golden["g1g1_running_mean"]
    q = b.Gaussian1DGaussian1D(*g["ctor"])
    for x in g["add"]:
        q.additem(x)
    assert q.mu == 3.0999999999999996 and q.nn == 3
    assert abs(q.logpredictive(3.2) - (-0.7246443915)) < 1e-9
    q.delitem(2.8)
    assert q.mu == 3.2499999999999996
    p = q.posterior()
    assert isinstance(p, b.Gaussian1D)
    e = b.Gaussian1DGaussian1D(2, 10, 0.5)
    with pytest.raises(RuntimeError):
        e.delitem(1.0)
    md = b.MultinomialDirichlet(5, 2.0)
    assert md.aa == golden["md_ctor"]["aa_stored"]
    c = orc.Comps.md(1, 5, 2.0)
    d = orc.Data.words(np.array([0, 0, 3]))
    for i, w in enumerate([0, 0, 3]):
        md.additem(w)
        orc.additem(c, 0, d, i)
    assert md.logpredictive(0) == orc.logpredictive(c, 0, d, 0)
    s = b.Sent([4, 1], [2, 1])
    sd = orc.Data.sents(np.array([0, 2]), np.array([4, 1]), np.array([2, 1]))
    assert abs(md.logpredictive(s) - orc.logpredictive(c, 0, sd, 0)) < 1e-12
    assert abs(md.loglikelihood(s) - orc.loglikelihood(c, 0, sd, 0)) < 1e-15
    assert np.allclose(b.Dirichlet(5, 4).alpha, golden["dirichlet"]["alpha"])
    assert np.allclose(b.Dirichlet(5, 4).mean(), golden["dirichlet"]["mean"])
    gp = golden["gaussian1d_pdf"]
    assert abs(b.Gaussian1D(gp["mu"], gp["vv"]).pdf(gp["x"]) - 0.2650035323) < 1e-9
    with pytest.raises(ValueError):
        b.Gaussian1D(0.0, 0.0)


def test_gen_bars_matches_oracle_and_doc(golden):
    g = golden["gen_bars_4_25"]
    bars = b.gen_bars(4, 25, 0.0)
    assert np.array_equal(bars, orc.gen_bars(4, 25, 0.0))
    assert np.allclose(bars[:, :10], np.array(g["rows_head"]))
    assert np.allclose(bars[:, 16:], np.array(g["rows_tail"]))
    assert np.allclose(b.gen_bars(10, 25, 0.1), orc.gen_bars(10, 25, 0.1))


def test_sparsify_sentence_and_flatten():
    s = b.sparsify_sentence([9, 8, 9, 6, 8, 9])
    assert s.w.tolist() == [9, 8, 6] and s.c.tolist() == [3, 2, 1]          # first-appearance order, src/common.jl:94-103
    md = b.MultinomialDirichlet(12, 1.2)
    fd = b.FlatData(md, [[b.Sent([3, 1, 3], [1, 2, 4])], [s, b.Sent([], [])]], grouped=True)
    assert fd.datum == 1 and fd.group_ptr.tolist() == [0, 1, 3]
    assert fd.sent_ptr.tolist() == [0, 2, 5, 5]
    assert fd.sent_word.tolist() == [1, 3, 6, 8, 9] and fd.sent_count.tolist() == [2, 5, 1, 2, 3]   # duplicates merged
    fd = b.FlatData(md, [np.array([1, 2]), np.array([], dtype=np.int64), np.array([5])], grouped=True)
    assert fd.datum == 0 and fd.word.tolist() == [1, 2, 5] and fd.group_ptr.tolist() == [0, 2, 2, 3]
    fd = b.FlatData(b.Gaussian1DGaussian1D(0, 1, 1), np.array([0.5, 1.5]), grouped=False)
    assert fd.datum == 2 and fd.group_ptr is None


def test_init_zz_ranges():
    rng = np.random.default_rng(0)
    lda = b.LDA(b.MultinomialDirichlet(5, 1.0), 7, 0.1)
    zz = [np.zeros(10, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(3, dtype=np.int64)]
    b.init_zz(lda, zz, rng)
    assert [len(z) for z in zz] == [10, 0, 3] and all(((z >= 0) & (z < 7)).all() for z in zz)
    bmm = b.BMM(b.MultinomialDirichlet(5, 1.0), 4, 0.1)
    z = np.zeros(100, dtype=np.int64)
    b.init_zz(bmm, z, rng)
    assert set(np.unique(z)) <= {0, 1, 2, 3} and len(np.unique(z)) == 4


def test_shard_bounds_cover_everything():
    from bias_jl_b200.sharding import shard_bounds, shard_csr
    for n, w in [(10, 3), (7, 8), (0, 2), (500000, 8)]:
        cuts = [shard_bounds(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(w - 1))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1
    ptr = np.array([0, 3, 3, 10, 12])
    lo, hi, i0, i1, lp = shard_csr(ptr, 1, 2)
    assert (lo, hi, i0, i1) == (2, 4, 3, 12) and lp.tolist() == [0, 7, 9]


@pytest.mark.timeout(400)
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on host cores only (the serial sampler, one instance per core) and prints one
    JSON line with the keys the driver reads."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=380, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lda_gibbs_token_samples_per_sec_K1024" and d["unit"] == "tokens/s"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
