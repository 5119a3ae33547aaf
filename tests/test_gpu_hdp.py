"""GPU parity tests for the HDP direct-assignment sampler (reference src/HDP.jl:166-399):
Stirling table, table-count conditionals, K+1 assignment conditionals, and whole sweeps."""
import numpy as np
import pytest

import bias_jl_b200 as b
from bias_jl_b200.models import hdp_table_conditionals, stirling_rows
from oracle import oracle as orc
from helpers import assert_conditionals_close, assert_draws_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=4321)
    yield c
    c.close()


def test_stirling_table_matches_oracle(ctx):
    n = 300
    dev = stirling_rows(ctx, n)
    ref = orc.stirling_table(n)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(dev), fin)
    assert np.allclose(dev[fin], ref[fin], rtol=1e-12, atol=1e-9)
    row5 = np.exp(dev[10:15])
    assert np.allclose(row5 / row5.max(), np.array([24, 50, 35, 10, 1]) / 50.0, rtol=1e-12)
    # every row is max-normalised
    for r in (1, 2, 17, 300):
        assert dev[r * (r - 1) // 2: r * (r + 1) // 2].max() == 0.0


def test_table_count_conditionals_match_oracle(ctx):
    rng = np.random.default_rng(0)
    nmax = 400
    ref_tab = orc.stirling_table(nmax)
    n = np.concatenate([[1, 2, 3, nmax], rng.integers(1, nmax, 60)]).astype(np.int32)
    a = np.concatenate([[0.5, 3.0, 1e-3, 2.5], rng.gamma(1.0, 2.0, 60) + 1e-3])
    u = rng.random(len(n))
    prob, draws = hdp_table_conditionals(ctx, n, a, u)
    for q in range(len(n)):
        m_ref, p_ref = orc.hdp_table_conditional(ref_tab, int(n[q]), float(a[q]), float(u[q]))
        assert np.allclose(prob[q, :n[q]], p_ref, rtol=1e-9, atol=1e-300)
        assert (prob[q, n[q]:] == 0).all()
        if draws[q] != m_ref:
            assert np.min(np.abs(np.cumsum(p_ref) - u[q])) < 1e-9
    with pytest.raises(b.BiasError) as e:
        hdp_table_conditionals(ctx, [10001], [1.0], [0.5])          # the reference's table has 10k rows: BoundsError
    assert e.value.code == 5


def test_production_table_pass_matches_the_conditionals_in_distribution(ctx):
    """hdp_tables_kernel (one thread per (group, component) cell up to 128 items, one warp per larger cell) against the
    exact expectation of M_jk under the oracle's conditional (src/HDP.jl:348-361): sums over all groups, 120 passes."""
    from bias_jl_b200.sharding import device_tensor
    rng = np.random.default_rng(5)
    KK, G = 5, 600
    lens = rng.integers(1, 90, G)
    lens[:6] = [400, 300, 260, 1, 2, 129]                     # cells beyond the per-thread limit go to the warp kernel
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    N = int(ptr[-1])
    z0 = rng.integers(0, KK, N).astype(np.int32)
    z0[:400] = 2                                              # one cell of 400 items
    x = 2.0 * (z0 + 1) + 0.05 * rng.standard_normal(N)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.01), KK, 1.0, 0.1, 0.1, 1.5, 0.1, 0.1, Kmax=16)
    xx = [x[ptr[j]:ptr[j + 1]] for j in range(G)]
    rs = b.ResidentSampler(ctx, hdp, b.FlatData(hdp.component, xx, True), z0, sample_hyperparam=False, n_internals=1)
    hp, beta = rs.hdp_state()
    nn = rs.group_counts()
    tab = orc.stirling_table(int(nn.max()))
    expect, var = np.zeros(KK), np.zeros(KK)
    cache = {}
    for j in range(G):
        for k in range(KK):
            n = int(nn[j, k])
            if n == 0:
                continue
            if (n, k) not in cache:
                _, p = orc.hdp_table_conditional(tab, n, float(hp.aa * beta[k]), 0.5)
                m = np.arange(1, n + 1)
                mu = float((m * p).sum())
                cache[(n, k)] = (mu, float((m * m * p).sum()) - mu * mu)
            expect[k] += cache[(n, k)][0]
            var[k] += cache[(n, k)][1]
    reps = 120
    tot = np.zeros(KK)
    for hh in range(reps):
        (pt, n_t, _), _ = rs.hdp_shard_tables(hh)
        ctx.sync()
        t = device_tensor(pt, n_t, "int64", "cuda:0").cpu().numpy()
        assert t[-1] == t[:KK].sum() and (t[KK:-1] == 0).all()
        tot += t[:KK]
    mean = tot / reps
    assert (np.abs(mean - expect) < 5.0 * np.sqrt(var / reps) + 1e-9).all(), (mean, expect, np.sqrt(var / reps))
    rs.close()


def _hdp_case(kind, seed):
    rng = np.random.default_rng(seed)
    KK, Kmax = 4, 16
    ptr = np.array([0, 30, 30, 75, 120], dtype=np.int64)
    N = int(ptr[-1])
    if kind == "f64":
        x = np.repeat([2.0, 4.0, 6.0, 8.0], 30) + 0.05 * rng.standard_normal(N)
        rng.shuffle(x)
        comp = b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.01)
        odata = orc.Data.reals(x)
        ocomp = orc.Comps.g1g1(Kmax, float(x.mean()), 10.0, 0.01)
        xx = [x[ptr[j]:ptr[j + 1]] for j in range(4)]
    elif kind == "int":
        V = 25
        w = rng.integers(0, V, N)
        comp = b.MultinomialDirichlet(V, 2.5)
        odata = orc.Data.words(w)
        ocomp = orc.Comps.md(Kmax, V, 2.5)
        xx = [w[ptr[j]:ptr[j + 1]] for j in range(4)]
    else:
        V = 25
        sents, true_zz, _ = b.data.bars_bmm_sentences(N, 12, 4, V, seed)
        comp = b.MultinomialDirichlet(V, 2.5)
        sp = np.concatenate([[0], np.cumsum([len(s) for s in sents])])
        odata = orc.Data.sents(sp, np.concatenate([s.w for s in sents]), np.concatenate([s.c for s in sents]))
        ocomp = orc.Comps.md(Kmax, V, 2.5)
        xx = [sents[ptr[j]:ptr[j + 1]] for j in range(4)]
    z0 = rng.integers(0, KK, N).astype(np.int32)
    hdp = b.HDP(comp, KK, 1.0, 0.1, 0.1, 1.5, 0.1, 0.1, Kmax=Kmax)
    return hdp, xx, odata, ocomp, ptr, z0


@pytest.mark.parametrize("kind", ["f64", "int", "sent"])
def test_assignment_conditionals_match_oracle(ctx, kind):
    hdp, xx, odata, ocomp, ptr, z0 = _hdp_case(kind, 3)
    data = b.FlatData(hdp.component, xx, True)
    rs = b.ResidentSampler(ctx, hdp, data, z0)
    st = orc.HDPState(ocomp, odata, ptr, z0, hdp.KK, hdp.gg, hdp.g1, hdp.g2, hdp.aa, hdp.a1, hdp.a2)
    N = data.n_items
    rng = np.random.default_rng(8)
    u = rng.random(N)
    raw, prob, draw = rs.conditionals(np.arange(N), u)
    assert raw.shape == (N, hdp.KK + 1)
    grp = np.searchsorted(ptr, np.arange(N), side="right") - 1
    oraw, oprob, odraw = np.zeros_like(raw), np.zeros_like(prob), np.zeros(N, dtype=np.int64)
    for i in range(N):
        odraw[i], oraw[i], oprob[i] = st.conditional(int(grp[i]), i, float(u[i]))
    assert_conditionals_close(raw, oraw, f"HDP {kind}")
    assert np.allclose(raw, oraw, rtol=1e-9, atol=1e-8)
    assert_draws_match(draw, odraw, oprob, u, 1e-9, "HDP verification draws")
    rs.close()


@pytest.mark.parametrize("kind", ["f64", "int", "sent"])
def test_sweeps_keep_state_consistent(ctx, kind):
    hdp, xx, odata, ocomp, ptr, z0 = _hdp_case(kind, 5)
    data = b.FlatData(hdp.component, xx, True)
    rs = b.ResidentSampler(ctx, hdp, data, z0, sample_hyperparam=True, n_internals=5)
    for _ in range(8):
        rs.sweep(1)
        s = rs.stats()
        hp, beta = rs.hdp_state()
        z = rs.get_zz()
        assert s.KK == hp.KK >= 1 and ((z >= 0) & (z < hp.KK)).all()
        assert set(np.unique(z)) == set(range(hp.KK))                 # every surviving component is non-empty
        assert len(beta) == hp.KK + 1 and abs(beta.sum() - 1.0) < 1e-9 and (beta > 0).all()
        assert hp.gg > 0 and hp.aa > 0
        comp = rs.components()
        assert np.array_equal(comp["nn"], np.bincount(z, minlength=hp.KK))
        nn = rs.group_counts()
        for j in range(len(ptr) - 1):
            assert np.array_equal(nn[j], np.bincount(z[ptr[j]:ptr[j + 1]], minlength=hp.KK))
    if kind != "f64":
        V = hdp.component.dd
        mi = np.zeros((hp.KK, V), dtype=np.int64)
        if kind == "int":
            np.add.at(mi, (z, np.asarray(odata.w)), 1)
        else:
            item = np.repeat(np.arange(data.n_items), np.diff(odata.sptr))
            np.add.at(mi, (z[item], odata.sw), odata.sc)
        assert np.array_equal(comp["mi"], mi)
    rs.close()


def test_hdp_gaussian_finds_the_number_of_atoms(ctx):
    """demo_HDP_Gaussian1DGaussian1D.jl: CRF-generated groups, atoms N(2k, 0.001); start from KK=1."""
    zz_true, n_atoms = b.data.gen_crf_data([100] * 30, 1.0, 0.5, seed=3)
    rng = np.random.default_rng(4)
    xx = [2.0 * (z + 1) + np.sqrt(0.001) * rng.standard_normal(len(z)) for z in zz_true]
    allx = np.concatenate(xx)
    make = lambda: b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=64)
    hdp = make()
    zz = [np.zeros(len(x), dtype=np.int64) for x in xx]
    KK_list, KK_dict, betas, gammas, alphas = b.collapsed_gibbs_sampler(hdp, xx, zz, 20, 1, 10, True, 5, ctx=ctx)
    assert len(KK_list) == 10 and len(betas) == 10 and hdp.KK == KK_list[-1]
    used = np.unique(np.concatenate(zz_true))
    assert abs(hdp.KK - len(used)) <= 2, (hdp.KK, len(used))
    # the serial reference sampler on the same data lands on a similar number of components
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])])
    Kmax = 64
    st = orc.HDPState(orc.Comps.g1g1(Kmax, float(allx.mean()), 10.0, 0.001), orc.Data.reals(allx), ptr,
                      np.zeros(len(allx), dtype=np.int64), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1)
    r = orc.Rng(5)
    tab = orc.stirling_table(128)
    for _ in range(15):
        assert not np.isnan(st.sweep(r))
        st.resample_beta(tab, 128, 5, True, r)
    # (the serial chain tends to carry a few more small components this early; the parallel one must not over-create)
    assert st.KK >= len(used) - 2 and hdp.KK <= st.KK + 2, (hdp.KK, st.KK, len(used))
    # every inferred component is (almost) pure in the generating labels; duplicates of one atom are
    # allowed this early in the chain, merged atoms are not
    z = np.concatenate(zz)
    zt = np.concatenate(zz_true)
    homogeneity = sum(np.bincount(zt[z == c]).max() for c in np.unique(z)) / len(z)
    assert homogeneity > 0.97, homogeneity


def test_kmax_exhaustion_is_an_error(ctx):
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 * k + 0.01 * rng.standard_normal(40) for k in range(6)])
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 100.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=3)
    with pytest.raises(b.BiasError) as e:
        b.collapsed_gibbs_sampler(hdp, [x], [np.zeros(len(x), dtype=np.int64)], 5, 0, 1, ctx=ctx)
    assert e.value.code == 5


@pytest.mark.parametrize("n_shards", [2, 3])
def test_sharded_hdp_on_one_gpu(ctx, n_shards):
    """The N>1 HDP protocol (sharding.ShardedHDP) with the ranks played by several samplers on one device: replicas
    of the component statistics, beta, alpha0, gamma and KK stay identical, counts equal the recomputation from the
    gathered zz, and the atoms are recovered as well as by the single-GPU sampler."""
    import torch
    from bias_jl_b200.sharding import LocalComm, ShardedHDP, shard_csr
    zz_true, n_atoms = b.data.gen_crf_data([80] * 36, 1.0, 0.5, seed=5)
    rng = np.random.default_rng(6)
    xx = [2.0 * (z + 1) + np.sqrt(0.001) * rng.standard_normal(len(z)) for z in zz_true]
    allx = np.concatenate(xx)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])]).astype(np.int64)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(allx.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=64)
    ctxs, samplers = [], []
    for r in range(n_shards):
        lo, hi, i0, i1, lptr = shard_csr(ptr, r, n_shards)
        data = b.FlatData.from_csr(b._lib.F64, i1 - i0, group_ptr=lptr, x=allx[i0:i1])
        c = b.Context(0, seed=900 + r)
        ctxs.append(c)
        samplers.append(b.ResidentSampler(c, hdp, data, np.zeros(i1 - i0, dtype=np.int32), sample_hyperparam=True,
                                          n_internals=5))
    sh = ShardedHDP(samplers, LocalComm(n_shards), torch.device("cuda", 0), host_seed=31)
    for sweep in range(25):
        sh.sweep(1)
        states = [rs.hdp_state() for rs in samplers]
        KK = states[0][0].KK
        z = np.concatenate([rs.get_zz() for rs in samplers])
        assert ((z >= 0) & (z < KK)).all() and set(np.unique(z)) == set(range(KK))
        for rs, (hp, beta) in zip(samplers, states):
            assert hp.KK == KK and hp.gg == states[0][0].gg and hp.aa == states[0][0].aa
            assert np.array_equal(beta, states[0][1]) and abs(beta.sum() - 1.0) < 1e-9
            comp = rs.components()
            assert np.array_equal(comp["nn"], np.bincount(z, minlength=KK))
            for k in range(KK):
                assert abs(comp["mu"][k] - allx[z == k].mean()) < 1e-9
    used = len(np.unique(np.concatenate(zz_true)))
    # the ranks take turns giving birth, so two ranks that meet the same new atom in one sweep do not both open a
    # component for it: KK stays where the single-GPU sampler puts it
    assert used - 1 <= KK <= used + 5, (KK, used)
    zt = np.concatenate(zz_true)
    homogeneity = sum(np.bincount(zt[z == c]).max() for c in np.unique(z)) / len(z)
    assert homogeneity > 0.97, homogeneity
    for rs in samplers:
        rs.close()
    for c in ctxs:
        c.close()


def test_sharded_hdp_multinomial_counts_stay_consistent(ctx):
    import torch
    from bias_jl_b200.sharding import LocalComm, ShardedHDP, shard_csr
    V = 25
    xx, _, _ = b.data.bars_lda_corpus(n_groups=40, n_tokens=60, seed=3)
    words = np.concatenate(xx)
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in xx])]).astype(np.int64)
    hdp = b.HDP(b.MultinomialDirichlet(V, 2.5), 2, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=48)
    z0 = np.random.default_rng(1).integers(0, 2, len(words)).astype(np.int32)
    samplers, ctxs = [], []
    for r in range(2):
        lo, hi, i0, i1, lptr = shard_csr(ptr, r, 2)
        data = b.FlatData.from_csr(b._lib.INT, i1 - i0, group_ptr=lptr, word=words[i0:i1])
        c = b.Context(0, seed=50 + r)
        ctxs.append(c)
        samplers.append(b.ResidentSampler(c, hdp, data, z0[i0:i1], sample_hyperparam=True, n_internals=3))
    sh = ShardedHDP(samplers, LocalComm(2), torch.device("cuda", 0), host_seed=7)
    sh.sweep(30)
    KK = samplers[0].hdp_state()[0].KK
    z = np.concatenate([rs.get_zz() for rs in samplers])
    mi = np.zeros((KK, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    for rs in samplers:
        comp = rs.components()
        assert np.array_equal(comp["mi"], mi) and np.array_equal(comp["mm"], mi.sum(1))
    assert 1 <= KK <= 46, KK
    for rs in samplers:
        rs.close()
    for c in ctxs:
        c.close()


def test_sharded_hdp_over_nccl_two_processes():
    """The same protocol with real ranks: two processes, two GPUs, NCCL (tools/hdp_sharded_check.py)."""
    import os
    import subprocess
    import sys
    if b.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HDP_GROUPS="4000")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "hdp_sharded_check.py")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["replicas_consistent"] and line["homogeneity"] > 0.97 and 20 <= line["KK"] <= 25
