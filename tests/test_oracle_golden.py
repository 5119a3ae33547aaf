"""Pins the CPU oracle against every RNG-independent known answer the reference
documents (tests/golden/reference_known_answers.json)."""
import math

import numpy as np

from oracle import oracle as orc


def _printed(value: float, printed: str) -> bool:
    """A Julia REPL transcript prints 6 decimals (rounded or truncated)."""
    nd = len(printed.split(".")[1])
    return abs(value - float(printed)) <= 10.0 ** (-nd) * 1.000001


def _g1g1_with(ctor, adds):
    c = orc.Comps.g1g1(1, *ctor)
    d = orc.Data.reals(np.array(adds + [0.0]))
    for i in range(len(adds)):
        orc.additem(c, 0, d, i)
    return c, d


def test_g1g1_running_mean(golden):
    g = golden["g1g1_running_mean"]
    c, d = _g1g1_with(g["ctor"], g["add"])
    assert c.nn[0] == g["nn_after_add"]
    assert c.mu[0] == 3.0999999999999996          # running-mean op order visible at 1 ulp
    assert _printed(c.mu[0], g["mu_after_add_printed"])
    assert orc.delitem(c, 0, d, 2) == 0
    assert c.nn[0] == g["nn_after_del"]
    assert c.mu[0] == 3.2499999999999996
    assert _printed(c.mu[0], g["mu_after_del_printed"])


def test_g1g1_delitem_edge_cases():
    c, d = _g1g1_with([2, 10, 0.5], [3.0])
    assert orc.delitem(c, 0, d, 0) == 0
    assert c.nn[0] == 0 and c.mu[0] == 2.0       # nn == 1 branch resets mu to m0, src/conjugates.jl:67-69
    assert orc.delitem(c, 0, d, 0) == -1         # the reference throws here, src/conjugates.jl:66


def test_g1g1_posterior_and_logpredictive(golden):
    g = golden["g1g1_posterior"]
    c, d = _g1g1_with(g["ctor"], g["add"])
    mu, vv = c.posterior(0)
    assert _printed(mu, g["mu_printed"]) and _printed(vv, g["vv_printed"])
    assert abs(mu - 3.0819672131147535) < 1e-15 and abs(vv - 0.16393442622950818) < 1e-16
    for case in golden["g1g1_logpredictive"]["cases"]:
        dd = orc.Data.reals(np.array([case["x"]]))
        lp = orc.logpredictive(c, 0, dd, 0)
        if case["typo"]:
            assert abs(lp - (-1.6555085890)) < 1e-9
        else:
            assert _printed(lp, case["printed"])
    # empty component: posterior is the prior (src/conjugates.jl:84-87)
    e = orc.Comps.g1g1(1, 2, 10, 0.5)
    assert e.posterior(0) == (2.0, 10.0)


def test_md_ctor_and_int_predictive(golden):
    g = golden["md_ctor"]
    c = orc.Comps.md(2, g["dd"], g["aa_in"])
    assert c.aa == g["aa_stored"]
    d = orc.Data.words(np.array([0, 0, 3]))
    for i in range(3):
        orc.additem(c, 0, d, i)
    assert list(c.mi[0]) == [2, 0, 0, 1, 0] and c.mm[0] == 3 and c.nn[0] == 3
    assert orc.logpredictive(c, 0, d, 0) == math.log((0.4 + 2) / (0.4 * 5 + 3))
    assert orc.logpredictive(c, 1, d, 0) == math.log(0.4 / 2.0)
    # diagnostic is the posterior MEAN, not a log (src/conjugates.jl:205)
    assert abs(orc.loglikelihood(c, 0, d, 2) - (1 + 0.4) / (3 + 2.0)) < 1e-15
    orc.delitem(c, 0, d, 0)
    assert list(c.mi[0]) == [1, 0, 0, 1, 0] and c.mm[0] == 2 and c.nn[0] == 2


def test_md_sent_predictive_matches_sequential_int_predictives():
    """A Sent's Dirichlet-multinomial predictive equals the multinomial coefficient times the
    chain of single-word predictives (each one conditioned on the previous words)."""
    rng = np.random.default_rng(0)
    V, K = 9, 3
    c = orc.Comps.md(K, V, 0.9)
    base = orc.Data.words(rng.integers(0, V, 40))
    for i in range(40):
        orc.additem(c, int(rng.integers(0, K)), base, i)
    sw, sc = np.array([4, 1, 7]), np.array([2, 1, 3])
    sent = orc.Data.sents(np.array([0, 3]), sw, sc)
    flat = orc.Data.words(np.repeat(sw, sc))
    for k in range(K):
        lp = orc.logpredictive(c, k, sent, 0)
        chain = 0.0
        for i in range(flat.n):
            chain += orc.logpredictive(c, k, flat, i)
            orc.additem(c, k, flat, i)
        for i in range(flat.n):
            orc.delitem(c, k, flat, i)
        coef = math.lgamma(6 + 1) - sum(math.lgamma(x + 1) for x in sc)
        assert abs(lp - (coef + chain)) < 1e-10
    # add/del of a Sent moves mm by the token total and nn by one (src/conjugates.jl:163-178)
    mm0, nn0 = int(c.mm[1]), int(c.nn[1])
    orc.additem(c, 1, sent, 0)
    assert c.mm[1] == mm0 + 6 and c.nn[1] == nn0 + 1
    orc.delitem(c, 1, sent, 0)
    assert c.mm[1] == mm0 and c.nn[1] == nn0


def test_gaussian1d_pdf(golden):
    g = golden["gaussian1d_pdf"]
    assert _printed(orc.gaussian1d_pdf(g["mu"], g["vv"], g["x"]), g["pdf_printed"])
    assert _printed(orc.gaussian1d_logpdf(g["mu"], g["vv"], g["x"]), g["logpdf_printed"])


def test_gen_bars(golden):
    g = golden["gen_bars_4_25"]
    bars = orc.gen_bars(4, 25, 0.0)
    assert np.allclose(bars[:, :10], np.array(g["rows_head"]), atol=1e-15)
    assert np.allclose(bars[:, 16:], np.array(g["rows_tail"]), atol=1e-15)
    assert np.allclose(bars.sum(1), 1.0)


def test_posterior_variance_closed_forms(golden):
    g = golden["bmm_gaussian_posterior_variances"]
    for n, pv in zip(g["nn"], g["post_vv"]):
        c = orc.Comps.g1g1(1, 0.0, g["v0"], g["vv"])
        c.nn[0] = n
        assert c.posterior(0)[1] == pv               # exact to the last printed digit
    g = golden["lda_gaussian_posterior_variances"]
    assert sum(g["implied_nn"]) == g["total"]
    for n, pv in zip(g["implied_nn"], g["post_vv"]):
        c = orc.Comps.g1g1(1, 0.0, g["v0"], g["vv"])
        c.nn[0] = n
        assert abs(c.posterior(0)[1] - pv) <= 1e-15 * pv * 4


def test_lognormalize_and_sample_semantics():
    pp = np.log(np.array([1.0, 2.0, 3.0, 4.0])) + 700.0       # would overflow without the max shift
    p = orc.lognormalize(pp)
    assert np.allclose(p, [0.1, 0.2, 0.3, 0.4], atol=1e-15)
    # strict `cw < r`: r exactly on a CDF edge stays in the lower bin (src/common.jl:74)
    w = np.array([0.25, 0.25, 0.5])
    assert orc.sample(w, 0.25) == 0 and orc.sample(w, np.nextafter(0.25, 1)) == 1
    assert orc.sample(w, 0.0) == 0 and orc.sample(w, 0.999999) == 2
    assert orc.sample(w, 1.5) == 2                            # clamps at n
    assert orc.sample(np.array([1.0]), 0.3) == 0


def test_julia_sum_pairwise_above_block():
    a = np.random.default_rng(1).random(5000)
    s = orc.julia_sum(a)
    assert abs(s - math.fsum(a)) < 1e-9
    assert orc.julia_sum(a[:7]) == ((((((a[0] + a[1]) + a[2]) + a[3]) + a[4]) + a[5]) + a[6])


def test_stirling_table_small_rows():
    t = orc.stirling_table(6)
    row5 = np.exp(t[10:15])
    assert np.allclose(row5 / row5.max(), np.array([24, 50, 35, 10, 1]) / 50.0, rtol=1e-12)
    row6 = np.exp(t[15:21])
    assert np.allclose(row6 / row6.max(), np.array([120, 274, 225, 85, 15, 1]) / 274.0, rtol=1e-12)


def test_table_conditional_matches_antoniak():
    """p(m | n, a) from the Stirling row equals the CRP table-count law s(n,m) a^m Gamma(a)/Gamma(a+n)."""
    nmax = 40
    t = orc.stirling_table(nmax)
    n, a = 17, 0.7
    m, prob = orc.hdp_table_conditional(t, n, a, 0.5)
    # brute force: distribution of the number of tables after n customers
    dist = np.zeros(n + 1)
    dist[1] = 1.0
    for i in range(1, n):
        new = np.zeros(n + 1)
        new[1:] += dist[:-1] * (a / (a + i))
        new += dist * (i / (a + i))
        dist = new
    assert np.allclose(prob, dist[1:], rtol=1e-10, atol=1e-300)
    assert 1 <= m <= n and prob[:m].sum() >= 0.5 > prob[:m - 1].sum()


def test_linear_domain_sweep_makes_the_same_draws_as_the_log_domain_restatement():
    """orc_lda_sweep_linear (convergence fixtures, tools/make_convergence_fixture.py) against orc_lda_sweep, the
    operation-by-operation restatement of src/LDA.jl:113-133: same uniforms, same randperm orders -> the same chain
    (a draw could only differ at a CDF tie of Float64 width)."""
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    for K, V, D, L in [(64, 300, 40, 100), (1024, 500, 12, 90), (10, 25, 30, 50)]:
        w = rng.integers(0, V, D * L)
        ptr = np.arange(0, D * L + 1, L)
        z0 = rng.integers(0, K, D * L)
        a = orc.LDAState(orc.Comps.md(K, V, 0.1 * V), orc.Data.words(w), ptr, z0, 0.1, with_diag=False)
        b_ = orc.LDAState(orc.Comps.md(K, V, 0.1 * V), orc.Data.words(w), ptr, z0, 0.1, with_diag=False)
        n_diff = 0
        for _ in range(4):
            u = rng.random(D * L)
            do = rng.permutation(D)
            to = np.concatenate([ptr[j] + rng.permutation(L) for j in do])
            a.sweep(u, doc_order=do, tok_order=to, diag_mode=0)
            b_.sweep_linear(u, doc_order=do, tok_order=to)
            n_diff += int((a.z != b_.z).sum())
            b_.z[:] = a.z                      # keep the states aligned should a tie have split them
            b_.c.mi[:] = a.c.mi
            b_.c.mm[:] = a.c.mm
            b_.c.nn[:] = a.c.nn
            b_.nn[:] = a.nn
        assert n_diff <= 1, (K, n_diff)
