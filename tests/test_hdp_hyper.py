"""sample_hyperparam!(hdp, n_group_j, m) (src/HDP.jl:124-159) and the my_beta draw (src/HDP.jl:362-364).

The host half of the update (alpha0 ~ Gamma, gamma by Escobar-West, beta ~ Dirichlet) is product code that needs no
device: it is compared in distribution with the oracle's restatement and with the closed forms here on the CPU.  The
device half (the O(groups) Beta / Bernoulli auxiliaries, hdp_hyper_kernel) is in the gpu-marked tests below:
sum_j s_j exactly for supplied uniforms, sum_j log w_j against the digamma closed form and single draws against the
Beta CDF.  Tolerances: Kolmogorov-Smirnov p > 1e-3 over 4,000 draws (fixed seeds: deterministic), z-scores below 5."""
import numpy as np
import pytest
from scipy import special, stats

import bias_jl_b200 as b
from bias_jl_b200 import models
from oracle import oracle as orc

N = 4000
PRI = dict(a1=0.7, a2=1.3, g1=2.5, g2=0.8)


def test_alpha0_draws_are_the_gamma_of_teh_eq_50():
    m, sum_logw, sum_s, KK = 57.0, -23.4, 11.0, 9
    aa, _ = models.hdp_hyper_draws(5, PRI["a1"], PRI["a2"], PRI["g1"], PRI["g2"], m, sum_logw, sum_s, KK, 1.9, 3.3, N)
    shape, rate = PRI["a1"] + m - sum_s, PRI["a2"] - sum_logw                       # src/HDP.jl:146-148
    assert stats.kstest(aa, stats.gamma(shape, scale=1.0 / rate).cdf).pvalue > 1e-3
    assert abs(aa.mean() - shape / rate) < 5 * np.sqrt(shape) / rate / np.sqrt(N)


@pytest.mark.parametrize("KK,m,gg", [(9, 57, 3.3), (1, 4, 0.4), (40, 900, 12.0)])
def test_gamma_draws_match_the_oracle_in_distribution(KK, m, gg):
    """Escobar-West update (src/HDP.jl:151-158): two-sample KS against the oracle's sample_hyperparam! called with no
    groups (then the alpha0 part has sum log w = sum s = 0 and the gamma part is what is compared), plus the exact
    mixture mean given eta averaged over the library's own draws is not needed: the distributions must coincide."""
    _, g_lib = models.hdp_hyper_draws(7, PRI["a1"], PRI["a2"], PRI["g1"], PRI["g2"], float(m), 0.0, 0.0, KK, 1.9, gg, N)
    a_lib, _ = models.hdp_hyper_draws(8, PRI["a1"], PRI["a2"], PRI["g1"], PRI["g2"], float(m), 0.0, 0.0, KK, 1.9, gg, N)
    rng = orc.Rng(123)
    ref = np.array([orc.hdp_sample_hyperparam(KK, gg, PRI["g1"], PRI["g2"], 1.9, PRI["a1"], PRI["a2"], [0], m, rng) for _ in range(N)])
    assert stats.ks_2samp(g_lib, ref[:, 1]).pvalue > 1e-3
    assert stats.ks_2samp(a_lib, ref[:, 0]).pvalue > 1e-3
    assert (g_lib > 0).all()


def test_full_update_matches_the_oracle_in_distribution_with_groups():
    """The whole of sample_hyperparam! with groups: the oracle draws its own w_j, s_j; the library side is fed auxiliaries
    drawn from their closed-form laws with numpy (the device kernel that draws them is tested below)."""
    sizes = np.array([3, 50, 1, 200, 17, 80, 5, 5, 120, 9])
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    aa0, gg0, KK, m = 2.2, 1.7, 6, 31
    rng = orc.Rng(77)
    ref = np.array([orc.hdp_sample_hyperparam(KK, gg0, PRI["g1"], PRI["g2"], aa0, PRI["a1"], PRI["a2"], ptr, m, rng) for _ in range(N)])
    r = np.random.default_rng(3)
    got = np.zeros((N, 2))
    for i in range(N):
        logw = np.log(r.beta(aa0 + 1.0, sizes)).sum()
        s = (r.random(len(sizes)) < sizes / (sizes + aa0)).sum()
        a, g = models.hdp_hyper_draws(1000 + i, PRI["a1"], PRI["a2"], PRI["g1"], PRI["g2"], float(m), float(logw), float(s), KK, aa0, gg0, 1)
        got[i] = a[0], g[0]
    assert stats.ks_2samp(got[:, 0], ref[:, 0]).pvalue > 1e-3
    assert stats.ks_2samp(got[:, 1], ref[:, 1]).pvalue > 1e-3


def test_invalid_gamma_parameters_are_errors_not_skipped_updates():
    """Distributions.Gamma(shape <= 0) throws in the reference; the library reports BIAS_EINVAL instead of keeping the old value."""
    with pytest.raises(b.BiasError) as e:
        models.hdp_hyper_draws(1, 0.5, 1.0, 1.0, 1.0, 3.0, -1.0, 9.0, 4, 1.0, 1.0, 1)       # a1 + m - sum s < 0
    assert e.value.code == 1
    with pytest.raises(b.BiasError):
        models.hdp_hyper_draws(1, 0.5, 1.0, -5.0, 1.0, 3.0, -1.0, 1.0, 4, 1.0, 1.0, 1)      # g1 + KK - 1 < 0


def test_beta_draws_are_dirichlet():
    """my_beta ~ Dirichlet([sum_j M_jk ..., gg]) (src/HDP.jl:362-364): marginals are Beta(a_k, a_0 - a_k), rows sum to one."""
    tables = np.array([40, 3, 1, 17, 250, 8])
    gg = 1.6
    x = models.hdp_beta_draws(11, tables, gg, N)
    assert x.shape == (N, 7) and np.allclose(x.sum(1), 1.0, atol=1e-12) and (x >= 0).all()
    a = np.concatenate([tables, [gg]]).astype(float)
    a0 = a.sum()
    for k in range(7):
        assert stats.kstest(x[:, k], stats.beta(a[k], a0 - a[k]).cdf).pvalue > 1e-3, k
    cov = np.cov(x[:, 0], x[:, 4])[0, 1]
    want = -a[0] * a[4] / (a0 ** 2 * (a0 + 1))
    assert abs(cov - want) < 0.2 * abs(want)
    # oracle: the same Dirichlet through its own generator
    rng = orc.Rng(5)
    g = np.array([[rng.gamma(ak) for ak in a] for _ in range(N)])
    ref = g / g.sum(1, keepdims=True)
    for k in (0, 2, 6):
        assert stats.ks_2samp(x[:, k], ref[:, k]).pvalue > 1e-3, k


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=2468)
    yield c
    c.close()


@pytest.mark.gpu
def test_bernoulli_auxiliaries_exact_for_supplied_uniforms(ctx):
    r = np.random.default_rng(1)
    sizes = r.integers(0, 400, 50_000)
    sizes[::7] = 0                                          # empty groups draw nothing (src/HDP.jl:136 would see Beta(a, 0))
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    u = r.random(len(sizes))
    for aa in (0.3, 4.5):
        _, s = models.hdp_hyper_aux(ctx, ptr, aa, 3, 1, uniforms=u)
        p = sizes / (sizes + aa)
        assert s == float(((u < p) & (sizes > 0)).sum())


@pytest.mark.gpu
def test_log_w_auxiliaries_follow_beta(ctx):
    """sum_j log w_j over 200,000 groups against E[log Beta(aa + 1, n)] = psi(aa + 1) - psi(aa + 1 + n) (CLT, |z| < 5,
    8 independent Philox streams), sum_j s_j against its binomial law, and 2,000 single draws against the Beta CDF."""
    r = np.random.default_rng(2)
    sizes = r.integers(1, 300, 200_000)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    aa = 1.7
    mean = (special.digamma(aa + 1) - special.digamma(aa + 1 + sizes)).sum()
    var = (special.polygamma(1, aa + 1) - special.polygamma(1, aa + 1 + sizes)).sum()
    p = sizes / (sizes + aa)
    zs = []
    for it in range(8):
        lw, s = models.hdp_hyper_aux(ctx, ptr, aa, sweep=it, internal=it % 3)
        zs.append((lw - mean) / np.sqrt(var))
        assert abs(s - p.sum()) < 5 * np.sqrt((p * (1 - p)).sum())
    assert np.abs(zs).max() < 5 and len(set(np.round(zs, 9))) == 8          # in law, and the streams differ
    n = 37
    w = np.exp([models.hdp_hyper_aux(ctx, [0, n], aa, sweep=it, internal=0)[0] for it in range(2000)])
    assert stats.kstest(w, stats.beta(aa + 1, n).cdf).pvalue > 1e-3
