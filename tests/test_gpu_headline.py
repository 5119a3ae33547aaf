"""Parity of the PRODUCTION LDA kernel (fp32, narrow tables) at the headline shape, K = 1024.

* the production sweep itself, token by token, against the serial oracle on one long document (the per-document row r
  is only ever derived from its previous value inside a sweep: this is where drift would show);
* fp32 conditionals and draws on the BASELINE config-3 state (100 M tokens, n_k ~ 1e5, 1024-term sums) against the
  fp64 verification kernel, which tests/test_gpu_lda.py pins to the oracle at small sizes;
* hot words (corpus frequency above 65,535) stay inside the same kernel on int32 rows.
"""
import numpy as np
import pytest

import bias_jl_b200 as b
from oracle import oracle as orc
from helpers import assert_draws_match

pytestmark = pytest.mark.gpu

EPS_FP32 = 2e-5      # a production (fp32) draw may differ from the Float64 one only when u is this close to a CDF edge


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=4242)
    yield c
    c.close()


@pytest.mark.parametrize("K,V,L,n_sweeps", [(1024, 2000, 60_000, 5), (1024, 300, 3_000, 6), (96, 50, 20_000, 4), (10, 25, 9_000, 4)])
def test_production_sweep_follows_the_serial_sampler_on_one_long_document(ctx, K, V, L, n_sweeps):
    """One document => the GPU sweep is as serial as src/LDA.jl:116-131: with the same uniforms every token must get
    the oracle's topic, except where u sits within EPS_FP32 of an edge of the oracle's CDF.  The oracle follows the
    GPU's choice at such a token, so one near-tie does not end the comparison.  L = 60,000 at K = 1024 is the longest
    document class this kernel accepts (< 65,536 tokens); thousands of moves rewrite the same slots of r."""
    rng = np.random.default_rng(K + L)
    # a few dominant topics and words, as a real long document has
    theta = rng.dirichlet(np.full(K, 0.05))
    phi_w = rng.dirichlet(np.full(V, 0.1), size=8)
    words = np.concatenate([rng.choice(V, L // 8, p=phi_w[i]) for i in range(8)] + [rng.integers(0, V, L - 8 * (L // 8))])
    words = rng.permutation(words).astype(np.int64)
    z0 = rng.choice(K, L, p=theta).astype(np.int32)
    alpha, aa = 0.1, 0.1 * V
    lda = b.LDA(b.MultinomialDirichlet(V, aa), K, alpha)
    data = b.FlatData(lda.component, [words], True)
    rs = b.ResidentSampler(ctx, lda, data, z0)
    st = orc.LDAState(orc.Comps.md(K, V, aa), orc.Data.words(words), data.group_ptr, z0, alpha, with_diag=False)
    n_ties = 0
    for s in range(n_sweeps):
        u = rng.random(L)
        rs.sweep_with_uniforms(u)
        z = rs.get_zz()
        n_diff, worst = st.sweep_follow(u, z.astype(np.int64))
        assert worst <= EPS_FP32, f"sweep {s}: a GPU draw differs from the oracle's with u {worst:.2e} away from any CDF edge"
        assert n_diff <= max(3, L // 300), f"sweep {s}: {n_diff} near-tie disagreements in {L} tokens"
        assert np.array_equal(st.z, z)
        n_ties += n_diff
    if rs.narrow_info()[2] > 0:
        assert rs.narrow_info()[0], "the sweeps were expected to run on the narrow tables"
    # the tables after the sweeps are the oracle's
    comp = rs.components()
    assert np.array_equal(comp["mi"], st.c.mi) and np.array_equal(comp["mm"], st.c.mi.sum(1))
    rs.close()


def test_hot_words_run_on_int32_rows_inside_the_16_bit_kernel(ctx):
    """Words above 65,535 occurrences get int32 rows next to the 16-bit rows of all other words and the SAME kernel
    launch serves both (per-token routing): counts are conserved exactly with cells far beyond 16 bits, frozen draws
    match the oracle on such a state, and the narrow tables stay current (no fall-back to the int32 table)."""
    K, V = 3, 9
    rng = np.random.default_rng(8)
    words = rng.permutation(np.concatenate([np.zeros(280_000, dtype=np.int64), np.full(70_000, 4, dtype=np.int64),
                                            rng.integers(1, V, 30_000)]))
    cut = np.sort(rng.choice(np.arange(1, len(words)), 600, replace=False))
    xx = [a for a in np.split(words, cut)]
    lda = b.LDA(b.MultinomialDirichlet(V, 1.2), K, 0.5)
    data = b.FlatData(lda.component, xx, True)
    z0 = rng.integers(0, K, data.n_items).astype(np.int32)
    rs = b.ResidentSampler(ctx, lda, data, z0)
    for _ in range(4):
        rs.sweep(1)
        assert rs.stats().n_moved > 0
    cur, n_hot, cap = rs.narrow_info()
    assert cur and n_hot == 2 and cap >= 2, (cur, n_hot, cap)
    z = rs.get_zz()
    comp = rs.components()
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    assert mi.max() > 65535
    assert np.array_equal(comp["mi"], mi) and np.array_equal(comp["mm"], mi.sum(1))
    u = rng.random(data.n_items)
    idx = rng.choice(data.n_items, 400, replace=False)
    fd = rs.frozen_draws(u)
    st = orc.LDAState(orc.Comps.md(K, V, 1.2), orc.Data.words(words), data.group_ptr, z, 0.5)
    grp = np.searchsorted(data.group_ptr, idx, side="right") - 1
    oprob = np.zeros((len(idx), K))
    odraw = np.zeros(len(idx), dtype=np.int64)
    for q, (i, j) in enumerate(zip(idx, grp)):
        odraw[q], _, oprob[q] = st.conditional(int(j), int(i), float(u[i]))
    assert_draws_match(fd[idx], odraw, oprob, u[idx], EPS_FP32)
    rs.sweep(2)                                   # and on after the verification calls
    z = rs.get_zz()
    mi[:] = 0
    np.add.at(mi, (z, words), 1)
    assert np.array_equal(rs.components()["mi"], mi)
    rs.close()


@pytest.mark.parametrize("K,V", [(1024, 500), (200, 64)])
def test_hot_and_cold_rows_share_a_warp_at_k1024(ctx, K, V):
    """Half of the tokens are hot words, so the two documents of a warp take different routes through the row loads and
    conversions most of the time (8 chunks, second half of the row streamed, selected chunk re-read)."""
    rng = np.random.default_rng(3)
    n_hot_tok, n_cold_tok = 210_000, 200_000
    hot_ids = rng.choice(3, n_hot_tok)                                   # words 0, 1, 2: 70k each
    words = rng.permutation(np.concatenate([hot_ids, rng.integers(3, V, n_cold_tok)])).astype(np.int64)
    xx = [words[i:i + 137] for i in range(0, len(words), 137)]
    alpha, aa = 0.1, 0.1 * V
    lda = b.LDA(b.MultinomialDirichlet(V, aa), K, alpha)
    data = b.FlatData(lda.component, xx, True)
    z0 = (rng.integers(0, K, data.n_items) // 7 * 7 % K).astype(np.int32)     # uneven start
    rs = b.ResidentSampler(ctx, lda, data, z0)
    rs.sweep(3)
    cur, n_hot, _ = rs.narrow_info()
    assert cur and n_hot == 3
    z = rs.get_zz()
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z, words), 1)
    comp = rs.components()
    assert np.array_equal(comp["mi"], mi) and np.array_equal(comp["mm"], mi.sum(1))
    # production draws on this state against the fp64 verification kernel
    u = rng.random(data.n_items)
    idx = rng.choice(data.n_items, 600, replace=False)
    _, prob, draw = rs.conditionals(idx, u[idx])
    fd = rs.frozen_draws(u)
    n_bad = assert_draws_match(fd[idx], draw, prob, u[idx], EPS_FP32)
    assert n_bad <= 3
    rs.close()


def _bisect_edges(rs, N, idx, edge, dev, iters=25):
    """CDF value of the production kernel at `edge[q]` for every sampled token idx[q]: bisection on the token's uniform
    (frozen counts: the draw is a step function of u; the step between topic edge and edge + 1 is cdf[edge])."""
    import torch
    u = torch.full((N,), 0.5, dtype=torch.float64, device=dev)
    d = torch.empty(N, dtype=torch.int32, device=dev)
    it, e = torch.as_tensor(idx, device=dev), torch.as_tensor(edge, device=dev, dtype=torch.int32)
    lo = torch.zeros(len(idx), dtype=torch.float64, device=dev)
    hi = torch.ones(len(idx), dtype=torch.float64, device=dev)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        u[it] = mid
        torch.cuda.synchronize()
        rs.frozen_draws_dev(u.data_ptr(), d.data_ptr())
        up = d[it] > e
        hi = torch.where(up, mid, hi)
        lo = torch.where(up, lo, mid)
    return (0.5 * (lo + hi)).cpu().numpy()


def test_c3_fp32_conditionals_on_the_config3_state(ctx):
    """BASELINE config 3 (K = 1024, V = 50,000, 100 M tokens, bench.py's corpus) after 30 sweeps: for 12,000 sampled
    tokens the production kernel's draws equal the fp64 verification kernel's except within EPS_FP32 of a CDF edge,
    its CDF (recovered black-box by bisection on the uniform) matches at every chunk boundary and at the edges
    around each token's most probable topics within 3e-6, and the probabilities of those topics match in log within
    1e-5 |log p| + 1e-6 plus the resolution of a 24-bit uniform (1.2e-7 / p)."""
    import torch
    import bench
    dev = torch.device("cuda:0")
    K, V, D, L = bench.K, 50_000, bench.D_TOTAL, bench.L_DOC
    N = D * L
    words_dev = bench.gen_docs_torch(bench.topic_cdf_torch(dev), 0, D, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    z_dev = torch.randint(0, K, (N,), device=dev, dtype=torch.int32, generator=g)
    words_h, z_h = words_dev.cpu().numpy(), z_dev.cpu().numpy()
    del words_dev, z_dev
    lda = b.LDA(b.MultinomialDirichlet(V, bench.BETA * V), K, bench.ALPHA)
    data = b.FlatData.from_csr(b._lib.INT, N, group_ptr=np.arange(D + 1, dtype=np.int64) * L, word=words_h)
    rs = b.ResidentSampler(ctx, lda, data, z_h)
    rs.sweep(30)
    assert rs.narrow_info()[0]
    rng = np.random.default_rng(17)
    nq = 12_000
    idx = np.sort(rng.choice(N, nq, replace=False))
    u_q = rng.random(nq)
    _, prob, draw64 = rs.conditionals(idx, u_q)                        # fp64, the reference's operation order
    cdf64 = np.cumsum(prob, axis=1)
    # (b) draws
    u = torch.rand(N, dtype=torch.float64, device=dev, generator=g)
    it = torch.as_tensor(idx, device=dev)
    u[it] = torch.as_tensor(u_q, device=dev)
    d = torch.empty(N, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    rs.frozen_draws_dev(u.data_ptr(), d.data_ptr())
    draw32 = d[it].cpu().numpy()
    n_bad = assert_draws_match(draw32, draw64, prob, u_q, EPS_FP32, "config-3 production draws")
    assert n_bad <= nq // 200, n_bad
    # (a) CDF at the chunk boundaries (the chunk-selection scan over 1024-term sums)
    for e in [127, 255, 383, 511, 639, 767, 895]:
        got = _bisect_edges(rs, N, idx, np.full(nq, e), dev)
        assert np.abs(got - cdf64[:, e]).max() < 3e-6, e
    # (a) probabilities of each token's two most probable topics, from the CDF edges on either side
    order = np.argsort(-prob, axis=1)
    worst = 0.0
    for rank in range(2):
        k = order[:, rank]
        hi_edge = _bisect_edges(rs, N, idx, k, dev)
        lo_edge = _bisect_edges(rs, N, idx, k - 1, dev)                 # edge -1: draw > -1 always -> 0
        hi_edge = np.where(k == K - 1, 1.0, hi_edge)
        p32 = hi_edge - lo_edge
        p64 = prob[np.arange(nq), k]
        assert np.abs(hi_edge - np.where(k == K - 1, 1.0, cdf64[np.arange(nq), k])).max() < 3e-6
        ok = p64 > 1e-2
        err = np.abs(np.log(p32[ok]) - np.log(p64[ok]))
        tol = 1e-5 * np.abs(np.log(p64[ok])) + 1e-6 + 2.5e-7 / p64[ok]
        worst = max(worst, float((err / tol).max()))
        assert (err <= tol).all(), (rank, float((err / tol).max()))
    assert ok.sum() > 200                                                 # the state has resolvable conditionals
    rs.close()
