"""Size-independent properties at BASELINE.json's full configuration sizes (the oracle cannot run these in
reasonable time): exact integer conservation of every count table against a recomputation from the returned
assignments, on 100 M tokens (C3), 1 M sentences (C2) and 10 M points (C4)."""
import numpy as np
import pytest

import bias_jl_b200 as b
from bias_jl_b200.sharding import device_tensor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b.Context(0, seed=97531)
    yield c
    c.close()


def test_c3_lda_100m_tokens_counts_conserved(ctx):
    import torch
    K, V, D, L = 1024, 50_000, 500_000, 200
    N = D * L
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    # Zipf-like word frequencies so that hot rows see heavy atomic contention
    w = (torch.rand(N, device=dev, generator=g) ** 3 * V).to(torch.int32).clamp_(0, V - 1)
    z0 = torch.randint(0, K, (N,), device=dev, dtype=torch.int32, generator=g)
    words_h, z_h = w.cpu().numpy(), z0.cpu().numpy()
    ptr = np.arange(D + 1, dtype=np.int64) * L
    lda = b.LDA(b.MultinomialDirichlet(V, 0.1 * V), K, 0.1)
    data = b.FlatData.from_csr(b._lib.INT, N, group_ptr=ptr, word=words_h)
    rs = b.ResidentSampler(ctx, lda, data, z_h)
    rs.sweep(2)
    s = rs.stats()
    assert 0 < s.n_moved <= N
    z = torch.from_numpy(rs.get_zz()).to(dev)
    assert int(z.min()) >= 0 and int(z.max()) < K
    # the live statistics as one flat int32 device buffer [n_wk | n_k | n_items] (zero base => delta == live)
    rs.delta_begin_zero()
    (p32, n32, _), = rs.delta_pack()
    live = device_tensor(p32, n32, "int32", dev)
    Kpad = 1024
    n_wk = live[:V * Kpad].view(V, Kpad)
    n_k = live[V * Kpad:V * Kpad + Kpad]
    ref = torch.bincount(w.to(torch.int64) * Kpad + z.to(torch.int64), minlength=V * Kpad).view(V, Kpad)
    assert torch.equal(n_wk.to(torch.int64), ref)
    assert torch.equal(n_k.to(torch.int64), torch.bincount(z.to(torch.int64), minlength=Kpad))
    assert int(n_k.sum()) == N
    ll, n = rs.log_likelihood()
    assert n == N and np.isfinite(ll) and ll < 0
    rs.close()


def test_c2_bmm_1m_sentences_counts_conserved(ctx):
    K, V, L, n_sent = 64, 1024, 64, 1_000_000
    sp, sw, sc, _, _ = b.data.bars_bmm_sentences_csr(n_sent, L, K, V, seed=123)
    bmm = b.BMM(b.MultinomialDirichlet(V, 1.0), K, 0.1)
    data = b.FlatData.from_csr(b._lib.SENT, n_sent, sent_ptr=sp, sent_word=sw, sent_count=sc)
    z0 = np.random.default_rng(0).integers(0, K, n_sent).astype(np.int32)
    rs = b.ResidentSampler(ctx, bmm, data, z0)
    rs.sweep(3)
    z = rs.get_zz()
    comp = rs.components()
    assert np.array_equal(comp["nn"], np.bincount(z, minlength=K))
    item = np.repeat(np.arange(n_sent), np.diff(sp))
    mi = np.zeros((K, V), dtype=np.int64)
    np.add.at(mi, (z[item], sw), sc)
    assert np.array_equal(comp["mi"], mi) and comp["mm"].sum() == n_sent * L
    rs.close()


def test_c4_hdp_10m_points_counts_conserved(ctx):
    rng = np.random.default_rng(123)
    n_groups, n_per, n_atoms = 100_000, 100, 20
    ks = rng.integers(0, n_atoms, (n_groups, 4))
    pick = rng.integers(0, 4, (n_groups, n_per))
    zt = np.take_along_axis(ks, pick, axis=1)
    x = (2.0 * (zt + 1) + np.sqrt(0.001) * rng.standard_normal(zt.shape)).reshape(-1)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=256)
    ptr = np.arange(n_groups + 1, dtype=np.int64) * n_per
    data = b.FlatData.from_csr(b._lib.F64, len(x), group_ptr=ptr, x=x)
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(len(x), dtype=np.int32), sample_hyperparam=True, n_internals=5)
    rs.sweep(6)
    hp, beta = rs.hdp_state()
    z = rs.get_zz()
    assert set(np.unique(z)) == set(range(hp.KK)) and abs(beta.sum() - 1.0) < 1e-9
    comp = rs.components()
    assert np.array_equal(comp["nn"], np.bincount(z, minlength=hp.KK))
    sums = np.bincount(z, weights=x, minlength=hp.KK)
    assert np.allclose(comp["mu"] * comp["nn"], sums, rtol=1e-9)
    assert n_atoms <= hp.KK <= n_atoms + 6
    # every inferred component is pure in the generating atoms
    homog = sum(np.bincount(zt.reshape(-1)[z == c]).max() for c in range(hp.KK)) / len(z)
    assert homog > 0.99
    rs.close()
