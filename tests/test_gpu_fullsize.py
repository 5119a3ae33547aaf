"""Size-independent properties at BASELINE.json's full configuration sizes (the oracle cannot run these in
reasonable time): exact integer conservation of every count table against a recomputation from the returned
assignments, on 100 M tokens (C3), 1 M sentences (C2), 10 M points (C4, both HDP samplers) and the 100-epoch RCRP /
RCRF workloads (C5), plus the seating invariants of the franchise samplers."""
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


def _sums_agree(got, want):
    """Component sums of a Gaussian model are Float64 atomics that every item passed through: a component that held
    5e5 points (sum ~1e7) and shrank to five keeps the rounding of what passed through it, so the tolerance is
    relative to the largest sum, not to each component's own (the reference's running mean drifts the same way,
    src/conjugates.jl:62, :71)."""
    return np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())


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
    assert _sums_agree(comp["mu"] * comp["nn"], sums)
    assert n_atoms <= hp.KK <= n_atoms + 10
    # every inferred component is pure in the generating atoms
    homog = sum(np.bincount(zt.reshape(-1)[z == c]).max() for c in range(hp.KK)) / len(z)
    assert homog > 0.99
    rs.close()


def _seating_is_consistent(z, ptr, tji, njt, kjt, n_tbls, KK):
    """The franchise invariants of tests/test_oracle_crf.py::check_crf_invariants, vectorised: every customer sits at
    an existing table of its restaurant and eats that table's dish, njt counts the customers of each table, no
    table is empty.  Returns the number of tables per (restaurant, dish slot) as flat (group, dish) arrays."""
    n_items, lens = len(z), np.diff(ptr)
    group = np.repeat(np.arange(len(lens)), lens)
    assert ((z >= 0) & (z < KK)).all() and (tji >= 0).all() and (tji < n_tbls[group]).all()
    assert ((n_tbls == 0) == (lens == 0)).all()
    slot = ptr[group] + tji                                       # the table's position in njt / kjt
    assert np.array_equal(kjt[slot], z)
    live = (np.arange(n_items) - ptr[group]) < n_tbls[group]      # table slots in use
    seated = np.bincount(slot, minlength=n_items)
    assert np.array_equal(seated[live], njt[live]) and (seated[~live] == 0).all() and (njt[live] > 0).all()
    return group[live], kjt[live]


def test_c4_crf_10m_points_seating_consistent(ctx):
    """Config 4 through the Chinese-restaurant-franchise sampler (CRF_gibbs_sampler!, src/HDP.jl:402-709)."""
    rng = np.random.default_rng(123)
    n_groups, n_per, n_atoms = 100_000, 100, 20
    ks = rng.integers(0, n_atoms, (n_groups, 4))
    zt = np.take_along_axis(ks, rng.integers(0, 4, (n_groups, n_per)), axis=1)
    x = (2.0 * (zt + 1) + np.sqrt(0.001) * rng.standard_normal(zt.shape)).reshape(-1)
    hdp = b.HDP(b.Gaussian1DGaussian1D(float(x.mean()), 10.0, 0.001), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, Kmax=256)
    ptr = np.arange(n_groups + 1, dtype=np.int64) * n_per
    data = b.FlatData.from_csr(b._lib.F64, len(x), group_ptr=ptr, x=x)
    rs = b.ResidentSampler(ctx, hdp, data, np.zeros(len(x), dtype=np.int32), sample_hyperparam=True, n_internals=5, crf=True)
    rs.sweep(6)
    hp, _ = rs.hdp_state()
    z, t, comp = rs.get_zz(), rs.crf_tables(), rs.components()
    _, dish = _seating_is_consistent(z, ptr, t["tji"], t["njt"], t["kjt"], t["n_tbls"], hp.KK)
    assert np.array_equal(t["mm"], np.bincount(dish, minlength=hp.KK)) and (t["mm"] > 0).all()
    assert np.array_equal(comp["nn"], np.bincount(z, minlength=hp.KK))
    assert _sums_agree(comp["mu"] * comp["nn"], np.bincount(z, weights=x, minlength=hp.KK))
    assert n_atoms <= hp.KK <= n_atoms + 20          # a few small extra dishes survive the first iterations
    homog = sum(np.bincount(zt.reshape(-1)[z == c]).max() for c in range(hp.KK)) / len(z)
    assert homog > 0.99
    rs.close()


def test_c5_rcrp_100_epochs_counts_conserved(ctx):
    """Config 5, gen_RCRP_data(fill(500, 100), 0.5) (src/generate_data.jl:57): per-epoch cluster sizes, component
    statistics and the chain rule (a cluster that skipped an epoch does not come back, src/RCRP.jl:309-331)."""
    n_epochs, n_per = 100, 500
    _, zz_true, _ = b.data.gen_rcrp_data([n_per] * n_epochs, 0.5, seed=123)
    rng = np.random.default_rng(5)
    xx = [10.0 * (z + 1) + 0.1 * rng.standard_normal(len(z)) for z in zz_true]
    x = np.concatenate(xx)
    rcrp = b.RCRP(b.Gaussian1DGaussian1D(float(xx[0].mean()), 2.0, 0.01), 1, 1.0, n_epochs, 1.0, 1.0, Kmax=256)
    rs = b.ResidentSampler(ctx, rcrp, b.FlatData(rcrp.component, xx, True), np.zeros(len(x), dtype=np.int32),
                           sample_hyperparam=True, n_internals=10)
    rs.sweep(12)
    KK = rs.current_KK()
    z, comp = rs.get_zz(), rs.components()
    _, aa = rs.rcrp_state()
    assert set(np.unique(z)) == set(range(KK)) and (aa > 0).all() and len(aa) == n_epochs
    assert np.array_equal(comp["nn"], np.bincount(z, minlength=KK))
    assert _sums_agree(comp["mu"] * comp["nn"], np.bincount(z, weights=x, minlength=KK))
    per_epoch = np.stack([np.bincount(z[t * n_per:(t + 1) * n_per], minlength=KK) for t in range(n_epochs)])
    for k in range(KK):                     # the epochs a cluster lives in are contiguous
        on = np.flatnonzero(per_epoch[:, k] > 0)
        assert len(on) == on[-1] - on[0] + 1, k
    zt = np.concatenate(zz_true)
    homog = sum(np.bincount(zt[z == c]).max() for c in range(KK)) / len(z)
    assert homog > 0.97
    rs.close()


def test_c5_rcrf_100_epochs_seating_consistent(ctx):
    """Config 5, the franchise over epochs (RCRF_gibbs_sampler!, src/RCRF.jl:122-938): 100 epochs x 200 restaurants x
    50 points; the seating invariants per restaurant, the tables per (epoch, dish), the component statistics."""
    n_epochs, n_rest, n_per = 100, 200, 50
    rng = np.random.default_rng(321)
    G = n_epochs * n_rest
    eptr = np.arange(n_epochs + 1, dtype=np.int64) * n_rest
    ptr = np.arange(G + 1, dtype=np.int64) * n_per
    epoch = np.repeat(np.arange(G) // n_rest, n_per)
    atoms = (rng.integers(0, 4, G * n_per) + epoch // 5) % 24
    x = 3.0 * (atoms + 1) + 0.05 * rng.standard_normal(len(atoms))
    rcrf = b.RCRF(b.Gaussian1DGaussian1D(float(x.mean()), 100.0, 0.0025), 1, 1.0, 0.1, 0.1, 1.0, 0.1, 0.1, n_epochs, Kmax=256)
    data = b.FlatData.from_csr(b._lib.F64, len(x), group_ptr=ptr, x=x)
    rs = b.ResidentSampler(ctx, rcrf, data, np.zeros(len(x), dtype=np.int32), sample_hyperparam=True, n_internals=5, epoch_ptr=eptr)
    rs.sweep(8)
    KK, gg, aa = rs.rcrf_state()
    z, t, comp = rs.get_zz(), rs.crf_tables(), rs.components()
    grp, dish = _seating_is_consistent(z, ptr, t["tji"], t["njt"], t["kjt"], t["n_tbls"], KK)
    mm = np.zeros((n_epochs, KK), dtype=np.int64)
    np.add.at(mm, (grp // n_rest, dish), 1)
    assert np.array_equal(np.asarray(t["mm"]).reshape(n_epochs, -1)[:, :KK], mm) and (mm.sum(0) > 0).all()
    assert (gg > 0).all() and (aa > 0).all()
    assert np.array_equal(comp["nn"], np.bincount(z, minlength=KK))
    assert _sums_agree(comp["mu"] * comp["nn"], np.bincount(z, weights=x, minlength=KK))
    homog = sum(np.bincount(atoms[z == c]).max() for c in range(KK)) / len(z)
    assert homog > 0.99
    rs.close()
