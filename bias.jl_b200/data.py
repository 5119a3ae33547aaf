"""Synthetic inputs for the hot path (host side; not accelerated): gen_bars (src/common.jl:157-175),
the bars corpora of the demos (doc/examples/demo_LDA_MultinomialDirichlet.jl:14-41,
demo_BMM_MultinomialDirichlet.jl), Gaussian mixtures (demo_BMM_Gaussian1DGaussian1D.jl) and the
CRF seating generator used by the HDP demos (src/generate_data.jl:116-296).
Indices are 0-based."""
from __future__ import annotations

import numpy as np

from .conjugates import Sent, sparsify_sentence


def gen_bars(n_bars: int, n_vocab: int, noise_level: float) -> np.ndarray:
    """Bar topics on a sqrt(V) x sqrt(V) grid, column-major flatten (src/common.jl:157-175)."""
    S = int(round(np.sqrt(n_vocab)))
    half = int(np.floor(n_bars / 2 + 0.5))
    bars = np.zeros((n_bars, n_vocab))
    for kk in range(n_bars):
        b = np.zeros((S, S)) + noise_level
        if kk < half:
            b[kk, :] = 1.0
        else:
            b[:, kk - half] = 1.0
        b /= b.sum()
        bars[kk, :] = b.flatten(order="F")
    return bars


def sample_categorical(p: np.ndarray, n: int, rng) -> np.ndarray:
    """`sample(w, n)` (src/common.jl:80-86) with numpy's generator."""
    cdf = np.cumsum(p)
    return np.minimum(np.searchsorted(cdf, rng.random(n) * cdf[-1], side="left"), len(p) - 1)


def bars_lda_corpus(n_groups=1000, n_tokens=100, true_KK=10, vocab_size=25, max_n_topic_doc=4, seed=123):
    """The LDA bars corpus (demo_LDA_MultinomialDirichlet.jl:14-41) at the size of BASELINE config 1.
    Returns (xx, true_zz, topics): xx[j] int64 word ids."""
    rng = np.random.default_rng(seed)
    topics = gen_bars(true_KK, vocab_size, 0.0)
    xx, true_zz = [], []
    for _ in range(n_groups):
        n_topic_doc = int(rng.integers(1, max_n_topic_doc + 1))
        topic_doc = rng.permutation(true_KK)[:n_topic_doc]
        kk = topic_doc[rng.integers(0, n_topic_doc, n_tokens)]
        words = np.empty(n_tokens, dtype=np.int64)
        for k in np.unique(kk):
            m = kk == k
            words[m] = sample_categorical(topics[k], int(m.sum()), rng)
        xx.append(words)
        true_zz.append(kk.astype(np.int64))
    return xx, true_zz, topics


def bars_bmm_sentences(n_sentences=200, n_tokens=15, true_KK=4, vocab_size=25, seed=123):
    """Sent observations from bar topics (demo_BMM_MultinomialDirichlet.jl). Returns (xx, true_zz, topics)."""
    rng = np.random.default_rng(seed)
    topics = gen_bars(true_KK, vocab_size, 0.0)
    true_zz = rng.integers(0, true_KK, n_sentences)
    xx = [sparsify_sentence(sample_categorical(topics[k], n_tokens, rng)) for k in true_zz]
    return xx, true_zz.astype(np.int64), topics


def bars_bmm_sentences_csr(n_sentences, n_tokens, true_KK, vocab_size, seed=123):
    """Vectorised CSR version of bars_bmm_sentences for large N (BASELINE config 2)."""
    rng = np.random.default_rng(seed)
    topics = gen_bars(true_KK, vocab_size, 0.0)
    cdf = np.cumsum(topics, axis=1)
    true_zz = rng.integers(0, true_KK, n_sentences)
    u = rng.random((n_sentences, n_tokens))
    words = np.empty((n_sentences, n_tokens), dtype=np.int64)
    for k in range(true_KK):
        m = true_zz == k
        words[m] = np.minimum(np.searchsorted(cdf[k], u[m].ravel()), vocab_size - 1).reshape(-1, n_tokens)
    words.sort(axis=1)
    first = np.ones_like(words, dtype=bool)
    first[:, 1:] = words[:, 1:] != words[:, :-1]
    nnz_per = first.sum(1)
    sent_ptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.int64)
    flat_first = first.ravel()
    sent_word = words.ravel()[flat_first].astype(np.int32)
    starts = np.flatnonzero(flat_first)
    ends = np.append(starts[1:], words.size)
    row_end = (np.arange(n_sentences) + 1) * n_tokens
    ends = np.minimum(ends, np.repeat(row_end, nnz_per))
    sent_count = (ends - starts).astype(np.int32)
    return sent_ptr, sent_word, sent_count, true_zz.astype(np.int64), topics


def gaussian_mixture(n=500, true_KK=5, vv=0.01, seed=123):
    """demo_BMM_Gaussian1DGaussian1D.jl: atoms at 1..KK, likelihood variance vv."""
    rng = np.random.default_rng(seed)
    atoms = np.arange(1, true_KK + 1, dtype=np.float64)
    zz = rng.integers(0, true_KK, n)
    return atoms[zz] + np.sqrt(vv) * rng.standard_normal(n), zz.astype(np.int64), atoms


def gaussian_groups(n_groups=50, n_per=100, true_KK=5, vv=0.01, max_per_group=3, seed=123):
    """Grouped Gaussian data (demo_LDA_Gaussian1DGaussian1D.jl): each group mixes a few atoms."""
    rng = np.random.default_rng(seed)
    atoms = np.arange(1, true_KK + 1, dtype=np.float64)
    xx, zz = [], []
    for _ in range(n_groups):
        ks = rng.permutation(true_KK)[:int(rng.integers(1, max_per_group + 1))]
        z = ks[rng.integers(0, len(ks), n_per)]
        xx.append(atoms[z] + np.sqrt(vv) * rng.standard_normal(n_per))
        zz.append(z.astype(np.int64))
    return xx, zz, atoms


def gen_crf_data(n_group_j, gg: float, aa: float, seed=123):
    """Chinese-restaurant-franchise seating (src/generate_data.jl:116-296, join_tables=true view):
    returns per-group dish labels zz[j] (0-based) and the number of dishes."""
    rng = np.random.default_rng(seed)
    m_k: list[int] = []                  # tables per dish
    zz = []
    for nj in n_group_j:
        tables_n, tables_k = [], []
        z = np.empty(nj, dtype=np.int64)
        for i in range(nj):
            w = np.array(tables_n + [aa], dtype=np.float64)
            t = int(np.searchsorted(np.cumsum(w), rng.random() * w.sum()))
            t = min(t, len(w) - 1)
            if t == len(tables_n):       # new table: choose its dish from the franchise
                wk = np.array(m_k + [gg], dtype=np.float64)
                k = int(np.searchsorted(np.cumsum(wk), rng.random() * wk.sum()))
                k = min(k, len(wk) - 1)
                if k == len(m_k):
                    m_k.append(0)
                m_k[k] += 1
                tables_n.append(0)
                tables_k.append(k)
            tables_n[t] += 1
            z[i] = tables_k[t]
        zz.append(z)
    return zz, len(m_k)


def dirichlet_lda_corpus(D, L, K, V, beta_conc=0.01, alpha_conc=0.1, seed=123):
    """Host version of the large synthetic corpus (BASELINE config 3 shape, any size):
    phi_k ~ Dirichlet(beta_conc), theta_d ~ Dirichlet(alpha_conc)
    (demo_LDA_MultinomialDirichlet_sparse.jl:26,33).  Returns doc_ptr, words (int32)."""
    rng = np.random.default_rng(seed)
    phi = rng.gamma(beta_conc, size=(K, V)) + 1e-300
    phi /= phi.sum(1, keepdims=True)
    cdf = np.cumsum(phi, axis=1)
    words = np.empty(D * L, dtype=np.int32)
    chunk = max(1, 2_000_000 // max(K, 1))
    for d0 in range(0, D, chunk):
        d1 = min(D, d0 + chunk)
        theta = rng.gamma(alpha_conc, size=(d1 - d0, K)) + 1e-300
        tc = np.cumsum(theta, axis=1)
        u = rng.random((d1 - d0, L)) * tc[:, -1:]
        z = np.empty((d1 - d0, L), dtype=np.int64)
        for r in range(d1 - d0):
            z[r] = np.minimum(np.searchsorted(tc[r], u[r]), K - 1)
        zf = z.ravel()
        uw = rng.random(zf.size)
        out = np.empty(zf.size, dtype=np.int32)
        order = np.argsort(zf, kind="stable")
        bounds = np.searchsorted(zf[order], np.arange(K + 1))
        for k in range(K):
            sel = order[bounds[k]:bounds[k + 1]]
            if sel.size:
                out[sel] = np.minimum(np.searchsorted(cdf[k], uw[sel] * cdf[k, -1]), V - 1)
        words[d0 * L:d1 * L] = out
    doc_ptr = np.arange(D + 1, dtype=np.int64) * L
    return doc_ptr, words


def gen_rcrp_data(N_t, aa: float, seed=123):
    """Recurrent-CRP seating (src/generate_data.jl:57-113): epoch 1 is a plain CRP; a customer of epoch t
    joins cluster k with weight n[t-1,k] + n[t,k] and opens a new one with weight aa.
    Returns (nn [T x KK], zz list of 0-based label arrays, KK)."""
    rng = np.random.default_rng(seed)
    TT = len(N_t)
    counts: list[np.ndarray] = []        # one length-TT column per cluster
    zz = []
    for tt in range(TT):
        z = np.empty(N_t[tt], dtype=np.int64)
        for ii in range(N_t[tt]):
            if tt == 0 and ii == 0:
                k = 0                                           # :66-70
                counts.append(np.zeros(TT, dtype=np.int64))
            else:
                w = np.array([c[tt] + (c[tt - 1] if tt > 0 else 0) for c in counts] + [aa], dtype=np.float64)
                w /= w.sum()
                k = int(min(np.searchsorted(np.cumsum(w), rng.random()), len(w) - 1))
                if k == len(counts):
                    counts.append(np.zeros(TT, dtype=np.int64))
            counts[k][tt] += 1
            z[ii] = k
        zz.append(z)
    nn = np.stack(counts, axis=1) if counts else np.zeros((TT, 0), dtype=np.int64)
    return nn, zz, len(counts)
