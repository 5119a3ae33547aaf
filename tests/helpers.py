"""Shared helpers for the parity tests."""
import numpy as np


def near_tie(prob_row: np.ndarray, u: float, tol: float) -> bool:
    """True when u sits within tol of a CDF edge (the only place a draw may legitimately differ)."""
    cdf = np.cumsum(prob_row)
    return bool(np.min(np.abs(cdf - u)) < tol)


def assert_draws_match(draw_gpu, draw_ref, probs, uniforms, tol, what=""):
    bad = np.flatnonzero(np.asarray(draw_gpu) != np.asarray(draw_ref))
    for q in bad:
        assert near_tie(probs[q], uniforms[q], tol), (
            f"{what}: draw {draw_gpu[q]} != oracle {draw_ref[q]} at query {q}, u={uniforms[q]!r} "
            f"is not within {tol} of a CDF edge")
    return len(bad)


def norm_logp(raw: np.ndarray) -> np.ndarray:
    m = raw.max(axis=-1, keepdims=True)
    return raw - (m + np.log(np.exp(raw - m).sum(axis=-1, keepdims=True)))


def assert_conditionals_close(raw_gpu, raw_ref, what=""):
    """Parity level (a): normalised log-probabilities within 1e-5 relative (+1e-6 absolute)."""
    a, b = norm_logp(np.asarray(raw_gpu)), norm_logp(np.asarray(raw_ref))
    err = np.abs(a - b)
    tol = 1e-5 * np.abs(b) + 1e-6
    assert (err <= tol).all(), f"{what}: max normalised log-prob error {err.max():.3e}"


def lda_token_loglik(ptr, words, z, K, V, alpha, beta):
    """Per-token predictive log-likelihood of the training tokens under the point estimates
    theta_dk = (n_dk+alpha)/(L_d+K alpha), phi_kw = (n_kw+beta)/(n_k+V beta).  (The reference has no
    perplexity function, src/notes.md:30; this is the evaluator both sides are scored with.)"""
    D = len(ptr) - 1
    n_kw = np.zeros((K, V))
    np.add.at(n_kw, (z, words), 1)
    phi = (n_kw + beta) / (n_kw.sum(1, keepdims=True) + V * beta)
    total = 0.0
    for j in range(D):
        s, e = ptr[j], ptr[j + 1]
        if e == s:
            continue
        n_dk = np.bincount(z[s:e], minlength=K)
        theta = (n_dk + alpha) / (e - s + K * alpha)
        total += np.log(theta @ phi[:, words[s:e]]).sum()
    return total / max(1, ptr[-1])


def lda_heldout_loglik(ptr, words, z, heldout, K, V, alpha, beta):
    """Document completion: sum over held-out tokens of log sum_k theta_dk phi_kw with theta from the training
    tokens of the document and phi from the training counts (same point estimates as lda_token_loglik)."""
    n_kw = np.zeros((K, V))
    np.add.at(n_kw, (z, words), 1)
    phi = (n_kw + beta) / (n_kw.sum(1, keepdims=True) + V * beta)
    total, n = 0.0, 0
    for j, h in enumerate(heldout):
        if len(h) == 0:
            continue
        s, e = ptr[j], ptr[j + 1]
        n_dk = np.bincount(z[s:e], minlength=K)
        theta = (n_dk + alpha) / (e - s + K * alpha)
        total += np.log(theta @ phi[:, np.asarray(h)]).sum()
        n += len(h)
    return total, n
