#!/usr/bin/env python
"""Generates tests/golden/lda_k1024_convergence.json: held-out (document-completion) perplexity trajectories of the
SERIAL reference sampler at the headline topic count, K = 1024: oracle/bias_oracle.c orc_lda_sweep_linear — the sweep of
src/LDA.jl:113-133 in Float64 with documents and tokens visited in `randperm` order as src/LDA.jl:113-114 does, and the
conditional of src/LDA.jl:26-42 evaluated as the normalised linear weight it equals (tests/test_oracle_golden.py checks
that this makes the same draws as the log-domain restatement orc_lda_sweep; it is what makes 80 sweeps at K = 1024 affordable).

  corpus   : bias_jl_b200.data.dirichlet_lda_corpus(D=10000, L=240, K_true=256, V=5000, beta_conc=0.01, alpha_conc=0.01,
             seed=2024); the first 200 tokens of every document are trained on (2,000,000 tokens), the last 40 held out
  model    : LDA(MultinomialDirichlet(5000, 0.1 * 5000), 1024, 0.1), zz0 = default_rng(77).integers(0, 1024, N)
  chains   : one per seed, uniforms / permutations from default_rng(1000 + seed)
  recorded : perplexity exp(-sum log p / n) of the held-out tokens after the sweeps in CHECKPOINTS

One sweep of the serial sampler over this corpus takes about 20 s on one core; run the chains in parallel:

  python tools/make_convergence_fixture.py --chain 0 &   (... --chain 3)   then   --merge

tests/test_gpu_convergence.py regenerates the same corpus from the same seeds on the GPU box and compares the GPU
chains (one GPU, and 8 shards with the per-sweep exchange) with these numbers.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SPEC = dict(D=10_000, L_train=200, L_held=40, K=1024, K_true=256, V=5000, beta_conc=0.01, alpha_conc=0.01, corpus_seed=2024,
            alpha=0.1, beta=0.1, z_seed=77)
CHECKPOINTS = [1, 5, 10, 20, 40, 60, 80]
OUT = os.path.join(ROOT, "tests", "golden", "lda_k1024_convergence.json")


def corpus(spec=SPEC):
    """-> (train_ptr, train_words int64, held list of arrays, z0 int64)"""
    from bias_jl_b200.data import dirichlet_lda_corpus
    L = spec["L_train"] + spec["L_held"]
    ptr, words = dirichlet_lda_corpus(spec["D"], L, spec["K_true"], spec["V"], beta_conc=spec["beta_conc"],
                                      alpha_conc=spec["alpha_conc"], seed=spec["corpus_seed"])
    docs = words.reshape(spec["D"], L).astype(np.int64)
    train = np.ascontiguousarray(docs[:, :spec["L_train"]]).reshape(-1)
    held = [docs[j, spec["L_train"]:] for j in range(spec["D"])]
    tptr = np.arange(spec["D"] + 1, dtype=np.int64) * spec["L_train"]
    z0 = np.random.default_rng(spec["z_seed"]).integers(0, spec["K"], train.size).astype(np.int64)
    return tptr, train, held, z0


def run_chain(seed: int):
    from oracle import oracle as orc
    from helpers import lda_heldout_loglik
    s = SPEC
    tptr, train, held, z0 = corpus()
    aa = s["beta"] * s["V"]
    st = orc.LDAState(orc.Comps.md(s["K"], s["V"], aa), orc.Data.words(train), tptr, z0, s["alpha"], with_diag=False)
    rng = np.random.default_rng(1000 + seed)
    out = {"seed": seed, "sweeps": [], "perplexity": [], "seconds": []}
    s0, n0 = lda_heldout_loglik(tptr, train, z0, held, s["K"], s["V"], s["alpha"], s["beta"])
    out["initial_perplexity"] = float(np.exp(-s0 / n0))
    t0 = time.time()
    for sweep in range(1, max(CHECKPOINTS) + 1):
        doc_order = rng.permutation(s["D"])
        tok_order = np.concatenate([tptr[j] + rng.permutation(s["L_train"]) for j in doc_order])
        st.sweep_linear(rng.random(train.size), doc_order=doc_order, tok_order=tok_order)
        if sweep in CHECKPOINTS:
            sl, n = lda_heldout_loglik(tptr, train, st.z, held, s["K"], s["V"], s["alpha"], s["beta"])
            out["sweeps"].append(sweep)
            out["perplexity"].append(float(np.exp(-sl / n)))
            out["seconds"].append(round(time.time() - t0, 1))
            with open(OUT.replace(".json", f".chain{seed}.part"), "w") as f:
                json.dump(out, f)
            print(seed, sweep, out["perplexity"][-1], out["seconds"][-1], flush=True)
    return out


def merge():
    chains = []
    for seed in range(16):
        p = OUT.replace(".json", f".chain{seed}.part")
        if os.path.exists(p):
            chains.append(json.load(open(p)))
    n = min(len(c["sweeps"]) for c in chains)
    doc = {"generator": "tools/make_convergence_fixture.py", "spec": SPEC,
           "sampler": "oracle/bias_oracle.c orc_lda_sweep_linear (serial, Float64, randperm document and token order)",
           "sweeps": chains[0]["sweeps"][:n], "initial_perplexity": chains[0]["initial_perplexity"],
           "chains": [{"seed": c["seed"], "perplexity": c["perplexity"][:n]} for c in chains]}
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", OUT, "with", len(chains), "chains,", n, "checkpoints")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain", type=int, default=None)
    ap.add_argument("--merge", action="store_true")
    a = ap.parse_args()
    if a.merge:
        merge()
    else:
        run_chain(a.chain or 0)
