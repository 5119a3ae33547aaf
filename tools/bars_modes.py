"""Spread of the bars-corpus chain over seeds (GPU kernel selected by BIAS_LDA_KERNEL) next to serial oracle chains."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bias_jl_b200 as b
from oracle import oracle as orc
from helpers import lda_token_loglik

K, V, alpha, aa = 10, 25, 0.1, 2.5
xx, _, _ = b.data.bars_lda_corpus(n_groups=200, n_tokens=100, seed=11)
words = np.concatenate(xx)
rng = np.random.default_rng(21)
z0 = rng.integers(0, K, len(words)).astype(np.int32)
lda = b.LDA(b.MultinomialDirichlet(V, aa), K, alpha)
data = b.FlatData(lda.component, xx, True)
out = []
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    ctx = b.Context(0, seed=1000 + seed)
    rs = b.ResidentSampler(ctx, lda, data, z0)
    rs.sweep(60)
    out.append(lda_token_loglik(data.group_ptr, words, rs.get_zz().astype(np.int64), K, V, alpha, aa / V))
    rs.close(); ctx.close()
print(os.environ.get("BIAS_LDA_KERNEL", "default"), "gpu", np.round(sorted(out), 4).tolist(), "mean", np.mean(out))
if len(sys.argv) > 2:
    ref = []
    for seed in range(int(sys.argv[2])):
        st = orc.LDAState(orc.Comps.md(K, V, aa), orc.Data.words(words), data.group_ptr, z0, alpha)
        r = np.random.default_rng(500 + seed)
        for _ in range(60):
            st.sweep(r.random(data.n_items), diag_mode=0)
        ref.append(lda_token_loglik(data.group_ptr, words, st.z, K, V, alpha, aa / V))
    print("serial", np.round(sorted(ref), 4).tolist(), "mean", np.mean(ref))
