"""bias.jl_b200 — B200-native collapsed-Gibbs engine behind BIAS.jl's sampler API.

The directory name is not a valid Python identifier; import it as `bias_jl_b200`
(the loader module at the repository root).  The compute path is libbias_cuda
(csrc/, built for sm_100a); there is no CPU fallback.
"""
from ._lib import (BiasError, Context, LIB_PATH, SYMBOLS, device_count, lib)                 # noqa: F401
from .conjugates import Gaussian1DGaussian1D, MultinomialDirichlet, Sent, sparsify_sentence  # noqa: F401
from .distributions import Dirichlet, Gaussian1D                                             # noqa: F401
from .models import (BMM, HDP, LDA, RCRP, RCRP_gibbs_sampler, CRF_gibbs_sampler, RCRF, RCRF_gibbs_sampler, FlatData, ResidentSampler, Group, lda_run_multi, collapsed_gibbs_sampler,      # noqa: F401
                     default_context, init_zz, posterior)
from .data import gen_bars                                                                   # noqa: F401
from . import data, sharding                                                                 # noqa: F401

__all__ = ["LDA", "BMM", "HDP", "RCRP", "RCRP_gibbs_sampler", "CRF_gibbs_sampler", "RCRF", "RCRF_gibbs_sampler", "MultinomialDirichlet", "Gaussian1DGaussian1D", "Gaussian1D", "Dirichlet", "Sent",
           "sparsify_sentence", "init_zz", "collapsed_gibbs_sampler", "posterior", "gen_bars", "Context",
           "ResidentSampler", "Group", "lda_run_multi", "FlatData", "BiasError", "device_count"]
