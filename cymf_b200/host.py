"""Host-side prologue shared by BPR.fit and WMF.fit (cymf/bpr.pyx:97-101 == cymf/wmf.pyx:88-92)."""
import numpy as np


def init_missing_factors(model, n_rows, n_cols):
    """Seed 4321 is set ONLY when W is absent; a preset W with a missing H draws H from the ambient RNG state."""
    K = model.num_components
    if model.W is None:
        np.random.seed(4321)
        model.W = np.random.uniform(low=-0.1, high=0.1, size=(n_rows, K)) / K
    if model.H is None:
        model.H = np.random.uniform(low=-0.1, high=0.1, size=(n_cols, K)) / K


def init_factors(n_rows, n_cols, K):
    class _M:
        num_components, W, H = K, None, None
    m = _M()
    init_missing_factors(m, n_rows, n_cols)
    return m.W, m.H
