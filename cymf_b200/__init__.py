"""cymf_b200 -- B200-native (sm_100a CUDA) implementation of CyMF's factor-update hot path behind CyMF's
own Python API (reference: minatosato/cymf, cymf/__init__.py:1-7).

    import cymf_b200 as cymf
    model = cymf.BPR(num_components=20, learning_rate=0.01, weight_decay=0.01)
    model.fit(train_csr, num_epochs=30, num_threads=8)
    cymf.evaluator.AverageOverAllEvaluator(test, train, k=5).evaluate(model.W, model.H)
    cymf.WMF(64).fit(train_csr, 5, 8); cymf.RelMF(20).fit(train_csr, 10, 8)
    X, i2w = cymf.glove.read_text("corpus.txt", min_count=5, window_size=10); cymf.GloVe(50).fit(X, 10, 8)

The CUDA library (cymf_b200/libcymf_b200.so, C ABI in include/cymf_b200.h) is loaded on first use; there is
no CPU fallback -- calls raise when it has not been built or no CUDA device is visible.
"""
from .bpr import BPR
from .wmf import WMF
from .relmf import RelMF
from .glove import GloVe
from . import evaluator
from .evaluator import Evaluator, AverageOverAllEvaluator, AoaEvaluator, UnbiasedEvaluator
from . import synth
from . import dataset

__version__ = "0.1.0"


def install_as_cymf():
    """Make `import cymf` resolve to this package (drop-in for scripts written against the reference)."""
    import sys
    sys.modules["cymf"] = sys.modules[__name__]
    return sys.modules[__name__]
