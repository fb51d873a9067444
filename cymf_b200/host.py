"""Host-side logic shared by BPR.fit and WMF.fit: the seeded prologue (cymf/bpr.pyx:97-101 == cymf/wmf.pyx:88-92)
and the epoch loop with per-epoch validation / early stopping (cymf/bpr.pyx:159-190 == cymf/wmf.pyx:109-132)."""
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


def run_epochs(model, sess, num_epochs, step, verbose, ncols):
    """The reference's epoch loop.  `sess` keeps the factors on the device; `step()` enqueues one epoch.

    With a `valid_evaluator` the reference evaluates the live factors after every epoch, keeps the best
    DCG@5, counts consecutive non-improving epochs and stops on the 12th (`count > 10`, strict `>` so ties count
    as improvements), then rebinds model.W / model.H to copies of the best epoch (bpr.pyx:173-190).  When the
    evaluator is cymf_b200's own, scoring reads the device-resident factors (dense f64 views) and the best-epoch
    snapshot stays on the device too: no factor leaves HBM until the fit ends.  Any other evaluator object gets
    the NumPy arrays, refreshed in place every epoch, exactly as the reference passes them."""
    from tqdm import tqdm
    from .evaluator import Evaluator
    valid_evaluator = getattr(model, "valid_evaluator", None)
    early_stopping = getattr(model, "early_stopping", False)
    W, H = model.W, model.H
    on_device = isinstance(valid_evaluator, Evaluator)
    best = None
    if valid_evaluator:
        best = sess.snapshot() if on_device else (W.copy(), H.copy())
    count = 0
    with tqdm(total=num_epochs, leave=True, ncols=ncols, disable=not verbose) as progress:
        for epoch in range(num_epochs):
            step()
            if valid_evaluator:
                if on_device:
                    valid_dcg = valid_evaluator.evaluate(*sess.dense_f64())["DCG@5"]
                else:
                    sess.download(W, H)
                    valid_dcg = valid_evaluator.evaluate(W, H)["DCG@5"]
                if early_stopping and model.valid_dcg > valid_dcg and count > 10:
                    break
                elif early_stopping and model.valid_dcg > valid_dcg:
                    count += 1
                else:
                    count = 0
                    model.valid_dcg = valid_dcg
                    best = sess.snapshot() if on_device else (W.copy(), H.copy())
            progress.set_description(
                f"EPOCH={epoch+1:{len(str(num_epochs))}} "
                f"{(', DCG@5=' + str(np.round(valid_dcg, 3))) if valid_evaluator else ''}")
            progress.update(1)
    sess.download(W, H)                                    # model.W / model.H hold the last epoch, in place
    if valid_evaluator and early_stopping:                 # ... and are rebound to copies of the best one
        if on_device:
            sess.restore(best)
            Wb, Hb = np.empty_like(W), np.empty_like(H)
            sess.download(Wb, Hb)
            model.W, model.H = Wb, Hb
        else:
            model.W, model.H = best[0].copy(), best[1].copy()
