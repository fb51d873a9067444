"""Stand-in for `cymf.dataset` (cymf/dataset/*.py are network downloaders and out of this build's scope): the same
attribute shape as `cymf.dataset.MovieLens` (`.train` / `.valid` / `.test` as scipy lil matrices, `.num_user`,
`.num_item`, `.*_size`; cymf/dataset/movielens.py:62-73) filled from the seeded synthetic generator, so that the
reference's example scripts run offline with `SyntheticMovieLens("ml-100k")` in place of `MovieLens("ml-100k")`."""
import numpy as np
from scipy import sparse

from . import synth


class SyntheticMovieLens(object):
    def __init__(self, dataset_type="ml-100k"):
        if dataset_type not in synth.CONFIGS:
            raise ValueError(f"{dataset_type} is invalid.")                    # movielens.py:24-25
        from sklearn.model_selection import train_test_split
        cfg = synth.CONFIGS[dataset_type]
        X = synth.synth_implicit(cfg["U"], cfg["I"], cfg["nnz"], cfg["seed"]).tocoo()
        self.num_user, self.num_item = X.shape
        pairs = np.stack([X.row, X.col], axis=1)
        train, test = train_test_split(pairs, test_size=0.1, random_state=12345)    # movielens.py:62-63
        train, valid = train_test_split(train, test_size=0.1, random_state=12345)
        self.train, self.valid, self.test = (self._to_matrix(p) for p in (train, valid, test))
        self.train_size, self.valid_size, self.test_size = self.train.nnz, self.valid.nnz, self.test.nnz

    def _to_matrix(self, pairs):
        m = sparse.csr_matrix((np.ones(pairs.shape[0]), (pairs[:, 0], pairs[:, 1])), shape=(self.num_user, self.num_item))
        return m.tolil()


class MovieLens(object):
    def __init__(self, *args, **kwargs):
        raise RuntimeError("cymf_b200 ships no downloader (the build has no network access): use "
                           "cymf_b200.dataset.SyntheticMovieLens(name) or pass your own scipy matrices to fit()")
