"""Device-side table post-processing: the notebook's MinMaxScaler step.

Reference (NB = raw line of channel_importance_hand_crafted_features.ipynb, cell 16):
    norm = MinMaxScaler().fit(X_train)        NB:389
    X_train = norm.transform(X_train)         NB:391
    X_test = norm.transform(X_test)           NB:394
``MinMaxScaler`` below has the same ``fit`` / ``transform`` / ``fit_transform`` calls and the
same fitted attributes (``data_min_``, ``data_max_``, ``data_range_``, ``scale_``, ``min_``) with
sklearn's semantics (NaN ignored when fitting and kept when transforming, zero ranges scale by 1,
default feature_range (0, 1)), but works on a feature table that lives on the GPU
(``FeatureExtractor.extract_planar`` output), so a table kept on the device for repeated
ablations is never copied to the host.  numpy input is accepted for convenience (uploaded,
result downloaded).  The arithmetic is in libimfeat.so (csrc/post_kernels.cuh); no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from .extractor import _ptr, _require_cuda, get_extractor


class MinMaxScaler:
    def __init__(self, device=None):
        self._ex = get_extractor(device=device)
        self._stats = None

    # -- helpers -------------------------------------------------------------------------------
    def _to_device(self, X):
        torch = _require_cuda()
        was_numpy = isinstance(X, np.ndarray)
        if was_numpy:
            X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(
                torch.device("cuda", self._ex.device))
        if not (X.is_cuda and X.dtype == torch.float64 and X.dim() == 2):
            raise _lib.ImfeatError("MinMaxScaler expects a 2-D float64 CUDA tensor (or a numpy array)")
        if X.stride(1) != 1:
            X = X.contiguous()
        return X, was_numpy

    # -- sklearn-shaped interface --------------------------------------------------------------
    def fit(self, X):
        torch = _require_cuda()
        X, _ = self._to_device(X)
        n, f = int(X.shape[0]), int(X.shape[1])
        stats = torch.empty((4, f), dtype=torch.float64, device=X.device)
        _lib.check(self._ex.lib.imfeat_minmax_fit_device(
            self._ex._ctx, _ptr(X), n, f, int(X.stride(0)) if n > 1 else f, _ptr(stats),
            self._ex._stream()), self._ex._ctx)
        self._stats = stats
        self.n_features_in_ = f
        self.n_samples_seen_ = n
        return self

    def transform(self, X, out=None):
        torch = _require_cuda()
        if self._stats is None:
            raise _lib.ImfeatError("MinMaxScaler.transform called before fit")
        X, was_numpy = self._to_device(X)
        n, f = int(X.shape[0]), int(X.shape[1])
        if f != self.n_features_in_:
            raise _lib.ImfeatError("X has %d features, the scaler was fitted with %d" % (f, self.n_features_in_))
        if out is None:
            out = torch.empty((n, f), dtype=torch.float64, device=X.device)
        _lib.check(self._ex.lib.imfeat_minmax_transform_device(
            self._ex._ctx, _ptr(X), n, f, int(X.stride(0)) if n > 1 else f, _ptr(self._stats), _ptr(out),
            int(out.stride(0)) if n > 1 else f, self._ex._stream()), self._ex._ctx)
        return out.cpu().numpy() if was_numpy else out

    def fit_transform(self, X):
        return self.fit(X).transform(X)

    # fitted attributes (device tensors; call .cpu().numpy() to compare with sklearn)
    @property
    def data_min_(self):
        return self._stats[0]

    @property
    def data_max_(self):
        return self._stats[1]

    @property
    def data_range_(self):
        return self._stats[1] - self._stats[0]

    @property
    def scale_(self):
        return self._stats[2]

    @property
    def min_(self):
        return self._stats[3]
