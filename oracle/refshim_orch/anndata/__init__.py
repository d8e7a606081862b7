"""Minimal AnnData stand-in (TEST INFRASTRUCTURE, container-only).

Just enough surface for the reference orchestrator (dds.py / ds.py) to run unmodified
when ``design`` is a DataFrame and ``contrast`` an ndarray.  Written from the attribute
usage listed in SURVEY.md Appendix A; not a copy of anndata.
"""
import numpy as np
import pandas as pd


class AnnData:
    def __init__(self, X=None, obs=None, var=None, obsm=None, varm=None, uns=None, layers=None):
        if isinstance(X, pd.DataFrame):
            if obs is not None and not X.index.equals(obs.index):
                raise ValueError("Index of obs must match index of X.")
            if var is None:
                var = pd.DataFrame(index=X.columns)
            X = X.values
        self.X = np.asarray(X)
        n_obs, n_vars = self.X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(n_obs).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(n_vars).astype(str))
        self.obsm = dict(obsm) if obsm is not None else {}
        self.varm = dict(varm) if varm is not None else {}
        self.uns = dict(uns) if uns is not None else {}
        self.layers = dict(layers) if layers is not None else {}

    n_obs = property(lambda self: self.X.shape[0])
    n_vars = property(lambda self: self.X.shape[1])
    shape = property(lambda self: self.X.shape)
    obs_names = property(lambda self: self.obs.index)
    var_names = property(lambda self: self.var.index)

    def _cols(self, cols):
        if isinstance(cols, pd.Series):
            cols = cols.values
        cols = np.asarray(cols) if not isinstance(cols, pd.Index) else cols
        if isinstance(cols, pd.Index) or cols.dtype.kind in "OUS":
            return self.var.index.get_indexer(cols)
        if cols.dtype == bool:
            return np.flatnonzero(cols)
        return cols.astype(int)

    def __getitem__(self, key):
        rows, cols = key
        assert rows == slice(None)
        idx = self._cols(cols)
        new = AnnData.__new__(AnnData)
        new.X = self.X[:, idx]
        new.obs = self.obs.copy()
        new.var = self.var.iloc[idx].copy()
        new.obsm = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.obsm.items()}
        new.varm = {k: (v.iloc[idx].copy() if hasattr(v, "iloc") else np.asarray(v)[idx]) for k, v in self.varm.items()}
        new.uns = dict(self.uns)
        new.layers = {k: v[:, idx] for k, v in self.layers.items()}
        return new

    def copy(self):
        return self[:, np.arange(self.n_vars)]
