"""The wiring a PyDESeq2 maintainer adds for the two steps that are NOT behind the ``Inference`` plugin API but have a device
implementation here: Cook's distances and median-of-ratios size factors.

``pydeseq2`` is imported lazily -- this module is only usable where the reference package is installed::

    from pydeseq2_b200 import B200Inference
    from pydeseq2_b200.integration import b200_dataset_class

    DeseqDataSet = b200_dataset_class()                      # subclass of pydeseq2.dds.DeseqDataSet
    dds = DeseqDataSet(counts=..., metadata=..., design=..., inference=B200Inference(device=0))
    dds.deseq2()                                             # size factors and Cook's distances now come from the GPU too

Everything else -- state handling, the refit, the statistics -- is the reference's own code, untouched.
"""
from __future__ import annotations

import numpy as np


def b200_dataset_class():
    """Returns a ``DeseqDataSet`` subclass whose ``fit_size_factors`` (plain median-of-ratios case) and ``calculate_cooks`` call the
    backend when it offers ``size_factors`` / ``calculate_cooks`` (``B200Inference`` does), and fall back to the reference's own
    implementation otherwise (other fit types, control genes, a different backend)."""
    from pydeseq2.dds import DeseqDataSet

    class B200DeseqDataSet(DeseqDataSet):
        def fit_size_factors(self, fit_type=None, control_genes=None):  # dds.py:584-711
            kind = self.size_factors_fit_type if fit_type is None else fit_type
            plain = (kind == "ratio" and control_genes is None and not hasattr(self, "control_genes")
                     and hasattr(self.inference, "size_factors") and isinstance(self.X, np.ndarray))
            if plain:
                try:
                    normed, sf, logmeans = self.inference.size_factors(self.X, return_logmeans=True)
                except ValueError:  # every gene holds a zero: the reference switches to its iterative estimator (dds.py:682-690)
                    return super().fit_size_factors(fit_type=fit_type, control_genes=control_genes)
                # what the reference keeps from deseq2_norm_fit (dds.py:693): vst_fit / vst_transform normalise NEW counts with them
                self.logmeans, self.filtered_genes = logmeans, ~np.isinf(logmeans)
                self.layers["normed_counts"] = normed
                self.obs["size_factors"] = sf
                self.var["_normed_means"] = normed.mean(0)
                return None
            return super().fit_size_factors(fit_type=fit_type, control_genes=control_genes)

        def calculate_cooks(self):  # dds.py:986-1040
            if not hasattr(self.inference, "calculate_cooks"):
                return super().calculate_cooks()
            if "dispersions" not in self.var:
                self.fit_MAP_dispersions()
            nz = self.var["non_zero"].values
            cooks = self.inference.calculate_cooks(self.X[:, nz], self.obs["size_factors"].values, self.obsm["design_matrix"].values,
                                                   self.obsm["_mu_LFC"], self.obsm["_hat_diagonals"])[0]
            if self.low_memory:
                del self.obsm["_mu_LFC"]
                del self.obsm["_hat_diagonals"]
            self.layers["cooks"] = np.full((self.n_obs, self.n_vars), np.nan)
            self.layers["cooks"][:, nz] = cooks
            return None

    return B200DeseqDataSet
