"""Drop-in `Model` (reference circuitmap/model.py:15-162) for method='caviar'.

Same constructor, same `fit(obs, stimuli, method='caviar', fit_options=dict())`, same `state` keys
(mu, beta, lam, shape, rate, phi, phi_cov, z, receptive_fields, alpha untouched), `history`, `trial_count`,
`time`.  The arithmetic runs in csrc/caviar.cu behind cm_caviar_fit.
"""
import time
from copy import deepcopy

import numpy as np

from . import optimise


class Model:
    def __init__(self, N, priors=None):
        """Initialise circuitmap model (model.py:16-34)."""
        self.N = N
        self.priors = priors if priors is not None else {}
        self.priors.setdefault("alpha", 1 / 4 * np.ones(N))
        self.priors.setdefault("phi", np.c_[1e-1 * np.ones(N), 5e0 * np.ones(N)])
        self.priors.setdefault("phi_cov", np.array([np.array([[1e-1, 0], [0, 1e0]]) for _ in range(N)]))
        self.priors.setdefault("mu", np.zeros(N))
        self.priors.setdefault("beta", 1e1 * np.ones(N))
        self.priors.setdefault("shape", 1.)
        self.priors.setdefault("rate", 1e-1)
        self.state = deepcopy(self.priors)

    def fit(self, obs, stimuli, method="caviar", fit_options=dict()):
        if method == "caviar":
            self._fit_caviar(obs, stimuli, fit_options)
        elif method in ("mbcs", "cavi_sns"):
            raise NotImplementedError("method=%r is outside the B200 hot path (only 'caviar' is implemented)" % method)
        else:
            raise Exception          # model.py:44

    def _fit_caviar(self, obs, stimuli, fit_options):
        t_start = time.time()
        result = optimise.caviar(
            obs, stimuli, self.state["mu"], self.state["beta"], self.state["shape"],
            self.state["rate"], self.state["phi"], self.state["phi_cov"], **fit_options)
        t_end = time.time()
        mu, beta, lam, shape, rate, phi, phi_cov, z, receptive_fields, mu_hist, beta_hist, lam_hist, shape_hist, \
            rate_hist, phi_hist, phi_cov_hist, z_hist = result
        # optimise.caviar returns fresh NumPy arrays: np.asarray keeps them (the reference's np.array() converts JAX
        # arrays here, model.py:138-146; copying the 8 N K bytes of lam again would only cost time)
        self.state["mu"] = np.asarray(mu)
        self.state["beta"] = np.asarray(beta)
        self.state["shape"] = np.asarray(shape)
        self.state["rate"] = np.asarray(rate)
        self.state["phi"] = np.asarray(phi)
        self.state["phi_cov"] = np.asarray(phi_cov)
        self.state["lam"] = np.asarray(lam)
        self.state["z"] = np.asarray(z)
        self.state["receptive_fields"] = np.array(receptive_fields)
        self.trial_count = lam.shape[1]
        self.time = t_end - t_start
        self.history = {
            "mu": np.asarray(mu_hist), "beta": np.asarray(beta_hist), "lam": np.asarray(lam_hist),
            "shape": np.asarray(shape_hist), "rate": np.asarray(rate_hist), "phi": np.asarray(phi_hist),
            "phi_cov": np.asarray(phi_cov_hist), "z": np.asarray(z_hist),
        }
        return
