"""Host mirror of src/coefficients/cosinus.jl: the parameter tables of StochasticCoefficientCosinus
(constructor :32-55).  Evaluation of a_m at quadrature points happens on the device
(csrc/assemble.cu, csrc/estimate.cu); `get_am` is kept for host-side users (e.g. plotting, sampling)."""
from __future__ import annotations

import numpy as np
from scipy.special import zeta


class StochasticCoefficientCosinus:
    def __init__(self, tau=1.0, start=2, decay=2.0, mean=0.0, maxm=100):
        self.decay = decay
        self.mean_value = float(mean)
        d = np.zeros(maxm)
        b1 = np.zeros(maxm, dtype=np.int64)
        b2 = np.zeros(maxm, dtype=np.int64)
        j = k = 0
        for m in range(1, maxm + 2):  # shifted loop of cosinus.jl:38-51
            if m > 1:
                d[m - 2] = float(m - 2 + start) ** (-decay)
                b1[m - 2], b2[m - 2] = j, k
            if k > 0:
                j, k = j + 1, k - 1
            else:
                j, k = 0, j + k + 1
        self.decay_factors = d * (tau / zeta(decay, start))  # Hurwitz zeta amplitude (:52-53)
        self.b1, self.b2 = b1, b2

    maxm = property(lambda self: len(self.decay_factors))

    def get_am(self, m, x, y):
        if m == 0:
            return np.full(np.broadcast(x, y).shape, self.mean_value)
        return self.decay_factors[m - 1] * np.cos(np.pi * self.b1[m - 1] * x) * np.cos(np.pi * self.b2[m - 1] * y)
