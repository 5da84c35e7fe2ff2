"""Oracle: KLE coefficient of cosinus type.

Test infrastructure only (see oracle/__init__.py).  Restates src/coefficients/cosinus.jl:32-55
(constructor incl. the shifted loop and the Hurwitz-zeta amplitude), 58-65 (get_am!), 67-76
(get_gradam!).  Parity unpinned by the reference's tests; SpecialFunctions.zeta(s,z) is the Hurwitz
zeta = scipy.special.zeta(s, q).
"""
from __future__ import annotations

import numpy as np
from scipy.special import zeta


class StochasticCoefficientCosinus:
    def __init__(self, tau=1.0, start=2, decay=2.0, mean=0.0, maxm=100):
        decay_factors = np.zeros(maxm)
        b1 = np.zeros(maxm, dtype=np.int64)
        b2 = np.zeros(maxm, dtype=np.int64)
        k = 0
        j = 0
        for m in range(1, maxm + 2):
            if m > 1:
                decay_factors[m - 2] = float(m - 2 + start) ** (-decay)
                b1[m - 2] = j
                b2[m - 2] = k
            if k > 0:
                j += 1
                k -= 1
            else:
                k = (j + k) + 1
                j = 0
        amp = tau / zeta(decay, start)
        decay_factors *= amp
        self.decay = decay
        self.mean_value = float(mean)
        self.decay_factors = decay_factors
        self.b1 = b1
        self.b2 = b2

    @property
    def maxm(self):
        return len(self.decay_factors)

    def am(self, m, x, y):
        """a_m at points (x, y) (arrays); m = 0 is the mean."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if m == 0:
            return np.full(np.broadcast(x, y).shape, self.mean_value)
        return self.decay_factors[m - 1] * np.cos(np.pi * self.b1[m - 1] * x) * np.cos(np.pi * self.b2[m - 1] * y)

    def gradam(self, m, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if m == 0:
            z = np.zeros(np.broadcast(x, y).shape)
            return z, z.copy()
        b1 = self.b1[m - 1]
        b2 = self.b2[m - 1]
        d = self.decay_factors[m - 1]
        gx = -b1 * np.pi * np.sin(b1 * np.pi * x) * np.cos(b2 * np.pi * y) * d
        gy = -b2 * np.pi * np.cos(b1 * np.pi * x) * np.sin(b2 * np.pi * y) * d
        return gx, gy
