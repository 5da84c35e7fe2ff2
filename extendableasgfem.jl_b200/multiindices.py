"""Host mirror of src/mopcontrol.jl.  generate/prepare are trivial list manipulations; the two routines with
O(N^2) searches in the reference (add_boundary_modes :60-132, classify_modes :168-241) run in the native
library (csrc/index.cpp) through the C ABI."""
from __future__ import annotations

import numpy as np

from . import context as _ctx


def generate_multiindices(M, deg):  # mopcontrol.jl:6-22
    mi = [[j] for j in range(deg + 1)]
    L = len(mi)
    for _ in range(M - 1):
        mi = [mi[i] + [j] for i in range(L) for j in range(deg + 1)]
        L *= deg + 1
    return mi


def prepare_multi_indices(multi_indices, minimal_length=0):  # mopcontrol.jl:29-37 (in place)
    L = max(minimal_length, max(len(m) for m in multi_indices))
    for m in multi_indices:
        m.extend([0] * (L - len(m)))


def add_boundary_modes(multi_indices, p_extension=1, tail_extension=(10, 2)):
    """Returns the extended set as a list of lists (active modes first, padded)."""
    mi = [list(m) for m in multi_indices]
    prepare_multi_indices(mi)
    return _ctx.add_boundary_modes(np.array(mi, dtype=np.int64), p_extension, tail_extension).tolist()


def classify_modes(multi_indices_extended, n_active):
    return _ctx.classify_modes(np.array(multi_indices_extended, dtype=np.int64), n_active)


def graded_lex_multiindices(M, N, maxdeg=8):
    """Synthetic benchmark set: total degree ascending, first component descending."""
    def fixed(m, d):
        if m == 1:
            yield [d]
            return
        for first in range(d, -1, -1):
            for rest in fixed(m - 1, d - first):
                yield [first] + rest
    out = []
    for d in range(maxdeg + 1):
        for mi in fixed(M, d):
            out.append(mi)
            if len(out) == N:
                return out
    return out
