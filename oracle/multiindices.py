"""Oracle: multi-index set management (integer only, restated literally incl. quirks).

Test infrastructure only (see oracle/__init__.py).  Restates

* src/mopcontrol.jl:6-22   generate_multiindices
* src/mopcontrol.jl:29-37  prepare_multi_indices!
* src/mopcontrol.jl:60-132 add_boundary_modes   (the sleep(1) at :74 is dropped)
* src/mopcontrol.jl:168-241 classify_modes
* src/estimate.jl:1-22     get_neighbours (PLUS/MINUS tables, 1-based, 0 = absent)

generate_multiindices is PINNED through runtests.jl:104-127 (ordering of the full tensor set);
the rest is parity-unpinned by the reference's tests.  Multi-indices are Python lists of ints;
all returned index lists are 1-based like the Julia originals.
"""
from __future__ import annotations

import copy
import itertools

import numpy as np


def generate_multiindices(M: int, deg: int):
    mi = [[j] for j in range(deg + 1)]
    L = len(mi)
    for _ in range(M - 1):
        for i in range(L):
            for j in range(deg + 1):
                mi.append(mi[i] + [j])
        mi = mi[L:]
        L *= deg + 1
    return mi


def prepare_multi_indices(multi_indices, minimal_length: int = 0):
    """In place, like the `!` original."""
    new_length = max(minimal_length, max(len(m) for m in multi_indices))
    for m in multi_indices:
        while len(m) < new_length:
            m.append(0)


def add_boundary_modes(multi_indices, p_extension: int = 1, tail_extension=(10, 2)):
    """Mutates `multi_indices` (padding) exactly like the reference and returns the extended list."""
    last_nonzero = 0
    maxdegree1 = 0
    nmodes = len(multi_indices)
    for j in range(1, nmodes):  # j in 2:length
        mj = multi_indices[j]
        maxdegree1 = max(maxdegree1, mj[0])
        # `for k in length:-1:(last_nonzero+1)`: the range is frozen at loop entry, and the
        # lowest qualifying nonzero position wins (quirk)
        for k in range(len(mj), last_nonzero, -1):
            if mj[k - 1] != 0:
                last_nonzero = k
    prepare_multi_indices(multi_indices, minimal_length=last_nonzero + tail_extension[0])
    ext = copy.deepcopy(multi_indices)
    have = {tuple(m) for m in ext}
    maxlength2 = max(len(m) for m in ext)

    def push(mi):
        t = tuple(mi)
        if t not in have:
            have.add(t)
            ext.append(list(mi))

    for k in range(1, maxlength2 + 1):
        new = list(multi_indices[0])
        new[k - 1] = 1
        push(new)
    for k in range(maxdegree1 + 1, maxdegree1 + p_extension + 1):
        new = list(multi_indices[0])
        new[0] = k
        push(new)
    for j in range(nmodes):
        mj = multi_indices[j]
        last_nonzero_pos = 1
        for k in range(len(mj), 0, -1):
            if mj[k - 1] != 0:
                last_nonzero_pos = k
                break
        for k in range(1, last_nonzero_pos + tail_extension[1] + 1):
            if k > last_nonzero + tail_extension[1]:
                break
            new = list(mj)
            new[k - 1] += 1
            push(new)
    return ext


def classify_modes(multi_indices, active_modes=None):
    """Returns (inactive_else, inactive_bnd, inactive_bnd2, active_bnd, active_int), 1-based."""
    if active_modes is None:
        active_modes = multi_indices
    act = {tuple(m) for m in active_modes}
    active_bnd, active_int, inactive_bnd, inactive_bnd2, inactive_else = [], [], [], [], []
    L0 = len(multi_indices[0])
    for j, mj in enumerate(multi_indices, start=1):
        last_nonzero_pos = 1
        for k in range(len(mj), 0, -1):
            if mj[k - 1] != 0:
                last_nonzero_pos = k
                break
        if tuple(mj) in act:
            if last_nonzero_pos == len(mj):
                active_bnd.append(j)
            else:
                active = True
                for k in range(1, last_nonzero_pos + 2):
                    new = list(mj)
                    new[k - 1] += 1
                    if tuple(new) not in act:
                        active = False
                        break
                (active_int if active else active_bnd).append(j)
        else:
            bnd_level = 0
            kmax = min(last_nonzero_pos + 1, L0)
            for k in range(1, kmax + 1):
                new = list(mj)
                if new[k - 1] > 0:
                    new[k - 1] -= 1
                    if tuple(new) in act:
                        bnd_level = 1
                        inactive_bnd.append(j)
                        break
            if bnd_level == 0:
                # a multi-range `for k in .., k2 in ..` is ONE loop nest in Julia: `break` leaves both
                for k, k2 in itertools.product(range(1, kmax + 1), repeat=2):
                    new = list(mj)
                    if new[k - 1] > 0 and new[k2 - 1] > 0:
                        new[k - 1] -= 1
                        new[k2 - 1] -= 1
                        if tuple(new) in act:
                            bnd_level = 2
                            inactive_bnd2.append(j)
                            break
                if bnd_level == 0:
                    inactive_else.append(j)
    return inactive_else, inactive_bnd, inactive_bnd2, active_bnd, active_int


def get_neighbours(multi_indices):
    """PLUS[m, j], MINUS[m, j] (M x N int64, 1-based mode ids, 0 = not in the set)."""
    M = len(multi_indices[0])
    N = len(multi_indices)
    pos = {}
    for k, m in enumerate(multi_indices, start=1):
        pos[tuple(m)] = k  # later duplicates overwrite, like the reference's full scan
    PLUS = np.zeros((M, N), dtype=np.int64)
    MINUS = np.zeros((M, N), dtype=np.int64)
    for j, mj in enumerate(multi_indices):
        for m in range(M):
            mu1 = list(mj)
            mu2 = list(mj)
            mu1[m] += 1
            mu2[m] -= 1
            PLUS[m, j] = pos.get(tuple(mu1), 0)
            MINUS[m, j] = pos.get(tuple(mu2), 0)
    return PLUS, MINUS


def graded_lex_multiindices(M: int, N: int, maxdeg: int = 8):
    """Synthetic set of SURVEY.md §8(d)/Appendix C: total degree ascending; inside one degree
    the first component descending, recursively; truncated to the first N (downward closed
    for the sizes used).  Not part of the reference - it defines the benchmark inputs."""

    def fixed_degree(m, d):
        if m == 1:
            yield [d]
            return
        for first in range(d, -1, -1):
            for rest in fixed_degree(m - 1, d - first):
                yield [first] + rest

    out = []
    for d in range(maxdeg + 1):
        for mi in fixed_degree(M, d):
            out.append(mi)
            if len(out) == N:
                return out
    return out
