import sys, numpy as np, itertools, collections
sys.path.insert(0, '/root/repo')
import asgfem_b200 as A
M, N = int(sys.argv[1]), int(sys.argv[2])
modes = np.array(A.graded_lex_multiindices(M, N), dtype=np.int64)
idx = {tuple(r): i for i, r in enumerate(modes)}
need = [set([0]) for _ in range(N)]   # directions (0=mean, 1..M)
for i, r in enumerate(modes):
    for m in range(M):
        if r[m] >= 1:
            need[i].add(m + 1)
        up = list(r); up[m] += 1
        if tuple(up) in idx:
            need[i].add(m + 1)
tot = sum(len(s) for s in need)
print("N", N, "pairs", tot, "deg hist", collections.Counter(int(r.sum()) for r in modes))
print("need-size hist", sorted(collections.Counter(len(s) for s in need).items()))

# ---- clustering prototype: units of U modes, D-sets of 8 keys
import time
def cluster(U):
    masks = np.array([sum(1 << d for d in s) for s in need], dtype=np.int64)
    un = set(range(N))
    units = []
    pc = lambda x: bin(x).count("1")
    order = sorted(range(N), key=lambda i: (-pc(int(masks[i])), i))
    t0 = time.time()
    for seed in order:
        if seed not in un: continue
        un.discard(seed)
        cur = [seed]; u = int(masks[seed])
        while len(cur) < U and un:
            best = None; bk = None
            for j in un:
                nu_ = u | int(masks[j])
                c = ((pc(nu_) + 7) // 8, pc(nu_), -pc(int(masks[j]) & u), abs(j - seed))
                if bk is None or c < bk: bk = c; best = j
            # do not grow the number of D-sets for sparse units
            if (bk[0] > (pc(u) + 7) // 8) and len(cur) >= U // 2 and False: break
            un.discard(best); cur.append(best); u |= int(masks[best])
        units.append((cur, u))
    steps = sum((pc(u) + 7) // 8 for _, u in units)
    print("U", U, "units", len(units), "steps", steps, "DMMA per row", steps * 2 * (U // 8), "time", time.time() - t0)
    return units
for U in (8, 16):
    cluster(U)
