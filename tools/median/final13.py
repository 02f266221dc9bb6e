"""Search a small min/max network for the last stage of the 5x5 median.

Input: the 13 candidate cells of a 5x5 matrix sorted along rows and columns (after column sorts and row sorts the
other 12 cells are certainly below / above the median).  The only 0/1 inputs that can occur are monotone
staircases (252 of them), so candidate networks are checked exhaustively in microseconds.
Objective: number of min/max ops after liveness pruning (a CE of which only one output is used costs one op).
"""
import itertools
import random
import sys

CAND = [(0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 1), (2, 2), (2, 3), (3, 0), (3, 1), (3, 2), (4, 0), (4, 1)]


def staircases():
    pats = []
    # h[c] = number of zeros at the top of column c ... monotone: zeros form a Young diagram in the top-left
    for h in itertools.product(range(6), repeat=5):
        if all(h[i] >= h[i + 1] for i in range(4)):
            m = [[1 if r >= h[c] else 0 for c in range(5)] for r in range(5)]
            pats.append(m)
    return pats


PATS = staircases()
NP = len(PATS)
FULL = (1 << NP) - 1


def wires_init():
    w = []
    for (r, c) in CAND:
        v = 0
        for k, m in enumerate(PATS):
            if m[r][c]:
                v |= 1 << k
        w.append(v)
    exp = 0
    for k, m in enumerate(PATS):
        if sum(sum(row) for row in m) >= 13:
            exp |= 1 << k
    return w, exp


W0, EXPECT = wires_init()


def run(ces):
    """returns (correct_wire_set, ops) ; ops computed for the best output wire"""
    w = list(W0)
    kept = []
    for (i, j) in ces:
        lo, hi = w[i] & w[j], w[i] | w[j]
        if lo == w[i] and hi == w[j]:
            continue
        swapped = (lo == w[j] and hi == w[i])
        w[i], w[j] = lo, hi
        kept.append((i, j, swapped))
    best = None
    for out in range(13):
        if w[out] != EXPECT:
            continue
        live = {out}
        ops = 0
        for (i, j, swapped) in reversed(kept):
            nl, nh = i in live, j in live
            if not (nl or nh):
                continue
            if swapped:
                live.discard(i); live.discard(j)
                if nl: live.add(j)
                if nh: live.add(i)
                continue
            ops += int(nl) + int(nh)
            live.add(i); live.add(j)
        if best is None or ops < best[0]:
            best = (ops, out)
    return best, kept


def minimise(ces, rng):
    """deletion + local mutation hill climbing keeping correctness"""
    cur = list(ces)
    best, _ = run(cur)
    improved = True
    while improved:
        improved = False
        idx = list(range(len(cur)))
        rng.shuffle(idx)
        for k in idx:
            trial = cur[:k] + cur[k + 1:]
            b, _ = run(trial)
            if b is not None and b[0] <= best[0]:
                if b[0] < best[0]:
                    improved = True
                cur, best = trial, b
                break
    return cur, best


def random_net(rng, n):
    ces = []
    for _ in range(n):
        i, j = rng.sample(range(13), 2)
        ces.append((min(i, j), max(i, j)) if rng.random() < 0.5 else (i, j))
    return ces


SORT13 = [(0, 12), (1, 10), (2, 9), (3, 7), (5, 11), (6, 8), (1, 6), (2, 3), (4, 11), (7, 9), (8, 10), (0, 4), (1, 2), (3, 6),
          (7, 8), (9, 10), (11, 12), (4, 6), (5, 9), (8, 11), (10, 12), (0, 5), (3, 8), (4, 7), (6, 11), (9, 10), (0, 1),
          (2, 5), (6, 9), (7, 8), (10, 11), (1, 3), (2, 4), (5, 6), (9, 10), (1, 2), (3, 4), (5, 7), (6, 8), (2, 3), (4, 5),
          (6, 7), (8, 9), (3, 4), (5, 6)]

if __name__ == "__main__":
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    print("patterns", NP)
    overall = None
    for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 300):
        perm = list(range(13))
        rng.shuffle(perm)
        start = [(perm[a], perm[b]) for (a, b) in SORT13]
        b0, _ = run(start)
        if b0 is None:
            continue
        net, b = minimise(start, rng)
        if overall is None or b[0] < overall[0]:
            _, kept = run(net)
            overall = (b[0], b[1], [(i, j) for (i, j, s) in kept if True])
            print("trial", t, "ops", b[0], "out wire", b[1], "CEs", len(kept), kept, flush=True)
