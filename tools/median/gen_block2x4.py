"""Generator of the 2x4-block 5x5 median: eight neighbouring outputs (y..y+1, x..x+3) from their shared 6x8 window.

Same building blocks as gen_block2x2.py (see there), with more sharing.  Rows 1..4 of every window column are sorted once
(c0..c7).  The 5x5 window of output column ox holds rows 1..4 of columns ox..ox+4 -- a sliding union of five sorted columns --
plus five cells of row 0 (upper output) or row 5 (lower output).  The unions are built from shared merges:
    M12 = c1 u c2, M34 = c3 u c4, M56 = c5 u c6;   core A = M12 u M34 (outputs 0, 1), core B = M34 u M56 (outputs 2, 3),
    of which only ranks 4..13 of 16 can matter;   T0 = c0 u A, T1 = A u c5, T2 = c2 u B, T3 = B u c7 (six middle ranks each).
Per output: the two middle ranks of T u (middle 4 of its row), then the corner clamped between them.
Emits median25_block2x4 into pysp_b200/csrc/median_block2x4.cuh; checked exhaustively (0-1 principle + monotonicity).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_block2x2 import SORT10, Prog, batcher, minimise, search_two_runs  # noqa: E402


def srt(lo, hi):
    return lambda p: all(p[i] <= p[i + 1] for i in range(lo, hi - 1))


def build():
    s4_ops, s4 = minimise(4, batcher(4), lambda p: True, [0, 1, 2, 3], [0, 1, 2, 3], tries=2)
    m44_ops, m44 = minimise(8, batcher(8), lambda p: srt(0, 4)(p) and srt(4, 8)(p), list(range(8)), list(range(8)), tries=20)
    print("merge(4,4):", len(m44), "CEs", m44_ops, "ops")
    best = None
    for t in range(6):
        ops, net = minimise(16, batcher(16), lambda p: srt(0, 8)(p) and srt(8, 16)(p), list(range(3, 13)), list(range(3, 13)),
                            tries=30, seed=t)
        if best is None or ops < best[0]:
            best = (ops, net)
    m88_ops, m88 = best
    print("merge(8,8) -> ranks 3..12:", len(m88), "CEs", m88_ops, "ops")
    mrg_ops, mrg = minimise(14, batcher(14), lambda p: srt(0, 4)(p) and srt(4, 14)(p), list(range(4, 10)), list(range(4, 10)), tries=60)
    print("merge(4,10) -> 6 middle ranks:", len(mrg), "CEs", mrg_ops, "ops")
    two_ops, two, two_groups = search_two_runs(10, [6, 4], [4, 5], [batcher(10), SORT10], trials=150)
    print("middle two of (6,4):", len(two), "CEs", two_ops, "ops")

    P = Prog()
    IN = [["in%d_%d" % (r, c) for c in range(8)] for r in range(6)]
    col = [P.net([IN[r][c] for r in range(1, 5)], s4) for c in range(8)]
    M12, M34, M56 = P.net(col[1] + col[2], m44), P.net(col[3] + col[4], m44), P.net(col[5] + col[6], m44)
    ZA = P.net(M12 + M34, m88)[3:13]
    ZB = P.net(M34 + M56, m88)[3:13]
    T = [P.net(col[0] + ZA, mrg)[4:10], P.net(col[5] + ZA, mrg)[4:10], P.net(col[2] + ZB, mrg)[4:10], P.net(col[7] + ZB, mrg)[4:10]]
    outs = [None] * 8
    for oy in (0, 1):
        row = 0 if oy == 0 else 5
        mids = {0: P.net([IN[row][c] for c in range(1, 5)], s4), 1: P.net([IN[row][c] for c in range(3, 7)], s4)}
        for ox in range(4):
            mid = mids[ox // 2]
            corner = IN[row][(0, 5, 2, 7)[ox]]
            w = [None] * 10
            for k, wire in enumerate(two_groups[0]):
                w[wire] = T[ox][k]
            for k, wire in enumerate(two_groups[1]):
                w[wire] = mid[k]
            w = P.net(w, two)
            m = P.new()
            P.ops.append((m, "min", corner, w[5]))
            o = P.new()
            P.ops.append((o, "max", w[4], m))
            outs[oy * 4 + ox] = o
    P.prune(outs)
    print("program: %d min/max ops for 8 outputs (%.2f per median)" % (len(P.ops), len(P.ops) / 8.0))
    return P, IN, outs


def check_exhaustive(P, IN, outs):
    nb = 25
    words = (1 << nb) // 64
    base = [0xAAAAAAAAAAAAAAAA, 0xCCCCCCCCCCCCCCCC, 0xF0F0F0F0F0F0F0F0, 0xFF00FF00FF00FF00, 0xFFFF0000FFFF0000, 0xFFFFFFFF00000000]
    idx = np.arange(words, dtype=np.uint64)
    wires = []
    for i in range(nb):
        if i < 6:
            wires.append(np.full(words, base[i], dtype=np.uint64))
        else:
            bit = (idx >> np.uint64(i - 6)) & np.uint64(1)
            wires.append(np.where(bit == 1, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0)))
    c = [np.zeros(words, dtype=np.uint64) for _ in range(5)]
    for x in wires:
        carry = x
        for k in range(5):
            t = c[k] & carry
            c[k] = c[k] ^ carry
            carry = t
    expect = c[4] | (c[3] & c[2] & (c[1] | c[0]))
    for oy in (0, 1):
        for ox in range(4):
            for fill in (np.uint64(0), np.uint64(0xFFFFFFFFFFFFFFFF)):
                env, n = {}, 0
                for r in range(6):
                    for cc in range(8):
                        if oy <= r < oy + 5 and ox <= cc < ox + 5:
                            env[IN[r][cc]] = wires[n]
                            n += 1
                        else:
                            env[IN[r][cc]] = np.full(words, fill, dtype=np.uint64)
                for (d, op, a, b) in P.ops:
                    env[d] = (env[a] & env[b]) if op == "min" else (env[a] | env[b])
                assert np.array_equal(env[outs[oy * 4 + ox]], expect), ("exhaustive check failed", oy, ox)
    print("exhaustive 0-1 check ok: 8 outputs x 2^25 windows x {outside all 0, outside all 1}")


def check_random(P, IN, outs, trials=20000):
    rng = np.random.default_rng(1)
    for mode in range(3):
        x = rng.standard_normal((trials, 6, 8)).astype(np.float32)
        if mode == 1:
            x = np.round(x * 2) / 2
        if mode == 2:
            x = rng.integers(0, 2, size=(trials, 6, 8)).astype(np.float32)
        env = {IN[r][c]: x[:, r, c] for r in range(6) for c in range(8)}
        for (d, op, a, b) in P.ops:
            env[d] = np.minimum(env[a], env[b]) if op == "min" else np.maximum(env[a], env[b])
        for oy in (0, 1):
            for ox in range(4):
                ref = np.partition(x[:, oy:oy + 5, ox:ox + 5].reshape(trials, 25), 12, axis=1)[:, 12]
                assert np.array_equal(env[outs[oy * 4 + ox]], ref), (mode, oy, ox)
    print("random check ok")


def emit(P, IN, outs, path):
    lines = ["// GENERATED by tools/median/gen_block2x4.py -- do not edit.",
             "// Exact 5x5 medians of a 2x4 block of outputs from the shared 6x8 window (%d min/max ops, %.2f per median)." % (
                 len(P.ops), len(P.ops) / 8.0),
             "// w[r][c]: rows y-2..y+3, cols x-2..x+5.  out[oy * 4 + ox] = median at (y + oy, x + ox).",
             "#pragma once", '#include "pysp_common.cuh"', "namespace pysp {",
             "PYSP_HD void median25_block2x4(const float (&w)[6][8], float (&out)[8]) {"]
    name = {IN[r][c]: "w[%d][%d]" % (r, c) for r in range(6) for c in range(8)}
    for (d, op, a, b) in P.ops:
        lines.append("    const float %s = %s(%s, %s);" % (d, "fminf" if op == "min" else "fmaxf", name.get(a, a), name.get(b, b)))
    for k, o in enumerate(outs):
        lines.append("    out[%d] = %s;" % (k, name.get(o, o)))
    lines += ["}", "}  // namespace pysp"]
    open(path, "w").write("\n".join(lines) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    P, IN, outs = build()
    check_exhaustive(P, IN, outs)
    check_random(P, IN, outs)
    here = os.path.dirname(os.path.abspath(__file__))
    emit(P, IN, outs, os.path.join(here, "..", "..", "pysp_b200", "csrc", "median_block2x4.cuh"))
