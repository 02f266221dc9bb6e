"""Join an ncu SASS source page (per-instruction samples / executed counts) with nvdisasm line info.

    python tools/ncu_by_line.py <report.ncu-rep> <kernel-regex> [lib.so] [--top N] [--inline]
Prints per source line (file:line): share of stall samples, share of executed warp instructions, and the
opcode mix.  The .so must be the build that was profiled.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel_re):
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(d, cub)], stderr=subprocess.DEVNULL).decode()
    out, cur, line, infn = {}, None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = re.search(kernel_re, m.group(1)) is not None
            cur = m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (line, m.group(2).strip())
    return out


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    so = next((a for a in sys.argv[3:] if a.endswith(".so")), "pysp_b200/libpysp_b200.so")
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                                  stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    sass = sass_lines(so, kre)
    base = None
    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    by_op = collections.Counter()
    tot_s = tot_i = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break               # only the first matching launch
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        samp, inst = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        line, txt = sass.get(off, (None, r[ix["Source"]]))
        op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        e = by_line[line]
        e[0] += samp
        e[1] += inst
        e[2][op] += inst
        by_op[op] += inst
        tot_s += samp
        tot_i += inst
    print("kernel %s: %d warp instructions, %d samples" % (kre, tot_i, tot_s))
    print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in by_op.most_common(18)))
    for line, (s, i, ops) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        name = "%s:%d" % line if line else "?"
        print("%5.1f%% samples %5.1f%% inst  %-24s %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), name,
                                                          " ".join("%s:%d%%" % (k, round(100.0 * v / max(i, 1))) for k, v in ops.most_common(5))))


if __name__ == "__main__":
    main()
