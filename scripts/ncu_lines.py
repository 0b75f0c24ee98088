"""Dev tool: attribute the SASS-level samples of an ncu report to source lines.

    ncu -i rep.ncu-rep --page source --csv -k regex:k_solve > solve_sass.csv
    python scripts/ncu_lines.py solve_sass.csv _Z7k_solve [min_pct]

The ncu CSV lists SASS instructions in address order; `nvdisasm -g` of the cubin inside the in-tree libplen_b200.so
(same sources, same flags -> same code) gives the source line of each instruction in the same order.  Output: per
source line, share of stall samples / executed warp instructions and the dominant stall reasons.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "plen_ml_walk_b200", "libplen_b200.so")
FRAME = os.environ.get("FRAME", "")      # e.g. FRAME=plen_solve.cuh: fold inlined helpers into their call site in that file


def sass_lines(func_prefix):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        on, cur, chain = False, ("?", 0), []
        for ln in txt.splitlines():
            if ln.startswith("\t.section"):
                on = (".text." + func_prefix) in ln
                continue
            if not on:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                chain.append((os.path.basename(m.group(1)), int(m.group(2))))      # inner frame first, outer frames follow
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                if chain:
                    # FRAME set: attribute to the outermost frame inside that file (the call site in the kernel body)
                    pick = [c for c in chain if c[0] == FRAME]
                    cur = pick[-1] if (FRAME and pick) else chain[0]
                    chain = []
                out.append((cur, m.group(2).strip()))
        if out:
            break
    return out


def main():
    path, func = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # n-th launch in the file
    rows = rows[starts[which]:starts[which + 1]]
    hdr = rows[1]
    i_src, i_smp, i_ins = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[2:] if len(r) > i_ins]
    sl = sass_lines(func)
    if len(sl) != len(body):
        print("warning: %d SASS instructions in the report, %d in the in-tree library" % (len(body), len(sl)))
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    src_text = {}
    tot_s = tot_i = 0
    for k, r in enumerate(body):
        key = sl[k][0] if k < len(sl) else ("?", 0)
        s, n = int(r[i_smp] or 0), int(r[i_ins] or 0)
        a = agg[key]
        a[0] += s; a[1] += n
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a[2][h[6:]] += v
        tot_s += s; tot_i += n
    files = {}
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    tot_st = collections.Counter()
    for a in agg.values():
        tot_st.update(a[2])
    print("stall mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(tot_st.values()))) for k, v in tot_st.most_common(8)))
    for key in sorted(agg, key=lambda k: (k[0], k[1])):
        s, n, st = agg[key]
        if 100.0 * s / max(1, tot_s) < min_pct and 100.0 * n / max(1, tot_i) < min_pct:
            continue
        f = key[0]
        if f not in files:
            cand = [os.path.join(ROOT, "plen_ml_walk_b200", "csrc", f), os.path.join(ROOT, "include", f)]
            files[f] = next((open(c).read().splitlines() for c in cand if os.path.exists(c)), [])
        text = files[f][key[1] - 1].strip()[:90] if 0 < key[1] <= len(files[f]) else ""
        top = ", ".join("%s %d" % kv for kv in st.most_common(3))
        print("%-16s %4d  smp %5.1f%%  inst %5.1f%%  [%s]  %s" % (f, key[1], 100.0 * s / max(1, tot_s), 100.0 * n / max(1, tot_i), top, text))


if __name__ == "__main__":
    main()
