#!/usr/bin/env python
"""Executed FP32 work of the physics step from the SASS itself.

ncu's smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on do NOT see the packed instructions of sm_100a (FFMA2 =
fma.rn.f32x2, FADD2, FMUL2): in k_solve FFMA2 is a third of all issued instructions and its op_ffma count equals the scalar
FFMA count alone (profiles/r2_flops_sass.md).  This tool counts from the per-instruction execution counts of the source
page instead:

    ncu --profile-from-start off --section SourceCounters --import-source on -o rep python scripts/profile_steady.py 131072 30 1
    ncu -i rep.ncu-rep --page source --csv > step_sass.csv
    python scripts/flops_sass.py step_sass.csv 131072

flop per predicated-on thread instruction: FFMA2 4, FFMA 2, FADD2 / FMUL2 2, FADD / FMUL 1 (MUFU, FMNMX, conversions: 0).
FMA-pipe cycles per warp instruction: FFMA2 / FADD2 / FMUL2 2, FFMA / FMUL / FADD 1 (checked against
sm__pipe_fma_cycles_active of the same launches).
"""
import collections
import csv
import json
import re
import sys

FLOP = {"FFMA2": 4, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FADD": 1, "FMUL": 1}
PIPE = {"FFMA2": 2, "FADD2": 2, "FMUL2": 2, "FFMA": 1, "FMUL": 1, "FADD": 1}


def main():
    path, robots = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    per = collections.defaultdict(lambda: collections.Counter())
    last = None
    for a, b in zip(starts[:-1], starts[1:]):
        block = tuple(tuple(r) for r in rows[a:b] if r)
        if block == last:       # ncu prints every launch twice (two views of the same SASS): count it once
            continue
        last = block
        name = re.sub(r"^void ", "", rows[a][1])
        name = re.match(r"[A-Za-z0-9_]+(<[^>]*>)?", name).group(0)
        hdr = next(j for j in range(a, b) if rows[j] and rows[j][0] == "Address")
        H = rows[hdr]
        si, wi, ti = H.index("Source"), H.index("Instructions Executed"), H.index("Predicated-On Thread Instructions Executed")
        k = per[name]
        k["launches"] += 1
        for r in rows[hdr + 1:b]:
            if len(r) <= ti:
                continue
            m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[si].strip())
            if not m:
                continue
            op = m.group(2)
            try:
                w, t = float(r[wi]), float(r[ti])
            except ValueError:
                continue
            k["warp_inst"] += w
            k["flop"] += FLOP.get(op, 0) * t
            k["pipe_cycles"] += PIPE.get(op, 0) * w
            if op in FLOP:
                k["warp_" + op] += w
                k["flop_" + op] += FLOP[op] * t
    out, tot = {}, collections.Counter()
    for name, k in per.items():
        out[name] = {"launches": int(k["launches"]), "warp_inst": k["warp_inst"], "flop": k["flop"],
                     "flop_seen_by_ncu_op_metrics": k["flop_FFMA"] + k["flop_FADD"] + k["flop_FMUL"],
                     "fma_pipe_cycles_per_warp_inst": k["pipe_cycles"] / max(k["warp_inst"], 1),
                     "share_of_warp_inst": {op: k["warp_" + op] / max(k["warp_inst"], 1) for op in FLOP if k["warp_" + op]}}
        tot["flop"] += k["flop"]; tot["seen"] += out[name]["flop_seen_by_ncu_op_metrics"]; tot["warp_inst"] += k["warp_inst"]
    out["per_env_step"] = {"robots": robots, "flop": tot["flop"] / robots, "flop_seen_by_ncu_op_metrics": tot["seen"] / robots,
                           "warp_inst": tot["warp_inst"] / robots,
                           "per_kernel_flop": {n: per[n]["flop"] / robots for n in per}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
