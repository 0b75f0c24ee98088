"""Dev tool: turn the files one `scripts/gpu_round.sh <tag>` pass left under gpurun_out/<tag> into the markdown tables of a
profiles/ summary (launch shares, ncu --set full metrics, bench lines).   python scripts/summarize_round.py gpurun_out/<tag>"""
import collections
import csv
import json
import os
import sys

D = sys.argv[1]


def launches():
    rows = [r for r in csv.reader(open(os.path.join(D, "launches.csv"))) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | device time | share | per launch |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.2f ms | %.1f %% | %.1f-%.1f us |" % (k[:48], len(v), sum(v) / 1e6, 100 * sum(v) / tot, min(v) / 1e3, max(v) / 1e3))


WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"), ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
        ("smsp__issue_active.avg.pct", "issue active %"), ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per issued instruction"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) pipe active %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "pipe tmem %"),
        ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "FADD thread ops"), ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "FMUL thread ops"),
        ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "FFMA thread ops"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %")]


def raw(fn):
    p = os.path.join(D, fn)
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        seen.setdefault(r[idx["Kernel Name"]].split("(")[0], r)
    names = list(seen)
    print("\n| metric | " + " | ".join("`%s`" % n for n in names) + " |\n|---|" + "---|" * len(names))
    for key, label in WANT:
        if key in idx:
            print("| %s | " % label + " | ".join("%s %s" % (seen[n][idx[key]], units[idx[key]]) for n in names) + " |")
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    for n in names:
        top = sorted(((float(seen[n][idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:5]
        print("stalls per issued instruction, `%s`: " % n + ", ".join("%s %.2f" % (b, a) for a, b in top))


def jline(fn, keys=None):
    p = os.path.join(D, fn)
    if os.path.exists(p) and os.path.getsize(p):
        try:
            d = json.loads(open(p).read().strip().splitlines()[-1])
            print("%s: %s" % (fn, json.dumps({k: d[k] for k in keys if k in d} if keys else d)))
        except Exception as e:      # noqa: BLE001
            print(fn, "unreadable", e)


launches()
raw("full_raw.csv")
raw("actor_tc_raw.csv")
print()
jline("bench.json", ["value", "ms_per_step", "e2e", "kernel_ms_per_step", "clocks", "gpu_launches", "cpu_baseline"])
jline("bench_ref.json", ["value", "cpu_baseline"])
for f in ("td3_bench.json", "actor_bench.json", "config3.json", "config4_u8.json", "config4_u8_b1024_fp16.json", "config4_nolearner.json",
          "walk_eval_fp32.json", "walk_eval_fp16.json"):
    jline(f)
for f in ("config2_4096.txt", "config5_1m.txt", "ab.txt", "smoke.log"):
    p = os.path.join(D, f)
    if os.path.exists(p):
        print(f + ":", open(p).read().strip().replace("\n", "\n    "))
p = os.path.join(D, "pytest_gpu.log")
if os.path.exists(p):
    print("pytest -m gpu:", [l for l in open(p).read().splitlines() if "passed" in l or "failed" in l][-1:])
