"""Dev tool: static single-warp issue model of a SASS region (no GPU needed).

    python scripts/sass_model.py <kernel-name-substring> [--lib path.so] [--list-branches] [--from 0xADDR --to 0xADDR]
                                 [--take 0xADDR ...] [--lat-shfl N] [--lat-lds N]

Decodes the scheduling control fields ptxas wrote into every 128-bit instruction (stall count, yield, write / read
scoreboard slot, wait mask -- the layout documented in /opt/skills/guides/B300_MICROARCH.md) and walks the region in
program order with the guide's single-warp model:

    T = max(T + stall, max(SB_done[s] for s in wait_mask));  SB_done[wbar] = T + latency(op)

Conditional branches fall through unless their address is listed with --take (backward branches always fall through).
Output: cycles of the region for ONE warp running alone, instructions issued, and which scoreboard waits were exposed
(by opcode of the producer).  It is a model for choosing between code shapes before spending GPU time, not a measurement.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAT = {"LDS": 29, "SHFL": 24, "LDG": 600, "LD": 600, "LDC": 40, "LDCU": 40, "S2R": 25, "S2UR": 25, "MUFU": 18, "LDL": 40, "STL": 10,
       "REDUX": 30, "VOTE": 10, "I2F": 18, "F2I": 18, "POPC": 18, "FLO": 18, "ATOMS": 60, "R2UR": 12, "BAR": 30, "MATCH": 30,
       "STG": 10, "STS": 10, "ST": 10}


def load(lib, name):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        out, on = [], False
        lines = txt.splitlines()
        i = 0
        while i < len(lines):
            ln = lines[i]
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                if on:
                    break
                on = name in m.group(1)
            elif on:
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", ln)
                if m and i + 1 < len(lines):
                    m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
                    hi = int(m2.group(1), 16)
                    text = m.group(2).strip()
                    pred = None
                    mp = re.match(r"(@!?U?P\d+)\s+(.*)", text)
                    if mp:
                        pred, text = mp.group(1), mp.group(2)
                    op = text.split()[0]
                    out.append(dict(addr=int(m.group(1), 16), text=text, pred=pred, op=op, base=op.split(".")[0],
                                    stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1, wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7,
                                    wait=(hi >> 52) & 0x3F))
                    i += 1
            i += 1
        if out:
            return out
    raise SystemExit("kernel %s not found in %s" % (name, lib))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--lib", default=os.path.join(ROOT, "plen_ml_walk_b200", "libplen_b200.so"))
    ap.add_argument("--list-branches", action="store_true")
    ap.add_argument("--from", dest="lo", default=None)
    ap.add_argument("--to", dest="hi", default=None)
    ap.add_argument("--take", nargs="*", default=[])
    ap.add_argument("--lat-shfl", type=int, default=LAT["SHFL"])
    ap.add_argument("--lat-lds", type=int, default=LAT["LDS"])
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    LAT["SHFL"], LAT["LDS"] = a.lat_shfl, a.lat_lds
    ins = load(a.lib, a.kernel)
    by_addr = {x["addr"]: k for k, x in enumerate(ins)}
    if a.list_branches:
        for x in ins:
            if x["base"] in ("BRA", "BRX", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "WARPSYNC"):
                print("%06x  %-6s %s" % (x["addr"], x["pred"] or "", x["text"]))
        return
    lo = int(a.lo, 16) if a.lo else ins[0]["addr"]
    hi = int(a.hi, 16) if a.hi else ins[-1]["addr"]
    take = {int(t, 16) for t in a.take}
    T, n, sb_done, sb_src = 0, 0, [0] * 6, [""] * 6
    exposed = collections.Counter()
    stall_sum = 0
    k = by_addr[lo]
    while k < len(ins) and ins[k]["addr"] <= hi:
        x = ins[k]
        arm, src = 0, ""
        for s in range(6):
            if (x["wait"] >> s) & 1 and sb_done[s] > arm:
                arm, src = sb_done[s], sb_src[s]
        t_issue = T
        if arm > t_issue:
            exposed[src] += arm - t_issue
            t_issue = arm
        if a.trace:
            print("%7d  %06x %-5s st%-2d w%02x wb%d  %s" % (t_issue, x["addr"], x["pred"] or "", x["stall"], x["wait"], x["wbar"], x["text"][:70]))
        if x["wbar"] < 6:
            d = t_issue + LAT.get(x["base"], 20)
            if d > sb_done[x["wbar"]]:
                sb_done[x["wbar"]], sb_src[x["wbar"]] = d, x["base"]
        if x["rbar"] < 6:
            d = t_issue + 6
            if d > sb_done[x["rbar"]]:
                sb_done[x["rbar"]], sb_src[x["rbar"]] = d, x["base"] + "(rd)"
        T = t_issue + max(1, x["stall"])
        stall_sum += max(1, x["stall"])
        n += 1
        if x["base"] == "BRA" and x["addr"] in take:
            m = re.search(r"0x([0-9a-f]+)", x["text"])
            k = by_addr[int(m.group(1), 16)]
            continue
        k += 1
    print("region %06x..%06x: %d instructions, T_1w = %d cycles (%.2f cycles / instruction); fixed-latency stall fields "
          "sum to %d cycles; exposed scoreboard waits: %s" % (lo, hi, n, T, T / max(1, n), stall_sum,
                                                             ", ".join("%s %d" % kv for kv in exposed.most_common(6)) or "none"))


if __name__ == "__main__":
    main()
