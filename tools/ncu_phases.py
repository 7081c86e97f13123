"""Per-phase summary of an `ncu --page source --csv` export: the SASS listing is cut at every block-wide barrier
(BAR.SYNC) and executed instructions, stall samples and shared-memory wavefronts are summed per segment.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_phases.py src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
segs, cur = [], None


def new(label):
    return {"label": label, "inst": 0, "samples": 0, "wave": 0, "ideal": 0, "n": 0, "ops": {}, **{s: 0 for s in stalls}}


cur = new("start")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[1] if src.startswith("@") and len(src.split()) > 1 else src.split()[0]
    f = lambda k: int(float(r[ix[k]] or 0))
    cur["inst"] += f("Instructions Executed")
    cur["samples"] += f("# Samples")
    cur["wave"] += f("L1 Wavefronts Shared")
    cur["ideal"] += f("L1 Wavefronts Shared Ideal")
    cur["n"] += 1
    base = op.split(".")[0]
    cur["ops"][base] = cur["ops"].get(base, 0) + f("Instructions Executed")
    for s in stalls:
        cur[s] += f(s)
    if op.startswith("BAR"):
        segs.append(cur)
        cur = new("after %s @%s" % (op, r[ix["Address"]][-5:]))
segs.append(cur)
ti = sum(s["inst"] for s in segs) or 1
ts = sum(s["samples"] for s in segs) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
for s in segs:
    if s["inst"] / ti < min_share and s["samples"] / ts < min_share:
        continue
    top = sorted(((s[k], k[6:]) for k in stalls), reverse=True)[:6]
    ops = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:9]
    print("%-28s sass %5d  inst %5.1f%%  samples %5.1f%%  smem wavefronts %9d (ideal %9d)" % (
        s["label"], s["n"], 100 * s["inst"] / ti, 100 * s["samples"] / ts, s["wave"], s["ideal"]))
    print("      stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / max(1, s["samples"])) for v, k in top))
    print("      ops:    " + ", ".join("%s %.1f%%" % (k, 100 * v / max(1, s["inst"])) for k, v in ops))
