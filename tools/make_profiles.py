"""Regenerates the profiles/ summaries of the current round from the files a measurement run left in gpurun_out/."""
import collections
import csv
import json
import os
import shutil
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = "gpurun_out"
for src, dst in (("bench_%s_final_g1_2p20.json", "%s_bench_g1_2p20.json"), ("bench_%s_final_g1_2p24.json", "%s_bench_g1_2p24.json"),
                 ("bench_%s_final_grumpkin_2p20.json", "%s_bench_grumpkin_2p20.json"), ("bench_%s_final_reference.json", "%s_bench_reference_arm.json"), ("bench_%s_final_g2_2p20.json", "%s_bench_g2_2p20.json"),
                 ("launches_%s_final.csv", "%s_launches_bench_2p20.csv")):
    if os.path.exists(os.path.join(G, src % R)):
        shutil.copy(os.path.join(G, src % R), os.path.join("profiles", dst % R))
rows = [r for r in csv.reader(open(f"profiles/{R}_launches_bench_2p20.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
# one segment per MSM: a segment starts at k_count; the host-buffer legs of bench.py cut a call into pipelined pieces (smaller MSMs), so only
# the segments whose accumulate kernel is full-size (>= 0.9 x the longest) enter the per-MSM table
segs, cur, setup = [], None, collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("kgr::", "")[:44]
    val = float(r[vi].replace(",", ""))
    if name.startswith("k_count"):
        cur = []
        segs.append(cur)
    if cur is not None and any(name.startswith(m) for m in ["k_count", "k_scan_tiles", "k_scan_sums", "k_scan_add", "k_fill", "k_accumulate", "k_fixup<", "k_fixup_long",
                                                             "k_reduce", "k_weight", "k_tree_sum", "k_fold<", "k_fold_tail", "k_vsum1", "k_vsum2", "k_fold_combine"]):
        cur.append((name, val))
    else:
        setup.setdefault(name, []).append(val)
acc_of = lambda seg: max([v for n, v in seg if n.startswith("k_accumulate")] or [0.0])
full = max(acc_of(sg) for sg in segs)
segs_full = [sg for sg in segs if acc_of(sg) >= 0.9 * full]
agg = collections.OrderedDict()
for sg in segs_full:
    for n, v in sg:
        agg.setdefault(n, []).append(v)
nacc = len(segs_full)
per = {k: sum(v) / nacc for k, v in agg.items()}
tot = sum(per.values())
lines = ["| kernel | launches per MSM | mean per launch (us) | per MSM (us) | share |", "|---|---|---|---|---|"]
for k, v in agg.items():
    if k in per:
        lines.append(f"| `{k}` | {len(v) / nacc:.0f} | {sum(v) / len(v) / 1e3:.1f} | {per[k] / 1e3:.1f} | {100 * per[k] / tot:.1f} % |")
lines.append(f"| **sum** | | | {tot / 1e3:.1f} | |")
other = [f"- `{k}`: {len(v)} launches, mean {sum(v) / len(v) / 1e3:.1f} us" for k, v in setup.items()]
other.append(f"- {len(segs) - len(segs_full)} smaller MSM segments (pieces of the host-buffer calls) are not in the table; {nacc} full-size MSMs are")
b = json.load(open(f"profiles/{R}_bench_g1_2p20.json"))
ph = b["phases_ms"]
acc_key = [x for x in per if x.startswith("k_accumulate")][0]
open(f"profiles/{R}_launches.md", "w").write(f"""# Round {R[1:]} — ncu launch list of one `bench.py` run (BN254 G1, 2^20 points, 1 x B200)

Command (on the GPU box): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_{R}_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-precompute`
Raw list: `profiles/{R}_launches_bench_2p20.csv`.  Times under ncu are serialised and cold-cache; compare shares.

""" + "\n".join(lines) + """

Other launches in the same capture (setup, not part of an MSM):
""" + "\n".join(other) + f"""

The same phases measured live with CUDA events inside `bench.py` (`profiles/{R}_bench_g1_2p20.json`, `phases_ms`):
count {ph['count']:.3f}, scan {ph['scan']:.3f}, fill {ph['fill']:.3f}, accumulate {ph['accumulate']:.3f}, fixup {ph['fixup']:.3f}, reduce {ph['reduce']:.3f} (folds + upper-half sums + combine + D2H), host finish {ph['host_finish']:.3f}; total {ph['total']:.3f} ms.
`k_accumulate` is {100 * per[acc_key] / tot:.1f} % of the MSM in the ncu list and {100 * ph['accumulate'] / ph['total']:.1f} % in the live timing: the shares agree.
""")
for name in ("g1_2p20", "g1_2p24", "grumpkin_2p20", "reference_arm"):
    d = json.load(open(f"profiles/{R}_bench_{name}.json"))
    print(name, round(d["value"], 2), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), (d.get("roofline") or {}).get("frac"))
