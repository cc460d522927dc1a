"""Round 2: regenerates profiles/r02_launches.md, profiles/r02_ncu_2p20_final_raw.csv and profiles/ncu_traffic.json from the files the final
measurement run left in gpurun_out/ (r02_launches_bench_2p20_final.csv, r02_ncu_full_2p20_raw_all.csv, r02_bench_n1_final.json)."""
import collections
import csv
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)
shutil.copy(os.path.join(G, "r02_bench_n1_final.json"), os.path.join(P, "r02_bench_n1.json"))
shutil.copy(os.path.join(G, "r02_bench_ref_n1_final.json"), os.path.join(P, "r02_bench_reference_arm_n1.json"))
shutil.copy(os.path.join(G, "r02_launches_bench_2p20_final.csv"), os.path.join(P, "r02_launches_bench_2p20.csv"))

# ---- launch list -> shares -------------------------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(P, "r02_launches_bench_2p20.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
nm = lambda r: r[ki].split("(")[0].replace("void ", "").replace("kgr::", "").split("<")[0]
MSM = ("k_sort_", "k_count", "k_scan_", "k_fill", "k_affine_", "k_batch_inv", "k_accumulate", "k_fixup", "k_fold", "k_vsum", "k_bucket_merge", "k_reduce", "k_weight", "k_tree_sum")
segs, cur, other = [], None, collections.OrderedDict()
for r in rows[1:]:
    name, val = nm(r), float(r[vi].replace(",", ""))
    if name in ("k_sort_digits", "k_count"):
        cur = []
        segs.append(cur)
    if cur is not None and name.startswith(MSM):
        cur.append((name, val))
    else:
        other.setdefault(name, []).append(val)
fulls = [s for s in segs if not any(k == "k_bucket_merge" for k, v in s) and len(s) >= 30]
agg = collections.OrderedDict()
for s in fulls:
    for k, v in s:
        agg.setdefault(k, []).append(v)
n = len(fulls)
per = {k: sum(v) / n for k, v in agg.items()}
tot = sum(per.values())
lines = ["| kernel | launches per MSM | mean per launch (us) | per MSM (us) | share |", "|---|---|---|---|---|"]
for k, v in agg.items():
    lines.append(f"| `{k}` | {len(v) / n:.0f} | {sum(v) / len(v) / 1e3:.1f} | {per[k] / 1e3:.1f} | {100 * per[k] / tot:.1f} % |")
lines.append(f"| **sum** | {sum(len(v) for v in agg.values()) / n:.0f} | | {tot / 1e3:.1f} | |")
grp = lambda pre: sum(v for k, v in per.items() if k.startswith(pre))
g_sort, g_aff, g_acc, g_fix, g_red = grp(("k_sort_",)), grp(("k_affine_", "k_batch_inv", "k_scan_")), grp(("k_accumulate",)), grp(("k_fixup",)), grp(("k_fold", "k_vsum"))
b = json.load(open(os.path.join(P, "r02_bench_n1.json")))
ph = b["phases_ms"]
tl = ph["total"] - ph["host_finish"]
sort_live = ph["count"] + ph["scan"] + ph["fill"]
gl = ["| phase | ncu, per MSM (us) | share under ncu | live CUDA events in `bench.py` (ms) | live share |", "|---|---|---|---|---|",
      f"| sort | {g_sort / 1e3:.0f} | {100 * g_sort / tot:.1f} % | {sort_live:.3f} | {100 * sort_live / tl:.1f} % |",
      f"| accumulate phase (batched-affine levels {g_aff / 1e3:.0f} + `k_accumulate` {g_acc / 1e3:.0f}) | {(g_aff + g_acc) / 1e3:.0f} | {100 * (g_aff + g_acc) / tot:.1f} % | {ph['accumulate']:.3f} | {100 * ph['accumulate'] / tl:.1f} % |",
      f"| fix-up | {g_fix / 1e3:.0f} | {100 * g_fix / tot:.1f} % | {ph['fixup']:.3f} | {100 * ph['fixup'] / tl:.1f} % |",
      f"| reduce | {g_red / 1e3:.0f} | {100 * g_red / tot:.1f} % | {ph['reduce']:.3f} | {100 * ph['reduce'] / tl:.1f} % |"]
oth = [f"- `{k}`: {len(v)} launches, mean {sum(v) / len(v) / 1e3:.1f} us" for k, v in other.items()]
open(os.path.join(P, "r02_launches.md"), "w").write(f"""# Round 2 — ncu launch list of one `bench.py` run (BN254 G1, 2^20 points, 1 x B200, final build of the round)

Command (GPU box): `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_2p20_final.csv python bench.py --steps 2 --warmup 1 --quick`
Raw list: `profiles/r02_launches_bench_2p20.csv`; this file is made by `tools/make_profiles_r02.py`.  Times under ncu are serialised and cold-cache; compare SHARES with the live timing.
{len(segs)} MSM segments in the capture; the table averages the {n} resident-input MSMs of 2^20 points (the host-buffer legs run as streamed pieces
— `k_bucket_merge` — and are left out).

""" + "\n".join(lines) + """

""" + "\n".join(gl) + f"""

Live total without the host finish: {tl:.3f} ms.  The dominant phase is {100 * (g_aff + g_acc) / tot:.1f} % of the MSM under ncu and {100 * ph['accumulate'] / tl:.1f} % live: the shares agree.
The reduction's kernels sum to MORE under ncu ({g_red / 1e3:.0f} us) than its live phase ({ph['reduce'] * 1e3:.0f} us): live, the early `k_vsum1` runs beside the fold
levels on a second stream, and the short latency-bound kernels do not pay the profiler's serialisation.

Other launches in the same capture (generators, L2 flush, the extras of the streamed pieces):
""" + "\n".join(oth) + "\n")

# ---- ncu --set full: the kernels of the last MSM of the capture; DRAM traffic of the dominant phase -----------------------------------------
rows = list(csv.reader(open(os.path.join(G, "r02_ncu_full_2p20_raw_all.csv"))))
hdr, units, data = rows[0], rows[1], rows[2:]
kn = hdr.index("Kernel Name")
last = [i for i, r in enumerate(data) if "k_sort_digits" in r[kn]][-1]
sel = data[last:]
csv.writer(open(os.path.join(P, "r02_ncu_2p20_final_raw.csv"), "w")).writerows([hdr, units] + sel)
col = lambda name: hdr.index(name)
tobytes = lambda val, unit: float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
traffic = collections.OrderedDict()
for r in sel:
    name = r[kn].split("(")[0].replace("void ", "").replace("kgr::", "").split("<")[0]
    traffic.setdefault(name, [0.0, 0])
    traffic[name][0] += tobytes(r[rd], units[rd]) + tobytes(r[wr], units[wr])
    traffic[name][1] += 1
from bench import kernel_sources_sha  # noqa: E402
acc_phase = sum(v[0] for k, v in traffic.items() if k.startswith(("k_affine_", "k_batch_inv", "k_scan_", "k_accumulate")))
json.dump({"workload": "bn254_g1 MSM, 2^20 points (tools/one_msm.py 20), one launch of the pipeline", "capture": "ncu --set full --clock-control none, profiles/r02_ncu_2p20_final_raw.csv",
           "kernel_sources_sha": kernel_sources_sha(), "accumulate_phase_dram_bytes": acc_phase,
           "per_kernel_dram_bytes": {k: {"bytes": v[0], "launches": v[1]} for k, v in traffic.items()}},
          open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print("accumulate phase DRAM bytes per MSM:", acc_phase / 1e9, "GB")
