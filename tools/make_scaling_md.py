"""profiles/r02_scaling.md from profiles/r02_bench_n{1,2,4,8}.json (the last JSON line of each bench.py run)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def main():
    P = lambda n: os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    n1 = last(P(1))
    rows = {n: last(P(n)) for n in (2, 4, 8) if os.path.exists(P(n))}
    b26, b24 = n1["strong_scaling_base"]["2p26"], n1["strong_scaling_base"]["2p24"]
    out = ["# Round 2 — scaling over 1 / 2 / 4 / 8 B200 of one box (`torchrun ... bench.py --gpus N`; files `r02_bench_n{1,2,4,8}.json`)", "",
           "STRONG scaling (BASELINE configs[4]): ONE BN254 G1 MSM over the same seeded vector at every N, sharded evenly; a step = every rank's MSM + `all_gather` of one",
           "96-byte point per rank + the host sum on rank 0; wall clock between barriers, max over ranks.  `resident`: scalars already in HBM; `host scalars`: uploaded",
           "from pinned memory every step (bases registered).  `checksum`: the combined point equals (sum k_i s_i) G with the sum accumulated per rank by the oracle.", "",
           "| N | 2^26 resident ms (Mpoints/s, x vs N = 1) | 2^26 host scalars ms (x) | 2^24 resident ms (x) | 2^24 host scalars ms | checksum 2^26 / 2^24 | one-process `kgr_msm` over N GPUs, 2^26 / 2^24 ms (same point as torchrun) |",
           "|---|---|---|---|---|---|---|",
           f"| 1 | {b26['ms_per_step']:.1f} ({b26['value']:.0f}, 1.00) | {b26['e2e']['ms_per_step']:.1f} (1.00) | {b24['ms_per_step']:.2f} (1.00) | {b24['e2e']['ms_per_step']:.2f} | {b26['checksum_ok']} / {b24['checksum_ok']} | — |"]
    for n, d in sorted(rows.items()):
        s24, ip = d["strong_2p24"], d.get("inprocess", {})
        ipt = (f"{ip['2p26']['ms_per_step']:.1f} / {ip['2p24']['ms_per_step']:.1f} ({ip['2p26']['same_point_as_torchrun']}, {ip['2p24']['same_point_as_torchrun']})"
               if "2p26" in ip else str(ip))
        out.append(f"| {n} | {d['ms_per_step']:.2f} ({d['value']:.0f}, {b26['ms_per_step'] / d['ms_per_step']:.2f}) | {d['e2e']['ms_per_step']:.2f} ({b26['e2e']['ms_per_step'] / d['e2e']['ms_per_step']:.2f}) | "
                   f"{s24['ms_per_step']:.2f} ({b24['ms_per_step'] / s24['ms_per_step']:.2f}) | {s24['e2e']['ms_per_step']:.2f} | {d['checksum_ok']} / {s24['checksum_ok']} | {ipt} |")
    out += ["", "Where the efficiency goes (N = 8, 2^26): a 2^23-point shard runs c = 17 with 15 windows where the 2^26-point MSM runs c = 20 with 13 (+15 % bucket additions per point),",
            "its bucket reduction and sort keep their fixed parts, and gather + host sum add ~0.8 ms to the slowest rank's device time.  No collective is on the data path.", "",
            "WEAK scaling (round 1's figure; an independent 2^20-point MSM per GPU, per-rank CUDA events / wall clock, max over ranks):", "",
            "| N | Mpoints/s resident (x) | end to end, points + scalars uploaded (x) | checksum on every rank |", "|---|---|---|---|",
            f"| 1 | {n1['value']:.0f} | {n1['e2e']['value']:.0f} | {n1['checksum_ok']} |"]
    for n, d in sorted(rows.items()):
        w = d["weak_scaling"]
        out.append(f"| {n} | {w['value']:.0f} ({w['value'] / n1['value']:.2f}) | {w['e2e_value']:.0f} ({w['e2e_value'] / n1['e2e']['value']:.2f}) | {w['checksum_ok_all_ranks']} |")
    out += ["", "The box is one NUMA node with 32 vCPUs (`nvidia-smi topo -m`: every GPU `CPU Affinity 0-31, NUMA Affinity 0`), so NUMA-local staging (VERDICT r1 item 2) cannot be",
            "arranged from inside it."]
    open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
