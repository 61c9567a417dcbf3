"""Turn ncu outputs brought back in gpurun_out/ into the small, committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python profiles/summarize.py kernels  gpurun_out/prof_r1_top.ncu-rep profiles/r1_kernels.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    mu = h.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] in ("ns", "nsecond") else v   # -> microseconds
        name = r[kn].split("(")[0].replace("snnflow::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                f"(cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write(f"total {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:70]}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |\n")
    print(open(out).read())


def kernels(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    seen = defaultdict(int)
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\nOne row per captured launch (first two per kernel).\n\n")
        f.write("| kernel | " + " | ".join(label for _, label in KEYS) + " |\n|---|" + "---|" * len(KEYS) + "\n")
        for r in rows[2:]:
            name = r[h.index("Kernel Name")].split("(")[0].replace("snnflow::", "")
            seen[name] += 1
            if seen[name] > 2:
                continue
            cells = []
            for k, _ in KEYS:
                if k in h:
                    i = h.index(k)
                    cells.append(f"{r[i]} {units[i]}".strip())
                else:
                    cells.append("-")
            f.write(f"| `{name}` | " + " | ".join(cells) + " |\n")
    print(open(out).read())


BENCH_NAME = [("wt_fwd_kernel<1", "win_fwd_seq"), ("wt_fwd_kernel<(bool)1", "win_fwd_seq"), ("wt_fwd_kernel<0", "win_fwd_rec"),
              ("wt_fwd_kernel<(bool)0", "win_fwd_rec"), ("wt_dgpw", "win_dgrad_pw"), ("wt_dgrad", "win_dgrad"), ("win_reduce", "win_reduce"), ("wt_recbwd", "win_rec_bwd"),
              ("wg_planes", "win_wgrad"), ("pw_seq", "win_pw_seq")]


def traffic(out, *paths):
    """profiles/traffic.json: DRAM bytes (read + write) per launch of each window kernel, mean over the captured launches,
    keyed by the name bench.py's per-launch profiler uses (bench.py reads it into roofline.traffic)."""
    import json
    acc = defaultdict(list)
    for path in paths:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        h, units = rows[0], rows[1]
        ir, iw, kn = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("Kernel Name")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = next((b for k, b in BENCH_NAME if k in r[kn]), None)
            if name:
                acc[name].append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
    res = {k: round(sum(v) / len(v)) for k, v in acc.items()}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    print(res)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], *sys.argv[3:])
    else:
        {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2], sys.argv[3])
