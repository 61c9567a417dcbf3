"""profiles/traffic.json from an ncu metrics pass: average DRAM bytes (read + write) and device time per launch of every
window-engine / encode / IWE kernel, keyed by the name the library's live profiler uses (bench.py reads it into
`roofline.traffic`).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/traffic_train.csv python bench.py --steps 1 --warmup 1 --no-graph --no-eval --no-cpu-baseline
    python profiles/make_traffic.py train=gpurun_out/traffic_train.csv eval=gpurun_out/traffic_eval.csv micro=...

Sections: `train` (BASELINE configs[1] step), `eval` (configs[2] window), `micro` (configs[4]); only the LAST launches of a
kernel are averaged (`--tail N`, default: the second half), i.e. a steady-state step after the warm-up.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel function name (ncu) -> name used by the library's profiler (prof_begin)
RULES = [
    (r"wt_fwd_kernel<(\(bool\))?(1|true)", "win_fwd_seq"), (r"wt_fwd_kernel<(\(bool\))?(0|false)", "win_fwd_rec"),
    (r"wt_recbwd_kernel", "win_rec_bwd"), (r"wt_dgpw_kernel", "win_dgrad_pw"), (r"wt_dgrad_kernel", "win_dgrad"),
    (r"wg_planes_kernel", "win_wgrad"), (r"pw_seq", "win_pw_seq"), (r"win_reduce_kernel", "win_reduce"),
    (r"pred_fwd_planes", "win_pred_fwd"), (r"pack_planes_kernel", "win_pack"), (r"window_pack_weights", "win_pack_weights"),
    (r"encode_cnt_kernel", "encode_cnt"), (r"encode_voxel_kernel", "encode_voxel"), (r"encode_image_last", "encode_image_last"),
    (r"encode_image_acc", "encode_image_acc"), (r"iwe_splat_fwd_kernel", "iwe_splat_fwd"), (r"iwe_splat_bwd_kernel", "iwe_splat_bwd"),
    (r"iwe_fix_to_float", "iwe_fix_to_float"), (r"flow_gather_fwd", "flow_gather_fwd"), (r"wl_gather_splat", "loss_gather_splat"),
    (r"wl_splat_bwd", "loss_splat_bwd"), (r"wl_sums", "loss_sums"), (r"wl_gimg", "loss_gimg"), (r"wl_final", "loss_final"),
    (r"wl_smooth", "loss_smooth"), (r"ld_scatter", "loader_scatter"),
    (r"convlif_fwd_tc_kernel", "convlif_fwd_tc"), (r"convlif_fwd_simt_kernel", "convlif_fwd"), (r"dp_allreduce", "dp_allreduce"),
    (r"opt_clip_adam", "opt_clip_adam"), (r"opt_sumsq", "opt_sumsq"),
]


def parse(path):
    launches = {}   # id -> dict(name, metrics)
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[1:]:
        d = launches.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1,
                 "msecond": 1e6}.get(unit, 1)
        d[r[ix["Metric Name"]]] = v * scale
    return [launches[k] for k in sorted(launches, key=int)]


def summarize(launches):
    by = {}
    for l in launches:
        for pat, name in RULES:
            if re.search(pat, l["name"]):
                by.setdefault(name, []).append(l)
                break
    out = {}
    for name, ls in by.items():
        tail = ls[len(ls) // 2:] if len(ls) > 1 else ls
        n = len(tail)
        out[name] = {
            "dram_bytes_per_launch": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in tail) / n,
            "dram_read_bytes_per_launch": sum(l.get("dram__bytes_read.sum", 0) for l in tail) / n,
            "ncu_us_per_launch": sum(l.get("gpu__time_duration.sum", 0) for l in tail) / n / 1e3,
            "launches_averaged": n,
        }
    return out


if __name__ == "__main__":
    res = {}
    dst = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(dst):
        try:
            old = json.load(open(dst))
            res = {k: v for k, v in old.items() if isinstance(v, dict) and "dram_bytes_per_launch" not in v}
        except Exception:  # noqa: BLE001
            res = {}
    for arg in sys.argv[1:]:
        sec, path = arg.split("=", 1)
        res[sec] = summarize(parse(path))
    with open(dst, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    for sec, d in res.items():
        for k, v in sorted(d.items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches_averaged"]):
            print(f"{sec:6s} {k:20s} {v['dram_bytes_per_launch'] / 1e6:9.2f} MB/launch  {v['ncu_us_per_launch']:8.1f} us  x{v['launches_averaged']}")
