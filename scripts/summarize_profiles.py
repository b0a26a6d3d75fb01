"""Turns the ncu reports that scripts/gpu_evidence.sh left in gpurun_out/ into the tracked summaries
under profiles/: a per-kernel table (duration, DRAM bytes, issue/occupancy, pipe utilisation), the
launch list of the bench command with each kernel's share of the step, and traffic.json (DRAM bytes
per launch of every kernel family -- bench.py reads it for roofline.traffic).

Usage (in the build container, which has ncu but no GPU):  python scripts/summarize_profiles.py r1
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]

FAMILY = [  # kernel-name substring -> bench.py breakdown family
    ("preprocess_fwd", "preprocess_fwd"), ("preprocess_bwd", "preprocess_bwd"), ("blend_fwd", "blend_fwd"),
    ("blend_bwd", "blend_bwd"), ("emit_kernel", "emit"), ("ranges_kernel", "tile_ranges"),
    ("rs_pass_kernel<unsigned short", "tile_sort"), ("rs_tile_hist", "tile_sort"), ("rs_tile_scan", "tile_sort"),
    ("rs_pass_kernel<unsigned int", "depth_sort_scan"), ("rs_hist_kernel", "depth_sort_scan"),
    ("scan_gather", "depth_sort_scan"), ("kmeans_assign_seg", "kmeans_assign_seg"), ("kmeans_assign", "kmeans_assign"),
    ("adam_kernel", "adam"),
    ("mask_pack_kernel", "mask_iou"), ("mask_pair_kernel", "mask_iou"), ("footprint_vote", "footprint"),
    ("kmeans_assign_seg", "kmeans_assign_seg"), ("multimem_allreduce", "multimem_allreduce"), ("peer_allreduce", "peer_allreduce"),
]


def family(name):
    for sub, fam in FAMILY:
        if sub in name:
            return fam
    return None


def to_float(v, unit, want):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    if want == "time_us":
        return v * scale.get(unit, 1.0)
    if want.endswith("_MB"):
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    return v


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        d = {"kernel": r[ki]}
        for m, short in METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[short] = to_float(r[i], units[i], short)
        out.append(d)
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    traffic = {}
    table = []
    for part in ("raster", "kmeans", "stage3", "footprint"):
        rep = os.path.join(SRC, f"{tag}_{part}.ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        rows = raw_rows(rep)
        table += rows
    cols = ["kernel"] + [s for _, s in METRICS]
    with open(os.path.join(OUT, f"{tag}_kernels_full.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        for d in table:
            w.writerow([d["kernel"][:80]] + [("%.4g" % d[c]) if c in d else "" for c in cols[1:]])
    # one Stage-1 training step (scripts/gpu_evidence_r2.sh part c): its own table, not part of traffic.json
    for suffix in ("stage1", "stage1_cached"):
        rep = os.path.join(SRC, f"{tag}_{suffix}.ncu-rep")
        if os.path.exists(rep):
            with open(os.path.join(OUT, f"{tag}_{suffix}_kernels_full.csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(cols)
                for d in raw_rows(rep):
                    w.writerow([d["kernel"][:80]] + [("%.4g" % d[c]) if c in d else "" for c in cols[1:]])
    # per-family DRAM traffic per launch group: kernels of one family inside ONE frame are summed
    fam_acc = {}
    for d in table:
        fam = family(d["kernel"])
        if fam is None:
            continue
        a = fam_acc.setdefault(fam, {"bytes": 0.0, "time_us": 0.0, "launches": 0, "names": {}, "insts": 0.0})
        a["insts"] += d.get("warp_insts", 0.0)
        a["bytes"] += (d.get("dram_read_MB", 0.0) + d.get("dram_write_MB", 0.0)) * 1e6
        a["time_us"] += d.get("time_us", 0.0)
        a["launches"] += 1
        a["names"][d["kernel"].split("(")[0][:60]] = a["names"].get(d["kernel"].split("(")[0][:60], 0) + 1
    # number of frames captured = launches of blend_fwd (one per frame)
    frames = max(1, fam_acc.get("blend_fwd", {}).get("launches", 1))
    for fam, a in fam_acc.items():
        per_launch = fam in ("kmeans_assign", "kmeans_assign_seg", "adam", "mask_iou", "footprint")
        n = max(1, a["launches"]) if per_launch else frames
        traffic[fam] = {"dram_bytes_per_frame": a["bytes"] / n, "ncu_time_us_per_frame": a["time_us"] / n,
                        "warp_insts_per_frame": a["insts"] / n,
                        "kernels": a["names"], "frames_captured": n}
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump({"source": f"ncu --set full, gpurun_out/{tag}_raster.ncu-rep + {tag}_kmeans.ncu-rep "
                             "(scripts/gpu_evidence.sh, scripts/summarize_profiles.py)", "tag": tag,
                   "families": traffic}, f, indent=1)
    # launch list of the bench command -> share of each kernel in the step
    lcsv = os.path.join(SRC, f"{tag}_launches.csv")
    if os.path.exists(lcsv):
        rows = [r for r in csv.reader(open(lcsv)) if len(r) > 5]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = {}
        for r in rows[1:]:
            name = r[ki].split("(")[0][:70]
            t = to_float(r[vi], r[ui], "time_us")
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += t
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(OUT, f"{tag}_launch_shares.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "share_of_captured_time"])
            for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                w.writerow([name, n, "%.1f" % t, "%.4f" % (t / tot)])
        import shutil
        shutil.copy(lcsv, os.path.join(OUT, f"{tag}_launches.csv"))
    for fam, a in sorted(traffic.items()):
        print(f"{fam:18s} dram {a['dram_bytes_per_frame'] / 1e6:8.1f} MB/frame  ncu time {a['ncu_time_us_per_frame']:7.1f} us")


if __name__ == "__main__":
    main()
