"""Condense the ncu reports of tools/r2_profiles.sh (gpurun_out/r2_full_*.ncu-rep) into profiles/r2/ncu_full_summary.json and
refresh profiles/kernel_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = {
    "gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__shared_mem_per_block_dynamic": "dyn_smem", "launch__cluster_size": "cluster",
}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    res = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("ctvio::", "")
        rec = {}
        for m, key in METRICS.items():
            if m not in idx:
                continue
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                continue
            rec[key] = v * UNIT.get(units[idx[m]], 1.0) if key in ("duration_us", "dram_read", "dram_write") else v
        res[name] = rec  # the last (warmest) launch of each kernel wins
    return res


summary = {}
for tag in ("c2", "c4", "c5"):
    rep = os.path.join(ROOT, "gpurun_out", f"r2_full_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    for k, v in load(rep).items():
        v["dram_bytes"] = v.get("dram_read", 0.0) + v.get("dram_write", 0.0)
        summary[f"{tag}:{k}"] = v
json.dump(summary, open(os.path.join(ROOT, "profiles", "r2", "ncu_full_summary.json"), "w"), indent=1)
tr_path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
tr = json.load(open(tr_path)) if os.path.exists(tr_path) else {}
for k, v in summary.items():
    tag, name = k.split(":", 1)
    for short, pat in (("chol_dag", "chol_dag_kernel"), ("visual", "visual_kernel<1>"), ("schur_tile", "schur_tile_kernel"),
                       ("jacobi_blocked", "jacobi_blocked_kernel")):
        if name.startswith(pat):
            tr[f"{tag}_{short}_dram_bytes_per_launch"] = v["dram_bytes"]
tr["source"] = "profiles/r2/ncu_full_summary.json (ncu --set full --clock-control none, tools/r2_profiles.sh + tools/ncu_summary.py, round 2 final kernels)"
json.dump(tr, open(tr_path, "w"), indent=1)
print(json.dumps({k: {kk: v.get(kk) for kk in ("duration_us", "dram_bytes", "regs", "grid", "cluster")} for k, v in summary.items()}, indent=1))
