#!/usr/bin/env python
"""Turns the .ncu-rep captures in gpurun_out/ into profiles/<round>_ncu_summary.{json,md} (run in the container:
ncu reads reports without a GPU). bench.py reads profiles/ncu_summary.json for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
}
UNIT_SCALE = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def read(rep):
    """rep: a .ncu-rep capture, or the `ncu -i ... --page raw --csv` export of one (made on the GPU box when the captures
    themselves are too large to bring back)."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[name] = v * UNIT_SCALE.get(units[i], 1.0) if name.endswith(("_MB", "_us")) else v
        res.append(d)
    return res


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out_dir = os.path.join(ROOT, "gpurun_out")
    summary = {}
    for wl in ("life", "mean", "mean_halo", "kernel", "kernel_fma", "circle", "positional", "scatter", "window3d", "diffusion", "diffusion2"):
        rep = os.path.join(out_dir, f"{rnd}_{wl}.ncu-rep")
        if not os.path.exists(rep):
            rep = os.path.join(out_dir, f"{rnd}_{wl}_raw.csv")
        if not os.path.exists(rep):
            continue
        ks = read(rep)
        if not ks:
            continue
        k = ks[0]
        k["dram_bytes_per_launch"] = (k.get("dram_read_MB", 0) + k.get("dram_write_MB", 0)) * 1e6
        summary[wl] = k
    json.dump(summary, open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_summary.json"), "w"), indent=1)
    # profiles/ncu_summary.json (read by bench.py for `roofline.traffic`) keeps the latest capture of every workload
    latest_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    latest = json.load(open(latest_path)) if os.path.exists(latest_path) else {}
    latest.update(summary)
    json.dump(latest, open(latest_path, "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_summary.md"), "w") as f:
        f.write(f"# ncu --set full summaries ({rnd}); one launch per workload, cold cache, serialised\n\n")
        f.write("| workload | kernel | duration µs | DRAM read MB | DRAM write MB | DRAM % of peak | regs | grid x block | warps active % | issue active % | L2 hit % |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
        for wl, k in summary.items():
            f.write(f"| {wl} | {k['kernel'][:48]} | {k.get('duration_us', 0):.1f} | {k.get('dram_read_MB', 0):.1f} | "
                    f"{k.get('dram_write_MB', 0):.1f} | {k.get('dram_pct_of_peak', 0):.1f} | {int(k.get('registers', 0))} | "
                    f"{int(k.get('grid', 0))} x {int(k.get('block', 0))} | {k.get('warps_active_pct', 0):.1f} | "
                    f"{k.get('issue_active_pct', 0):.1f} | {k.get('l2_hit_pct', 0):.1f} |\n")
    print(json.dumps(summary, indent=1)[:3000])


if __name__ == "__main__":
    main()
