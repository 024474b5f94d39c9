"""Turn gpurun_out/<tag>_prof.ncu-rep and <tag>_launches.csv into small committed summaries under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]

KEEP += ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
         "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]

# secondary kernels: gpurun_out/<tag>_sec_<name>.ncu-rep -> profiles/<tag>_sec_kernels_ncu_full.csv (one column per kernel)
sec = sorted((ROOT / "gpurun_out").glob(f"{tag}_sec_*.ncu-rep"))
if sec:
    cols = {}
    for repf in sec:
        raw = subprocess.run(["ncu", "-i", str(repf), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        cols[repf.stem.replace(f"{tag}_sec_", "")] = (hdr, units, data[0])
    if cols:
        stall_all = sorted({h for hdr, _, _ in cols.values() for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h})
        with open(out / f"{tag}_sec_kernels_ncu_full.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + list(cols))
            for k in KEEP + stall_all:
                unit, vals = "", []
                for hdr, units, row in cols.values():
                    if k in hdr:
                        i = hdr.index(k)
                        unit = units[i]
                        vals.append(row[i])
                    else:
                        vals.append("")
                if any(vals):
                    w.writerow([k, unit] + vals)
        print("wrote", out / f"{tag}_sec_kernels_ncu_full.csv")

rep = ROOT / "gpurun_out" / f"{tag}_prof.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    with open(out / f"{tag}_stage_kernel_ncu_full.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for k in KEEP + stall:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
    print("wrote", out / f"{tag}_stage_kernel_ncu_full.csv")

lc = ROOT / "gpurun_out" / f"{tag}_launches.csv"
if lc.exists():
    lines = [l for l in lc.read_text().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out / f"{tag}_launch_shares.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_of_gpu_time"])
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, c, round(t / 1e6, 4), round(t / tot, 4)])
    (out / f"{tag}_launches.csv").write_text("\n".join(lines) + "\n")
    print(open(out / f"{tag}_launch_shares.csv").read())
