"""Summarise an Nsight Compute report for profiles/: `python scripts/ncu_summary.py rep.ncu-rep "command"`.
Reads the report with `ncu -i ... --page raw --csv` (works on a box without a GPU) and keeps, per
captured kernel launch, the metrics the DESIGN tables quote: DRAM bytes, duration, pipe
utilisation (FP64 / DMMA), issue activity, occupancy limits and the stall breakdown."""
import csv, io, json, subprocess, sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, command = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        rec = dict(zip(header, r))
        out = {"Kernel Name": rec.get("Kernel Name", "")}
        for k in KEEP:
            if k in rec and rec[k] != "":
                out[k] = f"{rec[k]} {units[header.index(k)]}".strip()
        stalls = {}
        for k, v in rec.items():
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a"):
                val = float(v.replace(",", ""))
                if val >= 0.05:
                    stalls[k[len(STALL):-len("_per_issue_active.ratio")]] = round(val, 3)
        out["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        kernels.append(out)
    if "--last-per-kernel" in sys.argv:   # one record per kernel: its last captured launch
        last = {}
        for k in kernels:
            last[k["Kernel Name"]] = k
        kernels = list(last.values())
    print(json.dumps({"command": command, "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
