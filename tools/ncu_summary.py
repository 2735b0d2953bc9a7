"""Summarise an .ncu-rep (ncu --set full) into the small JSON kept under profiles/:

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "what this capture is" > profiles/x_ncu_full.json

One record per launch with the metrics DESIGN.md quotes (duration, DRAM bytes / throughput, tensor-pipe and issue
activity, warps, registers, instructions, occupancy limits, grid)."""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__cluster_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rep, what = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
launches = []
for r in data:
    d = {"kernel": r[hdr.index("Kernel Name")][:100]}
    for k in KEYS:
        idx = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
        if idx and r[idx[0]] != "":
            d[k] = f"{r[idx[0]]} {units[idx[0]]}".strip()
    launches.append(d)
print(json.dumps({"what": what, "launches": launches}, indent=1))
