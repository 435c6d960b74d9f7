"""Key figures of one or more .ncu-rep files (read offline with `ncu -i`): duration, DRAM bytes, L2 traffic, occupancy,
issue utilisation and the warp-state (stall) breakdown.  python scripts/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", path, d.get("Kernel Name", "")[:80])
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        for v, k in stalls[:8]:
            print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:8.2f}")
