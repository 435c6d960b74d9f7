"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/<name>.csv      (for r1_fused, r1_ea_fwd, r1_wgrad_group)
    python scripts/summarize_profiles.py
"""
import csv, json, os, re, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"), ("sm__cycles_elapsed.max", "cycles"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__inst_executed.sum", "inst")]


def kname(s):
    return re.sub(r"\(.*", "", s).replace("unnamed>::", "").replace("void ", "").replace("pfn::<", "")


def full_sets():
    out = []
    for f in ("r1_fused", "r1_ea_fwd", "r1_wgrad_group"):
        path = f"/tmp/{f}.csv"
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            rec = {"kernel": kname(d["Kernel Name"]), "grid": d["Grid Size"], "block": d["Block Size"]}
            for k, a in WANT:
                rec[a] = f"{d.get(k)} {u.get(k, '')}".strip()
            st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v.replace(",", ""))) for k, v in d.items()
                  if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and v not in ("", "n/a")]
            tot = sum(v for _, v in st) or 1
            rec["stalls"] = ", ".join(f"{k} {v / tot:.0%}" for k, v in sorted(st, key=lambda x: -x[1])[:5])
            out.append(rec)
    return out


def launch_list():
    src = os.path.join(ROOT, "gpurun_out", "r1_launches_fused.csv")
    shutil.copy(src, os.path.join(OUT, "r1_launches_ncu.csv"))
    lines = [l for l in open(src) if l.startswith('"')]
    seq = [(kname(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3) for r in csv.DictReader(lines)]
    starts = [i for i, s in enumerate(seq) if "k_find_reverse" in s[0]]
    return seq[starts[2]:starts[3]]


def main():
    full = full_sets()
    json.dump(full, open(os.path.join(OUT, "r1_ncu_full_summary.json"), "w"), indent=1)
    step = launch_list()
    tot = sum(s[3] for s in step)
    md = ["# Round 1 — ncu evidence (B200, sm_100a)\n",
          "All captures: `gpurun` on one B200, `ncu --clock-control none`, bench workload (case118v2 × 128 graphs, configs/standard.json, "
          "train mode).  Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `kernel_time`, not absolutes.  "
          "Regenerate with `scripts/summarize_profiles.py`.\n",
          "## 1. Launch list of one optimisation step (`profiles/r1_launches_ncu.csv`)\n",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph`; "
          "the 3rd step of the capture:\n",
          "| # | kernel | grid | block | µs | share |\n|---|---|---|---|---|---|"]
    for i, s in enumerate(step):
        md.append(f"| {i} | `{s[0]}` | {s[1]} | {s[2]} | {s[3]:.2f} | {s[3] / tot:.1%} |")
    md.append(f"| | **total** | | | **{tot:.1f}** | {len(step)} launches |\n")
    md.append("`k_mpn_fused_fwd<HB, MODE>`: MODE 0 = whole forward, 3 = whole backward data path (1 / 2 = one TAGConv / EdgeAggregation "
              "backward per launch, `PFN_BWD_CHAIN=0`), csrc/fused_fwd.cu.  bench.py (CUDA events, warm, 200 steps): 0.734 ms/step; the shares agree "
              "with its `kernel_time` (forward kernel 28 %, backward tile kernel 38 %, grouped weight gradients 29 %, graph prep 5 %).\n")
    md.append("## 2. `ncu --set full` of the main kernels (`profiles/r1_ncu_full_summary.json`)\n")
    md.append("| kernel | µs | DRAM read / written | tensor pipe | issue slots | LSU pipe | smem wavefronts | regs | top stall reasons |\n|---|---|---|---|---|---|---|---|---|")
    seen = set()
    for r in full:
        if r["kernel"] in seen:
            continue
        seen.add(r["kernel"])
        md.append(f"| `{r['kernel']}` {r['grid']}×{r['block']} | {float(r['time'].split()[0]):.1f} | {r['dram_rd']} / {r['dram_wr']} | "
                  f"{float(r['tensor%'].split()[0]):.1f} % | {float(r['issue%'].split()[0]):.1f} % | {float(r['lsu%'].split()[0]):.1f} % | "
                  f"{int(r['smem_wavefronts']) / 1e6:.1f} M | {r['regs'].split()[0]} | {r['stalls']} |")
    md.append(open(os.path.join(OUT, "r1_summary_notes.md")).read())
    open(os.path.join(OUT, "r1_summary.md"), "w").write("\n".join(md))
    print("wrote", os.path.join(OUT, "r1_summary.md"), f"({len(step)} launches, {tot:.1f} us)")


if __name__ == "__main__":
    main()
