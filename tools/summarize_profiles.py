#!/usr/bin/env python3
"""Turn the ncu outputs a GPU trip brings back (gpurun_out/<dir>/) into the tracked summaries under
profiles/:  python tools/summarize_profiles.py r2_v3 [gpurun_out subdirectory] [batch of the captured run]
  launches.csv        -> profiles/<tag>_launches.csv, <tag>_launch_shares.txt
  prof_sweeps.ncu-rep -> profiles/<tag>_ncu_full_raw.csv (ncu --page raw --csv), <tag>_ncu_sweeps_summary.json
                         (incl. stall reasons, shared-memory bank conflicts, sectors per request of the global
                         loads), dram_traffic.json (dram bytes per launch; read by bench.py)"""
import collections
import csv
import json
import pathlib
import shutil
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
NAMES = {"k_multiply_rows": "S5_multiply_rows", "k_propagate_cols": "S6_propagate_cols", "k_density_rows": "S1_density_rows",
         "k_potential_cols": "S2_potential_cols", "k_transmit_rows": "S3_transmit_rows", "k_bandlimit_cols": "S4_bandlimit_cols"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}


def main(tag, sub="", batch=10):
    out, prof = ROOT / "gpurun_out" / sub, ROOT / "profiles"
    launch_shares(tag, out, prof)
    summarize_rep(out / "prof_sweeps.ncu-rep", prof, f"{tag}_ncu_full_raw.csv", f"{tag}_ncu_sweeps_summary.json", batch, True, tag)
    if (out / "prof_sweeps_2048.ncu-rep").exists():      # the same capture on the Au 2048^2 workload
        summarize_rep(out / "prof_sweeps_2048.ncu-rep", prof, None, f"{tag}_ncu_sweeps_2048_summary.json", batch, False, tag)


def launch_shares(tag, out, prof):
    shutil.copy(out / "launches.csv", prof / f"{tag}_launches.csv")
    rows = list(csv.reader(l for l in open(prof / f"{tag}_launches.csv") if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1)
    tot = sum(a[1] for a in agg.values())
    with open(prof / f"{tag}_launch_shares.txt", "w") as f:
        f.write("kernel launches of a short `python bench.py` run (ncu --metrics gpu__time_duration.sum --clock-control none; "
                "cold-cache serialised times: compare shares)\n\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[:52]:52s} n={a[0]:4d} avg={a[1] / a[0]:8.2f}us share={100 * a[1] / tot:5.1f}%\n")


def summarize_rep(rep, prof, raw_name, summary_name, batch, write_traffic, tag):
    text = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    if raw_name:
        (prof / raw_name).write_text(text)
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    val = lambda r, n: float(r[hdr.index(n)].replace(",", "")) * SCALE.get(units[hdr.index(n)], 1)
    summary, traffic = [], {}
    stall_keys = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]

    def opt(r, n):
        return val(r, n) if n in hdr and r[hdr.index(n)] not in ("", "n/a") else None
    for r in rows[2:]:
        key = [v for k, v in NAMES.items() if k in r[hdr.index("Kernel Name")]][0]
        if "_tma" in r[hdr.index("Kernel Name")]:
            key += " (TMA pipeline)"
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        summary.append(dict(kernel=key, grid=r[hdr.index("Grid Size")], block=r[hdr.index("Block Size")],
                            dur_us=val(r, "gpu__time_duration.sum"), dram_read_MB=round(rd / 1e6, 2), dram_write_MB=round(wr / 1e6, 2),
                            regs=int(val(r, "launch__registers_per_thread")),
                            warps_active_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                            issue_active_pct=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            eligible_warps_per_cycle=val(r, "smsp__warps_eligible.avg.per_cycle_active"),
                            inst_executed=val(r, "smsp__inst_executed.sum"),
                            fma_pipe_pct=val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                            dram_throughput_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
        st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): val(r, k) for k in stall_keys}
        tot_st = sum(st.values()) or 1.0
        summary[-1]["stall_pct"] = {k: round(100 * v / tot_st, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6]}
        summary[-1]["smem_bank_conflicts"] = opt(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        summary[-1]["smem_wavefronts"] = opt(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        req, sec = opt(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), opt(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        summary[-1]["global_load_sectors_per_request"] = round(sec / req, 2) if req and sec else None
        summary[-1]["l2_hit_pct"] = opt(r, "lts__t_sector_hit_rate.pct")
        traffic.setdefault(key.split(" ")[0], int(rd + wr))
    json.dump(summary, open(prof / summary_name, "w"), indent=1)
    if not write_traffic:
        print("wrote", summary_name)
        return
    json.dump({"source": f"profiles/{tag}_ncu_full_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                         "S1..S4 launches cover a slice pair)", "batch": batch, "per_launch_bytes": traffic},
              open(prof / "dram_traffic.json", "w"), indent=1)
    print("wrote", tag, traffic)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2_v3", sys.argv[2] if len(sys.argv) > 2 else "",
         int(sys.argv[3]) if len(sys.argv) > 3 else 10)
