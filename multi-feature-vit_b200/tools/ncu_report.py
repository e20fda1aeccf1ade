"""Compact text summary of an `ncu --set full` report (one kernel launch), for profiles/.

    python multi-feature-vit_b200/tools/ncu_report.py gpurun_out/r02_gemm_fc1.ncu-rep [--source N] > profiles/r02_ncu_fc1.md

Reads the report with `ncu -i <rep> --page raw --csv` (and `--page source --csv` for the N hottest source lines by stall
samples).  Prints: duration, tensor-pipe / issue / XU / FMA / ALU / LSU utilisation (elapsed AND active), DRAM and L2
bytes, achieved occupancy, and the warp-stall sample histogram.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("duration_us", "gpu__time_duration.sum", 1),
    ("sm_active_cycles", "sm__cycles_active.avg", 1),
    ("tensor_pipe_pct_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("tensor_pipe_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("issue_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("xu_pipe_pct_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", 1),
    ("xu_pipe_pct_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    ("fma_pipe_pct_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("fmaheavy_pipe_pct_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("alu_pipe_pct_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("lsu_pipe_pct_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    ("dram_read_MB", "dram__bytes_read.sum", 1e-6),
    ("dram_write_MB", "dram__bytes_write.sum", 1e-6),
    ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("l2_to_l1_read_MB", "lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum", 1e-6),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("registers_per_thread", "launch__registers_per_thread", 1),
    ("dyn_smem_per_block_KB", "launch__shared_mem_per_block_dynamic", 1e-3),
    ("grid_size", "launch__grid_size", 1),
    ("inst_executed", "sm__inst_executed.sum", 1),
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    n_src = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else 0
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## %s" % d.get("Kernel Name", "?"))
        print("grid %s block %s" % (d.get("Grid Size"), d.get("Block Size")))
        for label, key, scale in KEYS:
            if key in d and d[key] not in ("", "n/a"):
                try:
                    print("%-28s %12.3f   (%s)" % (label, float(d[key].replace(",", "")) * scale, key))
                except ValueError:
                    print("%-28s %12s   (%s)" % (label, d[key], key))
        stalls = []
        for k in hdr:
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    stalls.append((int(float(d[k])), k[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        total = sum(c for c, _ in stalls) or 1
        print("warp stall samples (total %d):" % total)
        for c, name in sorted(stalls, reverse=True)[:10]:
            print("    %-28s %6d  %5.1f %%" % (name, c, 100.0 * c / total))
        print()
    if n_src:
        src = page(rep, "source")
        while src and "Address" not in src[0]:
            src = src[1:]
        if len(src) > 2:
            h = src[0]
            col = [i for i, k in enumerate(h) if "Sampling" in k and "All" in k]
            scol = [i for i, k in enumerate(h) if k.strip() in ("Source", "#", "stall_long_sb", "stall_wait", "stall_barrier",
                                                                 "stall_mio", "stall_math", "stall_short_sb")]
            print("columns: samples | " + " | ".join(h[i] for i in scol))
            if col:
                ci = col[0]
                ranked = []
                for r in src[1:]:
                    try:
                        ranked.append((int(float(r[ci])), r))
                    except (ValueError, IndexError):
                        pass
                ranked.sort(key=lambda t: -t[0])
                print("hottest source lines by stall samples (%s):" % h[ci])
                for c, r in ranked[:n_src]:
                    print("    %6d  %s" % (c, " | ".join(r[i] for i in scol)[:150]))


if __name__ == "__main__":
    main()
