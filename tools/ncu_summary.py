"""Summarise an .ncu-rep: key metrics per kernel and the hottest source lines / stall reasons.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_xu.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==== %s" % r[idx["Kernel Name"]][:110])
        for w in WANT:
            if w in idx:
                print("  %-82s %s %s" % (w, r[idx[w]], units[idx[w]]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"],
                             capture_output=True, text=True).stdout
        blocks = src.split("Kernel Name")
        rows = list(csv.reader(io.StringIO(src)))
        # find header row
        hi = [i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r)]
        for h in hi:
            hdr = rows[h]
            si = hdr.index("Source")
            ci = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
            ii = [i for i, c in enumerate(hdr) if c == "Instructions Executed"]
            body = []
            for r in rows[h + 1:]:
                if len(r) != len(hdr):
                    break
                try:
                    body.append((int(r[ci[0]]), r[si], r[ii[0]] if ii else ""))
                except Exception:
                    pass
            tot = sum(b[0] for b in body) or 1
            print("---- hottest source lines (of %d samples)" % tot)
            for s, line, ie in sorted(body, key=lambda b: -b[0])[:n]:
                print("  %5.1f%%  inst=%-10s %s" % (100.0 * s / tot, ie, line.strip()[:120]))


if __name__ == "__main__":
    main()
