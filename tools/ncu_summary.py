"""Markdown summary of ncu --set full reports (one section per kernel launch in the report).
    python tools/ncu_summary.py title report.ncu-rep [report2.ncu-rep ...] > profiles/<name>.md
Reads the raw page through `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import subprocess
import sys

KEYS = [("time", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs", "launch__registers_per_thread"), ("SM clock", "smsp__cycles_elapsed.avg.per_second"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
        ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor datapath active % (of elapsed)", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("warp instructions", "smsp__inst_executed.sum"),
        ("shared-memory wavefronts (LSU)", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("shared-memory store wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
        ("L2 hit rate %", "lts__t_sector_hit_rate.pct"), ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM throughput %", "dram__throughput.avg.pct_of_peak_sustained_elapsed")]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "branch_resolving", "barrier", "not_selected"]


def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(out.splitlines()))
    return r[0], r[1], r[2:]


def main():
    title, reports = sys.argv[1], sys.argv[2:]
    print("# " + title + "\n")
    for path in reports:
        h, u, data = rows(path)
        print("# report " + path.split("/")[-1] + "\n")
        for v in data:
            name = v[h.index("Kernel Name")]
            name = name.split("(")[0].replace("void ", "").replace("skr::", "")
            print("## " + name)
            for label, key in KEYS:
                if key in h:
                    i = h.index(key)
                    print("- %s: %s %s" % (label, v[i], u[i]))
            tot = 0
            st = {}
            for s in STALLS:
                key = "smsp__pcsamp_warps_issue_stalled_" + s
                if key in h:
                    st[s] = int(float(v[h.index(key)] or 0))
            for i, x in enumerate(h):
                if x.startswith("smsp__pcsamp_warps_issue_stalled_") and not x.endswith("_not_issued"):
                    tot += int(float(v[i] or 0))
            if tot:
                print("- stall samples (of %d): " % tot + ", ".join("%s %.0f%%" % (s, 100.0 * c / tot) for s, c in sorted(st.items(), key=lambda t: -t[1])))
            print()


if __name__ == "__main__":
    main()
