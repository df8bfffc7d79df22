"""Summarise ncu captures into profiles/ (run in the build container, no GPU needed).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches.md
    python tools/summarize_ncu.py kernels  gpurun_out/prof.ncu-rep profiles/r1_kernels.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        d.setdefault(r[ki].split("(")[0], []).append(v)
    tot = sum(sum(v) / len(v) for v in d.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("source: `%s`\n\n| kernel | launches captured | avg us | share of one step |\n|---|---|---|---|\n" % src)
        for k, v in d.items():
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k, len(v), sum(v) / len(v), sum(v) / len(v) / tot * 100))
        f.write("| **sum of per-kernel averages (one step)** | | %.1f | 100 %% |\n" % tot)
    print(open(dst).read())


def kernels(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none (one launch per kernel)\n\nsource: `%s`\n\n" % src)
        for r in rows[2:]:
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]])
            for m in METRICS:
                if m in idx:
                    f.write("| %s | %s | %s |\n" % (m, r[idx[m]], units[idx[m]]))
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2], sys.argv[3])
