"""Summarise an .ncu-rep: headline raw metrics per kernel and the hottest CUDA source lines.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] [--lines N]

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source
sass,cuda` (needs -lineinfo builds and --import-source on at capture time).
"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_no_instructions",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_sleeping",
    "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_drain", "smsp__pcsamp_warps_issue_stalled_tex_throttle",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def raw_page(rep, regex):
    args = ["-i", rep, "--page", "raw", "--csv"]
    if regex:
        args += ["--kernel-name", f"regex:{regex}"]
    rows = list(csv.reader(io.StringIO(ncu(args))))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"\n## {r[ki][:70]}  (id {r[0]})\n")
        print("| metric | unit | value |\n|---|---|---|")
        for m in RAW:
            if m in hdr:
                print(f"| {m} | {units[hdr.index(m)]} | {r[hdr.index(m)]} |")


def source_page(rep, regex, n_lines):
    args = ["-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
    if regex:
        args += ["--kernel-name", f"regex:{regex}"]
    rows = list(csv.reader(io.StringIO(ncu(args))))
    cur, out, kernel = None, [], None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            if kernel is not None and r[1] != kernel and out:
                emit(kernel, out, n_lines)
                out = []
            kernel = r[1]
        elif r[0] not in ("", "Line No") and cur and len(r) > 8:
            try:
                out.append((int(r[6] or 0), int(r[7] or 0), cur, r[0], r[1].strip()[:100]))
            except ValueError:
                pass
    if out:
        emit(kernel, out, n_lines)


def emit(kernel, out, n_lines):
    tot = sum(o[0] for o in out) or 1
    toti = sum(o[1] for o in out) or 1
    print(f"\n### hottest source lines of {kernel[:60]} ({tot} warp samples, {toti} warp instructions)\n")
    print("| samples % | instr % | line | source |\n|---|---|---|---|")
    for o in sorted(out, key=lambda o: -o[0])[:n_lines]:
        print(f"| {100 * o[0] / tot:.1f} | {100 * o[1] / toti:.1f} | {o[2]}:{o[3]} | `{o[4]}` |")


def main():
    rep = sys.argv[1]
    regex = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    n_lines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    raw_page(rep, regex)
    source_page(rep, regex, n_lines)


if __name__ == "__main__":
    main()
