"""Summarise an .ncu-rep (ncu --set full) into the small JSON files kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_train_kernel_ncu_full.json [profiles/train_kernel_traffic.json]"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_barriers", "launch__waves_per_multiprocessor",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALLS = ["wait", "long_scoreboard", "short_scoreboard", "barrier", "lg_throttle", "math_pipe_throttle", "not_selected", "selected",
          "branch_resolving", "no_instruction", "dispatch_stall", "mio_throttle", "membar"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[head.index("Kernel Name")], "grid": r[head.index("Grid Size")], "block": r[head.index("Block Size")]}
        for k in KEYS:
            if k in head:
                d[k] = f"{r[head.index(k)]} {units[head.index(k)]}".strip()
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in head:
                d["stall_" + s + "_per_issue"] = r[head.index(k)]
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    if len(sys.argv) > 3:
        def num(x):
            return float(x.replace(",", ""))
        per = []
        for r in data:
            rd, wr = r[head.index("dram__bytes_read.sum")], r[head.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            per.append(num(rd) * scale[units[head.index("dram__bytes_read.sum")]] + num(wr) * scale[units[head.index("dram__bytes_write.sum")]])
        # only launches of the headline shape (one global step): a capture window may also hold a fused warm-up launch
        dur = [num(r[head.index("gpu__time_duration.sum")]) for r in data]
        per = [p for p, d in zip(per, dur) if d <= 2.0 * min(dur)]
        json.dump({"kernel": res[0]["kernel"], "dram_bytes_per_launch": sum(per) / len(per), "launches_averaged": len(per),
                   "source": f"ncu --set full, {rep.split('/')[-1]} -> {out}"}, open(sys.argv[3], "w"))
    print(json.dumps(res[0], indent=1))


if __name__ == "__main__":
    main()
