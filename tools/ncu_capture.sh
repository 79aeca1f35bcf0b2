#!/bin/bash
# One `ncu --set full` capture per kernel of interest (one launch each, after the warm-up launches), exported on the GPU
# box as raw-page CSVs (the .ncu-rep files are too large to bring back; pass KEEP_REP=1 to keep them).
#   tools/ncu_capture.sh <tag> <kernel base-name regex> <launches to skip> <command ...>
# e.g. tools/ncu_capture.sh suite_tqp_f64 ew_kernel 3 python bench.py --workload suite_tqp_o1280x137_f64 --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity
# (ncu matches the regex against the function's base name: every streaming kernel is "ew_kernel", so pick the launch by
# running a command that launches only the kernel of interest and skipping its warm-up launches)
set -u
tag=$1; re=$2; skip=$3; shift 3
mkdir -p gpurun_out
rep=gpurun_out/r02_ncu_full_${tag}
ncu --set full --clock-control none --import-source on -k "regex:${re}" --launch-skip ${skip} -c 1 -f -o ${rep} "$@" > ${rep}.log 2>&1
ncu -i ${rep}.ncu-rep --page raw --csv > ${rep}_raw.csv 2>> ${rep}.log
python - <<PY
import csv
rows = list(csv.reader(open("${rep}_raw.csv")))
if len(rows) >= 3:
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    try:
        print("== ${tag}")
        for k in keys:
            if k in d:
                print(f"  {k:90s} {d[k]:>20s} {u[k]}")
    except BrokenPipeError:  # `| head`
        pass
else:
    print("== ${tag}: no kernel captured; see ${rep}.log")
PY
if [ "${KEEP_REP:-0}" != "1" ]; then rm -f ${rep}.ncu-rep; fi
