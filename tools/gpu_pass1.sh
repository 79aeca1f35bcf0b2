#!/bin/bash
# first GPU pass of round 2: tests, parity report, benches (outputs under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_box.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
python tools/parity_report.py gpu --out gpurun_out/parity_gpu.json > gpurun_out/r02_parity_gpu.log 2>&1; tail -2 gpurun_out/r02_parity_gpu.log
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 600 gpurun_out/r02_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 1500 gpurun_out/r02_bench_default.json; tail -3 gpurun_out/r02_bench_default.err
for w in suite7_tqp_o1280x137_f64 suite7_tqp_o1280x137_f32 single_pass_tqp_o1280x137_f64 ept_wbpt_o1280x137_f64 suite_tq_hybrid_o1280x137_f64; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_$w.json"))
    print("$w", "frac=%.3f"%d["roofline"]["frac"], "Gpt/s=%.1f"%(d["value"]/1e9), "parity", d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel"], d["parity"]["n_over_limit"], d["parity"]["nan_mismatches"]), "e2e", d["e2e"] and d["e2e"]["value"]/1e9, "pageable", d["e2e_pageable"] and d["e2e_pageable"]["value"]/1e9, d["clocks"])
except Exception as e:
    print("$w FAILED", e)
PY
  tail -2 gpurun_out/r02_bench_$w.err
done
python tools/kbench.py --dtype f64 > gpurun_out/r02_kbench_f64.log 2>&1; cat gpurun_out/r02_kbench_f64.log
python tools/kbench.py --dtype f32 > gpurun_out/r02_kbench_f32.log 2>&1; cat gpurun_out/r02_kbench_f32.log
