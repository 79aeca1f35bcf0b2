#!/bin/bash
# Round-end measurement pass on one B200 box: tests, default bench, the other workloads, per-kernel timings, ncu evidence.
# Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_reference_1gpu_box.json 2>&1
: > gpurun_out/bench_workloads.jsonl
for w in theta_rh_era5_f64 suite_tqp_o1280x137_f32 suite_ttdp_o1280x137_f64 ept_wbpt_o1280x137_f64 ept_wbpt_o1280x137_f32 conv_ens_o640_shard_f64 suite_tq_hybrid_o1280x137_f64; do
  python bench.py --workload $w --no-cpu 2> gpurun_out/bench_$w.err | tail -1 >> gpurun_out/bench_workloads.jsonl
done
python tools/kbench.py --realistic --dtype f64 > gpurun_out/kbench_final_f64.log 2>&1
python tools/kbench.py --realistic --dtype f32 > gpurun_out/kbench_final_f32.log 2>&1
python tools/kbench.py --smooth --dtype f64 > gpurun_out/kbench_smooth_f64.log 2>&1
python tools/kbench_hybrid.py > gpurun_out/kbench_hybrid.log 2>&1
python tools/kbench_hybrid.py --f32 >> gpurun_out/kbench_hybrid.log 2>&1
python tools/hostbench.py > gpurun_out/hostbench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ew_kernel|hybrid|column|bisect" -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 8 --warmup 3 > gpurun_out/ncu_default.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 139 -c 1 -o gpurun_out/suite_full -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_suite.log 2>&1; echo "ncu suite rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 139 -c 1 -o gpurun_out/ept_full -f python bench.py --workload ept_wbpt_o1280x137_f64 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ept.log 2>&1; echo "ncu ept rc=$?"
