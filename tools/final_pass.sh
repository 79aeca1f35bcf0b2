#!/bin/bash
# Round-2 measurement pass on one B200 box: tests, default bench + reference arm, every workload, per-kernel timings, host paths,
# ncu evidence.  Everything lands in gpurun_out/r02f_*; the summaries are copied to profiles/ by hand.
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -q -rs ) > $O/r02f_pytest_gpu.log 2>&1; tail -3 $O/r02f_pytest_gpu.log
python bench.py --impl reference > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err
python bench.py > $O/r02f_bench_default.json 2> $O/r02f_bench_default.err; echo "bench rc=$?"
: > $O/r02f_bench_workloads.jsonl
for w in suite_tqp_o1280x137_f64 suite_tqp_o1280x137_f32 suite_ttdp_o1280x137_f64 single_pass_tqp_o1280x137_f64 suite7_tqp_o1280x137_f64 suite7_tqp_o1280x137_f32 suite7_ttdp_o1280x137_f64 ept_wbpt_o1280x137_f64 ept_wbpt_o1280x137_f32 suite_tq_hybrid_o1280x137_f64 conv_ens_o640_shard_f64; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2> $O/r02f_bench_$w.err | tail -1 >> $O/r02f_bench_workloads.jsonl
done
# one ERA5 level is a 13 us step: enough steps for a steady state
python bench.py --workload theta_rh_era5_f64 --steps 2000 --warmup 200 --no-cpu 2> $O/r02f_bench_theta_rh_era5_f64.err | tail -1 >> $O/r02f_bench_workloads.jsonl
python - <<PY
import json
for ln in open("$O/r02f_bench_workloads.jsonl"):
    d=json.loads(ln); p=d['parity']
    print(f"{d['config']['workload']:34s} frac={d['roofline']['frac']:.3f} {d['value']/1e9:7.1f} Gpt/s parity={p['ok']} max={p['max_rel']:.1e} e2e={d['e2e']['value']/1e9:.2f} pg={d['e2e_pageable'] and round(d['e2e_pageable']['value']/1e9,2)} {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
python tools/kbench.py --realistic --dtype f64 > $O/r02f_kbench_f64.log 2>&1
python tools/kbench.py --realistic --dtype f32 > $O/r02f_kbench_f32.log 2>&1
python tools/kbench.py --smooth --dtype f64 > $O/r02f_kbench_smooth_f64.log 2>&1
python tools/kbench_hybrid.py > $O/r02f_kbench_hybrid.log 2>&1
python tools/kbench_hybrid.py --f32 >> $O/r02f_kbench_hybrid.log 2>&1
python tools/kbench_wind.py > $O/r02f_kbench_wind.log 2>&1
python tools/kbench_wind.py --f32 >> $O/r02f_kbench_wind.log 2>&1
python tools/kbench_scalar_p.py > $O/r02f_kbench_scalar_p.log 2>&1
python tools/kbench_levels.py > $O/r02f_kbench_levels.log 2>&1
python tools/kbench_levels.py --f32 >> $O/r02f_kbench_levels.log 2>&1
python tools/hostbench.py --points 105594880 > $O/r02f_hostbench.log 2>&1
python tools/dma_ceiling.py > $O/r02f_dma_ceiling_1gpu.log 2>&1
python tools/parity_report.py gpu --out $O/parity_gpu.json > $O/r02f_parity_gpu.log 2>&1
python tools/sweep.py --dtype f64 > $O/r02f_sweep_f64.jsonl 2> $O/r02f_sweep_f64.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ew_kernel|hybrid|column|bisect" -c 400 --csv --log-file $O/r02f_launches_default_bench_command.csv python bench.py --steps 8 --warmup 3 --no-e2e-pageable > $O/r02f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
B="--steps 2 --warmup 3 --no-cpu --no-e2e --no-parity"
tools/ncu_capture.sh suite_tqp_f64 ew_kernel 3 python bench.py --workload suite_tqp_o1280x137_f64 $B | head -12
tools/ncu_capture.sh suite_tqp_f32 ew_kernel 3 python bench.py --workload suite_tqp_o1280x137_f32 $B | head -12
tools/ncu_capture.sh suite_ttdp_f64 ew_kernel 3 python bench.py --workload suite_ttdp_o1280x137_f64 $B | head -12
tools/ncu_capture.sh suite7_tqp_f64 ew_kernel 3 python bench.py --workload suite7_tqp_o1280x137_f64 $B | head -12
tools/ncu_capture.sh single_pass_tqp_f64 ew_kernel 3 python bench.py --workload single_pass_tqp_o1280x137_f64 $B | head -12
tools/ncu_capture.sh ept_wbpt_f64 ew_kernel 3 python bench.py --workload ept_wbpt_o1280x137_f64 $B | head -12
tools/ncu_capture.sh suite_tq_hybrid_f64 suite_hybrid_kernel 3 python bench.py --workload suite_tq_hybrid_o1280x137_f64 $B | head -12
tools/ncu_capture.sh wbpt_newton_f64 ew_kernel 3 python tools/kbench.py --realistic --only wbpt_newton --iters 2 | head -12
tools/ncu_capture.sh wbpt_bisect_f64 ew_kernel 3 python tools/kbench.py --realistic --only wbpt_bisect --iters 2 | head -12
tools/ncu_capture.sh es_mixed_f64 ew_kernel 3 python tools/kbench.py --realistic --only es_mixed --iters 2 | head -12
tools/ncu_capture.sh thickness_f64 column_geopotential_kernel 2 python tools/kbench_hybrid.py | head -12
tools/ncu_capture.sh geometric_height_f64 column_geopotential_kernel 9 python tools/kbench_hybrid.py | head -12
tools/ncu_capture.sh wind_speed_f64 ew_kernel 3 python tools/kbench_wind.py | head -12
tools/ncu_capture.sh wind_direction_f64 ew_kernel 16 python tools/kbench_wind.py | head -12
tools/ncu_capture.sh ept_wbpt_f32 ew_kernel 3 python bench.py --workload ept_wbpt_o1280x137_f32 $B | head -12
tools/ncu_capture.sh suite7_ttdp_f64 ew_kernel 3 python bench.py --workload suite7_ttdp_o1280x137_f64 $B | head -12
tail -3 $O/r02f_kbench_f64.log
