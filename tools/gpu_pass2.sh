#!/bin/bash
# second GPU pass of round 2: tests, bench of every workload, ncu captures
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -rs ) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
rm -f gpurun_out/r02_bench_workloads.jsonl
for w in suite_tqp_o1280x137_f64 suite_tqp_o1280x137_f32 suite_ttdp_o1280x137_f64 single_pass_tqp_o1280x137_f64 suite7_tqp_o1280x137_f64 suite7_tqp_o1280x137_f32 suite7_ttdp_o1280x137_f64 ept_wbpt_o1280x137_f64 ept_wbpt_o1280x137_f32 suite_tq_hybrid_o1280x137_f64 conv_ens_o640_shard_f64 theta_rh_era5_f64; do
  extra="--no-cpu"; [ $w = suite_tqp_o1280x137_f64 ] && extra=""
  python bench.py --workload $w --steps 20 --warmup 5 $extra 2> gpurun_out/r02_bench_$w.err | tee -a gpurun_out/r02_bench_workloads.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['parity']
print('$w', 'frac=%.3f'%d['roofline']['frac'], 'Gpt/s=%.1f'%(d['value']/1e9), 'parity', p['ok'], '%.1e'%p['max_rel'], p['n_over_limit'], p['nan_mismatches'], 'e2e %.2f'%(d['e2e']['value']/1e9), 'pageable', d['e2e_pageable'] and '%.2f'%(d['e2e_pageable']['value']/1e9), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  tail -2 gpurun_out/r02_bench_$w.err | grep -v "^$" | head -2
done
B="--steps 2 --warmup 3 --no-cpu --no-e2e --no-parity"
tools/ncu_capture.sh suite_tqp_f64 ew_kernel 3 python bench.py --workload suite_tqp_o1280x137_f64 $B
tools/ncu_capture.sh suite_tqp_f32 ew_kernel 3 python bench.py --workload suite_tqp_o1280x137_f32 $B
tools/ncu_capture.sh suite_ttdp_f64 ew_kernel 3 python bench.py --workload suite_ttdp_o1280x137_f64 $B
tools/ncu_capture.sh suite7_tqp_f64 ew_kernel 3 python bench.py --workload suite7_tqp_o1280x137_f64 $B
tools/ncu_capture.sh ept_wbpt_f64 ew_kernel 3 python bench.py --workload ept_wbpt_o1280x137_f64 $B
tools/ncu_capture.sh suite_tq_hybrid_f64 suite_hybrid_kernel 3 python bench.py --workload suite_tq_hybrid_o1280x137_f64 $B
tools/ncu_capture.sh wbpt_newton_f64 ew_kernel 3 python tools/kbench.py --realistic --only wbpt_newton --iters 2
tools/ncu_capture.sh wbpt_bisect_f64 ew_kernel 3 python tools/kbench.py --realistic --only wbpt_bisect --iters 2
tools/ncu_capture.sh thickness_f64 column_geopotential_kernel 2 python tools/kbench_hybrid.py
