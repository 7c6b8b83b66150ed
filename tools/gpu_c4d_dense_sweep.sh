#!/bin/bash
# Dense-operator path (C4 dense) on one B200: bench lines, stage_parts / update_u sweeps, ncu capture of the update kernel and
# a launch list.  Run through gpurun from the repository root; everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python bench.py --workload c4d --dtype f32 > gpurun_out/s7a_c4d_f32.json 2> gpurun_out/s7a_e1
for parts in 1 2 8; do
  python bench.py --workload c4d --dtype f32 --no-cpu-baseline --stage-parts $parts > gpurun_out/s7a_c4d_f32_parts$parts.json 2>> gpurun_out/s7a_e2
done
for u in 2 8; do
  python bench.py --workload c4d --dtype f32 --no-cpu-baseline --option update_u=$u > gpurun_out/s7a_c4d_f32_u$u.json 2>> gpurun_out/s7a_e2
done
python bench.py --workload c4d --dtype f64 --no-cpu-baseline > gpurun_out/s7a_c4d_f64.json 2>> gpurun_out/s7a_e2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stage_update_kernel -s 300 -c 8 -f -o gpurun_out/s7a_prof_upd python bench.py --workload c4d --dtype f32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s7a_ncu1.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 800 --csv --log-file gpurun_out/s7a_launches_c4d_f32.csv python bench.py --workload c4d --dtype f32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s7a_ncu2.log 2>&1
for f in gpurun_out/s7a_c4d_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],2), json.dumps(d['roofline_hbm'].get('update_kernels',{}))[:600])"; done
