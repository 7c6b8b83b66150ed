mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_large_gpu.py -x -q -m gpu > gpurun_out/r02aa_pytest_large.log 2>&1
tail -3 gpurun_out/r02aa_pytest_large.log
B="timeout 300 python bench.py --workload c4d --steps 4 --warmup 2 --no-cpu-baseline --configs none --c5-strong off"
run() { name=$1; shift; $B "$@" > gpurun_out/r02aa_c4d_$name.json 2> gpurun_out/r02aa_err_$name || tail -3 gpurun_out/r02aa_err_$name; }
run f32_pdl_p2 --dtype f32
run f32_nopdl_p2 --dtype f32 --option pdl=off
run f32_pdl_p1 --dtype f32 --stage-parts 1
run f32_nopdl_p1 --dtype f32 --option pdl=off --stage-parts 1
run f32_pdl_p3 --dtype f32 --stage-parts 3
run f32_pdl_pair_p2 --dtype f32 --option tc_pair=on
run f64_pdl_p2 --dtype f64
run f64_nopdl_p2 --dtype f64 --option pdl=off
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aa_c4d_*.json')):
    try:
        d=json.load(open(f)); u=d.get('roofline_hbm',{}).get('update_kernels',{})
        pk=d['roofline'].get('per_kernel',{})
        print(f.split('c4d_')[1], round(d['value']), round(d['ms_per_step'],2), {k:round(v['us_per_launch'],1) for k,v in u.items()}, {k:round(v['avg_launch_ms']*1e3,1) for k,v in pk.items()}, round(d['roofline_hbm'].get('frac',0),3))
    except Exception as e: print(f,'ERR',e)
P
