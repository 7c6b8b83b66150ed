mkdir -p gpurun_out
timeout 120 build/probe_gemm_parts > gpurun_out/r02z_probe_gemm_parts_v3.txt 2>&1
cat gpurun_out/r02z_probe_gemm_parts_v3.txt
timeout 600 python -m pytest tests/test_large_gpu.py -x -q -m gpu > gpurun_out/r02z_pytest_large.log 2>&1
tail -3 gpurun_out/r02z_pytest_large.log
B="timeout 300 python bench.py --workload c4d --steps 4 --warmup 2 --no-cpu-baseline --configs none --c5-strong off --dtype f32"
run() { name=$1; shift; $B "$@" > gpurun_out/r02z_c4d_f32_$name.json 2> gpurun_out/r02z_err_$name || tail -3 gpurun_out/r02z_err_$name; }
run base_p2
run base_p3 --stage-parts 3
run pair_p2 --option tc_pair=on
run pair_nt1_p2 --option tc_pair=on --option tc_pair_nt=1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02z_c4d_*.json')):
    try:
        d=json.load(open(f)); u=d.get('roofline_hbm',{}).get('update_kernels',{})
        pk=d['roofline'].get('per_kernel',{})
        print(f.split('f32_')[1], round(d['value']), round(d['ms_per_step'],2), {k:round(v['us_per_launch'],1) for k,v in u.items()}, {k:round(v['avg_launch_ms']*1e3,1) for k,v in pk.items()}, round(d['roofline_hbm'].get('frac',0),3))
    except Exception as e: print(f,'ERR',e)
P
