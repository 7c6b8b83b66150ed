mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "wide_tile or gemm_kernels" > gpurun_out/r02w_pytest_pair.log 2>&1
tail -15 gpurun_out/r02w_pytest_pair.log
B="timeout 300 python bench.py --workload c4d --steps 3 --warmup 2 --no-cpu-baseline --configs none --c5-strong off --dtype f32"
$B --option tc_pair=off > gpurun_out/r02w_c4d_f32_pairoff.json 2> gpurun_out/r02w_e1
$B --option tc_pair=on > gpurun_out/r02w_c4d_f32_pairon.json 2> gpurun_out/r02w_e2
$B --option tc_pair=on --option tc_pair_nt=1 > gpurun_out/r02w_c4d_f32_pairon_nt1.json 2> gpurun_out/r02w_e3
tail -3 gpurun_out/r02w_e2
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02w_c4d_*.json')):
    try:
        d=json.load(open(f)); u=d.get('roofline_hbm',{}).get('update_kernels',{})
        pk=d['roofline'].get('per_kernel',{})
        print(f, round(d['value']), round(d['ms_per_step'],2), {k:(round(v['us_per_launch'],1), round(v['frac_of_hbm_peak'],3)) for k,v in u.items()}, {k:round(v['avg_launch_ms']*1e3,1) for k,v in pk.items()}, round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
P
