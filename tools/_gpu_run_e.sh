mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "wide_tile or gemm_kernels or stream_update or full_size" > gpurun_out/r02y_pytest.log 2>&1
tail -3 gpurun_out/r02y_pytest.log
B="timeout 300 python bench.py --workload c4d --steps 4 --warmup 2 --no-cpu-baseline --configs none --c5-strong off --dtype f32"
run() { name=$1; shift; $B "$@" > gpurun_out/r02y_c4d_f32_$name.json 2> gpurun_out/r02y_err_$name || tail -3 gpurun_out/r02y_err_$name; }
run base_p2
run s3_p2 --option tc_stages=3 --option stream_smem_kb=80 --option stream_grid=1
run s3_p2_d --option tc_stages=3 --option stream_smem_kb=80 --option stream_grid=1 --option stage_dephase=on
run s3_p3_d --option tc_stages=3 --option stream_smem_kb=80 --option stream_grid=1 --option stage_dephase=on --stage-parts 3
run s3_p4_d --option tc_stages=3 --option stream_smem_kb=80 --option stream_grid=1 --option stage_dephase=on --stage-parts 4
run s3_p4 --option tc_stages=3 --option stream_smem_kb=80 --option stream_grid=1 --stage-parts 4
run s3_p8_d --option tc_stages=3 --option stream_smem_kb=80 --option stream_grid=1 --option stage_dephase=on --stage-parts 8
run s2_p4_d --option tc_stages=2 --option stream_smem_kb=128 --option stream_grid=1 --option stage_dephase=on --stage-parts 4
run s2_p2_d --option tc_stages=2 --option stream_smem_kb=128 --option stream_grid=1 --option stage_dephase=on --stage-parts 2
run base_p2_d --option stage_dephase=on
run base_p4_d --option stage_dephase=on --stage-parts 4
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02y_c4d_*.json')):
    try:
        d=json.load(open(f)); u=d.get('roofline_hbm',{}).get('update_kernels',{})
        pk=d['roofline'].get('per_kernel',{})
        print(f.split('f32_')[1], round(d['value']), round(d['ms_per_step'],2), {k:round(v['us_per_launch'],1) for k,v in u.items()}, {k:round(v['avg_launch_ms']*1e3,1) for k,v in pk.items()}, round(d['roofline_hbm'].get('frac',0),3))
    except Exception as e: print(f,'ERR',e)
P
