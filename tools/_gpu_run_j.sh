mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "stream_update or launch_variants or two_overlapping or replicas" > gpurun_out/r02ae_pytest.log 2>&1; tail -3 gpurun_out/r02ae_pytest.log
B="timeout 300 python bench.py --workload c4d --steps 4 --warmup 2 --no-cpu-baseline --configs none --c5-strong off --dtype f32"
run() { name=$1; shift; $B "$@" > gpurun_out/r02ae_c4d_f32_$name.json 2> gpurun_out/r02ae_err_$name || tail -3 gpurun_out/r02ae_err_$name; }
run static_p2 --option stream_dynamic=off
run dyn_p2
run dyn_p2_d --option stage_dephase=on
run dyn_p2_d_g168 --option stage_dephase=on --option stream_grid_ctas=168
run dyn_p2_d_g200 --option stage_dephase=on --option stream_grid_ctas=200
run dyn_p2_d_g148 --option stage_dephase=on --option stream_grid_ctas=148
run dyn_p2_g168 --option stream_grid_ctas=168
run dyn_p3_d_g200 --option stage_dephase=on --option stream_grid_ctas=200 --stage-parts 3
run dyn_p4_d_g200 --option stage_dephase=on --option stream_grid_ctas=200 --stage-parts 4
run dyn_p4_d_g232 --option stage_dephase=on --option stream_grid_ctas=232 --stage-parts 4
run dyn_p4_d --option stage_dephase=on --stage-parts 4
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ae_c4d_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('f32_')[1], round(d['value']), round(d['ms_per_step'],2), round(d['roofline_hbm'].get('frac',0),3))
    except Exception as e: print(f,'ERR',e)
P
