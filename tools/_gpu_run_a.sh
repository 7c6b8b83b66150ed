set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "stream_update or two_overlapping or full_size or replicas or offsite" > gpurun_out/r02v_pytest_stream.log 2>&1
tail -5 gpurun_out/r02v_pytest_stream.log
B="python bench.py --workload c4d --steps 3 --warmup 2 --no-cpu-baseline --configs none --c5-strong off"
for dt in f32 f64; do
  timeout 300 $B --dtype $dt > gpurun_out/r02v_c4d_${dt}_c2p4.json 2> gpurun_out/r02v_e1
  timeout 300 $B --dtype $dt --stage-parts 2 > gpurun_out/r02v_c4d_${dt}_c2p2.json 2> gpurun_out/r02v_e2
  timeout 300 $B --dtype $dt --option stream_ctas=1 > gpurun_out/r02v_c4d_${dt}_c1p4.json 2> gpurun_out/r02v_e3
  timeout 300 $B --dtype $dt --option stream_ctas=1 --stage-parts 2 > gpurun_out/r02v_c4d_${dt}_c1p2.json 2> gpurun_out/r02v_e4
  timeout 300 $B --dtype $dt --option stream_ctas=1 --stage-parts 1 > gpurun_out/r02v_c4d_${dt}_c1p1.json 2> gpurun_out/r02v_e5
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02v_c4d_*.json')):
    try:
        d=json.load(open(f)); u=d.get('roofline_hbm',{}).get('update_kernels',{})
        print(f, round(d['value']), round(d['ms_per_step'],2), {k:(round(v['us_per_launch'],1), round(v['frac_of_hbm_peak'],3)) for k,v in u.items()}, round(d['roofline'].get('gemm_us_per_launch',0),1))
    except Exception as e: print(f,'ERR',e)
P
