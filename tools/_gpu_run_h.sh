mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02ac_pytest.log 2>&1; tail -3 gpurun_out/r02ac_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; tail -1 gpurun_out/r02ac_smoke.log
B="python bench.py --workload c4d --length 8 --steps 1 --warmup 1 --no-cpu-baseline --configs none --c5-strong off --dtype f32 --stage-parts 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|stage_update_stream' -s 60 -c 8 -f -o /tmp/r02ac_dense $B > gpurun_out/r02ac_ncu1.log 2>&1
ncu -i /tmp/r02ac_dense.ncu-rep --page raw --csv > gpurun_out/r02ac_dense_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'split_kernel' -s 30 -c 2 -f -o /tmp/r02ac_split $B --option tc_pair=off > gpurun_out/r02ac_ncu2.log 2>&1
ncu -i /tmp/r02ac_split.ncu-rep --page raw --csv > gpurun_out/r02ac_split_raw.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02ac_launches_c4d_f32.csv python bench.py --workload c4d --dtype f32 --steps 1 --warmup 1 --no-cpu-baseline --configs none --c5-strong off > gpurun_out/r02ac_ncu3.log 2>&1
for dt in f32 f64; do
  timeout 300 python bench.py --workload c4d --dtype $dt --no-cpu-baseline --configs none --c5-strong off > gpurun_out/r02ac_bench_c4d_$dt.json 2> gpurun_out/r02ac_e_$dt
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ac_bench_c4d_*.json')):
    try:
        d=json.load(open(f)); u=d.get('roofline_hbm',{}).get('update_kernels',{})
        pk=d['roofline'].get('per_kernel',{})
        print(f, round(d['value']), round(d['ms_per_step'],2), {k:round(v['us_per_launch'],1) for k,v in u.items()}, {k:round(v['avg_launch_ms']*1e3,1) for k,v in pk.items()}, round(d['roofline_hbm'].get('frac',0),3))
    except Exception as e: print(f,'ERR',e)
P
ls -la gpurun_out | tail -12
