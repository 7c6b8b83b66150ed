mkdir -p gpurun_out
B="timeout 300 python bench.py --workload c4d --steps 4 --warmup 2 --no-cpu-baseline --configs none --c5-strong off --dtype f32"
run() { name=$1; shift; $B "$@" > gpurun_out/r02af_c4d_f32_$name.json 2> gpurun_out/r02af_err_$name || tail -3 gpurun_out/r02af_err_$name; }
run p1 --stage-parts 1
run p2
run p4 --stage-parts 4
run p4_d --stage-parts 4 --option stage_dephase=on
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02af_c4d_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('f32_')[1], round(d['value']), round(d['ms_per_step'],2), 'enqueue ms', d.get('host_enqueue_ms_per_step'), 'launches', d.get('gpu_launches'))
    except Exception as e: print(f,'ERR',e)
P
