mkdir -p gpurun_out
timeout 120 build/probe_gemm_parts > gpurun_out/r02ad_probe.txt 2>&1; head -3 gpurun_out/r02ad_probe.txt; tail -3 gpurun_out/r02ad_probe.txt
timeout 600 python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "launch_variants or wide_tile or full_size" > gpurun_out/r02ad_pytest.log 2>&1; tail -5 gpurun_out/r02ad_pytest.log
B="timeout 300 python bench.py --workload c4d --steps 6 --warmup 2 --no-cpu-baseline --configs none --c5-strong off"
$B --dtype f32 > gpurun_out/r02ad_c4d_f32.json 2> gpurun_out/r02ad_e1
python - <<'P'
import json
d=json.load(open('gpurun_out/r02ad_c4d_f32.json')); print(round(d['value']), round(d['ms_per_step'],2))
P
