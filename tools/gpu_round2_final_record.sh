#!/bin/bash
# Final record of round 2 in one gpurun call on one B200: GPU suite, memcheck of the dense-operator GEMM variants, smoke(), reference
# arm and the default bench line (everything lands in gpurun_out/r02final_*).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02final_pytest.log 2>&1; tail -2 gpurun_out/r02final_pytest.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_large_gpu.py -x -q -m gpu -k "wide_tile or launch_variants or two_overlapping" > gpurun_out/r02final_memcheck.log 2>&1
grep -v "^=========     at\|Host Frame\|^=========         " gpurun_out/r02final_memcheck.log | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1; tail -1 gpurun_out/r02final_smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02final_bench_reference_arm.json 2> gpurun_out/r02final_e_ref
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02final_bench_default_n1.json 2> gpurun_out/r02final_e_n1
python - <<'P'
import json
d=json.load(open('gpurun_out/r02final_bench_default_n1.json'))
print('C3', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], d['roofline'].get('whole_step'), d['clocks'])
for k,v in d.get('configs',{}).items():
    print(k, round(v.get('value',0),1), v.get('unit'), round(v.get('ms_per_step',0),2), 'fma', (v.get('roofline') or {}).get('whole_step'), 'hbm', (v.get('roofline_hbm') or {}).get('frac'))
c=d.get('c5_strong'); print('c5', c and (round(c['value']), c['ms_per_step']))
r=json.load(open('gpurun_out/r02final_bench_reference_arm.json')); print('ref', r.get('value'), r.get('unit'))
P
