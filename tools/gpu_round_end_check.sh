#!/bin/bash
# What the driver runs at round end, in one gpurun call on one B200: GPU test suite, smoke(), the default bench line and
# the reference arm, plus the C4 dense lines.  Everything lands in gpurun_out/ (tag = $1).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
tag=${1:-final}
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_c3_reference_arm.json 2> gpurun_out/${tag}_e_ref
timeout 400 python bench.py > gpurun_out/${tag}_bench_c3_n1.json 2> gpurun_out/${tag}_e_c3
for dt in f32 f64; do
  timeout 200 python bench.py --workload c4d --dtype $dt > gpurun_out/${tag}_bench_c4d_$dt.json 2> gpurun_out/${tag}_e_c4d_$dt
done
python - <<PY
import json
for n in ("bench_c3_reference_arm", "bench_c3_n1", "bench_c4d_f32", "bench_c4d_f64"):
    try:
        d = json.load(open("gpurun_out/${tag}_%s.json" % n))
        print(n, round(d["value"], 1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("clocks") or {}).get("reasons"))
    except Exception as ex:
        print(n, "FAILED", ex)
PY
