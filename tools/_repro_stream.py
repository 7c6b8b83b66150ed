import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from conftest import run_case
from pasif_b200 import capi
from test_large_gpu import chain_case
prec = sys.argv[1]; tt = sys.argv[2]; ctas = sys.argv[3]; E = int(sys.argv[4])
case = chain_case(32, True, E=E, L=6)
case["traj_save_steps"] = None
ref = run_case(capi.engine_class(prec), case, options={"update_stream": "off"})
a = run_case(capi.engine_class(prec), case, options={"update_stream": "on", "stream_tt": tt, "stream_ctas": ctas})
for k in ("amplitudes", "gradient"):
    print(prec, tt, ctas, E, k, np.array_equal(a[k], ref[k]), float(np.max(np.abs(a[k] - ref[k]))))
