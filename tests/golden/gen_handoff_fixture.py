#!/usr/bin/env python3
"""Record what the UNMODIFIED reference front end (/root/reference/PASIf.py) hands to its native driver for a few
systems -> tests/golden/handoff_reference_frontend.json.  Runs in the build container (needs /root/reference)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import pasif_oracle as po  # noqa: E402
from pasif_b200 import workloads as wl  # noqa: E402
import test_frontend as tf  # noqa: E402
from conftest import load_reference_frontend  # noqa: E402

ref = load_reference_frontend(po.RecordingDriver)
out = {name: tf._jsonable(tf._handoff(ref.PASIf, ref, wl.golden_case(name))) for name in tf.CASES}
json.dump(out, open(tf.FIXTURE, "w"))
print("wrote", tf.FIXTURE)
