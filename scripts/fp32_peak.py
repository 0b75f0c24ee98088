"""Dev tool: FP32 FMA-pipe peak of cuda:0 through plen_measure_fp32_peak (scalar FFMA and packed FFMA2)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plen_ml_walk_b200 import _abi

lib = _abi.load_library()
out = {}
for mode, name in ((0, "ffma"), (1, "ffma2")):
    tf, mhz = C.c_float(), C.c_float()
    rc = lib.plen_measure_fp32_peak(0, mode, C.byref(tf), C.byref(mhz))
    assert rc == 0, lib.plen_last_error(None)
    out[name + "_tflops"] = tf.value
    out[name + "_implied_mhz_at_128_lanes"] = mhz.value
print(json.dumps(out))
