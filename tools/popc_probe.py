"""Measure the integer-pipe denominators for the kNN roofline on the GPU box (writes JSON)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orbslam_in_practice_b200 import _lib
popc, mixed = _lib.popc_peak(0)
out = {"popc_per_s": popc, "popc_with_xor_per_s": mixed,
       "gpairs_per_s_at_8_popc": popc / 8 / 1e9, "gpairs_per_s_at_8_popc_xor": mixed / 8 / 1e9}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/popc_peak.json", "w"))
