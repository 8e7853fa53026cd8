import sys, json
sys.path.insert(0, "/root/repo")
import torch, bench
from luxpythonenvgym_b200 import _lib
lib = _lib.load()
for spec in ["", "two_class=1,slice=6144", "two_class=1,slice=7168", "two_class=1,slice=8192", "two_class=0"]:
    for k, v in (("two_class", -1), ("slice", 0)):
        lib.lux_b200_set_option(k.encode(), v)
    for kv in spec.split(","):
        if "=" in kv:
            k, v = kv.split("="); lib.lux_b200_set_option(k.encode(), int(v))
    d = bench.dense_leg("cuda:0", 16384, 6554.9)
    print("%-28s ms_per_step %.4f kernel %.4f units %.1f tiles %.1f" % (spec or "auto", d["ms_per_step"], d["kernel_ms"]["lux_step_kernel"], d["mean_units_per_match"], d["mean_city_tiles_per_match"]))
