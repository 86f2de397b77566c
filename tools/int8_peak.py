"""Measured int8 tensor peak of this GPU (tcgen05.mma kind::i8 on every SM for >= 2 s) with the clocks and
power seen meanwhile.  Writes one JSON line; bench.py measures the same figure live for roofline.peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plink_ng_b200 as p
from bench import ClockSampler

out = {}
with p.GpuContext(0) as ctx:
    for name, n, form, secs in (("ts_n128", 128, 1, 2.5), ("ts_n64", 64, 1, 1.0), ("ts_n192", 192, 1, 1.0), ("ss_n240", 240, 0, 2.5)):
        s = ClockSampler(0)
        s.start()
        tops, t = ctx.int8_peak(n, form, secs)
        out[name] = {"tops": tops, "seconds": t, "clocks": s.stop()}
out["note"] = "all 148 SMs, one issuer warp per SM, 64 back-to-back UMMAs (M=128, K=32) per commit; ts = A operand in tensor memory"
print(json.dumps(out))
