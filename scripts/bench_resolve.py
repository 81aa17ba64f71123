#!/usr/bin/env python3
"""HBM roofline of the resolve kernel (PostProcessImage + AntiAliasImage, sr_resolve.cu): a supersampled
surface in HBM -> the final frame in HBM.  Algorithmic bytes = 4 B per source pixel read + 4 B per
destination pixel written.  Inputs are larger than L2 (126 MB) for the 8K cases; L2 is flushed between
iterations anyway.  usage: python scripts/bench_resolve.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from softray_b200 import lib

ctx = lib.Context(0)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
hbm = peaks.get("hbm_gbs", 6650.0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for (w, h, aa, style) in [(3840, 2160, 2, 0), (7680, 4320, 2, 0), (3840, 2160, 4, 0), (7680, 4320, 1, 2), (1920, 1080, 8, 1)]:
    src = torch.randint(0, 2**31 - 1, (h * aa, w * aa), dtype=torch.int32, device="cuda")
    dst = torch.empty((h, w), dtype=torch.int32, device="cuda")
    times = []
    for it in range(8):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        ctx.resolve_device(src.data_ptr(), dst.data_ptr(), w, h, aa, style, 0xFF00FF, stream=stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            times.append(a.elapsed_time(b))
    ms = sum(times) / len(times)
    nbytes = 4.0 * w * h * (aa * aa + 1)
    out.append({"dst": [w, h], "aa": aa, "style": style, "ms": ms, "bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6,
                "peak_gbs": hbm, "frac": nbytes / ms / 1e6 / hbm})
print(json.dumps({"kernel": "sr::resolve_kernel", "bound": "hbm", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                  "cases": out}))
