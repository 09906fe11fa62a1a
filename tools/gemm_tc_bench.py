#!/usr/bin/env python
"""Times gode_gemm_tc_f32 on the three products of the QC edge encoder's wide layer (QC/layers.py:76-86, hidden 73:
[E, 2667] x [2667, 5329] forward, its input gradient and its weight gradient) for several tile-order band heights
(GODE_GEMM_BAND; a huge band = the old row-tile-fastest order).  One JSON line per (shape, band)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import graph_odenet_b200  # noqa: E402,F401
from graph_odenet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
E_, K_, N_ = int(os.environ.get("QC_E", "146618")), 2667, 5329
torch.manual_seed(0)
h = torch.randn(E_, K_, device=dev)
w = torch.randn(K_, N_, device=dev) / K_ ** 0.5
g = torch.randn(E_, N_, device=dev)
hp, wt = ops._pad4(h), ops._pad4(w.t().contiguous())       # operands as LinearFn hands them over (copies outside the timing)
gp, wp = ops._pad4(g), ops._pad4(w)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flop = 2.0 * E_ * K_ * N_
bands = [int(b) for b in os.environ.get("BANDS", "1000000,32,16,8,4").split(",")]
for band in bands:
    os.environ["GODE_GEMM_BAND"] = str(band)
    for name, fn in (("fwd  [E,2667]x[2667,5329]", lambda: ops.gemm_tc(hp, wt)),
                     ("dX   [E,5329]x[5329,2667]", lambda: ops.gemm_tc(gp, wp)),
                     ("dW   [2667,E]x[E,5329]", lambda: ops.gemm_tc_reduce_rows(h, g))):
        ms = timed(fn)
        print(json.dumps({"product": name, "band": band, "ms": round(ms, 2), "tflops_fp32_equiv": round(flop / ms / 1e9, 1),
                          "tflops_tf32_3pass": round(3 * flop / ms / 1e9, 1)}), flush=True)
