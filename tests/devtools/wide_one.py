"""One wide shape, device-resident, a few steps (for ncu): python tests/devtools/wide_one.py d n r log2m [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tt-irt_b200")); sys.path.insert(0, ROOT)
import torch
from tt_irt_py import synth, tt_irt
d, n, r, log2m = (int(v) for v in sys.argv[1:5])
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
M = 1 << log2m
ns, xs, rk, c = synth.make_tt(d, n, r, seed=3)
Wf = synth.flops_per_sample(ns, rk)
md = tt_irt.Model(ns, xs, rk, c, device=0)
q = torch.rand((d, M), dtype=torch.float64, device="cuda:0")
z = torch.empty_like(q); l = torch.empty((M,), dtype=torch.float64, device="cuda:0")
st = torch.cuda.current_stream()
md.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, l.data_ptr(), None, tt_irt.MODE_FAST, st.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(st)
for _ in range(steps):
    md.sample_device(M, q.data_ptr(), M, z.data_ptr(), M, l.data_ptr(), None, tt_irt.MODE_FAST, st.cuda_stream)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"shape": [d, n, r, M], "ms": ms, "samples_per_s": M / ms * 1e3, "tflops": M * Wf / ms / 1e9, "finite": bool(torch.isfinite(l).all())}))
md.close()
