"""Aggregate host <-> device copy rate of the box with all GPUs copying at once (pinned memory, both directions):
the ceiling of any end-to-end figure that streams q in and Z out."""
import json, time, torch
G = torch.cuda.device_count()
n = 1 << 27  # 1 GiB of doubles per buffer
bufs = []
for g in range(G):
    with torch.cuda.device(g):
        bufs.append((torch.empty(n, dtype=torch.float64, pin_memory=True), torch.empty(n, dtype=torch.float64, pin_memory=True),
                     torch.empty(n, dtype=torch.float64, device="cuda:%d" % g), torch.empty(n, dtype=torch.float64, device="cuda:%d" % g),
                     torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)))
def run(gpus, reps=3):
    for g in gpus: torch.cuda.synchronize(g)
    t = time.perf_counter()
    for _ in range(reps):
        for g in gpus:
            h1, h2, d1, d2, s1, s2 = bufs[g]
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    for g in gpus: torch.cuda.synchronize(g)
    dt = time.perf_counter() - t
    return 2 * reps * len(gpus) * n * 8 / dt / 1e9
run(range(G), 1)
print(json.dumps({"gpus": G, "duplex_GBps_total": {str(k): round(run(range(k)), 1) for k in (1, 2, 4, 8) if k <= G}}))
