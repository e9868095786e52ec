"""Multi-process (gloo, world_size 2) test of the sample-sharding logic used at N > 1 GPUs: ranks own
contiguous row ranges, there is no data-path collective, and the union of the shards equals the
single-process result.  Runs the CPU oracle in place of the GPU kernels (test infrastructure only)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tt-irt_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from tt_irt_py import synth
    M, d = 1000, 5
    ns, xs, rk, c = synth.make_tt(d, 9, 4, seed=3)
    q = synth.make_q(M, d, seed=4)                     # every rank regenerates the same seeds
    m0, m1 = M * rank // world, M * (rank + 1) // world  # the contiguous split of ttirt_run_host
    Z, l = oracle.oracle_run(ns, xs, rk, c, q[m0:m1])
    # the only communication of the bench: barrier + max-over-ranks of the timing scalar
    t = torch.tensor([float(rank + 1)])
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    cnt = torch.tensor([float(m1 - m0)])
    dist.all_reduce(cnt)
    assert cnt.item() == float(M)
    np.savez(os.path.join(out_dir, "shard%d.npz" % rank), Z=Z, l=l, m0=m0, m1=m1)
    dist.destroy_process_group()


def test_row_shards_reassemble_to_the_full_result(tmp_path, oracle_mod):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from tt_irt_py import synth
    ns, xs, rk, c = synth.make_tt(5, 9, 4, seed=3)
    q = synth.make_q(1000, 5, seed=4)
    Z, l = oracle_mod.oracle_run(ns, xs, rk, c, q)
    for r in range(world):
        s = np.load(os.path.join(str(tmp_path), "shard%d.npz" % r))
        assert np.array_equal(s["Z"], Z[int(s["m0"]):int(s["m1"])])
        assert np.array_equal(s["l"], l[int(s["m0"]):int(s["m1"])])
