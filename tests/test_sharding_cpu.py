"""Multi-process (gloo, world_size 2) test of the sample-sharding logic used at N > 1 GPUs: ranks own
contiguous row ranges, there is no data-path collective, and the union of the shards equals the
single-process result.  The row split and the device policy are the PRODUCT's own host code (ttirt_shard_rows,
ttirt_auto_devices, exported by the C-ABI library and used by ttirt_run_host / tt_irt1 themselves; they need no device);
the CPU oracle stands in for the GPU kernels on each shard (test infrastructure only)."""
import ctypes
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
    from tt_irt_py import tt_irt
    lib = tt_irt.load_library()
    a, b = ctypes.c_longlong(-1), ctypes.c_longlong(-1)
    lib.ttirt_shard_rows.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    assert lib.ttirt_shard_rows(M, world, rank, ctypes.byref(a), ctypes.byref(b)) == 0   # the split ttirt_run_host uses
    m0, m1 = a.value, b.value
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


def test_product_shard_arithmetic_and_device_policy():
    """ttirt_shard_rows / ttirt_auto_devices (host code of the product library, no device needed): shards are contiguous,
    ordered, cover [0, M) exactly and differ by at most one row, up to the int32 ABI's largest batch; the default
    device policy shards a batch over N GPUs once it holds 2^22 seed points per GPU."""
    from tt_irt_py import tt_irt
    lib = tt_irt.load_library()
    LL = ctypes.c_longlong
    lib.ttirt_shard_rows.argtypes = [LL, ctypes.c_int, ctypes.c_int, ctypes.POINTER(LL), ctypes.POINTER(LL)]
    lib.ttirt_auto_devices.argtypes = [LL, ctypes.c_int]
    for M in (0, 1, 7, 1000, 10001, (1 << 26), (1 << 31) - 1, (1 << 40) + 3):
        for G in (1, 2, 3, 4, 8):
            prev, sizes = 0, []
            for g in range(G):
                a, b = LL(-1), LL(-1)
                assert lib.ttirt_shard_rows(M, G, g, ctypes.byref(a), ctypes.byref(b)) == 0
                assert a.value == prev and b.value >= a.value
                sizes.append(b.value - a.value); prev = b.value
            assert prev == M and max(sizes) - min(sizes) <= 1
    a, b = LL(0), LL(0)
    assert lib.ttirt_shard_rows(10, 2, 2, ctypes.byref(a), ctypes.byref(b)) != 0
    assert lib.ttirt_shard_rows(-1, 2, 0, ctypes.byref(a), ctypes.byref(b)) != 0
    pol = lib.ttirt_auto_devices
    assert [pol(1 << 14, 8), pol(1 << 22, 8), pol(1 << 24, 8), pol(1 << 25, 8), pol(1 << 26, 8), pol(1 << 26, 2), pol(1 << 26, 0)] == [1, 1, 4, 8, 8, 2, 1]
