"""CPU, world_size 2 over gloo: the multi-GPU path's host logic — contiguous link sharding and the time-ordered merge of
per-rank TP lists on rank 0. Per-rank TP production is stood in for by the oracle (test infrastructure); on GPUs the
same code path is fed by TPGenerator (tests/test_gpu_parity.py::test_sharded_equals_unsharded)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from fdreadoutlibs_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import fdreadoutlibs_b200 as S
    from fdreadoutlibs_b200 import frames as F
    from oracle import binding as B

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_links, n_units = 80, 6  # two APAs
    link0, n = sharding.shard_links(n_links, world, rank)
    p = S.gen_params(5, 0.3)
    fr = S.gen_wibeth_host(p, n, n_units, link0=link0)  # this rank generates only its shard
    cfg = B.make_config(threshold=30)
    local, _ = B.oracle_process_links(cfg, fr)
    local = S.sort_tps(sharding.globalise(local, link0))
    merged = sharding.gather_and_merge(local)
    if rank == 0:
        full, _ = B.oracle_process_links(cfg, S.gen_wibeth_host(p, n_links, n_units))
        full = S.sort_tps(full)
        ok = merged.size == full.size and bool((merged == full).all())
        key = list(zip(merged["time_start"].tolist(), merged["link"].tolist(), merged["channel"].tolist()))
        q.put((ok, key == sorted(key), int(merged.size), (link0, n)))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_links_blocks():
    # whole APAs stay together
    assert [sharding.shard_links(6000, 8, r) for r in range(8)][0] == (0, 720)
    spans = [sharding.shard_links(6000, 8, r) for r in range(8)]
    assert sum(n for _, n in spans) == 6000 and all(l0 % 40 == 0 and n % 40 == 0 for l0, n in spans)
    assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(7))
    # fewer APAs than ranks: split links evenly
    assert [sharding.shard_links(40, 8, r)[1] for r in range(8)] == [5] * 8
    assert sharding.shard_links(7, 2, 0) == (0, 3) and sharding.shard_links(7, 2, 1) == (3, 4)
    with pytest.raises(ValueError):
        sharding.shard_links(10, 2, 2)


def test_two_rank_shard_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, ordered, n, span = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and ordered and n > 100 and span == (0, 40)
