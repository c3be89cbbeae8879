"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: parameter broadcast at construction,
gradient all-reduce (average) of the flat arena, and that ranks stay bit-identical afterwards.
The CUDA kernels are not involved; the N>1 numerics are checked on GPUs in tests/test_gpu_ddp.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pixelssl_b200.nn.arena import ParamArena, EngineParallel
        from pixelssl_b200.nn.modules import Conv2d, BatchNorm2d
        torch.manual_seed(100 + rank)                      # different init per rank on purpose
        net = torch.nn.Sequential(Conv2d(8, 8, 3, bias=False), BatchNorm2d(8), Conv2d(8, 4, 1, bias=True))
        wrap = EngineParallel(net)
        wrap.arena = ParamArena(net)                       # what .cuda() does, minus the device move
        wrap._setup_distributed()
        flat = wrap.arena.data.clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same_after_broadcast = all(torch.equal(g, gathered[0]) for g in gathered)
        bn_synced = all(m.sync_group is not None for m in net.modules() if isinstance(m, BatchNorm2d))
        # gradient averaging
        wrap.arena.zero_grad()
        wrap.arena.grad.fill_(float(rank + 1))
        wrap.arena.all_reduce_grads()
        avg = float(wrap.arena.grad[0])
        # the rank-aware sampler reads rank / world size from the process group; gloo ranks do not get the
        # CUDA-IPC statistics exchange
        import numpy as np
        from pixelssl_b200 import ops
        from pixelssl_b200.nn.data import TwoStreamBatchSampler
        np.random.seed(7)
        smp = TwoStreamBatchSampler(list(range(12)), list(range(100, 140)), 2, 3)
        batches = [list(map(int, b)) for b in smp]
        q.put((rank, same_after_broadcast, bn_synced, avg, float(wrap.arena.grad.min()), float(wrap.arena.grad.max()),
               (smp.rank, smp.world_size), batches, len(ops._peer_exchanges)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_broadcast_and_grad_average():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    by_rank = {}
    for rank, same, bn_synced, avg, lo, hi, who, batches, n_peer in res:
        assert same, 'parameters differ after the rank-0 broadcast'
        assert bn_synced, 'BN layers were not put into cross-rank statistics mode'
        assert avg == lo == hi == 1.5          # mean of 1 and 2
        assert who == (rank, 2) and n_peer == 0
        by_rank[rank] = batches
    # same permutations on both ranks, disjoint slices: together they are the reference's global batches (4 + 6)
    import numpy as np
    from pixelssl_b200.nn.data import TwoStreamBatchSampler
    np.random.seed(7)
    whole = [list(map(int, b)) for b in TwoStreamBatchSampler(list(range(12)), list(range(100, 140)), 4, 6, rank=0, world_size=1)]
    assert len(whole) == len(by_rank[0]) == len(by_rank[1])
    for gb, b0, b1 in zip(whole, by_rank[0], by_rank[1]):
        assert b0 == gb[0:2] + gb[4:7] and b1 == gb[2:4] + gb[7:10]
