"""2-GPU NCCL test: one process per GPU, each with (1 labeled + 1 unlabeled) images, must reproduce
the single-GPU step on the combined (2 + 2) batch: same losses, same averaged gradients, same BN
running statistics (the reference synchronises BN statistics across replicas,
sync_batchnorm/batchnorm.py:55-78).  Skipped unless two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg(bs, ubs):
    return {'ssl_algorithm': 'ssl_mt', 'cons_for_labeled': False, 'cons_scale': 1.0, 'cons_rampup_epochs': 0,
            'ema_decay': 0.99, 'lr': 0.00025, 'momentum': 0.9, 'weight_decay': 0.0005, 'epochs': 2,
            'batch_size': bs, 'unlabeled_batch_size': ubs, 'log_freq': 10 ** 6}


def _run_step(alg, img, lab):
    alg._train([((img,), (lab,))], 0)
    sp = dict(alg.s_model.module.model.named_parameters())
    grads = torch.cat([p.grad.contiguous().reshape(-1) for p in sp.values()]).cpu()
    params = alg.s_model.arena.data.cpu().clone()
    bufs = torch.cat([b.reshape(-1).float() for n, b in alg.s_model.named_buffers() if 'num_batches' not in n]).cpu()
    return (float(alg.meters['s_task_loss'].val), float(alg.meters['cons_loss'].val), grads, params, bufs)


def _worker(rank, world, port, size, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import logging
        logging.getLogger('PixelSSL').setLevel(logging.ERROR)
        from pixelssl_b200 import runner, ops
        from oracle import sseg_oracle as O
        ops.set_conv_precision('fp32')
        alg = runner.build_algorithm(runner.build_args(_cfg(2, 1), iters_per_epoch=5))
        st = {k: v for k, v in O.randomize_bn_affine(O.init_deeplabv2(71, cls_bias_std=0.01), 72).items()}
        alg.s_model.load_state_dict({'module.model.' + k: v for k, v in st.items()})
        alg.t_model.load_state_dict({'module.model.' + k: v for k, v in st.items()})
        img, lab = O.synthetic_batch(500, 4, 2, size, size)          # global batch [L0, L1, U0, U1]
        idx = [rank, 2 + rank]
        out = _run_step(alg, img[idx].contiguous(), lab[idx].contiguous())
        if rank == 0:
            q.put(tuple(o if not torch.is_tensor(o) else o.numpy() for o in out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_step_equals_single_gpu_big_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    size = 65
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, size, q)) for r in range(2)]
    for p in procs:
        p.start()
    l2, c2, g2, p2, b2 = q.get(timeout=280)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single GPU, combined batch
    import logging
    logging.getLogger('PixelSSL').setLevel(logging.ERROR)
    from pixelssl_b200 import runner, ops
    from oracle import sseg_oracle as O
    ops.set_conv_precision('fp32')
    alg = runner.build_algorithm(runner.build_args(_cfg(4, 2), iters_per_epoch=5))
    st = O.randomize_bn_affine(O.init_deeplabv2(71, cls_bias_std=0.01), 72)
    alg.s_model.load_state_dict({'module.model.' + k: v for k, v in st.items()})
    alg.t_model.load_state_dict({'module.model.' + k: v for k, v in st.items()})
    # the reference's multi-replica BN path uses clamp(var, eps) where the single-replica path uses
    # var + eps (batchnorm.py:50-53 vs :125); the engine mirrors both, so force the same formula here
    from pixelssl_b200.nn.modules import BatchNorm2d
    for m in list(alg.s_model.modules()) + list(alg.t_model.modules()):
        if isinstance(m, BatchNorm2d):
            m.multi_replica_formula = True
    img, lab = O.synthetic_batch(500, 4, 2, size, size)
    l1, c1, g1, p1, b1 = _run_step(alg, img, lab)
    g1, p1, b1 = g1.numpy(), p1.numpy(), b1.numpy()
    # rank-0 loss is the mean over ITS samples; gradients/params/BN buffers must match the big batch
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    print('ddp vs big batch: grad %.2e params %.2e bn buffers %.2e' % (rel(g2, g1), rel(p2, p1), rel(b2, b1)))
    assert rel(b2, b1) <= 1e-4        # synchronised batch statistics
    assert rel(p2, p1) <= 5e-4        # = lr * gradient noise (10x lr on the classifier), measured 1.9e-4
    assert rel(g2, g1) <= 5e-2        # whole-network gradient: fp32-noise floor of this net (see test_gpu_model)


def test_batch_permutation_noise_is_the_yardstick_for_ddp_parity():
    """Root cause of the 2-rank vs big-batch gradient difference (2.8e-2 max-relative at 65x65, exact-fp32 mode):
    the two runs add the SAME numbers in a different order - BN statistics are fp32 partial sums per 128-row tile
    finished in fp64, and the row -> tile assignment of a 4-image batch differs from that of two 2-image shards - and
    the 101-layer random-init network on 5x5 feature maps (BN over 50..100 samples) amplifies that last-bit noise
    chaotically (tests/test_gpu_model.py).  This single-GPU test shows it without any second GPU: the big-batch step
    on [L0, L1, U0, U1] against the same step on the permuted batch [L1, L0, U1, U0] is mathematically the identical
    loss, yet its gradient moves by the same order of magnitude.  The data-parallel test above and bench.py's
    ddp_check are held to a small multiple of this measured noise."""
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    import logging
    logging.getLogger('PixelSSL').setLevel(logging.ERROR)
    from pixelssl_b200 import runner, ops
    from oracle import sseg_oracle as O
    from pixelssl_b200.nn.modules import BatchNorm2d
    ops.set_conv_precision('fp32')
    st = O.randomize_bn_affine(O.init_deeplabv2(71, cls_bias_std=0.01), 72)
    img, lab = O.synthetic_batch(500, 4, 2, 65, 65)
    outs = []
    for perm in ([0, 1, 2, 3], [1, 0, 3, 2]):
        alg = runner.build_algorithm(runner.build_args(_cfg(4, 2), iters_per_epoch=5))
        alg.s_model.load_state_dict({'module.model.' + k: v for k, v in st.items()})
        alg.t_model.load_state_dict({'module.model.' + k: v for k, v in st.items()})
        for m in list(alg.s_model.modules()) + list(alg.t_model.modules()):
            if isinstance(m, BatchNorm2d):
                m.multi_replica_formula = True
        outs.append(_run_step(alg, img[perm].contiguous(), lab[perm].contiguous()))
        del alg
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    noise_g, noise_p, noise_b = (rel(outs[1][i], outs[0][i]) for i in (2, 3, 4))
    print('batch-permutation noise: grad %.2e params %.2e bn buffers %.2e' % (noise_g, noise_p, noise_b))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-5 * abs(outs[0][0])            # the loss itself is well conditioned
    assert noise_b <= 1e-4 and noise_p <= 5e-4
    assert 1e-4 <= noise_g <= 5e-2, noise_g        # the gradient is not: same order as the 2-rank difference (2.8e-2)


def _peer_worker(rank, world, port, q):
    import ctypes
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from pixelssl_b200.nn.peer import PeerExchange
        from pixelssl_b200._lib import call
        px = PeerExchange(dist.group.WORLD)
        ok = True
        worst = 0.0
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for it, C in enumerate([4, 64, 256, 2048, 64, 1024, 8, 2048, 512, 64]):     # > NSLOT exchanges: slots recycle
            g = torch.Generator().manual_seed(100 * it + rank)
            rows = 1000.0
            x = torch.randn(int(rows), C, generator=g, dtype=torch.float64) * (1 + rank) + 0.3 * rank
            local = torch.cat((x.sum(0), (x * x).sum(0))).cuda()
            want = local.clone()
            dist.all_reduce(want)                                   # NCCL result as the yardstick
            # plain exchange (backward dsums)
            got = px.allreduce_bn(local.clone())
            worst = max(worst, float((got - want).abs().max() / want.abs().max()))
            # fused with finalize
            gamma = torch.rand(C, generator=g).cuda() + 0.5
            beta = torch.randn(C, generator=g).cuda()
            rm_a, rv_a = torch.zeros(C).cuda(), torch.ones(C).cuda()
            rm_b, rv_b = rm_a.clone(), rv_a.clone()
            ca = torch.empty(4, C).cuda()
            cb = torch.empty(4, C).cuda()
            s2 = px.allreduce_bn(local.clone(), (rows * world, C, gamma, beta, rm_a, rv_a, 0.1, 1e-5, 1, ca[0], ca[1], ca[2], ca[3]))
            call('pxl_bn_finalize', P(s2), rows * world, C, P(gamma), P(beta), P(rm_b), P(rv_b), 0.1, 1e-5, 1,
                 P(cb[0]), P(cb[1]), P(cb[2]), P(cb[3]), st)
            torch.cuda.synchronize()
            ok = ok and torch.equal(ca, cb) and torch.equal(rm_a, rm_b) and torch.equal(rv_a, rv_b)
            # every rank must hold bit-identical totals
            both = [torch.empty_like(s2) for _ in range(world)]
            dist.all_gather(both, s2)
            ok = ok and all(torch.equal(both[0], b) for b in both)
        status = px.status()
        dist.barrier()
        px.close()
        if rank == 0:
            q.put((ok, worst, status))
    finally:
        dist.destroy_process_group()


def test_peer_memory_bn_exchange_two_gpus():
    """csrc/peer_exchange.cu: the NVLink mailbox all-reduce equals NCCL's (fp64, 1e-15), its fused finalize equals
    pxl_bn_finalize bit for bit, all ranks get identical totals, slots recycle, watchdog silent."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, worst, status = q.get(timeout=200)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == 0
    assert worst <= 1e-14, worst
    assert ok
