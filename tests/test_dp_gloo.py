"""CPU, world_size 2 over gloo: the data-parallel gradient sync (bucketed all-reduce of the flat gradient buffer)
produces the MEAN of the per-shard gradients, which is what DistributedDataParallel gives the reference
(train.py:346)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import unet_pytorch_b200 as b2u
        from unet_pytorch_b200.trainer import FlatBuckets, GradientSync, _backward_order
        shapes = {k: v for k, v in b2u.vgg_unet_param_shapes(2).items()}
        order = _backward_order(list(shapes))
        lay = FlatBuckets(shapes, order, torch.device("cpu"), bucket_bytes=4 << 20)
        flat = lay.new_buffer()
        g = torch.Generator().manual_seed(100 + rank)
        flat.copy_(torch.randn(lay.total, generator=g))
        mine = flat.clone()
        sync = GradientSync(lay, flat)
        assert sync.enabled and sync.world == world
        sync.reset()
        # gradients become ready in backward order, two names (weight, bias) at a time like the engine reports them
        for i in range(0, len(order), 2):
            sync.ready(order[i:i + 2])
        scale = sync.finish()
        assert all(p == 0 for p in sync._pending)
        flat.mul_(scale)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = sum(gathered) / world
        torch.testing.assert_close(flat, expect, rtol=0, atol=1e-6)
        # frozen backbone: only decoder + head names are reported; encoder buckets must not fire or hang
        flat.copy_(mine)
        active = [n for n in order if not n.startswith("vgg.")]
        sync.reset(active)
        for i in range(0, len(active), 2):
            sync.ready(active[i:i + 2])
        sync.finish()
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_gradient_sync_is_mean_of_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"
