"""N > 1 host logic on CPU: world_size-2 gloo processes -- one packed weight broadcast, disjoint
contiguous frame shards, no other collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from impersonator_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init on every rank
    net = torch.nn.Sequential(torch.nn.Conv2d(6, 8, 3, bias=False), torch.nn.InstanceNorm2d(8, affine=True))
    extra = torch.full((1, 3, 4, 4), float(rank))
    got = sharding.broadcast_module(net, extras=[extra], src=0)
    sig = float(sum(p.double().sum() for p in net.state_dict().values()))
    a, b = sharding.shard_range(37, rank, world)
    q.put((rank, sig, float(got[0].mean()), a, b))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_shards_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                       # identical weights after the one broadcast
    assert res[0][2] == 0.0 and res[1][2] == 0.0        # extras came from rank 0
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 19, 19, 37)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    sd = {"b.weight": torch.randn(4, 3), "a.bias": torch.randn(5)}
    flat, layout = sharding.pack_state(sd, [torch.ones(2, 2)])
    back = sharding.unpack_state(flat, layout)
    assert torch.equal(back["b.weight"], sd["b.weight"]) and torch.equal(back["a.bias"], sd["a.bias"])
    assert torch.equal(back["__extra0"], torch.ones(2, 2))
