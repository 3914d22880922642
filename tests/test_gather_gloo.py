"""N>1 host logic on CPU: world_size-2 gloo run of the packed-detection all-gather (bench.py --gpus N uses the same
function over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dd3d_b200.gather import all_gather_detections, unpack


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake(rank, B, cap):
    g = torch.Generator().manual_seed(100 + rank)
    out = torch.randn(B, cap, 24, generator=g)
    counts = torch.randint(0, cap + 1, (B, ), generator=g, dtype=torch.int32)
    return out, counts


def _worker(rank, world, port, B, cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out, counts = _fake(rank, B, cap)
    g_out, g_cnt = all_gather_detections(out, counts)
    ok = True
    for r in range(world):
        o, c = _fake(r, B, cap)
        ok &= torch.equal(g_out[r * B:(r + 1) * B], o) and torch.equal(g_cnt[r * B:(r + 1) * B], c)
    per_image = unpack(g_out, g_cnt)
    ok &= len(per_image) == world * B and all(p.shape[0] == int(n) for p, n in zip(per_image, g_cnt))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_gather_detections_world2_gloo():
    world, B, cap = 2, 3, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
