"""N>1 host logic on CPU: world_size-2 gloo run of the packed-detection all-gather -- ONE collective on the packed
[dets | counts | flags] buffer (bench.py --gpus N runs the same classes over NCCL through the C ABI)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dd3d_b200.gather import DetectionGatherer, PackedDetections, all_gather_detections, unpack


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
    # the packed path the bench uses: views alias one buffer, a single collective moves dets + counts + flags
    calls = []
    orig = dist.all_gather_into_tensor
    dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    packed = PackedDetections(B, cap, "cpu")
    packed.out.copy_(out)
    packed.counts.copy_(counts)
    packed.flags.fill_(rank + 1)
    gat = DetectionGatherer(B, cap, "cpu")
    p_out, p_cnt, p_flags = gat.gather(packed)
    dist.all_gather_into_tensor = orig
    ok &= len(calls) == 1
    ok &= torch.equal(p_out, g_out) and torch.equal(p_cnt, g_cnt) and p_flags.tolist() == [1, 2]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_packed_layout_matches_the_c_abi():
    from dd3d_b200 import lib
    L = lib.load()
    for B, cap in ((1, 128), (32, 128), (7, 5000), (64, 160)):
        assert PackedDetections.packed_bytes(B, cap) == L.dd3d_packed_bytes(B, cap)
    p = PackedDetections(3, 8, "cpu")
    p.out[2, 7, 23] = 5.0
    p.counts[2] = 9
    p.flags[0] = 3
    assert p.buf.view(torch.float32)[3 * 8 * 24 - 1] == 5.0
    assert p.buf[3 * 8 * 96:].view(torch.int32)[:4].tolist() == [0, 0, 9, 3]


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
