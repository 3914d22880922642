"""Whole-batch gather of packed detections across ranks.

Replaces the reference's `d2_comm.gather` of pickled python lists (kitti_3d_evaluator.py:152-164,
nuscenes_evaluator.py:255) with ONE collective on the fixed-stride `[B][out_cap][24]` detection buffer (+ counts):
`ncclAllGather` on GPUs (torch.distributed "nccl"), `gloo` for the CPU tests.  Images are sharded by batch, rank r
owns global images [r*B, (r+1)*B)."""
import torch
import torch.distributed as dist


def all_gather_detections(out, counts, group=None):
    """out: [B, cap, 24] fp32, counts: [B] int32 (same device).  Returns ([world*B, cap, 24], [world*B]) on every rank."""
    world = dist.get_world_size(group)
    g_out = torch.empty((world * out.shape[0], ) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    g_cnt = torch.empty((world * counts.shape[0], ), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(g_out, out.contiguous(), group=group)
    dist.all_gather_into_tensor(g_cnt, counts.contiguous(), group=group)
    return g_out, g_cnt


def unpack(g_out, g_cnt):
    """Per-image list of [n_i, 24] views in global image order."""
    return [g_out[i, :int(n)] for i, n in enumerate(g_cnt.tolist())]
