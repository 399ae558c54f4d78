"""Metrics of the hot path as device reductions (reference lib/utils/metric.py:8-37 and
lib/dataset/AMASS.py:263-324).  ``average_pairwise_distance`` runs on one GPU unless the caller opts into
sharding with an explicit ``group=`` (then every rank of the group must call it)."""
import numpy as np
import torch

from . import _lib as L
from .misc import BodyPartIndices, shard_range


def average_pairwise_distance(joints3d, group=None):
    """APD = mean over ordered pairs (i != j) of mean-over-joints L2 distance (metric.py:8-37).

    Single-process by default, like the reference (which calls it on one rank only).  Sharding is OPT-IN: pass
    ``group=`` (a process group, or ``torch.distributed.group.WORLD``) and call it on EVERY rank of that group with
    the same ``joints3d``; each rank then reduces its contiguous row block and the partial sums are all-reduced."""
    L.require_cuda(joints3d, 'joints3d')
    j = joints3d.detach().to(torch.float32).contiguous()
    B, nj = j.shape[0], j.shape[1]
    world, rank = 1, 0
    if group is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    row0, nrows = shard_range(B, world, rank)
    rows = torch.empty(max(nrows, 1), dtype=torch.float32, device=j.device)
    L.check(L.load().dpb_apd_partial(L.ptr(j), B, nj, row0, nrows, L.ptr(rows), L.current_stream(j.device)))
    out = rows[:nrows].double().sum().reshape(1)          # per-row fp32 sums, added in double
    if world > 1:
        dist.all_reduce(out, group=group)
    return (out / (B * (B - 1)))[0].float()


def mean_point_error_mm(a, b, idx=None):
    """Per-sample 1000 * mean_k ||a[:,k]-b[:,k]|| over an optional point subset (AMASS.py:286-296)."""
    L.require_cuda(a, 'a')
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(torch.float32).contiguous()
    B, n = a.shape[0], a.shape[1]
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    ix = None if idx is None else torch.as_tensor(idx, dtype=torch.int32, device=a.device).contiguous()
    L.check(L.load().dpb_mean_point_error(L.ptr(a), L.ptr(b), B, n, L.ptr(ix), 0 if ix is None else ix.numel(),
                                          L.ptr(out), L.current_stream(a.device)))
    return out


class Evaler:
    """lib/dataset/AMASS.py:263-324 -- MPVPE / MPJPE per sample (mm), min over hypotheses."""

    def __init__(self, body_model, part=None, vert_idx=None):
        self.body_model = body_model
        self.part = part
        if part is not None:
            self.joint_idx = np.array(getattr(BodyPartIndices, part)) + 1     # skip pelvis
            self.vert_idx = None if vert_idx is None else np.asarray(vert_idx)
        else:
            self.joint_idx, self.vert_idx = None, None

    def eval_bodys(self, outs, gts):
        body_gt = self.body_model(pose_body=gts)
        body_out = self.body_model(pose_body=outs)
        return {'mpvpe_all': mean_point_error_mm(body_out.v, body_gt.v, self.vert_idx).cpu().numpy(),
                'mpjpe_body': mean_point_error_mm(body_out.Jtr, body_gt.Jtr, self.joint_idx).cpu().numpy()}

    def multi_eval_bodys(self, outs, gts):
        res = [self.eval_bodys(outs[:, h], gts) for h in range(outs.shape[1])]
        return {k: np.min([r[k] for r in res], axis=0) for k in ('mpvpe_all', 'mpjpe_body')}
