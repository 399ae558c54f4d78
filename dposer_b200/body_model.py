"""Body-model layer with the reference's call surface.

``BodyModel`` mirrors lib/body_model/body_model.py:8-112 and ``SMPLX`` mirrors
lib/body_model/smpl.py:49-78; both wrap third-party ``smplx==0.1.28`` in the reference.
Here the skinning arithmetic (shape/pose blendshapes, batched Rodrigues, kinematic chain,
skinning, vertex joints, landmarks) runs in libdposer_b200 (``dpb_lbs_forward`` /
``dpb_lbs_backward``); this file only maps arguments and keeps smplx's conventions
(default zero parameters of ``batch_size`` rows, pose concatenation order, output fields).

``bm_path`` may be a dict of body tensors (synthetic models: SMPL/SMPL-X files are licensed
and unavailable offline) or a path to a ``.npz`` / ``.pkl`` model file in the smplx layout.
"""
import ctypes as C
import os
import pickle

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
SMPLH_PARENTS = SMPL_PARENTS[:22] + [20, 22, 23, 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35] + \
    [21, 37, 38, 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50]
SMPLX_PARENTS = SMPL_PARENTS[:22] + [15, 15, 15] + \
    [20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38] + \
    [21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]
# smplx/vertex_ids.py: nose,reye,leye,rear,lear, LBigToe,LSmallToe,LHeel,RBigToe,RSmallToe,RHeel, 5+5 finger tips
SMPL_EXTRA_VIDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                   2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]
SMPLX_EXTRA_VIDS = [9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
                    5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022]
NUM_BODY_JOINTS = {'smpl': 23, 'smplh': 21, 'smplx': 21}
NUM_HAND_JOINTS = 15


class Struct(object):
    """smplx.utils.Struct."""

    def __init__(self, **kwargs):
        for key, val in kwargs.items():
            setattr(self, key, val)


class _LbsHandle:
    def __init__(self, ptr, device):
        self.ptr, self.device = ptr, device

    def __del__(self):
        try:
            if self.ptr:
                L.load().dpb_lbs_destroy(self.ptr)
        except Exception:
            pass


class _LbsFn(torch.autograd.Function):
    """verts, joints = LBS(betas, full_pose, transl).  need_verts=False skins only the vertices that the
    extra joints / landmarks read (SMPLify only uses joints, run/smplify.py:243-256)."""

    @staticmethod
    def forward(ctx, betas, full_pose, transl, core, need_verts, const_tail=False):
        h = core.handle(betas.device)
        B = betas.shape[0]
        dev = betas.device
        b = betas.detach().to(torch.float32).contiguous()
        p = full_pose.detach().to(torch.float32).contiguous()
        t = None if transl is None else transl.detach().to(torch.float32).contiguous()
        verts = torch.empty(B, core.V, 3, dtype=torch.float32, device=dev) if need_verts else None
        joints = torch.empty(B, core.n_out, 3, dtype=torch.float32, device=dev)
        wants_grad = betas.requires_grad or full_pose.requires_grad or (transl is not None and transl.requires_grad)
        # the workspace carries what backward reads: a forward that a backward may follow gets its own (another forward of
        # the same module must not overwrite it); grad mode is always off inside Function.forward, so ask the inputs
        ws = core.workspace(B, dev, fresh=wants_grad)
        flags = core.engine | (L.LBS_CONST_TAIL if (const_tail and core.tail is not None) else 0)
        if not wants_grad:
            flags |= L.LBS_NO_SAVE      # no backward will follow: skip the per-pose fp32 side outputs
        L.check(L.load().dpb_lbs_forward(h.ptr, L.ptr(b), L.ptr(p), L.ptr(t), L.ptr(verts), L.ptr(joints), B,
                                         flags, L.ptr(ws), ws.numel(), L.current_stream(dev)))
        ctx.core, ctx.need_verts, ctx.has_transl = core, need_verts, transl is not None
        ctx.tail_flag = flags & L.LBS_CONST_TAIL
        # an output the loss never touches arrives as None in backward (not as a [B,V,3] tensor of zeros): with no
        # vertex gradient the backward runs over the <=174 vertices the extra joints read instead of all of them
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(b, p)
        ctx.ws = ws
        if verts is None:
            verts = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(verts)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        core = ctx.core
        b, p = ctx.saved_tensors
        B, dev = b.shape[0], b.device
        h = core.handle(dev)
        gv = g_verts.contiguous().float() if (ctx.need_verts and g_verts is not None) else None
        gj = g_joints.contiguous().float() if g_joints is not None else None
        if gv is None and gj is None:
            return None, None, None, None, None, None
        g_pose = torch.empty_like(p)
        g_betas = torch.empty_like(b)
        g_transl = torch.empty(B, 3, dtype=torch.float32, device=dev) if ctx.has_transl else None
        ws = ctx.ws
        scratch, n_scratch = None, 0
        # scratch of the tensor-core vertex pass: all vertices (vertex cotangents) or the few the joints read
        if gv is not None:
            n_scratch = int(L.load().dpb_lbs_backward_scratch_bytes(h.ptr, B))
        elif B >= 64:
            n_scratch = int(L.load().dpb_lbs_backward_scratch_bytes_joints(h.ptr, B))
        if n_scratch:
            scratch = torch.empty(n_scratch, dtype=torch.uint8, device=dev)
        L.check(L.load().dpb_lbs_backward(h.ptr, L.ptr(b), L.ptr(p), L.ptr(gv), L.ptr(gj), L.ptr(g_pose),
                                          L.ptr(g_betas), L.ptr(g_transl), B, core.engine | ctx.tail_flag, L.ptr(ws), ws.numel(),
                                          L.ptr(scratch), n_scratch, L.current_stream(dev)))
        return g_betas, g_pose, g_transl, None, None, None


class LbsCore:
    """Immutable body tensors + the device handle (one per device)."""

    def __init__(self, tensors):
        t = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in tensors.items()}
        self.t = t
        self.V = int(t['v_template'].shape[0])
        self.J = int(t['J_regressor'].shape[0])
        self.S = int(t['shapedirs'].shape[2])
        self.parents = [int(p) for p in t['parents']]
        self.extra_vids = [int(v) for v in t.get('extra_vids', [])]
        lf = t.get('lmk_faces')
        self.n_lmk = 0 if lf is None else int(lf.shape[0])
        self.n_out = self.J + len(self.extra_vids) + self.n_lmk
        self.engine = L.ENGINE_AUTO       # tcgen05 blend for batches >= 64, fp32 blend otherwise (ENGINE_FP32 forces exact)
        self._handles = {}
        self._ws = {}
        self.tail = None                  # (n_var, [ (J-n_var)*3 ] axis-angle) declared with set_const_tail

    def set_const_tail(self, n_var, tail_pose):
        """Joints n_var..J-1 always carry ``tail_pose`` (hands / jaw / eyes at their default or mean pose): their
        pose-blend features are constants, folded into the template by ``dpb_lbs_set_const_tail``; calls that pass
        ``const_tail=True`` then run the blend with K = S + 9(n_var-1) instead of S + 9(J-1)."""
        tail = torch.as_tensor(tail_pose, dtype=torch.float32).detach().cpu().reshape(-1).contiguous()
        assert tail.numel() == (self.J - n_var) * 3
        self.tail = (int(n_var), tail)
        for h in self._handles.values():
            self._declare_tail(h)

    def _declare_tail(self, h):
        n_var, tail = self.tail
        a, p = L.host_f32(tail)
        L.check(L.load().dpb_lbs_set_const_tail(h.ptr, n_var, p), 'dpb_lbs_set_const_tail')

    def handle(self, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('dposer_b200 LBS runs on CUDA only (no CPU fallback)')
        key = device.index or 0
        h = self._handles.get(key)
        if h is not None:
            return h
        t = self.t
        m = L.BodyTensors()
        keep = []

        def f(x):
            a, p = L.host_f32(x)
            keep.append(a)
            return p

        def i(x):
            a, p = L.host_i32(x)
            keep.append(a)
            return p
        m.V, m.J, m.S = self.V, self.J, self.S
        m.v_template, m.shapedirs, m.posedirs = f(t['v_template']), f(t['shapedirs']), f(t['posedirs'])
        m.J_regressor, m.lbs_weights = f(t['J_regressor']), f(t['lbs_weights'])
        m.parents = i(np.asarray(self.parents, np.int32))
        m.n_extra = len(self.extra_vids)
        m.extra_vids = i(np.asarray(self.extra_vids if self.extra_vids else [0], np.int32))
        m.n_lmk = self.n_lmk
        if self.n_lmk:
            m.lmk_faces, m.lmk_bary = i(t['lmk_faces']), f(t['lmk_bary'])
        out = C.c_void_p()
        L.check(L.load().dpb_lbs_create(C.byref(out), C.byref(m), key), 'dpb_lbs_create')
        h = _LbsHandle(out, device)
        self._handles[key] = h
        if self.tail is not None:
            self._declare_tail(h)
        return h

    def workspace(self, B, device, fresh=False):
        """A fresh workspace per call when autograd will keep it alive (``fresh``); cached otherwise."""
        n = L.load().dpb_lbs_workspace_bytes(self.handle(device).ptr, int(B), 0)
        if fresh:
            return torch.empty(int(n), dtype=torch.uint8, device=device)
        key = (int(B), str(device))
        ws = self._ws.get(key)
        if ws is None:
            ws = torch.empty(int(n), dtype=torch.uint8, device=device)
            self._ws = {key: ws}
        return ws

    def __call__(self, betas, full_pose, transl=None, need_verts=True, const_tail=False):
        return _LbsFn.apply(betas, full_pose, transl, self, need_verts, const_tail)


def load_body_tensors(path, model_type, num_betas=10, num_expressions=10):
    """Read an smplx-layout .npz/.pkl model file into the tensor dict LbsCore expects.
    (Untested against real files in this repository: SMPL/SMPL-X are licensed and unavailable offline.)"""
    if os.path.isdir(path):
        raise NotImplementedError('pass the model file itself (e.g. SMPLX_NEUTRAL.npz), not a directory')
    if path.endswith('.npz'):
        d = dict(np.load(path, allow_pickle=True))
    else:
        with open(path, 'rb') as fh:
            d = pickle.load(fh, encoding='latin1')
    shapedirs = np.asarray(d['shapedirs'], np.float32)
    sd = shapedirs[:, :, :num_betas]
    if model_type == 'smplx':
        sd = np.concatenate([sd, shapedirs[:, :, 300:300 + num_expressions]], axis=2)
    V = shapedirs.shape[0]
    posedirs = np.asarray(d['posedirs'], np.float32).reshape(V * 3, -1).T      # [P, 3V]
    parents = np.asarray(d['kintree_table'])[0].astype(np.int64)
    parents[0] = -1
    J_reg = d['J_regressor']
    J_reg = np.asarray(J_reg.todense() if hasattr(J_reg, 'todense') else J_reg, np.float32)
    out = dict(v_template=torch.tensor(np.asarray(d['v_template'], np.float32)), shapedirs=torch.tensor(sd),
               posedirs=torch.tensor(np.ascontiguousarray(posedirs)), J_regressor=torch.tensor(J_reg),
               lbs_weights=torch.tensor(np.asarray(d['weights'], np.float32)), parents=parents.tolist(),
               faces=torch.tensor(np.asarray(d['f'], np.int64)),
               extra_vids=SMPLX_EXTRA_VIDS if model_type == 'smplx' else SMPL_EXTRA_VIDS)
    if model_type == 'smplx' and 'lmk_faces_idx' in d:
        faces = np.asarray(d['f'], np.int64)
        out['lmk_faces'] = torch.tensor(faces[np.asarray(d['lmk_faces_idx'], np.int64)].astype(np.int32))
        out['lmk_bary'] = torch.tensor(np.asarray(d['lmk_bary_coords'], np.float32))
    if 'hands_meanl' in d and 'hands_meanr' in d:
        out['hands_mean'] = torch.tensor(np.concatenate([np.asarray(d['hands_meanl'], np.float32).reshape(-1),
                                                         np.asarray(d['hands_meanr'], np.float32).reshape(-1)]))
    return out


class BodyModel(nn.Module):
    """lib/body_model/body_model.py:8-112 (wrapper semantics) on the native LBS kernels."""

    def __init__(self, bm_path, num_betas=10, batch_size=1, num_expressions=10, model_type='smplx'):
        super().__init__()
        assert (model_type in ['smpl', 'smplh', 'smplx'])
        tensors = bm_path if isinstance(bm_path, dict) else load_body_tensors(bm_path, model_type, num_betas,
                                                                             num_expressions)
        self.core = LbsCore(tensors)
        self.bm = self.core                       # the reference exposes the smplx layer as .bm
        self.model_type = model_type
        self.batch_size = batch_size
        self.num_betas = num_betas
        self.num_expressions = num_expressions if model_type == 'smplx' else 0
        self.num_joints = {'smpl': 23, 'smplh': 51, 'smplx': 54}[model_type]      # smplx NUM_JOINTS constants
        exp_J = {'smpl': 24, 'smplh': 52, 'smplx': 55}[model_type]
        if self.core.J != exp_J:
            raise ValueError(f'{model_type} expects {exp_J} joints, body tensors have {self.core.J}')
        if self.core.S != num_betas + self.num_expressions:
            raise ValueError(f'shapedirs has {self.core.S} components, expected num_betas+num_expressions='
                             f'{num_betas + self.num_expressions}')
        self.J_regressor = tensors['J_regressor'].detach().cpu().numpy()
        self.J_regressor_idx = {'pelvis': 0, 'lwrist': 20, 'rwrist': 21, 'neck': 12}
        faces = tensors.get('faces')
        self.register_buffer('faces_tensor', faces if faces is not None else torch.zeros(0, 3, dtype=torch.long))
        self.register_buffer('_dev', torch.zeros(1))
        if model_type in ('smplh', 'smplx'):
            # hands / jaw / eyes default to zero parameters (smplx, flat_hand_mean=True): a constant tail after
            # the 22 body joints whenever the caller omits them (every hot-path call of the reference does)
            self.core.set_const_tail(22, torch.zeros((self.core.J - 22) * 3))

    def _default(self, val, width):
        """smplx creates zero nn.Parameters [batch_size, width] for omitted inputs."""
        if val is not None:
            return val
        return torch.zeros(self.batch_size, width, dtype=torch.float32, device=self._dev.device)

    def forward(self, root_orient=None, pose_body=None, pose_hand=None, pose_jaw=None, pose_eye=None, betas=None,
                trans=None, dmpls=None, expression=None, return_dict=False, need_verts=True, **kwargs):
        assert (dmpls is None)
        nb = NUM_BODY_JOINTS[self.model_type]
        global_orient = self._default(root_orient, 3)
        body_pose = self._default(pose_body, nb * 3)
        betas_ = self._default(betas, self.num_betas)
        B = max(global_orient.shape[0], body_pose.shape[0], betas_.shape[0])
        if betas_.shape[0] != B and betas_.shape[0] == 1:
            betas_ = betas_.expand(B, -1)
        parts = [global_orient, body_pose]
        lh = rh = jaw = None
        if self.model_type in ('smplh', 'smplx'):
            lh = self._default(None if pose_hand is None else pose_hand[:, :NUM_HAND_JOINTS * 3], 45)
            rh = self._default(None if pose_hand is None else pose_hand[:, NUM_HAND_JOINTS * 3:], 45)
        if self.model_type == 'smplx':
            jaw = self._default(pose_jaw, 3)
            leye = self._default(None if pose_eye is None else pose_eye[:, :3], 3)
            reye = self._default(None if pose_eye is None else pose_eye[:, 3:], 3)
            parts += [jaw, leye, reye]
        if lh is not None:
            parts += [lh, rh]
        parts = [p if p.shape[0] == B else p.expand(B, -1) for p in parts]
        full_pose = torch.cat(parts, dim=1)
        shape = betas_
        if self.model_type == 'smplx':
            expr = self._default(expression, self.num_expressions)
            shape = torch.cat([betas_, expr if expr.shape[0] == B else expr.expand(B, -1)], dim=1)
        # smplx applies transl when given OR when the default (zero) parameter exists -> identical result
        const_tail = pose_hand is None and pose_jaw is None and pose_eye is None
        verts, joints = self.core(shape, full_pose, trans, need_verts, const_tail)
        out = {'v': verts if need_verts else None, 'f': self.faces_tensor, 'betas': betas_, 'Jtr': joints,
               'body_joints': joints[:22],           # reference slices the BATCH dim here (body_model.py:95)
               'pose_body': body_pose, 'full_pose': full_pose}
        if self.model_type in ['smplh', 'smplx']:
            out['pose_hand'] = torch.cat([lh, rh], dim=-1)
        if self.model_type == 'smplx':
            out['pose_jaw'] = jaw
            out['pose_eye'] = pose_eye
        return out if return_dict else Struct(**out)


# ---------------------------------------------------------------------------------------------
# SMPLify wrapper  (lib/body_model/smpl.py:49-78)
# ---------------------------------------------------------------------------------------------
JOINT_MAP_49 = [55, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65,
                8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]


def rot6d_to_axis_angle(rot6d):
    """Gram-Schmidt -> rotation matrix -> axis-angle (used once for the constant mean pose)."""
    x = rot6d.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = torch.nn.functional.normalize(a1, dim=1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=1)
    b3 = torch.cross(b1, b2, dim=1)
    R = torch.stack([b1, b2, b3], dim=-1)
    cos = ((R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]) - 1) / 2
    angle = torch.acos(cos.clamp(-1, 1))
    axis = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], dim=1)
    axis = axis / (2 * torch.sin(angle).clamp_min(1e-8))[:, None]
    return axis * angle[:, None]


SMPLOutput = Struct


class SMPLX(nn.Module):
    """lib/body_model/smpl.py:49-78: SMPL-X layer whose joints are re-indexed to the 49 SMPLify joints.
    The reference constructs ``smplx.SMPLX(model_path, **kw)`` with the smplx defaults (use_pca=True with zero PCA
    coefficients, flat_hand_mean=False), so ``full_pose`` = cat(orient, body, jaw, eyes, hands) + pose_mean carries the
    model file's constant NON-ZERO mean hand pose (``hands_meanl`` | ``hands_meanr``); jaw / eyes / expression are
    zero.  ``hand_mean`` [90] overrides it; ``flat_hand_mean=True`` gives zeros (what ``BodyModel`` uses)."""

    def __init__(self, model_path, batch_size=1, hand_mean=None, flat_hand_mean=False, **kwargs):
        super().__init__()
        tensors = model_path if isinstance(model_path, dict) else load_body_tensors(model_path, 'smplx')
        self.core = LbsCore(tensors)
        self.bm = self.core
        self.batch_size = batch_size
        mean = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'smpl_mean_params.npz'))
        self.register_buffer('mean_poses', rot6d_to_axis_angle(torch.tensor(mean['pose'], dtype=torch.float32))
                             .reshape(-1))                                        # [72]
        self.register_buffer('mean_shape', torch.tensor(mean['shape'], dtype=torch.float32))
        if hand_mean is None:
            hm = tensors.get('hands_mean')
            hand_mean = torch.zeros(90) if (flat_hand_mean or hm is None) else torch.as_tensor(hm)
        self.register_buffer('hand_mean', hand_mean.detach().float().reshape(90).clone())
        # jaw / eyes zero and hands at the constant mean pose in EVERY call: fold their pose blend into the template
        self.core.set_const_tail(22, torch.cat([torch.zeros(9), self.hand_mean.detach().cpu()]))
        faces = tensors.get('faces')
        self.faces = None if faces is None else faces.numpy()
        self.joint_map = torch.tensor(JOINT_MAP_49, dtype=torch.long)

    def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, transl=None, need_verts=False,
                **kwargs):
        dev = self.mean_shape.device
        B = body_pose.shape[0] if body_pose is not None else self.batch_size
        z = lambda w: torch.zeros(B, w, device=dev)           # noqa: E731
        global_orient = z(3) if global_orient is None else global_orient
        body_pose = z(63) if body_pose is None else body_pose
        betas = z(10) if betas is None else betas
        full_pose = torch.cat([global_orient, body_pose, z(9), self.hand_mean[None].expand(B, -1)], dim=1)
        S = self.core.S
        shape = torch.cat([betas, z(S - betas.shape[1])], dim=1) if S > betas.shape[1] else betas
        verts, joints = self.core(shape, full_pose, transl, need_verts, True)
        joints = joints[:, self.joint_map.to(dev), :]
        return SMPLOutput(vertices=verts if need_verts else None, global_orient=global_orient, body_pose=body_pose,
                          joints=joints, betas=betas, full_pose=full_pose)
