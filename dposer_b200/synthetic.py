"""Synthetic stand-ins for the assets that cannot ship: random-init score
weights, SMPL / SMPL-X shaped body tensors, and the benchmark inputs derived from the AMASS
poses the reference does ship (examples/toy_data.npz, extracted to dposer_b200/data/).
Used by tests, bench.py and smoke() on both the GPU and the CPU-oracle side.
"""
import os
import types

import numpy as np
import torch

from .body_model import SMPL_PARENTS, SMPLX_PARENTS, SMPL_EXTRA_VIDS, SMPLX_EXTRA_VIDS

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def default_config():
    """configs/subvp/amass_scorefc_continuous.py + configs/default_amass_configs.py values that the
    hot path reads (attribute-style object; ml_collections is not needed)."""
    ns = types.SimpleNamespace
    return ns(
        data=ns(normalize=True, rot_rep='axis', min_max=False),
        training=ns(sde='subvpsde', continuous=True, batch_size=1280, n_iters=400001, auxiliary_loss=False, denoise_steps=10,
                    likelihood_weighting=False, reduce_mean=True),
        optim=ns(weight_decay=0, optimizer='Adam', lr=2e-4, beta1=0.9, eps=1e-8, warmup=5000, grad_clip=1.),
        sampling=ns(method='pc', predictor='euler_maruyama', corrector='none', n_steps_each=1, noise_removal=True,
                    probability_flow=False, snr=0.16),
        model=ns(type='ScoreModelFC', HIDDEN_DIM=1024, EMBED_DIM=512, N_BLOCKS=2, dropout=0.1, fourier_scale=16,
                 scale_by_sigma=True, ema_rate=0.9999, nonlinearity='swish', embedding_type='positional',
                 sigma_min=0.01, sigma_max=50, num_scales=1000, beta_min=0.1, beta_max=20.),
        eval=ns(batch_size=50, num_samples=500),
        seed=42,
        device=torch.device('cuda:0') if torch.cuda.is_available() else torch.device('cpu'),
    )


def make_score_model(seed=42, config=None):
    """SURVEY 8(d) "Score weights": default init under torch.manual_seed(seed) in the reference's module
    order, then every GroupNorm weight~U(0.5,1.5), bias~N(0,0.1^2) (same generator, module order)."""
    from .model import ScoreModelFC
    config = config or default_config()
    torch.manual_seed(seed)
    m = ScoreModelFC(config, n_poses=21, pose_dim=3, hidden_dim=1024, embed_dim=512, n_blocks=2)
    for mod in m.modules():
        if isinstance(mod, torch.nn.GroupNorm):
            with torch.no_grad():
                mod.weight.copy_(torch.rand(1024) + 0.5)
                mod.bias.copy_(torch.randn(1024) * 0.1)
    return m.eval()


def toy_poses():
    return torch.tensor(np.load(os.path.join(_DATA, 'toy_poses.npz'))['pose_samples'])      # [500,63] AMASS


def gesture_sequences():
    d = np.load(os.path.join(_DATA, 'gestures.npz'))
    return torch.tensor(d['pose_body']), torch.tensor(d['root_orient'])                     # [240,63], [240,3]


def make_body_tensors(model_type='smpl', seed=None, nnz=4, reg_nnz=32):
    """Random body model with SMPL (6890/24/10) or SMPL-X (10475/55/20) shapes -- recipe of SURVEY 8(d) c2/c4."""
    if model_type == 'smpl':
        V, J, S, parents, extra = 6890, 24, 10, SMPL_PARENTS, SMPL_EXTRA_VIDS
        seed = 7 if seed is None else seed
    elif model_type == 'smplx':
        V, J, S, parents, extra = 10475, 55, 20, SMPLX_PARENTS, SMPLX_EXTRA_VIDS
        seed = 8 if seed is None else seed
    else:
        raise ValueError(model_type)
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor([-0.9, -1.2, -0.2])
    hi = torch.tensor([0.9, 0.6, 0.2])
    v_template = lo + (hi - lo) * torch.rand(V, 3, generator=g)
    shapedirs = torch.randn(V, 3, S, generator=g) * 0.01
    posedirs = torch.randn((J - 1) * 9, V * 3, generator=g) * 2e-3

    def dirichlet(rows, k):
        e = -torch.log(torch.rand(rows, k, generator=g).clamp_min(1e-12))
        return e / e.sum(1, keepdim=True)
    J_regressor = torch.zeros(J, V)
    for j in range(J):
        idx = torch.randperm(V, generator=g)[:reg_nnz]
        J_regressor[j, idx] = dirichlet(1, reg_nnz)[0]
    lbs_weights = torch.zeros(V, J)
    cols = torch.stack([torch.randperm(J, generator=g)[:nnz] for _ in range(V)])
    lbs_weights.scatter_(1, cols, dirichlet(V, nnz))
    out = dict(v_template=v_template, shapedirs=shapedirs, posedirs=posedirs, J_regressor=J_regressor,
               lbs_weights=lbs_weights, parents=list(parents), extra_vids=list(extra),
               faces=torch.randint(0, V, (2 * V, 3), generator=g))
    if model_type == 'smplx':
        out['lmk_faces'] = torch.randint(0, V, (51, 3), generator=g).to(torch.int32)
        out['lmk_bary'] = dirichlet(51, 3)
        # smplx model files carry hands_meanl / hands_meanr (45 each); lib/body_model/smpl.py's SMPLX runs with the
        # smplx defaults (flat_hand_mean=False) so this constant NON-ZERO hand pose is part of every forward.
        # Separate generator: the tensors above stay bit-identical to the round-1 recipe.
        out['hands_mean'] = 0.2 * torch.randn(90, generator=torch.Generator().manual_seed(seed + 1000))
    return out


def lbs_inputs(B, model_type='smpl', seed=11):
    """c2 inputs: AMASS toy poses tiled to B + N(0,0.05^2) jitter (hands zero), root/betas/trans ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    toy = toy_poses()
    reps = (B + toy.shape[0] - 1) // toy.shape[0]
    body = toy.repeat(reps, 1)[:B] + 0.05 * torch.randn(B, 63, generator=g)
    root = torch.randn(B, 3, generator=g)
    trans = torch.randn(B, 3, generator=g)
    if model_type == 'smpl':
        betas = torch.randn(B, 10, generator=g)
        pose_body = torch.cat([body, torch.zeros(B, 6)], dim=1)           # [B,69], hand joints zero
        return dict(root_orient=root, pose_body=pose_body, betas=betas, trans=trans)
    betas = torch.randn(B, 10, generator=g)
    return dict(root_orient=root, pose_body=body, betas=betas, trans=trans)


def full_pose_from(inputs, model_type='smpl'):
    """The concatenated [B, J*3] pose / [B,S] shape the oracle's body_forward takes."""
    B = inputs['pose_body'].shape[0]
    if model_type == 'smpl':
        return torch.cat([inputs['root_orient'], inputs['pose_body']], dim=1), inputs['betas']
    pose = torch.cat([inputs['root_orient'], inputs['pose_body'], torch.zeros(B, 9 + 90)], dim=1)
    return pose, torch.cat([inputs['betas'], torch.zeros(B, 10)], dim=1)


def completion_inputs(n_partial=4096, hypotheses=10, seed=21, normalizer=None):
    """c3 inputs: normalised toy poses tiled (+N(0,0.05^2)), legs masked, x hypotheses."""
    from .misc import create_mask, Posenormalizer
    normalizer = normalizer or Posenormalizer(None, device='cpu', normalize=True, min_max=False, rot_rep='axis')
    g = torch.Generator().manual_seed(seed)
    toy = normalizer.offline_normalize(toy_poses())
    reps = (n_partial + toy.shape[0] - 1) // toy.shape[0]
    poses = toy.repeat(reps, 1)[:n_partial] + 0.05 * torch.randn(n_partial, 63, generator=g)
    torch.manual_seed(seed)
    mask, obs = create_mask(poses, part='legs')
    return poses, mask.repeat(hypotheses, 1), obs.repeat(hypotheses, 1)
