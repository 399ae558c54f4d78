"""Training step on the native path against the REAL reference (tests/golden/train_golden.npz, made by
tests/golden/make_golden_train.py with replayed draws) and against the CPU oracle (oracle/train_ref.py)."""
import ctypes as C
import types

import numpy as np
import pytest
import torch

from conftest import golden
from dposer_b200 import _lib as L
from dposer_b200 import losses, sde_lib, synthetic
from dposer_b200.ema import ExponentialMovingAverage

pytestmark = pytest.mark.gpu
B, STEPS = 96, 3
SAMPLE = 48


def sample_idx(n):
    return torch.linspace(0, n - 1, min(SAMPLE, n)).long()


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu().reshape(-1), torch.as_tensor(b).double().cpu().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (96, 63, 1024), (1280, 1024, 1024), (1024, 512, 1280), (63, 1024, 96),
                                   (300, 5120, 512), (257, 129, 200)])
def test_split_gemm_matches_fp64(M, N, K):
    """dpb_gemm_nt (the GEMM every contraction of the training step uses): C = A B^T + bias vs float64, ragged sizes."""
    g = torch.Generator().manual_seed(M * 7 + N)
    A, Bm, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    ref = A.double() @ Bm.double().T + bias.double()
    Ad, Bd, bd = A.cuda(), Bm.cuda(), bias.cuda()
    out = torch.full((M, N), float('nan'), device='cuda')
    lib = L.load()
    ws = torch.empty(int(lib.dpb_gemm_nt_workspace_bytes(M, N, K)), dtype=torch.uint8, device='cuda')
    L.check(lib.dpb_gemm_nt(L.ptr(Ad), L.ptr(Bd), L.ptr(bd), L.ptr(out), M, N, K, L.ptr(ws), ws.numel(),
                            L.current_stream(Ad.device)))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert rel(out, ref) < 1e-5


def _state(cfg, model):
    opt = losses.get_optimizer(cfg, model.parameters())
    ema = ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
    return dict(optimizer=opt, model=model, ema=ema, step=0)


def _masks(g, s):
    return torch.tensor(np.unpackbits(g[f's{s}_masks'], axis=-1)).reshape(5, B, 1024)


def test_train_steps_vs_reference_golden():
    """Three get_step_fn(train=True) steps (dropout, warm-up, clipping, Adam, EMA) with the reference's draws replayed:
    loss, gradients, gradient norm after every step; parameter / EMA / Adam-moment deltas at the end."""
    g = golden('train_golden.npz')
    cfg = synthetic.default_config()
    model = synthetic.make_score_model(42).cuda()
    model.train()
    names = [n for n, _ in model.named_parameters()]
    p0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    state = _state(cfg, model)
    state['step'] = int(g['step0'])
    sde = sde_lib.subVPSDE(0.1, 20., 1000)
    step_fn = losses.get_step_fn(sde, train=True, optimize_fn=losses.optimization_manager(cfg), reduce_mean=True,
                                 continuous=True, likelihood_weighting=False)
    data = torch.tensor(g['data']).cuda()
    for s in range(STEPS):
        ld = step_fn(state, data[s * B:(s + 1) * B], t=torch.tensor(g[f's{s}_t']), z=torch.tensor(g[f's{s}_z']),
                     drop_mask=_masks(g, s))
        assert abs(float(ld['step_loss']) - float(g[f's{s}_loss'])) < 2e-4 * float(g[f's{s}_loss']), s
        assert abs(state['optimizer'].grad_norm() - float(g[f's{s}_gnorm'])) < 2e-4 * float(g[f's{s}_gnorm']), s
        for n, p in model.named_parameters():
            if n.startswith('pre_dense_cond'):
                assert float(p.grad.abs().max()) == 0.0
                continue
            flat = p.grad.reshape(-1)
            gs = g[f's{s}_g_{n}']
            gmax = float(flat.abs().max())
            assert float((flat[sample_idx(flat.numel()).cuda()].cpu() - torch.tensor(gs)).abs().max()) < 3e-4 * gmax, (s, n)
            tot, atot = g[f's{s}_gsum_{n}']
            assert abs(float(flat.double().sum()) - tot) < 3e-4 * atot + 1e-12, (s, n)
    assert state['step'] == int(g['step0']) + STEPS and state['ema'].num_updates == int(g['ema_num_updates'])
    opt = state['optimizer']
    for i, (n, p) in enumerate(model.named_parameters()):
        d = (p.detach() - p0[n]).reshape(-1)
        idx = sample_idx(d.numel()).cuda()
        dmax = max(float(np.abs(g[f'dp_{n}']).max()), 1e-12)
        # Adam turns an element whose three gradients nearly cancel into a +-lr step: tolerance relative to the largest delta
        assert float((d[idx].cpu() - torch.tensor(g[f'dp_{n}'])).abs().max()) < 3e-2 * dmax + 1e-9, n
        assert abs(float(d.double().norm()) - float(g[f'dpnorm_{n}'])) < 1e-2 * float(g[f'dpnorm_{n}']) + 1e-12, n
        e = (state['ema'].shadow_params[i] - p0[n]).reshape(-1)
        assert float((e[idx].cpu() - torch.tensor(g[f'dema_{n}'])).abs().max()) < 3e-2 * dmax + 1e-9, n
        if f'm_{n}' in g.files:
            st = opt.state[p]
            m = st['exp_avg'].reshape(-1)[idx].cpu()
            v = st['exp_avg_sq'].reshape(-1)[idx].cpu()
            assert float((m - torch.tensor(g[f'm_{n}'])).abs().max()) < 3e-4 * float(st['exp_avg'].abs().max()), n
            assert float((v - torch.tensor(g[f'v_{n}'])).abs().max()) < 6e-4 * float(st['exp_avg_sq'].abs().max()), n
    # ---- evaluation step: EMA weights, no dropout, parameters restored afterwards (losses.py:263-271)
    model.eval()
    before = [p.detach().clone() for p in model.parameters()]
    eval_fn = losses.get_step_fn(sde, train=False, reduce_mean=True, continuous=True, likelihood_weighting=False)
    ld = eval_fn(state, data[:B], t=torch.tensor(g['eval_t']), z=torch.tensor(g['eval_z']))
    assert abs(float(ld['step_loss']) - float(g['eval_loss'])) < 3e-4 * float(g['eval_loss'])
    assert all(torch.equal(a, p.detach()) for a, p in zip(before, model.parameters()))
    # ---- loss variants on the updated weights (eval mode): value, total gradient norm, one gradient in full
    vp, ve = sde_lib.VPSDE(0.1, 20., 1000), sde_lib.VESDE(0.01, 50., 1000)
    cases = [('lw', losses.get_sde_loss_fn(sde, True, reduce_mean=True, likelihood_weighting=True), 't'),
             ('sum', losses.get_sde_loss_fn(sde, True, reduce_mean=False, likelihood_weighting=False), 't'),
             ('ddpm', losses.get_ddpm_loss_fn(vp, True, reduce_mean=True), 'labels'),
             ('smld', losses.get_smld_loss_fn(ve, True, reduce_mean=False), 'labels')]
    model.config.model.dropout = 0.0           # the reference evaluated these in eval mode (gradients still wanted)
    for tag, fn, key in cases:
        loss = fn(model, data[:B], None, None, z=torch.tensor(g[f'{tag}_z']), **{key: torch.tensor(g[f'{tag}_{key}'])})
        assert abs(float(loss) - float(g[f'{tag}_loss'])) < 3e-4 * abs(float(g[f'{tag}_loss'])), tag
        assert abs(opt.grad_norm() - float(g[f'{tag}_gnorm'])) < 5e-4 * float(g[f'{tag}_gnorm']), tag
        assert rel(model.post_dense.bias.grad, g[f'{tag}_g_post']) < 5e-4, tag


def test_train_step_vs_oracle_other_batch_sizes():
    """Native loss + gradients vs autograd through the CPU oracle at batch sizes that are not tile multiples."""
    from oracle import score_ref as S
    from oracle import train_ref as T
    sde, osde = sde_lib.subVPSDE(0.1, 20., 1000), S.SubVP(0.1, 20., 1000)
    sd = S.make_state_dict(42)
    model = synthetic.make_score_model(42).cuda()
    model.train()
    for Bn, lw in [(1, False), (37, True), (130, False)]:
        gen = torch.Generator().manual_seed(Bn)
        batch = torch.randn(Bn, 63, generator=gen)
        t = torch.rand(Bn, generator=gen) * (1 - 1e-5) + 1e-5
        z = torch.randn(Bn, 63, generator=gen)
        masks = (torch.rand(5, Bn, 1024, generator=gen) >= 0.1).to(torch.uint8)
        names = T.param_names(sd)
        leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
        full = dict(sd)
        full.update(leaves)
        ol = T.sde_loss(full, osde, batch, t, z, masks, 0.1, True, lw)
        used = [k for k in names if not k.startswith('pre_dense_cond')]
        og = dict(zip(used, torch.autograd.grad(ol, [leaves[k] for k in used])))
        fn = losses.get_sde_loss_fn(sde, True, reduce_mean=True, likelihood_weighting=lw)
        loss = fn(model, batch.cuda(), None, None, t=t, z=z, drop_mask=masks)
        ol = ol.detach()
        assert abs(float(loss) - float(ol)) < 2e-4 * abs(float(ol)), Bn
        for n, p in model.named_parameters():
            if n in og:
                assert rel(p.grad, og[n]) < 5e-4, (Bn, n)


def test_train_philox_mode_is_deterministic_and_updates_the_sampler_weights():
    """Device-side draws (Philox z and dropout masks): same seed -> bit-identical loss and gradients; after a step the
    inference handle is rebuilt from the new weights; optimiser / EMA state dicts interoperate with torch's."""
    cfg = synthetic.default_config()
    cfg.optim.warmup = 0
    model = synthetic.make_score_model(42).cuda()
    model.train()
    state = _state(cfg, model)
    sde = sde_lib.subVPSDE(0.1, 20., 1000)
    data = synthetic.toy_poses()[:200].cuda()
    fn = losses.get_sde_loss_fn(sde, True, reduce_mean=True)
    runs = []
    for _ in range(2):
        torch.manual_seed(7)
        loss = fn(model, data, None, None)
        runs.append((float(loss), state['optimizer'].flat_g.clone()))
    assert np.isfinite(runs[0][0]) and runs[0][0] == runs[1][0] and torch.equal(runs[0][1], runs[1][1])
    torch.manual_seed(8)
    assert float(fn(model, data, None, None)) != runs[0][0]
    # one optimiser step changes what the sampler's score network computes
    model.eval()
    x, lab = torch.randn(5, 63).cuda(), torch.full((5,), 300.)
    before = model(x, lab).clone()
    model.train()
    step_fn = losses.get_step_fn(sde, train=True, optimize_fn=losses.optimization_manager(cfg), reduce_mean=True)
    step_fn(state, data)
    model.eval()
    after = model(x, lab)
    assert float((after - before).abs().max()) > 0
    from oracle import score_ref as S
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    model.engine = L.ENGINE_FP32
    assert rel(model(x, lab), S.score_model_forward(sd, x.cpu(), lab)) < 2e-5
    model.engine = L.ENGINE_AUTO
    # state dicts: torch.optim.Adam on CPU copies accepts ours and vice versa; EMA round trip
    cpu_params = [torch.nn.Parameter(p.detach().cpu().clone()) for p in model.parameters()]
    ref_opt = torch.optim.Adam(cpu_params, lr=cfg.optim.lr)
    sdict = state['optimizer'].state_dict()
    ref_opt.load_state_dict(sdict)
    assert float(ref_opt.state[cpu_params[0]]['step']) == 1.0
    assert torch.equal(ref_opt.state[cpu_params[0]]['exp_avg'], state['optimizer'].state[next(model.parameters())]['exp_avg'].cpu())
    state['optimizer'].load_state_dict(ref_opt.state_dict())
    assert state['optimizer']._step == 1
    ema2 = ExponentialMovingAverage(model.parameters(), decay=0.5)
    ema2.load_state_dict(state['ema'].state_dict())
    assert ema2.decay == cfg.model.ema_rate and ema2.num_updates == 1
    assert all(torch.equal(a, b) for a, b in zip(ema2.shadow_params, state['ema'].shadow_params))


def test_graphed_train_step_matches_eager_step():
    """get_step_fn(graph=True): the captured step (loss + backward + clip + Adam + EMA as one CUDA graph, schedule values and
    the Philox seed refreshed through one device buffer) reproduces the eager native step bit for bit over several steps,
    including the warm-up learning rate and the EMA decay schedule."""
    cfg = synthetic.default_config()
    cfg.optim.warmup = 4
    sde = sde_lib.subVPSDE(0.1, 20., 1000)
    data = synthetic.toy_poses()[:320].cuda()
    res = []
    for graph in (False, True):
        model = synthetic.make_score_model(42).cuda()
        model.train()
        state = _state(cfg, model)
        step_fn = losses.get_step_fn(sde, train=True, optimize_fn=losses.optimization_manager(cfg), reduce_mean=True,
                                     graph=graph)
        torch.manual_seed(5)
        ls = [float(step_fn(state, data[(i % 2) * 160:(i % 2) * 160 + 160])['step_loss']) for i in range(6)]
        res.append((ls, state['optimizer'].flat_p.clone(), state['ema']._flat.clone(), state['optimizer'].flat_m.clone(),
                    state['step'], state['ema'].num_updates, state['optimizer']._step))
    a, b = res
    assert a[0] == b[0], (a[0], b[0])
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert a[4:] == b[4:] == (6, 6, 6)


def test_checkpoint_resume_continues_identically(tmp_path):
    """run/train.py:182-191,392-403: a checkpoint in the reference's layout (model_state_dict / optimizer_state_dict / ema /
    step) written after two steps and loaded into a fresh model + optimiser + EMA continues bit for bit."""
    cfg = synthetic.default_config()
    cfg.optim.warmup = 3
    sde = sde_lib.subVPSDE(0.1, 20., 1000)
    data = synthetic.toy_poses()[:128].cuda()

    def fresh(seed):
        model = synthetic.make_score_model(seed).cuda()
        model.train()
        return _state(cfg, model)
    step_fn = losses.get_step_fn(sde, train=True, optimize_fn=losses.optimization_manager(cfg), reduce_mean=True)
    a = fresh(42)
    torch.manual_seed(1)
    for _ in range(2):
        step_fn(a, data)
    path = str(tmp_path / 'checkpoint-step2.pth')
    torch.save({'epoch': 1, 'model_state_dict': a['model'].state_dict(), 'optimizer_state_dict': a['optimizer'].state_dict(),
                'ema': a['ema'].state_dict(), 'step': a['step']}, path)
    torch.manual_seed(2)
    la = float(step_fn(a, data)['step_loss'])
    b = fresh(7)                                            # different weights: everything must come from the file
    ck = torch.load(path, map_location='cuda', weights_only=False)
    b['model'].load_state_dict(ck['model_state_dict'])
    b['optimizer'].load_state_dict(ck['optimizer_state_dict'])
    b['ema'].load_state_dict(ck['ema'])
    b['step'] = ck['step']
    torch.manual_seed(2)
    lb = float(step_fn(b, data)['step_loss'])
    assert la == lb and b['step'] == 3 and b['ema'].num_updates == 3
    assert torch.equal(a['optimizer'].flat_p, b['optimizer'].flat_p)
    assert torch.equal(a['optimizer'].flat_m, b['optimizer'].flat_m) and torch.equal(a['optimizer'].flat_v, b['optimizer'].flat_v)
    assert all(torch.equal(x, y) for x, y in zip(a['ema'].shadow_params, b['ema'].shadow_params))
    # the EMA update after a load runs on the re-allocated flat shadow buffer too
    assert b['ema']._flat.numel() == a['ema']._flat.numel()


def _aux_setup(B):
    from dposer_b200.body_model import BodyModel
    from dposer_b200.misc import Posenormalizer
    model = synthetic.make_score_model(42).cuda()
    model.train()
    bm = BodyModel(synthetic.make_body_tensors('smplx'), num_betas=10, batch_size=B, model_type='smplx').cuda()
    norm = Posenormalizer(None, device='cuda', normalize=True, min_max=False, rot_rep='axis')
    return model, bm, norm


def test_auxiliary_loss_step_vs_reference_golden():
    """get_step_fn(auxiliary_loss=True) (losses.py:91-121,244-258: DDIM chain under the optimiser + body model) against the
    REAL reference (train_aux_golden.npz: its step_fn run with a BodyModel-compatible object over the LBS restatement):
    the four losses, the gradient norm and sampled gradients, with the reference's t, z and 15 dropout masks replayed."""
    g = golden('train_aux_golden.npz')
    B, N = g['data'].shape[0], int(g['nsteps'])
    cfg = synthetic.default_config()
    model, bm, norm = _aux_setup(B)
    state = _state(cfg, model)
    state['step'] = 4000
    sde = sde_lib.subVPSDE(0.1, 20., 1000)
    step_fn = losses.get_step_fn(sde, train=True, optimize_fn=lambda *a, **k: None, reduce_mean=True, continuous=True,
                                 auxiliary_loss=True, denormalize=norm.offline_denormalize, body_model=bm, rot_rep='axis',
                                 denoise_steps=N)
    masks = torch.tensor(np.unpackbits(g['masks'], axis=-1)).reshape(N, 5, B, 1024)
    ld = step_fn(state, torch.tensor(g['data']).cuda(), t=torch.tensor(g['t']), z=torch.tensor(g['z']), drop_mask=masks)
    for k, tol in (('step_loss', 3e-4), ('score_loss', 3e-4), ('v2v_loss', 2e-3), ('j2j_loss', 2e-3)):
        assert abs(float(ld[k]) - float(g[k])) < tol * abs(float(g[k])), (k, float(ld[k]), float(g[k]))
    assert abs(state['optimizer'].grad_norm() - float(g['gnorm'])) < 3e-4 * float(g['gnorm'])
    for n, p in model.named_parameters():
        if f'g_{n}' not in g.files:
            continue
        flat = p.grad.reshape(-1)
        got = flat[sample_idx(flat.numel()).cuda()].cpu()
        assert float((got - torch.tensor(g[f'g_{n}'])).abs().max()) < 5e-4 * float(g[f'gmax_{n}']), n


def test_auxiliary_body_terms_vs_oracle():
    """The body-model terms alone (score term switched off, so that the chain's backward pass is what is measured): native
    gradients against autograd through the CPU oracle (DDIM chain + LBS restatement)."""
    from oracle import lbs_ref
    from oracle import score_ref as S
    from oracle import train_ref as T
    B, N = 5, 3
    model, bm, norm = _aux_setup(B)
    sde, osde = sde_lib.subVPSDE(0.1, 20., 1000), S.SubVP(0.1, 20., 1000)
    gen = torch.Generator().manual_seed(12)
    batch = norm.offline_normalize(synthetic.toy_poses()[:B].cuda()).cpu()
    t = torch.rand(B, generator=gen) * 0.6 + 0.2
    z = torch.randn(B, 63, generator=gen)
    masks = (torch.rand(N, 5, B, 1024, generator=gen) >= 0.1).to(torch.uint8)
    ld = losses.auxiliary_loss_grad(model, sde, batch.cuda(), norm.offline_denormalize, bm, denoise_steps=N, reduce_mean=True,
                                    t=t, z=z, drop_mask=masks, score_scale=0.0)
    m = synthetic.make_body_tensors('smplx')
    mean, std = norm.mean_poses.cpu(), norm.std_poses.cpu()

    def body_fn(pose):
        full = torch.cat([torch.zeros(B, 3), pose, torch.zeros(B, 99)], 1)
        return lbs_ref.body_forward(m, torch.zeros(B, m['shapedirs'].shape[2]), full)
    sd = S.make_state_dict(42)
    names = T.param_names(sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    _, _, ov, oj = T.aux_loss(full, osde, batch, t, z, masks, 0.1, lambda v: v * std + mean, body_fn, N, reduce_mean=True)
    used = [k for k in names if not k.startswith('pre_dense_cond')]
    og = dict(zip(used, torch.autograd.grad(ov + oj, [leaves[k] for k in used])))
    ov, oj = ov.detach(), oj.detach()
    assert abs(float(ld['v2v_loss']) - float(ov)) < 2e-3 * float(ov) and abs(float(ld['j2j_loss']) - float(oj)) < 2e-3 * float(oj)
    for n, p in model.named_parameters():
        if n in og:
            assert rel(p.grad, og[n]) < 2e-3, n
