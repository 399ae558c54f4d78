"""CPU: the oracle must reproduce every golden vector produced by the REAL reference
(tests/golden/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import numpy as np
import torch

from conftest import golden
from oracle import fitting_ref as Fr
from oracle import score_ref as S


def test_known_answers_appendix_d():
    k = golden('known_answers.npz')
    sde = S.SubVP(0.1, 20., 1000)
    ts = torch.linspace(1, 1e-3, 1000)
    t5 = ts[torch.tensor(k['i'])]
    assert np.array_equal(t5.numpy(), k['t'])
    assert np.array_equal((t5 * 999).long().numpy(), k['idx'])
    assert np.array_equal(sde.sde(torch.zeros(5, 1), t5)[1].numpy(), k['g'])
    assert np.array_equal(sde.alpha_sigma(t5)[1].numpy(), k['std'])
    assert np.array_equal(S.sigma_table()[k['idx']].numpy(), k['sigmas'])
    assert np.array_equal(S.timestep_embedding(torch.tensor([499.5]))[0].numpy(), k['temb_4995'])
    assert np.array_equal(sde.alphas.numpy(), k['sde_alphas'])
    assert np.array_equal((ts * 999).long().numpy(), k['all_idx'])
    # SURVEY Appendix D printed values
    assert k['idx'].tolist() == [999, 500, 499, 1, 0]
    np.testing.assert_allclose(k['g'][0], 4.472136021, rtol=1e-7)
    np.testing.assert_allclose(k['std'][4], 1.099705696e-04, rtol=1e-6)


def test_weight_recipe_fingerprint(oracle_sd):
    fp = golden('weights_fingerprint.npz')
    for k in fp.files:
        assert oracle_sd[k].double().abs().sum().item() == float(fp[k]), k


def test_score_net_matches_reference(oracle_sd):
    g = golden('score_golden.npz')
    x = torch.tensor(g['x'])
    sde = S.SubVP()
    for tv in ['1.0', '0.5', '0.1', '0.01', '0.001']:
        vt = torch.ones(7) * float(tv)
        assert np.array_equal(S.score_fn(oracle_sd, sde, x, vt).numpy(), g[f'score_{tv}'])
        assert np.array_equal(S.score_model_forward(oracle_sd, x, vt * 999).numpy(), g[f'model_{tv}'])
    vt = torch.tensor(g['t_mixed'])
    assert np.array_equal(S.score_model_forward(oracle_sd, x, vt * 999).numpy(), g['model_mixed'])


def _noise_dict(g, name, start):
    n = g[f'{name}_noise_pred'].shape[0]
    out = {}
    for i in range(n):
        d = {'pred': torch.tensor(g[f'{name}_noise_pred'][i])}
        for k in ('corr', 'imp_c', 'imp_p'):
            if f'{name}_noise_{k}' in g.files:
                d[k] = torch.tensor(g[f'{name}_noise_{k}'][i])
        out[start + i] = d
    return out


def test_pc_sampler_matches_reference(oracle_sd):
    g = golden('sampler_golden.npz')
    cases = [('em8', 8, 'none', None, False, 0), ('em32', 32, 'none', None, False, 0),
             ('lang8', 1000, 'langevin', 'denoise', False, 992), ('comp8', 8, 'none', 'completion', False, 0),
             ('ode8', 8, 'none', None, True, 0), ('den16', 16, 'none', 'denoise', False, 10)]
    for name, N, corr, task, pf, start in cases:
        obs = torch.tensor(g[f'{name}_obs']) if f'{name}_obs' in g.files else None
        mask = torch.tensor(g[f'{name}_mask']) if f'{name}_mask' in g.files else None
        traj, out = S.pc_sample(oracle_sd, S.SubVP(0.1, 20., N), torch.tensor(g[f'{name}_z0']), 1e-3,
                                noise=_noise_dict(g, name, start), corrector=corr, observation=obs, mask=mask,
                                task=task, start_step=start, probability_flow=pf, keep_traj=True)
        assert np.array_equal(out.numpy(), g[f'{name}_out']), name
        assert np.array_equal(traj[-1].numpy(), g[f'{name}_traj_last']), name


def test_prior_loss_matches_reference(oracle_sd):
    g = golden('prior_golden.npz')
    x0 = torch.tensor(g['x0'])
    sde = S.SubVP()
    ts = torch.linspace(1, 1e-3, 1000)
    for name, qt in [('q799', 799), ('q998', 998), ('q400', 400)]:
        l, gr = S.prior_loss(oracle_sd, sde, x0, torch.ones(7) * ts[qt], torch.tensor(g[f'{name}_z']), True, 'mean')
        assert np.array_equal(l.numpy(), g[f'{name}_loss']) and np.array_equal(gr.numpy(), g[f'{name}_grad'])
    for name, weighted, multi in [('md_plain', False, False), ('md_weighted', True, False), ('md_ddim', False, True)]:
        l, gr = S.prior_loss(oracle_sd, sde, x0, torch.ones(7) * ts[450], torch.tensor(g[f'{name}_z']), weighted,
                             'sum', divisor=7, multi_denoise=multi)
        assert np.array_equal(l.numpy(), g[f'{name}_loss']) and np.array_equal(gr.numpy(), g[f'{name}_grad'])


def test_integer_tables_bit_exact():
    g = golden('int_tables.npz')
    for part in ['legs', 'arms', 'trunk', 'hands', 'left_leg', 'right_leg', 'left_arm', 'right_arm']:
        assert Fr.BODY_PARTS[part] == g[f'part_{part}'].tolist()
        assert np.array_equal(np.sort(Fr.mask_indices(part).numpy()), g[f'maskzero_{part}'])
    assert g['maskzero_legs'].tolist() == [0, 1, 2, 3, 4, 5, 9, 10, 11, 12, 13, 14, 18, 19, 20, 21, 22, 23, 27, 28,
                                           29, 30, 31, 32]
    for N, total, trun, off, nm in [(1000, 200, 5.0, 2, 'completion'), (500, 180, 4.0, 2, 'denoise'),
                                    (500, 500, 20.0, 5, 'smplify')]:
        assert S.quan_t_schedule(N, total, trun, off) == g[f'quan_t_{nm}'].tolist()
    for total, world, r, first, n in g['shards'].tolist():
        st, cnt = Fr.shard_range(total, world, r)
        assert cnt == n and (n == 0 or st == first)


def test_fitting_losses_and_metrics_match_reference():
    g = golden('fitting_golden.npz')
    t = lambda k: torch.tensor(g[k])         # noqa: E731
    ints = golden('int_tables.npz')
    center = torch.full((6, 2), 512.)
    body = Fr.body_fitting_loss(t('pose'), t('betas'), t('joints'), center, t('kp'), t('conf'), t('prior'))
    cam = Fr.camera_fitting_loss(t('joints'), t('cam_t'), t('cam_est'), center, t('kp'), t('conf'),
                                 op_ind=ints['op_ind'], gt_ind=ints['gt_ind'])
    np.testing.assert_allclose(body.numpy(), g['body_loss'], rtol=1e-6)
    np.testing.assert_allclose(cam.numpy(), g['cam_loss'], rtol=1e-6)
    np.testing.assert_allclose(Fr.gaussian_smoothing(t('smooth_in'), 3, 2).numpy(), g['smooth_out'], rtol=1e-6,
                               atol=1e-7)
    a = golden('apd_golden.npz')
    np.testing.assert_allclose(Fr.apd(torch.tensor(a['joints'])).numpy(), a['apd'], rtol=1e-6)


def test_lbs_oracle_invariants():
    """LBS parity is unpinned (smplx absent); check the invariants any correct LBS satisfies."""
    from dposer_b200 import synthetic
    from oracle import lbs_ref
    m = synthetic.make_body_tensors('smpl')
    B = 3
    g = torch.Generator().manual_seed(0)
    betas = torch.randn(B, 10, generator=g)
    # zero pose: vertices equal the shaped template, joints equal the regressed rest joints
    v, j = lbs_ref.body_forward(m, betas, torch.zeros(B, 72))
    v_shaped = m['v_template'][None] + torch.einsum('bl,mkl->bmk', betas, m['shapedirs'])
    assert (v - v_shaped).abs().max() < 1e-6
    jrest = torch.einsum('bik,ji->bjk', v_shaped, m['J_regressor'])
    assert (j[:, :24] - jrest).abs().max() < 1e-6
    # root-only rotation: rigid motion about the pelvis rest joint
    pose = torch.zeros(B, 72)
    pose[:, :3] = torch.randn(B, 3, generator=g)
    v2, _ = lbs_ref.body_forward(m, betas, pose)
    R = lbs_ref.batch_rodrigues(pose[:, :3])
    expect = torch.einsum('bij,bvj->bvi', R, v_shaped - jrest[:, :1]) + jrest[:, :1]
    assert (v2 - expect).abs().max() < 2e-6
    assert j.shape == (B, 45, 3)
