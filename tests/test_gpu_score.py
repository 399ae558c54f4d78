"""GPU parity: score network / score_fn through the C ABI against the oracle and the golden vectors.
Tolerances: fp32 engine 2e-5 relative (fp32 accumulation order differs from ATen);
tcgen05 engine (fp16 operands, fp32 accumulate) 1e-3 relative -- BASELINE.json's stated bound."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err, max_rel
from dposer_b200 import _lib as L
from dposer_b200 import sde_lib, utils as mutils

pytestmark = pytest.mark.gpu
TOL = {L.ENGINE_FP32: 2e-5, L.ENGINE_TC: 1e-3}


@pytest.mark.parametrize('engine', [L.ENGINE_FP32, L.ENGINE_TC])
def test_score_fn_and_model_vs_golden(gpu_model, engine):
    g = golden('score_golden.npz')
    x = torch.tensor(g['x']).cuda()
    sde = sde_lib.subVPSDE(0.1, 20., 1000)
    gpu_model.engine = engine
    try:
        score_fn = mutils.get_score_fn(sde, gpu_model, train=False, continuous=True)
        for tv in ['1.0', '0.5', '0.1', '0.01', '0.001']:
            vt = torch.ones(7, device='cuda') * float(tv)
            assert max_rel(score_fn(x, vt, None, None), g[f'score_{tv}']) < TOL[engine], tv
            assert max_rel(gpu_model(x, vt * 999), g[f'model_{tv}']) < TOL[engine], tv
    finally:
        gpu_model.engine = L.ENGINE_AUTO


def test_model_forward_per_row_t(gpu_model):
    g = golden('score_golden.npz')
    x = torch.tensor(g['x']).cuda()
    vt = torch.tensor(g['t_mixed']).cuda()
    out = gpu_model(x, vt * 999)
    # rows have different magnitudes (sigma table): compare row by row
    ref = torch.tensor(g['model_mixed'])
    for r in range(7):
        assert max_rel(out[r], ref[r]) < 2e-5, r


@pytest.mark.parametrize('engine', [L.ENGINE_FP32, L.ENGINE_TC])
@pytest.mark.parametrize('B', [1, 7, 128, 500, 4096])
def test_score_vs_oracle_batches(gpu_model, oracle_sd, engine, B):
    from oracle import score_ref as S
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, 63, generator=gen) * 1.5
    sde, osde = sde_lib.subVPSDE(0.1, 20., 1000), S.SubVP()
    gpu_model.engine = engine
    try:
        score_fn = mutils.get_score_fn(sde, gpu_model, train=False, continuous=True)
        for tv in ([1.0, 0.5, 0.1, 0.01, 1e-3] if B <= 500 else [0.5]):
            ref = S.score_fn(oracle_sd, osde, x, torch.ones(B) * tv)
            out = score_fn(x.cuda(), torch.ones(B, device='cuda') * tv, None, None)
            assert rel_err(out, ref) < TOL[engine], (B, tv)
            # every row individually within 3x the bound (norm-wise)
            row = ((out.cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
            assert row < 3 * TOL[engine], (B, tv, row)
    finally:
        gpu_model.engine = L.ENGINE_AUTO


def test_tc_engine_agrees_with_fp32_engine_on_device(gpu_model):
    gen = torch.Generator().manual_seed(99)
    x = torch.randn(1000, 63, generator=gen).cuda()
    lab = torch.ones(1000, device='cuda') * 499.5
    gpu_model.engine = L.ENGINE_FP32
    a = gpu_model(x, lab)
    gpu_model.engine = L.ENGINE_TC
    b = gpu_model(x, lab)
    gpu_model.engine = L.ENGINE_AUTO
    assert rel_err(b, a) < 1e-3


@pytest.mark.parametrize('B', [127, 129, 18944, 18945, 37889])
def test_tc_engine_tile_and_wave_boundaries(gpu_model, B):
    """Row counts around the 128-row tile and the one-tile-per-SM wave (148 x 128 = 18 944): ragged last tile,
    ghost tile of an odd pair, more tile pairs than CTA pairs.  Every row must match the exact fp32 engine."""
    gen = torch.Generator().manual_seed(B)
    x = (torch.randn(B, 63, generator=gen) * 1.3).cuda()
    lab = torch.full((B,), 250.25, device='cuda')
    try:
        gpu_model.engine = L.ENGINE_FP32
        a = gpu_model(x, lab)
        gpu_model.engine = L.ENGINE_TC
        b = gpu_model(x, lab)
    finally:
        gpu_model.engine = L.ENGINE_AUTO
    row = ((b - a).norm(dim=1) / a.norm(dim=1)).max()
    assert float(row) < 1e-3, float(row)    # fp16 operands / fp32 accumulate: 1e-3 relative per row
    assert torch.isfinite(b).all()


def test_time_table_against_oracle(gpu_model, oracle_sd):
    """The hoisted time path: W_lt temb + b_lt + b_l for a few labels."""
    import torch.nn.functional as F
    from oracle import score_ref as S
    labels = torch.tensor([999.0, 499.5, 250.25, 0.999])
    tab = gpu_model.time_table(labels.cuda()).cpu()
    temb = F.silu(F.linear(S.timestep_embedding(labels), oracle_sd['shared_time_embed.0.weight'],
                           oracle_sd['shared_time_embed.0.bias']))
    names = ['pre_dense', 'b1_dense1', 'b1_dense2', 'b2_dense1', 'b2_dense2']
    for l, n in enumerate(names):
        ref = F.linear(temb, oracle_sd[n + '_t.weight'], oracle_sd[n + '_t.bias']) + oracle_sd[n + '.bias']
        assert max_rel(tab[:, l], ref) < 5e-6, n


def test_empty_batch_and_bad_args(gpu_model):
    out = gpu_model(torch.zeros(0, 63, device='cuda'), torch.zeros(0, device='cuda'))
    assert out.shape == (0, 63)
    with pytest.raises(RuntimeError):
        gpu_model(torch.zeros(2, 63), torch.ones(2))          # CPU tensor: no fallback


@pytest.mark.parametrize('engine', [L.ENGINE_FP32, L.ENGINE_TC])
def test_score_fn_vpsde_vesde_vs_golden(gpu_model, engine):
    """The other two SDEs of sde_lib.py (VPSDE :122-181, VESDE :234-291) through get_score_fn (utils.py:127-186),
    against the real reference (tests/golden/make_golden_sde_variants.py)."""
    g = golden('sde_variants_golden.npz')
    x = torch.tensor(g['x']).cuda()
    gpu_model.engine = engine
    try:
        for name, sde in [('vp', sde_lib.VPSDE(0.1, 20., 1000)), ('ve', sde_lib.VESDE(0.01, 50., 1000))]:
            score_fn = mutils.get_score_fn(sde, gpu_model, train=False, continuous=True)
            for tv in ['1.0', '0.5', '0.01']:
                vt = torch.ones(7, device='cuda') * float(tv)
                assert max_rel(score_fn(x, vt, None, None), g[f'{name}_score_{tv}']) < TOL[engine], (name, tv)
    finally:
        gpu_model.engine = L.ENGINE_AUTO
