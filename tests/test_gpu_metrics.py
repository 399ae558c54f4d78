"""GPU parity: APD and per-sample point errors vs the reference's numbers / the oracle."""
import numpy as np
import pytest
import torch

from conftest import golden
from dposer_b200 import metric
from oracle import fitting_ref as Fr

pytestmark = pytest.mark.gpu


def test_apd_vs_reference_golden():
    a = golden('apd_golden.npz')
    out = metric.average_pairwise_distance(torch.tensor(a['joints']).cuda())
    np.testing.assert_allclose(float(out), float(a['apd']), rtol=2e-6)
    j = torch.randn(500, 22, 3, generator=torch.Generator().manual_seed(0))
    np.testing.assert_allclose(float(metric.average_pairwise_distance(j.cuda())), float(Fr.apd(j)), rtol=1e-5)


def test_point_errors_vs_oracle():
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(9, 300, 3, generator=g), torch.randn(9, 300, 3, generator=g)
    idx = [3, 5, 8, 100, 299]
    mp, _ = Fr.eval_bodies(a, b, a, b, vert_idx=idx)
    np.testing.assert_allclose(metric.mean_point_error_mm(a.cuda(), b.cuda(), idx).cpu().numpy(), mp, rtol=1e-5)
    mp, _ = Fr.eval_bodies(a, b, a, b)
    np.testing.assert_allclose(metric.mean_point_error_mm(a.cuda(), b.cuda()).cpu().numpy(), mp, rtol=1e-5)
