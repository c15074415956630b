"""GPU: K6 point-cloud metrics (nof/criteria/pointcloud_metrics.py, metrics.py) against the oracle's exact kd-tree."""
import numpy as np
import pytest
import torch

import pcnerf_oracle as orc
from gpu_util import dev

pytestmark = pytest.mark.gpu


def _clouds(n1, n2, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-20, 20, size=(n1, 3)) * np.array([1, 1, 0.05])
    b = a[rng.integers(0, n1, size=n2)] + rng.normal(0, 0.15, size=(n2, 3))
    return a, b


@pytest.mark.parametrize("n1,n2", [(1, 1), (7, 1030), (5000, 7000), (40000, 1000)])
def test_nn_correspondance_exact(n1, n2):
    from pcnerf_b200.nof.criteria import pointcloud_metrics as pm
    a, b = _clouds(n1, n2, n1 + n2)
    idx, dist = pm.nn_correspondance(a, b)
    ridx, rdist = orc.nn_correspondance(a, b)
    np.testing.assert_allclose(np.asarray(dist), rdist, rtol=1e-12, atol=1e-15)
    same = np.asarray(idx) == ridx
    # a different index is only acceptable for an exact tie
    if not same.all():
        d_alt = np.linalg.norm(a[np.asarray(idx)[~same]] - b[~same], axis=1)
        np.testing.assert_allclose(d_alt, rdist[~same], rtol=1e-12)


def test_eval_pts_and_scalar_metrics():
    from pcnerf_b200.nof.criteria import metrics as m
    a, b = _clouds(20000, 18000, 3)
    cd, f = m.eval_points(torch.from_numpy(a).to(dev()), torch.from_numpy(b).to(dev()))
    rcd, rf = orc.eval_pts(a, b)
    np.testing.assert_allclose([cd, f], [rcd, rf], rtol=1e-12)
    pred = torch.tensor([1.0, 2.0, 3.3, 10.0], device=dev())
    gt = torch.tensor([1.1, 2.0, 3.0, 9.5], device=dev())
    np.testing.assert_allclose(float(m.abs_error(pred, gt)), (0.1 + 0 + 0.3 + 0.5) / 4, rtol=1e-6)
    assert float(m.acc_thres(pred, gt)) == 50.0
    mask = torch.tensor([True, True, False, False], device=dev())
    assert float(m.acc_thres(pred, gt, mask)) == 100.0


def test_empty_clouds():
    from pcnerf_b200.nof.criteria import pointcloud_metrics as pm
    assert pm.nn_correspondance(np.zeros((0, 3)), np.zeros((5, 3))) == ([], [])
    assert pm.nn_correspondance(np.zeros((5, 3)), np.zeros((0, 3))) == ([], [])
