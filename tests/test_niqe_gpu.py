"""NIQE on the device (csrc/niqe.cu + resr_b200/iqa.py) against the unmodified reference run with synthetic
pristine statistics (tests/golden/niqe.npz, oracle/make_golden_niqe.py) and against the numpy oracle on another image."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rows(a):
    return np.array(sorted(map(tuple, a)))


def test_niqe_matches_reference_golden():
    import resr_b200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "niqe.npz"))
    mu, cov = g["mu_prisparam"], g["cov_prisparam"]
    for ci in range(3):
        x = torch.from_numpy(g[f"x{ci}"].astype(np.float32) / np.float32(255.0)).cuda()
        border = int(g[f"border{ci}"])
        feat = resr_b200.iqa.niqe_features(x, border).cpu().numpy()
        ref_feat = g[f"feat{ci}"]
        assert feat.shape == ref_feat.shape
        worst = max(np.abs(_rows(feat[b]) - _rows(ref_feat[b])).max() for b in range(feat.shape[0]))
        metric = resr_b200.iqa.NIQE(border, (mu, cov))(x).cpu().numpy().reshape(-1)
        rel = np.abs(metric - g[f"niqe{ci}"]).max() / g[f"niqe{ci}"].max()
        print(f"case {ci}: features max diff {worst:.2e}; NIQE {metric} vs {g[f'niqe{ci}']} (rel {rel:.1e})")
        assert worst <= 1e-6
        assert rel <= 1e-6


def test_niqe_matches_oracle_on_a_large_image_and_rejects_small_ones():
    import resr_b200
    from oracle import niqe as on
    rng = np.random.default_rng(3)
    x = (rng.integers(0, 256, (1, 3, 400, 520)).astype(np.float32) / np.float32(255.0))
    a = rng.normal(0, 0.3, (36, 36))
    mu, cov = rng.normal(0, 1, 36), a @ a.T + 0.05 * np.eye(36)
    score, feat = on.niqe(x, 4, mu, cov)
    got_feat = resr_b200.iqa.niqe_features(torch.from_numpy(x).cuda(), 4).cpu().numpy()
    assert np.abs(_rows(got_feat[0]) - _rows(feat[0])).max() <= 1e-6
    got = resr_b200.iqa.NIQE(4, (mu, cov))(torch.from_numpy(x).cuda()).item()
    assert abs(got - score[0]) <= 1e-6 * score[0]
    with pytest.raises(ValueError):
        resr_b200.iqa.niqe_features(torch.zeros(1, 3, 90, 200).cuda(), 0)
