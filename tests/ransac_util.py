"""Shared inputs of the RANSAC scoring tests: matched point pairs related by a true homography
plus outliers, and hypotheses spread around it (generated with a fixed MT19937, the seed
policy the engine asks of the host)."""
import numpy as np


def ransac_case(n_match, n_hyp, seed, w=1500, h=1112):
    rng = np.random.RandomState(seed)
    H = np.array([[1.0, 0.02, 480.0], [-0.015, 1.0, 6.0], [1e-5, -2e-5, 1.0]])
    kp2 = np.stack([rng.uniform(-w / 2, w / 2, n_match), rng.uniform(-h / 2, h / 2, n_match)], 1)
    p = np.concatenate([kp2, np.ones((n_match, 1))], 1) @ H.T
    kp1 = p[:, :2] / p[:, 2:3] + rng.randn(n_match, 2) * 1.2
    out = rng.rand(n_match) < 0.3
    kp1[out] = np.stack([rng.uniform(-w / 2, w / 2, out.sum()), rng.uniform(-h / 2, h / 2, out.sum())], 1)
    homos = np.repeat(H.reshape(1, 9), n_hyp, 0)
    homos = homos + rng.randn(n_hyp, 9) * np.array([1e-3, 1e-3, 1.5, 1e-3, 1e-3, 1.5, 1e-7, 1e-7, 0.0])
    if n_hyp > 3:
        homos[n_hyp // 2] = homos[1]          # a tie: the FIRST maximum must win
    thres = np.float32((w + h) * 0.5 / 800 * 3.5)      # transform_estimate.cc:47 with RANSAC_INLIER_THRES 3.5
    return kp1, kp2, homos, float(thres)
