"""Synthetic bundle-adjustment problems for the Jacobian tests: cameras looking along a
panning arc, point matches generated through the true homographies plus noise."""
import numpy as np


def rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def ba_case(n_cam, matches_per_pair, seed, w=1500, h=1112, extra_pairs=0):
    """-> cams [n_cam, 12] (focal, ppx, ppy, R row-major), pairs [(from, to, n_match)], pts [n, 4] (to.xy, from.xy).
    Every camera appears in at least one pair (a chain), plus `extra_pairs` random ones, some reversed."""
    rng = np.random.RandomState(seed)
    cams = np.zeros((n_cam, 12))
    Ks, Rs = [], []
    for i in range(n_cam):
        f = 900.0 + 200.0 * rng.rand()
        ppx, ppy = rng.randn(2) * 3.0
        R = rodrigues(np.array([0.02 * rng.randn(), 0.25 * (i - n_cam / 2) + 0.02 * rng.randn(), 0.02 * rng.randn()]))
        if i == 0:
            R = np.eye(3)                       # the identity camera: dRdvi's small-angle branch
        cams[i, :3] = f, ppx, ppy
        cams[i, 3:] = R.reshape(-1)
        Ks.append(np.array([[f, 0, ppx], [0, f, ppy], [0, 0, 1.0]])); Rs.append(R)
    plist = [(i, i + 1) for i in range(n_cam - 1)]
    for _ in range(extra_pairs):
        a, b = rng.choice(n_cam, 2, replace=False)
        plist.append((int(a), int(b)))
    pairs, pts = [], []
    for k, (a, b) in enumerate(plist):
        if k % 3 == 2:
            a, b = b, a
        nm = int(matches_per_pair * (0.5 + rng.rand())) if matches_per_pair > 1 else matches_per_pair
        H = Ks[a] @ Rs[a] @ Rs[b].T @ np.linalg.inv(Ks[b])      # to -> from
        to = np.stack([rng.uniform(-w / 2, w / 2, nm), rng.uniform(-h / 2, h / 2, nm)], 1)
        hom = (H @ np.concatenate([to, np.ones((nm, 1))], 1).T).T
        frm = hom[:, :2] / hom[:, 2:3] + rng.randn(nm, 2) * 0.7
        pairs.append((a, b, nm))
        pts.append(np.concatenate([to, frm], 1))
    return cams, pairs, np.concatenate(pts, 0) if pts else np.zeros((0, 4))


def numpy_pair_mats(cams, pairs):
    """The 13 per-pair matrices of pano_ba_pair from numpy products — NOT bit-identical to the
    reference's Eigen products, which does not matter where the matrices are only INPUTS (the
    GPU-vs-oracle tests); the golden fixture carries the reference's own."""
    cams = np.asarray(cams, np.float64).reshape(-1, 12)

    def K(c):
        return np.array([[c[0], 0, c[1]], [0, c[0], c[2]], [0, 0, 1.0]])

    def cross(v):
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])

    def drdvi(R):
        # a compact formula for the derivative of a rotation in exponential coordinates (Gallego & Yezzi)
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
        s = np.linalg.norm(w)
        if s < 1e-7:
            return [cross(e) for e in np.eye(3)]
        v = w / s * np.arccos(np.clip((np.trace(R) - 1) * 0.5, -1, 1))
        out = []
        for i in range(3):
            e = np.eye(3)[i]
            out.append((v[i] * cross(v) + cross(np.cross(v, (np.eye(3) - R) @ e))) / (v @ v) @ R)
        return out

    dK = [np.array([[1, 0, 0], [0, 1, 0], [0, 0, 0.0]]), np.array([[0, 0, 1], [0, 0, 0], [0, 0, 0.0]]),
          np.array([[0, 0, 0], [0, 0, 1], [0, 0, 0.0]])]
    out = np.zeros((len(pairs), 13, 9))
    for p, (a, b, _) in enumerate(pairs):
        Ka, Ra = K(cams[a]), cams[a, 3:].reshape(3, 3)
        Kbi, Rbi = np.linalg.inv(K(cams[b])), cams[b, 3:].reshape(3, 3).T
        da, db = drdvi(Ra), [m.T for m in drdvi(cams[b, 3:].reshape(3, 3))]
        m = Ka @ Ra @ Rbi @ Kbi
        mats = [m, Ra @ Rbi @ Kbi, Rbi @ Kbi] + [Ka @ d for d in da] + [Kbi] + [m @ d for d in dK] + [Ka @ Ra @ d for d in db]
        out[p] = np.array([x.reshape(-1) for x in mats])
    return out
