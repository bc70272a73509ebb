"""Seeded input generators shared by the CPU and GPU tests."""
import numpy as np
import torch

# reduced PN2_CLS-shaped configuration of the tiny golden fixture (tests/golden/pn2cls_tiny.npz)
TINY_CONFIG = dict(
    score_classes=3,
    num_centroids=(256, 64, 16),
    radius=(0.1, 0.2, 0.4),
    num_neighbours=(16, 16, 8),
    sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128)),
    fp_channels=((128, 128), (64, 64), (32, 32, 32)),
    num_fp_neighbours=(3, 3, 3),
    seg_channels=(64, 32, 32, 16),
    num_removal_directions=5,
    dropout_prob=0.5,
)


def uniform_cloud(B, N, seed, scale=1.0):
    rs = np.random.RandomState(seed)
    return torch.from_numpy((rs.rand(B, 3, N) * scale).astype(np.float32))


def lattice_cloud(B, N, seed, side=8):
    """Points on an integer lattice / side: massive exact distance ties (and duplicates when N > side^3)."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy((rs.randint(0, side, size=(B, 3, N)) / float(side)).astype(np.float32))


def duplicated_cloud(B, N, seed):
    """50 % duplicated points, as np.random.choice(replace=True) produces (grasp_detector.py:89)."""
    rs = np.random.RandomState(seed)
    base = rs.rand(B, 3, max(N // 2, 1)).astype(np.float32)
    sel = rs.randint(0, base.shape[2], size=N)
    return torch.from_numpy(np.ascontiguousarray(base[:, :, sel]))


def identical_cloud(B, N):
    return torch.full((B, 3, N), 0.25, dtype=torch.float32)


def tabletop_scene(seed, n_points=25600):
    """Procedural single-view tabletop cloud in the fixture's camera frame:
    a 0.8 x 0.7 m plane patch plus 5-12 random boxes / cylinders / spheres (5-20 cm), visible-side
    surface samples, 1 mm depth noise, 5 mm voxel de-duplication, random choice to n_points."""
    rs = np.random.RandomState(seed)
    pts = []
    n_plane = 60000
    plane = np.stack([rs.uniform(-0.4, 0.4, n_plane), rs.uniform(-0.35, 0.35, n_plane), np.zeros(n_plane)], 1)
    pts.append(plane)
    for _ in range(rs.randint(5, 13)):
        kind = rs.randint(3)
        c = np.array([rs.uniform(-0.3, 0.3), rs.uniform(-0.25, 0.25), 0.0])
        s = rs.uniform(0.05, 0.20, size=3)
        n = 8000
        if kind == 0:  # box: top + 2 visible sides
            u = rs.uniform(-0.5, 0.5, size=(n, 2))
            face = rs.randint(3, size=n)
            p = np.zeros((n, 3))
            p[face == 0] = np.c_[u[face == 0] * s[:2], np.full((face == 0).sum(), s[2])]
            p[face == 1] = np.c_[np.full((face == 1).sum(), -0.5 * s[0]), u[face == 1, 0] * s[1],
                                 (u[face == 1, 1] + 0.5) * s[2]]
            p[face == 2] = np.c_[u[face == 2, 0] * s[0], np.full((face == 2).sum(), -0.5 * s[1]),
                                 (u[face == 2, 1] + 0.5) * s[2]]
        elif kind == 1:  # upright cylinder: top disc + front half of the wall
            r, h = 0.5 * s[0], s[2]
            top = rs.rand(n) < 0.3
            ang = rs.uniform(np.pi, 2 * np.pi, n)
            rad = np.where(top, r * np.sqrt(rs.rand(n)), r)
            ang = np.where(top, rs.uniform(0, 2 * np.pi, n), ang)
            p = np.c_[rad * np.cos(ang), rad * np.sin(ang), np.where(top, h, rs.uniform(0, h, n))]
        else:  # sphere: upper-front part
            r = 0.5 * s[0]
            v = rs.randn(n, 3)
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            v[:, 2] = np.abs(v[:, 2])
            p = v * r + np.array([0, 0, r])
        pts.append(p + c)
    p = np.concatenate(pts, 0)
    p[:, 2] += rs.randn(p.shape[0]) * 0.001
    # 5 mm voxel de-duplication (processing_config.py:20)
    key = np.floor(p / 0.005).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    p = p[np.sort(first)]
    # camera frame of the fixture: mean about (-0.06, -0.02, -1.18)
    p = p[:, [1, 0, 2]] * np.array([1.0, 1.0, -1.0]) + np.array([-0.057, -0.017, -1.18])
    replace = p.shape[0] < n_points
    sel = rs.choice(p.shape[0], n_points, replace=replace)
    return np.ascontiguousarray(p[sel].T.astype(np.float32))


def tabletop_batch(B, first_seed=1000, n_points=25600):
    return torch.from_numpy(np.stack([tabletop_scene(first_seed + i, n_points) for i in range(B)]))
