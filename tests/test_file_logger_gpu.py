"""GPU: ``loggin_to_file`` (s4g_release_b200/file_logger.py) — the files of the reference's per-step dump
(utils/file_logger_cls.py:12-246) with their formats, the numbers against the reference's formulas restated in numpy,
and the collision-checked top frames against the CPU oracle's per-pose check."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _preds(n, seed):
    g = torch.Generator().manual_seed(seed)
    return {"score": torch.randn(1, 3, n, generator=g).cuda() * 2, "frame_R": torch.randn(1, 9, n, generator=g).cuda(),
            "frame_t": torch.randn(1, 4, n, generator=g).cuda(), "movable_logits": torch.rand(1, 5, n, generator=g).cuda()}


def test_files_and_numbers(tmp_path):
    from s4g_release_b200.file_logger import jet_colors, loggin_to_file
    from tests.inputs import tabletop_scene
    n = 4096
    pts = torch.from_numpy(tabletop_scene(1003, n))[None].cuda()
    preds = _preds(n, 1)
    labels = {"scene_points": pts, "scene_score": torch.rand(1, n).cuda(), "scene_score_labels": torch.randint(0, 3, (1, n)).cuda()}
    assert loggin_to_file(labels, preds, 7, str(tmp_path), prefix="val", with_label=True) is None
    d = tmp_path / "val_step00007"
    names = {p.name for p in d.iterdir()}
    assert names == {"scene_points.xyz", "gt_scene_score.txt", "gt_scene_score_labels.txt", "scene_score_logits.txt",
                     "pred_frame_R.txt", "pred_frame_t.txt", "pred_frame.ply", "pred_pts.ply", "pred_scene_score.txt"}
    xyz = pts[0].cpu().numpy().T
    np.testing.assert_allclose(np.loadtxt(d / "scene_points.xyz"), xyz, atol=5.1e-5)
    prob = torch.softmax(preds["score"][0], 0).cpu().numpy().T
    np.testing.assert_allclose(np.loadtxt(d / "scene_score_logits.txt"), prob, atol=5.1e-5)
    R = preds["frame_R"][0].cpu().numpy().T.reshape(-1, 3, 3)
    tp = torch.softmax(preds["frame_t"][0], 0).cpu().numpy().T
    want_t = -(tp * np.array([[0.08, 0.06, 0.04, 0.02]])).sum(1, keepdims=True) * R[:, :, 0] + xyz  # reference :41-46
    np.testing.assert_allclose(np.loadtxt(d / "pred_frame_t.txt"), want_t, atol=5.1e-5)
    score = (np.linspace(0, 1, 4)[:-1][None] * prob).sum(1)                                         # reference :67-69
    np.testing.assert_allclose(np.loadtxt(d / "pred_scene_score.txt"), score, atol=5.1e-5)
    head = (d / "pred_frame.ply").read_text().splitlines()[:12]
    assert head[0] == "ply" and "element vertex %d" % (12 * (n // 2)) in head and "element face %d" % (n // 2) in head
    head = (d / "pred_pts.ply").read_text().splitlines()
    assert "element vertex %d" % n in head[:10]
    # jet: the published end points / mid colour
    c = jet_colors(np.array([0.0, 0.5, 1.0]))
    assert np.allclose(c[0], [0, 0, 0.5], atol=2e-3) and np.allclose(c[2], [0.5, 0, 0], atol=2e-3) and c[1][1] > 0.99


def test_top_frames_against_the_oracle_collision_check(tmp_path):
    from oracle import model_cpu
    from s4g_release_b200.file_logger import loggin_to_file
    from tests.inputs import tabletop_scene
    n = 8192
    cloud = tabletop_scene(1004, n)
    pts = torch.from_numpy(cloud)[None].cuda()
    preds = _preds(n, 2)
    top_H, score = loggin_to_file({"scene_points": pts}, preds, 1, str(tmp_path), with_label=False, work_dir=str(tmp_path))
    # restate :186-232 on the host: top 50 by the logged score, Gram-Schmidt, per-pose oracle check
    prob = torch.softmax(preds["score"][0], 0).cpu().numpy().T
    pred = (np.linspace(0, 1, 4)[:-1][None] * prob).sum(1)
    top = np.argsort(-pred)[:50]
    R = preds["frame_R"][0].cpu().numpy().T.reshape(-1, 3, 3).astype(np.float64)
    tp = torch.softmax(preds["frame_t"][0], 0).cpu().numpy().T
    t = -(tp * np.array([[0.08, 0.06, 0.04, 0.02]])).sum(1, keepdims=True) * R[:, :, 0] + cloud.T
    H = model_cpu.orthogonalization(R[top], t[top])
    keep, _ = model_cpu.collision_filter(H, cloud.T)
    assert len(top_H) == len(keep) and len(score) == len(keep)
    if len(keep):
        np.testing.assert_allclose(top_H, H[keep], atol=1e-5)  # the logger decodes in fp32 like the reference (:37-46)
        assert os.path.exists(tmp_path / "top_frames.npy")
    assert (tmp_path / "postprocess_time_ours.txt").exists()
