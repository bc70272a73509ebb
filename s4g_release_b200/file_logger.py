"""``loggin_to_file`` — the reference's per-step dump of a PN2_CLS prediction (utils/file_logger_cls.py:12-246) without
open3d / matplotlib, and with the top-K collision check on the device.

Files written into ``<output_dir>/<prefix>_step<step:05d>/`` (same names, same ``numpy.savetxt`` formats as the reference):
  scene_points.xyz            (N, 3)   %.4f                                   :26-27
  gt_scene_score.txt, gt_scene_score_labels.txt   when ``with_label``         :28-32
  scene_score_logits.txt      (N, C)   softmax of the score head, %.4f        :33-35
  pred_frame_R.txt            (N, 9)   raw rotation head, %.4f                :37-39
  pred_frame_t.txt            (N, 3)   point - (softmax(frame_t) · [.08,.06,.04,.02]) * R[:, :, 0]   :41-48
  pred_scene_score.txt        (N,)     sum_c softmax_c * linspace(0, 1, C + 1)[:-1]  (the reference's "TODO" weights:
                                        they start at 0, unlike grasp_detector.py:144 which uses [1:])     :67-69, 171
  pred_pts.ply                the cloud coloured by that score through the "jet" colour map (1024 levels)   :71-72, 165-170
  pred_frame.ply              per second point a small triangle spanning the predicted y axis (12 vertices, 1 face)  :121-163
and, when ``with_label`` is False (real experiments, :186-244): the K = 50 best-scoring points are turned into gripper
frames (Gram-Schmidt on the first two columns), checked against the cloud with the gripper collision test — here ONE
launch of s4g_grasp_collision_f32 for all 50 instead of a python loop with two host syncs per pose — and the
collision-free ones are saved to ``top_frames.npy``; returns ``(top_H, score)``.

PLY files are written as ASCII PLY (the reference goes through open3d's writer; any PLY reader, open3d included, loads
both).  The colour map restates matplotlib's published "jet" segment table; matplotlib is not imported.
"""
import os
import time

import numpy as np
import torch

T_SCORE = np.array([0.08, 0.06, 0.04, 0.02])
_JET = {  # matplotlib _cm.py `_jet_data`: (x, y) break points, linear in between
    "r": ((0.0, 0.0), (0.35, 0.0), (0.66, 1.0), (0.89, 1.0), (1.0, 0.5)),
    "g": ((0.0, 0.0), (0.125, 0.0), (0.375, 1.0), (0.64, 1.0), (0.91, 0.0), (1.0, 0.0)),
    "b": ((0.0, 0.5), (0.11, 1.0), (0.34, 1.0), (0.65, 0.0), (1.0, 0.0)),
}


def jet_colors(values, levels=1024):
    """plt.get_cmap("jet", levels)(values)[:, :3] for values in [0, 1] (out-of-range values clamp to the end colours)."""
    grid = np.linspace(0.0, 1.0, levels)
    lut = np.stack([np.interp(grid, *zip(*_JET[c])) for c in "rgb"], axis=1)
    v = np.asarray(values, dtype=np.float64)
    idx = np.clip((v * levels).astype(np.int64), 0, levels - 1)
    idx[v == 1.0] = levels - 1
    return lut[idx]


def write_ply_points(path, points, colors):
    """ASCII PLY point cloud: float xyz + uchar rgb (colors in [0, 1])."""
    rgb = np.clip(np.round(np.asarray(colors) * 255.0), 0, 255).astype(np.uint8)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(points))
        for p, c in zip(np.asarray(points, dtype=np.float64), rgb):
            f.write("%.6f %.6f %.6f %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2]))


def write_ply_mesh(path, vertices, colors, triangles):
    rgb = np.clip(np.round(np.asarray(colors) * 255.0), 0, 255).astype(np.uint8)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %d\n"
                "property list uchar int vertex_indices\nend_header\n" % (len(vertices), len(triangles)))
        for p, c in zip(np.asarray(vertices, dtype=np.float64), rgb):
            f.write("%.6f %.6f %.6f %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2]))
        for t in triangles:
            f.write("3 %d %d %d\n" % tuple(t))


def frame_glyphs(scene_points, frame_t, frame_R, stride=2):
    """The 12-vertex glyph per `stride`-th point of file_logger_cls.py:121-158 (vectorised): point -> predicted origin
    (red), x / y / z axis markers (green / yellow / blue); only the y-axis triangle is a face."""
    j = np.arange(0, scene_points.shape[0], stride)
    p, t, R = scene_points[j], frame_t[j], frame_R[j]
    ax = lambda a, b: (t + R[:, :, a] * 0.01 + R[:, :, b] * 0.001, t + R[:, :, a] * 0.01)
    x1, x2 = ax(0, 1)
    y1, y2 = ax(1, 2)
    z1, z2 = ax(2, 0)
    verts = np.stack([p, t * 0.5 + p * 0.5 + 0.0001, t, t, x1, x2, t, y1, y2, t, z1, z2], axis=1).reshape(-1, 3)
    col = np.tile(np.repeat(np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1]], dtype=np.float64), 3, axis=0), (len(j), 1))
    tri = 12 * np.arange(len(j))[:, None] + np.array([[6, 7, 8]])
    return verts, col, tri


def top_frames(scene_points, scene_pred, frame_R, frame_t, k=50, device=None):
    """file_logger_cls.py:186-232: frames of the k best-scoring points, orthonormalised, collision-checked against the
    cloud.  Returns (H (n, 4, 4) float64, scores list).  The collision test runs on the device for all k poses at once
    (postprocess.GraspPostProcessor.collision_free = cloud_processor/view_collision_checker.py:37-65, the same test as
    EvalExpCloud.view_non_collision, eval_point_cloud.py:115-144)."""
    from .postprocess import GraspPostProcessor
    top = np.argsort(-scene_pred)[:k]
    R = frame_R[top]
    x = R[:, :, 0] / np.linalg.norm(R[:, :, 0], axis=1, keepdims=True)
    y = R[:, :, 1] - np.sum(x * R[:, :, 1], axis=1, keepdims=True) * x
    y = y / np.linalg.norm(y, axis=1, keepdims=True)
    H = np.tile(np.eye(4), (len(top), 1, 1))
    H[:, :3, :3] = np.stack([x, y, np.cross(x, y)], axis=2)
    H[:, :3, 3] = frame_t[top]
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    ok = GraspPostProcessor().collision_free(torch.from_numpy(H).to(device),
                                             torch.from_numpy(np.ascontiguousarray(scene_points, dtype=np.float32)).to(device))
    ok = ok.cpu().numpy()
    return H[ok], [float(s) for s in scene_pred[top][ok]]


def loggin_to_file(data_batch, preds, step, output_dir, prefix="", with_label=True, work_dir="."):
    """Reference signature (utils/file_logger_cls.py:12) + ``work_dir`` for the two files the reference drops into the
    current directory (``top_frames.npy``, ``postprocess_time_ours.txt``)."""
    step_dir = os.path.join(output_dir, "{}_step{:05d}".format(prefix, step))
    os.makedirs(step_dir, exist_ok=True)
    if "score" not in preds:
        return None
    save = lambda name, arr, fmt: np.savetxt(os.path.join(step_dir, name), arr, fmt=fmt)
    scene_points = data_batch["scene_points"][0].detach().cpu().numpy().T
    save("scene_points.xyz", scene_points, "%.4f")
    if with_label:
        save("gt_scene_score.txt", data_batch["scene_score"][0].cpu().numpy(), "%.4f")
        save("gt_scene_score_labels.txt", data_batch["scene_score_labels"][0].cpu().numpy(), "%d")
    probs = torch.softmax(preds["score"][0].float(), dim=0).detach().cpu().numpy().T
    save("scene_score_logits.txt", probs, "%.4f")
    raw_R = preds["frame_R"][0].float().transpose(0, 1).detach().cpu().numpy()
    save("pred_frame_R.txt", raw_R, "%.4f")
    frame_R = raw_R.reshape(-1, 3, 3)
    t_prob = torch.softmax(preds["frame_t"][0].float(), dim=0).transpose(0, 1).detach().cpu().numpy()
    frame_t = -(t_prob * T_SCORE[None, :]).sum(1, keepdims=True) * frame_R[:, :, 0] + scene_points
    save("pred_frame_t.txt", frame_t, "%.4f")
    classes = probs.shape[1]
    scene_pred = np.sum(np.linspace(0, 1, classes + 1)[:-1][None, :] * probs, axis=1)
    write_ply_mesh(os.path.join(step_dir, "pred_frame.ply"), *frame_glyphs(scene_points, frame_t, frame_R))
    write_ply_points(os.path.join(step_dir, "pred_pts.ply"), scene_points, jet_colors(scene_pred))
    save("pred_scene_score.txt", scene_pred, "%.4f")
    if with_label:
        return None
    tic = time.time()
    dev = preds["score"].device if preds["score"].is_cuda else None
    top_H, score = top_frames(scene_points, scene_pred, frame_R, frame_t, k=50, device=dev)
    with open(os.path.join(work_dir, "postprocess_time_ours.txt"), "a+") as f:
        f.write("{:.4f}\n".format((time.time() - tic) * 1000.0))
    if len(top_H) > 0:
        np.save(os.path.join(work_dir, "top_frames.npy"), top_H)
    return top_H, score
