"""Score / SE(3)-pose parity metrics between two PN2_CLS prediction dicts for ONE scene (test + bench helper; numpy
fp64, no product and no oracle imports).  The quantities are the ones the reference's consumer derives from the raw
heads (grasp_detector.py:137-185): expected grasp score (softmax · linspace), the ``score > 0.7`` decision, the
approach-offset class / expected offset (softmax · [0.08, 0.06, 0.04, 0.02] m), the Gram-Schmidt'd rotation and the
gripper translation ``point - t · x_axis``."""
import numpy as np

T_SCORE = np.array([0.08, 0.06, 0.04, 0.02])


def _softmax(x, axis=0):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def expected_score(score_logits):
    """(C, N) logits -> (N,) expected score, grasp_detector.py:142-146."""
    p = _softmax(np.asarray(score_logits, dtype=np.float64), 0)
    c = p.shape[0]
    return (np.linspace(0, 1, c + 1)[1:, None] * p).sum(0)


def frames(frame_R):
    """(9, N) raw head output -> (N, 3, 3) orthonormal frames: row-major 3x3 per point, Gram-Schmidt on columns 0, 1,
    z = x × y (grasp_detector.py:124-135)."""
    R = np.asarray(frame_R, dtype=np.float64).T.reshape(-1, 3, 3)
    x = R[:, :, 0]
    x = x / np.linalg.norm(x, axis=1, keepdims=True)
    y = R[:, :, 1]
    y = y - (x * y).sum(1, keepdims=True) * x
    y = y / np.linalg.norm(y, axis=1, keepdims=True)
    return np.stack([x, y, np.cross(x, y)], axis=2)


def geodesic_deg(Ra, Rb):
    tr = np.einsum("nij,nij->n", Ra, Rb)
    return np.degrees(np.arccos(np.clip((tr - 1.0) / 2.0, -1.0, 1.0)))


def scene_metrics(cloud_3n, ref, got, threshold=0.7):
    """ref / got: dicts of numpy arrays for one scene — score (3,N), frame_R (9,N), frame_t (4,N), movable_logits (5,N).
    Returns a flat dict of floats."""
    pts = np.asarray(cloud_3n, dtype=np.float64).T
    s_r, s_g = expected_score(ref["score"]), expected_score(got["score"])
    ds = np.abs(s_r - s_g)
    dec_r, dec_g = s_r > threshold, s_g > threshold
    flips = dec_r != dec_g
    pt_r, pt_g = _softmax(np.asarray(ref["frame_t"], np.float64)), _softmax(np.asarray(got["frame_t"], np.float64))
    t_r, t_g = (pt_r * T_SCORE[:, None]).sum(0), (pt_g * T_SCORE[:, None]).sum(0)
    cls_r, cls_g = pt_r.argmax(0), pt_g.argmax(0)
    # a class flip only matters when the top two probabilities are not a near tie in the reference
    top2 = np.sort(pt_r, axis=0)[-2:]
    margin = top2[1] - top2[0]
    R_r, R_g = frames(ref["frame_R"]), frames(got["frame_R"])
    ang = geodesic_deg(R_r, R_g)
    pos_r = pts - t_r[:, None] * R_r[:, :, 0]
    pos_g = pts - t_g[:, None] * R_g[:, :, 0]
    dpos = np.linalg.norm(pos_r - pos_g, axis=1) * 1e3
    top_r, top_g = np.argsort(-s_r, kind="stable")[:50], np.argsort(-s_g, kind="stable")[:50]
    hi = dec_r | dec_g
    sel = hi if hi.any() else np.ones_like(hi)
    return {
        "score_abs_err_max": float(ds.max()), "score_abs_err_mean": float(ds.mean()),
        "threshold_flip_frac": float(flips.mean()),
        "threshold_flip_max_margin": float(np.abs(s_r - threshold)[flips].max()) if flips.any() else 0.0,
        "n_above_threshold_ref": int(dec_r.sum()),
        "t_class_flip_frac": float((cls_r != cls_g).mean()),
        "t_class_flip_max_margin": float(margin[cls_r != cls_g].max()) if (cls_r != cls_g).any() else 0.0,
        "t_offset_err_mm_max": float(np.abs(t_r - t_g).max() * 1e3),
        "rot_err_deg_max": float(ang.max()), "rot_err_deg_mean": float(ang.mean()),
        "rot_err_deg_p99": float(np.percentile(ang, 99)),
        "rot_err_deg_max_candidates": float(ang[sel].max()),
        "translation_err_mm_max": float(dpos.max()), "translation_err_mm_mean": float(dpos.mean()),
        "movable_abs_err_max": float(np.abs(np.asarray(ref["movable_logits"], np.float64) -
                                            np.asarray(got["movable_logits"], np.float64)).max()),
        "top50_overlap": float(len(set(top_r.tolist()) & set(top_g.tolist())) / 50.0),
        "logit_rel_err_max": float(max(np.abs(np.asarray(ref[k], np.float64) - np.asarray(got[k], np.float64)).max() /
                                       max(1.0, np.abs(np.asarray(ref[k])).max())
                                       for k in ("score", "frame_R", "frame_t"))),
    }


def summarize(per_scene):
    """worst case over scenes for error-like keys, minimum for overlaps."""
    out = {}
    for k in per_scene[0]:
        vals = [m[k] for m in per_scene]
        out[k] = min(vals) if "overlap" in k or k.startswith("n_") else max(vals)
    out["scenes"] = len(per_scene)
    return out
