"""CPU suite: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls
without a GPU), the python surface refuses CPU tensors loudly, and the model surface is
reference-compatible (state_dict layout, parameter count, seeded init)."""
import ctypes
import glob
import os
import re

import pytest
import torch

from tests.conftest import ROOT


def _declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(s4g_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from s4g_release_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 12
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), "libs4g_b200.so does not export %s" % name
    assert dll.s4g_version() >= 1
    # the ctypes table and the header agree
    assert set(_lib.exported_symbols()) == set(declared)


def test_argument_errors_do_not_need_a_gpu():
    from s4g_release_b200 import _lib
    lib = _lib.lib
    # num_centroids > num_points: CHECK_GE in sampling_kernel.cu:139 -> S4G_E_ARG before any CUDA call
    dummy = ctypes.c_void_p(16)
    assert lib.s4g_farthest_point_sample_f32(dummy, 1, 8, 9, dummy, None) == -1
    assert b"num_points < num_centroids" in lib.s4g_last_error()
    assert lib.s4g_point_search_f32(dummy, dummy, 1, 8, 8, 2, dummy, dummy, None) == -1
    assert lib.s4g_point_search_f32(dummy, dummy, 1, 8, 2, 3, dummy, dummy, None) == -1
    assert lib.s4g_ball_query_f32(None, dummy, 1, 8, 8, 0.1, 4, dummy, dummy, None) == -1


def test_cpu_tensors_are_refused():
    from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext
    x = torch.zeros(1, 3, 16)
    idx = torch.zeros(1, 4, 2, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pn2_ext.farthest_point_sample(x, 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pn2_ext.ball_query(x, x, 0.1, 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pn2_ext.group_points_forward(x, idx)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pn2_ext.point_search(x, x, 3)


def test_model_surface_is_reference_compatible(golden_full):
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2, PointNet2Loss
    torch.manual_seed(0)
    net = PointNet2(**PN2_CLS_CONFIG)
    sd = net.state_dict()
    assert len(sd) == 200
    assert sum(p.numel() for p in net.parameters()) == int(golden_full["n_params"]) == 6632213
    assert tuple(sd["sa_modules.0.mlp.0.conv.weight"].shape) == (128, 3, 1, 1)
    assert tuple(sd["sa_modules.2.mlp.2.conv.weight"].shape) == (1024, 512, 1, 1)
    assert tuple(sd["fp_modules.0.mlp.0.conv.weight"].shape) == (1024, 1536, 1)
    assert tuple(sd["fp_modules.2.mlp.0.conv.weight"].shape) == (256, 512, 1)
    assert tuple(sd["mlp_seg.0.conv.weight"].shape) == (512, 256, 1)
    assert tuple(sd["movable_logit.0.weight"].shape) == (5, 128, 1)
    for k in ("seg_logit.bias", "R_logit.weight", "t_logit.bias", "mlp_R.3.bn.running_var",
              "mlp_movable.0.bn.num_batches_tracked"):
        assert k in sd
    # a DataParallel-style checkpoint ("module." prefix, utils/checkpoint.py:81-89) loads after stripping
    net.load_state_dict({k: v for k, v in sd.items()}, strict=True)
    # loss restatement (PointNet2_tcls.py:162-219) on random predictions
    B, N, n = 2, 64, 40
    g = torch.Generator().manual_seed(3)
    preds = {"score": torch.randn(B, 3, N, generator=g), "frame_R": torch.randn(B, 9, N, generator=g),
             "frame_t": torch.randn(B, 4, N, generator=g), "movable_logits": torch.rand(B, 5, N, generator=g)}
    labels = {"scene_score_labels": torch.randint(0, 3, (B, N), generator=g),
              "scene_movable_labels": torch.randint(0, 2, (B, 5, N), generator=g).float(),
              "best_frame_R": torch.randn(B, 9, n, generator=g),
              "best_frame_t": torch.randint(0, 4, (B, n), generator=g),
              "scene_score": torch.rand(B, N, generator=g)}
    loss = PointNet2Loss(neg_weight=0.5)(preds, labels)
    # literal reference formula
    gt = labels["best_frame_R"]
    inv = gt.clone()
    for a, b in ((1, 3), (4, 6), (7, 9)):
        inv[:, a:b] = -inv[:, a:b]
    p = preds["frame_R"][:, :, :n]
    l1 = ((p - gt) ** 2).mean(1, True)
    l2 = ((p - inv) ** 2).mean(1, True)
    ref_R = (torch.min(torch.cat([l1, l2], 1), dim=1)[0] * labels["scene_score"][:, :n]).mean() * 5.0
    assert torch.allclose(loss["R_loss"], ref_R)
    assert set(loss) == {"cls_loss", "R_loss", "t_loss", "mov_loss"}
