"""CPU: PN2_CLS PointNet2Loss (label smoothing off / on), PointNet2Metric and the builder triple against
tests/golden/pn2cls_loss.npz — outputs of the REFERENCE's own classes (tests/golden/make_pn2cls_loss_golden.py;
reference PointNet2_tcls.py:156-268, nn_utils/functional.py:91-114)."""
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "pn2cls_loss.npz")))


def _io(gold):
    preds = {k[5:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("pred/")}
    labels = {k[6:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("label/")}
    return preds, labels


@pytest.mark.parametrize("tag,smoothing", [("loss", 0.0), ("loss_smooth", 0.1)])
def test_loss_terms_match_the_reference_class(gold, tag, smoothing):
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2Loss
    preds, labels = _io(gold)
    out = PointNet2Loss(label_smoothing=smoothing, neg_weight=0.5)(preds, labels)
    assert set(out) == {"cls_loss", "R_loss", "t_loss", "mov_loss"}
    for k, v in out.items():
        np.testing.assert_allclose(v.item(), gold[tag + "/" + k].item(), rtol=2e-6, atol=1e-7, err_msg=k)


def test_loss_accepts_the_key_the_forward_emits(gold):
    """the reference's loss reads "scene_score_logits" although its forward emits "score" (:142, :163)"""
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2Loss
    preds, labels = _io(gold)
    preds["score"] = preds.pop("scene_score_logits")
    out = PointNet2Loss(neg_weight=0.5)(preds, labels)
    np.testing.assert_allclose(out["cls_loss"].item(), gold["loss/cls_loss"].item(), rtol=2e-6)


def test_metric_matches_the_reference_class(gold):
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2Metric
    preds, labels = _io(gold)
    out = PointNet2Metric()(preds, labels)
    assert set(out) == {"cls_acc", "mov_acc", "R_err", "t_acc"}
    for k in ("cls_acc", "mov_acc", "t_acc"):  # per-element 0/1 tensors
        assert np.array_equal(out[k].numpy(), gold["metric/" + k]), k
    np.testing.assert_allclose(out["R_err"].item(), gold["metric/R_err"].item(), rtol=1e-5)


def test_builder_returns_the_reference_triple():
    from s4g_release_b200.network_models.models.PointNet2_tcls import (PointNet2, PointNet2Loss, PointNet2Metric,
                                                                        build_pointnet2_cls)
    net, loss, metric = build_pointnet2_cls()
    assert isinstance(net, PointNet2) and isinstance(loss, PointNet2Loss) and isinstance(metric, PointNet2Metric)
    assert loss.neg_weight == 0.5 and net.fusable()


def test_fusable_guard_and_engine_invalidation():
    """default ctor arguments (global level, num_fp_neighbours 0) must not reach the fused planner; cached
    engines must not survive load_state_dict / .to / deepcopy."""
    import copy
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2
    from tests.inputs import TINY_CONFIG
    assert not PointNet2(score_classes=3).fusable()  # the class defaults: last level num_centroids = 0
    assert PointNet2(**PN2_CLS_CONFIG).fusable() and PointNet2(**TINY_CONFIG).fusable()
    odd = dict(TINY_CONFIG, num_neighbours=(16, 16, 12))
    assert not PointNet2(**odd).fusable()
    net = PointNet2(**TINY_CONFIG).eval()
    sentinel = object()
    net.attach_engine(sentinel)
    assert net.fused_engine.__self__ is net and net._engine is sentinel
    key = net._engine_key
    with torch.no_grad():
        net.seg_logit.bias.add_(1.0)  # in-place update (optimizer / EMA style)
    assert net._param_fingerprint() != key
    net.attach_engine(sentinel)
    net.load_state_dict(net.state_dict())
    assert net._engine is None
    net.attach_engine(sentinel)
    clone = copy.deepcopy(net)
    assert clone._engine is None and net._engine is sentinel
    net.float()
    assert net._engine is None
