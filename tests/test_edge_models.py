"""EDGEPN2D / EDGEPN2DU model files, the PN2 metric and the model factory against tests/golden/edge_models.npz — outputs
of the REFERENCE's own EdgePointNet2Down and PointNet2Metric classes (tests/golden/make_edge_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

CFG = dict(score_classes=3, num_centroids=(256, 64, 16, 0), radius=(0.1, 0.2, 0.4, -1.0), num_neighbours=(16, 16, 8, -1),
           sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256)),
           fp_channels=((64, 64), (64, 32), (32, 32), (32, 32, 16)), num_fp_neighbours=(0, 3, 3, 3), seg_channels=(32,),
           dropout_prob=0.5)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "edge_models.npz")))


def test_edge_down_module_surface(gold):
    from s4g_release_b200.network_models.models.EdgePointNet2Down import EdgePointNet2Down
    torch.manual_seed(0)
    net = EdgePointNet2Down(**CFG)
    sd = net.state_dict()
    ref = {k[3:]: v for k, v in gold.items() if k.startswith("sd/")}
    assert list(sd) == list(ref) and all(tuple(sd[k].shape) == ref[k].shape for k in sd)
    for k in ("sa_modules.1.mlp.0.conv.weight", "fp_modules.2.mlp.1.conv.weight", "seg_logit.weight"):
        assert np.array_equal(sd[k].numpy(), ref[k]), k  # same creation order: identical seeded init
    assert not net.fusable()


def test_pn2_metric_matches_the_reference_class(gold):
    from s4g_release_b200.network_models.models.PointNet2 import PointNet2Metric
    preds = {k[4:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("out/")}
    labels = {k[6:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("label/")}
    out = PointNet2Metric()(preds, labels)
    assert np.array_equal(out["cls_acc"].numpy(), gold["metric/cls_acc"])
    assert np.array_equal(out["mov_acc"].numpy(), gold["metric/mov_acc"])
    np.testing.assert_allclose(out["R_err"].item(), gold["metric/R_err"].item(), rtol=1e-5)
    np.testing.assert_allclose(out["t_err"].item(), gold["metric/t_err"].item(), rtol=1e-6)


def _cfg_node(model_type, key):
    node = types.SimpleNamespace(NUM_CENTROIDS=CFG["num_centroids"], RADIUS=CFG["radius"], NUM_NEIGHBOURS=CFG["num_neighbours"],
                                 SA_CHANNELS=CFG["sa_channels"], FP_CHANNELS=CFG["fp_channels"],
                                 NUM_FP_NEIGHBOURS=CFG["num_fp_neighbours"], SEG_CHANNELS=CFG["seg_channels"],
                                 DROPOUT_PROB=0.5, LABEL_SMOOTHING=0.0, NEG_WEIGHT=0.5)
    return types.SimpleNamespace(MODEL=types.SimpleNamespace(TYPE=model_type, **{key: node}),
                                 DATA=types.SimpleNamespace(SCORE_CLASSES=3, NUM_REMOVAL_DIRECTIONS=5))


@pytest.mark.parametrize("model_type,key,cls", [("PN2", "PN2", "PointNet2"), ("PN2_CLS", "PN2", "PointNet2"),
                                                ("PN2_LOCAL", "PN2", "PointNet2"), ("EDGEPN2D", "EDGEPN2D", "EdgePointNet2Down"),
                                                ("EDGEPN2DU", "EDGEPN2DU", "EdgePointNet2DownUp")])
def test_build_model_dispatch(model_type, key, cls):
    """reference build_model.py:13-31: MODEL.TYPE -> (net, loss, metric)"""
    from s4g_release_b200.network_models.models.build_model import build_model
    net, loss, metric = build_model(_cfg_node(model_type, key))
    assert type(net).__name__ == cls and isinstance(loss, torch.nn.Module) and isinstance(metric, torch.nn.Module)


def test_build_model_rejects_what_it_does_not_serve():
    from s4g_release_b200.network_models.models.build_model import build_model
    with pytest.raises(ValueError, match="baselines"):
        build_model(_cfg_node("GPD", "PN2"))
    with pytest.raises(ValueError, match="Unknown model"):
        build_model(_cfg_node("nope", "PN2"))


@pytest.mark.gpu
def test_edge_down_forward_on_the_sm100a_operators(gold):
    from s4g_release_b200.network_models.models.EdgePointNet2Down import EdgePointNet2Down
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = EdgePointNet2Down(**CFG)
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("sd/")}, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net({"scene_points": torch.from_numpy(gold["points"]).cuda()})
    for k in ("scene_score_logits", "frame_R", "frame_t", "movable_logits"):
        want = gold["out/" + k]
        np.testing.assert_allclose(out[k].cpu().numpy(), want, atol=2e-4 * max(1.0, np.abs(want).max()), rtol=0)


@pytest.mark.gpu
def test_edge_downup_constructs_and_runs(gold):
    """the reference's class raises NameError at construction (recorded in the golden); ours builds the same module tree
    and runs its two heads"""
    from s4g_release_b200.network_models.models.EdgePointNet2DownUp import EdgePointNet2DownUp
    assert "NameError" in str(gold["downup_in_reference"])
    net = EdgePointNet2DownUp(**CFG).cuda().eval()
    with torch.no_grad():
        out = net({"scene_points": torch.from_numpy(gold["points"]).cuda()})
    assert tuple(out["scene_score_logits"].shape) == (2, 3, 1024) and tuple(out["frame_R"].shape) == (2, 9, 1024)
    assert all(torch.isfinite(v).all() for v in out.values())
