"""The PN2 sibling model (network_models/models/PointNet2.py) against tests/golden/pn2_sibling.npz, which holds the outputs
of the REFERENCE's own PointNet2 / PointNet2Loss / toRotMatrix classes (tests/golden/make_pn2_golden.py)."""
import os

import numpy as np
import pytest
import torch

CFG = dict(score_classes=3, num_centroids=(256, 64, 16, 0), radius=(0.1, 0.2, 0.4, -1.0), num_neighbours=(16, 16, 8, -1),
           sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256)),
           fp_channels=((64, 64), (64, 32), (32, 32), (32, 32, 16)), num_fp_neighbours=(0, 3, 3, 3), seg_channels=(32,),
           num_removal_directions=5, dropout_prob=0.5)


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "pn2_sibling.npz")))


def test_to_rot_matrix_bit_exact(gold):
    from s4g_release_b200.network_models.functions.functions import toRotMatrix
    got = toRotMatrix(torch.from_numpy(gold["rot6d"]))
    assert np.array_equal(got.numpy(), gold["rot9"])
    R = got.reshape(3, 3, 3, 50).permute(0, 3, 1, 2)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3).expand_as(R), atol=1e-5)


def test_module_surface_and_seeded_init(gold):
    from s4g_release_b200.network_models.models.PointNet2 import PointNet2
    torch.manual_seed(0)
    net = PointNet2(**CFG)
    sd = net.state_dict()
    ref = {k[3:]: v for k, v in gold.items() if k.startswith("sd/")}
    assert list(sd) == list(ref) and all(tuple(sd[k].shape) == ref[k].shape for k in sd)
    # same parameter creation order => identical default init under the same seed (BN / t_logit were re-seeded later)
    for k in ("sa_modules.0.mlp.0.conv.weight", "sa_modules.3.mlp.2.conv.weight", "fp_modules.0.mlp.1.conv.weight",
              "seg_logit.weight", "R_logit.bias", "movable_logit.0.weight"):
        assert np.array_equal(sd[k].numpy(), ref[k]), k
    assert float(net.t_logit.weight.detach().abs().max()) == 0.0 and float(net.t_logit.bias.detach().abs().max()) == 0.0  # :150-152


def test_loss_matches_reference(gold):
    from s4g_release_b200.network_models.models.PointNet2 import PointNet2Loss
    preds = {k[4:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("out/")}
    labels = {k[6:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("label/")}
    loss = PointNet2Loss()(preds, labels)
    for k, v in loss.items():
        np.testing.assert_allclose(v.item(), gold["loss/" + k].item(), rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_forward_on_the_sm100a_operators(gold):
    """module path (global set-abstraction level, zero-neighbour propagation, 6-D rotation head) on the GPU operators"""
    from s4g_release_b200.network_models.models.PointNet2 import PointNet2
    net = PointNet2(**CFG)
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("sd/")}, strict=True)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net({"scene_points": torch.from_numpy(gold["points"]).cuda()})
    for k in ("scene_score_logits", "frame_R", "frame_t", "movable_logits"):
        want = gold["out/" + k]
        np.testing.assert_allclose(out[k].cpu().numpy(), want, atol=2e-4 * max(1.0, np.abs(want).max()), rtol=0)


@pytest.mark.gpu
def test_fused_engine_serves_the_sibling_heads():
    """without a global level the sibling runs on the fused tcgen05 engine: 6-channel / 3-channel logits epilogues"""
    from s4g_release_b200.network_models.models.PointNet2 import PointNet2
    cfg = dict(CFG, num_centroids=(256, 64, 16), radius=(0.1, 0.2, 0.4), num_neighbours=(16, 16, 8),
               sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128)), fp_channels=((128, 128), (64, 64), (32, 32, 32)),
               num_fp_neighbours=(3, 3, 3), seg_channels=(64, 32))
    torch.manual_seed(3)
    net = PointNet2(**cfg)
    torch.nn.init.normal_(net.t_logit.weight, std=0.05)
    net = net.cuda().eval()
    pts = torch.rand(2, 3, 1024, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        fused = net({"scene_points": pts})
        ref = net({"scene_points": pts}, fused=False)
    for k in ref:
        err = (fused[k] - ref[k]).abs().max().item()
        assert err <= 6e-2 * max(1.0, ref[k].abs().max().item()), (k, err)
