"""GPU suite for the training (autograd) path: the module stack on the sm_100a ops gives the same loss and
parameter gradients as the same stack with the two differentiable gathers replaced by torch-native
gathers (autograd of torch.gather), and a few Adam steps reduce the loss.  fp32, TF32 off; tolerance 2e-3
relative to the largest gradient entry (the backward scatters use atomicAdd, as the reference's do)."""
import pytest
import torch

from tests.inputs import TINY_CONFIG

pytestmark = pytest.mark.gpu


def _torch_group_points(points, index):
    B, C, N = points.shape
    _, M, K = index.shape
    return torch.gather(points.unsqueeze(2).expand(B, C, M, N), 3, index.unsqueeze(1).expand(B, C, M, K))


def _torch_feature_interpolate(feature, index, weight):
    g = _torch_group_points(feature, index)  # (B,C,Nq,3)
    return (g * weight.unsqueeze(1)).sum(-1)


def _make(seed=0):
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2
    cfg = dict(TINY_CONFIG)
    cfg["dropout_prob"] = 0.0
    torch.manual_seed(seed)
    return PointNet2(**cfg).cuda()


def test_gradients_match_torch_native_gathers(monkeypatch):
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2Loss
    from s4g_release_b200.network_models.models.pointnet2_utils import functions as F_
    from s4g_release_b200.train import synthetic_labels
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _make().train()
    loss_fn = PointNet2Loss(neg_weight=0.5)
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(2, 3, 1024, generator=g).cuda()
    labels = synthetic_labels(2, 1024, num_frame=200, device="cuda")

    def run():
        net.zero_grad(set_to_none=True)
        losses = loss_fn(net({"scene_points": pts}), labels)
        total = sum(losses.values())
        total.backward()
        return total.item(), {n: p.grad.clone() for n, p in net.named_parameters()}

    # BatchNorm running stats change between the two runs but do not enter the train-mode forward
    loss_a, grads_a = run()
    monkeypatch.setattr(F_, "group_points", _torch_group_points)
    monkeypatch.setattr(F_, "feature_interpolate", _torch_feature_interpolate)
    loss_b, grads_b = run()
    assert abs(loss_a - loss_b) <= 1e-4 * max(1.0, abs(loss_b))
    for n in grads_a:
        scale = max(grads_b[n].abs().max().item(), 1e-6)
        assert (grads_a[n] - grads_b[n]).abs().max().item() <= 2e-3 * scale, n


def test_adam_steps_reduce_the_loss():
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2Loss
    from s4g_release_b200.train import Trainer, synthetic_labels
    net = _make(1)
    trainer = Trainer(net, PointNet2Loss(neg_weight=0.5), lr=1e-3)
    pts = torch.rand(2, 3, 1024, generator=torch.Generator().manual_seed(4)).cuda()
    labels = synthetic_labels(2, 1024, num_frame=200, device="cuda")
    first = sum(trainer.step({"scene_points": pts}, labels).values()).item()
    for _ in range(8):
        last = sum(trainer.step({"scene_points": pts}, labels).values()).item()
    assert last < first
    # after training, the eval-mode fused path must see the new parameters (engine is rebuilt)
    net.eval()
    with torch.no_grad():
        out = net({"scene_points": pts})
        ref = net({"scene_points": pts}, fused=False)
    for k in out:
        err = (out[k] - ref[k]).abs().max().item()
        assert err <= 6e-2 * max(1.0, ref[k].abs().max().item()), k
