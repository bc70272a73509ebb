"""CPU: CheckPointer writes / reads the reference's checkpoint layout (utils/checkpoint.py) — and, when the reference
tree is present (build container), a checkpoint written by the REFERENCE's own CheckPointer loads into our model
and vice versa."""
import importlib.util
import os

import pytest
import torch

from tests.inputs import TINY_CONFIG


def _model():
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2
    torch.manual_seed(0)
    return PointNet2(**TINY_CONFIG)


def test_roundtrip_and_layout(tmp_path):
    from s4g_release_b200.checkpoint import CheckPointer
    net = _model()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=20, gamma=0.5)
    cp = CheckPointer(net, opt, sch, save_dir=str(tmp_path))
    assert not cp.has_checkpoint() and cp.load() == {}
    cp.save("model_003", epoch=3, best_metric=0.5)
    raw = torch.load(tmp_path / "model_003.pth", map_location="cpu")
    assert set(raw) == {"model", "optimizer", "scheduler", "epoch", "best_metric"}
    assert list(raw["model"]) == list(net.state_dict())
    assert (tmp_path / "last_checkpoint").read_text() == str(tmp_path / "model_003.pth")
    net2 = _model()
    for p in net2.parameters():
        p.data.add_(1.0)
    extra = CheckPointer(net2, save_dir=str(tmp_path)).load()
    assert extra["epoch"] == 3 and extra["best_metric"] == 0.5
    for (k, a), b in zip(net.state_dict().items(), net2.state_dict().values()):
        assert torch.equal(a, b), k


def test_dataparallel_prefix_is_stripped(tmp_path):
    from s4g_release_b200.checkpoint import CheckPointer
    net = _model()
    torch.save({"model": {"module." + k: v for k, v in net.state_dict().items()}}, tmp_path / "dp.pth")
    net2 = _model()
    for p in net2.parameters():
        p.data.zero_()
    CheckPointer(net2, save_dir="").load(str(tmp_path / "dp.pth"), resume=False)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))


REF = "/root/reference/inference/grasp_proposal/utils/checkpoint.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_interoperates_with_the_reference_checkpointer(tmp_path):
    from s4g_release_b200.checkpoint import CheckPointer
    spec = importlib.util.spec_from_file_location("ref_checkpoint", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    net = _model()
    ref.CheckPointer(net, save_dir=str(tmp_path)).save("from_ref", epoch=7)
    net2 = _model()
    for p in net2.parameters():
        p.data.zero_()
    assert CheckPointer(net2, save_dir=str(tmp_path)).load()["epoch"] == 7
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    ours = tmp_path / "ours"
    ours.mkdir()
    CheckPointer(net, save_dir=str(ours)).save("from_ours", epoch=9)
    net3 = _model()
    for p in net3.parameters():
        p.data.zero_()
    assert ref.CheckPointer(net3, save_dir=str(ours)).load()["epoch"] == 9
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net3.state_dict().values()))
