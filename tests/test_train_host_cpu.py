"""CPU suite for the HOST logic of the fused training step (s4g_release_b200/train_engine.py): what can be checked
without a kernel — the weight re-ordering of a grouped block, the order in which a chain's blocks hand gradients to
each other (with and without the reduce fused into the input-gradient GEMM), and the assumption behind the head-by-head
schedule (every term of PointNet2Loss depends on ONE head).  The kernels themselves: tests/test_train_engine_gpu.py."""
import pytest
import torch

from s4g_release_b200 import train_engine as te


def _conv_block(cin, cout):
    from s4g_release_b200.network_models.nn_utils.conv import Conv2d
    torch.manual_seed(cin + cout)
    return Conv2d(cin, cout, 1)


@pytest.mark.parametrize("cf", [0, 32])
def test_grouped_block_weight_rows_match_the_reference_channel_order(cf):
    """reference grouping concatenates [relative xyz (3) | features (cf)] (pointnet2_utils/modules.py:49-53); the row
    layout is [features | xyz | 0 x 5]: the re-ordered bf16 weights must give the same products"""
    blk = _conv_block(cf + 3, 16)
    b = te.Block(blk, grouped_cf=cf)
    wb = b._weight_rows().float()
    assert wb.shape == (16, cf + 8)
    g = torch.Generator().manual_seed(1)
    xyz, feat = torch.randn(50, 3, generator=g), torch.randn(50, cf, generator=g)
    ref_in = torch.cat([xyz, feat], dim=1)                       # the reference's channel order
    rows = torch.cat([feat, xyz, torch.zeros(50, 5)], dim=1)     # the row layout of s4g_train_group_rows_bf16
    w = blk.conv.weight.detach().reshape(16, cf + 3).to(torch.bfloat16).float()
    assert torch.allclose(rows @ wb.t(), ref_in @ w.t(), atol=1e-6)


def test_plain_block_weight_rows_are_padded_to_eight_columns():
    blk = _conv_block(13, 8)
    wb = te.Block(blk)._weight_rows()
    assert wb.shape == (8, 16) and wb.dtype == torch.bfloat16
    assert torch.equal(wb[:, :13].float(), blk.conv.weight.detach().reshape(8, 13).to(torch.bfloat16).float())
    assert (wb[:, 13:] == 0).all()


class _FakeBlock:
    """records how chain_backward calls it"""

    def __init__(self, name, log):
        self.name, self.log = name, log

    def backward(self, dz, need_dx=True, prev=None, pre=None):
        self.log.append((self.name, dz, need_dx, None if prev is None else prev.name, pre))
        out = "d" + self.name
        return (out, "sums_for_" + prev.name) if prev is not None else out


@pytest.mark.parametrize("fused", [False, True])
def test_chain_backward_hands_gradients_from_last_to_first(fused, monkeypatch):
    monkeypatch.setattr(te, "FUSED_BWD_REDUCE", fused)
    log = []
    blocks = [_FakeBlock(n, log) for n in ("b0", "b1", "b2")]
    out = te.chain_backward(blocks, "dz", need_dx=False)
    assert out == "db0"
    assert [e[0] for e in log] == ["b2", "b1", "b0"]
    assert [e[1] for e in log] == ["dz", "db2", "db1"]
    assert [e[2] for e in log] == [True, True, False]            # only the first block may skip its input gradient
    if fused:   # block l masks / reduces for block l-1 and hands the sums over
        assert [e[3] for e in log] == ["b1", "b0", None]
        assert [e[4] for e in log] == [None, "sums_for_b1", "sums_for_b0"]
    else:
        assert all(e[3] is None and e[4] is None for e in log)


def test_every_loss_term_depends_on_one_head_only():
    """TrainEngine._separable: the heads run forward + loss term + backward one after the other, feeding ZEROS for the
    other heads' predictions — valid only if d term_k / d prediction_j == 0 for j != k (PointNet2_tcls.py:162-219)"""
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2Loss
    from s4g_release_b200.train import synthetic_labels
    B, N, M = 2, 64, 40
    labels = synthetic_labels(B, N, num_frame=M, first_seed=3)
    g = torch.Generator().manual_seed(0)
    shapes = {"score": (B, 3, N), "frame_R": (B, 9, N), "frame_t": (B, 4, N), "movable_logits": (B, 5, N)}
    leaves = {k: torch.randn(s, generator=g).requires_grad_(True) for k, s in shapes.items()}
    preds = dict(leaves, movable_logits=torch.sigmoid(leaves["movable_logits"]))
    losses = PointNet2Loss(neg_weight=0.5)(preds, labels)
    assert tuple(losses) == te.TrainEngine.LOSS_KEYS
    for k, key in enumerate(te.TrainEngine.LOSS_KEYS):
        grads = torch.autograd.grad(losses[key], list(leaves.values()), retain_graph=True, allow_unused=True)
        for j, gr in enumerate(grads):
            if j == k:
                assert gr is not None and gr.abs().sum() > 0, key
            else:
                assert gr is None or gr.abs().sum() == 0, (key, te.TrainEngine.PRED_KEYS[j])
    # and with the other predictions replaced by zeros the term's value is unchanged (what step_loss does)
    for k, key in enumerate(te.TrainEngine.LOSS_KEYS):
        only = {n: (preds[n] if j == k else torch.zeros(shapes[n])) for j, n in enumerate(te.TrainEngine.PRED_KEYS)}
        assert torch.allclose(PointNet2Loss(neg_weight=0.5)(only, labels)[key], losses[key])


def test_engine_refuses_a_model_without_a_fused_plan():
    class _NoPlan(torch.nn.Module):
        def fusable(self):
            return False

    with pytest.raises(RuntimeError):
        te.TrainEngine(_NoPlan(), None)
