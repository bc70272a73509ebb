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


# ---------------------------------------------------------------- the GEMM's launch planning (host arithmetic, no GPU)
def _plan(P, N, K, mode, sms=148):
    import ctypes
    from s4g_release_b200._lib import check, lib
    out = (ctypes.c_longlong * 7)()
    check(lib.s4g_gemm_bf16_plan(P, N, K, mode, sms, ctypes.cast(out, ctypes.c_void_p)), "gemm_bf16_plan")
    return dict(zip(("bn", "groups", "ws", "stages", "grid", "threads", "smem"), list(out)))


def _model_gemm_shapes(scenes):
    """(rows, K, N) of every conv of PN2_CLS at `scenes` scenes: set abstraction, propagation, heads"""
    r0, r1, r2, rp = scenes * 5120 * 64, scenes * 1024 * 64, scenes * 256 * 64, scenes * 25600
    fwd = [(r0, 8, 128), (r0, 128, 128), (r0, 128, 256), (r1, 264, 256), (r1, 256, 256), (r1, 256, 512), (r2, 520, 512),
           (r2, 512, 512), (r2, 512, 1024), (scenes * 1024, 1536, 1024), (scenes * 1024, 1024, 512), (scenes * 5120, 768, 512),
           (scenes * 5120, 512, 512), (rp, 512, 256), (rp, 256, 256), (rp, 256, 512), (rp, 512, 256), (rp, 256, 128)]
    return fwd


@pytest.mark.parametrize("scenes", [1, 2, 32, 64])
def test_gemm_plans_fit_the_shared_memory_of_an_sm(scenes):
    limit = 227 * 1024 - 256  # dynamic + ~176 B static
    for rows, K, N in _model_gemm_shapes(scenes):
        for mode, (k, n) in ((0, (K, N)), (1, (K, N)), (0, (N, K)), (2, (N, (K + 7) // 8 * 8))):  # forward, dX, dX + reduce
            p = _plan(rows, n, k, mode)
            assert 0 < p["smem"] <= limit, (rows, k, n, mode, p)
            assert p["stages"] >= 2 and p["threads"] in (224, 352) and 1 <= p["grid"] <= 148, (rows, k, n, mode, p)
            assert p["bn"] in (128, 256) and p["groups"] in (1, 2)
            if p["bn"] == 256:
                assert p["groups"] == 2 and p["ws"] == 0 and n > 128
            if mode == 2:
                assert p["bn"] == 128 and p["groups"] == 1


def test_gemm_plan_rules_follow_the_measurements():
    """the per-shape choices DESIGN.md §3 "Training" reports (profiles/r02/gemm_layers_v6.txt)"""
    R0, R1, R2 = 32 * 5120 * 64, 32 * 1024 * 64, 32 * 256 * 64
    p = _plan(R0, 128, 8, 1)        # one K slab: epilogue-bound -> two groups alternating tiles, weights resident
    assert (p["bn"], p["groups"], p["ws"]) == (128, 2, 1)
    p = _plan(R0, 256, 128, 1)
    assert (p["bn"], p["groups"], p["ws"]) == (128, 2, 1)
    p = _plan(R1, 256, 264, 1)      # a fifth slab of 8 columns: stays on the weight-stationary 128-column schedule
    assert (p["bn"], p["groups"], p["ws"]) == (128, 1, 1)
    for rows, K, N in ((R1, 256, 512), (R2, 512, 1024), (R2, 520, 512), (32 * 25600, 512, 256)):
        p = _plan(rows, N, K, 1)    # many K steps per tile: one N = 256 MMA per step, the groups split the tile
        assert (p["bn"], p["groups"], p["ws"], p["stages"]) == (256, 2, 0, 3), (rows, K, N, p)
    p = _plan(100, 16, 128, 0)      # a tiny problem: one CTA per tile
    assert p["grid"] == 1
