"""GPU suite for the whole PN2_CLS forward: module (drop-in / training) path and fused inference engine
against the golden fixtures produced by the reference's python modules (tests/golden/make_golden.py) and
the CPU oracle.  Geometry must be bit-exact; head outputs within the stated tolerance:
  * module path, fp32 (TF32 off):                 |err| <= 2e-3 * max(1, max|ref|)
  * fused engine, torch fp32 MLP reference:        same
  * fused engine, tcgen05 (bf16 operands, fp32 acc, bf16 activations between layers, 17 layers deep):
                                                   |err| <= 6e-2 * max(1, max|ref|)   (engine.tolerance())"""
import numpy as np
import pytest
import torch

from tests.inputs import TINY_CONFIG

pytestmark = pytest.mark.gpu
HEADS = ("score", "frame_R", "frame_t", "movable_logits")


@pytest.fixture(scope="module")
def tiny_net(golden_tiny):
    from s4g_release_b200.network_models.models.PointNet2_tcls import PointNet2
    net = PointNet2(**TINY_CONFIG)
    sd = {k[3:]: torch.from_numpy(v) for k, v in golden_tiny.items() if k.startswith("sd/")}
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


def _rel_err(got, want):
    want = torch.as_tensor(want)
    return ((got.float().cpu() - want).abs().max() / max(1.0, want.abs().max().item())).item()


def test_module_path_reproduces_reference_modules(tiny_net, golden_tiny):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        out = tiny_net({"scene_points": torch.from_numpy(golden_tiny["points"]).cuda()}, fused=False)
    for k in HEADS:
        assert _rel_err(out[k], golden_tiny["out/" + k]) <= 2e-3, k


@pytest.mark.parametrize("backend", ["torch", "tcgen05"])
def test_fused_engine_tiny(tiny_net, golden_tiny, backend):
    from s4g_release_b200.engine import FusedPointNet2
    torch.backends.cuda.matmul.allow_tf32 = False
    eng = FusedPointNet2(tiny_net, mlp_backend=backend)
    out, trace = eng.forward(torch.from_numpy(golden_tiny["points"]).cuda(), return_trace=True)
    torch.cuda.synchronize()
    for i in range(3):
        assert np.array_equal(trace["fps"][i].cpu().numpy(), golden_tiny[f"sa{i}/fps_index"])
        assert np.array_equal(trace["ball"][i].cpu().numpy(), golden_tiny[f"sa{i}/ball_index"])
        sa = trace["sa_feature"][i].float().cpu().transpose(1, 2)
        assert _rel_err(sa, golden_tiny[f"sa{i}/new_feature"]) <= eng.tolerance(), "sa%d" % i
    for k in HEADS:
        assert _rel_err(out[k], golden_tiny["out/" + k]) <= eng.tolerance(), k


def test_default_forward_is_the_fused_tcgen05_path(tiny_net, golden_tiny):
    with torch.no_grad():
        out = tiny_net({"scene_points": torch.from_numpy(golden_tiny["points"]).cuda()})
    assert tiny_net.fused_engine().mlp_backend == "tcgen05"
    for k in HEADS:
        assert _rel_err(out[k], golden_tiny["out/" + k]) <= tiny_net.fused_engine().tolerance(), k


@pytest.mark.parametrize("backend", ["torch", "tcgen05"])
def test_full_model_on_fixture_cloud(cloud_2638, golden_full, backend):
    """BASELINE config 1: 2638_view_0 subsample, shipped architecture, seeded weights."""
    from s4g_release_b200.engine import FusedPointNet2
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2
    from tests.golden.make_golden import seed_reference_weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = seed_reference_weights(PointNet2(**PN2_CLS_CONFIG)).cuda().eval()
    eng = FusedPointNet2(net, mlp_backend=backend)
    out, trace = eng.forward(torch.from_numpy(cloud_2638)[None].cuda(), return_trace=True)
    torch.cuda.synchronize()
    for i in range(3):
        assert np.array_equal(trace["fps"][i].cpu().numpy(), golden_full[f"sa{i}/fps_index"])
        assert np.array_equal(trace["ball"][i].sum(dim=2).cpu().numpy(), golden_full[f"sa{i}/ball_index_sum"])
        sa = trace["sa_feature"][i].float().cpu().transpose(1, 2)[:, ::8, ::8]
        assert _rel_err(sa, golden_full[f"sa{i}/new_feature_s"]) <= eng.tolerance(), "sa%d" % i
    stride = int(golden_full["stride"])
    for k in HEADS:
        assert _rel_err(out[k][:, :, ::stride], golden_full["out/" + k]) <= eng.tolerance(), k


def test_fused_engine_batch_consistency(tiny_net, golden_tiny):
    """A scene inside a batch gives the same result as the scene alone (scenes are independent units)."""
    eng = tiny_net.fused_engine()
    pts = torch.from_numpy(golden_tiny["points"]).cuda()
    both = eng.forward(pts)
    one = eng.forward(pts[1:2].contiguous())
    for k in HEADS:
        assert torch.equal(both[k][1:2], one[k]), k


def test_host_out_streams_identical_predictions():
    """forward(..., host_out=pinned dict): the results copied to the host head by head equal the returned tensors"""
    import bench
    net = bench.seeded_model().cuda()
    x = bench.synthetic_scenes(2, 1000)[:, :, :8192].contiguous().cuda()
    with torch.no_grad():
        ref = net({"scene_points": x})
        host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in ref.items()}
        got = net({"scene_points": x}, host_out=host)
        torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(got[k], ref[k]) and torch.equal(host[k], ref[k].cpu()), k


def test_host_out_accepts_bf16_buffers():
    """pinned result buffers in bf16: the streamed copies convert on the device (half the D2H bytes)"""
    import bench
    net = bench.seeded_model().cuda()
    x = bench.synthetic_scenes(2, 1000)[:, :, :8192].contiguous().cuda()
    with torch.no_grad():
        ref = net({"scene_points": x})
        host = {k: torch.empty(v.shape, dtype=torch.bfloat16).pin_memory() for k, v in ref.items()}
        net({"scene_points": x}, host_out=host)
        torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(host[k], ref[k].to(torch.bfloat16).cpu()), k


def test_side_stream_three_nn_is_bit_identical():
    """the 3-NN searches that run on a side stream beside the next level's sampling (engine.overlap_geometry) produce
    exactly the serial schedule's predictions, call after call"""
    import bench
    net = bench.seeded_model().cuda().eval()
    x = bench.synthetic_scenes(3, 1000)[:, :, :8192].contiguous().cuda()
    eng = net.fused_engine()
    assert eng.overlap_geometry
    with torch.no_grad():
        a = eng.forward(x)
        b = eng.forward(x)
        eng.overlap_geometry = False
        c = eng.forward(x)
        torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k


def test_tf32_tight_parity_backend(cloud_2638, golden_full):
    """mlp_backend="tf32" (csrc/linear_tf32.cu, one tcgen05 kind::tf32 launch per layer): the tight-parity mode of
    SURVEY.md §7 on BASELINE config 1.  Stated tolerance: head outputs within 4e-3 * max(1, max|ref|) of the fp32
    oracle golden (10-bit mantissa operands, 17 layers deep); geometry bit-exact."""
    from s4g_release_b200.engine import FusedPointNet2
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2
    from tests.golden.make_golden import seed_reference_weights
    torch.manual_seed(0)
    net = seed_reference_weights(PointNet2(**PN2_CLS_CONFIG)).cuda().eval()
    eng = FusedPointNet2(net, mlp_backend="tf32")
    out, trace = eng.forward(torch.from_numpy(cloud_2638)[None].cuda(), return_trace=True)
    torch.cuda.synchronize()
    for i in range(3):
        assert np.array_equal(trace["fps"][i].cpu().numpy(), golden_full[f"sa{i}/fps_index"])
    stride = int(golden_full["stride"])
    worst = max(_rel_err(out[k][:, :, ::stride], golden_full["out/" + k]) for k in HEADS)
    print("tf32 backend: worst head error %.3e" % worst)
    assert worst <= eng.tolerance() == 4e-3


def test_fp_linear_split_matches_the_concat_order(cloud_2638):
    """The finest propagation level runs its first conv on the sparse points and interpolates the pre-activations
    (engine.fp_linear_split; conv and interpolation commute).  Against the interpolate-then-conv order of the reference
    the difference is only where the bf16 rounding happens: <= 2e-2 * max|ref| on every head (measured ~3e-3)."""
    from s4g_release_b200.engine import FusedPointNet2
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2
    from tests.golden.make_golden import seed_reference_weights
    torch.manual_seed(0)
    net = seed_reference_weights(PointNet2(**PN2_CLS_CONFIG)).cuda().eval()
    x = torch.from_numpy(cloud_2638)[None].cuda()
    a = FusedPointNet2(net, fp_linear_split=True)
    b = FusedPointNet2(net, fp_linear_split=False)
    assert a.fp_pre[2] is not None and a.fp_pre[0] is None and all(p is None for p in b.fp_pre)
    oa, ob = a.forward(x), b.forward(x)
    torch.cuda.synchronize()
    for k in HEADS:
        assert _rel_err(oa[k], ob[k].float().cpu()) <= 2e-2, k
