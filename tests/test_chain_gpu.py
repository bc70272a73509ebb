"""GPU numerics suite for the tcgen05 MLP-chain kernel (csrc/mlp_chain.cu) against a plain torch fp32
reference of the same op on the same bf16-rounded operands.  Tolerances are stated per test: the kernel
multiplies bf16 operands exactly and accumulates in fp32, so against a reference that (a) uses the same
bf16-rounded weights/inputs and (b) rounds hidden activations to bf16 at the same places, the only
difference is fp32 summation order:  |err| <= 2e-2 * max|ref| is a loose bound, 1 bf16 ulp = 2^-8."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _ref_chain(x, layers, round_hidden=True):
    """x fp32 [P, Cin] (already bf16-representable); layers [(W, b, relu)] fp32."""
    for i, (w, b, relu) in enumerate(layers):
        x = x @ _bf(w).t() + b
        if relu:
            x = torch.relu(x)
        if round_hidden and i + 1 < len(layers):
            x = _bf(x)
    return x


def _layers(dims, seed, relu_last=True, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(len(dims) - 1):
        w = torch.randn(dims[i + 1], dims[i], generator=g) * (scale / np.sqrt(dims[i]))
        b = torch.randn(dims[i + 1], generator=g) * 0.1
        out.append((w, b, relu_last or i + 2 < len(dims)))
    return out


def _check(got, want, tol=2e-2):
    err = (got.float().cpu() - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= tol * max(ref, 1e-3), "max abs err %.4e vs max|ref| %.4e" % (err, ref)


@pytest.mark.parametrize("dims,P", [
    ([64, 64], 128), ([64, 128], 128), ([16, 32], 128), ([128, 256], 256), ([256, 512], 384), ([512, 256], 1024),
    ([64, 64], 100), ([64, 64], 1), ([128, 96], 333), ([1536, 1024], 512), ([1024, 1024], 640), ([1280, 512], 256),
])
def test_single_layer_rows(dims, P):
    from s4g_release_b200.chain import MlpChain
    layers = _layers(dims, seed=sum(dims))
    x = _bf(torch.randn(P, dims[0], generator=torch.Generator().manual_seed(P)))
    ch = MlpChain(layers, "cuda")
    got = ch.run_rows(x.cuda().to(torch.bfloat16))
    torch.cuda.synchronize()
    _check(got, _bf(_ref_chain(x, layers)))


@pytest.mark.parametrize("dims,P", [
    ([64, 64, 64], 256), ([512, 256, 256, 256], 1000), ([1280, 512, 512], 512), ([256, 512, 256, 256, 128], 2048),
    ([32, 16, 32], 77),
])
def test_multi_layer_rows(dims, P):
    from s4g_release_b200.chain import MlpChain
    layers = _layers(dims, seed=sum(dims) + 1)
    x = _bf(torch.randn(P, dims[0], generator=torch.Generator().manual_seed(P)))
    got = MlpChain(layers, "cuda").run_rows(x.cuda().to(torch.bfloat16))
    torch.cuda.synchronize()
    _check(got, _bf(_ref_chain(x, layers)))


@pytest.mark.parametrize("dims,n_out,B,n_points,sigmoid", [([256, 512, 256, 256, 128], 3, 2, 1280, False),
                                                           ([64, 32], 9, 3, 100, False), ([64, 32], 5, 1, 4096, True)])
def test_logits_head(dims, n_out, B, n_points, sigmoid):
    from s4g_release_b200.chain import OUT_LOGITS, MlpChain
    layers = _layers(dims + [n_out], seed=n_out, relu_last=False)
    P = B * n_points
    x = _bf(torch.randn(P, dims[0], generator=torch.Generator().manual_seed(P)))
    got = MlpChain(layers, "cuda", out_mode=OUT_LOGITS, sigmoid=sigmoid).run_rows(x.cuda().to(torch.bfloat16), n_points)
    torch.cuda.synchronize()
    want = _ref_chain(x, layers).reshape(B, n_points, n_out).transpose(1, 2)
    if sigmoid:
        want = torch.sigmoid(want)
    assert tuple(got.shape) == (B, n_out, n_points) and got.dtype == torch.float32
    _check(got, want)


@pytest.mark.parametrize("xyz_cuda", [False, True])
@pytest.mark.parametrize("feat_c,dims,B,N,M,K", [(0, [128, 128, 256], 2, 2048, 256, 64), (64, [64, 64, 128], 2, 1024, 64, 16),
                                                  (256, [256, 256, 512], 1, 2048, 128, 64), (512, [512, 512, 1024], 1, 512, 64, 64),
                                                  (32, [32, 64], 3, 300, 50, 8), (0, [16, 32], 1, 100, 7, 32),
                                                  (0, [64, 128], 2, 700, 33, 16), (0, [256, 128], 1, 400, 20, 8), (0, [48], 1, 90, 9, 8)])
def test_set_abstraction_gather_maxpool(feat_c, dims, B, N, M, K, xyz_cuda):
    """gather + centroid subtraction + concat + chain + max over K, vs torch fp32 on the same operands."""
    from s4g_release_b200.chain import IN_GATHER, OUT_MAXPOOL, MlpChain
    g = torch.Generator().manual_seed(feat_c + M)
    layers = _layers([feat_c + 3] + dims, seed=M, scale=2.0)
    xyz = torch.rand(B, 3, N, generator=g)
    sel = torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(B)])
    ctr = torch.gather(xyz, 2, sel.unsqueeze(1).expand(-1, 3, -1)).contiguous()
    nbr = torch.randint(0, N, (B, M, K), generator=g, dtype=torch.int32)
    feat = _bf(torch.randn(B * N, feat_c, generator=g)) if feat_c else None
    if xyz_cuda and feat_c:
        pytest.skip("the CUDA-core xyz layer only exists for chains without gathered features")
    ch = MlpChain(layers, "cuda", IN_GATHER, feat_c, OUT_MAXPOOL, group=K, xyz_layer_on_cuda_cores=xyz_cuda)
    got = ch.run_gather(feat.cuda().to(torch.bfloat16) if feat_c else None, xyz.cuda(), ctr.cuda(), nbr.cuda())
    torch.cuda.synchronize()
    # reference: reference channel order [rel_xyz | features] (modules.py:48)
    idx = nbr.long().reshape(B, M * K)
    gx = torch.gather(xyz.transpose(1, 2), 1, idx.unsqueeze(-1).expand(-1, -1, 3)).reshape(B, M, K, 3)
    rel32 = gx - ctr.transpose(1, 2).unsqueeze(2)
    rel = _bf(rel32)
    x = rel
    if feat_c:
        gf = torch.gather(feat.reshape(B, N, feat_c), 1, idx.unsqueeze(-1).expand(-1, -1, feat_c)).reshape(B, M, K, feat_c)
        x = torch.cat([rel, gf], dim=-1)
    if ch.in_mode == 5:
        # IN_XYZ_MLP: the 3 -> C layer runs in fp32 on the CUDA cores (unrounded coordinates and weights)
        w0, b0, _ = layers[0]
        h = _bf(torch.relu(rel32.reshape(-1, 3) @ w0.t() + b0))
        want = _ref_chain(h, layers[1:]).reshape(B * M, K, -1).max(dim=1)[0]
    else:
        want = _ref_chain(x.reshape(-1, feat_c + 3), layers).reshape(B * M, K, -1).max(dim=1)[0]
    assert tuple(got.shape) == (B * M, dims[-1])
    _check(got, _bf(want))


@pytest.mark.parametrize("Nq,Nk", [(700, 90), (3000, 1200)])
def test_fp_front_end(Nq, Nk):
    from s4g_release_b200.engine import FusedPointNet2 as E
    from oracle import pn2_ext_cpu as ora
    g = torch.Generator().manual_seed(4)
    B, C2, C1 = 2, 64, 32
    q = torch.rand(B, 3, Nq, generator=g)
    k = torch.rand(B, 3, Nk, generator=g)
    idx, w = E.three_nn_weights(q.cuda(), k.cuda())
    o_idx, o_d = ora.point_search(q, k, 3)
    assert torch.equal(idx.cpu().long(), o_idx)
    inv = 1.0 / torch.clamp(o_d, min=1e-10)
    np.testing.assert_allclose(w.cpu().numpy(), (inv / inv.sum(2, keepdim=True)).numpy(), rtol=1e-6)
    sparse = _bf(torch.randn(B * Nk, C2, generator=g))
    dense = _bf(torch.randn(B * Nq, C1, generator=g))
    out = E.interp_concat(sparse.cuda().to(torch.bfloat16), idx, w, dense.cuda().to(torch.bfloat16), B, Nk, Nq)
    want = ora.interpolate_forward(sparse.reshape(B, Nk, C2).transpose(1, 2).contiguous(), o_idx, w.cpu())
    _check(out[:, :C2], _bf(want.transpose(1, 2).reshape(B * Nq, C2)), tol=1e-2)
    assert torch.equal(out[:, C2:].float().cpu(), dense)


def _two_block_chain(*args, **kw):
    """MlpChain with two row blocks per tile, or None when no plan fits (wide chains)."""
    from s4g_release_b200.chain import MlpChain
    try:
        return MlpChain(*args, subs=2, **kw)
    except RuntimeError:
        return None


@pytest.mark.parametrize("dims,P", [([64, 64, 64], 256), ([64, 64, 64], 700), ([32, 16, 32], 77), ([128, 128, 128, 64], 1300),
                                    ([64, 64], 129), ([256, 128, 128], 513)])
def test_two_row_blocks_per_tile_rows(dims, P):
    """tiles of 2 x 128 rows, the two blocks interleaved layer by layer (WorkerJob::sub): same results, ragged tails"""
    layers = _layers(dims, seed=sum(dims) + 3)
    x = _bf(torch.randn(P, dims[0], generator=torch.Generator().manual_seed(P)))
    ch = _two_block_chain(layers, "cuda")
    assert ch is not None
    got = ch.run_rows(x.cuda().to(torch.bfloat16))
    torch.cuda.synchronize()
    _check(got, _bf(_ref_chain(x, layers)))


@pytest.mark.parametrize("feat_c,dims,B,N,M,K", [(0, [128, 128, 256], 2, 2048, 256, 64), (0, [128, 128, 256], 1, 3000, 333, 64),
                                                  (64, [64, 64, 128], 2, 1024, 65, 16), (32, [32, 64], 3, 300, 50, 8)])
def test_two_row_blocks_per_tile_gather_maxpool(feat_c, dims, B, N, M, K):
    from s4g_release_b200.chain import IN_GATHER, OUT_MAXPOOL
    g = torch.Generator().manual_seed(feat_c + M)
    layers = _layers([feat_c + 3] + dims, seed=M, scale=2.0)
    xyz = torch.rand(B, 3, N, generator=g)
    sel = torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(B)])
    ctr = torch.gather(xyz, 2, sel.unsqueeze(1).expand(-1, 3, -1)).contiguous()
    nbr = torch.randint(0, N, (B, M, K), generator=g, dtype=torch.int32)
    feat = _bf(torch.randn(B * N, feat_c, generator=g)) if feat_c else None
    ch = _two_block_chain(layers, "cuda", IN_GATHER, feat_c, OUT_MAXPOOL, group=K)
    assert ch is not None
    got = ch.run_gather(feat.cuda().to(torch.bfloat16) if feat_c else None, xyz.cuda(), ctr.cuda(), nbr.cuda())
    torch.cuda.synchronize()
    idx = nbr.long().reshape(B, M * K)
    gx = torch.gather(xyz.transpose(1, 2), 1, idx.unsqueeze(-1).expand(-1, -1, 3)).reshape(B, M, K, 3)
    x = _bf(gx - ctr.transpose(1, 2).unsqueeze(2))
    if feat_c:
        gf = torch.gather(feat.reshape(B, N, feat_c), 1, idx.unsqueeze(-1).expand(-1, -1, feat_c)).reshape(B, M, K, feat_c)
        x = torch.cat([x, gf], dim=-1)
    want = _ref_chain(x.reshape(-1, feat_c + 3), layers).reshape(B * M, K, -1).max(dim=1)[0]
    _check(got, _bf(want))


def test_two_row_blocks_logits_head():
    from s4g_release_b200.chain import IN_ROWS, OUT_LOGITS
    dims, n_out, B, n_points = [64, 32], 5, 3, 1000
    layers = _layers(dims + [n_out], seed=5, relu_last=False)
    x = _bf(torch.randn(B * n_points, dims[0], generator=torch.Generator().manual_seed(1)))
    ch = _two_block_chain(layers, "cuda", IN_ROWS, 0, OUT_LOGITS, sigmoid=True)
    assert ch is not None
    got = ch.run_rows(x.cuda().to(torch.bfloat16), n_points=n_points)
    torch.cuda.synchronize()
    want = torch.sigmoid(_ref_chain(x, layers)).reshape(B, n_points, n_out).transpose(1, 2)
    _check(got, want)


# ---- input blocks fetched by TMA tensor copies into the 128-byte-swizzled layout (s4g_chain_create_tuned_in) ----
@pytest.mark.parametrize("dims,P,stride_pad,subs", [
    ([64, 64], 128, 0, 1), ([64, 64, 64], 700, 0, 1), ([128, 256], 333, 0, 1), ([256, 512, 256, 256, 128], 2048, 0, 1),
    ([192, 128, 64], 1000, 0, 1), ([384, 256, 256], 20000, 0, 1), ([512, 256, 256, 256], 1000, 64, 1),
    ([1280, 512, 512], 513, 0, 1), ([64, 64, 64], 1300, 0, 2), ([128, 128, 128, 64], 77, 8, 2),
])
def test_tma_input_rows(dims, P, stride_pad, subs):
    from s4g_release_b200.chain import MlpChain
    layers = _layers(dims, seed=sum(dims) + 7)
    x = _bf(torch.randn(P, dims[0], generator=torch.Generator().manual_seed(P)))
    try:
        ch = MlpChain(layers, "cuda", subs=subs, tma_in=1)
    except RuntimeError:
        assert subs == 2  # the planner may refuse two row blocks per tile; one block must always plan
        pytest.skip("no two-row-block plan for this chain")
    xd = torch.zeros(P, dims[0] + stride_pad, dtype=torch.bfloat16, device="cuda")
    xd[:, :dims[0]] = x.cuda().to(torch.bfloat16)
    xd[:, dims[0]:] = 7.0  # channels past cin belong to someone else: they must not leak in
    got = ch.run_rows(xd)
    torch.cuda.synchronize()
    _check(got, _bf(_ref_chain(x, layers)))
    same = MlpChain(layers, "cuda", subs=subs).run_rows(xd)  # same plan, cp.async input: bit-identical
    assert torch.equal(got, same)


def test_tma_input_logits_head():
    from s4g_release_b200.chain import OUT_LOGITS, MlpChain
    dims, n_out, B, n_points = [256, 512, 256, 256, 128], 3, 2, 1280
    layers = _layers(dims + [n_out], seed=n_out, relu_last=False)
    x = _bf(torch.randn(B * n_points, dims[0], generator=torch.Generator().manual_seed(3)))
    got = MlpChain(layers, "cuda", out_mode=OUT_LOGITS, tma_in=1).run_rows(x.cuda().to(torch.bfloat16), n_points)
    torch.cuda.synchronize()
    want = _ref_chain(x, layers).reshape(B, n_points, n_out).transpose(1, 2)
    _check(got, want)


def test_tma_input_refused_where_it_cannot_apply():
    from s4g_release_b200.chain import IN_GATHER, OUT_MAXPOOL, MlpChain
    with pytest.raises(RuntimeError):
        MlpChain(_layers([32, 64], seed=1), "cuda", tma_in=1)  # 32 input channels: not a whole 128-byte row
    with pytest.raises(RuntimeError):
        MlpChain(_layers([16, 64, 64], seed=1), "cuda", in_mode=IN_GATHER, feat_c=0, out_mode=OUT_MAXPOOL, group=16, tma_in=1)


@pytest.mark.parametrize("subs", [1, 2])
@pytest.mark.parametrize("feat_c,dims,B,N,M,K", [(64, [64, 64, 128], 2, 1024, 64, 16), (256, [256, 256, 512], 1, 2048, 128, 64),
                                                  (512, [512, 512, 1024], 1, 512, 64, 64), (128, [128, 128], 3, 300, 50, 8),
                                                  (192, [64, 64], 2, 500, 33, 8), (256, [256, 256, 512], 3, 4096, 1000, 32)])
def test_tma_gather4_input_maxpool(feat_c, dims, B, N, M, K, subs):
    """Neighbour feature rows fetched with tile::gather4 copies (tma_in=1): bit-identical to the cp.async path (which
    test_set_abstraction_gather_maxpool checks against torch), including ragged last tiles."""
    from s4g_release_b200.chain import IN_GATHER, OUT_MAXPOOL, MlpChain
    g = torch.Generator().manual_seed(feat_c + M)
    layers = _layers([feat_c + 3] + dims, seed=M, scale=2.0)
    xyz = torch.rand(B, 3, N, generator=g).cuda()
    ctr = xyz[:, :, :M].contiguous()
    nbr = torch.randint(0, N, (B, M, K), generator=g, dtype=torch.int32).cuda()
    feat = torch.randn(B * N, feat_c, generator=g).cuda().to(torch.bfloat16)
    try:
        ch = MlpChain(layers, "cuda", IN_GATHER, feat_c, OUT_MAXPOOL, group=K, subs=subs, tma_in=1)
        ref = MlpChain(layers, "cuda", IN_GATHER, feat_c, OUT_MAXPOOL, group=K, subs=subs)
    except RuntimeError:
        assert subs == 2
        pytest.skip("no two-row-block plan for this chain")
    got = ch.run_gather(feat, xyz, ctr, nbr)
    want = ref.run_gather(feat, xyz, ctr, nbr)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all() and got.float().abs().max() > 0
    assert torch.equal(got, want)


@pytest.mark.parametrize("kind", ["rows", "logits", "gather", "gather_feat", "rows_linear"])
def test_every_tunable_plan_is_bit_identical(kind):
    """Every plan constraint set the engine's tuner may pin (engine.candidate_plans: ring sizes, accumulator pairing,
    cooperative epilogues, one / two row blocks per tile, cp.async / TMA input) yields the SAME bits: each accumulator
    sums its K-blocks in the same order in all of them, so the tuned choice can never change a result."""
    from s4g_release_b200.chain import IN_GATHER, IN_ROWS, OUT_LOGITS, OUT_MAXPOOL, OUT_ROWS, MlpChain
    from s4g_release_b200.engine import candidate_plans
    g = torch.Generator().manual_seed(5)
    if kind in ("rows", "logits", "rows_linear"):
        dims = {"rows": [512, 256, 256, 256], "logits": [256, 512, 256, 256, 128, 9], "rows_linear": [512, 256]}[kind]
        in_mode, feat_c, group = IN_ROWS, 0, 1
        out_mode = OUT_LOGITS if kind == "logits" else OUT_ROWS
        layers = _layers(dims, seed=3, relu_last=(kind == "rows"))
        P = 128 * 37 + 61
        x = torch.randn(P, dims[0], generator=g).cuda().to(torch.bfloat16)
        run = lambda ch: ch.run_rows(x, n_points=P if kind == "logits" else 0)
    else:
        feat_c = 256 if kind == "gather_feat" else 0
        dims = [256, 256, 512] if feat_c else [128, 128, 256]
        in_mode, out_mode, group = IN_GATHER, OUT_MAXPOOL, 64
        layers = _layers([feat_c + 3] + dims, seed=4, scale=2.0)
        B, N, M = 2, 4096, 300
        xyz = torch.rand(B, 3, N, generator=g).cuda()
        ctr = xyz[:, :, :M].contiguous()
        nbr = torch.randint(0, N, (B, M, group), generator=g, dtype=torch.int32).cuda()
        feat = torch.randn(B * N, feat_c, generator=g).cuda().to(torch.bfloat16) if feat_c else None
        run = lambda ch: ch.run_gather(feat, xyz, ctr, nbr)
    want, seen = None, set()
    for slots, pairs, coop, subs, tma in candidate_plans(in_mode, feat_c, out_mode):
        try:
            ch = MlpChain(layers, "cuda", in_mode, feat_c, out_mode, group=group, slots=slots, pairs=pairs, coop=coop,
                          subs=subs, tma_in=tma)
        except RuntimeError:
            continue  # the planner has no deadlock-free plan under these constraints
        key = (ch.describe(), tma)
        if key in seen:
            continue
        seen.add(key)
        got = run(ch)
        torch.cuda.synchronize()
        if want is None:
            want = got
        assert torch.equal(got, want), "plan %r differs from the planner's default" % ((slots, pairs, coop, subs, tma),)
    assert len(seen) >= 4, "only %d distinct plans were exercised" % len(seen)


@pytest.mark.parametrize("P,N,K,relu", [(128, 128, 32, True), (300, 70, 45, True), (1000, 260, 515, False),
                                        (4096, 512, 256, True), (77, 9, 128, False), (129, 1024, 1536, True)])
def test_linear_tf32_layer(P, N, K, relu):
    """csrc/linear_tf32.cu (tcgen05 kind::tf32) against fp64 on the SAME TF32-rounded operands: what is left is fp32
    accumulation order, |err| <= 2e-5 * sum|x||w| scale (stated as 1e-4 * max|ref|)."""
    from s4g_release_b200.engine import FusedPointNet2 as E
    g = torch.Generator().manual_seed(P + N + K)
    x = E._round_tf32(torch.randn(P, K, generator=g))
    w = E._round_tf32(torch.randn(N, K, generator=g) / np.sqrt(K))
    b = torch.randn(N, generator=g) * 0.1
    K4 = (K + 3) // 4 * 4
    wp = torch.zeros(N, K4)
    wp[:, :K] = w
    got = E.linear_tf32(x.cuda(), wp.cuda(), b.cuda(), relu=relu, round_out=False)
    torch.cuda.synchronize()
    want = x.double() @ w.double().t() + b.double()
    if relu:
        want = torch.relu(want)
    err = (got.double().cpu() - want).abs().max().item()
    assert tuple(got.shape) == (P, N) and err <= 1e-4 * max(1.0, want.abs().max().item()), err
    # round_out: the stored value is the nearest TF32 of the same result
    got_r = E.linear_tf32(x.cuda(), wp.cuda(), b.cuda(), relu=relu, round_out=True)
    assert torch.equal(got_r.cpu(), E._round_tf32(got.cpu()))


@pytest.mark.parametrize("dims", [[64, 32], [16, 32], [64, 64], [32, 32, 32]])
def test_many_tiles_of_a_one_block_chain(dims):
    """Regression (round 2): chains whose tile is a single activation block ran into a slot-phase ambiguity from the 4th
    tile of a CTA on (mbarrier dead-lock -> trap) under plans with more slots than blocks.  40 tiles per CTA, every plan
    the tuner may try, result checked against torch."""
    from s4g_release_b200.chain import IN_ROWS, OUT_ROWS, MlpChain
    from s4g_release_b200.engine import candidate_plans
    layers = _layers(dims, seed=11)
    P = 148 * 40 * 128 + 77
    x = _bf(torch.randn(P, dims[0], generator=torch.Generator().manual_seed(3)))
    want = _bf(_ref_chain(x[:4096], layers))
    xd = x.cuda().to(torch.bfloat16)
    seen = set()
    for slots, pairs, coop, subs, tma in candidate_plans(IN_ROWS, 0, OUT_ROWS):
        try:
            ch = MlpChain(layers, "cuda", IN_ROWS, 0, OUT_ROWS, slots=slots, pairs=pairs, coop=coop, subs=subs, tma_in=tma)
        except RuntimeError:
            continue
        key = (ch.describe(), tma)
        if key in seen:
            continue
        seen.add(key)
        for _ in range(3):
            got = ch.run_rows(xd)
        torch.cuda.synchronize()
        _check(got[:4096], want)
    assert seen
