"""CPU tests of the MLP-chain planner (csrc/chain_plan.cu): every chain shape of PN2_CLS and of the GPU
numerics suite must get a plan that the planner's own simulation proves deadlock-free, fits in 227 KB of
shared memory, and keeps the job tables within the kernel-parameter budget.  No GPU is touched."""
import ctypes

import pytest

from s4g_release_b200._lib import lib

IN_ROWS, IN_GATHER = 0, 1
OUT_ROWS, OUT_MAXPOOL, OUT_LOGITS = 2, 3, 4


def _plan(dims, in_mode=IN_ROWS, feat_c=0, out_mode=OUT_ROWS, group=1):
    n = len(dims) - 1
    arr = lambda v: (ctypes.c_int * n)(*v)
    h = lib.s4g_chain_create(n, arr(dims[:-1]), arr(dims[1:]), arr([1] * n), in_mode, feat_c, out_mode, dims[-1], group, 0)
    if not h:
        return None, lib.s4g_last_error().decode()
    v = [ctypes.c_int() for _ in range(7)]
    assert lib.s4g_chain_info(h, *[ctypes.byref(x) for x in v]) == 0
    keys = ("n_jobs", "slots", "stages", "load_depth", "smem_bytes", "sim_cycles", "mma_cycles")
    info = {k: x.value for k, x in zip(keys, v)}
    buf = ctypes.create_string_buffer(1 << 16)
    lib.s4g_chain_describe(h, buf, len(buf))
    info["weight_bytes"] = lib.s4g_chain_weight_bytes(h)
    lib.s4g_chain_destroy(h)
    return info, buf.value.decode()


MODEL_CHAINS = [
    ([3, 128, 128, 256], IN_GATHER, 0, OUT_MAXPOOL, 64),
    ([259, 256, 256, 512], IN_GATHER, 256, OUT_MAXPOOL, 64),
    ([515, 512, 512, 1024], IN_GATHER, 512, OUT_MAXPOOL, 64),
    ([1536, 1024], IN_ROWS, 0, OUT_ROWS, 1),
    ([1024, 1024], IN_ROWS, 0, OUT_ROWS, 1),
    ([1280, 512, 512], IN_ROWS, 0, OUT_ROWS, 1),
    ([512, 256, 256, 256], IN_ROWS, 0, OUT_ROWS, 1),
    ([256, 512, 256, 256, 128, 3], IN_ROWS, 0, OUT_LOGITS, 1),
    ([256, 512, 256, 256, 128, 9], IN_ROWS, 0, OUT_LOGITS, 1),
]


@pytest.mark.parametrize("dims,in_mode,feat_c,out_mode,group", MODEL_CHAINS)
def test_model_chains_plan(dims, in_mode, feat_c, out_mode, group):
    info, text = _plan(dims, in_mode, feat_c, out_mode, group)
    assert info is not None, text
    assert info["smem_bytes"] <= 227 * 1024
    assert 2 <= info["slots"] <= 6 and 2 <= info["stages"] <= 8
    assert 1 <= info["load_depth"] <= 3
    # weights: every (cin_pad x cout_pad) element exactly once per tile
    cin_pad = [((c + 15) // 16) * 16 for c in dims[:-1]]
    if in_mode == IN_GATHER:
        cin_pad[0] = feat_c + 16
    cout_pad = [((c + 15) // 16) * 16 for c in dims[1:]]
    if out_mode == OUT_MAXPOOL:
        cout_pad[-1] = ((dims[-1] + 127) // 128) * 128
    if out_mode == OUT_LOGITS:
        cout_pad[-1] = 16
    passes = 2 if dims == [1536, 1024] or dims == [1024, 1024] else 1
    assert info["weight_bytes"] == 2 * sum(a * b for a, b in zip(cin_pad, cout_pad)), text.splitlines()[0]
    assert passes >= 1
    # planner estimate: wide chains keep the tensor pipe busy most of the tile
    if min(dims[:-1]) >= 256:
        assert info["mma_cycles"] >= 0.6 * info["sim_cycles"], text.splitlines()[0]


@pytest.mark.parametrize("dims", [[64, 64], [64, 128], [16, 32], [128, 256], [256, 512], [512, 256], [128, 96], [64, 64, 64],
                                  [32, 16, 32], [1280, 512], [512, 256, 256, 256]])
def test_small_row_chains_plan(dims):
    info, text = _plan(dims)
    assert info is not None, text


@pytest.mark.parametrize("feat_c,dims,group", [(64, [64, 64, 128], 16), (32, [32, 64], 8), (0, [16, 32], 32)])
def test_small_gather_chains_plan(feat_c, dims, group):
    info, text = _plan([feat_c + 3] + dims, IN_GATHER, feat_c, OUT_MAXPOOL, group)
    assert info is not None, text
    assert "LOAD_XYZ" in text and "EPI_MAXPOOL" in text


def test_rejects_unsupported_shapes():
    assert _plan([64, 1024, 64])[0] is None          # hidden activation wider than 512 channels
    assert _plan([67, 64], IN_GATHER, 60, OUT_MAXPOOL, 64)[0] is None  # feat_c not a multiple of 16
    assert _plan([64, 64], IN_ROWS, 0, OUT_MAXPOOL, 128)[0] is None    # group 128 unsupported
    assert _plan([64, 64, 40], IN_ROWS, 0, OUT_LOGITS)[0] is None      # > 16 logits


@pytest.mark.parametrize("dims", [[64, 32], [16, 32], [64, 64], [128, 96], [512, 256], [32, 32, 32], [1536, 1024]])
def test_never_more_slots_than_blocks_per_tile(dims):
    """Round-2 planner fix: a block must inherit its slot from the SAME or the PREVIOUS tile — with more slots than
    activation blocks per tile the loader's first slot wait is several barrier completions behind and passes on the
    wrong one-bit phase (a 64 -> 32 chain dead-locked from the 4th tile of a CTA on).  Every constraint set the tuner
    may pin must respect slots <= blocks per tile."""
    n = len(dims) - 1
    arr = lambda v: (ctypes.c_int * n)(*v)
    checked = 0
    for slots in (0, 1, 2, 3, 4, 5, 6):
        for subs in (1, 2):
            for tma in (0, 1):
                h = lib.s4g_chain_create_tuned_in(n, arr(dims[:-1]), arr(dims[1:]), arr([1] * n), IN_ROWS, 0, OUT_ROWS, dims[-1],
                                                  1, 0, slots, -1, -1, subs, tma)
                if not h:
                    continue
                buf = ctypes.create_string_buffer(1 << 16)
                lib.s4g_chain_describe(h, buf, len(buf))
                lib.s4g_chain_destroy(h)
                head = dict(kv.split("=") for kv in buf.value.decode().splitlines()[0].split() if "=" in kv)
                assert int(head["slots"]) <= int(head["n_act"]), (slots, subs, tma, head)
                checked += 1
    assert checked >= 2
