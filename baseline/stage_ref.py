"""Stage the UNMODIFIED python side of the reference's hot path under ``baseline/_ref/`` (git-ignored, NOT
gpurun-ignored: it travels to the GPU box exactly like ``oracle/_ref/``).  Run in the build container, where
/root/reference exists; ``__graft_entry__.build()`` calls it.  The reference is not pip-installable (no setup.py /
pyproject at its root — only the two in-place extension builds, PN2U/setup.py and NM/functions/setup.py), so "install"
here means: the package files of ``grasp_proposal.network_models`` copied byte for byte, a manifest with their
sha256, and the reference's own CUDA extension compiled by oracle/build_ref.py (``oracle/_ref/ref_pn2_ext.so``).

Nothing here is product code and nothing under ``baseline/_ref`` is committed.  Users:
  * tests/test_reference_dropin_gpu.py — the Level-1 drop-in test: the reference's PointNet2 (its own modules.py /
    functions.py / nn_utils) on (i) its own CUDA ops and (ii) this repo's ``pn2_ext``;
  * bench.py ``reference_cuda`` — the reference model + reference CUDA kernels timed on the B200.
"""
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/inference/grasp_proposal"
DST = os.path.join(HERE, "_ref", "grasp_proposal")
PN2_EXT = "grasp_proposal.network_models.models.pointnet2_utils.pn2_ext"
FILES = [
    "__init__.py",
    "network_models/__init__.py",
    "network_models/models/__init__.py",
    "network_models/models/PointNet2_tcls.py",
    "network_models/models/pointnet2_utils/__init__.py",
    "network_models/models/pointnet2_utils/functions.py",
    "network_models/models/pointnet2_utils/modules.py",
    "network_models/nn_utils/__init__.py",
    "network_models/nn_utils/conv.py",
    "network_models/nn_utils/functional.py",
    "network_models/nn_utils/init.py",
    "network_models/nn_utils/linear.py",
    "network_models/nn_utils/mlp.py",
    "network_models/functions/__init__.py",
    "network_models/functions/functions.py",
    "network_models/functions/gather_knn.py",
]


def stage(verbose=False):
    """Copies the files when the reference tree is present; returns the staged root or None."""
    if not os.path.isdir(SRC):
        return root() if available() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("baseline/_ref: staged %d reference files" % len(FILES))
    return root()


def root():
    return os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(DST, "network_models", "models", "PointNet2_tcls.py"))


def import_reference_model(pn2_ext_module):
    """Imports the staged reference package with ``pn2_ext_module`` standing in for its compiled extension (the
    reference does ``from . import pn2_ext`` unconditionally, PN2U/functions.py:2) and returns its PointNet2 class.
    Re-importable: a second call with another extension module swaps the ops under the SAME reference classes."""
    if not available():
        raise ImportError("baseline/_ref is not staged (run python baseline/stage_ref.py in the build container)")
    if root() not in sys.path:
        sys.path.insert(0, root())
    shim = types.ModuleType(PN2_EXT)
    for fn in ("farthest_point_sample", "ball_query", "group_points_forward", "group_points_backward",
               "point_search", "interpolate_forward", "interpolate_backward"):
        setattr(shim, fn, getattr(pn2_ext_module, fn))
    sys.modules[PN2_EXT] = shim
    import contextlib
    # (the reference prints "Please compile source files ..." to stdout when its optional dgcnn_ext is absent,
    # functions/gather_knn.py:4-7 — keep stdout clean for callers that print one JSON line)
    with contextlib.redirect_stdout(sys.stderr):
        import grasp_proposal.network_models.models.pointnet2_utils as pkg
        pkg.pn2_ext = shim
        fmod = sys.modules.get("grasp_proposal.network_models.models.pointnet2_utils.functions")
        if fmod is not None:  # already imported with another extension: rebind the name its functions look up
            fmod.pn2_ext = shim
        from grasp_proposal.network_models.models.PointNet2_tcls import PointNet2
    return PointNet2


if __name__ == "__main__":
    print(stage(verbose=True))
