"""s4g_release_b200 — B200-native (sm_100a) PointNet++ hot path of S4G behind the reference's
``pn2_ext`` operator boundary and ``network_models`` module surface.  See DESIGN.md."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is not built)

__all__ = ["_lib"]
