"""rbdreference_b200 - B200-native batched rigid-body dynamics (rnea / rnea_grad / minv).

Drop-in for the hot path of A2R-Lab/RBDReference: `RBDReference(robot).rnea(...)`,
`.rnea_grad(...)`, `.minv(...)` and the per-pass helpers, evaluated over batches of knot
points by hand-written sm_100a CUDA kernels behind the C ABI in include/rbd_b200.h.
"""
from .model import RobotModel, compile_model, MAX_DOF  # noqa: F401
from . import robots  # noqa: F401


def __getattr__(name):
    # engine (and with it the CUDA library) is loaded on first use so that the host-only
    # pieces (model compiler, robots) import without the built .so
    if name == "RBDReference":
        from .engine import RBDReference
        return RBDReference
    if name in ("shard_bounds", "gather_to_all", "gather_to_rank"):
        from . import dist
        return getattr(dist, name)
    raise AttributeError(name)


__all__ = ["RBDReference", "RobotModel", "compile_model", "robots", "MAX_DOF"]
