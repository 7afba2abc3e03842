"""`RBDReference(robot).ALGORITHM(...)` - drop-in host API over the sm_100a kernels.

Same method names, positional order and defaults as /root/reference/RBDReference.py
(README.md:9-17), for the hot path: rnea, rnea_grad, minv and their eight per-pass helpers.

Accepted inputs (every array argument of one call must be of the same kind):

* reference shapes - numpy `(n,)` vectors (and `(6,NB)`, `(6,n,NB)`, ... intermediates):
  one knot point, results returned as fresh numpy arrays of the reference's shapes, except the
  documented in-place cases which update and return the caller's array (`rnea_bpass` f,
  `minv_fpass` Minv/F, `rnea_grad_bpass_*` df), exactly like the reference;
* batched - a leading batch axis `B` on every argument: numpy arrays (copied to the device and
  back) or torch tensors (CUDA tensors are used in place, outputs stay on the device).
  `out[k]` equals the reference called on element k.

All arithmetic runs in hand-written CUDA kernels through the C ABI in include/rbd_b200.h;
PyTorch only owns device memory and streams.  There is no CPU fallback: without the built
library or without a CUDA device the calls raise.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _capi
from .hostpipe import HostPipeline, pinned_empty
from .model import FbModel, RobotModel, compile_ee_model, compile_fb_model, compile_model, select_end_effector_joints

__all__ = ["RBDReference"]

_SUFFIX = {torch.float64: "f64", torch.float32: "f32"}


class RBDReference:
    """B200-native counterpart of the reference class (RBDReference.py:5-7)."""

    def __init__(self, robotObj, dtype=torch.float64, device=None):
        if isinstance(dtype, str):
            dtype = {"float64": torch.float64, "f64": torch.float64, "fp64": torch.float64,
                     "float32": torch.float32, "f32": torch.float32, "fp32": torch.float32}[dtype]
        if dtype not in _SUFFIX:
            raise ValueError("dtype must be torch.float64 or torch.float32")
        self.robot = robotObj
        self.dtype = dtype
        self._suffix = _SUFFIX[dtype]
        self._np_dtype = np.float64 if dtype == torch.float64 else np.float32
        self._lib = _capi.load_library()            # raises if the CUDA library is missing
        self.floating_base = bool(getattr(robotObj, "floating_base", False)) or isinstance(robotObj, FbModel)
        if self.floating_base:
            # rnea / rnea_grad / minv of the reference's floating-base branches
            self.model = robotObj if isinstance(robotObj, FbModel) else compile_fb_model(robotObj)
            self.NB, self.n, self.nq = self.model.NB, self.model.nv, self.model.nq
            self._handle = _capi.FbModelHandle(self.model)
        else:
            self.model = robotObj if isinstance(robotObj, RobotModel) else compile_model(robotObj)
            self.n = self.nq = self.NB = self.model.n
            self._handle = _capi.ModelHandle(self.model)
        self._fb = "fb_" if self.floating_base else ""       # C-ABI prefix of the per-pass helpers
        self._device = torch.device(device) if device is not None else None
        if self._device is not None and self._device.type == "cuda" and torch.cuda.is_available():
            # the scratch pool of an explicitly requested device exists before any (possibly captured) call needs it
            idx = self._device.index if self._device.index is not None else torch.cuda.current_device()
            _capi.check(self._lib.rbd_prepare_device(int(idx)), "rbd_prepare_device")
        self._ee_handles = {}                       # (names, offset) -> compiled end-effector handle
        self._pipes = {}                            # device -> HostPipeline (streams + staging buffers)

    # ------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------
    def _default_device(self) -> torch.device:
        if self._device is not None:
            return self._device
        if not torch.cuda.is_available():
            raise RuntimeError("rbdreference_b200: no CUDA device available - the hot path has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    class _Ctx:
        """Normalises one call's arguments to contiguous device tensors with a batch axis."""

        def __init__(self, eng: "RBDReference", lead, base_ndim: int):
            self.eng = eng
            self.kind = "torch" if isinstance(lead, torch.Tensor) else "numpy"
            arr_ndim = lead.dim() if self.kind == "torch" else np.ndim(lead)
            if arr_ndim == base_ndim:
                self.batched = False
            elif arr_ndim == base_ndim + 1:
                self.batched = True
            else:
                raise ValueError("expected an array with %d or %d dimensions" % (base_ndim, base_ndim + 1))
            if self.kind == "torch" and lead.is_cuda:
                self.device = lead.device
            else:
                self.device = eng._default_device()
            self.in_device = lead.device if self.kind == "torch" else None
            self.B: Optional[int] = None

        def dev(self, x, shape_tail, name):
            """-> contiguous (B, *shape_tail) tensor of the engine dtype on the compute device."""
            if x is None:
                return None
            eng = self.eng
            if isinstance(x, torch.Tensor):
                t = x
            else:
                t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=eng._np_dtype))
            if not self.batched:
                t = t.unsqueeze(0)
            if tuple(t.shape[1:]) != tuple(shape_tail):
                raise ValueError("%s: expected trailing shape %s, got %s" % (name, tuple(shape_tail), tuple(t.shape[1:])))
            if self.B is None:
                self.B = int(t.shape[0])
            elif int(t.shape[0]) != self.B:
                raise ValueError("%s: batch size %d does not match %d" % (name, t.shape[0], self.B))
            return t.to(device=self.device, dtype=eng.dtype).contiguous()

        def empty(self, *shape_tail):
            return torch.empty((self.B,) + tuple(shape_tail), dtype=self.eng.dtype, device=self.device)

        def ret(self, t):
            """device tensor -> what the caller expects (numpy / torch, batched or not)."""
            if t is None:
                return None
            if not self.batched:
                t = t[0]
            if self.kind == "numpy":
                return t.cpu().numpy()
            if self.in_device is not None and self.in_device != t.device:
                return t.to(self.in_device)
            return t

        def write_back(self, target, t):
            """In-place contract: put device result `t` into the caller's array and return it."""
            if not self.batched:
                t = t[0]
            if isinstance(target, torch.Tensor):
                if target.data_ptr() != t.data_ptr():
                    target.copy_(t)
                return target
            np.copyto(target, t.cpu().numpy().astype(target.dtype, copy=False))
            return target

    def _fixed_only(self, what: str):
        if self.floating_base:
            raise NotImplementedError(
                "%s: no floating-base path (rnea, rnea_grad, minv, their eight per-pass helpers, forward_dynamics and "
                "forward_dynamics_grad are the floating-base entry points, SURVEY.md 8f rank 3)" % what)

    # ---- caller-supplied result buffers ---------------------------------------------------------
    def _check_out(self, t, ctx: "_Ctx", shape, name: str):
        """`out=` / `c_out=` of a batched CUDA-tensor call: the kernels write through the raw pointer, so
        anything but a contiguous tensor of the engine dtype, exact shape and compute device is refused."""
        if t is None:
            return None
        if ctx.kind != "torch" or not ctx.batched:
            raise ValueError("%s: a torch result buffer needs batched torch inputs (numpy calls take a numpy `out`)" % name)
        if not isinstance(t, torch.Tensor):
            raise ValueError("%s must be a torch tensor" % name)
        if not t.is_cuda or t.device != ctx.device:
            raise ValueError("%s must live on the compute device %s, got %s" % (name, ctx.device, t.device))
        if t.dtype != self.dtype:
            raise ValueError("%s must have dtype %s, got %s" % (name, self.dtype, t.dtype))
        if tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)
        return t

    # ---- host-buffer path (numpy in, numpy out): chunked, pinned, three streams -------------------
    @staticmethod
    def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
        """numpy array in page-locked host memory: the DMA engines read / write it directly, so passing
        such arrays (inputs and `out=`) to the batched numpy calls skips every host-side copy."""
        return pinned_empty(shape, dtype)

    def _host_batched(self, lead) -> bool:
        if isinstance(lead, torch.Tensor) or np.ndim(lead) != 2:
            return False
        self._default_device()                      # raises without a CUDA device: there is no CPU path
        return True

    def _host_in(self, x, tail, name, B=None):
        if x is None:
            return None
        a = np.asarray(x)
        if a.dtype != self._np_dtype or not a.flags.c_contiguous:
            a = np.ascontiguousarray(a, dtype=self._np_dtype)
        if a.ndim != 1 + len(tail) or tuple(a.shape[1:]) != tuple(tail):
            raise ValueError("%s: expected trailing shape %s, got %s" % (name, tuple(tail), tuple(a.shape[1:])))
        if B is not None and a.shape[0] != B:
            raise ValueError("%s: batch size %d does not match %d" % (name, a.shape[0], B))
        return a

    def _host_out(self, out, B, tail, name):
        shape = (B,) + tuple(tail)
        if out is None:
            return pinned_empty(shape, self._np_dtype)
        if isinstance(out, torch.Tensor) or not isinstance(out, np.ndarray):
            raise ValueError("%s: numpy inputs take a numpy result buffer" % name)
        if out.dtype != self._np_dtype or tuple(out.shape) != shape or not out.flags.c_contiguous or not out.flags.writeable:
            raise ValueError("%s must be a writable C-contiguous %s array of shape %s" % (name, np.dtype(self._np_dtype), shape))
        return out

    def _pipeline(self) -> HostPipeline:
        dev = self._default_device()
        p = self._pipes.get(dev)
        if p is None:
            p = self._pipes[dev] = HostPipeline(dev, self.dtype)
        return p

    def _host_run(self, name, B, ins, in_tails, outs, out_tails, make_args, handle=None):
        """Run C-ABI driver `name` over host arrays through the pipeline.  `make_args(dins, douts)` returns
        the argument list between the batch size and the stream."""
        fn = getattr(self._lib, "rbd_%s_%s" % (name, self._suffix))
        h = (handle or self._handle).ptr
        pipe = self._pipeline()

        def launch(m, dins, douts):
            args = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in make_args(dins, douts)]
            rc = fn(h, m, *args, torch.cuda.current_stream(pipe.device).cuda_stream)
            _capi.check(rc, "rbd_%s_%s" % (name, self._suffix))

        pipe.run(B, ins, in_tails, outs, out_tails, launch)

    def set_variant(self, variant: int) -> None:
        """Kernel family for THIS engine's handle only (-1 = follow `set_kernel_variant`).  Floating-base robots know
        two families: 0 / 3 warp-cooperative kernels in base coordinates, 1 / 2 one knot point per thread."""
        if self.floating_base:
            _capi.check(self._lib.rbd_fb_model_set_kernel_variant(self._handle.ptr, int(variant)), "rbd_fb_model_set_kernel_variant")
            return
        _capi.check(self._lib.rbd_model_set_kernel_variant(self._handle.ptr, int(variant)), "rbd_model_set_kernel_variant")

    def _call(self, name: str, ctx: "_Ctx", *args, handle=None):
        if ctx.B == 0:
            return                                   # empty batch: nothing to launch
        fn = getattr(self._lib, "rbd_%s_%s" % (name, self._suffix))
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                conv.append(a.data_ptr())
            else:
                conv.append(a)
        with torch.cuda.device(ctx.device):
            stream = torch.cuda.current_stream(ctx.device).cuda_stream
            rc = fn((handle or self._handle).ptr, ctx.B, *conv, stream)
        _capi.check(rc, "rbd_%s_%s" % (name, self._suffix))

    # ------------------------------------------------------------------------------------
    # RNEA
    # ------------------------------------------------------------------------------------
    def rnea_fpass(self, q, qd, qdd=None, GRAVITY=-9.81):
        """RBDReference.py:559-598 -> (v, a, f), each (6, NB)."""
        ctx = self._Ctx(self, q, 1)
        n, NB = self.n, self.NB
        dq, dqd, dqdd = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(qdd, (n,), "qdd")
        v, a, f = ctx.empty(6, NB), ctx.empty(6, NB), ctx.empty(6, NB)
        self._call(self._fb + "rnea_fpass", ctx, dq, dqd, dqdd, float(GRAVITY), v, a, f)
        return ctx.ret(v), ctx.ret(a), ctx.ret(f)

    def rnea_bpass(self, q, f):
        """RBDReference.py:600-621 -> (c, f); `f` is accumulated in place and returned."""
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq, df = ctx.dev(q, (self.nq,), "q"), ctx.dev(f, (6, self.NB), "f")
        c = ctx.empty(n)
        self._call(self._fb + "rnea_bpass", ctx, dq, df, c)
        return ctx.ret(c), ctx.write_back(f, df)

    def rnea(self, q, qd, qdd=None, GRAVITY=-9.81, f_ext=None, outputs="all"):
        """RBDReference.py:623-628 -> (c, v, a, f).  `f_ext` is accepted and ignored as upstream.

        `outputs="c"` (extension) skips writing v, a, f and returns only c.
        """
        n, NB = self.n, self.NB
        op = "fb_rnea" if self.floating_base else "rnea"
        if self._host_batched(q):
            hq = self._host_in(q, (self.nq,), "q")
            B = hq.shape[0]
            hqd, hqdd = self._host_in(qd, (n,), "qd", B), self._host_in(qdd, (n,), "qdd", B)
            g = float(GRAVITY)
            if outputs == "c":
                c = self._host_out(None, B, (n,), "c")
                self._host_run(op, B, [hq, hqd, hqdd], [(self.nq,), (n,), (n,)], [c], [(n,)],
                               lambda di, do: [di[0], di[1], di[2], g, do[0], None, None, None])
                return c
            res = [self._host_out(None, B, t, "out") for t in ((n,), (6, NB), (6, NB), (6, NB))]
            self._host_run(op, B, [hq, hqd, hqdd], [(self.nq,), (n,), (n,)], res, [(n,), (6, NB), (6, NB), (6, NB)],
                           lambda di, do: [di[0], di[1], di[2], g, do[0], do[1], do[2], do[3]])
            return tuple(res)
        ctx = self._Ctx(self, q, 1)
        dq, dqd, dqdd = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(qdd, (n,), "qdd")
        c = ctx.empty(n)
        if outputs == "c":
            self._call(op, ctx, dq, dqd, dqdd, float(GRAVITY), c, None, None, None)
            return ctx.ret(c)
        v, a, f = ctx.empty(6, NB), ctx.empty(6, NB), ctx.empty(6, NB)
        self._call(op, ctx, dq, dqd, dqdd, float(GRAVITY), c, v, a, f)
        return ctx.ret(c), ctx.ret(v), ctx.ret(a), ctx.ret(f)

    # ------------------------------------------------------------------------------------
    # Minv
    # ------------------------------------------------------------------------------------
    def minv_bpass(self, q):
        """RBDReference.py:630-735 -> (Minv, F, U, Dinv) with Dinv = D (:698)."""
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq = ctx.dev(q, (self.nq,), "q")
        Minv, F, U, D = ctx.empty(n, n), ctx.empty(n, 6, n), ctx.empty(n, 6), ctx.empty(n)
        self._call(self._fb + "minv_bpass", ctx, dq, Minv, F, U, D)
        return ctx.ret(Minv), ctx.ret(F), ctx.ret(U), ctx.ret(D)

    def minv_fpass(self, q, Minv, F, U, Dinv):
        """RBDReference.py:737-783 -> Minv (the caller's array, updated in place; F is rewritten)."""
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq = ctx.dev(q, (self.nq,), "q")
        dM, dF = ctx.dev(Minv, (n, n), "Minv"), ctx.dev(F, (n, 6, n), "F")
        dU, dD = ctx.dev(U, (n, 6), "U"), ctx.dev(Dinv, (n,), "Dinv")
        self._call(self._fb + "minv_fpass", ctx, dq, dM, dF, dU, dD)
        ctx.write_back(F, dF)
        return ctx.write_back(Minv, dM)

    def minv(self, q, output_dense=True, out=None):
        """RBDReference.py:785-806 -> Minv (n, n)."""
        n = self.n
        op = "fb_minv" if self.floating_base else "minv"
        if self._host_batched(q):
            hq = self._host_in(q, (self.nq,), "q")
            B = hq.shape[0]
            res = self._host_out(out, B, (n, n), "out")
            dense = 1 if output_dense else 0
            self._host_run(op, B, [hq], [(self.nq,)], [res], [(n, n)], lambda di, do: [di[0], dense, do[0]])
            return res
        ctx = self._Ctx(self, q, 1)
        dq = ctx.dev(q, (self.nq,), "q")
        Minv = self._check_out(out, ctx, (ctx.B, n, n), "out") if out is not None else ctx.empty(n, n)
        self._call(op, ctx, dq, 1 if output_dense else 0, Minv)
        return ctx.ret(Minv)

    # ------------------------------------------------------------------------------------
    # RNEA gradient
    # ------------------------------------------------------------------------------------
    def rnea_grad_fpass_dq(self, q, qd, v, a, GRAVITY=-9.81):
        """RBDReference.py:1127-1187 -> (dv_dq, da_dq, df_dq), each (6, n, NB)."""
        ctx = self._Ctx(self, q, 1)
        n, NB = self.n, self.NB
        dq, dqd = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd")
        dv_, da_ = ctx.dev(v, (6, NB), "v"), ctx.dev(a, (6, NB), "a")
        dv, da, df = ctx.empty(6, n, NB), ctx.empty(6, n, NB), ctx.empty(6, n, NB)
        self._call(self._fb + "rnea_grad_fpass_dq", ctx, dq, dqd, dv_, da_, float(GRAVITY), dv, da, df)
        return ctx.ret(dv), ctx.ret(da), ctx.ret(df)

    def rnea_grad_fpass_dqd(self, q, qd, v):
        """RBDReference.py:1189-1255 -> (dv_dqd, da_dqd, df_dqd)."""
        ctx = self._Ctx(self, q, 1)
        n, NB = self.n, self.NB
        dq, dqd, dv_ = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(v, (6, NB), "v")
        dv, da, df = ctx.empty(6, n, NB), ctx.empty(6, n, NB), ctx.empty(6, n, NB)
        self._call(self._fb + "rnea_grad_fpass_dqd", ctx, dq, dqd, dv_, dv, da, df)
        return ctx.ret(dv), ctx.ret(da), ctx.ret(df)

    def rnea_grad_bpass_dq(self, q, f, df_dq):
        """RBDReference.py:1257-1297 -> dc_dq (n, n); `df_dq` is accumulated in place."""
        ctx = self._Ctx(self, q, 1)
        n, NB = self.n, self.NB
        dq, df_, ddf = ctx.dev(q, (self.nq,), "q"), ctx.dev(f, (6, NB), "f"), ctx.dev(df_dq, (6, n, NB), "df_dq")
        dc = ctx.empty(n, n)
        self._call(self._fb + "rnea_grad_bpass_dq", ctx, dq, df_, ddf, dc)
        ctx.write_back(df_dq, ddf)
        return ctx.ret(dc)

    def rnea_grad_bpass_dqd(self, q, df_dqd, USE_VELOCITY_DAMPING=False):
        """RBDReference.py:1299-1343 -> dc_dqd (n, n); `df_dqd` is accumulated in place."""
        ctx = self._Ctx(self, q, 1)
        n, NB = self.n, self.NB
        dq, ddf = ctx.dev(q, (self.nq,), "q"), ctx.dev(df_dqd, (6, n, NB), "df_dqd")
        dc = ctx.empty(n, n)
        self._call(self._fb + "rnea_grad_bpass_dqd", ctx, dq, ddf, 1 if USE_VELOCITY_DAMPING else 0, dc)
        ctx.write_back(df_dqd, ddf)
        return ctx.ret(dc)

    def rnea_grad(self, q, qd, qdd=None, GRAVITY=-9.81, USE_VELOCITY_DAMPING=False, out=None, c_out=None):
        """RBDReference.py:1345-1368 -> dc_du (n, 2n) = [dc_dq | dc_dqd], one fused launch.

        `out` / `c_out` (extension) receive dc_du (B,n,2n) / c (B,n): CUDA tensors for batched CUDA-tensor
        calls, numpy arrays for batched numpy calls (pinned ones are filled by DMA directly).
        """
        n = self.n
        op = "fb_rnea_grad" if self.floating_base else "rnea_grad"
        if self._host_batched(q):
            hq = self._host_in(q, (self.nq,), "q")
            B = hq.shape[0]
            hqd, hqdd = self._host_in(qd, (n,), "qd", B), self._host_in(qdd, (n,), "qdd", B)
            res = self._host_out(out, B, (n, 2 * n), "out")
            g, damp = float(GRAVITY), 1 if USE_VELOCITY_DAMPING else 0
            if c_out is not None:
                hc = self._host_out(c_out, B, (n,), "c_out")
                self._host_run(op, B, [hq, hqd, hqdd], [(self.nq,), (n,), (n,)], [res, hc], [(n, 2 * n), (n,)],
                               lambda di, do: [di[0], di[1], di[2], g, damp, do[0], do[1]])
            else:
                self._host_run(op, B, [hq, hqd, hqdd], [(self.nq,), (n,), (n,)], [res], [(n, 2 * n)],
                               lambda di, do: [di[0], di[1], di[2], g, damp, do[0], None])
            return res
        ctx = self._Ctx(self, q, 1)
        dq, dqd, dqdd = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(qdd, (n,), "qdd")
        dc_du = self._check_out(out, ctx, (ctx.B, n, 2 * n), "out") if out is not None else ctx.empty(n, 2 * n)
        c_out = self._check_out(c_out, ctx, (ctx.B, n), "c_out")
        self._call(op, ctx, dq, dqd, dqdd, float(GRAVITY), 1 if USE_VELOCITY_DAMPING else 0, dc_du, c_out)
        return ctx.ret(dc_du)

    def rnea_grad_passes(self, q, qd, qdd=None, GRAVITY=-9.81, USE_VELOCITY_DAMPING=False):
        """rnea_grad composed from the per-pass entry points (the reference's own call sequence,
        RBDReference.py:1353-1367).  Debug path: materialises every (6,n,NB) intermediate."""
        c, v, a, f = self.rnea(q, qd, qdd, GRAVITY)
        _, _, df_dq = self.rnea_grad_fpass_dq(q, qd, v, a, GRAVITY)
        _, _, df_dqd = self.rnea_grad_fpass_dqd(q, qd, v)
        dc_dq = self.rnea_grad_bpass_dq(q, f, df_dq)
        dc_dqd = self.rnea_grad_bpass_dqd(q, df_dqd, USE_VELOCITY_DAMPING)
        if isinstance(dc_dq, torch.Tensor):
            return torch.cat((dc_dq, dc_dqd), dim=-1)
        return np.concatenate((dc_dq, dc_dqd), axis=-1)

    def minv_passes(self, q, output_dense=True):
        """minv composed from minv_bpass + minv_fpass (+ mirror), RBDReference.py:793-804."""
        Minv, F, U, D = self.minv_bpass(q)
        Minv = self.minv_fpass(q, Minv, F, U, D)
        if output_dense:
            n = self.NB                                 # :799-804 loops over range(NB): with a floating base only that block
            iu = np.triu_indices(n, 1)
            if isinstance(Minv, torch.Tensor):
                Minv[..., iu[1], iu[0]] = Minv[..., iu[0], iu[1]]
            else:
                Minv[..., iu[1], iu[0]] = Minv[..., iu[0], iu[1]]
        return Minv

    # ------------------------------------------------------------------------------------
    # compositions of the hot-path kernels
    # ------------------------------------------------------------------------------------
    def forward_dynamics(self, q, qd, u):
        """RBDReference.py:1369-1372: Minv @ (u - c) with c = rnea(q, qd) (qdd omitted upstream).
        rnea + minv + one fused (u - c) / matrix-vector kernel, all inside the C library."""
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq, dqd, du = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(u, (n,), "u")
        qdd = ctx.empty(n)
        self._call("fb_forward_dynamics" if self.floating_base else "forward_dynamics", ctx, dq, dqd, du, qdd, None)
        return ctx.ret(qdd)

    def forward_dynamics_grad(self, q, qd, u):
        """RBDReference.py:1374-1384 -> (qdd_dq, qdd_dqd) = (-Minv dc_dq, -Minv dc_dqd)."""
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq, dqd, du = ctx.dev(q, (self.nq,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(u, (n,), "u")
        o1, o2 = ctx.empty(n, n), ctx.empty(n, n)
        self._call("fb_forward_dynamics_grad" if self.floating_base else "forward_dynamics_grad", ctx, dq, dqd, du,
                   o1, o2, None)
        return ctx.ret(o1), ctx.ret(o2)

    def aba(self, q, qd, tau, f_ext=None, GRAVITY=-9.81):
        """RBDReference.py:817 (fixed-base branch :940-1024) -> qdd (n,).  Reproduces the reference's
        aba() including its :984 bias-force quirk; f_ext is ignored, as upstream."""
        self._fixed_only("aba")
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq, dqd, dtau = ctx.dev(q, (n,), "q"), ctx.dev(qd, (n,), "qd"), ctx.dev(tau, (n,), "tau")
        qdd = ctx.empty(n)
        self._call("aba", ctx, dq, dqd, dtau, float(GRAVITY), qdd)
        return ctx.ret(qdd)

    def crba(self, q, out=None):
        """RBDReference.py:1026-1124 (fixed-base branch) -> joint-space inertia matrix H (n, n)."""
        self._fixed_only("crba")
        n = self.n
        if self._host_batched(q):
            hq = self._host_in(q, (n,), "q")
            B = hq.shape[0]
            res = self._host_out(out, B, (n, n), "out")
            self._host_run("crba", B, [hq], [(n,)], [res], [(n, n)], lambda di, do: [di[0], do[0]])
            return res
        ctx = self._Ctx(self, q, 1)
        dq = ctx.dev(q, (n,), "q")
        H = self._check_out(out, ctx, (ctx.B, n, n), "out") if out is not None else ctx.empty(n, n)
        self._call("crba", ctx, dq, H)
        return ctx.ret(H)

    # ------------------------------------------------------------------------------------
    # end-effector kinematics
    # ------------------------------------------------------------------------------------
    def select_end_effector_joints(self, ee_joint_names):
        """RBDReference.py:190-211 -> (ee_jids, fixed_jids); ValueError for an unknown name."""
        if isinstance(self.robot, RobotModel):
            raise ValueError("end-effector kinematics need the robot object, not a compiled RobotModel")
        return select_end_effector_joints(self.robot, ee_joint_names)

    def _ee_handle(self, ee_joint_names, ee_offsets):
        if isinstance(self.robot, RobotModel):
            raise ValueError("end-effector kinematics need the robot object, not a compiled RobotModel")
        off = (0.0, 0.0, 0.0, 1.0) if ee_offsets is None else tuple(
            float(x) for x in np.asarray(ee_offsets[0], dtype=np.float64).reshape(-1))
        key = (None if ee_joint_names is None else tuple(ee_joint_names), off)
        h = self._ee_handles.get(key)
        if h is None:
            h = _capi.EeModelHandle(compile_ee_model(self.robot, ee_joint_names, [off]))
            self._ee_handles[key] = h
        return h

    def end_effector_pose(self, q, ee_joint_names=None, ee_offsets=None):
        """RBDReference.py:220-283.  One knot point -> list over end effectors of (6, 1) arrays
        [x y z roll pitch yaw] (the reference returns np.matrix objects of that shape); batched
        q (B, n) -> (B, n_ee, 6).  `ee_offsets=None` is the reference's default [[0, 0, 0, 1]]; as
        upstream only ee_offsets[0] is used."""
        self._fixed_only("end_effector_pose")
        h = self._ee_handle(ee_joint_names, ee_offsets)
        ctx = self._Ctx(self, q, 1)
        dq = ctx.dev(q, (self.n,), "q")
        pose = ctx.empty(h.n_ee, 6)
        self._call("end_effector_pose", ctx, dq, pose, handle=h)
        out = ctx.ret(pose)
        if not ctx.batched:
            return [out[e].reshape(6, 1) for e in range(h.n_ee)]
        return out

    def end_effector_pose_gradient(self, q, ee_joint_names=None, ee_offsets=None, return_pose=False):
        """RBDReference.py:295-386.  One knot point -> list over end effectors of (6, n) arrays;
        batched -> (B, n_ee, 6, n).  `return_pose=True` (extension) also returns the pose computed by
        the same launch."""
        self._fixed_only("end_effector_pose_gradient")
        h = self._ee_handle(ee_joint_names, ee_offsets)
        ctx = self._Ctx(self, q, 1)
        n = self.n
        dq = ctx.dev(q, (n,), "q")
        grad = ctx.empty(h.n_ee, 6, n)
        pose = ctx.empty(h.n_ee, 6) if return_pose else None
        self._call("end_effector_pose_gradient", ctx, dq, grad, pose, handle=h)
        g = ctx.ret(grad)
        p = ctx.ret(pose)
        if not ctx.batched:
            g = [g[e] for e in range(h.n_ee)]
            p = None if p is None else [p[e].reshape(6, 1) for e in range(h.n_ee)]
        return (g, p) if return_pose else g

    # ------------------------------------------------------------------------------------
    def uses_world_kernels(self) -> bool:
        """True if the fused drivers run the world-frame kernels for this robot."""
        if self.floating_base:
            return False
        return bool(self._lib.rbd_model_uses_world_kernels(self._handle.ptr))

    @staticmethod
    def set_kernel_variant(variant: int) -> None:
        """Process-wide default: 0 = automatic kernel choice; 1 = force the generic body-frame kernels; 2.. see
        include/rbd_b200.h.  `set_variant` overrides it for one engine."""
        lib = _capi.load_library()
        _capi.check(lib.rbd_set_kernel_variant(int(variant)), "rbd_set_kernel_variant")

    def trim_scratch(self, keep_bytes: int = 0) -> None:
        """Return the library's cached scratch memory on this engine's device to the driver (forward_dynamics(_grad)
        and the hybrid minv kernel keep their temporaries in a private pool; see include/rbd_b200.h)."""
        dev = self._default_device()
        torch.cuda.synchronize(dev)
        with torch.cuda.device(dev):
            _capi.check(self._lib.rbd_trim_scratch(int(keep_bytes)), "rbd_trim_scratch")

    def launch_count(self) -> int:
        return int(self._lib.rbd_launch_count())
