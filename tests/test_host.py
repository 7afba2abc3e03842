"""CPU tests of the host side: model compiler, C-ABI library surface, sharding, loud failures."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, make_robot
from rbdreference_b200 import _capi, compile_model, robots
from rbdreference_b200.dist import shard_bounds


def test_library_loads_and_exports_every_declared_symbol():
    lib = _capi.load_library()
    header = open(os.path.join(ROOT, "include", "rbd_b200.h")).read()
    declared = set(re.findall(r"\b(rbd_[a-z0-9_]+)\s*\(", header))
    declared -= {"rbd_model", "rbd_ee_model", "rbd_fb_model"}
    assert declared == set(_capi.exported_symbols()), declared ^ set(_capi.exported_symbols())
    for sym in sorted(declared):
        assert hasattr(lib, sym), "missing export " + sym
    assert lib.rbd_abi_version() == 1


def test_model_create_validates_arguments():
    lib = _capi.load_library()
    handle = ctypes.c_void_p()
    assert lib.rbd_model_create(None, ctypes.byref(handle)) == -1
    assert b"null" in lib.rbd_last_error_string()
    m = compile_model(robots.iiwa14())
    h = _capi.ModelHandle(m)
    assert lib.rbd_model_num_dof(h.ptr) == 7
    bad = compile_model(robots.iiwa14())
    bad.parent = bad.parent.copy(); bad.parent[0] = 3
    with pytest.raises(_capi.RbdError):
        _capi.ModelHandle(bad)


@pytest.mark.parametrize("name", ["iiwa14", "hyq", "atlas", "tree9", "tree13"])
def test_model_compiler_reproduces_transforms(name):
    rb = make_robot(name)
    m = compile_model(rb)
    n = m.n
    rng = np.random.default_rng(3)
    for i in range(n):
        for t in rng.uniform(-3, 3, 3):
            f1, f2 = (np.cos(t), np.sin(t)) if m.kind[i] == 0 else (t, 0.0)
            x18 = m.XA[i] + f1 * m.XB[i] + f2 * m.XC[i]
            X = np.zeros((6, 6))
            X[:3, :3] = x18[:9].reshape(3, 3); X[3:, 3:] = X[:3, :3]; X[3:, :3] = x18[9:].reshape(3, 3)
            assert np.max(np.abs(X - rb.get_Xmat_Func_by_id(i)(t))) < 1e-13
    for i in range(n):
        assert sorted(rb.get_subtree_by_id(i)) == m.subtree[i]
        anc = set(rb.get_ancestors_by_id(i)) | {i}
        assert anc == {c for c in range(n) if (int(m.anc_mask[i]) >> c) & 1}


@pytest.mark.parametrize("name", ["iiwa14", "hyq", "atlas", "tree9", "tree13"])
def test_ee_model_compiler_reproduces_hom_transforms(name):
    """compile_ee_model recovers T(q) and dT(q) of every joint from the robot's own callables and
    resolves the end-effector selection in the reference's order (RBDReference.py:190-211)."""
    from rbdreference_b200.model import compile_ee_model
    rb = make_robot(name)
    ee = compile_ee_model(rb)
    assert list(ee.ee_joint) == rb.get_leaf_nodes()
    rng = np.random.default_rng(3)
    for i in range(ee.n):
        for t in rng.uniform(-3, 3, 3):
            f1, f2 = (np.cos(t), np.sin(t)) if ee.kind[i] == 0 else (t, 0.0)
            T = (ee.TA[i] + f1 * ee.TB[i] + f2 * ee.TC[i]).reshape(3, 4)
            D = (ee.DA[i] + f1 * ee.DB[i] + f2 * ee.DC[i]).reshape(3, 4)
            assert np.max(np.abs(T - rb.get_Xmat_hom_Func_by_id(i)(t)[:3])) < 1e-13
            assert np.max(np.abs(D - rb.get_dXmat_hom_Func_by_id(i)(t)[:3])) < 1e-13
    h = _capi.EeModelHandle(ee)
    assert _capi.load_library().rbd_ee_model_num_ee(h.ptr) == len(rb.get_leaf_nodes())


def test_ee_selection_order_and_errors():
    from rbdreference_b200.model import compile_ee_model, select_end_effector_joints
    rb = robots.iiwa14()
    # moving joints first, then fixed joints, whatever the order of the names (:201-210)
    assert select_end_effector_joints(rb, ["iiwa_joint_ee", "iiwa_joint_4", "iiwa_tool_tip"]) == ([3], [0, 1])
    ee = compile_ee_model(rb, ["iiwa_joint_ee", "iiwa_joint_4", "iiwa_tool_tip"], [[0.1, 0.2, 0.3, 1.0]])
    assert list(ee.ee_joint) == [3, 6, 6] and list(ee.offset) == [0.1, 0.2, 0.3, 1.0]
    assert np.array_equal(ee.ee_final[0], np.eye(4)[:3].reshape(12))
    assert np.allclose(ee.ee_final[1].reshape(3, 4)[:, 3], [0, 0, 0.045])
    with pytest.raises(ValueError, match="Could not find joint or fixed joint named: nope"):
        compile_ee_model(rb, ["nope"])
    lib = _capi.load_library()
    handle = ctypes.c_void_p()
    assert lib.rbd_ee_model_create(None, ctypes.byref(handle)) == -1
    bad = compile_ee_model(rb)
    bad.ee_joint = np.array([9], dtype=np.int32)
    with pytest.raises(_capi.RbdError):
        _capi.EeModelHandle(bad)


def test_fb_model_compiler_finds_the_base_layout():
    """compile_fb_model re-discovers position / quaternion layout and rotation sense of the base
    transform by probing the robot's own callable, and rejects what it cannot represent."""
    from rbdreference_b200.model import compile_fb_model
    rb = robots.by_name("hyq_fb")
    fb = compile_fb_model(rb)
    assert (fb.NB, fb.nv, fb.nq) == (13, 18, 19)
    assert (fb.pos_off, fb.quat_off, fb.w_first, fb.transpose) == (0, 3, 0, 0)
    assert list(fb.parent[:5]) == [-1, 0, 1, 2, 0] and np.array_equal(fb.I[0].reshape(6, 6), rb.get_Imat_by_id(0))
    assert np.array_equal(fb.XA[1:], compile_model(robots.hyq()).XA)
    h = _capi.FbModelHandle(fb)
    assert _capi.load_library().rbd_fb_model_num_vel(h.ptr) == 18

    class QuatFirstW(robots.FloatingBaseRobot):       # q[0:7] = (w, x, y, z, px, py, pz), E = R
        def get_Xmat_Func_by_id(self, i):
            if i != 0:
                return super().get_Xmat_Func_by_id(i)
            return lambda q7: robots.xrot(robots.quat_rotation([q7[1], q7[2], q7[3], q7[0]])) @ robots.xlt(q7[4:7])
    fb2 = compile_fb_model(QuatFirstW(robots.hyq()))
    assert (fb2.pos_off, fb2.quat_off, fb2.w_first, fb2.transpose) == (4, 0, 1, 1)

    class Odd(robots.FloatingBaseRobot):
        def get_Xmat_Func_by_id(self, i):
            if i != 0:
                return super().get_Xmat_Func_by_id(i)
            return lambda q7: 2.0 * robots.FloatingBaseRobot.base_transform(q7)
    with pytest.raises(ValueError, match="supported layout"):
        compile_fb_model(Odd(robots.hyq()))
    bad = compile_fb_model(rb)
    bad.parent = bad.parent.copy(); bad.parent[3] = -1
    with pytest.raises(_capi.RbdError):
        _capi.FbModelHandle(bad)


def test_model_compiler_flop_model_matches_survey_table():
    """algorithmic flops / bytes per evaluation."""
    expect = {"iiwa14": (2854, 26806, 9738), "hyq": (4440, 21912, 10464), "atlas": (12486, 144510, 53097)}
    for name, (r, g, mi) in expect.items():
        m = compile_model(make_robot(name))
        assert (m.flops("rnea"), m.flops("rnea_grad"), m.flops("minv")) == (r, g, mi)
    m = compile_model(make_robot("iiwa14"))
    assert m.io_bytes("rnea_grad") == 952 and m.io_bytes("minv") == 448 and m.io_bytes("rnea") == 224
    assert m.io_bytes("rnea", full_rnea=True) == 1232


def test_model_compiler_rejects_unsupported_robots():
    rb = robots.iiwa14()
    rb.floating_base = True
    with pytest.raises(NotImplementedError):
        compile_model(rb)
    big = robots.random_tree(40, seed=0)
    with pytest.raises(ValueError):
        compile_model(big)

    class Weird(robots.Robot):
        def get_Xmat_Func_by_id(self, i):
            base = super().get_Xmat_Func_by_id(i)
            return lambda q: base(q * q)        # not of 1-DoF A + B cos + C sin form
    with pytest.raises(ValueError):
        compile_model(Weird("weird", robots.iiwa14().joints))


def test_S_shape_variants_are_normalised():
    """The reference accepts S as (6,), (6,1) ndarray or (6,1) np.matrix."""
    class ColS(robots.Robot):
        def get_S_by_id(self, i):
            return super().get_S_by_id(i).reshape(6, 1)
    a = compile_model(robots.iiwa14())
    b = compile_model(ColS("iiwa14", robots.iiwa14().joints))
    assert np.array_equal(a.S, b.S)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_without_gpu_fails_loudly():
    from rbdreference_b200 import RBDReference
    eng = RBDReference(robots.iiwa14())
    with pytest.raises(RuntimeError, match="no CUDA device"):
        eng.rnea_grad(np.zeros(7), np.zeros(7), np.zeros(7))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        eng.minv(np.zeros((4, 7)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rbdreference_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]", src, re.M), fn + " uses the oracle"


def test_shard_bounds_partition_the_batch():
    for B in (0, 1, 7, 1000, 1 << 20):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["RBD_ROOT"])
import torch, torch.distributed as dist
from rbdreference_b200.dist import shard_bounds, gather_to_all, gather_to_rank
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
for B in (10, 7, 1):                      # even, ragged, fewer knot points than ranks
    full = torch.arange(B * 3, dtype=torch.float64).reshape(B, 3)
    lo, hi = shard_bounds(B, rank, world)
    local = full[lo:hi].clone()
    assert torch.equal(gather_to_all(local, B), full)
    for dst in range(world):
        got = gather_to_rank(local, B, dst=dst)
        assert (got is None) == (rank != dst)
        if rank == dst:
            assert torch.equal(got, full)
if world >= 3:                            # a sub-group that does not start at global rank 0: group ranks != global ranks
    members = list(range(1, world))
    grp = dist.new_group(members)
    if rank in members:
        gr, gw = dist.get_rank(grp), len(members)
        for B in (9, 4):
            full = torch.arange(B * 2, dtype=torch.float64).reshape(B, 2)
            lo, hi = shard_bounds(B, gr, gw)
            local = full[lo:hi].clone()
            assert torch.equal(gather_to_all(local, B, group=grp), full)
            got = gather_to_rank(local, B, dst=gw - 1, group=grp)
            assert (got is None) == (gr != gw - 1)
            if gr == gw - 1:
                assert torch.equal(got, full)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_gather(tmp_path, world):
    """world_size-2 and -3 runs of the sharding + gather plumbing on CPU (gloo): ragged batches, every destination
    rank, and (world 3) a sub-group whose group ranks differ from the global ranks."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, RBD_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29613 + world), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_reference_arm_prints_the_contract_line_with_our_config():
    """`bench.py --impl reference` (CPU only): one JSON line with impl / cpu_baseline / e2e, and the same `config`
    dict our arm prints for the same workload (the driver compares them)."""
    import json
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0

    class A:
        robot, op, dtype, batch = "iiwa14", "rnea_grad", "f64", 1 << 20
    assert line["config"] == bench.config_for(A, 7)
    assert line["config"]["n_dof"] == 7 and line["config"]["batch_per_gpu"] == 1 << 20
