"""Error behaviour of the C ABI (include/mrb200.h): failures are loud negative codes with a message, never a
"free" answer (the reference's convention, P/problems/mujoco_env.py:797-807), and there is no CPU fallback."""
import ctypes as C

import numpy as np
import pytest

from multirobot_pathplanning_benchmark_b200 import _lib
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES

ERR_ARG, ERR_BLOB, ERR_CUDA, ERR_NO_DEVICE = -1, -2, -3, -4


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour of a machine WITHOUT a GPU")
def test_no_device_is_a_loud_error_not_a_fallback(cuda_lib):
    h = C.c_void_p()
    rc = cuda_lib.mrb200_scene_create(4, C.byref(h))
    assert rc == ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in cuda_lib.mrb200_last_error()
    from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
    with pytest.raises(Exception):
        SceneBackend(max_modes=2)
    radii = (C.c_double * 2)(0.1, 0.1)
    assert cuda_lib.mrb200_abstract_create(2, 2, radii, 0, None, 0, None, C.byref(h)) < 0


def test_pure_argument_checks_need_no_device(cuda_lib):
    assert cuda_lib.mrb200_scene_create(0, C.byref(C.c_void_p())) == ERR_ARG
    assert cuda_lib.mrb200_scene_create(4, None) == ERR_ARG
    assert cuda_lib.mrb200_check_configs(None, 0, None, 10, -1.0, None, None, 0, None) == ERR_ARG
    assert cuda_lib.mrb200_scene_destroy(None) == 0 and cuda_lib.mrb200_abstract_destroy(None) == 0
    assert cuda_lib.mrb200_version() >= 1
    out3 = (C.c_int32 * 3)()
    assert cuda_lib.mrb200_scene_set_two_phase(None, 0, 1) == ERR_ARG
    assert cuda_lib.mrb200_scene_get_two_phase(None, 0, out3) == ERR_ARG
    assert b"two_phase" in cuda_lib.mrb200_last_error()


@pytest.mark.gpu
def test_scene_argument_and_blob_errors(cuda_lib):
    import torch
    h = C.c_void_p()
    assert cuda_lib.mrb200_scene_create(2, C.byref(h)) == 0
    try:
        mk, kw = SCENES["2d_handover"]
        cs = S.compile_blob(mk(), kw["tol"])
        blob = cs.blob32.copy()
        q = torch.zeros((64, cs.dof), dtype=torch.float32, device="cuda")
        out = torch.full((64,), 7, dtype=torch.uint8, device="cuda")
        # empty slot
        assert cuda_lib.mrb200_check_configs(h, 0, q.data_ptr(), 64, -1.0, out.data_ptr(), None, 0, None) == ERR_ARG
        assert b"empty mode slot" in cuda_lib.mrb200_last_error()
        # wrong magic / wrong version / truncated blob
        bad = blob.copy(); bad[S.H_MAGIC] ^= 1
        assert cuda_lib.mrb200_scene_set_mode(h, 0, bad.ctypes.data, bad.nbytes, None) == ERR_BLOB
        bad = blob.copy(); bad[S.H_VERSION] += 1
        assert cuda_lib.mrb200_scene_set_mode(h, 0, bad.ctypes.data, bad.nbytes, None) == ERR_BLOB
        assert cuda_lib.mrb200_scene_set_mode(h, 0, blob.ctypes.data, blob.nbytes // 2, None) == ERR_BLOB
        # slot out of range, null blob
        assert cuda_lib.mrb200_scene_set_mode(h, 2, blob.ctypes.data, blob.nbytes, None) == ERR_ARG
        assert cuda_lib.mrb200_scene_set_mode(h, 0, None, blob.nbytes, None) == ERR_ARG
        assert cuda_lib.mrb200_scene_set_mode(h, 0, blob.ctypes.data, blob.nbytes, None) == 0
        # null buffers, negative sizes; an empty batch is fine and launches nothing
        assert cuda_lib.mrb200_check_configs(h, 0, None, 64, -1.0, out.data_ptr(), None, 0, None) == ERR_ARG
        assert cuda_lib.mrb200_check_configs(h, 0, q.data_ptr(), -1, -1.0, out.data_ptr(), None, 0, None) == ERR_ARG
        assert cuda_lib.mrb200_check_configs(h, 0, q.data_ptr(), 0, -1.0, out.data_ptr(), None, 0, None) == 0
        assert cuda_lib.mrb200_check_edges(h, 0, q.data_ptr(), None, 8, 0.01, None, 0, -1, 0, -1.0, out.data_ptr(), None, None) == ERR_ARG
        torch.cuda.synchronize()
        assert bool((out == 7).all()), "a failed call must not write answers"
        # per-robot rule needs one flag per shape
        flags = (C.c_uint8 * 1)(1)
        assert cuda_lib.mrb200_check_configs_for_robot(h, 0, q.data_ptr(), 64, -1.0, flags, flags, 1, out.data_ptr(), None) == ERR_ARG
    finally:
        assert cuda_lib.mrb200_scene_destroy(h) == 0


@pytest.mark.gpu
def test_neighbour_search_argument_errors(cuda_lib):
    import torch
    from multirobot_pathplanning_benchmark_b200 import knn as K
    pts = torch.rand((100, 6), dtype=torch.float64, device="cuda")
    with pytest.raises(Exception):
        K.batch_knn(pts, pts, [[0, 3], [3, 6]], "max_euclidean", 0)           # k < 1
    with pytest.raises(Exception):
        K.batch_knn(pts, pts, [[0, 3], [2, 7]], "max_euclidean", 3)           # slice outside D
    with pytest.raises(Exception):
        K.batch_knn(pts, pts, [[0, 3], [3, 6]], "chebyshev", 3)               # unknown metric
    with pytest.raises(Exception):
        K.batch_knn(pts.float(), pts, [[0, 3], [3, 6]], "euclidean", 3)       # fp32 queries: the path is fp64 like the reference
    with pytest.raises(Exception):
        K.batch_knn(pts, pts, [[0, 3], [3, 6]], "euclidean", 3, mode="tensor")  # tensor path needs N >= 1024
    # k larger than the corpus: clipped like the reference (prm_graph.py:440-445) -> -1 padding or N columns
    res = K.batch_knn(pts[:5], pts[:3], [[0, 3], [3, 6]], "euclidean", 3)
    idx = res[0] if isinstance(res, tuple) else res
    assert idx.shape == (5, 3) and int(idx.min()) >= 0
