"""Layout invariants of the compiled scene blob (scene.py compile_blob <-> csrc/scene_blob.h, version 7): what the
kernels stage in shared memory, what stays in the global tail, and how a broadphase record finds its pair ids."""
import numpy as np
import pytest

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES

QUEUED, DIRECT = range(6), (6, 7)


@pytest.fixture(scope="module", params=list(SCENES))
def compiled(request):
    mk, kw = SCENES[request.param]
    return request.param, S.compile_blob(mk(), kw["tol"])


def test_header_matches_the_c_header():
    import os
    import re
    h = open(os.path.join(os.path.dirname(S.__file__), "csrc", "scene_blob.h")).read()
    c = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+MRB_(\w+)\s+(0x[0-9A-Fa-f]+|\d+)\b", h)}
    for py, cname in (("BLOB_VERSION", "BLOB_VERSION"), ("HDR_WORDS", "HDR_WORDS"), ("H_STAGED_WORDS", "H_STAGED_WORDS"),
                      ("H_REC_BASE", "H_REC_BASE"), ("H_IDS_BASE", "H_IDS_BASE"), ("H_IDS_STAGED", "H_IDS_STAGED"),
                      ("H_GPTR", "H_GPTR"), ("H_BP", "H_BP"), ("H_BP_IDS", "H_BP_IDS"), ("BP_SUBLISTS", "BP_SUBLISTS"),
                      ("H_OFF_PAIRS", "H_OFF_PAIRS"), ("H_N_PAIRS", "H_N_PAIRS"), ("H_OFF_SCENTRE", "H_OFF_SCENTRE"),
                      ("FRAME_WORDS", "FRAME_WORDS"), ("SHAPE_WORDS", "SHAPE_WORDS")):
        assert getattr(S, py) == c[cname], (py, getattr(S, py), c[cname])
    assert S.H_GPTR % 2 == 0 and S.H_GPTR + 2 <= S.H_BP           # an aligned, otherwise unused 64-bit header slot
    assert S.H_BP_IDS + 8 * S.BP_SUBLISTS <= S.HDR_WORDS


def test_staged_prefix_and_tail(compiled):
    name, cs = compiled
    b = cs.blob32.astype(np.int64)
    total, staged = int(b[S.H_TOTAL_WORDS]), int(b[S.H_STAGED_WORDS])
    assert total == len(cs.blob32) and staged == cs.staged_words
    assert staged % 4 == 0 and total % 4 == 0 and S.HDR_WORDS <= staged <= total     # 16-byte multiples for cp.async.bulk
    # everything the kernels touch through shared memory lies inside the prefix
    for off, n in ((b[S.H_OFF_FRAMES], cs.n_frames * S.FRAME_WORDS), (b[S.H_OFF_SHAPES], (cs.n_moving + cs.n_static) * S.SHAPE_WORDS),
                   (b[S.H_OFF_CHAINS], 2 * b[S.H_NCHAINS]), (b[S.H_OFF_SCENTRE], 4 * cs.n_static)):
        assert S.HDR_WORDS <= off and off + n <= staged
    assert b[S.H_OFF_SCENTRE] % 4 == 0 and b[S.H_REC_BASE] % 2 == 0                   # float4 / 8-byte record reads
    n_rec = 0
    for t in range(S.NUM_PAIR_TYPES):
        for k in range(S.BP_SUBLISTS):
            off, n = b[S.H_BP + (t * S.BP_SUBLISTS + k) * 2], b[S.H_BP + (t * S.BP_SUBLISTS + k) * 2 + 1]
            if n:
                assert b[S.H_REC_BASE] <= off and off + 2 * n <= staged
                assert b[S.H_BP_IDS + t * S.BP_SUBLISTS + k] == b[S.H_IDS_BASE] + (off - b[S.H_REC_BASE]) // 2
            n_rec += n
    # the pair lists of the directly evaluated types are staged, those of the queued types are not
    for t in range(S.NUM_PAIR_TYPES):
        off, n = b[S.H_OFF_PAIRS + t], b[S.H_N_PAIRS + t]
        if n:
            assert (off + n <= staged) if t in DIRECT else (off >= staged and off + n <= total)
    assert b[S.H_OFF_STATIC_PAIRS] >= staged
    # pair ids: inside the prefix on small scenes, in the tail on large ones, never straddling
    ids0, ids1 = b[S.H_IDS_BASE], b[S.H_IDS_BASE] + n_rec
    if b[S.H_IDS_STAGED]:
        assert ids1 <= staged and staged * 4 <= S.STAGE_IDS_MAX_BYTES + 16
    else:
        assert ids0 >= staged and ids1 <= total
    assert b[S.H_GPTR] == 0 and b[S.H_GPTR + 1] == 0                                   # the device's slot


def test_every_record_names_a_pair_of_its_type(compiled):
    name, cs = compiled
    b = cs.blob32.astype(np.int64)
    offS = int(b[S.H_OFF_SHAPES])
    for t in QUEUED:
        lo, n = b[S.H_OFF_PAIRS + t], b[S.H_N_PAIRS + t]
        listed = {(int(p) & 0xffff, (int(p) >> 16) & 0xfff) for p in b[lo:lo + n]}
        seen = set()
        for k in range(S.BP_SUBLISTS):
            off, m = b[S.H_BP + (t * S.BP_SUBLISTS + k) * 2], b[S.H_BP + (t * S.BP_SUBLISTS + k) * 2 + 1]
            ids = b[S.H_BP_IDS + t * S.BP_SUBLISTS + k]
            for i in range(m):
                pk = int(b[ids + i])
                a, c = pk & 0xffff, (pk >> 16) & 0xfff
                assert (a, c) in listed and (a, c) not in seen
                seen.add((a, c))
                x, y = (a, c) if a < cs.n_moving else (c, a)
                rec = int(b[off + 2 * i])
                assert (rec & 0xffff) == int(b[offS + x * S.SHAPE_WORDS + 2]) * 128        # X: byte offset of its W row
                if k == 0:
                    assert (rec >> 16) == int(b[offS + y * S.SHAPE_WORDS + 2]) * 128       # Y moving: byte offset
                else:
                    assert (rec >> 16) == y - cs.n_moving                                  # Y static: static index
        dynamic = {p for p in listed if p[0] < cs.n_moving or p[1] < cs.n_moving}
        idx = {n_: i for i, n_ in enumerate(cs.shape_names)}
        unreachable = {frozenset((idx[a], idx[c])) for a, c in cs.unreachable_pairs}
        assert {frozenset(p) for p in dynamic} == {frozenset(p) for p in seen} | {frozenset(p) for p in dynamic if frozenset(p) in unreachable}
