"""Test-only stand-ins for the CUDA devices, backed by the CPU oracle.  They let the CPU suite
exercise the host logic of env.py (mode relinking, masks, planner compatibility) on a machine
without a GPU.  Never imported by the product."""
import numpy as np

from oracle import oracle_abstract as OA
from oracle import oracle_scene as O


class OracleSceneDevice:
    def __init__(self):
        self.cs = {}
        self.calls = {"configs": 0, "edges": 0, "robot": 0}

    def set_mode(self, slot, cs):
        self.cs[slot] = cs

    def check_configs(self, slot, q, tol=None):
        self.calls["configs"] += 1
        q = np.asarray(q, np.float32).astype(np.float64)
        return O.check_configs(self.cs[slot].blob64, q, -1.0 if tol is None else tol)[0]

    def check_configs_for_robot(self, slot, q, rel, oth, tol=None):
        self.calls["robot"] += 1
        q = np.asarray(q, np.float32).astype(np.float64)
        return O.check_configs(self.cs[slot].blob64, q, -1.0 if tol is None else tol, rel=rel, oth=oth)[0]

    def check_edges(self, slot, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False, tol=None):
        self.calls["edges"] += 1
        q1 = np.asarray(q1, np.float32).astype(np.float64)
        q2 = np.asarray(q2, np.float32).astype(np.float64)
        f, p, _ = O.check_edges(self.cs[slot].blob64, q1, q2, resolution, Ns=N, n_start=n_start,
                                n_max=-1 if n_max is None else n_max, include_endpoints=include_endpoints,
                                tol=-1.0 if tol is None else tol)
        return f, p


class OracleAbstractDevice:
    def __init__(self):
        self.sc = OA.AbstractScene.abstract_test()
        self.calls = {"configs": 0, "edges": 0}

    def check_configs(self, q):
        self.calls["configs"] += 1
        return self.sc.batch_flags(np.asarray(q, np.float64))

    def check_edges(self, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False):
        self.calls["edges"] += 1
        return self.sc.batch_edge_flags(q1, q2, resolution, include_endpoints=include_endpoints, N_start=n_start, N_max=n_max, Ns=N)
