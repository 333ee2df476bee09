"""CPU stand-ins for the CUDA devices of env.py, answered by the fp64 oracle.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference,
the CPU arm of the `planners` block) may import this.  It is what "the reference's CPU backend" means wherever rai
itself cannot run: the same environment classes, the same unmodified reference planners, with every
collision query answered on the host, one query per call like rai (P/problems/rai_base_env.py:442-676), single
threaded like the reference.  Never imported by the product package."""
import time

import numpy as np

from oracle import oracle_abstract as OA
from oracle import oracle_scene as O


class OracleSceneDevice:
    def __init__(self, nthreads: int = 1):
        self.cs = {}
        self.nthreads = nthreads
        self.calls = {"configs": 0, "edges": 0, "robot": 0}
        self.items = {"configs": 0, "edges": 0, "robot": 0}
        self.seconds = 0.0

    def set_mode(self, slot, cs):
        self.cs[slot] = cs

    def check_configs(self, slot, q, tol=None):
        t = time.perf_counter()
        self.calls["configs"] += 1
        q = np.asarray(q, np.float32).astype(np.float64)
        self.items["configs"] += len(q)
        out = O.check_configs(self.cs[slot].blob64, q, -1.0 if tol is None else tol, nthreads=self.nthreads)[0]
        self.seconds += time.perf_counter() - t
        return out

    def check_configs_for_robot(self, slot, q, rel, oth, tol=None):
        t = time.perf_counter()
        self.calls["robot"] += 1
        q = np.asarray(q, np.float32).astype(np.float64)
        self.items["robot"] += len(q)
        out = O.check_configs(self.cs[slot].blob64, q, -1.0 if tol is None else tol, rel=rel, oth=oth, nthreads=self.nthreads)[0]
        self.seconds += time.perf_counter() - t
        return out

    def check_edges(self, slot, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False, tol=None):
        t = time.perf_counter()
        self.calls["edges"] += 1
        q1 = np.asarray(q1, np.float32).astype(np.float64)
        q2 = np.asarray(q2, np.float32).astype(np.float64)
        self.items["edges"] += len(q1)
        f, p, _ = O.check_edges(self.cs[slot].blob64, q1, q2, resolution, Ns=N, n_start=n_start,
                                n_max=-1 if n_max is None else n_max, include_endpoints=include_endpoints,
                                tol=-1.0 if tol is None else tol, nthreads=self.nthreads)
        self.seconds += time.perf_counter() - t
        return f, p

    def __deepcopy__(self, memo):
        return self


class OracleAbstractDevice:
    def __init__(self):
        self.sc = OA.AbstractScene.abstract_test()
        self.calls = {"configs": 0, "edges": 0}
        self.seconds = 0.0

    def check_configs(self, q):
        t = time.perf_counter()
        self.calls["configs"] += 1
        out = self.sc.batch_flags(np.asarray(q, np.float64))
        self.seconds += time.perf_counter() - t
        return out

    def check_edges(self, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False):
        t = time.perf_counter()
        self.calls["edges"] += 1
        out = self.sc.batch_edge_flags(q1, q2, resolution, include_endpoints=include_endpoints, N_start=n_start, N_max=n_max, Ns=N)
        self.seconds += time.perf_counter() - t
        return out

    def __deepcopy__(self, memo):
        return self


class OraclePrefetchDevice(OracleSceneDevice):
    """the same CPU device with the asynchronous edge-batch seam of the CUDA device (mrb200_submit_edges_host /
    mrb200_collect_edges_host), answered synchronously: lets the CPU suite exercise env.py's candidate-edge speculation"""

    def __init__(self, nthreads: int = 1):
        super().__init__(nthreads)
        self.tickets = {}
        self.next_ticket = 0
        self.calls["prefetch"] = 0
        self.items["prefetch"] = 0

    def prefetch_edges(self, slot, q1, q2, resolution, N=None, include_endpoints=False, tol=None):
        q2 = np.asarray(q2, np.float32)
        q1 = np.asarray(q1, np.float32).reshape(-1, q2.shape[1])
        if len(q1) == 1:
            q1 = np.repeat(q1, len(q2), 0)
        t = self.next_ticket
        self.next_ticket += 1
        n_edges = self.calls["edges"]
        self.tickets[t] = self.check_edges(slot, q1, q2, resolution, N=N, include_endpoints=include_endpoints, tol=tol)
        self.calls["edges"] = n_edges
        self.items["edges"] -= len(q2)
        self.calls["prefetch"] += 1
        self.items["prefetch"] += len(q2)
        return t

    def collect_edges(self, ticket, E):
        f, p = self.tickets.pop(ticket)
        assert len(f) == E
        return f, p
