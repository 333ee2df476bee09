"""The margin argument of the parity bar, shared by the GPU tests (BASELINE.json north_star: flags identical for
every sample whose signed clearance exceeds 1e-5).  An edge result (free flag, first colliding position in the
reference's binary order) may differ from the oracle's ONLY if one of the interpolated configurations the reference
would have visited lies inside the margin; no percentage thresholds."""
import numpy as np

from oracle import oracle_scene as O

MARGIN = 1e-5


def edge_disagreements_are_inside_margin(cs, q1, q2, resolution, free, first, ofree, ofirst, Ns=None, n_start=0, n_max=None,
                                         include_endpoints=False, tol=None):
    """Asserts the rule above for one batch of edges; returns the number of edges that differed (all of them excused by
    a margin sample).  q1, q2: fp32 arrays as handed to the device."""
    tol = cs.tol if tol is None else tol
    free, first, ofree, ofirst = (np.asarray(x) for x in (free, first, ofree, ofirst))
    differ = np.nonzero((free.astype(bool) != ofree.astype(bool)) | (first != ofirst))[0]
    for e in differ:
        a, b = q1[e].astype(np.float64), q2[e].astype(np.float64)
        N = int(Ns[e]) if Ns is not None else max(2, int(np.max(np.abs(a - b)) / resolution) + 1)
        order = O.binary_indices(N)
        hi = N if n_max is None or n_max < 0 else min(n_max, N)
        idx = order[n_start:hi]
        if not include_endpoints:
            idx = idx[(idx != 0) & (idx != N - 1)]
        assert len(idx), f"edge {e}: results differ although its window holds no sample"
        qs = a[None] + ((b - a) / (N - 1))[None] * idx.astype(np.float64)[:, None]
        _, p, md = O.check_configs(cs.blob64, qs, tol)
        m = np.min(np.abs(O.margin(p, md, tol)))
        assert m <= MARGIN, (f"edge {e}: device (free={bool(free[e])}, first={first[e]}) vs oracle (free={bool(ofree[e])}, "
                             f"first={ofirst[e]}) with every sample margin-clear (closest {m:.3g})")
    return len(differ)
