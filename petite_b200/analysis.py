"""On-device analysis of a shower batch: the reference's ``detector_cut`` (src/PETITE/shower.py:815-864).

``detector_cut(batch, owner, detector_positions, detector_radius, method, energy_cut, detector_inner_radius)`` keeps the
reference's argument meaning: every entry of ``detector_positions`` is used as the z of a plane (that is what
``transverse_position(p0, detector_positions)`` does with the list it is handed, shower.py:852), particles are
extrapolated in a straight line from their creation point, and a particle passes if its transverse radius lies in
(detector_inner_radius, detector_radius).  ``batch`` is a ShowerBatch or DarkBatch still resident in HBM.
"""
import ctypes as C

import numpy as np

from . import _capi as capi


def transverse_position(p0, r0, z):
    """(x, y) of a straight track at plane z (shower.py:815-823)."""
    T = (z - r0[2]) / p0[3]
    return [r0[0] + T * p0[1], r0[1] + T * p0[2]]


def detector_cut(batch, owner, detector_positions, detector_radius, method="Sample", energy_cut=None, detector_inner_radius=0.0):
    torch = owner._torch
    z = np.ascontiguousarray(np.atleast_1d(detector_positions), dtype=np.float64)
    nd = len(z)
    lo, hi = (-np.inf, np.inf) if energy_cut is None else (float(energy_cut[0]), float(energy_cut[1]))
    t = batch._t
    from .shower import stack_struct
    st = stack_struct(t)
    wpass, wall = np.zeros(nd), np.zeros(1)
    want_mask = method in ("Sample", "SampleW")
    mask = torch.zeros((max(batch.n, 1), nd), dtype=torch.uint8, device=t["p0"].device) if want_mask else None
    stream = torch.cuda.current_stream(owner._device).cuda_stream
    capi.check(owner._engine, capi.lib.pb_detector_cut(
        owner._engine, C.byref(st), 0, batch.n, capi.dptr(z), nd, float(detector_radius), float(detector_inner_radius), lo, hi,
        capi.dptr(wpass), capi.dptr(wall), C.c_void_p(mask.data_ptr() if want_mask else 0), C.c_void_p(stream)))
    if method == "TotalWeight":
        return [float(v) for v in wpass]
    if method == "Efficiency":
        return [float(v / wall[0]) for v in wpass] if wall[0] != 0 else [0.0] * nd
    m = mask[: batch.n].cpu().numpy().astype(bool)
    if method == "SampleW":
        return m.T
    if method == "Sample":
        return [np.nonzero(m[:, k])[0] for k in range(nd)]       # record indices passing each detector
    raise ValueError(method)
