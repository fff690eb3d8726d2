"""The QAGS restatement (petite_b200/csrc/quadpack.cuh, what pb_quad_batch runs one GPU thread per integral) against
scipy.integrate.quad - the QUADPACK build the reference calls - on the real set-up integrands of a DarkShower: host build of the
same header and the same integrand code (tests/csrc/quad_host.cpp, g++), no GPU.  Same subdivision decisions -> the same numbers
to rounding, including the integrals QAGS gives up on (ier != 0) with errors far above its tolerance."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT


@pytest.fixture(scope="module")
def host_quad(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("quad") / "libquad_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "csrc", "quad_host.cpp")], check=True)
    lib = C.CDLL(out)
    lib.quad_host_batch.restype = C.c_int
    return lib.quad_host_batch


def test_gauss_kronrod_rule_is_exact_for_polynomials(host_quad):
    """(10, 21) Gauss-Kronrod: the 21-point rule integrates polynomials up to degree 31 exactly."""
    from petite_b200 import dark_setup as ds
    from petite_b200.shower import LinearTable
    # f = a straight line through a table: one panel, exact, error estimate 0 -> QAGS returns after 21 evaluations
    t = LinearTable(np.array([0.0, 10.0]), np.array([1.0, 21.0]))
    run = ds.c_abi_runner(host_quad)
    got = run([t], 1.0, np.array([ds._call(0, 0, 1.0, 7.0)], dtype=ds.CALL_DTYPE))
    assert abs(got[0] - (6.0 + (49 - 1))) < 1e-13 and run.last_ier[0] == 1          # ier 0, one interval


@pytest.mark.parametrize("material", ["lead", "graphite"])
def test_set_up_integrals_equal_scipy_quad(host_quad, material):
    from petite_b200 import dark_setup as ds
    from tests.test_dark_setup_cpu import _host_only_dark_shower
    sh = _host_only_dark_shower(material, 0.03)
    seen = {}

    def both(tables, dEdx_m, calls):
        rng = np.random.default_rng(len(calls))
        sel = np.sort(rng.choice(len(calls), size=min(len(calls), 160), replace=False))     # scipy needs ~10 ms per call
        a = ds.c_abi_runner(host_quad)
        got = a(tables, dEdx_m, calls)
        want = ds.scipy_runner(tables, dEdx_m, calls[sel])
        kind = int(calls[0]["kind"])
        scale = np.maximum(np.abs(want), 1e-300)
        rel = np.abs(got[sel] - want) / scale
        seen[kind] = (float(rel.max()), int((a.last_ier[sel] >= 1000).sum()), len(sel))
        assert np.all((rel < 1e-9) | (np.abs(got[sel] - want) < 1e-30)), (kind, float(rel.max()))
        return got
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        ds.build(sh, os.path.join(d, "s.npz"), runner=both)
    print(material, "max rel difference to scipy.quad, calls QAGS flagged (ier != 0), calls compared:", seen)
    assert set(seen) == {0, 1}
