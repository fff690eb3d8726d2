"""ctypes binding of the C ABI declared in ``include/petite_b200.h``.

The shared library is built in-tree by ``petite_b200.build.build_library()`` (nvcc, sm_100a).  There is no
CPU fallback: if the library is missing or fails to load, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PETITE_B200_LIB") or os.path.join(_HERE, "libpetite_b200.so")   # override: tuning builds only

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()' "
        "or python -m petite_b200.build).  petite_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)


class pb_config(C.Structure):
    _fields_ = [("Z_T", C.c_double), ("A_T", C.c_double), ("rho", C.c_double), ("dEdx_GeV_per_m", C.c_double),
                ("mT_sampler", C.c_double), ("min_energy", C.c_double), ("Eg_min", C.c_double), ("Ee_min", C.c_double),
                ("maxF_fudge", C.c_double), ("rescale_MCS", C.c_double), ("min_calc", C.c_double * 5),
                ("max_sweeps", C.c_int64),
                ("mV", C.c_double), ("g_e", C.c_double), ("kinetic_mixing", C.c_double), ("Zeff", C.c_double),
                ("E_res_ann", C.c_double), ("E_thr_comp", C.c_double),
                ("bound_electron", C.c_int32), ("reserved", C.c_int32)]


class pb_stack(C.Structure):
    _fields_ = [("p0", C.c_void_p), ("r0w", C.c_void_p), ("pf", C.c_void_p), ("rf", C.c_void_p),
                ("ids", C.c_void_p), ("aux", C.c_void_p), ("capacity", C.c_int64)]


class pb_primaries(C.Structure):
    _fields_ = [("p", c_double_p), ("r", c_double_p), ("weight", c_double_p), ("mass", c_double_p),
                ("pid", c_int32_p), ("flags", c_int32_p), ("n", C.c_int64), ("on_device", C.c_int64)]


class pb_counters(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_particles", "n_waves", "n_steps", "n_substeps", "n_samples", "n_trials",
                                         "n_no_sample", "n_launches", "max_wave", "n_charged")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class pb_dark_tables(C.Structure):
    _fields_ = [("w_E", c_double_p * 4), ("w_y", c_double_p * 4), ("w_n", C.c_int32 * 4),
                ("nsdark_comp_lx", c_double_p), ("nsdark_comp_ly", c_double_p), ("nsdark_comp_n", C.c_int32), ("pad", C.c_int32),
                ("d_E", c_double_p * 4), ("d_table", c_double_p * 4), ("d_n", C.c_int32 * 4), ("min_E", C.c_double * 4)]


class pb_profile(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_int64 * 8), ("trials", C.c_int64 * 16), ("samples", C.c_int64 * 16)]


KERNEL_NAMES = ["k_init_primaries", "k_loop", "k_bucket_scan", "k_bucket_fill", "k_sample", "k_emit", "k_finalize"]
TALLY_SIZE, TALLY_NSPECIES, TALLY_EBINS, TALLY_TBINS = 1024, 7, 64, 32
TALLY_COUNT, TALLY_WSUM, TALLY_WESUM, TALLY_EHIST, TALLY_THIST = 0, 8, 16, 32, 32 + 7 * 64

pb_engine = C.c_void_p

# every symbol include/petite_b200.h declares
SIGNATURES = {
    "pb_version": (C.c_char_p, []),
    "pb_last_error": (C.c_char_p, [pb_engine]),
    "pb_create": (C.c_int, [C.POINTER(pb_engine), C.c_int, C.POINTER(pb_config)]),
    "pb_destroy": (None, [pb_engine]),
    "pb_set_config": (C.c_int, [pb_engine, C.POINTER(pb_config)]),
    "pb_upload_nsigma": (C.c_int, [pb_engine, C.c_int, c_double_p, c_double_p, C.c_int]),
    "pb_upload_maps": (C.c_int, [pb_engine, C.c_int, c_double_p, C.c_int, C.c_int, c_int32_p, c_double_p, c_double_p, C.c_int]),
    "pb_run_showers": (C.c_int, [pb_engine, C.POINTER(pb_primaries), C.c_uint64, C.c_uint64, C.c_int,
                                 C.POINTER(pb_stack), C.POINTER(pb_counters), C.c_void_p]),
    "pb_upload_dark": (C.c_int, [pb_engine, C.POINTER(pb_dark_tables)]),
    "pb_run_dark": (C.c_int, [pb_engine, C.POINTER(pb_stack), C.c_int64, C.c_uint32, C.POINTER(pb_stack),
                              C.POINTER(pb_counters), C.c_void_p]),
    "pb_draw_samples": (C.c_int, [pb_engine, C.c_int, c_double_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, c_double_p,
                                  c_int32_p, C.c_void_p]),
    "pb_find_max": (C.c_int, [pb_engine, C.c_int, C.c_int, C.c_uint64, C.c_double, c_double_p, c_double_p]),
    "pb_train_accumulate": (C.c_int, [pb_engine, C.c_int, c_double_p, C.c_int, C.c_int, c_int32_p, c_double_p, C.c_int64, C.c_uint64,
                                      C.c_double, c_double_p, c_double_p, c_double_p]),
    "pb_train_accumulate_p": (C.c_int, [pb_engine, C.c_int, c_double_p, C.c_int, C.c_int, c_int32_p, c_double_p, C.c_int64, C.c_uint64,
                                        C.c_double, C.c_double, c_double_p, c_double_p, c_double_p]),
    "pb_tally": (C.c_int, [pb_engine, C.POINTER(pb_stack), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "pb_detector_cut": (C.c_int, [pb_engine, C.POINTER(pb_stack), C.c_int64, C.c_int64, c_double_p, C.c_int, C.c_double, C.c_double,
                                  C.c_double, C.c_double, c_double_p, c_double_p, C.c_void_p, C.c_void_p]),
    "pb_set_profiling": (C.c_int, [pb_engine, C.c_int]),
    "pb_get_profile": (C.c_int, [pb_engine, C.POINTER(pb_profile)]),
    "pb_measure_fp64_peak": (C.c_int, [pb_engine, c_double_p]),
    "pb_probe": (C.c_int, [pb_engine, C.c_int, C.c_int, c_double_p, C.c_int64, C.c_int, c_double_p, C.c_int]),
    "pb_replay": (C.c_int, [pb_engine, C.c_int64, c_double_p, c_double_p, C.POINTER(C.c_int64), c_double_p]),
    "pb_quad_batch": (C.c_int, [pb_engine, C.c_int, c_int32_p, C.POINTER(c_double_p), C.POINTER(c_double_p), c_double_p, C.c_double,
                                C.c_void_p, C.c_int64, c_double_p, c_double_p, c_int32_p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

PB_OK, PB_ERR_CUDA, PB_ERR_ARG, PB_ERR_CAPACITY, PB_ERR_STATE, PB_ERR_NO_SAMPLE = 0, -1, -2, -3, -4, -5
PB_FLAG_SHORT_LIVED, PB_FLAG_NO_SAMPLE, PB_FLAG_LONG_LIVED = 1, 2, 4
PROBE_DSIGMA, PROBE_NSIGMA, PROBE_MAP, PROBE_MCS, PROBE_KIN, PROBE_PHILOX, PROBE_HOTMATH, PROBE_MCS_FAST, PROBE_SUBSTEP, PROBE_DARKKIN, PROBE_PROPAGATE = range(11)


class EngineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"petite_b200 engine error {code}: {message}")
        self.code = code


def check(engine, rc):
    if rc != PB_OK:
        raise EngineError(rc, lib.pb_last_error(engine).decode())


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int32_p)
