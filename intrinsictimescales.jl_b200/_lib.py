"""ctypes binding of libitjl_b200.so (C ABI: include/itjl_b200.h).

There is no CPU fallback: if the shared library has not been built, or no CUDA
device is present when a context is created, this module raises.
"""
import ctypes
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_PKG, "libitjl_b200.so")
_lib = None


class ItjlError(RuntimeError):
    """Non-zero status from libitjl_b200 (message from itjl_last_error)."""

    def __init__(self, code, message):
        super().__init__(f"itjl_b200 status {code}: {message}")
        self.code = code
        self.message = message


class ModelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("summary", ctypes.c_int32),
                ("distance", ctypes.c_int32), ("update", ctypes.c_int32),
                ("n_trials", ctypes.c_int64), ("n_t", ctypes.c_int64),
                ("dt", ctypes.c_double), ("data_mean", ctypes.c_double),
                ("data_sd", ctypes.c_double), ("n_stats", ctypes.c_int64),
                ("target_stats", ctypes.c_void_p), ("missing_mask", ctypes.c_void_p),
                ("freq_bins", ctypes.c_void_p)]


class Tuning(ctypes.Structure):
    """itjl_tuning (include/itjl_b200.h): per-context knobs; the library never reads the environment."""
    _fields_ = [("abc_tc", ctypes.c_int32), ("tc_terms", ctypes.c_int32), ("tc_tail_max", ctypes.c_int32),
                ("tc_groups", ctypes.c_int32), ("tc_seg_trials", ctypes.c_int32), ("tc_bias", ctypes.c_int32),
                ("abc_threads", ctypes.c_int32), ("abc_bufs", ctypes.c_int32), ("acw_stream", ctypes.c_int32),
                ("acw_waves", ctypes.c_int32), ("pinned_upload", ctypes.c_int32), ("copy_threads", ctypes.c_int32),
                ("pmc_fp32", ctypes.c_int32), ("psd_v2", ctypes.c_int32), ("tc_error_bound", ctypes.c_double)]


KIND_OU, KIND_OU_OSC = 0, 1
SUMMARY_ACF, SUMMARY_ACF_MISSING, SUMMARY_PSD, SUMMARY_PGRAM = 0, 1, 2, 3
DIST_LINEAR, DIST_LOG = 0, 1
UPDATE_EXACT, UPDATE_EM = 0, 1
ACW0, ACW50, ACWEULER, AUC, TAU, TAU3 = 1, 2, 4, 8, 16, 32

_vp, _i32, _i64, _u32, _u64, _f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64,
                                      ctypes.c_uint32, ctypes.c_uint64, ctypes.c_double)

# name -> (restype, argtypes): every symbol include/itjl_b200.h declares
SIGNATURES = {
    "itjl_version": (ctypes.c_char_p, []),
    "itjl_device_count": (ctypes.c_int, [_vp]),
    "itjl_ctx_create": (ctypes.c_int, [ctypes.c_int, _vp]),
    "itjl_ctx_destroy": (ctypes.c_int, [_vp]),
    "itjl_last_error": (ctypes.c_char_p, [_vp]),
    "itjl_ctx_synchronize": (ctypes.c_int, [_vp]),
    "itjl_ctx_stream": (_u64, [_vp]),
    "itjl_ctx_launch_count": (_i64, [_vp]),
    "itjl_ctx_sm_count": (ctypes.c_int, [_vp]),
    "itjl_model_create": (ctypes.c_int, [_vp, _vp, _vp]),
    "itjl_model_destroy": (ctypes.c_int, [_vp]),
    "itjl_abc_distances": (ctypes.c_int, [_vp, _vp, _i64, _i32, _u64, _u64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "itjl_abc_generation": (ctypes.c_int, [_vp, _vp, _i64, _i32, _u64, _u64, _i64, _f64, _i64, _vp, _vp, _vp, _vp]),
    "itjl_abc_accept": (ctypes.c_int, [_vp, _vp, _i64, _f64, _i64, _vp, _vp, _vp]),
    "itjl_abc_distances_dev": (ctypes.c_int, [_vp, _vp, _i64, _i32, _u64, _u64, _i64, _vp]),
    "itjl_abc_distances_sharded_dev": (ctypes.c_int, [_vp, _vp, _i64, _i32, _u64, _u64, _i64, _vp]),
    "itjl_model_work": (ctypes.c_int, [_vp, _vp, _vp]),
    "itjl_generate_ou": (ctypes.c_int, [_vp, _f64, _f64, _f64, _i64, _i64, _i32, _i32, _u64, _vp, _vp, _vp]),
    "itjl_generate_ou_osc": (ctypes.c_int, [_vp, _f64, _f64, _f64, _f64, _i64, _i64, _f64, _f64, _i32, _u64,
                                            _vp, _vp, _vp, _vp]),
    "itjl_acf": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "itjl_acf_missing": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "itjl_psd_hamming": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _vp]),
    "itjl_acw": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _u32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "itjl_fit_expdecay": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _vp]),
    "itjl_fit_expdecay_3_parameters": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _vp]),
    "itjl_model_kernel_info": (ctypes.c_int, [_vp, _vp, _vp]),
    "itjl_tc_plan": (ctypes.c_int, [_i64, _i64, _i64, _i32, _i32, _vp]),
    "itjl_fp32_peak": (ctypes.c_int, [_vp, _i32, _vp]),
    "itjl_ctx_timing_reset": (ctypes.c_int, [_vp]),
    "itjl_ctx_timing_get": (ctypes.c_int, [_vp, _vp, _vp]),
    "itjl_tc_probe": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i32, _u32, _u32, _u32, _vp]),
    "itjl_tuning_default": (ctypes.c_int, [_vp]),
    "itjl_ctx_set_tuning": (ctypes.c_int, [_vp, _vp]),
    "itjl_ctx_get_tuning": (ctypes.c_int, [_vp, _vp]),
    "itjl_ctx_create_multi": (ctypes.c_int, [ctypes.c_int, _vp, _vp]),
    "itjl_comm_unique_id": (ctypes.c_int, [_vp]),
    "itjl_ctx_attach_comm": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp]),
    "itjl_ctx_world": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "itjl_pmc_weights": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp]),
    "itjl_weighted_covar": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "itjl_tc_predicted_error": (_f64, [_i64, _i64, _i32, _i32, _f64, _f64]),
    "itjl_acf_strided": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i32, _vp]),
    "itjl_acw_strided": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f64, _u32, _i32, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _i64]),
    "itjl_data_upload": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64]),
    "itjl_acw_resident": (ctypes.c_int, [_vp, _f64, _u32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "itjl_acw_acf": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "itjl_ctx_timing_stop": (ctypes.c_int, [_vp]),
    "itjl_psd_periodogram": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _vp]),
    "itjl_nextfastfft": (_i64, [_i64]),
    "itjl_psd_lombscargle": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _f64, _f64, _i64, _vp]),
    "itjl_fit_knee": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _f64, _f64, _i32, _i32, _vp, _vp, _vp]),
    "itjl_acw_knee": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f64, _f64, _f64, _u32, _i32, _vp, _vp, _i64, _vp, _vp]),
    "itjl_acw_knee_nfreq": (_i64, [_i64, _i32]),
    "itjl_abc_distances_combined": (ctypes.c_int, [_vp, _vp, _i64, _i32, _u64, _u64, _i64, _vp, _vp, _i32, _f64, _f64, _vp]),
    "itjl_pmc_propose": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _f64, _i64, _u64, _u64, _i64, _vp]),
    "itjl_order_stats": (ctypes.c_int, [_vp, _vp, _i64, _f64, _vp, _i32, _vp, _vp]),
    "itjl_generate_ou_two": (ctypes.c_int, [_vp, _f64, _f64, _f64, _f64, _i64, _i64, _f64, _i32, _u64, _vp]),
}


def library_path():
    return _SO


def build(verbose=False):
    """Compile libitjl_b200.so in-tree with nvcc for sm_100a (csrc/Makefile)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j8"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return _SO


def lib():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(
                f"{_SO} is missing: build it with `make -C {os.path.join(_PKG, 'csrc')}` "
                "(python __graft_entry__.py build). There is no CPU fallback.")
        l = ctypes.CDLL(_SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def ptr(a):
    """Host pointer of a numpy array (None -> NULL)."""
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["C" if order == "C" else "F", "ALIGNED"])


class Context:
    """One GPU + one stream (itjl_ctx), or - `devices=[...]` - one process driving several GPUs with the particles
    of a generation sharded inside the library (itjl_ctx_create_multi).  Calls on one context are serialised by
    the library's own lock."""

    def __init__(self, device=0, devices=None):
        self._h = ctypes.c_void_p()
        if devices is not None:
            ids = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            rc = lib().itjl_ctx_create_multi(len(devices), ids, ctypes.byref(self._h))
            device = devices[0]
        else:
            rc = lib().itjl_ctx_create(int(device), ctypes.byref(self._h))
        if rc != 0:
            raise ItjlError(rc, lib().itjl_last_error(None).decode())
        self.device = int(device)
        import weakref
        self._models = weakref.WeakSet()        # itjl_model handles created on this context: destroyed before it

    # ---- tuning (itjl_tuning) ----------------------------------------------------------
    def get_tuning(self):
        t = Tuning()
        self.check(lib().itjl_ctx_get_tuning(self._h, ctypes.byref(t)))
        return t

    def set_tuning(self, **kw):
        """Change knobs of itjl_tuning by name, e.g. set_tuning(abc_tc=0).  A model reads them when it is created."""
        t = self.get_tuning()
        for k, v in kw.items():
            if k not in dict(Tuning._fields_):
                raise KeyError(f"unknown tuning field {k!r}")
            setattr(t, k, v)
        self.check(lib().itjl_ctx_set_tuning(self._h, ctypes.byref(t)))

    def reset_tuning(self):
        t = Tuning()
        lib().itjl_tuning_default(ctypes.byref(t))
        self.check(lib().itjl_ctx_set_tuning(self._h, ctypes.byref(t)))

    def tuned(self, **kw):
        """Context manager: the given knobs inside the block, the previous values after it."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            old = self.get_tuning()
            self.set_tuning(**kw)
            try:
                yield self
            finally:
                self.check(lib().itjl_ctx_set_tuning(self._h, ctypes.byref(old)))
        return cm()

    # ---- communicator ------------------------------------------------------------------
    def attach_comm(self, rank, world, unique_id):
        """Join an NCCL communicator of `world` processes (one GPU each): collective (ncclCommInitRank)."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self.check(lib().itjl_ctx_attach_comm(self._h, int(rank), int(world), buf))

    def world(self):
        """(rank, world size, devices driven by this process)."""
        r, w, n = ctypes.c_int(0), ctypes.c_int(1), ctypes.c_int(1)
        self.check(lib().itjl_ctx_world(self._h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(n)))
        return r.value, w.value, n.value

    def timing_stop(self):
        self.check(lib().itjl_ctx_timing_stop(self._h))

    def check(self, rc):
        if rc != 0:
            raise ItjlError(rc, lib().itjl_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        self.check(lib().itjl_ctx_synchronize(self._h))

    @property
    def stream(self):
        return int(lib().itjl_ctx_stream(self._h))

    @property
    def launch_count(self):
        return int(lib().itjl_ctx_launch_count(self._h))

    @property
    def sm_count(self):
        return int(lib().itjl_ctx_sm_count(self._h))

    def timing_reset(self):
        self.check(lib().itjl_ctx_timing_reset(self._h))

    def timing_get(self):
        n = ctypes.c_int64(0)
        ms = ctypes.c_double(0.0)
        self.check(lib().itjl_ctx_timing_get(self._h, ctypes.byref(n), ctypes.byref(ms)))
        return n.value, ms.value

    def fp32_peak(self, iters=5):
        out = ctypes.c_double(0.0)
        self.check(lib().itjl_fp32_peak(self._h, int(iters), ctypes.byref(out)))
        return out.value

    def close(self):
        if self._h:
            for m in list(getattr(self, "_models", ())):     # a model holds a pointer to its context
                m.close()
            lib().itjl_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks)."""
    buf = ctypes.create_string_buffer(128)
    rc = lib().itjl_comm_unique_id(buf)
    if rc != 0:
        raise ItjlError(rc, lib().itjl_last_error(None).decode())
    return buf.raw


def tc_predicted_error(n_t, n_trials, n_terms, seg_rows, mask_frac=0.0, mean_ratio=0.0):
    """The error model of the tensor-core ACF kernel (itjl_tc_predicted_error; host only)."""
    return float(lib().itjl_tc_predicted_error(int(n_t), int(n_trials), int(n_terms), int(seg_rows), float(mask_frac),
                                               float(mean_ratio)))


_default_ctx = {}


def default_context(device=None):
    """Process-wide context for `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
