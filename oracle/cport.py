"""ctypes binding of oracle/c/itjl_oracle.c (the fast C restatement of the
reference's per-particle loop body; test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libitjl_oracle.so")
_lib = None

SUMMARY = {"acf": 0, "acf_missing": 1, "psd": 2}
DISTANCE = {"linear": 0, "logarithmic": 1}
UPDATE = {"exact": 0, "em": 1}


class ModelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("summary", ctypes.c_int32),
                ("distance", ctypes.c_int32), ("update", ctypes.c_int32),
                ("n_trials", ctypes.c_int64), ("n_t", ctypes.c_int64),
                ("dt", ctypes.c_double), ("data_mean", ctypes.c_double),
                ("data_sd", ctypes.c_double), ("n_stats", ctypes.c_int64),
                ("target_stats", ctypes.c_void_p), ("missing_mask", ctypes.c_void_p),
                ("freq_bins", ctypes.c_void_p)]


def build(force=False):
    src = os.path.join(_HERE, "c", "itjl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()                      # make is a no-op when the .so is up to date
        _lib = ctypes.CDLL(_SO)
        _lib.itjl_oracle_abc_distances.restype = ctypes.c_int
        _lib.itjl_oracle_abc_stats.restype = ctypes.c_int
        _lib.itjl_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().itjl_oracle_max_threads())


def model_arrays(model):
    """Flatten an oracle.models.Model into the C descriptor (+ keep-alive arrays)."""
    keep = {}
    keep["target"] = np.ascontiguousarray(model.data_sum_stats, dtype=np.float64)
    d = ModelDesc()
    if model.kind not in ("one_timescale", "one_timescale_with_missing", "one_timescale_and_osc",
                          "one_timescale_and_osc_with_missing") or getattr(model, "distance_combined", False):
        raise ValueError(f"the C restatement does not cover model kind {model.kind!r} / distance_combined")
    d.kind = 1 if model.kind in ("one_timescale_and_osc", "one_timescale_and_osc_with_missing") else 0
    if model.summary_method == "psd":
        if model.kind.endswith("with_missing"):
            raise ValueError("the C restatement has no Lomb-Scargle PSD")
        d.summary = SUMMARY["psd"]
    elif model.kind.endswith("with_missing"):
        d.summary = SUMMARY["acf_missing"]
    else:
        d.summary = SUMMARY["acf"]
    d.distance = DISTANCE[model.distance_method]
    d.update = UPDATE[model.update]
    d.n_trials = model.numTrials
    d.n_t = model.n_t
    d.dt = model.dt
    d.data_mean = model.data_mean
    d.data_sd = model.data_sd
    d.n_stats = keep["target"].shape[0]
    d.target_stats = keep["target"].ctypes.data
    if model.missing_mask is not None:
        keep["mask"] = np.asfortranarray(model.missing_mask.astype(np.uint8))
        d.missing_mask = keep["mask"].ctypes.data
    if model.freq_idx is not None:
        keep["bins"] = np.ascontiguousarray(np.nonzero(model.freq_idx)[0], dtype=np.int32)
        d.freq_bins = keep["bins"].ctypes.data
    return d, keep


def abc_distances(model, theta, seed=0, generation=0, particle_offset=0, u0=None, noise=None,
                  phases=None, n_threads=0, return_stats=False):
    """theta: (n_particles, p).  Returns float64 distances (and, with return_stats, the
    (n_particles, n_stats) trial-mean summary statistics they were computed from)."""
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    n, p = theta.shape
    th = np.asfortranarray(theta)
    d, keep = model_arrays(model)
    out = np.empty(n, dtype=np.float64)

    def ptr(a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep[id(a)] = a
        return ctypes.c_void_p(a.ctypes.data)

    stats = np.empty((n, int(d.n_stats)), dtype=np.float64) if return_stats else None
    rc = lib().itjl_oracle_abc_stats(ctypes.byref(d), ctypes.c_void_p(th.ctypes.data),
                                     ctypes.c_int64(n), ctypes.c_int32(p),
                                     ctypes.c_uint64(seed), ctypes.c_uint64(generation),
                                     ctypes.c_int64(particle_offset), ptr(u0), ptr(noise),
                                     ptr(phases), ctypes.c_void_p(out.ctypes.data),
                                     ctypes.c_void_p(stats.ctypes.data) if return_stats else None,
                                     ctypes.c_int32(n_threads))
    if rc != 0:
        raise RuntimeError(f"itjl_oracle_abc_stats -> {rc}")
    return (out, stats) if return_stats else out
