"""Tuning scripts only: translate the ITJL_* environment switches the scripts of this directory use into the
per-context itjl_tuning of the library (which never reads the environment itself).  Call `from_env(ctx)` before
a model is created."""
import os

_MAP = {"ITJL_ABC_TC": "abc_tc", "ITJL_TC_SEG": "tc_seg_trials", "ITJL_TC_TERMS": "tc_terms",
        "ITJL_TC_TAIL_MAX": "tc_tail_max", "ITJL_TC_GROUPS": "tc_groups", "ITJL_TC_BIAS": "tc_bias",
        "ITJL_ABC_THREADS": "abc_threads", "ITJL_ABC_BUFS": "abc_bufs", "ITJL_ACW_STREAM": "acw_stream",
        "ITJL_ACW_WAVES": "acw_waves", "ITJL_PINNED_UPLOAD": "pinned_upload", "ITJL_COPY_THREADS": "copy_threads", "ITJL_PSD_V2": "psd_v2"}


def from_env(ctx):
    ctx.reset_tuning()
    kw = {field: int(os.environ[name]) for name, field in _MAP.items() if name in os.environ}
    if "ITJL_TC_BOUND" in os.environ:
        kw["tc_error_bound"] = float(os.environ["ITJL_TC_BOUND"])
    if kw:
        ctx.set_tuning(**kw)
    return ctx
