#!/usr/bin/env python
"""Discover / confirm the tcgen05 shared-memory descriptor conventions on the B200 (diagnostic;
the asserted version lives in tests/test_gpu_tc_probe.py)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

T = 128


def idesc_f16_mn(m, n):
    return (1 << 4) | (1 << 15) | (1 << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def build_image(x, rows_total):
    """series x (ints) -> fp16 image in the MN-major no-swizzle layout, K rows 16 B apart."""
    sbo = rows_total * 16
    img = np.zeros(16 * sbo // 2, dtype=np.float16)
    n = len(x)
    for j in range(n):
        i, a = divmod(j, T)
        off = (a // 8) * sbo + i * 16 + (a % 8) * 2
        img[off // 2] = x[j]
    return img, sbo


def run(pkg, ctx, img, ops, lbo, sbo, idesc):
    out = np.zeros((128, 512), dtype=np.float32)
    ops = np.ascontiguousarray(np.asarray(ops, dtype=np.int32).reshape(-1, 4))
    lib = pkg._lib.lib()
    rc = lib.itjl_tc_probe(ctx.handle, img.ctypes.data_as(ctypes.c_void_p), img.nbytes,
                           ops.ctypes.data_as(ctypes.c_void_p), len(ops), lbo, sbo, idesc,
                           out.ctypes.data_as(ctypes.c_void_p))
    ctx.check(rc)
    return out


def main():
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    rng = np.random.default_rng(0)
    ksteps, n_shift = 2, 4
    rows = 16 * ksteps
    rows_total = rows + n_shift - 1
    x = rng.integers(-4, 5, size=rows_total * T).astype(np.float64)
    x[rows * T:] = rng.integers(-4, 5, size=(rows_total - rows) * T)     # non-zero tail rows: shifted reads must see them
    img, sbo = build_image(x, rows_total)
    X = x.reshape(rows_total, T)
    expect = np.zeros((128, 512))
    for s in range(n_shift):
        expect[:, 128 * s:128 * (s + 1)] = X[:rows].T @ X[s:s + rows]
    ops = []
    for k in range(ksteps):
        for s in range(n_shift):
            ops.append([k * 16 * 16, (k * 16 + s) * 16, 128 * s, 1 if k > 0 else 0])
    for name, lbo, sbo_f in [("lbo=K-group stride(128), sbo=MN-block stride", 128, sbo),
                             ("swapped", sbo, 128)]:
        try:
            out = run(pkg, ctx, img, ops, lbo, sbo_f, idesc_f16_mn(128, 128))
        except Exception as e:  # noqa: BLE001
            print(name, "FAILED", e)
            continue
        err = np.abs(out - expect).max()
        print(f"{name}: max|out-expect| = {err}  (expect max {np.abs(expect).max()})", flush=True)
        if err != 0:
            bad = np.argwhere(out != expect)
            print("  mismatches:", len(bad), "first", bad[:5].tolist(), out[tuple(bad[0])], expect[tuple(bad[0])])
    # N = 64 on the last tile and a 2-row K offset inside a K step (descriptor start granularity)
    ops2 = [[0, 3 * 16, 0, 0], [16 * 16, (16 + 3) * 16, 0, 1]]
    out = run(pkg, ctx, img, ops2, 128, sbo, idesc_f16_mn(128, 64))
    e2 = X[:rows].T @ X[3:3 + rows]
    print("N=64 shift=3: max err", np.abs(out[:, :64] - e2[:, :64]).max())


if __name__ == "__main__":
    main()
