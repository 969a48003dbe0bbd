#!/usr/bin/env python
"""tcgen05 experiments on the B200 (diagnostic): which TMEM lanes an M = 64 MMA writes, whether its
D address accepts a lane offset, and how the fp32 accumulator rounds."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

SBO = 256          # one K step (16 rows x 16 B) per 8-wide MN block
REGION = 16 * SBO  # 128 MN x 16 K fp16


def idesc(m, n):
    return (1 << 4) | (1 << 15) | (1 << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def put(img, region, mn, k, val):
    off = region * REGION + (mn // 8) * SBO + k * 16 + (mn % 8) * 2
    img[off // 2] = val


def run(pkg, ctx, img, ops, idsc):
    out = np.zeros((128, 512), dtype=np.float32)
    ops = np.ascontiguousarray(np.asarray(ops, dtype=np.int32).reshape(-1, 4))
    rc = pkg._lib.lib().itjl_tc_probe(ctx.handle, img.ctypes.data_as(ctypes.c_void_p), img.nbytes,
                                      ops.ctypes.data_as(ctypes.c_void_p), len(ops), 128, SBO, idsc,
                                      out.ctypes.data_as(ctypes.c_void_p))
    ctx.check(rc)
    return out


def describe(out, label):
    nz = np.argwhere(out != 0)
    if len(nz) == 0:
        print(label, ": all zero"); return
    lanes = sorted(set(nz[:, 0].tolist())); cols = sorted(set(nz[:, 1].tolist()))
    print(label, ": lanes", lanes[:8], "...", lanes[-8:], f"({len(lanes)} lanes)  cols {cols[0]}..{cols[-1]} ({len(cols)})")
    print("   lane->value at first col:", {int(l): float(out[l, cols[0]]) for l in lanes[:4] + lanes[14:20] + lanes[-2:]})


def main():
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    img = np.zeros(4 * REGION // 2, dtype=np.float16)
    for a in range(128):
        put(img, 0, a, 0, a + 1)          # region 0: A[k=0][a] = a + 1
        put(img, 1, a, 0, 1.0)            # region 1: B[k=0][b] = 1
        put(img, 2, a, 0, 4096.0)         # region 2: 4096
        put(img, 3, a, 0, 4096.0 if a % 2 == 0 else -4096.0)   # region 3: +-4096
    A, B = 0, REGION
    for label, ops, ids in [
        ("M=128 N=32", [[A, B, 0, 0]], idesc(128, 32)),
        ("M=64 N=32 lane 0", [[A, B, 0, 0]], idesc(64, 32)),
        ("M=64 N=32 A rows 64..127", [[A + 8 * SBO, B, 0, 0]], idesc(64, 32)),
    ]:
        try:
            out = run(pkg, ctx, img, ops, ids)
            describe(out, label)
        except Exception as e:  # noqa: BLE001
            print(label, "FAILED:", e)
            ctx = pkg.default_context(0)

    # accumulator rounding: D = +-2^24, then add +-x
    img2 = img.copy()
    for x in (1.75, 1.25, 1.0, 0.75, 3.0):
        for a in range(128):
            put(img2, 0, a, 0, x)                           # A2 = x
            put(img2, 1, a, 0, 1.0 if a % 2 == 0 else -1.0)  # B2 = +-1
        out = run(pkg, ctx, img2, [[2 * REGION, 3 * REGION, 0, 0], [A, B, 0, 1]], idesc(128, 32))
        print(f"2^24 + {x}: {out[0, 0]:.1f}   -2^24 - {x}: {out[0, 1]:.1f}   (RN would give "
              f"{np.float32(2.0**24 + x):.1f} / {np.float32(-2.0**24 - x):.1f})")
    # products inside one instruction: 2^24 + sixteen 0.25s in one K step (exact sum 4 -> visible if summed before rounding)
    img3 = img.copy()
    for a in range(128):
        for k in range(16):
            put(img3, 0, a, k, 0.25)
            put(img3, 1, a, k, 1.0)
    out = run(pkg, ctx, img3, [[2 * REGION, 3 * REGION, 0, 0], [A, B, 0, 1]], idesc(128, 32))
    print("2^24 + 16 x 0.25 in one MMA:", out[0, 0], " (-2^24 + 4):", out[0, 1])


def risky():
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    img = np.zeros(4 * REGION // 2, dtype=np.float16)
    for a in range(128):
        put(img, 0, a, 0, a + 1)
        put(img, 1, a, 0, 1.0)
    A, B = 0, REGION
    for label, ops, ids in [
        ("M=64 N=32 lane offset 16", [[A, B, (16 << 16) | 0, 0]], idesc(64, 32)),
        ("M=64 N=32 lane offset 16 col 64", [[A, B, (16 << 16) | 64, 0]], idesc(64, 32)),
        ("M=64 N=32 two halves", [[A, B, 0, 0], [A + 8 * SBO, B, (16 << 16) | 0, 0]], idesc(64, 32)),
        ("M=128 N=32 lane offset 16 (expected invalid)", [[A, B, (16 << 16) | 0, 0]], idesc(128, 32)),
    ]:
        try:
            out = run(pkg, ctx, img, ops, ids)
            describe(out, label)
        except Exception as e:  # noqa: BLE001
            print(label, "FAILED:", e, flush=True)
            return


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "risky":
        risky()
        sys.exit(0)
    main()
