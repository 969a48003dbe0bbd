#!/usr/bin/env python
"""Writes the fixed inputs of the reference-made golden vectors: tests/golden/reference_inputs_v1/*.f64
(raw little-endian Float64, COLUMN-MAJOR - what Julia's `read!` into an `Array{Float64}` expects) and
`manifest.txt` (name, dims).  NumPy only (no oracle, no GPU): anybody can regenerate the identical bytes.

    python tests/golden/make_reference_inputs.py            # rewrites the inputs (deterministic)
    julia --project=/path/to/IntrinsicTimescales.jl tests/golden/make_reference_golden.jl
                                                            # -> tests/golden/reference_v1.json
    python -m pytest tests/test_reference_golden.py         # oracle (CPU) and CUDA path (GPU) against it

The Julia step needs the reference package and cannot run in this image (no julia binary); until someone
runs it once, `reference_v1.json` is absent, the loader tests skip and parity stays "unpinned".
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "reference_inputs_v1")
FS = 500.0


def ou_rows(rng, n_rows, n_t, tau, dt):
    a = np.exp(-dt / tau)
    s = np.sqrt(tau / 2 * (1 - a * a))
    x = np.empty((n_rows, n_t))
    prev = rng.standard_normal(n_rows)
    for k in range(n_t):
        prev = a * prev + s * rng.standard_normal(n_rows)
        x[:, k] = prev
    return (x - x.mean(axis=1, keepdims=True)) / x.std(axis=1, ddof=1, keepdims=True)


def inputs():
    rng = np.random.default_rng(20261020)
    dt = 1.0 / FS
    out = {}
    x = ou_rows(rng, 4, 1200, 0.05, dt)                       # (trials x time), tau = 50 ms
    out["x_complete"] = x
    xm = x.copy()
    xm[0, 100:220] = np.nan                                    # contiguous gaps + scattered samples
    xm[1, 900:1020] = np.nan
    xm[2, rng.integers(0, 1200, 60)] = np.nan
    out["x_missing"] = xm
    t = dt * np.arange(1, 1201)
    osc = np.sqrt(2.0) * np.sin(2 * np.pi * 10.0 * t[None, :] + rng.uniform(0, 2 * np.pi, size=(4, 1)))
    y = np.sqrt(0.3) * osc + np.sqrt(0.7) * ou_rows(rng, 4, 1200, 0.05, dt)
    out["x_osc"] = 1.5 * (y - y.mean(axis=1, keepdims=True)) / y.std(axis=1, ddof=1, keepdims=True) + 0.3
    xw = rng.standard_normal((48, 1000))                       # short-memory acw regime
    # acw(:auc) integrates every slice over the FIRST slice's zero-crossing window and throws when that
    # window has one point (acw.jl:194 + Romberg): put a slice whose lag-1 and lag-2 autocorrelations are both
    # positive first
    def head_positive(r):
        c = r - r.mean()
        return all(np.dot(c[:-k], c[k:]) > 0 for k in (1, 2))
    first = next(i for i in range(48) if head_positive(xw[i]))
    xw[[0, first]] = xw[[first, 0]]
    out["x_white"] = xw
    lags = dt * np.arange(60)
    taus = np.array([0.01, 0.03, 0.08, 0.2])
    acf = np.exp(-lags[None, :] / taus[:, None]) + 0.01 * rng.standard_normal((4, 60))
    acf[:, 0] = 1.0
    out["acf_rows"] = acf                                      # fit_expdecay / acw_romberg inputs
    out["distances"] = np.concatenate([rng.gamma(2.0, 0.05, 500), [np.nan] * 7, [11.0, 25.0]])
    th_prev = np.stack([rng.uniform(0.05, 1.0, 50), rng.normal(10.0, 2.0, 50)], axis=1)
    out["theta_prev"] = th_prev
    out["theta_new"] = np.stack([rng.uniform(0.05, 1.0, 40), rng.normal(10.0, 2.0, 40)], axis=1)
    w = rng.uniform(0.5, 1.5, 50)
    out["weights_prev"] = w / w.sum()
    out["tau_squared"] = 2.0 * np.cov(th_prev, rowvar=False)
    out["dist_a"] = rng.uniform(0.1, 2.0, 30)
    out["dist_b"] = rng.uniform(0.1, 2.0, 30)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    arrays = inputs()
    lines = []
    for name, a in arrays.items():
        a = np.asarray(a, dtype="<f8")
        with open(os.path.join(OUT, name + ".f64"), "wb") as f:
            f.write(np.asfortranarray(a).tobytes(order="F"))
        lines.append(name + " " + " ".join(str(d) for d in a.shape))
    with open(os.path.join(OUT, "manifest.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", len(lines), "arrays to", OUT)


def load_inputs():
    """The committed inputs, as the loader tests read them."""
    out = {}
    with open(os.path.join(OUT, "manifest.txt")) as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            shape = tuple(int(p) for p in parts[1:])
            raw = np.fromfile(os.path.join(OUT, parts[0] + ".f64"), dtype="<f8")
            out[parts[0]] = raw.reshape(shape, order="F")
    return out


if __name__ == "__main__":
    main()
