"""Philox4x32 counter-based RNG (Salmon et al., SC'11; Random123) with ROUNDS = 7 rounds - the
Crush-resistant minimum of the paper (its table 2; Philox4x32-7 passes BigCrush, 10 rounds is the
conservative default) - and the normal / uniform transforms shared by the oracle and the CUDA kernels.
Round 1 of this project used 10 rounds; the stream is a specification of this project, not of the reference.

The reference draws its noise inside StochasticDiffEq's SOSRA solver
(/root/reference/src/utils/ou_process.jl:136-138), whose stream cannot be
reproduced; the B200 path therefore defines its own counter-based stream and this
file is its specification.  The CUDA implementation
(intrinsictimescales.jl_b200/csrc/philox.cuh) must produce the same uint32 words
bit for bit and the same normals to fp32 rounding.

Counter layout (128 bit) and key (64 bit):
    c0 = block index b            (sample j lives in block j >> 2, word j & 3)
    c1 = trial index
    c2 = global particle index (low 32 bit)
    c3 = (stream << 28) | (generation & 0x0FFFFFFF)
    key = (seed & 0xffffffff, seed >> 32)
Streams: 0 = OU driving noise xi_j, 1 = per-trial initial state u0 (word 0/1) and
oscillation phase (word 2).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_NOISE = 0
STREAM_INIT = 1


ROUNDS = 7


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=ROUNDS):
    """Vectorised Philox4x32-`rounds`.  All inputs broadcastable integer arrays; returns
    four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK32
                                           for c in (c0, c1, c2, c3)])
    c0 = c0.copy(); c1 = c1.copy(); c2 = c2.copy(); c3 = c3.copy()
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0 = p0 >> np.uint64(32)
        lo0 = p0 & MASK32
        hi1 = p1 >> np.uint64(32)
        lo1 = p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """The 10-round variant (Random123's default; only the known-answer tests use it)."""
    return philox4x32(c0, c1, c2, c3, k0, k1, rounds=10)


def u01(r):
    """uint32 -> open-interval uniform ((r >> 9) + 0.5) * 2^-23: 24 significant bits,
    exactly representable in fp32, never 0 or 1."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


def box_muller(ra, rb):
    """Two uint32 words -> two N(0,1) draws: R = sqrt(-2 ln u), z0 = R cos(2 pi v),
    z1 = R sin(2 pi v)."""
    u = u01(ra)
    v = u01(rb)
    rad = np.sqrt(-2.0 * np.log(u))
    ang = 2.0 * np.pi * v
    return rad * np.cos(ang), rad * np.sin(ang)


def c3_word(stream, generation):
    return ((int(stream) & 0xF) << 28) | (int(generation) & 0x0FFFFFFF)


def noise_normals(seed, generation, particle, trial, n):
    """xi_0 .. xi_{n-1} for one (particle, trial): block b -> xi_{4b..4b+3}."""
    nb = (n + 3) // 4
    b = np.arange(nb, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32(b, trial, int(particle) & 0xFFFFFFFF,
                                   c3_word(STREAM_NOISE, generation),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = box_muller(r0, r1)
    z2, z3 = box_muller(r2, r3)
    z = np.stack([z0, z1, z2, z3], axis=1).reshape(-1)
    return z[:n]


def init_state_and_phase(seed, generation, particle, trial):
    """(u0 ~ N(0,1), phase ~ U(0, 2 pi)) for one (particle, trial)."""
    r0, r1, r2, _ = philox4x32(0, trial, int(particle) & 0xFFFFFFFF,
                                  c3_word(STREAM_INIT, generation),
                                  seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, _ = box_muller(r0, r1)
    phase = 2.0 * np.pi * u01(r2)
    return float(z0), float(phase)
