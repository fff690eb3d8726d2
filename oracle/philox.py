"""Counter-based RNG shared by the oracle and the CUDA engine.  TEST INFRASTRUCTURE.

Philox4x32-10 (Salmon et al., SC'11).  The draw-addressing scheme below is THIS project's
(the reference uses two global sequential generators); it is what makes a GPU
wavefront and a sequential CPU walk consume identical uniforms ("replay"):

    particle key  (2 x u32)  = path-derived: root = philox(ctr=(shower_lo, shower_hi, 0, ST_KEY_ROOT), key=seed)[0:2]
                                              child = philox(ctr=(child_bit, ST_KEY_CHILD, 0, 0), key=parent_key)[0:2]
    one call      philox(ctr=(c0, stream, c2, c3), key=particle key) -> 4 x u32 -> two doubles + 24 spare bits
                  d0 = u52(out0, out1), d1 = u52(out2, out3),  u52(hi, lo) = (hi * 2^20 + (lo >> 12)) * 2^-52
                  spare = (out1 & 0xFFF) << 12 | (out3 & 0xFFF)   (the bits u52 leaves over)
                  u48(s0, s1) = (s0 * 2^24 + s1) * 2^-48          (uniform from the spare bits of two calls)

Streams and counter layout are listed in ``STREAMS`` and mirrored in petite_b200/csrc/rng.cuh.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

ST_SUBSTEP = 1    # ctr (i, ST, 0, 0)        -> (u_hard, u_dz)           shower.py:561-562
ST_FINAL = 2      # ctr (0, ST, 0, 0)        -> (distC, -)               shower.py:540,583
ST_MCS = 3        # ctr (i, ST, 0, pc)       -> (u_phi, u_radius); sign = spare bit 0 (the Box-Muller angle drops out of |z|)
ST_CHOICE = 4     # ctr (0, ST, 0, 0)        -> (u_choice, -)            shower.py:671-697
ST_VEGAS = 5      # ctr (t, ST, j, pc)       doubles D[2j], D[2j+1] of trial t; D = y_0..y_{dim-1}, u_accept (dim 4: u_accept = u48 of the two calls' spare bits)
ST_KIN = 6        # ctr (0, ST, 0, pc)       -> (u_az1, u_az2)           kinematics.py
ST_DECAY = 7      # ctr (0, ST, 0, pc)       -> (u_cos, u_phi)           particle.py:234-235
ST_DBIN = 8       # ctr (0, ST, 0, pc)       -> (u_bin, -)               dark_shower.py:752
ST_PE = 11        # ctr (i, ST, 0, pc)       -> (u_x, u_u)               dark_shower.py:715-716
ST_C0 = 12        # ctr (0, ST, 0, pc)       -> (u_c0, -)                dark_shower.py:776
ST_KEY_ROOT = 0xA0
ST_KEY_CHILD = 0xA1
MCS_FINAL_INDEX = 0xFFFFFFFF

STREAMS = {k: v for k, v in globals().items() if k.startswith("ST_")}


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable integer arrays; returns 4 uint64 arrays (< 2^32)."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(*[np.asarray(a, dtype=np.uint64) for a in (c0, c1, c2, c3, k0, k1)])
    c0, c1, c2, c3, k0, k1 = [a.copy() & MASK for a in (c0, c1, c2, c3, k0, k1)]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & MASK
        n1 = p1 & MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & MASK
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def u52(hi, lo):
    hi = np.asarray(hi, dtype=np.uint64)
    lo = np.asarray(lo, dtype=np.uint64)
    return ((hi << np.uint64(20)) | (lo >> np.uint64(12))).astype(np.float64) * (1.0 / 4503599627370496.0)


def u48(s0, s1):
    s0 = np.asarray(s0, dtype=np.uint64)
    s1 = np.asarray(s1, dtype=np.uint64)
    return ((s0 << np.uint64(24)) | s1).astype(np.float64) * (1.0 / 281474976710656.0)


def draw2s(key, c0, stream, c2=0, c3=0):
    """Two doubles in [0,1) and the call's 24 spare bits for (key, counter).  Vectorised over c0/c2."""
    o0, o1, o2, o3 = philox4x32(c0, stream, c2, c3, key[0], key[1])
    spare = ((o1 & np.uint64(0xFFF)) << np.uint64(12)) | (o3 & np.uint64(0xFFF))
    return u52(o0, o1), u52(o2, o3), spare


def draw2(key, c0, stream, c2=0, c3=0):
    """Two doubles in [0,1) for (key, counter).  Vectorised over c0/c2."""
    a, b, _ = draw2s(key, c0, stream, c2, c3)
    return a, b


def root_key(seed, shower_id):
    seed = int(seed)
    shower_id = int(shower_id)
    o = philox4x32(shower_id & 0xFFFFFFFF, (shower_id >> 32) & 0xFFFFFFFF, 0, ST_KEY_ROOT,
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return (int(o[0]), int(o[1]))


def child_key(key, bit):
    o = philox4x32(int(bit), ST_KEY_CHILD, 0, 0, key[0], key[1])
    return (int(o[0]), int(o[1]))
