"""Philox4x32-10 on the host (pure Python ints) -- only for single uniforms of the test hooks
(`Vertex.sampleNextVertex()` without an explicit x).  Same stream definition as the device code:
counter = (walk_id lo, walk_id hi, draw // 2, 0), key = seed, two 53-bit uniforms per block."""

_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for r in range(10):
        if r:
            k0 = (k0 + _W0) & _MASK
            k1 = (k1 + _W1) & _MASK
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
    return c0, c1, c2, c3


def uniform(seed, walk_id, draw):
    r = philox4x32_10((walk_id & _MASK, (walk_id >> 32) & _MASK, draw >> 1, 0), (seed & _MASK, (seed >> 32) & _MASK))
    lo, hi = (r[2], r[3]) if draw & 1 else (r[0], r[1])
    return float(((hi << 32) | lo) >> 11) * 2.0 ** -53
