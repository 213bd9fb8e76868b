"""ORACLE (test infrastructure, not product code).

Host restatement of the counter-based mask-id stream used by the CUDA kernel
`codae_mask_table_philox` (csrc/corrupt.cu).  The reference draws its mask-id table
`Corrupter.mask_to_use[N, nb_run]` with un-seeded CPython `random.sample`
(/root/reference/codae/tool/data_tool.py:222-226); the B200 path replaces that
draw with Philox4x32-10 (Salmon et al., SC'11 -- published algorithm, restated
here) so that the id of observation `i` depends only on (seed, i) and never on
batch position, thread or rank.  The table produced here can be injected into
the reference (`corrupter.mask_to_use = torch.from_numpy(table)`), which then
emits exactly the dense masks the CUDA kernel applies.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
STREAM_TAG = np.uint32(0x434F4441)  # "CODA": keeps this stream apart from any other Philox user
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments are uint32 arrays (or scalars) of one shape."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    k0 = np.asarray(k0, dtype=np.uint32).copy()
    k1 = np.asarray(k1, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return c0, c1, c2, c3


def philox_mask_table(seed, nb_observation, nb_run, first_observation=0):
    """int16 [nb_observation, nb_run]: row i is a Fisher-Yates permutation of range(nb_run).

    draw t of observation i = lane (t % 4) of Philox(counter=(i_lo, i_hi, t // 4, TAG), key=seed);
    j = t + ((draw * (nb_run - t)) >> 32); swap(perm[t], perm[j]) for t = 0 .. nb_run-2.
    Same semantics as `Corrupter.mask_to_use` (a permutation per observation, data_tool.py:222-226).
    """
    obs = np.arange(first_observation, first_observation + nb_observation, dtype=np.uint64)
    lo = (obs & MASK32).astype(np.uint32)
    hi = (obs >> np.uint64(32)).astype(np.uint32)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    perm = np.tile(np.arange(nb_run, dtype=np.int16), (nb_observation, 1))
    rows = np.arange(nb_observation)
    block = None
    for t in range(nb_run - 1):
        if t % 4 == 0:
            block = philox4x32_10(lo, hi, np.full_like(lo, t // 4), np.full_like(lo, STREAM_TAG),
                                  np.full_like(lo, k0), np.full_like(lo, k1))
        draw = block[t % 4].astype(np.uint64)
        j = t + ((draw * np.uint64(nb_run - t)) >> np.uint64(32)).astype(np.int64)
        a = perm[rows, t].copy()
        perm[rows, t] = perm[rows, j]
        perm[rows, j] = a
    return perm
