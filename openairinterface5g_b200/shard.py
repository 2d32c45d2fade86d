"""Code-block sharding across GPUs (one process per GPU).  Code blocks are independent units until the TB-level CRC
(reference phy_procedures_nr_gNB.c:271-300), so the data path needs no collective: every rank decodes its own slice.
In offload mode the HARQ soft buffer of (ulsch_id, segment r) must stay on one GPU across retransmissions, hence the
sticky hash."""


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` of `world` (first n_items % world ranks get one extra)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sticky_gpu(ulsch_id, segment, world):
    """GPU that owns the HARQ soft buffer of (ulsch_id, segment) for the lifetime of the HARQ process: the library's rule
    (nrb200_sticky_device in csrc/nrb200_abi.cu -- the C entry points of the offload convention use the same function), restated here for
    the one-process-per-GPU tools so that both agree; tests/test_abi_symbols.py checks the two against each other."""
    h = (ulsch_id * 0x9E3779B1 + segment * 0x85EBCA77) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & 0xFFFFFFFF
    h ^= h >> 12
    return h % world if world else 0
