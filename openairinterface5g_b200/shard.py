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
    """GPU that owns the HARQ soft buffer of (ulsch_id, segment) for the lifetime of the HARQ process."""
    return (ulsch_id * 131 + segment) % world
