"""Contiguous block sharding of independent slots / receiver streams across ranks."""


def shard_range(n_items: int, rank: int, world: int):
    """Items [lo, hi) owned by `rank`: blocks of ceil(n/world), the last ranks may be short or empty."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
