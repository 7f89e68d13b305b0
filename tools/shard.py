"""Contiguous block sharding of independent slots / receiver streams across ranks."""


def shard_range(n_items: int, rank: int, world: int):
    """Items [lo, hi) owned by `rank`: blocks of ceil(n/world), the last ranks may be short or empty."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


# ---- spot records of one step, staged for ONE all_gather (bench.py, multi-GPU) --------------------------------------------------
# A step is `chunks` executor batches of `bc` slots.  Each batch contributes bc * m decoder_results records (28 bytes each) and bc
# int32 counts; the stage holds them batch-major as uint8[chunks][bc*m*28 + 4*bc], so that one all_gather of the flat buffer moves
# everything and a batch's part can be written as soon as that batch has been collected.
REC_BYTES = 28


def stage_row_bytes(bc: int, m: int) -> int:
    return bc * m * REC_BYTES + 4 * bc


def stage_batch(stage, k: int, res, nres, bc: int, m: int):
    """Copy batch k's records (uint8[bc, m, 28] tensor) and counts (int32[bc] tensor) into the stage tensor uint8[chunks, row]."""
    import torch
    rec = bc * m * REC_BYTES
    stage[k, :rec].view(bc, m, REC_BYTES).copy_(res)
    stage[k, rec:].view(torch.int32).copy_(nres)


def unpack_gathered(g, world: int, chunks: int, bc: int, m: int):
    """numpy uint8[world, chunks, row] (the gathered stages) -> (records uint8[world*chunks*bc, m, 28], counts int32[world*chunks*bc])
    in (rank, slot) order."""
    import numpy as np
    rec = bc * m * REC_BYTES
    g = np.asarray(g).reshape(world, chunks, stage_row_bytes(bc, m))
    res = np.ascontiguousarray(g[:, :, :rec]).reshape(world * chunks * bc, m, REC_BYTES)
    nres = np.ascontiguousarray(g[:, :, rec:]).view(np.int32).reshape(world * chunks * bc)
    return res, nres
