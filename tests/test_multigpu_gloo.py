"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharding + spot-record gather that bench.py does over
NCCL (slots are independent: contiguous shard per rank, no data-path collective, one all_gather of fixed-size
records).  The per-rank "decode" is stubbed by the CPU oracle on tiny 3200 sps slots -- this test covers the
plumbing, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle, result_dtype
    from tools import ft8enc, synth
    from tools.shard import shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    lo, hi = shard_range(n_total, rank, world)
    per = (n_total + world - 1) // world
    M = 50
    res = np.zeros((per, M), result_dtype)
    nres = np.full(per, -1, np.int32)
    for k, s in enumerate(range(lo, hi)):
        rng = np.random.Generator(np.random.PCG64(s))
        msg = ("CQ", synth.random_call(rng), synth.random_grid(rng))
        i_s, q_s = synth.slot_f32([(ft8enc.tones(ft8enc.pack_std(*msg)), 300.0 + 100.0 * s, 0.5, -5.0)], s)
        i_s, q_s, _ = orc.condition(i_s, q_s, 48000)
        o = orc.subsystem(i_s, q_s)
        res[k] = o["results"]
        nres[k] = o["n"]
    t_res = torch.from_numpy(res.view(np.uint8).reshape(per, M * 28))
    t_n = torch.from_numpy(nres)
    g_res = torch.empty((world * per, M * 28), dtype=torch.uint8)
    g_n = torch.empty(world * per, dtype=torch.int32)
    dist.all_gather_into_tensor(g_res, t_res)
    dist.all_gather_into_tensor(g_n, t_n)
    if rank == 0:
        np.savez(out_path, res=g_res.numpy(), n=g_n.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _step_worker(rank, world, port, chunks, bc, m, out_path):
    """bench.py's multi-GPU record path on gloo: every batch's records are staged as soon as the batch is collected, ONE all_gather
    per step moves the flat stage, rank 0 unpacks (rank, slot)-ordered records."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from tools.shard import stage_batch, stage_row_bytes, unpack_gathered

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stage = torch.zeros((chunks, stage_row_bytes(bc, m)), dtype=torch.uint8)
    gathered = torch.empty((world, chunks, stage_row_bytes(bc, m)), dtype=torch.uint8)
    steps = []
    for step in range(2):  # the stage is reused by the next step
        for k in range(chunks):
            slot0 = (rank * chunks + k) * bc   # global slot index of the batch's first slot
            res = torch.zeros((bc, m, 28), dtype=torch.uint8)
            nres = torch.zeros(bc, dtype=torch.int32)
            for j in range(bc):
                res[j, :, 0] = (slot0 + j + 7 * step) % 251
                res[j, j % m, 27] = 200 + step
                nres[j] = slot0 + j + 1000 * step
            stage_batch(stage, k, res, nres, bc, m)
        dist.all_gather_into_tensor(gathered.view(-1), stage.view(-1))
        if rank == 0:
            steps.append(unpack_gathered(gathered.numpy().copy(), world, chunks, bc, m))
    if rank == 0:
        np.savez(out_path, res0=steps[0][0], n0=steps[0][1], res1=steps[1][0], n1=steps[1][1])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_gather(tmp_path):
    import torch.multiprocessing as mp
    world, chunks, bc, m = 2, 3, 4, 5
    out = str(tmp_path / "steps.npz")
    mp.spawn(_step_worker, args=(world, _free_port(), chunks, bc, m, out), nprocs=world, join=True)
    g = np.load(out)
    total = world * chunks * bc
    for step in range(2):
        res, n = g[f"res{step}"], g[f"n{step}"]
        assert res.shape == (total, m, 28) and n.shape == (total,)
        assert n.tolist() == [s + 1000 * step for s in range(total)], "counts in (rank, slot) order"
        for s in range(total):
            assert (res[s, :, 0] == (s + 7 * step) % 251).all() and res[s, (s % bc) % m, 27] == 200 + step


def test_shard_ranges_cover_everything_once():
    sys.path.insert(0, ROOT)
    from tools.shard import shard_range
    for n in (0, 1, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))


def test_two_rank_gather(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import result_dtype
    n_total, world = 5, 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    g = np.load(out)
    per = (n_total + world - 1) // world
    res = g["res"].reshape(world * per, 50, 28).view(result_dtype).reshape(world * per, 50)
    n = g["n"]
    calls = []
    for r in range(world):
        lo = r * per
        for k in range(per):
            s = r * per + k
            if s < n_total:
                assert n[lo + k] >= 1, f"slot {s} did not decode"
                calls.append(res[lo + k][0]["call"].decode())
            else:
                assert n[lo + k] == -1  # padding record of the short last shard
    assert len(calls) == n_total and len(set(calls)) == n_total
