#!/usr/bin/env python3
"""bench.py -- FT8 15 s-slots decoded per second, from raw 2.4 Msps uint8 RTL IQ (BASELINE.json config #2:
one 15 s slot = 36 M complex samples through the full CIC+FIR decimation and FT8 decode), batched.

  python bench.py --gpus N --steps K --warmup W                     this repository's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W    the reference's own CPU implementation, all host cores

A "step" is one pass of the whole hot path (decimate -> condition -> waterfall -> Costas sync/top-K ->
LLR/LDPC/CRC/unpack -> spot table) over one batch of synthetic slots.  `value` times it with the batch already
resident in HBM (the batch is larger than L2); `e2e` times the same path through the C-ABI call that takes HOST
buffers (pinned), host->device copies and the device->host read of the spot records inside the timed region.
`roofline` is for the dominant kernel (cic_block_sums, HBM-bound), timed live with CUDA events on the launching
stream; `cpu_baseline` is the CPU checker (oracle/_ref = the unmodified reference when it was built, else the
restatement) timed on this box's host cores on a bounded sample of the same batch.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAW_SLOT_BYTES = 72_000_000
ALGO_BYTES_PER_SLOT = 72_000_000 + 47_936 * 8  # 2 B/sample in + 8 B per output
METRIC = "FT8 15s-slots decoded/sec from raw 2.4 Msps uint8 IQ"


# --------------------------------------------------------------------------------------------- inputs
def slot_params(seed: int):
    """Deterministic message / frequency / time offset of synthetic slot `seed`."""
    from tools import ft8enc, synth
    rng = np.random.Generator(np.random.PCG64(0xF78 + seed))
    to, de, ex = synth.random_message(rng)
    if seed % 2 == 0:
        to = "CQ"
        ex = synth.random_grid(rng)
    return dict(text=f"{to} {de} {ex}", f_hz=float(rng.uniform(200.0, 1400.0)), t0=float(0.5 + rng.uniform(-0.3, 0.3)),
                amp=20.0, noise=30.0)


def gen_batch(n_slots: int, first_seed: int, device, ctx=None):
    """n_slots x 72 MB of uint8 IQ made ON THE DEVICE by the library's synthesiser (csrc/synth.cu): one FT8 message per
    slot as phase-continuous 8-FSK at (f - 600 kHz), 20 LSB over 30 LSB of noise, offset 128, saturating."""
    from ft8b200_loader import load
    pkg = load()
    own = ctx is None
    if own:
        ctx = pkg.Context(device.index or 0)
    params = [slot_params(first_seed + s) for s in range(n_slots)]
    sig = pkg.make_signals((pkg.pack77_std(*p["text"].split()), p["f_hz"], p["t0"], p["amp"]) for p in params)
    buf = ctx.synth_raw(sig, np.arange(n_slots + 1, dtype=np.int32), params[0]["noise"], 0xF78, first_slot_index=first_seed)
    if own:
        ctx.close()
    return buf, [p["text"] for p in params]


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region: NVML polled every 2 ms from a thread
    (nvidia-smi -lms cannot resolve a region that lasts milliseconds)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def loop():
            while not self.stop_flag:
                try:
                    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        why = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        why = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.rows.append((time.time(), float(mhz), int(why)))
                except Exception:
                    pass
                time.sleep(0.002)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1)

    def summary(self, t0: float, t1: float):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if not rows:  # region shorter than one poll: take the nearest samples
            rows = sorted(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:3]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        why = 0
        for r in rows:
            why |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, bit in self.REASONS.items() if why & bit), "samples": len(rows)}


# --------------------------------------------------------------------------------------------- CPU arms
_CPU_SLOTS = None  # numpy uint8 [n, 72e6], shared with forked workers


def cpu_kind():
    from oracle.pyoracle import Reference
    return "reference" if Reference.available("k120") else "port"


_CPU_PHASE_S = {"decimator": 0.0, "decode": 0.0}   # seconds spent per phase by _cpu_one_slot in THIS process (single-thread baseline)


def _cpu_one_slot(idx: int):
    """Raw slot -> spots on the CPU: rtlsdr_callback() in 65536-byte calls, decoder() conditioning, ft8_subsystem()."""
    from oracle.pyoracle import Oracle, Reference
    raw = _CPU_SLOTS[idx]
    orc = Oracle()
    t0 = time.perf_counter()
    if cpu_kind() == "reference":
        ref = Reference("k120", fresh=True)  # private copy: the daemon's decimator state is function-static
        for o in range(0, raw.size, 65536):
            ref.callback(raw[o:o + 65536])
        i_s, q_s, n = ref.rx()
        t1 = time.perf_counter()
        i_s, q_s, _ = orc.condition(i_s, q_s, n)  # decoder() itself is thread-bound in the daemon (rtlsdr_ft8d.c:221-285)
        count = int(ref.subsystem(i_s, q_s)["n"])
    else:
        oi, oq = orc.decimate_slot(raw)
        t1 = time.perf_counter()
        i_s = np.zeros(48000, np.float32); q_s = np.zeros(48000, np.float32)
        i_s[:oi.size] = oi; q_s[:oq.size] = oq
        i_s, q_s, _ = orc.condition(i_s, q_s, oi.size)
        count = int(orc.subsystem(i_s, q_s)["n"])
    _CPU_PHASE_S["decimator"] += t1 - t0
    _CPU_PHASE_S["decode"] += time.perf_counter() - t1
    return count


def _cpu_synth_slot(idx: int):
    """Slot `idx` of the bench batch into _CPU_SLOTS (shared memory), by the CPU twin of ft8b200_synth_raw."""
    from oracle.pyoracle import Oracle, signal_dtype
    orc = Oracle()
    p = slot_params(idx)
    sig = np.zeros(1, signal_dtype)
    sig[0]["payload"] = np.frombuffer(orc.pack_std(*p["text"].split()), np.uint8)
    sig[0]["f0_hz"], sig[0]["t0_sec"], sig[0]["amp"] = p["f_hz"], p["t0"], p["amp"]
    _CPU_SLOTS[idx] = orc.synth_raw(sig, p["noise"], 0xF78, idx, RAW_SLOT_BYTES // 2)
    return 0


def cpu_run(n_slots: int, workers: int):
    """Seconds to push slots 0..n_slots-1 of _CPU_SLOTS through the CPU path with `workers` processes."""
    t0 = time.perf_counter()
    if workers <= 1:
        out = [_cpu_one_slot(k) for k in range(n_slots)]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            out = pool.map(_cpu_one_slot, range(n_slots), chunksize=1)
    return time.perf_counter() - t0, out


# --------------------------------------------------------------------------------------------- main
_JSON_FD = None


def bind_to_gpu_numa(local: int):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e measurement
    (first touch) are local to the PCIe root the copies go through.  Returns a note for `config`, or None when it does not apply."""
    if os.environ.get("BENCH_NUMA", "1") == "0":
        return "off"
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "GPU %s on NUMA node %d: rank bound to its %d CPUs" % (bus, node, len(cpus))
    except Exception as exc:
        return "not bound (%s)" % type(exc).__name__


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slots", type=int, default=512, help="15 s slots per GPU per step (36.9 GB of raw IQ per GPU at 512)")
    ap.add_argument("--e2e-slots", type=int, default=8, help="slots per step of the host-buffer (e2e) measurement")
    ap.add_argument("--chunks", type=int, default=4, help="executor batches per step: a step's slots are submitted to ft8b200_pipe_t in this many batches")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight in the pipelined executor")
    ap.add_argument("--back-sms", type=int, default=32, help="SMs of the back-end partition (waterfall/sync/LDPC of batch n next to the decimator "
                                                             "of batch n+1 on the other SMs); 0 = no partition, kernels of consecutive batches run serially")
    ap.add_argument("--overlap", action="store_true", help="(without a partition) let the back end of batch n time-share the GPU with the decimator of batch n+1")
    ap.add_argument("--k1-variant", type=int, default=0, help="0 = streaming cic_block_sums kernel, 1..6 = bulk-copy (TMA) variants")
    ap.add_argument("--cpu-slots", type=int, default=96, help="bounded CPU-baseline sample (slots)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line, the JSON: anything a library prints to fd 1 on the way (NCCL's version banner at
    # communicator creation, for one) is sent to stderr, and the JSON line is written to the original stdout at the end
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "BASELINE config #2 batched: %d x (one 15 s slot of raw 2.4 Msps uint8 RTL IQ, 36 M complex samples, one FT8 message "
                          "at 20 LSB over 30 LSB noise) per GPU per step -> CIC+FIR decimation -> decoder() conditioning -> waterfall -> "
                          "sync (K=120) -> LDPC/CRC/unpack -> spot table" % args.slots,
              "slots_per_gpu_per_step": args.slots, "input_bytes_per_step_per_gpu": args.slots * RAW_SLOT_BYTES,
              "l2": "inputs larger than L2 (%.1f GB per step per GPU vs 126 MB)" % (args.slots * RAW_SLOT_BYTES / 1e9),
              "max_candidates": 120, "max_messages": 50, "ldpc_iterations": 20,
              "executor": "ft8b200_pipe_t depth %d, %d batches of %d slots per step, %s" % (
                  args.depth, args.chunks, args.slots // max(args.chunks, 1),
                  ("SM partition: back end of batch n on >= %d SMs, decimator of batch n+1 on the others" % args.back_sms) if args.back_sms > 0 else
                  ("overlap (time-shared)" if args.overlap else "serial (kernels of consecutive batches do not share the GPU)")),
              "parallelism": "slots sharded across GPUs, no data-path collective; "
              "spot records gathered with one NCCL all_gather per step" if world > 1 else "single GPU"}

    if args.impl == "reference":
        return reference_arm(args, rank, world, config)

    import torch
    import torch.distributed as dist
    from ft8b200_loader import load
    pkg = load()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if numa:
        config["numa"] = numa
    if world > 1:
        # the all_gather of the spot records must not queue behind the decimator's 190k-CTA grid: high-priority NCCL stream
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=device, pg_options=opts)

    B = args.slots
    if args.chunks < 1 or B % args.chunks:
        raise SystemExit("--slots must be a multiple of --chunks")
    Bc = B // args.chunks   # slots per executor batch
    batch, texts = gen_batch(B, 100_000 * rank, device)
    torch.cuda.synchronize()
    # The product's batch executor (ft8b200_pipe_t): `depth` batches in flight; in SERIAL mode the kernels of consecutive
    # batches never share the GPU (each kernel is timed alone), only D2H of the records and host work overlap them.
    pipe = pkg.Pipe(local, args.depth)
    pipe.set_mode(serial=not args.overlap, decimator_variant=args.k1_variant)
    if args.back_sms > 0:
        # green contexts: disjoint SM sets for the HBM-bound decimator and the issue-bound back end (ft8b200_pipe_set_partition)
        try:
            config["sm_partition"] = dict(zip(("front_sms", "back_sms"), pipe.set_partition(args.back_sms)))
        except Exception as exc:  # a driver without green contexts: same kernels, consecutive batches back to back on the whole GPU
            config["sm_partition"] = "unavailable (%s): running serial" % exc
            config["executor"] = "ft8b200_pipe_t depth %d, %d batches of %d slots per step, serial" % (args.depth, args.chunks, args.slots // args.chunks)
            args.back_sms = 0
    M = pipe.M
    # Multi-GPU: the spot records of a step's batches are staged on the device (a local copy, so a lane is free again as soon
    # as its records are copied) and gathered to every rank with ONE NCCL all_gather per step; rank 0 reads them on the host.
    from tools.shard import stage_batch, stage_row_bytes, unpack_gathered
    if world > 1:
        stage = torch.empty((args.chunks, stage_row_bytes(Bc, M)), dtype=torch.uint8, device=device)  # per batch: records, then counts
        gathered = torch.empty((world, args.chunks, stage_row_bytes(Bc, M)), dtype=torch.uint8, device=device)
        copied = torch.cuda.Event()
        # rank 0 reads every step's gathered records into pinned host memory without blocking its submit loop
        host_rec = [torch.empty(gathered.shape, dtype=torch.uint8).pin_memory() for _ in range(2)] if rank == 0 else None
        host_ev = [torch.cuda.Event(), torch.cuda.Event()]
    n_staged = [0]

    def last_gathered():
        """rank 0: records of the last gathered step, all ranks, in (rank, slot) order."""
        if rank != 0 or n_staged[0] < args.chunks:
            return None
        i = (n_staged[0] // args.chunks - 1) % 2
        host_ev[i].synchronize()
        return unpack_gathered(host_rec[i].numpy(), world, args.chunks, Bc, M)

    def collect():
        """Oldest batch -> host records on rank 0 (multi-GPU: one NCCL all_gather of the fixed-size spot records per step)."""
        if world > 1 and os.environ.get("BENCH_NOGATHER"):
            return pipe.collect(Bc)   # diagnostic only: how fast would the ranks run without the collective
        if world > 1:
            res_dev, nres_dev = pipe.collect_device()
            stage_batch(stage, n_staged[0] % args.chunks, res_dev, nres_dev, Bc, M)
            copied.record()
            pipe.depend_on(copied)   # the lane's buffers are rewritten only after they have been copied out (ordered on the device)
            n_staged[0] += 1
            if n_staged[0] % args.chunks == 0:
                dist.all_gather_into_tensor(gathered.view(-1), stage.view(-1))   # spot records over NVLink, once per step
                if rank == 0:
                    i = (n_staged[0] // args.chunks - 1) % 2
                    host_ev[i].synchronize()   # the copy issued two steps ago (long finished) owns this buffer
                    host_rec[i].copy_(gathered, non_blocking=True)
                    host_ev[i].record()
            return None
        return pipe.collect(Bc)

    def run(steps):
        """`steps` passes over the B resident slots, each submitted as args.chunks executor batches; every batch's spot
        records are read back.  Returns the records of the last pass in slot order."""
        outs = []
        for _ in range(steps):
            for c in range(args.chunks):
                if pipe.in_flight() == pipe.depth:
                    outs.append(collect())
                pipe.submit(batch[c * Bc:(c + 1) * Bc], Bc)
        while pipe.in_flight():
            outs.append(collect())
        if world > 1:
            return last_gathered()
        last = outs[-args.chunks:]
        return np.concatenate([np.asarray(o[0]) for o in last]), np.concatenate([np.asarray(o[1]) for o in last])

    out = run(args.warmup)
    # correctness guard on the first batch: every synthetic slot must decode to its own message
    if world == 1:
        res, nres = out
        def slot_ok(s):  # CQ messages must come back with their call sign; other messages only count as a decode (a15)
            if nres[s] < 1:
                return False
            to, de = texts[s].split()[:2]
            return to != "CQ" or any(r["call"] == de.encode() for r in res[s][:nres[s]])
        n_good = sum(1 for s in range(B) if slot_ok(s))
    else:
        n_good = -1
        if rank == 0 and out is not None:  # gathered records of every rank: slots that produced at least one message
            config["decoded_slots_all_ranks"] = "%d of %d" % (int((out[1] >= 1).sum()), world * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    pipe.set_profiling(True)
    launches0 = pipe.launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    run(args.steps)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = pipe.launches() - launches0
    stage_acc, n_prof = pipe.stage_times()
    pipe.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- e2e: host buffers in, host results out, through the C-ABI calls that take HOST memory
    Be = min(args.e2e_slots, B)
    host = torch.empty((Be, RAW_SLOT_BYTES), dtype=torch.uint8, pin_memory=True)
    host.copy_(batch[:Be])
    host_np = host.numpy()

    def run_host(steps):
        out = None
        for _ in range(steps):
            if pipe.in_flight() == pipe.depth:
                out = pipe.collect(Be)
            pipe.submit_host(host_np, Be)
        while pipe.in_flight():
            out = pipe.collect(Be)
        return out

    run_host(args.warmup)
    barrier()
    e0.record()
    r_e2e, n_e2e = run_host(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    sampler.stop()
    e2e = {"value": world * Be * args.steps / (ms_e2e * 1e-3), "unit": "slots/s", "h2d_bytes_per_step": Be * RAW_SLOT_BYTES,
           "d2h_bytes_per_step": Be * (M * 28 + 4), "slots_per_step": Be, "same_results_as_device_path": bool(world > 1 or (
               np.array_equal(n_e2e, nres[:Be]) and r_e2e.tobytes() == res[:Be].tobytes())),
           "h2d_gbs": world * Be * RAW_SLOT_BYTES * args.steps / (ms_e2e * 1e-3) / 1e9,
           "api": "ft8b200_pipe_submit_host / ft8b200_pipe_collect (pinned host IQ in, decoder_results out, %d batches in flight)" % pipe.depth}

    # ---- roofline of the dominant kernel (cic_block_sums): algorithmic bytes / CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_prof = max(n_prof, 1)
    k1_ms = stage_acc.get("block_sums", 0.0) / n_prof
    k2_ms = stage_acc.get("comb_fir", 0.0) / n_prof
    achieved = Bc * ALGO_BYTES_PER_SLOT / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None
    roofline = {"kernel": "cic_block_sums_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured, burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": Bc * ALGO_BYTES_PER_SLOT, "slots_per_launch": Bc, "launch_ms": k1_ms,
                "decimator_ms_incl_comb_fir": k1_ms + k2_ms,
                "decimator_msps": Bc * 36.0 / ((k1_ms + k2_ms) * 1e-3) if k1_ms > 0 else None,
                "timed": "CUDA events around every launch of the kernel inside the timed region (%d launches)" % n_prof,
                "stage_ms_per_launch": {k: v / n_prof for k, v in stage_acc.items()},
                "stage_note": "with an SM partition the back-end stages (comb_fir ... spots) run on the back-end SMs concurrently with the "
                              "next batch's block sums: their times overlap it and do not add to the step" if args.back_sms > 0 else "stages run back to back"}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))  # one `ncu --set full` capture, per slot
        roofline["traffic"] = tr.get("dram_bytes_per_slot", 0) * Bc or None
        roofline["traffic_source"] = tr.get("source")
    except Exception:
        pass

    out = {"metric": METRIC, "value": value, "unit": "slots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32",
           "data": "synthetic (generated on the device by ft8b200_synth_raw)", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
           "decoded_ok_slots_in_first_batch": n_good}

    if rank == 0 and world == 1:
        global _CPU_SLOTS
        n_cpu = min(args.cpu_slots, B)
        _CPU_SLOTS = batch[:n_cpu].cpu().numpy()
        _CPU_PHASE_S["decimator"] = _CPU_PHASE_S["decode"] = 0.0
        secs, n_dec = cpu_run(n_cpu, 1)
        gpu_n = [int(x) for x in nres[:n_cpu]]
        out["cpu_baseline"] = {"value": n_cpu / secs, "unit": "slots/s", "cores": 1, "kind": cpu_kind(),
                               "sample": "%d of the step's %d slots, single thread: rtlsdr_callback in 65536-byte calls + decoder() "
                                         "conditioning + ft8_subsystem (%.2f s)" % (n_cpu, B, secs),
                               "same_spot_counts_as_gpu": n_dec == gpu_n,
                               # the two halves of the path on one host thread
                               "decimator_msps": n_cpu * (RAW_SLOT_BYTES // 2) / max(_CPU_PHASE_S["decimator"], 1e-9) / 1e6,
                               "decode_slots_per_s": n_cpu / max(_CPU_PHASE_S["decode"], 1e-9)}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def reference_arm(args, rank, world, config):
    """The reference's own CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return 0
    global _CPU_SLOTS
    cores = os.cpu_count() or 1
    n = max(4 * cores, 16)  # a few slots per worker per step
    # Inputs: the same synthetic slots as the CUDA arm, made by the synthesiser's bit-identical CPU twin (oracle/ft8_oracle_synth.c)
    # in the worker processes -- nothing of libft8b200 and no GPU is involved anywhere in this arm.
    try:
        import multiprocessing as mp
        shared = mp.RawArray("B", n * RAW_SLOT_BYTES)
        _CPU_SLOTS = np.frombuffer(shared, np.uint8).reshape(n, RAW_SLOT_BYTES)
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_synth_slot, range(n), chunksize=1)
    except Exception as exc:  # pragma: no cover
        emit({"impl": "reference", "unavailable": f"input synthesis failed: {exc}"})
        return 0
    for _ in range(min(args.warmup, 1)):
        cpu_run(min(n, cores), cores)
    t = 0.0
    for _ in range(args.steps):
        secs, _ = cpu_run(n, cores)
        t += secs
    value = n * args.steps / t
    kind = cpu_kind()
    cfg = dict(config)
    cfg["reference_sample"] = "%d slots per step over %d worker processes" % (n, cores)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "slots/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f32",
           "data": "synthetic (same slots as the CUDA arm, made on the host by the synthesiser's bit-identical CPU twin)", "config": cfg,
           "cpu_baseline": {"value": value, "unit": "slots/s", "cores": cores, "kind": kind,
                            "sample": "%d slots per step, one forked process per slot on %d cores (the reference itself is single-threaded)" % (n, cores)},
           "e2e": {"value": value, "unit": "slots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
