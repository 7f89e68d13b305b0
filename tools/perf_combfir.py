#!/usr/bin/env python3
"""A/B of cic_comb_fir_kernel builds (run under gpurun).  Each library under rtlsdr-ft8d_b200/build/ab/lib_<threads>_<tiles>.so (built with
-DFT8B200_COMB_THREADS / -DFT8B200_COMB_TILES) runs in its own process: comb+FIR launch time at 128 raw slots on the whole GPU (CUDA events
of the stage-wise API, minimum of 8) and inside the SM-partitioned executor (32 back-end SMs, mean over 36 batches), with a digest of the
3200 sps outputs so that every build is seen to produce the same samples.  usage: tools/perf_combfir.py [TAG]; child: --one
Building a variant (in rtlsdr-ft8d_b200/, after `make`):
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -DFT8B200_COMB_THREADS=256 \
       -DFT8B200_COMB_TILES=1 -c csrc/decimator.cu -o build/ab/decimator_256_1.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/lib_256_1.so build/ab/decimator_256_1.o \
       $(ls build/*.o | grep -v decimator.o) -lcudart -ldl
and `cp libft8b200.so build/ab/lib_base.so` for the build under test (FT8B200_LIB_PATH selects the library the harness loads)."""
import glob, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import time
    import numpy as np, torch, bench
    from ft8b200_loader import load
    pkg = load()
    dev = torch.device("cuda:0")
    B = 128
    iq, _ = bench.gen_batch(B, 0, dev)
    ctx = pkg.Context(0)
    ctx.set_profiling(True)
    t = []
    for _ in range(8):
        ctx.process_raw(iq, B)
        torch.cuda.synchronize()
        t.append(ctx.stage_times()["comb_fir"])
    res, n = ctx.fetch_results(B)
    dec = ctx.decimate(iq, B, pkg.RAW_SLOT_BYTES, pkg.RAW_SLOT_BYTES)
    h = hashlib.sha256()
    for x in dec[:4]:
        h.update(np.ascontiguousarray(x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)).tobytes())
    h.update(res.tobytes())
    out = {"whole_gpu_ms_min": min(t), "whole_gpu_ms_all": t, "digest": h.hexdigest()[:16]}
    ctx.close()
    pipe = pkg.Pipe(0, 3)
    pipe.set_mode(serial=False, decimator_variant=0)
    pipe.set_partition(32)
    wall = 0.0
    for rep in range(2):
        pipe.set_profiling(rep == 1)
        t0 = time.perf_counter()
        for i in range(36):
            while pipe.in_flight() >= 3:
                pipe.collect(B)
            pipe.submit(iq, B)
        while pipe.in_flight():
            pipe.collect(B)
        wall = time.perf_counter() - t0
    ms, nb = pipe.stage_times()
    out["partition32_stage_ms"] = {k: v / nb for k, v in ms.items()}
    out["partition32_ms_per_batch_wall"] = wall * 1e3 / 36
    print(json.dumps(out))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "x"
    libs = sorted(glob.glob(os.path.join(ROOT, "rtlsdr-ft8d_b200", "build", "ab", "lib_*.so")))
    allr = {}
    for lib in libs:
        env = dict(os.environ, FT8B200_LIB_PATH=lib)
        r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
        name = os.path.basename(lib)[4:-3]
        try:
            allr[name] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            allr[name] = {"error": (r.stderr or r.stdout)[-400:]}
        v = allr[name]
        print(name, v.get("whole_gpu_ms_min"), (v.get("partition32_stage_ms") or {}).get("comb_fir"), v.get("partition32_ms_per_batch_wall"), v.get("digest"), v.get("error", ""))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(allr, open(os.path.join(ROOT, "gpurun_out", "perf_combfir_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    one() if "--one" in sys.argv else main()
