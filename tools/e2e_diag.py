#!/usr/bin/env python
"""Diagnostic for the end-to-end leg's run-to-run spread: for several freshly allocated host arenas report how much of
each is backed by 2 MB pages (/proc/self/smaps), the plain cudaMemcpy bandwidth to and from it, and the Mpps of
bng_prog_run(BNG_MEM_HOST) over 2^20 64-byte frames in it.
    gpurun -- 'python tools/e2e_diag.py'"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from bng_b200 import MEM_HOST, Dataplane, workloads as W  # noqa: E402
from bng_b200.layouts import as_bytes  # noqa: E402


def huge_kb(addr, nbytes):
    """AnonHugePages (kB) of the mappings that overlap [addr, addr + nbytes)."""
    tot = 0
    cur = None
    for line in open("/proc/self/smaps"):
        p = line.split()
        if "-" in p[0] and len(p) >= 5 and ":" not in p[0]:
            lo, hi = (int(x, 16) for x in p[0].split("-"))
            cur = lo < addr + nbytes and hi > addr
        elif cur and p[0] == "AnonHugePages:":
            tot += int(p[1])
    return tot


def main():
    numa = bench.bind_to_gpu_numa_node(0)
    torch.cuda.set_device(0)
    n = 1 << 20
    wl = W.build("pipeline_64", n)
    dp = Dataplane(device=0, max_batch=n, **W.sizing(wl))
    for m, k, v in wl.maps:
        assert dp.update_batch(m, as_bytes(k), as_bytes(v)) == 0
    dev = torch.device("cuda", 0)
    for prog, h, l in wl.prewarm:
        dp.run(prog, torch.from_numpy(h).to(dev).reshape(-1), torch.from_numpy(l.astype(np.int32)).to(dev), wl.now0 - 1,
               stride=64, mem=0)
        dp.sync()
    hdr = torch.from_numpy(wl.headers)
    lens = torch.from_numpy(wl.lens.astype(np.int32))
    dbuf = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    out = {"host_affinity": numa, "thp": open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
           "defrag": open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip(), "arenas": []}
    for mode in ("thp", "thp", "thp", "thp", "pinned", "pinned"):
        os.environ["BNG_HOST_ARENA"] = "pinned" if mode == "pinned" else ""
        if mode != "pinned":
            os.environ.pop("BNG_HOST_ARENA")
        t0 = time.perf_counter()
        arena = bench.host_arena(n * 64)
        t_alloc = time.perf_counter() - t0
        len_h = bench.host_arena(n * 4).view(torch.int32)[:n]
        ver_h = bench.host_arena(n)
        hk = huge_kb(arena.data_ptr(), n * 64)
        arena.view(n, 64)[:] = hdr
        bw = {}
        for name, fn in (("h2d", lambda: dbuf.copy_(arena, non_blocking=True)), ("d2h", lambda: arena.copy_(dbuf, non_blocking=True))):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            bw[name] = round(5 * n * 64 / (time.perf_counter() - t0) / 1e9, 1)
        steps = []
        dp.prof_enable(True)
        for s in range(6):
            arena.view(n, 64)[:] = hdr
            len_h.copy_(lens)
            for ring in ("spoof_events", "nat_log_rb"):
                dp.drain(ring)
            dp.sync()
            t0 = time.perf_counter()
            dp.run(wl.prog, arena, len_h, wl.now0 + s * wl.now_step, stride=64, verdict=ver_h, mem=MEM_HOST, arena_bytes=n * 64)
            steps.append(round((time.perf_counter() - t0) * 1e3, 3))
        prof = dp.prof_read()
        dp.prof_enable(False)
        out["arenas"].append({"mode": mode, "alloc_ms": round(t_alloc * 1e3, 1), "huge_kb": hk, "of_kb": n * 64 // 1024,
                              "memcpy_GBps": bw, "step_ms": steps, "Mpps_best": round(n / min(steps) / 1e3, 1),
                              "kernels_ms": {k: round(v[1] / v[0], 4) for k, v in prof.items()}})
        print(json.dumps(out["arenas"][-1]), flush=True)
    print(json.dumps({k: v for k, v in out.items() if k != "arenas"}))


if __name__ == "__main__":
    main()
