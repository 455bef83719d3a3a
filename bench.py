#!/usr/bin/env python
"""bench.py — Mpps of the subscriber-dataplane hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pipeline_imix] [--impl reference]

A *step* is one batch of 2^22 synthetic frames per GPU through the program of
the chosen workload (default: the full pipeline antispoof -> NAT44 -> QoS on
IMIX frames over 10 k subscribers, BASELINE.json configs[3]).  `value` is
whole-job Mpps with frames resident in HBM when the timed region starts
(CUDA events on the library's stream, max over ranks); `e2e` is the same
metric through the C-ABI call with pinned HOST buffers, host<->device copies
inside the timed region.  `--impl reference` times the reference's own eBPF C
(oracle/_ref, or the port where that library is absent) on the host cores.
Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bng_b200 import workloads as W  # noqa: E402

METRIC = "Mpps"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------
def _oracle_kind():
    from oracle import pyoracle
    if pyoracle.available("reference"):
        return "reference"
    if not pyoracle.available("port"):
        pyoracle.build("port")
    return "port"


def _cpu_worker(args):
    """One host core: its MAC-hash shard of the subscribers, private map set (as the kernel's per-CPU,
    per-RX-queue execution of the eBPF programs), `steps` passes over a bounded sample."""
    workload, n, rank, world, steps, warmup, kind = args
    from oracle.pyoracle import Oracle
    wl = W.BUILDERS[workload](n, rank, world)
    o = Oracle(kind)
    for m, k, v in wl.maps:
        from bng_b200.layouts import as_bytes
        o.update_batch(m, as_bytes(k), as_bytes(v))
    translated = []
    for prog, h, l in wl.prewarm:
        pa = o.arena(h.shape[0] * 64 + 64)
        pa[: h.shape[0] * 64] = h.reshape(-1)
        o.run(prog, pa, l.copy(), wl.now0 - 1, stride=64)
        if wl.derive is not None:
            translated.append(np.array(pa[: h.shape[0] * 64]))
    if wl.derive is not None:
        wl.headers, wl.lens = wl.derive(translated)
    off16, stride, total16 = W.slot16(wl.lens, wl.imix, wl.headers.shape[1])
    arena = o.arena(total16 * 16 + 64)
    hw = wl.headers.shape[1]

    def restore():
        if off16 is None:
            arena[: wl.n * stride].reshape(wl.n, stride)[:, :hw] = wl.headers
        else:
            a16 = arena[: total16 * 16].reshape(total16, 16)
            for g in range(hw // 16):
                a16[off16.astype(np.int64) + g] = wl.headers[:, 16 * g: 16 * g + 16]

    t_total = 0.0
    for s in range(warmup + steps):
        restore()
        lens = wl.lens.copy()
        t0 = time.perf_counter()
        o.run(wl.prog, arena, lens, wl.now0 + s * wl.now_step, off16=off16, stride=stride)
        dt = time.perf_counter() - t0
        if s >= warmup:
            t_total += dt
    return wl.n * steps, t_total


def cpu_run(workload: str, n: int, procs: int, steps: int, warmup: int):
    import multiprocessing as mp
    kind = _oracle_kind()
    jobs = [(workload, n, r, procs, steps, warmup, kind) for r in range(procs)]
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    pk = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    return pk / tmax / 1e6, kind, tmax


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 128))
    n = 1 << 18
    t0 = time.time()
    mpps, kind, tmax = cpu_run(a.workload, n, procs, a.steps, a.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": round(mpps, 3), "unit": "Mpps", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(tmax / a.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32/u64 integer", "data": "synthetic",
        "config": {"workload": a.workload, "frames_per_step": n * procs, "host_procs": procs,
                   "sharding": "subscriber MAC hash, one private map set per core"},
        "cpu_baseline": {"value": round(mpps, 3), "unit": "Mpps", "cores": procs,
                         "kind": "reference" if kind == "reference" else "port",
                         "sample": f"{procs} cores x {n} frames x {a.steps} steps of {a.workload}; "
                                   "reference eBPF C compiled natively (gcc -O2) over a userspace map runtime"},
        "e2e": {"value": round(mpps, 3), "unit": "Mpps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.time() - t0, 1),
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
_ARENAS = []  # keeps registered mappings alive


def host_arena(nbytes: int, kind: str):
    """Pinned, GPU-mapped host buffer for the frame arena.  kind 'thp': anonymous memory advised to use
    2 MB transparent huge pages, touched, then cudaHostRegister'ed; falls back to cudaHostAlloc."""
    import mmap
    import torch
    if kind == "thp" and hasattr(mmap, "MADV_HUGEPAGE"):
        try:
            huge = 2 << 20
            size = (nbytes + huge - 1) // huge * huge
            mm = mmap.mmap(-1, size + huge, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            base = np.frombuffer(mm, dtype=np.uint8)
            addr = base.ctypes.data
            off = (-addr) % huge
            mm.madvise(mmap.MADV_HUGEPAGE, 0, size + huge)
            arr = base[off:off + size]
            arr[:] = 0  # first touch on this (NUMA-bound) thread
            t = torch.from_numpy(arr)
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), size, 1 | 2)  # portable | mapped
            if int(rc) == 0:
                _ARENAS.append((mm, base, t))
                return t[:nbytes]
        except Exception as e:  # noqa: BLE001
            print(f"[bench] thp arena unavailable ({e}); using cudaHostAlloc", file=sys.stderr)
    return torch.zeros(nbytes, dtype=torch.uint8).pin_memory()


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that first-touch places
    the pinned host buffers on the GPU's NUMA node (PCIe traffic then stays off the socket interconnect)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(len(mask)) for b in range(64) if (mask[w] >> b) & 1]
        cpus = [c for c in cpus if c < os.cpu_count()]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}", "count": len(cpus)}
    except Exception as e:  # affinity is an optimisation, never a requirement
        return {"error": str(e)[:80]}
    return None
class DevPtr:
    """__cuda_array_interface__ wrapper so torch can view library-owned device memory."""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_gpu(a):
    import torch
    import torch.distributed as dist
    from bng_b200 import MEM_DEVICE, MEM_HOST, Dataplane
    from bng_b200.layouts import as_bytes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever native libraries print there while we run (NCCL's version
    # banner, for one) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    numa = bind_to_gpu_numa_node(local)  # pinned host buffers must live next to the GPU's PCIe root
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.frames
    wl = W.BUILDERS[a.workload](n, rank, world)
    n = wl.n  # a workload may hold fewer frames than asked for (nat_cold: one frame per flow)
    dp = Dataplane(device=local, max_batch=max(n, 1 << 20), rank=rank, world=world,
                   **({} if a.reference_capacities else W.sizing(wl)))
    for m, k, v in wl.maps:
        r = dp.update_batch(m, as_bytes(k), as_bytes(v))
        assert r == 0, (m, r)
    translated = []
    for prog, h, l in wl.prewarm:  # e.g. create the NAT sessions of every flow once (cold start), untimed
        ph = torch.from_numpy(h).to(dev).reshape(-1)
        pl = torch.from_numpy(l.astype(np.int32)).to(dev)
        torch.cuda.synchronize()
        dp.run(prog, ph, pl, wl.now0 - 1, stride=64, mem=MEM_DEVICE)
        dp.sync()
        if wl.derive is not None:
            translated.append(ph.cpu().numpy())
    if wl.derive is not None:  # frames that depend on what the prewarm did (return traffic of translated flows)
        wl.headers, wl.lens = wl.derive(translated)
    hw = wl.headers.shape[1]
    off16, stride, total16 = W.slot16(wl.lens, wl.imix, hw, a.align)
    hdr_d = torch.from_numpy(wl.headers).to(dev)
    len0_d = torch.from_numpy(wl.lens.astype(np.int32)).to(dev)
    len_d = len0_d.clone()
    arena_d = torch.zeros(total16 * 16 + 64, dtype=torch.uint8, device=dev)
    a16 = arena_d[: total16 * 16].view(total16, 16)
    off_d = None
    if off16 is not None:
        off_d = torch.from_numpy(off16.astype(np.int32)).to(dev)
        gidx = off_d.long()[:, None] + torch.arange(hw // 16, device=dev)[None, :]
    verdict_d = torch.zeros(n, dtype=torch.uint8, device=dev)
    lib_stream = torch.cuda.ExternalStream(dp.stream, device=dev)

    def restore():
        dp.sync()  # the previous step (asynchronous on the library's stream) must be done with the arena
        if off16 is None:
            arena_d[: n * stride].view(n, stride)[:, :hw] = hdr_d
        else:
            a16[gidx.reshape(-1)] = hdr_d.view(-1, 16)
        len_d.copy_(len0_d)
        torch.cuda.synchronize()
        reset_state()

    def reset_state():
        for ring in ("spoof_events", "nat_log_rb"):  # the event consumer keeps the staging rings empty (untimed)
            dp.drain(ring)
        if wl.name == "nat_cold_64":  # every step starts from empty flow tables and fresh port blocks
            for m in ("nat_sessions", "nat_reverse", "eim_table"):
                dp.clear(m)
            for m, k, v in wl.maps:
                if m == "subscriber_nat":
                    dp.update_batch(m, as_bytes(k), as_bytes(v))

    step_no = [0]

    def step():
        now = wl.now0 + step_no[0] * wl.now_step
        step_no[0] += 1
        dp.run(wl.prog, arena_d, len_d, now, off16=off_d, stride=stride, verdict=verdict_d, mem=MEM_DEVICE)

    for _ in range(a.warmup):
        restore()
        step()
        dp.sync()
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = dp.launch_count
    evs = []
    for _ in range(a.steps):
        restore()  # fresh frames for this step (untimed: stands in for the NIC filling the arena)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        step()
        e1.record(lib_stream)
        evs.append((e0, e1))
    dp.sync()
    torch.cuda.synchronize()
    launches = dp.launch_count - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = sum(step_ms)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    clocks = sampler.result()
    drops = int((verdict_d == 2).sum().item())
    value = world * n * a.steps / (total_ms_max * 1e-3) / 1e6

    # ---- per-kernel timing for the roofline (separate pass, events around every launch) ----
    dp.prof_enable(True)
    for _ in range(3):
        restore()
        step()
        dp.sync()
    prof = dp.prof_read()
    dp.prof_enable(False)
    top = max(prof.items(), key=lambda kv: kv[1][1])
    top_ms = top[1][1] / top[1][0]
    step_prof_ms = sum(v[1] for v in prof.values()) / 3
    peak, peak_src = peaks()
    algo = W.ALGO_BYTES[wl.name]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(wl.name, {}).get(top[0])
    achieved = algo * n / (top_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": algo, "kernel_ms": round(top_ms, 4),
                "kernel_share_of_step": round(top[1][1] / 3 / step_prof_ms, 3),
                "step_frac": round(algo * n / (total_ms_max / a.steps * 1e-3) / 1e9 / peak, 4),
                "kernels_ms": {k: round(v[1] / v[0], 4) for k, v in prof.items()}}

    # ---- end to end through the C ABI with pinned host buffers ----
    # One call = one batch of n_e frames (the first n_e of the workload).  The batch is capped at --e2e-frames
    # (2^20) because what bounds this leg is the host side: behind the box's IOMMU the GPU's scattered header
    # reads need the arena in 2 MB pages, and a 1.6 GB arena often cannot get them all (72 vs 290 Mpps seen).
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    n_e = min(n, max(1, a.e2e_frames))
    lens_e, hdrs_e = wl.lens[:n_e], wl.headers[:n_e]
    off16_e = off16[:n_e] if off16 is not None else None
    total16_e = total16 if n_e == n else (int(off16[n_e]) if off16 is not None else n_e * stride // 16)
    arena_h = host_arena(total16_e * 16 + 64, a.arena)

    def host_like(t):
        h = host_arena(t.numel() * t.element_size(), a.arena).view(t.dtype)[: t.numel()]
        h.copy_(t)
        return h

    len_h = host_like(torch.from_numpy(lens_e.astype(np.int32)))
    off_h = host_like(torch.from_numpy(off16_e.astype(np.int32))) if off16 is not None else None
    verdict_h = host_like(torch.zeros(n_e, dtype=torch.uint8))
    hdr_h = torch.from_numpy(np.ascontiguousarray(hdrs_e))
    len0_h = torch.from_numpy(lens_e.astype(np.int32))
    h16 = arena_h[: total16_e * 16].view(total16_e, 16)
    gidx_h = gidx[:n_e].reshape(-1).cpu() if off16 is not None else None

    def restore_host():
        if off16 is None:
            arena_h[: n_e * stride].view(n_e, stride)[:, :hw] = hdr_h
        else:
            h16.index_copy_(0, gidx_h, hdr_h.view(-1, 16))
        len_h.copy_(len0_h)

    arena_bytes = total16 * 16
    arena_bytes_e = total16_e * 16
    tc_prog = wl.prog != "dhcp_fastpath_prog"
    hb = 64 if tc_prog else 448  # bytes of each frame a program can touch = what crosses PCIe from a pinned arena

    def e2e_run(arena_t, off_t, strd, restore_fn, nbytes):
        tot = 0.0
        for s in range(1 + e2e_steps):
            restore_fn()
            reset_state()
            dp.sync()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            now = wl.now0 + step_no[0] * wl.now_step
            step_no[0] += 1
            dp.run(wl.prog, arena_t, len_h, now, off16=off_t, stride=strd, verdict=verdict_h, mem=MEM_HOST, arena_bytes=nbytes)
            dt = time.perf_counter() - t0
            if s >= 1:
                tot += dt
        et = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        return world * n_e * e2e_steps / float(et.item()) / 1e6

    # (a) the frames as they sit in the host arena (full frames, IMIX on 64-byte boundaries)
    e2e_val = e2e_run(arena_h, off_h, stride, restore_host, arena_bytes_e)
    per_frame_in = float(np.minimum(lens_e, hb).mean())
    h2d = int(n_e * per_frame_in) + n_e * 4 + (n_e * 4 if off16 is not None else 0)
    d2h = int(n_e * (per_frame_in - (16 if tc_prog else 0))) + n_e + (n_e * 4 if not tc_prog else 0)
    e2e_extra = None
    if tc_prog and wl.imix:
        # (b) header-split receive: the NIC put the first 64 bytes of every frame in a contiguous ring
        # (len[] still carries the full frame length); that ring is all the TC programs ever touch
        ring_h = host_arena(n_e * 64, a.arena)

        def restore_ring():
            ring_h.view(n_e, 64)[:, :hw] = hdr_h
            len_h.copy_(len0_h)

        v = e2e_run(ring_h, None, 64, restore_ring, n_e * 64)
        e2e_extra = {"value": round(v, 2), "unit": "Mpps", "layout": "header-split ring (64 B per frame, len = full frame)",
                     "h2d_bytes_per_step": n_e * 64 + n_e * 4, "d2h_bytes_per_step": n_e * 64 + n_e}

    # ---- counter reconciliation over NCCL (outside the timed region, as in production) ----
    ptr, nst = dp.stats_device_ptr()
    stats_local = torch.as_tensor(DevPtr(ptr, nst), device=dev).clone()
    stats_global = stats_local.clone()
    if world > 1:
        dist.all_reduce(stats_global, op=dist.ReduceOp.SUM)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        c_n = 1 << 20
        mpps, kind, tt = cpu_run(a.workload, c_n, 1, 4, 1)
        cpu = {"value": round(mpps, 3), "unit": "Mpps", "cores": 1, "kind": kind,
               "sample": f"1 core x {c_n} frames x 4 passes of {a.workload} ({tt:.1f} s), reference eBPF C run natively"}
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mpps", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": round(total_ms_max / a.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32/u64 integer", "data": "synthetic",
            "config": {"workload": wl.name, "program": wl.prog, "frames_per_gpu_per_step": n,
                       "subscribers_this_gpu": wl.n_subs_local, "sharding": f"splitmix64(mac) % {world}",
                       "avg_frame_bytes": round(float(wl.lens.mean()), 1), "frame_align": a.align if wl.imix else stride,
                       "step_ms_min_med_max": [round(float(x), 4) for x in
                                               (min(step_ms), float(np.median(step_ms)), max(step_ms))],
                       "step_ms_all": [round(float(x), 3) for x in step_ms],
                       "l2_policy": "inputs larger than L2 (arena %.0f MB + tables) and rewritten between steps" % (arena_bytes / 1e6),
                       **wl.info},
            "wire_gbps": round(value * 1e6 * float(wl.lens.mean()) * 8 / 1e9, 1),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_val, 2), "unit": "Mpps", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "frames_per_step": n_e,
                    "layout": "pinned host arena, full frames; only the bytes a program can touch cross PCIe"},
            "e2e_header_split": e2e_extra, "host_affinity": numa,
            "gpu_launches": int(launches), "clocks": clocks,
            "verdict_drop_fraction_last_step": round(drops / n, 4),
            "stats_allreduce": {"antispoof_allowed": int(stats_global[0].item()), "nat_snat": int(stats_global[10].item()),
                                "qos_dropped": int(stats_global[7].item())},
            "lru_overflow": int(dp.lru_overflow), "events_lost": int(dp.events_lost),
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    dp.close()
    if world > 1:
        dist.destroy_process_group()


def run_dhcp_slow(a):
    """BASELINE.json configs[0]: 1 000 DHCP DISCOVERs through the slow path with a 256-entry lease map, CPU only.
    The reference's pkg/dhcp is Go and cannot be built in this image; what is timed is its C++ restatement
    (bng_b200/host/bng_dhcp_slow.hpp, one thread — the reference serialises on the pool mutex).  Plumbing: no GPU,
    no roofline, gpu_launches 0 by construction."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_host_mirror
    test_host_mirror.build_host_test()
    rounds = max(1, a.steps) * 100
    j = json.loads(subprocess.run([test_host_mirror.SLOW_BIN, str(rounds)], capture_output=True, text=True, check=True).stdout)
    print(json.dumps({
        "metric": "DHCP DISCOVER/s (slow path, CPU only)", "value": round(j["req_per_s"], 1), "unit": "requests/s", "n_gpus": 0,
        "steps": rounds, "warmup": 0, "ms_per_step": round(j["seconds"] / rounds * 1e3, 4), "higher_is_better": True,
        "scaling": "n/a", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "dhcp_slow", "requests_per_step": 1000, "clients_with_lease": 256, "pool": "10.0.0.0/22",
                   "implementation": "C++ restatement of pkg/dhcp Server.handleDiscover + Pool.Allocate (Go toolchain absent)"},
        "roofline": None, "gpu_launches": 0,
        "cpu_baseline": {"value": round(j["req_per_s"], 1), "unit": "requests/s", "cores": 1, "kind": "port",
                         "sample": f"{j['requests']} DISCOVERs in {j['seconds']:.2f} s"},
        "e2e": {"value": round(j["req_per_s"], 1), "unit": "requests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "replies_fnv1a": j["replies_fnv1a"]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pipeline_imix", choices=sorted(W.BUILDERS) + ["dhcp_slow"],
                    help="dhcp_slow = BASELINE config #1: the DHCP slow path (CPU only, plumbing; no GPU involved)")
    ap.add_argument("--frames", type=int, default=1 << 22)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-frames", type=int, default=1 << 20, help="frames per bng_prog_run call in the end-to-end leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reference-capacities", action="store_true",
                    help="size every table for the reference's compile-time max_entries instead of the workload")
    ap.add_argument("--align", type=int, default=64, help="frame placement granularity in the IMIX arena (16 or 64)")
    ap.add_argument("--arena", default="thp", choices=["thp", "pinned"],
                    help="host frame arena of the e2e leg: 2 MB transparent huge pages registered with CUDA (what a DPDK-style "
                         "receive ring uses: far fewer IOMMU translations for the GPU's scattered header reads), or cudaHostAlloc")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.workload == "dhcp_slow":
        run_dhcp_slow(a)
    elif a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
