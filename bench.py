#!/usr/bin/env python
"""bench.py — headline benchmark of the state-validity hot path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2|config3|config4|config5]

One "step" = one pass of the fused FK + collision kernel (K1F) over one batch of 2^20 interpolated 8-DOF
states against the resident 256^3 voxel world W256, map check on.  With N > 1 (torchrun,
one rank per GPU) every rank checks its own 2^20-state shard, writes its mask straight into its slot of a
gather buffer, and the mask is gathered on every rank — `--gather nccl` (default): one ncclAllGather per step
on a communication stream; `--gather p2p`: fused into the kernel over peer memory (weak scaling).
`--workload config4` runs the 2^26-state batch of configs[3] sharded over the ranks instead (strong scaling),
`config3` the near-neighbour queries, `config5` the planning queries.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port (the reference itself cannot be
built here: ROS/KDL/FCL/octomap are absent — DESIGN.md) on the host cores.
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

S_CONFIG2 = 1 << 20
S_CONFIG4 = 1 << 26
BYTES_PER_STATE = 64.125          # 8 x f64 in + 1 bit out
METRIC = "collision-checked 8-DOF configs/sec"
UNIT = "configs/s"
WORKLOAD = "config2: 2^20 interpolated 8-DOF states per GPU vs W256 (256^3 voxels, res 0.05), map check, conservative 326-sphere model"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense bf16/fp16 TFLOP/s of this pool's B200 (cuBLAS, burst) — the denominator for the tensor-core NN scan."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json, cuBLAS bf16 burst)"
    return 2250.0, "fallback (nominal dense bf16)"


def bind_to_gpu_numa(index):
    """Run this process on the CPU cores local to GPU `index` (sysfs local_cpulist of its PCI function) so that pinned host
    buffers are allocated on the GPU's own NUMA node: the host-buffer (e2e) path is PCIe-bound and a remote node costs up to
    25 % of it.  Returns the cpulist string, or None when it cannot be determined."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bus[-12:].lower()
        cpulist = open(path).read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpulist
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, util, int(reasons)))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["nvml unavailable"]}
        nv = self.nv
        loaded = [s for s in self.samples if s[1] > 0] or self.samples
        sm = sorted(s[0] for s in loaded)[len(loaded) // 2]
        bits = 0
        for s in loaded:
            bits |= s[2]
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        return {"sm_mhz": sm, "sm_max_mhz": self.max_sm, "reasons": [k for k, v in names.items() if bits & v],
                "samples": len(loaded)}


def build_inputs(n_states, seed):
    from squirrel_motion_planner_b200 import synthetic as syn
    return syn.state_batch(n_states, seed=seed)


def device_states(torch, dev, n_states, seed):
    """n_states interpolated edge states (the config-2 generator) in device memory, built 2^20 at a time so that the host
    never holds more than one 64 MiB chunk."""
    from squirrel_motion_planner_b200 import synthetic as syn
    out = torch.empty((n_states, 8), dtype=torch.float64, device=dev)
    chunk = 1 << 20
    for i, s0 in enumerate(range(0, n_states, chunk)):
        n = min(chunk, n_states - s0)
        out[s0:s0 + n].copy_(torch.from_numpy(syn.state_batch(n, seed=seed + 104729 * i)))
    return out


def make_world():
    from squirrel_motion_planner_b200 import robot_desc as rd, synthetic as syn
    w = syn.box_world(g=256, n_boxes=64, seed=1234)
    origin = w.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET])   # CC:87
    return w, origin


def host_threads():
    """All the host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would make the
    CPU arm single-threaded under the driver's multi-GPU launches: the oracle takes its thread count as an argument
    (num_threads clause), so the affinity mask is what counts."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_sample(n_threads, target_s=12.0):
    """Oracle port of the reference checker (early-exit path) on a bounded sample of the same workload."""
    import oracle
    from squirrel_motion_planner_b200 import robot_desc as rd
    lt, fx = rd.load_fixture()
    spheres = [(n, tuple(c), r) for n, c, r in fx["spheres"]]
    orc = oracle.Oracle(lt, spheres)
    w, origin = make_world()
    orc.set_grid(w.bits, w.dims, origin, w.res)
    q_all = build_inputs(S_CONFIG2, 42)
    n, dt = 21 * 2000, 0.0
    while True:                                   # grow the sample until it is worth target_s of CPU work
        q = q_all[:n]
        t0 = time.perf_counter()
        orc.check_states(q, oracle.FLAG_MAP, nthreads=n_threads)
        dt = time.perf_counter() - t0
        if dt >= 0.5 * target_s or n >= S_CONFIG2:
            break
        n = int(min(S_CONFIG2, max(2 * n, n * target_s / max(dt, 1e-3))))
    reps = 1
    if dt < 0.5 * target_s:                       # the whole batch is too quick: repeat it
        reps = int(max(1, target_s / max(dt, 1e-3)))
        t0 = time.perf_counter()
        for _ in range(reps):
            orc.check_states(q, oracle.FLAG_MAP, nthreads=n_threads)
        dt = (time.perf_counter() - t0) / reps
    return orc, n / dt, n, dt * reps


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = host_threads()
    orc, rate, n_probe, _ = cpu_baseline_sample(threads, target_s=3.0)
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    n = int(min(S_CONFIG2, max(2100, rate * min(budget_s, 8.0))))
    q = build_inputs(S_CONFIG2, 42)[:n]
    for _ in range(args.warmup):
        orc.check_states(q, oracle.FLAG_MAP, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.check_states(q, oracle.FLAG_MAP, nthreads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "first %d of the 2^20 config-2 states per step (W256, map check), oracle early-exit path" % n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "CPU oracle port of CollisionChecker::isInCollision on host cores; the reference itself "
                           "needs ROS/KDL/FCL/octomap, absent here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_own(args):
    import torch
    import torch.distributed as dist
    from squirrel_motion_planner_b200 import api, synthetic as syn

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    numa = bind_to_gpu_numa(local_rank)

    ctx = api.default_context(local_rank)      # raises if libsqmp_b200.so or the GPU is missing
    ctx.set_option("launch_overlap", 1)        # opt-in: every batch below is complete in HBM before its call is enqueued
    w, origin = make_world()
    ctx.set_voxel_grid(w.bits, w.dims, origin, w.res)

    if args.workload == "config4":
        total = S_CONFIG4
        s_rank = total // world_size
        scaling = "strong"
        n_rot = 1                                  # 2^26/8 states x 64 B = 512 MiB per rank > 126 MiB L2
        batches = None
        wl = "config4: 2^26 interpolated states sharded over %d GPU(s), W256, map check, NCCL mask all-gather" % world_size
    else:
        s_rank = S_CONFIG2
        scaling = "weak"
        n_rot = 4                                  # 4 x 64 MiB distinct batches = 256 MiB > 126 MiB L2
        batches = [build_inputs(s_rank, 42 + 7919 * (rank * n_rot + i)) for i in range(n_rot)]
        wl = WORKLOAD
    if batches is None:
        d_q = [device_states(torch, dev, s_rank, 42 + 15485863 * rank)]
        batches = [d_q[0][:S_CONFIG2].cpu().numpy()]          # the e2e leg below times one 2^20-state host batch
    else:
        d_q = [torch.from_numpy(b).to(dev) for b in batches]
    words = (s_rank + 31) // 32
    # gathered mask, every rank; NBUF buffers in rotation so that the gather of step i (comm stream) overlaps the kernels
    # of the following steps (compute stream).  The kernel writes straight into the rank's slot of the gather buffer.
    # Buffer reuse is throttled on the HOST (the launch of step i blocks until the gather of step i - NBUF has completed):
    # a stream-side wait on the other stream's event between two validity launches would cancel their launch overlap.
    NBUF = 4 if world_size > 1 else 1
    d_all = [torch.zeros(world_size * words, dtype=torch.int32, device=dev) for _ in range(NBUF)]
    d_mine = [a[rank * words:(rank + 1) * words] for a in d_all]
    flags = api.CHECK_MAP
    comm = torch.cuda.Stream(device=dev) if world_size > 1 else None
    kernel_done = [torch.cuda.Event() for _ in range(NBUF)]
    gather_done = [torch.cuda.Event() for _ in range(NBUF)]
    gathers_issued = [False] * NBUF

    # --gather p2p (default when it can be set up): the all-gather is fused into the validity kernel — every rank's
    # kernel stores its mask words straight into every rank's gather buffer through peer-mapped pointers (CUDA IPC, NVLink
    # stores); on the communication stream, behind the kernel, a one-warp kernel publishes the step in every rank's flag
    # array and another waits for all ranks' flags (= the mask of the step has reached this rank).
    # --gather nccl: one ncclAllGather per step.
    pg, gather_mode, d_status = None, "nccl", None
    if world_size > 1 and args.gather in ("auto", "p2p"):
        try:
            from squirrel_motion_planner_b200.peer_gather import PeerGather
            pg = PeerGather(rank, world_size, words, NBUF)
            d_status = torch.zeros(1, dtype=torch.int32, device=dev)
            gather_mode = "p2p-fused"
        except Exception as e:                        # e.g. CUDA IPC not permitted in this container
            if args.gather == "p2p":
                raise
            if rank == 0:
                print("bench: fused peer gather unavailable (%s), using NCCL" % e, file=sys.stderr)
            pg = None
        ok = torch.tensor([1 if pg is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # all ranks or none
        if int(ok.item()) == 0 and pg is not None:
            pg.close()
            pg, gather_mode = None, "nccl"
    step_no = [0]

    def step(i):
        b = i % NBUF
        cur = torch.cuda.current_stream()
        if world_size > 1 and gathers_issued[b]:
            gather_done[b].synchronize()              # buffer b is free again once its previous gather has finished
        if pg is not None:
            step_no[0] += 1
            ctx.check_states_gather_dev(d_q[i % n_rot], s_rank, flags, pg.structs[b], stream)
            kernel_done[b].record(cur)
            with torch.cuda.stream(comm):             # behind the kernel: publish the step, then wait for every rank's
                comm.wait_event(kernel_done[b])
                ctx.gather_signal_dev(pg.structs[b], step_no[0], comm.cuda_stream)
                ctx.gather_wait_dev(pg.flags_ptr, world_size, step_no[0], d_status, comm.cuda_stream)
                gather_done[b].record(comm)
            gathers_issued[b] = True
            return
        ctx.check_states_dev(d_q[i % n_rot], s_rank, flags, d_mine[b], stream)
        if world_size > 1:
            kernel_done[b].record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(kernel_done[b])
                dist.all_gather_into_tensor(d_all[b], d_mine[b])
                gather_done[b].record(comm)
            gathers_issued[b] = True

    if pg is not None:                                # the fused gather must deliver what ncclAllGather delivers
        ctx.check_states_dev(d_q[0], s_rank, flags, d_mine[0], stream)
        dist.all_gather_into_tensor(d_all[0], d_mine[0])
        step(0)
        torch.cuda.current_stream().wait_stream(comm)
        torch.cuda.synchronize()
        got = torch.empty_like(d_all[0])
        pg.copy_buffer_to(0, got)
        same = torch.tensor([1 if (torch.equal(got, d_all[0]) and int(d_status.item()) == 0) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if int(same.item()) == 0:
            raise RuntimeError("fused peer gather does not reproduce the NCCL all-gather")

    def drain():                                      # every all-gather issued so far has completed
        if world_size > 1:
            torch.cuda.current_stream().wait_stream(comm)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    drain()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
        torch.cuda.synchronize()
    l0 = ctx.kernel_launches
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for i in range(args.steps):                       # the timed region: nothing but the steps (no events in between,
        step(i)                                       # so consecutive K1F grids may overlap their tails as in production)
    drain()                                           # the region ends when the last mask has reached every rank
    e_stop.record()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
        torch.cuda.synchronize()
    launches = ctx.kernel_launches - l0
    total_ms = e_start.elapsed_time(e_stop)
    if pg is not None and int(d_status.item()) != 0:
        raise RuntimeError("fused peer gather: a rank's flags did not arrive (timeout)")
    if world_size > 1:                                # the kernel alone (roofline): same launches without the collective
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(args.steps):
            ctx.check_states_dev(d_q[i % n_rot], s_rank, flags, d_mine[0], stream)
        k1.record()
        torch.cuda.synchronize()
        kern_ms = k0.elapsed_time(k1) / args.steps
    else:
        kern_ms = total_ms / args.steps               # the region holds exactly `steps` launches of the one kernel
    if world_size > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = s_rank * world_size * args.steps / (total_ms * 1e-3)

    # ---- e2e: the host-buffer C-ABI call (pinned host memory in, mask back on the host), same workload ----
    s_e2e = len(batches[0])
    h_q = torch.from_numpy(batches[0]).pin_memory()
    h_mask = torch.zeros((s_e2e + 31) // 32, dtype=torch.int32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        ctx.check_states(h_q, flags, out=h_mask)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.check_states(h_q, flags, out=h_mask)      # synchronous: H2D + K1 + D2H inside
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world_size > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = s_e2e * world_size * e2e_steps / e2e_s
    n_colliding = int(np.unpackbits(h_mask.numpy().view(np.uint8)).sum())
    # the ceiling of that leg: what the host -> device copy of the same pinned batch reaches with all ranks copying at once
    d_tmp = torch.empty_like(d_q[0][:s_e2e])
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d_tmp.copy_(h_q, non_blocking=True)
    torch.cuda.synchronize()
    h2d_s = (time.perf_counter() - t0) / 5
    if world_size > 1:
        t = torch.tensor([h2d_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_s = float(t.item())
    h2d_gbs = s_e2e * 64 / h2d_s / 1e9
    del d_tmp

    extra = {}
    if not args.no_extra and world_size > 1 and args.workload != "config4":
        extra.update(multi_gpu_extras(ctx, torch, dist, dev, stream, rank, world_size, args))   # collective: every rank takes part
    if rank == 0 and not args.no_extra:
        extra.update(side_metrics(ctx, torch, dev, stream, d_q, min(s_rank, S_CONFIG2)))
    extra["h2d_gbs_per_gpu_all_ranks_copying"] = h2d_gbs
    extra["e2e_frac_of_h2d_ceiling"] = (e2e_value / world_size) * 64 / 1e9 / h2d_gbs
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = s_rank * BYTES_PER_STATE / (kern_ms * 1e-3) / 1e9
        cpu = None
        if world_size == 1 and not args.no_cpu:
            import oracle
            threads = host_threads()
            orc_, rate, n, dt = cpu_baseline_sample(threads, target_s=12.0)
            q1 = build_inputs(S_CONFIG2, 42)[:21 * 400]
            t0 = time.perf_counter()
            orc_.check_states(q1, oracle.FLAG_MAP, nthreads=1)
            single = len(q1) / (time.perf_counter() - t0)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "first %d of the 2^20 config-2 states (repeated), W256, map check, oracle early-exit path, "
                             "%.1f s of CPU work timed" % (n, dt),
                   "single_thread": {"value": single, "unit": UNIT, "sample": "first %d states, one thread (the reference's "
                                     "checker is single-threaded: CC:141-142)" % len(q1)}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "states_per_gpu": s_rank, "flags": "MAP",
                       "l2": "inputs rotate over %d distinct batches (%d MiB > 126 MiB L2); the 18 MiB grid + distance "
                             "field stay resident by design" % (n_rot, n_rot * s_rank * 64 >> 20),
                       "colliding_fraction": n_colliding / s_rank,
                       "entry": "sqmp_check_states_gather_dev" if pg is not None else "sqmp_check_states_dev",
                       "mask_gather": gather_mode if world_size > 1 else "none (1 GPU)",
                       "kernel": "K1F (bound hierarchy + slab fields, fp32 proof stages with a proven error bound, fp64 exact "
                                 "stage; verdicts identical to the all-fp64 flat kernel)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "traffic_source": None, "peak_source": peak_src,
                         "kernel": "k1_fast<false,false,SigRobotino> (K1F)", "kernel_ms": kern_ms,
                         "bytes_per_state": BYTES_PER_STATE},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": s_rank * 64, "d2h_bytes_per_step": words * 4,
                    "entry": "sqmp_check_states (pinned host buffers)", "steps": e2e_steps, "host_cpus": numa},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "extra": extra,
        }
        # dram bytes of ONE 2^20-state launch of this kernel from the round's `ncu --set full` capture (profiles/, written
        # by tools/ncu_traffic.py from the .ncu-rep; the build tag must match, else the key stays null)
        traffic_file = os.path.join(ROOT, "profiles", "r2_k1f_traffic.json")
        if os.path.exists(traffic_file) and s_rank == S_CONFIG2:
            with open(traffic_file) as f:
                tr = json.load(f)
            if tr.get("library") == api.load_library().sqmp_version().decode():
                line["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
                line["roofline"]["traffic_source"] = tr.get("source")
        print(json.dumps(line))
    if pg is not None:
        torch.cuda.synchronize()
        dist.barrier()
        pg.close()
    ctx.close()
    if world_size > 1:
        dist.destroy_process_group()


def multi_gpu_extras(ctx, torch, dist, dev, stream, rank, world_size, args):
    """N > 1, same run: (1) BASELINE.json configs[3] — the 2^26-state batch sharded over the ranks (strong scaling), every
    rank's mask all-gathered with NCCL; (2) configs[2] — the 2^16 queries sharded over the ranks against the replicated
    2^20-node tree, ids all-gathered.  Device-timed, max over ranks."""
    from squirrel_motion_planner_b200 import api, synthetic as syn
    out = {}
    s_rank = S_CONFIG4 // world_size
    d_q = device_states(torch, dev, s_rank, 42 + 15485863 * rank)
    words = s_rank // 32
    d_all = torch.zeros(world_size * words, dtype=torch.int32, device=dev)
    d_mine = d_all[rank * words:(rank + 1) * words]

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step4():
        ctx.check_states_dev(d_q, s_rank, api.CHECK_MAP, d_mine, stream)
        dist.all_gather_into_tensor(d_all, d_mine)
    ms = timed(step4, 5, 2)
    out["config4_states_per_s"] = S_CONFIG4 / ms * 1e3
    out["config4_ms_per_batch"] = ms
    out["config4_workload"] = "2^26 interpolated states sharded over %d GPUs (strong), W256, map check, ncclAllGather of the %d MiB mask" % (
        world_size, S_CONFIG4 // 8 >> 20)
    del d_q, d_all
    torch.cuda.empty_cache()
    N, Q, K = 1 << 20, 1 << 16, 20
    nodes, cost = syn.random_tree(N, 7, 9)
    ctx.tree_reset(0)
    ctx.tree_append_batch(0, nodes, cost)
    q_rank = Q // world_size
    d_qs = torch.from_numpy(syn.sample_configs(Q, 8)[rank * q_rank:(rank + 1) * q_rank]).to(dev)
    d_ids_all = torch.zeros(Q, dtype=torch.int32, device=dev)
    d_ids = d_ids_all[rank * q_rank:(rank + 1) * q_rank]
    d_low = torch.empty((q_rank, K), dtype=torch.int32, device=dev)
    d_high = torch.empty_like(d_low)
    d_cnt = torch.empty(q_rank, dtype=torch.int32, device=dev)
    d_ov = torch.zeros(1, dtype=torch.int32, device=dev)

    def step3():
        ctx.nn_nearest_dev(0, d_qs, q_rank, d_ids, stream)
        ctx.nn_radius_dev(0, d_qs, None, q_rank, 1.5, K, d_low, d_high, d_cnt, d_ov, stream)
        dist.all_gather_into_tensor(d_ids_all, d_ids)
    ms = timed(step3, 3, 1)
    out["nn_sharded_queries_per_s"] = Q / ms * 1e3
    out["nn_sharded_workload"] = "config3: 2^20-node tree replicated, 2^16 queries sharded over %d GPUs, nearest + radius/top-k per query, ids all-gathered" % world_size
    ctx.tree_reset(0)
    return out


def side_metrics(ctx, torch, dev, stream, d_qs_rot, S):
    """Secondary numbers of the same run (not the headline): FK-only, self-collision, edges entry, NN."""
    d_q = d_qs_rot[0][:S]
    from squirrel_motion_planner_b200 import api, synthetic as syn
    peak, _ = measured_peaks()

    def t_ms(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = {}
    d_mask = torch.zeros((S + 31) // 32, dtype=torch.int32, device=dev)
    # K2 is the HBM-bound kernel: its inputs and outputs rotate over a working set larger than L2 (>= 4 x 64 MiB in, 4 x 56 MiB out)
    ins = [t[:S] for t in d_qs_rot]
    while len(ins) < 4:
        ins.append(ins[0].clone())
    outs = [torch.empty((S, 7), dtype=torch.float64, device=dev) for _ in range(4)]
    d_pose = outs[0]
    k2i = [0]

    def k2():
        ctx.fk_dev(ins[k2i[0] % len(ins)], S, outs[k2i[0] % 4], stream)
        k2i[0] += 1
    ms = t_ms(k2, n=24, warm=4)
    out["fk_states_per_s"] = S / ms * 1e3
    out["fk_hbm_frac"] = S * 120 / (ms * 1e-3) / 1e9 / peak          # K2, fp64 chain (default): 120 B/state
    out["fk_working_set"] = "%d MiB in + %d MiB out, rotating (> 126 MiB L2)" % (len(ins) * S * 64 >> 20, 4 * S * 56 >> 20)
    ctx.set_option("fk_precision", 32)
    ms = t_ms(k2, n=24, warm=4)
    out["fk_fp32chain_states_per_s"] = S / ms * 1e3
    out["fk_fp32chain_hbm_frac"] = S * 120 / (ms * 1e-3) / 1e9 / peak  # K2, fp32 chain + fp64 base (<= 2e-6 m error)
    ctx.set_option("fk_precision", 64)
    ms = t_ms(lambda: ctx.check_states_dev(d_q, S, api.CHECK_MAP | api.CHECK_SELF, d_mask, stream))
    out["map_plus_self_states_per_s"] = S / ms * 1e3
    E = S // 21
    a, b = syn.edge_batch(E, 42)
    da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d_first = torch.empty(E, dtype=torch.int32, device=dev)
    ms = t_ms(lambda: ctx.check_edges_dev(da, db, E, 20, api.CHECK_MAP, d_mask, d_first, stream))
    out["edges_entry_states_per_s"] = E * 21 / ms * 1e3
    # the same batch through the host-buffer edge entry point: 2 x 64 B per 21 states over PCIe instead of 21 x 64 B
    ha, hb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    hm = torch.zeros((E * 21 + 31) // 32, dtype=torch.int32).pin_memory()
    hf = torch.zeros(E, dtype=torch.int32).pin_memory()
    for _ in range(2):
        ctx.check_edges(ha, hb, 20, api.CHECK_MAP, out_mask=hm, out_first=hf)
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.check_edges(ha, hb, 20, api.CHECK_MAP, out_mask=hm, out_first=hf)
    out["e2e_edges_entry_states_per_s"] = E * 21 * 10 / (time.perf_counter() - t0)
    d_s = torch.empty((S, 8), dtype=torch.float64, device=dev)
    d_t = torch.empty(S, dtype=torch.int32, device=dev)
    lo = np.array([-5.0, -5.0, -np.pi, -1.2, -1.7, -1.8, -2.4, -2.9])
    hi = np.array([5.0, 5.0, np.pi, 1.5, 2.6, 1.8, 2.4, 2.9])
    ms = t_ms(lambda: ctx.sample_configs_dev(1, 0, S, lo, hi, 64, d_s, d_t, stream), n=10, warm=2)
    out["sample_configs_per_s"] = S / ms * 1e3                      # K5: uniform draw + FK rejection (ee above platform)
    N, Q = 1 << 20, 1 << 16
    nodes, cost = syn.random_tree(N, 7, 9)
    ctx.tree_reset(0)
    ctx.tree_append_batch(0, nodes, cost)
    d_qs = torch.from_numpy(syn.sample_configs(Q, 8)).to(dev)
    d_ids = torch.empty(Q, dtype=torch.int32, device=dev)
    ms = t_ms(lambda: ctx.nn_nearest_dev(0, d_qs, Q, d_ids, stream), n=2, warm=1)
    out["nn_nearest_queries_per_s"] = Q / ms * 1e3
    out["nn_nearest_pairs_per_s"] = N * Q / ms * 1e3
    # the scan's estimates come out of two M128 N128 K16 fp16 MMAs per 128 x 128 pairs (nn.cu: k_nn_tc): 64 issued flop per pair
    tpeak, tsrc = measured_tensor_peak()
    out["nn_nearest_ms"] = ms
    out["nn_nearest_roofline"] = {"bound": "tensor", "achieved": N * Q * 64 / (ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                                  "frac": N * Q * 64 / (ms * 1e-3) / 1e12 / tpeak, "peak_source": tsrc,
                                  "kernel": "k_nn_tc<false> (tcgen05 fp16-split pre-filter + exact fp64 evaluation of what passes)",
                                  "note": "issued MMA flop (operands are split in two fp16 pieces); the epilogue reads every estimate "
                                          "from TMEM, which is what bounds the kernel (profiles/r2_k3_tc_nearest_ncu_summary.txt)"}
    d_low = torch.empty((Q, 20), dtype=torch.int32, device=dev)
    d_high = torch.empty_like(d_low)
    d_cnt = torch.empty(Q, dtype=torch.int32, device=dev)
    d_ov = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = t_ms(lambda: ctx.nn_radius_dev(0, d_qs, None, Q, 1.5, 20, d_low, d_high, d_cnt, d_ov, stream), n=2, warm=1)
    out["nn_radius_queries_per_s"] = Q / ms * 1e3
    out["nn_radius_ms"] = ms
    out["nn_config"] = "N=2^20 nodes, Q=2^16 queries, r=1.5, k=20"
    return out


def run_config3(args):
    """BASELINE.json configs[2]: near-neighbour search over a 2^20-node tree with 2^16 queries.  Multi-GPU per SURVEY.md
    §8e: the tree is replicated, the queries are sharded, the ids are all-gathered (strong scaling).  A step = nearest
    (K3) + radius / top-k (K4) for every query of the rank's shard."""
    import torch
    import torch.distributed as dist
    from squirrel_motion_planner_b200 import api, synthetic as syn
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = api.default_context(local_rank)
    N, Q, K = 1 << 20, 1 << 16, 20
    nodes, cost = syn.random_tree(N, 7, 9)
    ctx.tree_reset(0)
    ctx.tree_append_batch(0, nodes, cost)
    q_rank = Q // world_size
    d_qs = torch.from_numpy(syn.sample_configs(Q, 8)[rank * q_rank:(rank + 1) * q_rank]).to(dev)
    d_all = torch.zeros(Q, dtype=torch.int32, device=dev)
    d_ids = d_all[rank * q_rank:(rank + 1) * q_rank]
    d_low = torch.empty((q_rank, K), dtype=torch.int32, device=dev)
    d_high = torch.empty_like(d_low)
    d_cnt = torch.empty(q_rank, dtype=torch.int32, device=dev)
    d_ov = torch.zeros(1, dtype=torch.int32, device=dev)

    def step():
        ctx.nn_nearest_dev(0, d_qs, q_rank, d_ids, stream)
        ctx.nn_radius_dev(0, d_qs, None, q_rank, 1.5, K, d_low, d_high, d_cnt, d_ov, stream)
        if world_size > 1:
            dist.all_gather_into_tensor(d_all, d_ids)

    steps, warm = max(1, min(args.steps, 10)), max(3, min(args.warmup, 3))
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
        torch.cuda.synchronize()
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
        torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - l0
    if world_size > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    if rank == 0:
        tpeak, tsrc = measured_tensor_peak()
        ms = total_ms / steps
        algo = N * (64 + 72) + Q * (64 + 68) + Q * (4 + (2 * K + 1) * 4)      # SURVEY.md §8d rows K3 + K4
        print(json.dumps({
            "metric": "near-neighbour queries/sec (nearest + radius/top-k per query)", "value": Q / (ms * 1e-3),
            "unit": "queries/s", "n_gpus": world_size, "steps": steps, "warmup": warm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: N=2^20-node tree (replicated), Q=2^16 queries sharded over %d GPU(s), nearest + radius r=1.5 k=20, "
                                   "ids all-gathered" % world_size,
                       "l2": "the 48 MiB fp16 operand image of the tree is L2-resident by design and re-read by every 256-query tile"},
            "pairs_per_s": 2.0 * N * Q / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": 2.0 * N * Q * 64 / (ms * 1e-3) / 1e12 / world_size, "peak": tpeak, "unit": "TFLOP/s",
                         "frac": 2.0 * N * Q * 64 / (ms * 1e-3) / 1e12 / world_size / tpeak, "traffic": None, "peak_source": tsrc,
                         "kernel": "k_nn_tc<false> + k_nn_tc<true> (tcgen05 fp16-split pre-filter, fp64 exact)", "kernel_ms": ms,
                         "algorithmic_hbm_bytes": algo,
                         "note": "per GPU; issued MMA flop (64 per pair: two K=16 fp16 MMAs over split operands); the scan is bounded by "
                                 "the TMEM reads of its epilogue, the HBM traffic is tiny by construction"},
            "gpu_launches": int(launches), "clocks": sampler.summary()}))
    ctx.close()
    if world_size > 1:
        dist.destroy_process_group()


def run_config5(args):
    """BASELINE.json configs[4]: independent planning queries in a cluttered 512^3 scene, one query stream per GPU, no
    collective.  Each rank runs the C++ planner front-end (shim/sqmp_plan, GPU backend) over its share of the queries
    with a fixed iteration budget, so every GPU (and the CPU baseline) does identical work per query."""
    import subprocess
    import tempfile
    import torch
    import torch.distributed as dist
    from squirrel_motion_planner_b200 import api, robot_desc as rd, synthetic as syn
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import planner_workdir as pw

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")            # only a barrier and a max over ranks: no data-path collective
    torch.cuda.set_device(local_rank)
    shim = os.path.join(ROOT, "squirrel_motion_planner_b200", "shim")
    desc = os.path.join(ROOT, "squirrel_motion_planner_b200", "data", "robotino_plan.sqmpdesc")
    n_queries, iters = args.queries, args.iterations
    if args.threads <= 0:   # the planner front-end is host-bound (SQMP_PLAN_PROFILE=1): give every GPU its share of the host cores
        args.threads = max(1, min(8, (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)) // max(1, world_size)))
    w = syn.box_world(g=512, n_boxes=2048, seed=4321)                       # W512, SURVEY.md §8d
    # collision-free (start, goal) pairs: rejection sampling with the CUDA checker itself (self + map), seed 5
    ctx = api.default_context(local_rank)
    ctx.set_voxel_grid(w.bits, w.dims, w.origin + np.array([0.0, 0.0, rd.OCTREE_Z_OFFSET]), w.res)
    def in_collision(q):
        mask = ctx.check_states(q, api.CHECK_MAP | api.CHECK_SELF)
        return np.unpackbits(mask.view(np.uint8), bitorder="little")[:len(q)].astype(bool)
    starts, goals = pw.config5_queries(in_collision, n_queries, seed=5)
    ctx.close()
    d = tempfile.mkdtemp(prefix="sqmp_cfg5_")
    pw.write_world(d, w)
    pw.write_queries(d, starts, goals)
    per = (n_queries + world_size - 1) // world_size
    first, count = rank * per, max(0, min(per, n_queries - rank * per))
    if world_size > 1:
        dist.barrier()
    # the rank's stream of queries runs `--batch` at a time in lock step (shim/sqmp_multi.hpp: one library call per request
    # kind and round for the whole batch); --batch 1 is the strictly sequential runner (one query in flight per GPU)
    if args.batch > 1:
        cmd = [os.path.join(shim, "sqmp_plan_multi"), desc, d, "gpu", str(iters), "1", str(local_rank), str(first), str(count), str(args.batch), str(args.threads)]
    else:
        cmd = [os.path.join(shim, "sqmp_plan"), desc, d, "gpu", str(iters), "1", str(local_rank), str(first), str(count)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("sqmp_plan failed: " + r.stderr[-500:])
    if os.environ.get("SQMP_PLAN_PROFILE") == "1" and rank == 0:          # where the runner's wall time goes (shim/sqmp_multi.hpp)
        sys.stderr.write("".join(l + "\n" for l in r.stderr.splitlines() if l.startswith("#")))
    tail = [l for l in r.stdout.splitlines() if l.startswith("# queries")][-1].split()
    secs, states, its = float(tail[4]), int(tail[6]), int(tail[8])
    rows = [l.split() for l in r.stdout.splitlines() if l and not l.startswith("#") and "invalid" not in l]
    solved = sum(row[1] == "1" for row in rows)
    near_calls, rewires, informed_n = sum(int(row[5]) for row in rows), sum(int(row[9]) for row in rows), sum(int(row[16]) for row in rows)
    optimising = sum(int(row[5]) > 0 for row in rows)            # queries that reached the optimisation phase
    agg = torch.tensor([secs, float(states), float(its), float(solved), float(count), float(near_calls), float(rewires), float(informed_n),
                        float(optimising)], dtype=torch.float64)
    if world_size > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        secs = float(mx[0])
    else:
        secs = float(agg[0])
    if rank == 0:
        cpu = None
        if world_size == 1 and not args.no_cpu:
            import oracle  # noqa: F401  (the baseline binary links the oracle library)
            lt, fx = rd.load_fixture()
            pw.write_oracle_model(d, lt, [(n, tuple(c), r_) for n, c, r_ in fx["spheres"]])
            nq_cpu = max(1, min(4, count))
            t0 = time.perf_counter()
            rc = subprocess.run([os.path.join(ROOT, "oracle", "sqmp_plan_oracle"), desc, d, "oracle", str(iters), "1", "0", "0", str(nq_cpu)],
                                capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if rc.returncode == 0:
                same = [a for a in rc.stdout.splitlines() if a and not a.startswith("#")] == \
                       [a for a in r.stdout.splitlines() if a and not a.startswith("#")][:nq_cpu]
                cpu = {"value": nq_cpu / dt, "unit": "planning queries/s", "cores": 1, "kind": "port",
                       "sample": "first %d queries, same planner binary over the CPU oracle, %.1f s; decisions identical to the GPU run: %s" % (nq_cpu, dt, same)}
        total_q = float(agg[4])
        print(json.dumps({
            "metric": "planning queries/sec (fixed %d BiRRT* iterations each)" % iters, "value": total_q / secs,
            "unit": "planning queries/s", "n_gpus": world_size, "steps": 1, "warmup": 0, "ms_per_step": secs * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5: %d independent planning queries, W512 (512^3 voxels, 2048 boxes), one query stream per GPU, no collective" % n_queries,
                       "iterations_per_query": iters, "flags": "SELF|MAP", "queries_in_flight_per_gpu": max(1, args.batch) * (args.threads if args.batch > 1 else 1),
                       "host_threads_per_gpu": args.threads if args.batch > 1 else 1},
            "checked_states_per_s": float(agg[1]) / secs, "planner_iterations_per_s": float(agg[2]) / secs,
            "solved": int(agg[3]), "queries_in_optimisation_phase": int(agg[8]), "near_calls": int(agg[5]), "rewires": int(agg[6]),
            "informed_samples": int(agg[7]), "cpu_baseline": cpu, "gpu_launches": None}))
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4", "config5"])
    ap.add_argument("--queries", type=int, default=1024, help="config5: number of planning queries")
    ap.add_argument("--iterations", type=int, default=1000, help="config5: BiRRT* iterations per query")
    ap.add_argument("--threads", type=int, default=0, help="config5: host threads per GPU, each with its own context (0 = host cores / GPUs, at most 8)")
    ap.add_argument("--batch", type=int, default=512, help="config5: queries in flight per GPU (1 = sequential runner)")
    ap.add_argument("--gather", default="nccl", choices=["auto", "p2p", "nccl"],
                    help="N > 1: mask all-gather fused into the kernel over peer memory (p2p) or ncclAllGather (nccl)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config5":
        run_config5(args)
    elif args.workload == "config3":
        run_config3(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
