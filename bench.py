#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native FK + planning-objective hot path.

Metric (BASELINE.json): Panda FK + collision cost + gradient, configurations per second, whole job.
Workload per GPU and step (weak scaling): one fused launch over 65,536 trajectories x 64 waypoints = 4,194,304 Panda
configurations (the >=1M-config variant of BASELINE config 3): FK of all links, obstacle-sphere collision (5 points x 10
spheres, EnvSpheres3D), self-collision (10 pairs), workspace bounds, EE SE(3) goal cost, and the analytic gradient w.r.t.
q.  q + cost + grad_q = 251 MB per step > 126 MB L2, so every step streams from HBM.  Every step also reduces the cost:
a device sum of the shard's costs and (N > 1) the NCCL all-reduce of the scalar run on a side stream, overlapped with the
next step's kernel, inside the timed region (`distributed.total_cost`).

Under torchrun (N > 1) the same JSON line also carries, per N:
  * `strong`  - the headline batch (4,194,304 configurations in total) split N ways + all-reduce;
  * `cfg4`    - BASELINE config 4: Adam IK, 1 M seeds sharded, fused kernel + NCCL all-gather of the best solutions;
  * `cfg5`    - BASELINE config 5: Stretch objective, 16 M configurations in total sharded N ways + all-reduce;
  * `shard_identity` - every rank's shard equals (bit for bit, compared by hash) the slice of rank 0's one-GPU result.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA kernels)
    python bench.py --impl reference ...                          the reference's own CPU path (unmodified, eager PyTorch)
    torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU

Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRAJ, HORIZON = 65536, 64
FLOP_PER_CFG = 2142.0        # BASELINE.md section 4 (FMA = 2), fused objective fwd+bwd, Panda
BYTES_PER_CFG = 60.0         # 28 (q) + 4 (cost) + 28 (grad_q)
CPU_SAMPLE = 4096 * 64       # BASELINE config 3
HEADLINE_KERNEL = "k_objective_hot"
METRIC = "panda_fk_collision_cost_grad_configs_per_s"


def panda_objective_descriptors():
    """Panda + EnvSpheres3D objective of BASELINE config 3, built through the public drop-in classes."""
    if "v" not in _cache:
        from torch_robotics_b200.workloads import panda_spheres3d_objective
        _cache["v"] = panda_spheres3d_objective()
    return _cache["v"]


_cache = {}


def q_rows(lo, hi, row0, n, seed):
    """Rows [row0, row0 + n) of a seeded global batch: every row is a function of (seed, row index) only, so a shard of the
    batch is the same whatever rank generates it (counter-based Philox streams, one per block of 65,536 rows)."""
    d = len(lo)
    out = np.empty((n, d), np.float32)
    blk = 65536
    r = row0
    while r < row0 + n:
        b = r // blk
        r1 = min((b + 1) * blk, row0 + n)
        u = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, b])).random((blk, d), dtype=np.float32)
        out[r - row0:r1 - row0] = lo + u[r - b * blk:r1 - b * blk] * (hi - lo)
        r = r1
    return out


def sample_q(desc, n, seed, row0=0):
    return q_rows(desc.lo[desc.controlled].astype(np.float32), desc.hi[desc.controlled].astype(np.float32), row0, n, seed)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d["sm_max_mhz"]), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def build_hash():
    """Hash of the CUDA sources the library was built from: ties a stored ncu traffic figure to a build."""
    h = hashlib.sha256()
    cs = os.path.join(ROOT, "torch_robotics_b200", "csrc")
    for f in sorted(os.listdir(cs)):
        if f.endswith((".cu", ".cuh")) and f != "robots_generated.cuh":
            h.update(open(os.path.join(cs, f), "rb").read())
    return h.hexdigest()[:16]


def stored_traffic():
    """DRAM bytes per launch of the headline kernel from the committed ncu capture - only if it was taken on this kernel
    and this build (profiles/objective_traffic.json records both); otherwise null with the reason."""
    p = os.path.join(ROOT, "profiles", "objective_traffic.json")
    if not os.path.exists(p):
        return None, "no capture committed"
    d = json.load(open(p))
    if HEADLINE_KERNEL not in str(d.get("kernel", "")):
        return None, f"capture is of kernel {d.get('kernel')!r}, not {HEADLINE_KERNEL}"
    if d.get("build_hash") != build_hash():
        return None, f"capture was taken on build {d.get('build_hash')}, this is {build_hash()}"
    return d.get("dram_bytes_per_launch"), "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, " + str(d.get("source", ""))


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
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
        if self.nv is None or os.environ.get("TRK_BENCH_NO_SAMPLER"):
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "display_clocks": 0x100}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def workload_config(world, configs_per_step=N_TRAJ * HORIZON):
    return {"workload": "panda_fused_objective_fwd_bwd: FK(11 links) + obstacle spheres(5 pts x 10, EnvSpheres3D) + "
                        "self-collision(10 pairs) + ws bounds + EE SE3 cost and grad_q (BASELINE config 3 objective)",
            "configs_per_gpu_per_step": int(configs_per_step), "horizon": HORIZON, "trajectories": int(configs_per_step) // HORIZON,
            "parallelism": f"batch-sharded x{world}; per step the kernel's per-block partial sums of the cost are summed (fp64) and the "
                           "scalar is all-reduced over NCCL on a side stream (overlapped with the next step's kernel)",
            "l2": "q+cost+grad = 251 MB per step > 126 MB L2 (inputs larger than L2, no flush needed)"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (eager PyTorch, all host threads) and the C port of the path
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(desc, cost, n_sample, min_seconds, seed=11):
    """configs/s of the C port of the reference path (oracle/trk_oracle.c) on all host threads."""
    from oracle import c_oracle as co
    co.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    tree, cc = co.Tree(desc), co.Cost(cost)
    q = sample_q(desc, n_sample, seed)
    out = dict(total=np.empty(n_sample, np.float32), terms=None, grad_q=np.empty_like(q), occ=None, argmin=None)
    co.objective(tree, cc, q, out=out)   # warm-up
    t0, reps = time.perf_counter(), 0
    while True:
        co.objective(tree, cc, q, out=out)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return n_sample * reps / dt, co.max_threads(), reps


def reference_rate(budget_s, steps=None, warmup=1, device="cpu"):
    """configs/s of the UNMODIFIED reference on BASELINE config 3's objective.  A step is a bounded sample of the
    4096 x 64 batch: as many whole trajectories (of 64 waypoints) as fit `budget_s` seconds for `steps` + `warmup`
    steps, at most 4096.  Returns (rate, cores, trajectories per step, steps, seconds per step, source path) or None
    when the reference is not present."""
    from oracle import reference_runner as rr
    if rr.locate_reference() is None:
        return None
    import torch
    cores = os.cpu_count() or 1
    obj = rr.ReferenceObjective(device, threads=cores)
    lo, hi = obj.limits()
    probe_b = 64
    qp = torch.from_numpy(q_rows(lo, hi, 0, probe_b * HORIZON, 11).reshape(probe_b, HORIZON, -1)).to(device)
    t_probe = rr.time_steps(obj, qp, 1, 1)                 # seconds for 64 trajectories (after one warm-up step)
    n_steps = steps if steps else 3
    per_traj = t_probe / probe_b
    b = int(budget_s / ((n_steps + warmup) * per_traj))
    b = max(64, min(4096, (b // 64) * 64))
    q = torch.from_numpy(q_rows(lo, hi, 0, b * HORIZON, 11).reshape(b, HORIZON, -1)).to(device)
    s_step = rr.time_steps(obj, q, n_steps, warmup)
    return b * HORIZON / s_step, cores, b, n_steps, s_step, obj.path


def run_reference(args, rank, world):
    """--impl reference: the reference's own implementation of the path on the host cores.  The unmodified reference
    (eager PyTorch, /root/reference or the baseline/_ref install) when present, all host threads; otherwise the C port of
    the path (oracle/trk_oracle.c), saying so.  Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ref = None
    try:
        ref = reference_rate(budget_s=150.0, steps=args.steps, warmup=min(args.warmup, 2))
    except Exception as e:       # an import problem of the reference must not lose the arm: fall back to the port, say why
        ref_err = f"{type(e).__name__}: {e}"
    extra = {}
    if ref is not None:
        rate, cores, b, n_steps, s_step, path = ref
        kind, cfg_per_step = "reference", b * HORIZON
        sample = (f"{b} trajectories x {HORIZON} waypoints = {cfg_per_step} configurations per step (bounded sample of BASELINE config 3's "
                  f"4096 x 64; sized so that {n_steps} steps fit ~150 s), unmodified reference at {path}: "
                  "PlanningTask.compute_collision_cost + EESE3DistanceField, forward + cost.sum().backward(), torch.set_num_threads(all)")
        ms_per_step = 1e3 * s_step
        try:
            import torch
            if torch.cuda.is_available():      # informative: the reference's eager CUDA path on one B200
                r2 = reference_rate(budget_s=40.0, steps=3, warmup=1, device="cuda:0")
                extra["reference_eager_cuda"] = {"value": r2[0], "unit": "configs/s", "configs_per_step": r2[2] * HORIZON,
                                                 "ms_per_step": 1e3 * r2[4], "note": "same objective, device='cuda:0', eager PyTorch"}
        except Exception as e:
            extra["reference_eager_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"}
    else:
        from oracle import c_oracle as co
        co.set_threads(cores)
        desc, cost = panda_objective_descriptors()
        tree, cc = co.Tree(desc), co.Cost(cost)
        q = sample_q(desc, CPU_SAMPLE, 11)
        out = dict(total=np.empty(CPU_SAMPLE, np.float32), terms=None, grad_q=np.empty_like(q), occ=None, argmin=None)
        for _ in range(args.warmup):
            co.objective(tree, cc, q, out=out)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            co.objective(tree, cc, q, out=out)
        dt = time.perf_counter() - t0
        rate, cores, kind, cfg_per_step = CPU_SAMPLE * args.steps / dt, co.max_threads(), "port", CPU_SAMPLE
        ms_per_step = 1e3 * dt / args.steps
        sample = (f"{CPU_SAMPLE} configs/step (BASELINE config 3) - the reference is not present on this box "
                  f"({locals().get('ref_err', 'neither /root/reference nor baseline/_ref')}): C port of the path (oracle/trk_oracle.c), OpenMP")
    cfg = workload_config(world, cfg_per_step)
    cfg["parallelism"] = f"host CPU, {cores} threads"
    cfg["l2"] = "n/a (CPU)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "configs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": rate, "unit": "configs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, **extra}))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def tensor_hash(*tensors):
    """64-bit hash of the raw bytes of device tensors (bit-exact comparison across ranks without moving the data)."""
    import torch
    acc = 0
    for t in tensors:
        v = t.contiguous().view(torch.int32).to(torch.int64).reshape(-1)
        idx = torch.arange(v.numel(), device=v.device, dtype=torch.int64)
        # order-sensitive: sum of value * (odd multiplier derived from the index), wrapping int64 arithmetic
        acc = (acc * 1000003 + int(((v ^ (idx * 0x9E3779B1)) * (2 * idx + 1)).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--skip-extra", action="store_true", help="headline only (development)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from torch_robotics_b200 import distributed as D
    from torch_robotics_b200._lib import DeviceCost, DeviceModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_bound = D.bind_to_gpu_numa_node(local_rank) if world > 1 else False
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own version / debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # this level printf()s the banner straight to stdout
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    desc, cost = panda_objective_descriptors()
    model = DeviceModel(desc, dev)
    dc = DeviceCost(model, cost)
    n = N_TRAJ * HORIZON
    GLOBAL_SEED = 1000
    # weak scaling: rank r owns rows [r n, (r + 1) n) of one seeded global batch
    q_host = torch.from_numpy(sample_q(desc, n, GLOBAL_SEED, row0=rank * n)).pin_memory()
    q = q_host.to(dev)
    RING = 4
    tots = [torch.empty(n, device=dev) for _ in range(RING)]
    tot = tots[0]
    grad = torch.empty_like(q)
    main_stream = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    sums = torch.zeros(1, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    REDUCE_ON_MAIN = os.environ.get("TRK_BENCH_REDUCE", "side") == "main"          # development A/B switches
    NO_COLLECTIVE = os.environ.get("TRK_BENCH_REDUCE", "side") == "none"

    def run_steps(qd, tots, gradd, cost_obj, k):
        """k steps: fused kernel on the main stream; the cost reduction of step i (device sum + all-reduce) runs on the side
        stream while the following steps compute (the cost output is a ring of len(tots) arrays, so the ranks may drift by that
        many steps before a slow all-reduce holds anybody up).  Returns (per-launch
        kernel ms list, span ms first launch -> last reduction, the last reduced total, collective ms list)."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        c0 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        c1 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        last = torch.cuda.Event(enable_timing=True)
        depth = len(tots)
        red_done = [None] * depth
        parts = [torch.empty(cost_obj.MAX_BLOCK_SUMS, device=dev) for _ in range(depth)]
        for i in range(k):
            totd = tots[i % depth]
            if red_done[i % depth] is not None:
                main_stream.wait_event(red_done[i % depth])      # step i overwrites what step i - depth's reduction read
            starts[i].record(main_stream)
            # cost + gradient + the kernel's per-block partial sums of the cost (first stage of the reduction, fused)
            cost_obj.objective_sum(qd, out_grad=gradd, out_total=totd, out_block_sums=parts[i % depth])
            ends[i].record(main_stream)
            red_stream = main_stream if REDUCE_ON_MAIN else side
            red_stream.wait_event(ends[i])
            with torch.cuda.stream(red_stream):
                s = parts[i % depth].sum(dtype=torch.float64).reshape(1)
                c0[i].record(red_stream)
                if world > 1 and not NO_COLLECTIVE:
                    dist.all_reduce(s, op=dist.ReduceOp.SUM)
                c1[i].record(red_stream)
                sums.copy_(s)
                ev = torch.cuda.Event()
                ev.record(red_stream)
                red_done[i % depth] = ev
        main_stream.wait_stream(side)
        last.record(main_stream)
        torch.cuda.synchronize()
        kernel_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
        coll_ms = [a.elapsed_time(b) for a, b in zip(c0, c1)]
        return kernel_ms, starts[0].elapsed_time(last), float(sums.item()), coll_ms

    # ---------------- headline (weak scaling) ----------------
    # Warm-up: the W steps asked for plus enough untimed steps to keep every GPU busy for ~15 ms right up to the barrier that opens
    # the timed region.  (With 8 ranks the rendezvous around a short warm-up leaves some GPUs idle long enough to drop their clocks:
    # the first ~10 timed launches then ran at 0.35 ms instead of 0.20 ms and a 20-step run reported 0.31-0.37 ms per step.)
    sampler = ClockSampler(local_rank)
    sampler.start()
    warm_extra = max(0, 64 - args.warmup)
    run_steps(q, tots, grad, dc, args.warmup)
    barrier()
    run_steps(q, tots, grad, dc, warm_extra)
    barrier()
    t_wall0 = time.perf_counter()
    kernel_ms, span_ms, total_sum, coll_ms = run_steps(q, tots, grad, dc, args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks_headline = sampler.result()
    total_ms = float(sum(kernel_ms))
    r_total = tots[(args.steps - 1) % RING].clone()

    # ---------------- shard identity + strong scaling (the headline batch split N ways) ----------------
    strong = None
    shard_identity = None
    if world > 1:
        lo_r, hi_r = D.shard_bounds(N_TRAJ, rank, world)
        ns = (hi_r - lo_r) * HORIZON
        # every rank evaluates its shard of the GLOBAL batch of rank 0 (rows [0, n)); rank 0 also evaluates the whole batch
        qs = torch.from_numpy(sample_q(desc, ns, GLOBAL_SEED, row0=lo_r * HORIZON)).to(dev)
        tss, gs = [torch.empty(ns, device=dev) for _ in range(RING)], torch.empty_like(qs)
        run_steps(qs, tss, gs, dc, 64)
        barrier()
        k_ms, s_ms, s_sum, s_coll = run_steps(qs, tss, gs, dc, args.steps)
        barrier()
        ts = tss[(args.steps - 1) % RING]
        mine = torch.tensor([tensor_hash(ts, gs) - (1 << 63)], dtype=torch.int64, device=dev)
        hashes = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(hashes, mine)
        ok = None
        if rank == 0:
            dc.objective(q, want_terms=False, out_total=tot, out_grad=grad)      # the one-GPU result of the global batch
            ok = True
            for r in range(world):
                a, b = D.shard_bounds(N_TRAJ, r, world)
                ref_h = tensor_hash(tot[a * HORIZON:b * HORIZON], grad[a * HORIZON:b * HORIZON]) - (1 << 63)
                ok = ok and (ref_h == int(hashes[r].item()))
        flag = torch.tensor([1 if (ok or rank != 0) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        shard_identity = bool(flag.item())
        assert shard_identity, "a rank's shard differs from the slice of the one-GPU result"
        t = torch.tensor([s_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"total_configs": n, "configs_per_rank": ns, "ms_per_step": float(t.item()) / args.steps,
                  "value": n / (float(t.item()) / args.steps * 1e-3), "unit": "configs/s",
                  "collective_ms": float(np.median(s_coll)) if s_coll else None,
                  "note": "fixed total (the headline batch) split in contiguous blocks of whole trajectories; time = max over ranks "
                          "of first launch -> last all-reduce"}
        del qs, ts, tss, gs

    # ---------------- end to end through the C ABI with host buffers (pinned): H2D q, fused kernel, D2H cost+grad ----------------
    from torch_robotics_b200.host_pipeline import HostObjectivePipeline
    pipe = HostObjectivePipeline(dc, n)
    cost_host = torch.empty(n, dtype=torch.float32).pin_memory()
    grad_host = torch.empty((n, desc.n_dofs), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        pipe.run(q_host, cost_host, grad_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(q_host, cost_host, grad_host)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert torch.allclose(cost_host[:1024], r_total[:1024].cpu(), rtol=1e-6, atol=1e-6)
    # copy-only ceiling of the same call on this box: the same bytes both ways at once, no kernel, all ranks concurrently
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copies():
        with torch.cuda.stream(s_in):
            q.copy_(q_host, non_blocking=True)
        with torch.cuda.stream(s_out):
            cost_host.copy_(tot, non_blocking=True)
            grad_host.copy_(grad, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    for _ in range(2):
        copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        copies()
    barrier()
    copy_s = (time.perf_counter() - t0) / 5

    # ---------------- BASELINE configs 4 (IK) and 5 (Stretch) at this N ----------------
    cfg4 = cfg5 = None
    if not args.skip_extra:
        cfg4 = bench_cfg4(torch, dist, D, model, desc, dev, rank, world)
        cfg5 = bench_cfg5(torch, dist, D, dev, rank, world, run_steps)

    # ---------------- secondary kernels of the path (one GPU): FK forward / backward, link Jacobian ----------------
    extra = {}
    if rank == 0 and world == 1 and not args.skip_extra:
        extra = secondary_kernels(torch, model, desc, q, dev)
        try:
            extra["scenes"] = scene_lines(torch, dev)
        except Exception as e:      # informative block: never lose the headline line over it
            extra["scenes"] = {"unavailable": f"{type(e).__name__}: {e}"}

    sampler.stop_flag = True
    times = torch.tensor([span_ms, total_ms, e2e_s, copy_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    span_ms, total_ms, e2e_s, copy_s = [float(v) for v in times.tolist()]
    assert np.isfinite(total_sum)

    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = span_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        kernel_rate = n / (total_ms / args.steps * 1e-3)          # one GPU, kernel only (CUDA events per launch, max over ranks)
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        achieved_tf = kernel_rate * FLOP_PER_CFG / 1e12
        traffic, traffic_note = stored_traffic()
        h2d, d2h = n * desc.n_dofs * 4, n * (desc.n_dofs + 1) * 4
        out = {
            "metric": METRIC, "value": value, "unit": "configs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "warmup_extra_untimed_steps": warm_extra, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "kernel_family": model.kernel_family, "headline_kernel": HEADLINE_KERNEL, "build_hash": build_hash(),
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak, "traffic": traffic, "traffic_note": traffic_note,
                         "flop_per_config": FLOP_PER_CFG, "bytes_per_config": BYTES_PER_CFG,
                         "kernel_ms": total_ms / args.steps, "kernel_configs_per_s": kernel_rate,
                         "peak_source": f"148 SMs x 128 lanes x 2 x sm_max_mhz ({peaks['source']} MEASURED_PEAKS.json)",
                         "hbm": {"achieved": kernel_rate * BYTES_PER_CFG / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": kernel_rate * BYTES_PER_CFG / 1e9 / peaks["hbm_gbs"]}},
            "collective": {"op": "fused per-block partial sums (trk_objective_sums) -> fp64 sum of <= 8192 values + NCCL all-reduce(SUM) of the scalar, once per step, side stream",
                           "collective_ms": float(np.median(coll_ms)) if (coll_ms and world > 1) else 0.0,
                           "total_cost_last_step": total_sum},
            "e2e": {"value": world * n / e2e_s, "unit": "configs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3,
                    "copy_only_ms": copy_s * 1e3, "copy_only_configs_per_s": world * n / copy_s,
                    "frac_of_copy_ceiling": copy_s / e2e_s,
                    "api": "C ABI trk_objective_host(pinned q) -> pinned cost, grad_q (8 chunks, H2D / kernel / D2H on three streams); "
                           "copy_only = the same bytes both ways at once without the kernel, all ranks concurrently (the box's ceiling)"},
            "gpu_launches": args.steps,          # one launch of the fused objective kernel per step (+ a memset and an 8192-value sum of its partials)
            "numa_bound": numa_bound,
            "clocks": {**sampler.result(), "samples_during_headline": clocks_headline["samples"],
                       "note": "sampled at 1 kHz over every timed region of this run (headline, strong, e2e, configs 4 and 5)"},
            "wall_s_timed_region": t_wall,
        }
        if strong is not None:
            out["strong"] = strong
            out["shard_identity"] = shard_identity
        if cfg4 is not None:
            out["cfg4"] = cfg4
        if cfg5 is not None:
            out["cfg5"] = cfg5
        if extra:
            out["extra_kernels"] = extra
        if world == 1:
            port_rate, port_cores, reps = cpu_port_rate(desc, cost, CPU_SAMPLE, args.cpu_seconds)
            port = {"value": port_rate, "unit": "configs/s", "cores": port_cores, "kind": "port",
                    "sample": f"{reps} x {CPU_SAMPLE} configs (BASELINE config 3 size) of the same objective, "
                              "C port of the reference path (oracle/trk_oracle.c), OpenMP"}
            ref = None
            try:
                ref = reference_rate(budget_s=20.0, steps=2, warmup=1)
            except Exception as e:
                port["reference_unavailable"] = f"{type(e).__name__}: {e}"
            if ref is not None:
                rate, cores, b, n_steps, s_step, path = ref
                out["cpu_baseline"] = {"value": rate, "unit": "configs/s", "cores": cores, "kind": "reference",
                                       "sample": f"{n_steps} steps of {b} trajectories x {HORIZON} waypoints of BASELINE config 3's objective, "
                                                 f"unmodified reference at {path} (eager PyTorch, all host threads)",
                                       "port": port}
            else:
                out["cpu_baseline"] = port
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_cfg4(torch, dist, D, model, desc, dev, rank, world, S=1_000_000, iters=500, k_best=8):
    """BASELINE config 4: batched Adam IK for the Panda (examples/inverse_kinematics.py:37-42 settings), S seeds sharded over
    the ranks, fused kernel + all-gather of each rank's best solutions (robot_tree.py:303-384).  Device-timed end to end."""
    from torch_robotics_b200.workloads import ik_example_target
    lo_r, hi_r = D.shard_bounds(S, rank, world)
    s_loc = hi_r - lo_r
    lim = np.array([[desc.lo[b], desc.hi[b]] for b in desc.controlled], np.float32)
    lo, hi = lim[:, 0] + np.float32(np.pi / 64), lim[:, 1] - np.float32(np.pi / 64)
    q0 = torch.from_numpy(q_rows(lo, hi, lo_r, s_loc, 7)).to(dev)
    Ht = ik_example_target().reshape(4, 4)
    ee = desc.link_names.index("ee_link")
    peaks = measured_peaks()
    fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6

    def once():
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        res = model.ik_adam(q0, Ht, ee, lo, hi, n_iters=iters, lr=0.2, se3_eps=0.05, best_k=k_best)
        b.record()
        best = D.gather_best_solutions(res["q"], res["err"], res["valid"], k=k_best, local_keys=res["best_keys"])
        c.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b), b.elapsed_time(c), res, best
    once()
    runs = []
    res = best = None
    for _ in range(5):
        # only the timings are kept: holding every run's result tensors made the caching allocator go back to cudaMalloc on
        # some runs, which showed up as 60 - 120 ms "kernel" times on all ranks at once
        res = best = None
        if world > 1:
            dist.barrier()
        tk, tc, res, best = once()
        runs.append((tk, tc))
    ms_k = float(np.median([r[0] for r in runs]))
    ms_c = float(np.median([r[1] for r in runs]))
    print(f"[cfg4 rank {rank}] kernel ms per run {[round(r[0], 3) for r in runs]}, selection ms {[round(r[1], 3) for r in runs]}", file=sys.stderr)
    t = torch.tensor([ms_k, ms_c, ms_k + ms_c], dtype=torch.float64, device=dev)
    nv = res["valid"].sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nv, op=dist.ReduceOp.SUM)
    ms_k, ms_c, ms_t = [float(v) for v in t.tolist()]
    best_k = torch.tensor([min(r[0] for r in runs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(best_k, op=dist.ReduceOp.MAX)
    return {"workload": "panda_adam_ik", "seeds_total": S, "seeds_per_rank": s_loc, "iters": iters, "ms_kernel": ms_k, "ms_kernel_best_of_5": float(best_k.item()),
            "collective_ms": ms_c,
            "ms_total": ms_t, "seed_iters_per_s": S * iters / (ms_t * 1e-3), "fp32_frac_per_gpu": S * iters / (ms_k * 1e-3) * 1200 / fp32_peak / world,
            "flop_per_seed_iter": 1200, "valid_fraction": float(nv.item()) / S, "best_err": [float(v) for v in best[1].tolist()],
            "collective": f"per-block top-{k_best} in the IK kernel (trk_ik_adam_best) -> per-rank top-{k_best} + NCCL all-gather of [{k_best}, D+2] + "
                          f"top-{k_best} (distributed.gather_best_solutions)"}


def bench_cfg5(torch, dist, D, dev, rank, world, run_steps, n_total=1 << 24):
    """BASELINE config 5: Habitat Stretch (44 links, 14 DoF) objective with gradient, n_total configurations in total sharded
    over the ranks (contiguous blocks), total cost all-reduced.  Device-timed, max over ranks."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.descriptor import build_cost_descriptor, compile_urdf
    from torch_robotics_b200.environments.env_spheres_3d import SPHERE_CENTERS, SPHERE_RADIUS
    from torch_robotics_b200.workloads import ik_example_target
    urdf = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots", "habitat_stretch", "urdf", "hab_stretch.urdf")
    ds = compile_urdf(urdf, "stretch")
    mdl = DeviceModel(ds, dev)
    pts = list(ds.controlled) + [ds.link_names.index("link_straight_gripper")]
    cost = build_cost_descriptor(pt_link=pts, pt_radius=[0.05] * len(pts), cutoff_obj=0.01,
                                 objects=[(np.array(SPHERE_CENTERS, np.float32), np.full(10, SPHERE_RADIUS, np.float32), None, None)],
                                 ws_min=[-1, -1, -1], ws_max=[1, 1, 1], ee_link=ds.link_names.index("link_straight_gripper"),
                                 ee_target=ik_example_target().reshape(4, 4))
    dcs = DeviceCost(mdl, cost)
    lo_r, hi_r = D.shard_bounds(n_total, rank, world)
    ns = hi_r - lo_r
    has = ds.has_lim[ds.controlled] > 0
    lo5 = np.where(has, ds.lo[ds.controlled], -np.pi).astype(np.float32)
    hi5 = np.where(has, ds.hi[ds.controlled], np.pi).astype(np.float32)
    # generated on the device in slabs (a 16 M x 14 host array is 940 MB): uniform in the limits, seeded per global slab
    q5 = torch.empty((ns, ds.n_dofs), dtype=torch.float32, device=dev)
    lo_t, hi_t = torch.from_numpy(lo5).to(dev), torch.from_numpy(hi5).to(dev)
    slab = 1 << 20
    for s0 in range(0, ns, slab):
        s1 = min(ns, s0 + slab)
        g = torch.Generator(device=dev).manual_seed(5000 + (lo_r + s0) // slab)
        q5[s0:s1] = lo_t + torch.rand((s1 - s0, ds.n_dofs), generator=g, device=dev) * (hi_t - lo_t)
    tots, grad = [torch.empty(ns, device=dev) for _ in range(4)], torch.empty_like(q5)
    run_steps(q5, tots, grad, dcs, 3)
    if world > 1:
        dist.barrier()
    k = 4
    k_ms, span, s_sum, coll = run_steps(q5, tots, grad, dcs, k)
    t = torch.tensor([span / k, float(np.mean(k_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_k = [float(v) for v in t.tolist()]
    fp32_peak = 148 * 128 * 2 * measured_peaks()["sm_max_mhz"] * 1e6
    return {"workload": "stretch_fused_objective_fwd_bwd (15 collision points x 10 spheres, ws bounds, EE SE3; 44 links, 14 DoF)",
            "configs_total": n_total, "configs_per_rank": ns, "ms_per_step": ms, "ms_kernel": ms_k, "configs_per_s": n_total / (ms * 1e-3),
            "fp32_frac_per_gpu": ns / (ms_k * 1e-3) * 5880 / fp32_peak, "flop_per_config": 5880, "kernel_family": mdl.kernel_family,
            "collective_ms": float(np.median(coll)) if (coll and world > 1) else 0.0, "total_cost": s_sum, "scaling": "strong"}


def scene_lines(torch, dev):
    """PlanningTask.compute_cost_and_grad (the three collision fields, no EE term) on the shipped 3-D environments, 1,048,576
    configurations: the callers' usual entry point, served by the headline kernel's scene variants."""
    from torch_robotics_b200.environments.env_maze_boxes_3d import EnvMazeBoxes3D
    from torch_robotics_b200.environments.env_spheres_3d import EnvSpheres3D
    from torch_robotics_b200.environments.env_spheres_3d_extra_objects import EnvSpheres3DExtraObjects
    from torch_robotics_b200.environments.env_table_shelf import EnvTableShelf
    from torch_robotics_b200.robots.robot_panda import RobotPanda
    from torch_robotics_b200.tasks.tasks import PlanningTask
    TA = dict(device=str(dev), dtype=torch.float32)
    robot = RobotPanda(tensor_args=TA)
    b, h = 16384, 64
    q = robot.q_min + torch.rand(b, h, 7, device=dev) * (robot.q_max - robot.q_min)
    out = {}
    for name, mk in (("EnvSpheres3D", lambda: EnvSpheres3D(tensor_args=TA)),
                     ("EnvSpheres3DExtraObjects", lambda: EnvSpheres3DExtraObjects(tensor_args=TA)),
                     ("EnvTableShelf", lambda: EnvTableShelf(tensor_args=TA)),
                     ("EnvMazeBoxes3D", lambda: EnvMazeBoxes3D(tensor_args=TA)),
                     ("EnvTableShelf_voxel_sdf_1cm", lambda: EnvTableShelf(precompute_sdf_obj_fixed=True, sdf_cell_size=0.01, tensor_args=TA))):
        task = PlanningTask(env=mk(), robot=robot, tensor_args=TA)
        for _ in range(3):
            task.compute_cost_and_grad(q)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            task.compute_cost_and_grad(q)
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 10
        out[name] = {"ms": ms, "configs_per_s": b * h / ms * 1e3}
        del task
    out["note"] = "PlanningTask.compute_cost_and_grad (self + objects + workspace, no EE term), 16384 x 64 configurations, 10 calls"
    return out


def secondary_kernels(torch, model, desc, q, dev):
    """BASELINE config 2 and the link Jacobian on one GPU: HBM-bound kernels against the measured copy bandwidth."""
    from torch_robotics_b200 import _lib as L_
    n2 = 1 << 20
    q2 = q[:n2].contiguous()
    Hbuf = model.fk_forward(q2)
    Gbuf = torch.randn_like(Hbuf)
    L = desc.n_links

    def timed(fn, reps=5, inner=10):
        # `inner` back-to-back launches per event pair so that host launch overhead is hidden behind GPU work
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(inner):
                fn()
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / inner)
        return float(np.median(ts))
    gq = torch.empty_like(q2)

    def fwd():
        L_.check(L_.lib().trk_fk_forward(model.handle, q2.data_ptr(), n2, None, L, Hbuf.data_ptr(), L_.stream_ptr(dev)), "fk_forward")

    def bwd():
        L_.check(L_.lib().trk_fk_backward(model.handle, q2.data_ptr(), n2, None, L, Gbuf.data_ptr(), gq.data_ptr(), L_.stream_ptr(dev)), "fk_backward")
    ms_f, ms_b = timed(fwd), timed(bwd)
    del Hbuf, Gbuf
    n3 = 1 << 19
    Jbuf = torch.empty((n3, L, 7, desc.n_dofs), dtype=torch.float32, device=dev)

    def jac():
        L_.check(L_.lib().trk_link_jacobian(model.handle, q2.data_ptr(), n3, Jbuf.data_ptr(), L_.stream_ptr(dev)), "link_jacobian")
    ms_j = timed(jac)
    by_j = n3 * (desc.n_dofs * 4 + L * 7 * desc.n_dofs * 4)
    del Jbuf
    by_f, by_b = n2 * (desc.n_dofs * 4 + L * 64), n2 * (2 * desc.n_dofs * 4 + L * 64)
    hbm = measured_peaks()["hbm_gbs"]
    return {"fk_forward_1M": {"ms": ms_f, "configs_per_s": n2 / ms_f * 1e3, "GBps": by_f / ms_f / 1e6, "hbm_frac": by_f / ms_f / 1e6 / hbm,
                              "bytes_per_config": by_f / n2},
            "fk_backward_1M": {"ms": ms_b, "configs_per_s": n2 / ms_b * 1e3, "GBps": by_b / ms_b / 1e6, "hbm_frac": by_b / ms_b / 1e6 / hbm,
                               "bytes_per_config": by_b / n2},
            "fk_fwd_bwd_1M_configs_per_s": n2 / (ms_f + ms_b) * 1e3,
            "link_jacobian_512k": {"ms": ms_j, "configs_per_s": n3 / ms_j * 1e3, "GBps": by_j / ms_j / 1e6, "hbm_frac": by_j / ms_j / 1e6 / hbm,
                                   "bytes_per_config": by_j / n3},
            "note": "1,048,576 configs: H (738 MB) exceeds L2; CUDA events around 10 back-to-back launches, median of 5"}


if __name__ == "__main__":
    main()
