#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native FK + planning-objective hot path.

Metric (BASELINE.json): Panda FK + collision cost + gradient, configurations per second, whole job.
Workload per GPU and step: one fused launch over 65,536 trajectories x 64 waypoints = 4,194,304 Panda
configurations (the >=1M-config variant of BASELINE config 3): FK of all links, obstacle-sphere
collision (5 points x 10 spheres, EnvSpheres3D), self-collision (10 pairs), workspace bounds, EE
SE(3) goal cost, and the analytic gradient w.r.t. q.  q + cost + grad_q = 251 MB per step > 126 MB L2,
so every step streams from HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA kernels)
    python bench.py --impl reference ...                          CPU port of the reference path
    torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU, weak scaling

Prints ONE JSON line (rank 0).
"""
import argparse
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
CPU_SAMPLE = 4096 * 64       # BASELINE config 3, the bounded sample timed on the host cores

PANDA_URDF = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots", "franka_description", "robots",
                          "panda_arm_no_gripper.urdf")


def panda_objective_descriptors():
    """Panda + EnvSpheres3D objective of BASELINE config 3, built through the public drop-in classes."""
    from torch_robotics_b200.workloads import panda_spheres3d_objective
    return panda_objective_descriptors_cached(panda_spheres3d_objective)


_cache = {}


def panda_objective_descriptors_cached(fn):
    if "v" not in _cache:
        _cache["v"] = fn()
    return _cache["v"]


def sample_q(desc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = desc.lo[desc.controlled], desc.hi[desc.controlled]
    return (lo + rng.random((n, desc.n_dofs), dtype=np.float32) * (hi - lo)).astype(np.float32)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d["sm_max_mhz"]), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


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
        if self.nv is None:
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
            time.sleep(0.002)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


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


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path, restated in C (the Python reference cannot travel to
    the GPU box), all host threads, each step = BASELINE config 3 (4096 x 64 configurations)."""
    if rank != 0:
        return
    from oracle import c_oracle as co
    co.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
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
    rate = CPU_SAMPLE * args.steps / dt
    cores = co.max_threads()
    sample = f"{CPU_SAMPLE} configs/step (4096 traj x 64 waypoints, BASELINE config 3) of the same Panda objective"
    print(json.dumps({
        "impl": "reference", "metric": "panda_fk_collision_cost_grad_configs_per_s", "value": rate, "unit": "configs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": rate, "unit": "configs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(world):
    return {"workload": "panda_fused_objective_fwd_bwd: FK(11 links) + obstacle spheres(5 pts x 10, EnvSpheres3D) + "
                        "self-collision(10 pairs) + ws bounds + EE SE3 cost and grad_q",
            "configs_per_gpu_per_step": N_TRAJ * HORIZON, "trajectories": N_TRAJ, "horizon": HORIZON,
            "parallelism": f"batch-sharded x{world}, no data-path collective",
            "l2": "q+cost+grad = 251 MB per step > 126 MB L2 (inputs larger than L2, no flush needed)"}



def other_baseline_configs(dev, timed):
    """BASELINE configs 4 and 5 on one GPU: fused Adam IK (Panda, 1M seeds x 500 iterations) and the Stretch objective
    (44 links, 15 collision points, 2,097,152 configurations), each against its FP32 bound."""
    import torch
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.descriptor import build_cost_descriptor, compile_urdf
    from torch_robotics_b200.environments.env_spheres_3d import SPHERE_CENTERS, SPHERE_RADIUS
    from torch_robotics_b200.workloads import ik_example_target
    fp32_peak = 148 * 128 * 2 * measured_peaks()["sm_max_mhz"] * 1e6
    out = {}
    # config 4
    desc, _ = panda_objective_descriptors()
    model = DeviceModel(desc, dev)
    S, iters = 1_000_000, 500
    lim = np.array([[desc.lo[b], desc.hi[b]] for b in desc.controlled], np.float32)
    lo, hi = lim[:, 0] + np.float32(np.pi / 64), lim[:, 1] - np.float32(np.pi / 64)
    q0 = torch.as_tensor((lo + np.random.default_rng(7).random((S, desc.n_dofs), dtype=np.float32) * (hi - lo)).astype(np.float32), device=dev)
    Ht = ik_example_target().reshape(4, 4)
    ee = desc.link_names.index("ee_link")
    run = lambda: model.ik_adam(q0, Ht, ee, lo, hi, n_iters=iters, lr=0.2, se3_eps=0.05)
    ms = timed(run, reps=3, inner=1)
    res = run()
    out["ik_adam_1M_seeds_500_iters"] = {"ms": ms, "seed_iters_per_s": S * iters / ms * 1e3,
                                         "fp32_frac": S * iters / ms * 1e3 * 1200 / fp32_peak, "flop_per_seed_iter": 1200,
                                         "valid_fraction": float(res["valid"].float().mean().item())}
    # config 5
    urdf = os.path.join(ROOT, "torch_robotics_b200", "data", "urdf", "robots", "habitat_stretch", "urdf", "hab_stretch.urdf")
    ds = compile_urdf(urdf, "stretch")
    ms_model = DeviceModel(ds, dev)
    pts = list(ds.controlled) + [ds.link_names.index("link_straight_gripper")]
    cost = build_cost_descriptor(pt_link=pts, pt_radius=[0.05] * len(pts), cutoff_obj=0.01,
                                 objects=[(np.array(SPHERE_CENTERS, np.float32), np.full(10, SPHERE_RADIUS, np.float32), None, None)],
                                 ws_min=[-1, -1, -1], ws_max=[1, 1, 1], ee_link=ds.link_names.index("link_straight_gripper"), ee_target=Ht)
    dcs = DeviceCost(ms_model, cost)
    n5 = 1 << 21
    has = ds.has_lim[ds.controlled] > 0
    lo5 = np.where(has, ds.lo[ds.controlled], -np.pi).astype(np.float32)
    hi5 = np.where(has, ds.hi[ds.controlled], np.pi).astype(np.float32)
    q5 = torch.as_tensor((lo5 + np.random.default_rng(5).random((n5, ds.n_dofs), dtype=np.float32) * (hi5 - lo5)).astype(np.float32), device=dev)
    tot, grad = torch.empty(n5, device=dev), torch.empty_like(q5)
    ms5 = timed(lambda: dcs.objective(q5, want_terms=False, out_total=tot, out_grad=grad))
    out["stretch_objective_2M"] = {"ms": ms5, "configs_per_s": n5 / ms5 * 1e3, "fp32_frac": n5 / ms5 * 1e3 * 5880 / fp32_peak,
                                   "flop_per_config": 5880, "kernel_family": ms_model.kernel_family}
    return out

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
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
    from torch_robotics_b200._lib import DeviceCost, DeviceModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from torch_robotics_b200.distributed import bind_to_gpu_numa_node
    numa_bound = bind_to_gpu_numa_node(local_rank) if world > 1 else False
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
    q_host = torch.from_numpy(sample_q(desc, n, 1000 + rank)).pin_memory()
    q = q_host.to(dev)

    def step():
        return dc.objective(q, want_grad=True, want_terms=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        starts[i].record()
        r = step()
        ends[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    kernel_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(sum(kernel_ms))
    span_ms = starts[0].elapsed_time(ends[-1])       # first launch -> last completion, includes launch gaps

    # ---- end to end through the public API with host buffers (pinned): H2D q, fused kernel, D2H cost+grad
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

    # ---- secondary kernels of the path (BASELINE config 2): FK forward and FK fwd+bwd at 1,048,576 configs, HBM-bound
    extra = {}
    if rank == 0:
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
        lib, check, sp = None, None, None
        from torch_robotics_b200 import _lib as L_
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
        extra = {"fk_forward_1M": {"ms": ms_f, "configs_per_s": n2 / ms_f * 1e3, "GBps": by_f / ms_f / 1e6, "hbm_frac": by_f / ms_f / 1e6 / hbm,
                                   "bytes_per_config": by_f / n2},
                 "fk_backward_1M": {"ms": ms_b, "configs_per_s": n2 / ms_b * 1e3, "GBps": by_b / ms_b / 1e6, "hbm_frac": by_b / ms_b / 1e6 / hbm,
                                    "bytes_per_config": by_b / n2},
                 "fk_fwd_bwd_1M_configs_per_s": n2 / (ms_f + ms_b) * 1e3,
                 "link_jacobian_512k": {"ms": ms_j, "configs_per_s": n3 / ms_j * 1e3, "GBps": by_j / ms_j / 1e6, "hbm_frac": by_j / ms_j / 1e6 / hbm,
                                        "bytes_per_config": by_j / n3},
                 "note": "1,048,576 configs: H (738 MB) exceeds L2; CUDA events around 10 back-to-back launches, median of 5"}
        extra.update(other_baseline_configs(dev, timed))

    times = torch.tensor([span_ms, total_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    span_ms, total_ms, e2e_s = [float(v) for v in times.tolist()]

    # sanity: the result of the last step is finite and matches the e2e pipeline's output
    chk = float(r["total"][:1024].double().sum().item())
    assert np.isfinite(chk)
    assert torch.allclose(cost_host[:1024], r["total"][:1024].cpu(), rtol=1e-6, atol=1e-6)

    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = span_ms / args.steps
        value = world * n / (ms_per_step * 1e-3)
        kernel_rate = n / (total_ms / args.steps * 1e-3)          # one GPU, kernel only (CUDA events per launch)
        fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        achieved_tf = kernel_rate * FLOP_PER_CFG / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "objective_traffic.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        out = {
            "metric": "panda_fk_collision_cost_grad_configs_per_s", "value": value, "unit": "configs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "kernel_family": model.kernel_family,
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak, "traffic": traffic,
                         "flop_per_config": FLOP_PER_CFG, "bytes_per_config": BYTES_PER_CFG,
                         "kernel_ms": total_ms / args.steps, "kernel_configs_per_s": kernel_rate,
                         "peak_source": f"148 SMs x 128 lanes x 2 x sm_max_mhz ({peaks['source']} MEASURED_PEAKS.json)",
                         "hbm": {"achieved": kernel_rate * BYTES_PER_CFG / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": kernel_rate * BYTES_PER_CFG / 1e9 / peaks["hbm_gbs"]}},
            "e2e": {"value": world * n / e2e_s, "unit": "configs/s", "h2d_bytes_per_step": n * desc.n_dofs * 4,
                    "d2h_bytes_per_step": n * (desc.n_dofs + 1) * 4, "ms_per_step": e2e_s * 1e3,
                    "api": "C ABI trk_objective_host(pinned q) -> pinned cost, grad_q (8 chunks, H2D / kernel / D2H on three streams)"},
            "gpu_launches": args.steps,
            "numa_bound": numa_bound,
            "extra_kernels": extra,
            "clocks": sampler.result(),
            "wall_s_timed_region": t_wall,
        }
        if world == 1:
            rate, cores, reps = cpu_port_rate(desc, cost, CPU_SAMPLE, args.cpu_seconds)
            out["cpu_baseline"] = {"value": rate, "unit": "configs/s", "cores": cores, "kind": "port",
                                   "sample": f"{reps} x {CPU_SAMPLE} configs (BASELINE config 3 size) of the same objective, "
                                             "C port of the reference path (oracle/trk_oracle.c), OpenMP"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
