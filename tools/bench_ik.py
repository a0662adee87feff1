#!/usr/bin/env python
"""BASELINE config 4: batched Adam IK for the Panda (examples/inverse_kinematics.py settings), seeds sharded
across the ranks, NCCL all-gather of the best solutions.  Prints one JSON line (rank 0).
    python tools/bench_ik.py [--seeds 1000000] [--iters 500] [--loop]     (--loop: torch.optim.Adam loop over the FK op)"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_robotics_b200 import distributed as D  # noqa: E402
from torch_robotics_b200.torch_kinematics_tree.models.robots import DifferentiableFrankaPanda  # noqa: E402
from torch_robotics_b200.workloads import ik_example_target  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=500)
ap.add_argument("--loop", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lo_i, hi_i = D.shard_bounds(args.seeds, rank, world)
S = hi_i - lo_i
panda = DifferentiableFrankaPanda(device=dev)
Ht = ik_example_target().to(dev)
torch.manual_seed(1234 + rank)
kw = dict(link_name="ee_link", batch_size=S, max_iters=args.iters, lr=2e-1, se3_eps=5e-2, eps_joint_lim=torch.pi / 64,
          print_freq=-1, debug=args.loop)
panda.inverse_kinematics(Ht, **{**kw, "max_iters": 2})     # warm-up at full size (allocator, tables)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
q, idx = panda.inverse_kinematics(Ht, **kw)
torch.cuda.synchronize()
t_solve = time.perf_counter() - t0
H = panda.compute_forward_kinematics_all_links(q.detach(), link_list=["ee_link"]).squeeze(1)
from torch_robotics_b200.torch_kinematics_tree.geometrics.utils import SE3_distance  # noqa: E402
err = SE3_distance(H, Ht)
valid = torch.zeros(S, dtype=torch.bool, device=dev)
valid[idx] = True
bq, berr, bvalid = D.gather_best_solutions(q.detach(), err, valid, k=8)
torch.cuda.synchronize()
dt = torch.tensor([t_solve], dtype=torch.float64, device=dev)
t_total = time.perf_counter() - t0
nvalid = torch.tensor([idx.numel()], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(nvalid)
if rank == 0:
    print(json.dumps({"workload": "panda_adam_ik", "impl": "torch_loop_over_fk_op" if args.loop else "fused_kernel", "seeds": args.seeds,
                      "iters": args.iters, "n_gpus": world, "seconds": float(dt), "seconds_incl_fk_check_and_best_gather": t_total, "seed_iters_per_s": args.seeds * args.iters / float(dt),
                      "valid_fraction": int(nvalid) / args.seeds, "best_err": [float(v) for v in berr.cpu()]}))
if world > 1:
    dist.destroy_process_group()
