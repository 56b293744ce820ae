#!/usr/bin/env python3
"""The N > 1 extras of bench.py (sharded G1 MSM / Fr NTT with in-run identity checks) on their own, without the proof
pipeline in front: torchrun --nproc-per-node N tools/run_multi_extras.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import zk_apps_b200 as z

world = int(os.environ.get("WORLD_SIZE", 1))
rank, world, local, dist = bench.dist_setup(world)
import torch
torch.cuda.set_device(local)
ctx = z.Context(local)
out = bench.extras_multi_gpu(ctx, z, dist, rank, world, local)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
