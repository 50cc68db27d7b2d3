"""Runs a few resident evaluations of one bench workload (for ncu launch lists): one_eval.py <workload> [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "svgp_c4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hp = bench.host_problem(name, 0)
arm = bench.OurArm(name, hp, 0, 1)
arm.build_resident()
for _ in range(iters):
    v = arm.eval_resident()
torch.cuda.synchronize()
print(name, float(v))
