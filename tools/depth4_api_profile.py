#!/usr/bin/env python3
"""Where the depth-4 validation wall goes (BASELINE metric ii) through the public batch entry, per phase:
   torchrun --nproc-per-node N tools/depth4_api_profile.py      (N = 1: python tools/depth4_api_profile.py)"""
import gzip, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ["PDE_B200_PROFILE"] = "1"
import torch
import torch.distributed as dist
from pde_engine_b200.validator import GpuBatchValidator

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
u = json.load(gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt"))["depths"]["4"]["uniques"]
gv = GpuBatchValidator(None, "force_free", P=4096)
if rank == 0:
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bv = gv.prefilter(u)
        print(f"--- call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms, survivors {int(bv.survivor.sum())}", file=sys.stderr, flush=True)
    gv.shutdown()
else:
    gv.serve()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
