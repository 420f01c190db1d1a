"""Multi-GPU parity check of the Z-slab path, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/slab_check.py

Every rank steps its slab of a small dam break; rank 0 also steps the same scene unsharded on its GPU and compares
positions / velocities by uid (north_star tolerance 1e-4 relative), iteration counts and particle conservation
(sph_project_b200.slab.slab_parity_check, the same check bench.py runs at N > 1).  Prints one JSON line; exit code 1 on
a mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from sph_project_b200.slab import slab_parity_check
    method = os.environ.get("SLAB_CHECK_METHOD", "dfsph")
    steps = int(os.environ.get("SLAB_CHECK_STEPS", "30"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = slab_parity_check(rank, world, local, method=method, steps=steps, late_block=os.environ.get("SLAB_CHECK_LATE_BLOCK") == "1")
    ok = True
    if rank == 0:
        print(json.dumps(dict(res, slab_check=method)), flush=True)
        ok = res["ok"]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
