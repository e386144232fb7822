"""Run under torchrun with >= 2 ranks (one per GPU): the x-slab run must reproduce the single-GPU periodic run of the
same global system (atoms matched by global id, positions to 1e-9, pair counts equal, energy within 1e-9), with the
Langevin thermostat on (Philox keyed by the global atom id).  usage: mgpu_check.py [steps] [lj|adress|adress-cuts|tetramer]
[nve]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mrmd_b200 import api, slabs

    api.L().mrmd_b200_set_device(local)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    mode = sys.argv[2] if len(sys.argv) > 2 else "lj"
    langevin = not (len(sys.argv) > 3 and sys.argv[3] == "nve")
    res = slabs.parity_check(rank, world, steps=steps, mode=mode, langevin=langevin)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
