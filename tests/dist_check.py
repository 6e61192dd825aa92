"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_check.py
Every rank uploads the same particle set, the stage calls are collective (Morton-slice compute + NCCL
all-gather / all-reduce inside libsphb), and every rank must end with the full state equal to the golden
vectors of the unmodified reference."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_util as U  # noqa: E402

sys.path.insert(0, U.GOLDEN_DIR)
from make_golden import GOLDEN  # noqa: E402
from sphcode_b200 import sample_params, lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in ("evrard_c4", "khi_disph_ac", "gresho_gsph2", "shock_tube_c1"):
        g = np.load(U.golden_path(name))
        sample, over = GOLDEN[name]
        p = sample_params(sample, **over)
        c = lib.Context(p, p["DIM"], device=local)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        c.set_distributed_id(rank, world, uid.cpu().numpy().tobytes())
        c.upload(g["ic"])
        c.initialize()
        try:
            U.assert_fields(c.particles, g["state0"], U.PRE_FIELDS + U.FORCE_FIELDS, what=f"{name} rank {rank} initialize", params=p)
            for s in (1, 2):
                dt = c.integrate()
                assert abs(dt - float(g[f"dt{s}"])) <= U.RTOL * float(g[f"dt{s}"])
                U.assert_fields(c.particles, g[f"state{s}"], U.STEP_FIELDS, what=f"{name} rank {rank} step {s}", params=p)
            print(f"rank {rank}/{world} {name}: ok", flush=True)
        except AssertionError as e:
            ok = False
            print(f"rank {rank}/{world} {name}: FAIL {e}", flush=True)
        c.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
