"""torchrun --nproc-per-node 2 tools/nlhe_world_check.py — NCCL exchange parity: `world` ranks of `batch` trees each must end
with exactly the table of ONE process running world*batch trees (rank 0 runs that single-process reference too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from robopoker_b200.comm import Comm  # noqa: E402
from robopoker_b200.distributed import ShardedNlhe  # noqa: E402
from robopoker_b200.nlhe import Nlhe  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
batch, epochs = int(os.environ.get("RBP_CHECK_BATCH", 4096)), int(os.environ.get("RBP_CHECK_EPOCHS", 4))
whole = None
comm = Comm.from_torch(dist, device=local)
for mode in ("library", "owner", "replicated"):
    s = Nlhe(batch=batch, seed=17, table_slots=1 << 22, device=local)
    if mode == "library":  # the exchange inside librbp_b200 (rbp_nlhe_attach_comm): peer-memory stores + NCCL barriers
        s.attach_comm(comm).step(epochs)
    else:                  # the host-driven exchange over torch.distributed (kept as the cross-check)
        s.set_stream(torch.cuda.current_stream().cuda_stream)
        ShardedNlhe(s, dist, device=local, mode=mode).step(epochs)
    rows = s.profile()
    digest = torch.tensor([int.from_bytes(__import__("hashlib").sha256(rows.tobytes()).digest()[:7], "little")], device="cuda")
    every = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(every, digest)
    same = all(int(e.item()) == int(every[0].item()) for e in every)
    if rank == 0:
        if whole is None:
            w = Nlhe(batch=batch * world, seed=17, table_slots=1 << 22, device=local)
            w.step(epochs)
            whole = w.profile()
            w.close()
        print(mode, "| ranks identical:", same, "| equal to one process with", batch * world, "trees:", whole.tobytes() == rows.tobytes(), "| rows", len(rows), flush=True)
    s.close()
    dist.barrier()
dist.destroy_process_group()
