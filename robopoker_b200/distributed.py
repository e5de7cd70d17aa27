"""One process per GPU: the single exchange step of the hot path, over torch.distributed.

MCCFR (BATCHED fold): every rank samples its shard of the epoch's trees and reduces it to blocked partial sums
(`sample`), ranks all-gather those 48-byte-per-infoset partials (NCCL over NVLink on GPUs, gloo in the CPU tests), and
every rank folds the gathered buffer in rank order (`fold_gathered`) — so all tables stay bit-identical without ever
moving per-tree records.  NLHE MCCFR (ordered fold, sparse table): ranks all-gather their update records and fold
all of them (`ShardedNlhe`).  k-means: integer centroid accumulators are all-reduced (sum) between `step_local` and
`step_finish`.

The solver object only needs `sample() / fold_gathered()`; the GPU `Solver` and the CPU oracle both provide them, which
is how the world_size-2 gloo tests exercise this file without a GPU.
"""
import numpy as np


class _DeviceWords:
    """Zero-copy view of library-owned device memory for torch (`__cuda_array_interface__`)."""

    def __init__(self, ptr, nbytes, typestr="<i4", itemsize=4):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedSolver:
    """`Solver::step` across `world_size` ranks; rank r owns tree ids [r*batch, (r+1)*batch) of each epoch."""

    def __init__(self, solver, dist=None, device=None):
        import torch

        self.solver = solver
        self.dist = dist
        self.torch = torch
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.on_gpu = hasattr(solver, "delta_buffer")
        if self.on_gpu:
            solver.set_world(self.rank, self.world)
            ptr, nbytes = solver.delta_buffer()
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            self.local = torch.as_tensor(_DeviceWords(ptr, nbytes), device=self.device)
            self.gathered = torch.empty(self.world * self.local.numel(), dtype=self.local.dtype, device=self.device)
        else:
            solver.set_fold(1, self.rank, self.world)
            self.gathered = torch.empty(self.world * solver.partial_words(), dtype=torch.int32)

    def step(self, n=1):
        torch = self.torch
        for _ in range(n):
            if self.on_gpu:
                self.solver.sample()  # returns with the library stream drained
                if self.dist is not None and self.world > 1:
                    self.dist.all_gather_into_tensor(self.gathered, self.local)
                    torch.cuda.current_stream(self.device).synchronize()
                else:
                    self.gathered.copy_(self.local)
                    torch.cuda.current_stream(self.device).synchronize()
                self.solver.fold_gathered(self.gathered.data_ptr(), self.world)
            else:
                words = torch.from_numpy(self.solver.sample().view(np.int32))
                if self.dist is not None and self.world > 1:
                    self.dist.all_gather_into_tensor(self.gathered, words)
                else:
                    self.gathered.copy_(words)
                self.solver.fold_gathered(self.gathered.numpy().view(np.uint32), self.world)
        return self


class ShardedNlhe:
    """`Nlhe::step` across `world_size` ranks against replicated profile tables.  Rank r samples tree ids
    [r*batch, (r+1)*batch) of each epoch; the ranks all-gather their update records (ragged: counts first, then the padded
    record words) and every rank folds the whole epoch's records in (infoset, tree) order — tables stay bit-identical on
    all ranks and equal to a single process running world_size*batch trees.

    `solver` is the GPU `robopoker_b200.nlhe.Nlhe` or the CPU oracle (`sample_records/fold_records` on host arrays), which
    is how the world_size-2 gloo test exercises this class without a GPU."""

    def __init__(self, solver, dist=None, device=None):
        import torch

        self.solver, self.dist, self.torch = solver, dist, torch
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.on_gpu = hasattr(solver, "records")
        solver.set_world(self.rank, self.world)
        if self.on_gpu:
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            ptr, _, cap, words = solver.records()
            self.words = words
            self.local = torch.as_tensor(_DeviceWords(ptr, cap * words * 4), device=self.device).view(cap, words)

    def _gather(self, local, count):
        """Ragged all-gather of [count, words] int32 rows → one [total, words] tensor (rank order)."""
        torch, dist = self.torch, self.dist
        if dist is None or self.world == 1:
            return local[:count]
        counts = torch.zeros(self.world, dtype=torch.int64, device=local.device)
        mine = torch.tensor([count], dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(counts, mine)
        counts = counts.tolist()
        width = max(counts)
        if local.shape[0] < width:  # host path: pad this rank's rows up to the widest shard
            local = torch.cat([local, torch.zeros(width - local.shape[0], local.shape[1], dtype=local.dtype, device=local.device)])
        padded = torch.empty(self.world * width, local.shape[1], dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(padded, local[:width].contiguous())
        return torch.cat([padded[r * width:r * width + counts[r]] for r in range(self.world)]).contiguous()

    def step(self, n=1):
        torch = self.torch
        for _ in range(n):
            if self.on_gpu:
                self.solver.sample()  # returns with the library stream drained
                _, count, _, _ = self.solver.records()
                every = self._gather(self.local, count)
                torch.cuda.current_stream(self.device).synchronize()
                self.solver.fold_records(every.data_ptr(), every.shape[0])
            else:
                mine = torch.from_numpy(self.solver.sample_records())
                every = self._gather(mine, mine.shape[0])
                self.solver.fold_records(every.numpy())
        return self


def sharded_kmeans_step(layer, dist):
    """`Elkan::step_elkan` over point shards for any layer object exposing `step_local / exchange_arrays / step_finish`
    with host (numpy) accumulators — the CPU oracle in the gloo tests.  Integer sums: exact in any reduction order."""
    import torch

    layer.step_local()
    acc, tally = layer.exchange_arrays()
    if dist is not None and dist.get_world_size() > 1:
        ta = torch.from_numpy(acc.view(np.int64))
        tt = torch.from_numpy(tally.view(np.int32))
        dist.all_reduce(ta)
        dist.all_reduce(tt)
    return layer.step_finish()


def allreduce_kmeans_step(layer, dist, device=None):
    """`Elkan::step_elkan` for point-sharded ranks: local point pass, integer all-reduce of the centroid
    accumulators and tallies, then the centroid/drift update — identical on every rank."""
    import torch

    layer.step_local()
    acc_ptr, acc_bytes, sizes_ptr, re_ptr, _ = layer.exchange_buffers()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    acc = torch.as_tensor(_DeviceWords(acc_ptr, acc_bytes, "<i8", 8), device=dev)
    sizes = torch.as_tensor(_DeviceWords(sizes_ptr, 4 * layer.k, "<i4", 4), device=dev)
    re = torch.as_tensor(_DeviceWords(re_ptr, 4, "<i4", 4), device=dev)
    torch.cuda.synchronize(dev)
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(acc)
        dist.all_reduce(sizes)
        dist.all_reduce(re)
        torch.cuda.current_stream(dev).synchronize()
    return layer.step_finish()
