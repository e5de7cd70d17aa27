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
    [r*batch, (r+1)*batch) of each epoch.  Two exchanges are implemented, both ending with tables that hold, on every
    rank, exactly the rows of ONE process running world_size*batch trees:

    * "owner" (default): infosets are owned by rank hash(key) mod world.  Records go to their owner (ragged all-to-all),
      each rank folds only what it owns — the fold's work per rank does not grow with the world — and the touched rows
      (176 B each) are all-gathered and written into every replica.
    * "replicated": ranks all-gather every record and each folds the whole epoch (simplest; fold work grows with world).

    `solver` is the GPU `robopoker_b200.nlhe.Nlhe` or the CPU oracle (same method names on host arrays), which is how the
    world_size-2 gloo tests exercise this class without a GPU."""

    def __init__(self, solver, dist=None, device=None, mode="owner"):
        import torch

        assert mode in ("owner", "replicated")
        self.solver, self.dist, self.torch, self.mode = solver, dist, torch, mode
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.on_gpu = hasattr(solver, "records")
        solver.set_world(self.rank, self.world)
        if self.on_gpu:
            self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            ptr, _, cap, words = solver.records()
            self.words = words
            self.local = torch.as_tensor(_DeviceWords(ptr, cap * words * 4), device=self.device).view(cap, words)

    def _counts(self, mine, device):
        """all-gather of one integer per rank."""
        torch = self.torch
        counts = torch.zeros(self.world, dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(counts, torch.tensor([mine], dtype=torch.int64, device=device))
        return counts.tolist()

    def _gather(self, local, count):
        """Ragged all-gather of [count, words] int32 rows → one [total, words] tensor (rank order)."""
        torch, dist = self.torch, self.dist
        if dist is None or self.world == 1:
            return local[:count]
        counts = self._counts(count, local.device)
        width = max(max(counts), 1)
        if local.shape[0] < width:  # pad this rank's rows up to the widest shard
            local = torch.cat([local, torch.zeros(width - local.shape[0], local.shape[1], dtype=local.dtype, device=local.device)])
        padded = torch.empty(self.world * width, local.shape[1], dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(padded, local[:width].contiguous())
        return torch.cat([padded[r * width:r * width + counts[r]] for r in range(self.world)]).contiguous()

    def _to_owners(self, grouped, counts):
        """Ragged all-to-all: rows [sum(counts), words] grouped by destination rank → the rows destined to this rank."""
        torch, dist = self.torch, self.dist
        send = torch.tensor(counts, dtype=torch.int64, device=grouped.device)
        recv = torch.zeros_like(send)
        dist.all_to_all_single(recv, send)
        recv = recv.tolist()
        out = torch.empty(sum(recv), grouped.shape[1], dtype=grouped.dtype, device=grouped.device)
        dist.all_to_all_single(out, grouped[:sum(counts)].contiguous(), output_split_sizes=recv, input_split_sizes=counts)
        return out

    def step(self, n=1):
        torch = self.torch
        for _ in range(n):
            owner = self.mode == "owner" and self.world > 1
            if self.on_gpu:
                self.solver.sample()  # returns with the library stream drained
                if owner:
                    ptr, counts = self.solver.partition_records(self.world)
                    grouped = torch.as_tensor(_DeviceWords(ptr, max(sum(counts), 1) * self.words * 4), device=self.device).view(-1, self.words)
                    mine = self._to_owners(grouped, counts)
                    torch.cuda.current_stream(self.device).synchronize()
                    self.solver.fold_records(mine.data_ptr(), mine.shape[0])
                    rptr, rcount, rwords = self.solver.touched_rows()
                    rows = torch.as_tensor(_DeviceWords(rptr, max(rcount, 1) * rwords * 4), device=self.device).view(-1, rwords)
                    every = self._gather(rows, rcount)
                    torch.cuda.current_stream(self.device).synchronize()
                    self.solver.apply_rows(every.data_ptr(), every.shape[0])
                else:
                    _, count, _, _ = self.solver.records()
                    every = self._gather(self.local, count)
                    torch.cuda.current_stream(self.device).synchronize()
                    self.solver.fold_records(every.data_ptr(), every.shape[0])
            elif owner:
                grouped, counts = self.solver.partition_records(self.world)
                mine = self._to_owners(torch.from_numpy(grouped), counts)
                self.solver.fold_records(mine.numpy())
                rows = torch.from_numpy(self.solver.touched_rows())
                self.solver.apply_rows(self._gather(rows, rows.shape[0]).numpy())
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
