"""Multi-GPU plumbing of the inference path: independent shards, no data-path collective.

One process per GPU (torchrun); every rank synthesises its own batches of utterances.  The only communication is the
bookkeeping the benchmark contract asks for: a barrier on both sides of the timed region and a MAX all-reduce of the
per-rank device time.  The backend is NCCL on GPUs and gloo in the CPU tests (tests/test_dist_gloo.py).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_of(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` utterance batches for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def barrier(world):
    if world > 1:
        dist.barrier()


def max_over_ranks(value, device, world):
    """The slowest rank's time: what the whole job waits for."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device, world):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------------
# Training step: the one real exchange of the path is the gradient all-reduce of data-parallel
# training (the reference wraps the model in nn.DataParallel, train.py:38; one process per GPU replaces it).  Host logic
# only -- it runs on whatever backend the process group has (NCCL over NVLink on GPUs, gloo in tests/test_dist_gloo.py).
class GradientBuckets:
    """Flat fp32 buckets over the trainable parameters, filled in REVERSE registration order (the order in which backward
    produces gradients, so the first bucket is complete first and its all-reduce can overlap the rest of backward).

    * a Parameter registered under several names (tied: fastformer logits, conformer positional encodings) enters once;
    * frozen parameters (requires_grad False: sinusoid tables, bucket edges) are skipped;
    * a parameter whose .grad is None on this rank contributes zeros, so every rank issues identical collectives.
    """

    def __init__(self, parameters, bucket_bytes=25 << 20):
        seen, params = set(), []
        for prm in parameters:
            if prm.requires_grad and id(prm) not in seen:
                seen.add(id(prm))
                params.append(prm)
        params.reverse()
        self.buckets, cur, size = [], [], 0
        for prm in params:
            n = prm.numel() * 4
            if cur and size + n > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(prm)
            size += n
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)

    def launch(self, index, world):
        """Pack bucket `index` and start its (asynchronous) SUM all-reduce."""
        bucket = self.buckets[index]
        dev = bucket[0].device
        flat = torch.zeros(sum(p.numel() for p in bucket), device=dev, dtype=torch.float32)
        off = 0
        for p in bucket:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        self._flat[index] = flat
        self._work[index] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True) if world > 1 else None

    def finish(self, world):
        """Wait for every launched bucket, divide by the world size and write the averaged gradients back."""
        for i, bucket in enumerate(self.buckets):
            if self._flat[i] is None:
                self.launch(i, world)
            if self._work[i] is not None:
                self._work[i].wait()
            flat = self._flat[i] / world
            off = 0
            for p in bucket:
                g = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += p.numel()
            self._flat[i] = self._work[i] = None


def allreduce_gradients(parameters, world, bucket_bytes=25 << 20):
    """Average .grad over the ranks (call after backward): the non-overlapped form of GradientBuckets."""
    gb = GradientBuckets(parameters, bucket_bytes)
    for i in range(len(gb.buckets)):
        gb.launch(i, world)
    gb.finish(world)
    return len(gb.buckets)


# ---------------------------------------------------------------------------------------------------------------------
# The training step's own reducer: the gradients already live in ONE flat fp32 arena (train_engine.GradArena; param.grad are
# views of it), laid out in the order the backward pass finishes them.  So there is nothing to pack: each "bucket" is a
# contiguous slice of the arena, all-reduced IN PLACE (ncclAllReduce over NVLink / NVSwitch through torch.distributed) as
# soon as the backward pass has left the stage that writes it -- decoder + mel head first, then the variance adaptor,
# then the encoder -- on NCCL's own stream, overlapping the rest of the backward.  The 1/world of DDP's mean is folded into
# the seed of the backward pass (train_engine.run_backward), not applied as a separate pass.
STAGE_OF_PREFIX = (("postnet.", "decoder"), ("mel_linear.", "decoder"), ("decoder.", "decoder"),
                   ("variance_adaptor.", "variance_adaptor"))


def _stage_of(name):
    for prefix, stage in STAGE_OF_PREFIX:
        if name.startswith(prefix):
            return stage
    return "encoder"      # encoder.*, speaker_emb.*: final only when the whole tape has been replayed


class ArenaAllReduce:
    def __init__(self, process_group=None, max_bucket_bytes=64 << 20):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.max_bucket = max_bucket_bytes // 4
        self.works = []
        self.launched = []        # (lo, hi) element ranges, for tests / the benchmark record
        self.enabled = True

    def _runs(self, arena):
        """Contiguous slices of the arena whose slots all become final at the same stage."""
        runs = []
        for name, off, n in arena.order:
            stage = _stage_of(name)
            end = off + (n + 3) // 4 * 4
            if runs and runs[-1][0] == stage and runs[-1][2] == off:
                runs[-1][2] = end
            else:
                runs.append([stage, off, end])
        return runs

    def begin(self, ctx, arena):
        """Register the launch points of this backward pass on the tape."""
        self.works, self.launched = [], []
        ctx.hooks = []
        if self.world == 1 or not self.enabled:
            return
        runs = self._runs(arena)

        def launcher(stage):
            def fire():
                for st, lo, hi in runs:
                    if st != stage:
                        continue
                    pos = lo
                    while pos < hi:
                        nxt = min(hi, pos + self.max_bucket)
                        self.works.append(dist.all_reduce(arena.flat[pos:nxt], op=dist.ReduceOp.SUM, group=self.group,
                                                          async_op=True))
                        self.launched.append((pos, nxt))
                        pos = nxt
            return fire

        ctx.hooks.append((ctx.marks.get("decoder", 0), launcher("decoder")))
        ctx.hooks.append((ctx.marks.get("variance_adaptor", 0), launcher("variance_adaptor")))
        self._last = launcher("encoder")

    def finish(self):
        if self.world == 1 or not self.enabled:
            return
        self._last()
        for w in self.works:
            w.wait()
        self.works = []


class DistributedDataParallel(torch.nn.Module):
    """Stand-in for torch.nn.parallel.DistributedDataParallel at train.py:58 (`model = DistributedDataParallel(model,
    device_ids=[rank])`, then `model.module.state_dict()` at train.py:193): broadcasts parameters and buffers from rank 0
    at construction like DDP does, and attaches the arena reducer so that `loss.backward()` leaves the AVERAGED gradients in
    param.grad.  One import line changes in the reference's train.py."""

    def __init__(self, module, device_ids=None, output_device=None, process_group=None, bucket_cap_mb=64, **_ignored):
        super().__init__()
        self.module = module
        if dist.is_initialized() and dist.get_world_size(process_group) > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0, group=process_group)
        module._reducer = ArenaAllReduce(process_group, int(bucket_cap_mb) << 20)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
