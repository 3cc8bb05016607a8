"""CPU, world_size 2, gloo: the N > 1 plumbing of the inference path (independent shards + max-over-ranks timing)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_work():
    from ctts_b200.dist import shard_of
    for n in (1, 7, 16, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_of(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo_barrier_and_max():
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        from ctts_b200 import dist as cd, synth
        rank, world, _ = cd.env_rank_world()
        dist.init_process_group("gloo")
        lo, hi = cd.shard_of(5, rank, world)
        # every rank synthesises a different batch (seed = global batch index): shards never overlap
        frames = 0
        for i in range(lo, hi):
            b = synth.ljspeech_batch(batch=2, s_max=10, s_step=1, mode="infer", seed=i)
            frames += int(b["src_lens"].sum()) * 8
        cd.barrier(world)
        slowest = cd.max_over_ranks(10.0 + rank, torch.device("cpu"), world)
        total = cd.sum_over_ranks(frames, torch.device("cpu"), world)
        if rank == 0:
            print("RESULT", slowest, int(total), hi - lo)
        dist.destroy_process_group()
    """) % os.path.join(ROOT, "comprehensive-transformer-tts_b200")
    # torchrun cannot take -c: write the program to a temp file
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    try:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29631", path], capture_output=True,
                             text=True, timeout=300)
    finally:
        os.unlink(path)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stderr[-2000:]
    _, slowest, total, n0 = line[0].split()
    assert float(slowest) == 11.0          # max over ranks, not rank 0's own time
    assert int(total) == 5 * (10 + 9) * 8  # all five batches were synthesised exactly once
    assert int(n0) == 3


def test_two_ranks_gloo_gradient_buckets():
    """The training-step exchange: bucketed gradient averaging over 2 ranks -- tied parameters once,
    frozen ones skipped, a rank without a gradient contributes zeros, several buckets."""
    code = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        from ctts_b200 import dist as cd
        rank, world, _ = cd.env_rank_world()
        dist.init_process_group("gloo")
        torch.manual_seed(0)
        a = torch.nn.Parameter(torch.zeros(300, 7))
        b = torch.nn.Parameter(torch.zeros(11))
        frozen = torch.nn.Parameter(torch.zeros(5), requires_grad=False)
        c = torch.nn.Parameter(torch.zeros(1000))
        params = [a, b, frozen, a, c]                     # `a` is registered twice (tied)
        a.grad = torch.full_like(a, 1.0 + rank)           # ranks 0 / 1: 1 and 2 -> mean 1.5
        b.grad = torch.arange(11.0) * (rank + 1)          # -> 1.5 * arange
        if rank == 0:
            c.grad = torch.full_like(c, 4.0)              # rank 1 has no gradient for c -> mean 2
        n = cd.allreduce_gradients(params, world, bucket_bytes=4096)
        ok = (torch.allclose(a.grad, torch.full_like(a, 1.5)) and torch.allclose(b.grad, torch.arange(11.0) * 1.5)
              and torch.allclose(c.grad, torch.full_like(c, 2.0)) and frozen.grad is None)
        gb = cd.GradientBuckets(params, bucket_bytes=4096)
        order = [p is c for p in gb.buckets[0]]           # reverse registration order: c first
        if rank == 0:
            print("RESULT", int(ok), n, int(order[0]), sum(len(x) for x in gb.buckets))
        dist.destroy_process_group()
    """) % os.path.join(ROOT, "comprehensive-transformer-tts_b200")
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    try:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29632", path], capture_output=True,
                             text=True, timeout=300)
    finally:
        os.unlink(path)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stderr[-2000:]
    _, ok, n_buckets, c_first, n_params = line[0].split()
    assert int(ok) == 1
    assert int(n_buckets) == 2          # [c (4000 B) + b (44 B)] | [a (8400 B > bucket size: alone)]
    assert int(c_first) == 1
    assert int(n_params) == 3           # a once, b, c; frozen skipped
