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


def test_two_ranks_data_parallel_training_step_arena_allreduce():
    """world_size 2, gloo, kernels answered by the C-ABI emulator: the DistributedDataParallel stand-in leaves in
    param.grad the MEAN over ranks of the per-rank gradients (train.py:58 semantics), reduced in place on slices of the
    flat arena that are launched stage by stage during the backward pass."""
    code = textwrap.dedent("""
        import os, sys
        for p in %r:
            sys.path.insert(0, p)
        os.environ["CTTS_DROPOUT"] = "0"
        import torch, torch.distributed as dist
        import cases, train_checks
        from oracle import capi_emulator as emu
        from ctts_b200 import dist as cd, synth
        import ctts_b200

        class MP:
            def setattr(self, o, n, v): setattr(o, n, v)
        emu.install(MP())
        rank, world, _ = cd.env_rank_world()
        dist.init_process_group("gloo")

        def grads_for(seed, ddp):
            (p, m, t), sd, _ = cases.build_case("fs2_train")
            m["transformer_fs2"]["encoder_layer"] = 1
            m["transformer_fs2"]["decoder_layer"] = 1
            from ctts_b200 import spec
            entries, _, _ = spec.parameter_spec(p, m)
            sd = synth.synthetic_state_dict(entries, pin_frames_per_phoneme=None)
            if ddp and rank == 1:                       # DDP broadcasts rank 0's parameters at construction
                sd = {k: v + 1.0 if v.is_floating_point() and "running" not in k else v for k, v in sd.items()}
            net = ctts_b200.CompTransTTS(p, m, t)
            net.load_state_dict(sd, strict=True)
            net.train()
            model = cd.DistributedDataParallel(net) if ddp else net
            batch = synth.ljspeech_batch(batch=2, s_max=12, s_step=3, mode="teacher", seed=seed)
            args, kw = cases.call_kwargs(batch)
            out = model(*args, **kw)
            cases.train_objective(out).backward()
            return net, net.grad_arena().flat.clone()

        net, mine = grads_for(100 + rank, True)
        launched = list(net._reducer.launched)
        other = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        same = all(torch.equal(o, mine) for o in other)
        if rank == 0:
            _, g0 = grads_for(100, False)
            _, g1 = grads_for(101, False)
            ref = (g0 + g1) / 2
            err = (mine - ref).abs().max().item() / ref.abs().max().item()
            arena = net.grad_arena()
            covered = sorted(launched)
            contiguous = covered[0][0] == 0 and all(a[1] == b[0] for a, b in zip(covered, covered[1:])) \\
                and covered[-1][1] == arena.flat.numel()
            views = all(p.grad.data_ptr() == v.data_ptr() for p, v in arena.params)
            print("RESULT", same, err, len(launched), contiguous, views)
        dist.destroy_process_group()
    """) % ([ROOT, os.path.join(ROOT, "comprehensive-transformer-tts_b200"), os.path.join(ROOT, "tests", "golden")],)
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    try:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29633", path], capture_output=True,
                             text=True, timeout=600)
    finally:
        os.unlink(path)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    assert line, out.stderr[-3000:]
    _, same, err, n_launch, contiguous, views = line[0].split()
    assert same == "True"                    # every rank ends up with the same gradients
    assert float(err) < 1e-5                 # = the mean of the two single-process gradients
    assert int(n_launch) >= 3                # decoder / variance adaptor / encoder stages were launched separately
    assert contiguous == "True" and views == "True"
