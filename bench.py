#!/usr/bin/env python
"""Headline benchmark: valid mel frames / second of the CompTransTTS forward path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): transformer_fs2 + supervised-duration model (learn_alignment False),
LJSpeech shape, batch 16, src_lens 100..70, free-running inference with the duration predictor pinned to
8 frames / phoneme -> M = 800, 10 880 valid frames per step.  A "step" is one
`model(...)` call.  For N > 1 (torchrun, one rank per GPU) every rank runs its own batch of 16 utterances
(independent shards, no collective on the data path) -> weak scaling; value = all ranks' frames / max time.

`--impl reference` times the reference's CPU implementation of the same forward (the oracle port:
oracle/ctts_oracle.py, PyTorch fp32 on all host threads) on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "comprehensive-transformer-tts_b200"), os.path.join(ROOT, "tests", "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

FRAMES_PER_PHONEME = 8
BATCH = 16
METRIC = "mel-frames/sec (batch-16 LJSpeech shape)"
UNIT = "mel-frames/s"
FFN_FLOP_PER_FRAME = 2 * 256 * 9 * 1024  # the dominant kernel: Conv1d(256->1024, k=9) of one decoder FFN


def _port_over_reference():
    """Time of the oracle port / time of the unmodified reference on this workload, measured in the build container where
    the reference is mounted (profiles/r02_port_vs_reference.json; tools/port_vs_reference.py)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_port_vs_reference.json")))["port_over_reference"]
    except Exception:
        return None


PORT_OVER_REFERENCE = _port_over_reference()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(hbm=6650.0, tensor=1590.0, src="fallback (B200_PROFILING.md)")


# Workloads.  "fs2" is the headline (BASELINE.json configs[1]); the others are the same batch-16 LJSpeech shape through
# the other block types, and BASELINE configs[4] (fs2 + liu2021 prosody, every utterance S = M / 8 phonemes long).
WORKLOADS = {
    "fs2": dict(block="transformer_fs2", prosody=None, s_max=100, s_step=2,
                text="transformer_fs2 + supervised duration (learn_alignment False), LJSpeech shape, batch 16 per GPU, "
                     "S 100..70, 8 frames/phoneme (M 800, 10880 valid frames), free-running inference, random-init weights"),
    "transformer": dict(block="transformer", prosody=None, s_max=100, s_step=2),
    "fastformer": dict(block="fastformer", prosody=None, s_max=100, s_step=2),
    "conformer": dict(block="conformer", prosody=None, s_max=100, s_step=2),
    "liu2021_m64": dict(block="transformer_fs2", prosody="liu2021", s_max=8, s_step=0),
    "liu2021_m256": dict(block="transformer_fs2", prosody="liu2021", s_max=32, s_step=0),
    "liu2021_m1024": dict(block="transformer_fs2", prosody="liu2021", s_max=128, s_step=0),
}


def workload_text(name):
    w = WORKLOADS[name]
    if "text" in w:
        return w["text"]
    return "%s%s, supervised duration, LJSpeech config, batch 16 per GPU, S %d%s, 8 frames/phoneme (M %d), free-running " \
           "inference, random-init weights" % (w["block"], " + liu2021 prosody" if w["prosody"] else "", w["s_max"],
                                               "..%d" % (w["s_max"] - 15 * w["s_step"]) if w["s_step"] else " (all utterances)",
                                               8 * w["s_max"])


def build_workload(seed, name="fs2"):
    from ctts_b200 import configs, spec, synth
    w = WORKLOADS[name]
    p, m, t = configs.builtin_configs("LJSpeech", block_type=w["block"], learn_alignment=False, prosody=w["prosody"])
    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0], pin_frames_per_phoneme=FRAMES_PER_PHONEME)
    batch = synth.ljspeech_batch(batch=BATCH, s_max=w["s_max"], s_step=w["s_step"], mode="infer", seed=seed)
    frames = int(batch["src_lens"].sum()) * FRAMES_PER_PHONEME
    return (p, m, t), sd, batch, frames


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def pick_cpu_threads(cfgs, sd):
    """The reference gets the thread count that serves it best: a 4-utterance slice of the workload is timed at
    8 / 16 / 32 / 64 / all logical cores (oversubscribed hosts are much slower at 'all')."""
    from oracle import ctts_oracle as O
    from ctts_b200 import synth
    p, m, t = cfgs
    small = synth.ljspeech_batch(batch=4, s_max=100, s_step=2, mode="infer", seed=0)
    a = (small["speakers"], small["texts"], small["src_lens"], small["max_src_len"])
    ncpu = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for n in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu} or {ncpu}):
        torch.set_num_threads(n)
        with torch.no_grad():
            O.comp_trans_tts_forward(sd, p, m, t, *a)
            t0 = time.perf_counter()
            O.comp_trans_tts_forward(sd, p, m, t, *a)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def cpu_forward_timer(cfgs, sd, batch, frames, budget_s=20.0, min_iters=2):
    """Times oracle.comp_trans_tts_forward on the host; returns (frames/s, iters, threads, median s)."""
    from oracle import ctts_oracle as O
    p, m, t = cfgs
    threads = pick_cpu_threads(cfgs, sd)
    args = (batch["speakers"], batch["texts"], batch["src_lens"], batch["max_src_len"])
    with torch.no_grad():
        O.comp_trans_tts_forward(sd, p, m, t, *args)  # warm-up
        times = []
        t_end = time.perf_counter() + budget_s
        while len(times) < min_iters or (time.perf_counter() < t_end and len(times) < 8):
            t0 = time.perf_counter()
            O.comp_trans_tts_forward(sd, p, m, t, *args)
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return frames / med, len(times), threads, med


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfgs, sd, batch, frames = build_workload(seed=0)
    steps = max(args.steps, 1)
    from oracle import ctts_oracle as O
    p, m, t = cfgs
    threads = pick_cpu_threads(cfgs, sd)
    a = (batch["speakers"], batch["texts"], batch["src_lens"], batch["max_src_len"])
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            O.comp_trans_tts_forward(sd, p, m, t, *a)
        # the same number of steps as the GPU arm (0.5 - 2.5 s per step depending on the host: K = 20 ends within a minute)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.comp_trans_tts_forward(sd, p, m, t, *a)
        dt = time.perf_counter() - t0
    value = frames * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "transformer_fs2 + supervised duration, LJSpeech shape, batch 16, S 100..70, "
                               "8 frames/phoneme (M 800, 10880 valid frames), free-running inference",
                   "device": "host CPU, torch %s" % torch.__version__},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d full forward passes of the batch-16 workload (oracle/ctts_oracle.py); thread count = "
                                   "best of 8/16/32/64/all on a 4-utterance slice; host has %d logical cores"
                                   % (steps, os.cpu_count() or 1),
                         "port_over_reference_time": PORT_OVER_REFERENCE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
# Training step (BASELINE.json configs[2], [3]; train.py:104-123): forward in model.train() + objective + backward + the
# data-parallel gradient all-reduce.  Reported as a sub-record of the headline line.
TRAIN_CONFIGS = [
    # name, dataset, block, learn_alignment, mode, global batch, s_max, s_step
    ("transformer_fs2 LJSpeech-shape B16 supervised duration (configs[1] shape, training step)", "LJSpeech",
     "transformer_fs2", False, "teacher", 16, 100, 2),
    ("fastformer VCTK-shape B32 multi-speaker, unsupervised alignment (configs[3])", "VCTK", "fastformer", True, "unsup",
     32, 50, 1),
    ("conformer LJSpeech-shape B16 unsupervised alignment (configs[2])", "LJSpeech", "conformer", True, "unsup", 16, 100, 2),
]


def _train_objective(out):
    """A scalar on every differentiable output (stands in for model/loss.py: plain torch on the 14-tuple)."""
    terms = [out[0], out[1], out[4]]
    if out[2] is not None:
        terms += [out[2]["cwt"], out[2]["f0_mean"], out[2]["f0_std"]]
    if out[3] is not None:
        terms.append(out[3])
    if out[10][0] is not None:
        terms += [out[10][0], out[10][3]]
    return sum(t.float().pow(2).mean() for t in terms)


def train_record(rank, world, dev, steps=5, warmup=2, only=None):
    import torch.distributed as dist
    import ctts_b200
    from ctts_b200 import capi, configs, spec, synth
    from ctts_b200 import dist as cdist
    records = []
    for idx, (name, dataset, block, learn, mode, gbatch, s_max, s_step) in enumerate(TRAIN_CONFIGS):
        if only is not None and idx not in only:
            continue
        if gbatch % world:
            continue
        per = gbatch // world
        p, m, t = configs.builtin_configs(dataset, block_type=block, learn_alignment=learn)
        sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0], pin_frames_per_phoneme=None)
        net = ctts_b200.CompTransTTS(p, m, t)
        net.load_state_dict(sd, strict=True)
        net.to(dev).train()
        model = cdist.DistributedDataParallel(net) if world > 1 else net
        # the global batch is synthesised identically on every rank and each rank keeps its contiguous slice (train.py:237)
        full = synth.ljspeech_batch(batch=gbatch, s_max=s_max, s_step=s_step, mode=mode, seed=7,
                                    spk_dim=512 if dataset == "VCTK" else None)
        lo, hi = rank * per, (rank + 1) * per

        def cut(v):
            if torch.is_tensor(v):
                return v[lo:hi].contiguous().to(dev) if v.dim() > 0 and v.shape[0] == gbatch else v.to(dev)
            if isinstance(v, dict):
                return {k: cut(x) for k, x in v.items()}
            return v

        b = {k: cut(v) for k, v in full.items()}
        S = int(b["src_lens"].max())
        M = int(b["mel_lens"].max())
        b["texts"] = b["texts"][:, :S].contiguous()
        b["max_src_len"], b["max_mel_len"] = S, M
        b["mels"] = b["mels"][:, :M].contiguous()
        for k in list(b["p_targets"]):
            v = b["p_targets"][k]
            if v.dim() >= 2:
                b["p_targets"][k] = v[:, :M].contiguous()
        if mode == "teacher":
            b["e_targets"] = b["e_targets"][:, :S].contiguous()
            b["d_targets"] = b["d_targets"][:, :S].contiguous()
        else:
            b["e_targets"] = b["e_targets"][:, :M].contiguous()
            b["attn_priors"] = b["attn_priors"][:, :S, :M].contiguous()
        frames_global = int(full["mel_lens"].sum())
        kw = {k: b[k] for k in ("mels", "mel_lens", "max_mel_len", "p_targets", "e_targets", "d_targets", "attn_priors",
                                "spker_embeds") if k in b}
        if mode == "unsup":
            kw["step"] = 120000

        def one_step():
            import copy
            k2 = dict(kw)
            k2["p_targets"] = dict(kw["p_targets"])
            net.zero_grad(set_to_none=True)
            out = model(b["speakers"], b["texts"], b["src_lens"], b["max_src_len"], **k2)
            _train_objective(out).backward()

        def timed(n_steps, n_warm):
            nonlocal_step = lambda: one_step()      # late binding: the weak-scaling variant swaps the step function
            for _ in range(n_warm):
                nonlocal_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            l0 = capi.LAUNCHES
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_steps):
                nonlocal_step()
            e1.record()
            e1.synchronize()
            if world > 1:
                dist.barrier()
            return cdist.max_over_ranks(e0.elapsed_time(e1), dev, world) / n_steps, (capi.LAUNCHES - l0) // n_steps

        ms, launches = timed(steps, warmup)
        rec = {"config": name, "global_batch": gbatch, "per_gpu_batch": per, "scaling": "strong (global batch fixed, "
               "train.py:237)", "S_max": S, "M_max": M, "valid_mel_frames_global": frames_global, "steps": steps,
               "ms_per_step": ms, "mel_frames_per_s": frames_global / (ms * 1e-3), "kernel_launches_per_step": launches,
               "includes": "training-mode forward, objective, backward, gradient all-reduce (N > 1); no optimizer step",
               "dropout": os.environ.get("CTTS_DROPOUT", "1") != "0"}
        arena = net.grad_arena()
        rec["grad_arena_mb"] = arena.flat.numel() * 4 / 1e6
        if world > 1:
            net._reducer.enabled = False
            ms_off, _ = timed(steps, 3)      # a new CUDA-graph key: eager, capture, then replays
            net._reducer.enabled = True
            # the all-reduce alone: the whole arena in one call, on an otherwise idle GPU
            torch.cuda.synchronize()
            dist.barrier()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dist.all_reduce(arena.flat)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ar = cdist.max_over_ranks(sorted(ts)[len(ts) // 2], dev, world)
            nbytes = arena.flat.numel() * 4
            ms2, _ = timed(steps, 3)      # and once more with the reducer: the order of the two measurements must not matter
            ms = min(ms, ms2)
            rec["ms_per_step"] = ms
            rec["mel_frames_per_s"] = frames_global / (ms * 1e-3)
            n_buckets = len(net._reducer.launched)
            rec.update({"ms_per_step_without_allreduce": ms_off, "allreduce_exposed_ms": ms - ms_off,
                        "allreduce_standalone_ms": ar, "allreduce_bus_gbs": 2 * (world - 1) / world * nbytes / (ar * 1e-3) / 1e9,
                        "allreduce_buckets": n_buckets,
                        "nvlink_peak_gbs_per_direction": 900.0})
        if world > 1:
            # weak scaling: every rank keeps the WHOLE N = 1 batch (global batch = world x the quoted one)
            bw = {k: (v.to(dev) if torch.is_tensor(v) else ({kk: vv.to(dev) for kk, vv in v.items()} if isinstance(v, dict)
                                                              else v)) for k, v in full.items()}
            kww = {k: bw[k] for k in ("mels", "mel_lens", "max_mel_len", "p_targets", "e_targets", "d_targets", "attn_priors",
                                      "spker_embeds") if k in bw}
            if mode == "unsup":
                kww["step"] = 120000

            def weak_step():
                k2 = dict(kww)
                k2["p_targets"] = dict(kww["p_targets"])
                net.zero_grad(set_to_none=True)
                out = model(bw["speakers"], bw["texts"], bw["src_lens"], bw["max_src_len"], **k2)
                _train_objective(out).backward()

            one_step_saved = one_step
            one_step = weak_step
            ms_w, _ = timed(steps, 3)
            one_step = one_step_saved
            rec["weak"] = {"per_gpu_batch": gbatch, "global_batch": gbatch * world, "ms_per_step": ms_w,
                           "mel_frames_per_s": world * frames_global / (ms_w * 1e-3)}
        records.append(rec)
        del net, model
        torch.cuda.empty_cache()
    return records


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step sub-record")
    ap.add_argument("--workload", default="fs2", choices=sorted(WORKLOADS), help="fs2 = the headline (BASELINE configs[1])")
    ap.add_argument("--train-only", action="store_true", help="print only the training-step record (development)")
    ap.add_argument("--train-config", type=int, default=None, help="restrict the training record to one TRAIN_CONFIGS entry")
    ap.add_argument("--train-steps", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    import ctts_b200
    from ctts_b200 import capi, engine
    from ctts_b200 import dist as cdist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    if args.train_only:
        recs = train_record(rank, world, dev, steps=args.train_steps, warmup=2 if args.train_steps > 1 else 1,
                            only=None if args.train_config is None else [args.train_config])
        if rank == 0:
            print(json.dumps({"train": recs, "n_gpus": world}))
        if world > 1:
            dist.destroy_process_group()
        return

    cfgs, sd, batch, frames = build_workload(seed=rank, name=args.workload)
    m_max = FRAMES_PER_PHONEME * WORKLOADS[args.workload]["s_max"]
    conformer = WORKLOADS[args.workload]["block"] == "conformer"
    flop_per_frame = 2 * 256 * 1024 if conformer else FFN_FLOP_PER_FRAME
    if args.workload != "fs2":
        args.no_train = True
    net = ctts_b200.CompTransTTS(*cfgs).eval()
    net.load_state_dict(sd, strict=True)
    net.to(dev)
    dev_in = (batch["speakers"].to(dev), batch["texts"].to(dev), batch["src_lens"].to(dev), batch["max_src_len"])
    host_in = tuple(x.pin_memory() for x in (batch["speakers"], batch["texts"], batch["src_lens"]))
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)  # > 126 MB L2

    # --- per-launch timing of the dominant kernel (decoder FFN Conv1d k=9), live, on the launching stream ------------
    ffn_events = []
    orig_conv, orig_tc = engine.conv_gemm, engine.gemm_tc

    def timed(fn):
        def wrapper(x, w, *a, **k):
            # the decoder's widest GEMM: the k = 9 FFN conv (conformer: the first FFN linear 256 -> 1024)
            hit = (k.get("taps", 1) == 9) if not conformer else (k.get("taps", 1) == 1 and w.shape[0] == 1024)
            if hit and x.shape[1] == m_max:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = fn(x, w, *a, **k)
                e1.record()
                ffn_events.append((e0, e1))
                return y
            return fn(x, w, *a, **k)
        return wrapper

    ffn_kernel = ("ctts_gemm_split -> gemm_pair_kernel (tcgen05 cta_group::2, TMA, TMEM; bf16 hi/lo planes x3 MMAs)"
                  if net.decoder_math == "bf16x3"
                  else "ctts_conv1d_gemm (FP32 CUDA cores)")

    def step_device():
        return net(*dev_in)

    host_out = {}

    def step_e2e():
        spk, txt, lens = (h.to(dev, non_blocking=True) for h in host_in)
        out = net(spk, txt, lens, batch["max_src_len"])
        # device -> host read of the result (post-net mel + its lengths) into pinned landing buffers, as a serving loop would
        for k, t in (("mel", out[1]), ("mel_lens", out[9])):
            if k not in host_out or host_out[k].shape != t.shape:
                host_out[k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            host_out[k].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out["mel"], host_out["mel_lens"]

    def run_timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = 0.0
        launches0 = capi.LAUNCHES
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return cdist.max_over_ranks(total, dev, world), capi.LAUNCHES - launches0

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches = run_timed(step_device, args.steps, args.warmup)
    ms_e2e, _ = run_timed(step_e2e, args.steps, 2)
    # per-launch duration of the dominant kernel: the same steps again with graph replay off (a replayed graph cannot
    # be instrumented from the host), CUDA events around each launch on the launching stream, L2 flushed between steps
    graphs_were_on = net.use_cuda_graphs
    net.use_cuda_graphs = False
    engine.conv_gemm, engine.gemm_tc = timed(orig_conv), timed(orig_tc)
    from ctts_b200 import engine_blocks      # the other block types bind the two functions at import time
    engine_blocks.conv_gemm, engine_blocks.gemm_tc = engine.conv_gemm, engine.gemm_tc
    ffn_events.clear()
    # Without graphs the host is slower than the GPU, and an event pair would then include the wait for the launch.
    # A ~4 ms device-side spin at the start of each stage lets the host queue the whole stage ahead of the GPU, so the
    # pairs bracket kernel execution only.
    orig_stage_b = engine.variance_stage_b

    def spin_then_stage_b(*a, **k):
        torch.cuda._sleep(8_000_000)
        return orig_stage_b(*a, **k)

    engine.variance_stage_b = spin_then_stage_b

    def step_instrumented():
        torch.cuda._sleep(8_000_000)
        return step_device()

    _, eager_launches = run_timed(step_instrumented, args.steps, 1)
    engine.variance_stage_b = orig_stage_b
    torch.cuda.synchronize()
    per_step = len(ffn_events) // max(args.steps + 1, 1)
    ffn_ms = [a.elapsed_time(b) for a, b in ffn_events[-max(per_step, 1) * args.steps:]]
    engine.conv_gemm, engine.gemm_tc = orig_conv, orig_tc
    engine_blocks.conv_gemm, engine_blocks.gemm_tc = orig_conv, orig_tc
    net.use_cuda_graphs = graphs_were_on
    if graphs_were_on:
        launches = eager_launches   # kernels per step are the same; replayed steps do not pass through capi.call
    clocks = sampler.stop() if sampler else None
    train = None
    if not args.no_train:
        del flush
        torch.cuda.empty_cache()
        try:
            train = train_record(rank, world, dev)
        except Exception as e:      # the headline line must not be lost to a failure of the sub-record
            train = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    if rank == 0:
        pk = peaks()
        value = world * frames * args.steps / (ms_total / 1e3)
        e2e = world * frames * args.steps / (ms_e2e / 1e3)
        ffn_mean = sum(ffn_ms) / max(len(ffn_ms), 1)
        # median: in the eager re-run the host occasionally falls behind the GPU (descriptor encoding, allocator), and a
        # start event that completes on an idle GPU then also counts the host's delay before the launch
        ffn_avg = sorted(ffn_ms)[len(ffn_ms) // 2] if ffn_ms else 0.0
        traffic = tensor_pct = None
        summ = os.path.join(ROOT, "profiles", "ncu_ffn1_summary.json")
        if os.path.exists(summ) and net.decoder_math == "bf16x3" and args.workload == "fs2":   # one ncu --set full capture of this kernel (committed)
            nc = json.load(open(summ))
            traffic, tensor_pct = nc["dram_bytes_total"], nc["tensor_pipe_active_pct_of_peak_sustained_active"]
        ffn_tflops = frames * flop_per_frame / (ffn_avg * 1e-3) / 1e12 if ffn_ms else None
        burst = None
        try:
            burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (decoder/PostNet GEMMs: bf16 hi+lo planes x3 MMAs, fp32 accumulate)" if net.decoder_math == "bf16x3"
            else "f32", "data": "synthetic",
            "config": {"workload": workload_text(args.workload), "workload_name": args.workload,
                       "l2_flush": "256 MiB device write between timed steps", "timing": "CUDA events per step, max over ranks",
                       "parallelism": "independent shards x%d" % world,
                       "cuda_graphs": bool(net.use_cuda_graphs)},
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host_in)),
                    "d2h_bytes_per_step": int(BATCH * m_max * 80 * 4 + BATCH * 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "decoder FFN Conv1d(256->1024,k9)+GELU implicit GEMM: " + ffn_kernel,
                         "bound": "tensor", "achieved": ffn_tflops, "peak": pk["tensor"], "unit": "TFLOP/s",
                         "frac": (ffn_tflops / pk["tensor"]) if ffn_tflops else None,
                         "frac_of_burst_peak": (ffn_tflops / burst) if (ffn_tflops and burst) else None, "burst_peak": burst,
                         "traffic": traffic,
                         "traffic_source": "profiles/ncu_ffn1_summary.json (dram read + write bytes per launch)",
                         "tensor_pipe_active_pct": tensor_pct, "mma_per_algorithmic_product": 3,
                         "peak_source": pk["src"], "launch_ms": ffn_avg, "launch_ms_mean": ffn_mean, "launches_timed": len(ffn_ms),
                         "launch_timing": "median over CUDA-event pairs around each launch in an eager re-run of the timed steps "
                                          "(host kept ahead of the GPU by a device-side spin at the start of each stage)",
                         "flops_per_launch": frames * flop_per_frame},
        }
        if train is not None:
            line["train"] = train
        if world == 1 and not args.no_cpu_baseline:
            v, iters, threads, med = cpu_forward_timer(cfgs, sd, batch, frames, budget_s=20.0 if args.workload == "fs2" else 8.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "port_over_reference_time": PORT_OVER_REFERENCE,
                                    "sample": "%d full forward passes of the same batch-16 workload on the host "
                                              "(oracle/ctts_oracle.py, torch fp32, median %.0f ms; thread count = best of "
                                              "8/16/32/64/all on a 4-utterance slice; host has %d logical cores)"
                                              % (iters, med * 1e3, os.cpu_count() or 1)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
