"""Shared harness of the per-kernel GPU tests: call one C-ABI entry point on the device and its CPU restatement
(oracle/capi_emulator.py) on copies of the same arguments, then compare every tensor argument."""
import ctypes

import torch

from ctts_b200 import capi
from oracle import capi_emulator as emu

DEV = torch.device("cuda:0")


class PA:
    """A `const void* const*` plane-array argument (list of bf16 tensors)."""
    def __init__(self, tensors):
        self.t = list(tensors)


class Scratch:
    """A workspace argument: allocated on both sides, never compared."""
    def __init__(self, t):
        self.t = t


class LL:
    def __init__(self, *v):
        self.v = [int(x) for x in v]


class II:
    def __init__(self, *v):
        self.v = [int(x) for x in v]


def _conv(a, dev):
    if torch.is_tensor(a):
        return a.clone().to(dev)
    if isinstance(a, PA):
        ts = [t.clone().to(dev) for t in a.t]
        arr = capi.ptr_array(ts)
        return arr
    if isinstance(a, Scratch):
        return _NoCompare(a.t.clone().to(dev))
    if isinstance(a, LL):
        return (ctypes.c_longlong * len(a.v))(*a.v)
    if isinstance(a, II):
        return (ctypes.c_int * len(a.v))(*a.v)
    return a


class _NoCompare:
    def __init__(self, t):
        self.t = t


def run_both(name, args, atol=1e-5, rtol=1e-4, int_exact=True):
    cpu = [_conv(a, torch.device("cpu")) for a in args]
    gpu = [_conv(a, DEV) for a in args]
    unwrap = lambda c: c.t if isinstance(c, _NoCompare) else c
    emu.ENTRIES[name](*[(unwrap(c).detach() if torch.is_tensor(unwrap(c)) else c) for c in cpu], 0)
    capi.call(name, *[unwrap(x) for x in gpu], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    worst = 0.0
    for i, (c, g) in enumerate(zip(cpu, gpu)):
        pairs = []
        if torch.is_tensor(c):
            pairs = [(c, g)]
        elif hasattr(c, "_keepalive"):     # operand planes: the VALUE (sum of the planes) is the contract
            pairs = [(sum(t.float() for t in c._keepalive), sum(t.float() for t in g._keepalive))]
        for ct, gt in pairs:
            gt = gt.cpu()
            if ct.dtype in (torch.int64, torch.int32, torch.uint8):
                if int_exact:
                    assert torch.equal(ct, gt), "%s: integer argument %d differs" % (name, i)
                continue
            a, b = ct.float(), gt.float()
            assert torch.isfinite(b).all(), "%s: argument %d has non-finite values on the GPU" % (name, i)
            err = (a - b).abs()
            lim = atol + rtol * a.abs()
            bad = err > lim
            assert not bad.any(), "%s: argument %d differs: max err %.3g at |ref| %.3g (%d of %d elements)" % (
                name, i, err.max().item(), a.abs().flatten()[err.argmax()].item(), int(bad.sum()), a.numel())
            worst = max(worst, err.max().item())
    return worst


def g(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed + 17 * len(shape) + sum(shape))
    return torch.randn(*shape, generator=gen) * scale


def planes(x, n=2):
    out, rem = [], x.float().clone()
    for _ in range(n):
        h = rem.to(torch.bfloat16)
        out.append(h)
        rem = rem - h.float()
    return out


