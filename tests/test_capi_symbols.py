"""CPU: the C-ABI shared library loads and exports every symbol include/ctts_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ctts_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctts_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for need in ("ctts_conv1d_gemm", "ctts_attention", "ctts_length_scan", "ctts_length_expand", "ctts_layernorm",
                 "ctts_gemm_bf16x3", "ctts_cwt_to_pitch", "ctts_embed_tokens"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from ctts_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "libctts_b200.so does not export %s" % s
    assert sorted(capi.SIGNATURES) == declared_symbols(), "capi.SIGNATURES and the header disagree"
    assert capi.load().ctts_abi_version() == 1


def test_no_cpu_fallback():
    """The product path must refuse CPU tensors instead of silently computing somewhere else."""
    import torch
    import ctts_b200
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=False)
    m["transformer_fs2"]["encoder_layer"] = 1
    m["transformer_fs2"]["decoder_layer"] = 1
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    with pytest.raises(Exception):
        net(torch.zeros(1, dtype=torch.long), torch.ones(1, 4, dtype=torch.long), torch.tensor([4]), 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "comprehensive-transformer-tts_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                for line in txt.splitlines():
                    assert not re.match(r"\s*(from|import)\s+oracle", line), "%s imports the oracle" % f


def _prototypes():
    """name -> list of parameter type strings, parsed from the header (comments stripped)."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*)\s*(ctts_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        if params in ("", "void"):
            out[m.group(1)] = []
            continue
        out[m.group(1)] = [" ".join(p.split()) for p in params.split(",")]
    return out


def test_ctypes_signatures_match_the_header():
    """Every parameter of every prototype in include/ctts_b200.h has the matching ctypes kind in capi.SIGNATURES:
    pointer -> c_void_p, float -> c_float, long long -> c_longlong, size_t -> c_size_t, int -> c_int.  Catches ABI drift without a GPU."""
    from ctts_b200 import capi
    protos = _prototypes()
    assert sorted(protos) == sorted(capi.SIGNATURES)
    for name, params in protos.items():
        sig = capi.SIGNATURES[name]
        assert len(sig) == len(params), "%s: header has %d parameters, binding %d" % (name, len(params), len(sig))
        for i, (p, c) in enumerate(zip(params, sig)):
            if "*" in p:
                want = ctypes.c_void_p
            elif re.match(r"(const )?float\b", p):
                want = ctypes.c_float
            elif re.match(r"(const )?unsigned long long\b", p):
                want = ctypes.c_ulonglong
            elif re.match(r"(const )?long long\b", p):
                want = ctypes.c_longlong
            elif re.match(r"(const )?size_t\b", p):
                want = ctypes.c_size_t
            else:
                assert re.match(r"(const )?int\b", p), "%s: unexpected parameter type %r" % (name, p)
                want = ctypes.c_int
            assert c is want, "%s parameter %d (%s): binding uses %s" % (name, i, p, c.__name__)


def test_emulator_restates_every_kernel_entry_point():
    """oracle/capi_emulator.py is the per-kernel oracle of the GPU tests and the stand-in of the CPU host-logic tests: every
    kernel-launching entry point of the ABI has a torch restatement there with the same number of arguments."""
    import inspect
    from ctts_b200 import capi
    from oracle import capi_emulator as emu
    not_kernels = {"ctts_abi_version", "ctts_last_error", "ctts_device_arch", "ctts_debug_set_timing_buffer"}
    # entry points that only exist as thin aliases / legacy forms of a restated one
    missing = []
    for name, argtypes in capi.SIGNATURES.items():
        if name in not_kernels:
            continue
        fn = getattr(emu, name, None)
        if fn is None:
            missing.append(name)
            continue
        n_args = len(inspect.signature(fn).parameters)
        assert n_args == len(argtypes), "%s: emulator takes %d arguments, the ABI %d" % (name, n_args, len(argtypes))
    assert not missing, "no restatement in oracle/capi_emulator.py for: %s" % ", ".join(sorted(missing))
