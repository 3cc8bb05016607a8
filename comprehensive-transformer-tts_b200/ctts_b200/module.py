"""`CompTransTTS`: the reference's Python API surface over the B200 kernels.

Same constructor, same `forward` signature and 14-tuple, same state_dict key names as
model/CompTransTTS.py:12-152, so `utils/model.py:get_model`, `train.py`, `evaluate.py` and
`synthesize.py` of the reference can use this class in place of theirs.
The module tree only exists to carry parameters under the reference's names; the forward pass
is `engine.forward`, which calls libctts_b200 through its C ABI.
"""
import math
import os

import torch
import torch.nn as nn

from . import engine, spec


class _Tracked(nn.Module):
    """Counts structural changes of the parameter tree (a tensor / sub-module attached, replaced or removed anywhere) so
    that the engine can keep its name -> tensor table between calls: walking the tree costs ~0.25 ms per forward, an
    order of magnitude more than checking the storages and version counters of the tensors it already knows."""
    structure_epoch = 0

    def __setattr__(self, name, value):
        if isinstance(value, (torch.Tensor, nn.Module)) or name in self.__dict__.get("_parameters", ()) \
                or name in self.__dict__.get("_buffers", ()):
            _Tracked.structure_epoch += 1
        super().__setattr__(name, value)

    def __delattr__(self, name):
        _Tracked.structure_epoch += 1
        super().__delattr__(name)

    def register_parameter(self, name, param):
        _Tracked.structure_epoch += 1
        super().register_parameter(name, param)

    def register_buffer(self, name, tensor, persistent=True):
        _Tracked.structure_epoch += 1
        super().register_buffer(name, tensor, persistent=persistent)

    def add_module(self, name, module):
        _Tracked.structure_epoch += 1
        super().add_module(name, module)

    def load_state_dict(self, *args, **kwargs):
        _Tracked.structure_epoch += 1      # assign=True swaps the Parameter objects
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, recurse=True):
        _Tracked.structure_epoch += 1      # .to() / .float() may swap the Parameter objects (future overwrite flag)
        return super()._apply(fn, recurse)


class _Node(_Tracked):
    """Anonymous container: the parameter names are the contract, not the class tree."""


def _attach(root, name, tensor, kind):
    parts = name.split(".")
    node = root
    for p in parts[:-1]:
        nxt = node._modules.get(p)
        if nxt is None:
            nxt = _Node()
            node.add_module(p, nxt)
        node = nxt
    if kind == "buffer":
        node.register_buffer(parts[-1], tensor)
        return tensor
    param = tensor if isinstance(tensor, nn.Parameter) else nn.Parameter(tensor, requires_grad=(kind == "param"))
    node.register_parameter(parts[-1], param)
    return param


def _initial(shape, init):
    """Initialisers equivalent in distribution to the reference's (blocks.py:10-23,255-298 and torch defaults)."""
    if init == "count":
        return torch.zeros((), dtype=torch.long)
    if init == "ones":
        return torch.ones(shape)
    if init == "zeros":
        return torch.zeros(shape)
    if init == "normal01":
        return torch.randn(shape)
    if init == "normal05":  # STL token embeddings, modules.py:470
        return torch.randn(shape) * 0.5
    if init == "normal002":  # fastformer.py:285-290
        return torch.randn(shape) * 0.02
    if init == "emb1":
        t = torch.randn(shape)
        t[0] = 0
        return t
    if init == "sinusoid_interleaved":  # get_sinusoid_encoding_table, blocks.py:26-46
        import numpy as np
        _, rows, d = shape
        pos = np.arange(rows, dtype=np.float64)[:, None]
        tab = pos / np.power(10000, 2 * (np.arange(d)[None, :] // 2) / d)
        tab[:, 0::2] = np.sin(tab[:, 0::2])
        tab[:, 1::2] = np.cos(tab[:, 1::2])
        return torch.FloatTensor(tab).unsqueeze(0)
    if init.startswith("linspace") or init.startswith("logspace"):
        _, lo, hi = init.split(":")
        if init.startswith("logspace"):
            return torch.exp(torch.linspace(math.log(float(lo)), math.log(float(hi)), shape[0]))
        return torch.linspace(float(lo), float(hi), shape[0])
    if init.startswith("emb"):
        d = int(init.split(":")[1])
        t = torch.randn(shape) * d ** -0.5
        t[0] = 0
        return t
    if init.startswith("default"):
        bound = 1.0 / math.sqrt(int(init.split(":")[1]))
        return (torch.rand(shape) * 2 - 1) * bound
    if init.startswith("xavier"):
        gain = spec.gain_of(init.split(":")[1] if ":" in init else "")
        t = torch.empty(shape)
        nn.init.xavier_uniform_(t, gain=gain)
        return t
    raise ValueError(init)


class CompTransTTS(_Tracked):
    """ CompTransTTS (B200-native).  Reference: model/CompTransTTS.py:12-152. """

    def __init__(self, preprocess_config, model_config, train_config):
        super().__init__()
        self.preprocess_config = preprocess_config
        self.model_config = model_config
        self.train_config = train_config
        entries, d_enc, d_dec = spec.parameter_spec(preprocess_config, model_config)
        made = {}
        for name, shape, kind, init in entries:
            if init.startswith("tie:"):
                _attach(self, name, made[init[4:]], kind)
            else:
                made[name] = _attach(self, name, _initial(shape, init), kind)
        self.d_encoder, self.d_decoder = d_enc, d_dec
        self.has_speaker_emb = bool(model_config["multi_speaker"])
        self.embedder_type = preprocess_config["preprocessing"].get("speaker_embedder", "none") \
            if self.has_speaker_emb else None
        # arithmetic of the decoder / mel head (downstream of every quantiser): "bf16x3" = tcgen05 tensor cores with
        # bf16 hi/lo operand planes, "fp32" = CUDA-core FMA (same kernels as the encoder / predictors)
        self.decoder_math = os.environ.get("CTTS_DECODER_MATH", "bf16x3")
        if self.decoder_math not in ("bf16x3", "fp32"):
            raise ValueError("CTTS_DECODER_MATH must be 'bf16x3' or 'fp32'")
        # arithmetic of the encoder + variance predictors (UPSTREAM of the quantisers, must be FP32-equivalent):
        # "bf16x6" = tcgen05 with 3 bf16 planes per operand (24 mantissa bits, 6 MMAs per k-slice), "fp32" = CUDA cores
        self.encoder_math = os.environ.get("CTTS_ENCODER_MATH", "bf16x6")
        if self.encoder_math not in ("bf16x6", "fp32"):
            raise ValueError("CTTS_ENCODER_MATH must be 'bf16x6' or 'fp32'")
        # CUDA graphs: the forward is captured as two stages around its single host sync (engine.GraphCache)
        self.use_cuda_graphs = os.environ.get("CTTS_CUDA_GRAPHS", "1") != "0"
        self._graphs = engine.GraphCache()
        self._prepared = engine.Prepared(self)
        # training step (train_engine.py): flat gradient arena, dropout stream, optional data-parallel reducer
        self._arena = None
        self._arena_epoch = -1
        self._train_weights = None
        self._anchor = None
        self._reducer = None
        self._dropout_seed = int(os.environ.get("CTTS_DROPOUT_SEED", train_config.get("seed", 1234)
                                                if isinstance(train_config, dict) else 1234))
        self._dropout_offset = 0
        self._dropout_counter = None
        from . import train_engine
        self._train_graphs = train_engine.TrainGraphs()
        self._train_graphs.sig = None

    def dropout_counter(self, device):
        """Device-side step counter of the dropout stream (added to every site's offset; CUDA-graph mode)."""
        if self._dropout_counter is None or self._dropout_counter.device != device:
            self._dropout_counter = torch.zeros(1, device=device, dtype=torch.int64)
        return self._dropout_counter

    def grad_arena(self):
        """The flat fp32 buffer all parameter gradients live in (param.grad are views of it); rebuilt when the parameters
        move (`.to(device)`) or are replaced."""
        from . import train_engine
        if self._arena is None or self._arena_epoch != _Tracked.structure_epoch:
            # (the tree walk behind signature() costs ~0.25 ms: only after a structural change, not every step)
            if self._arena is None or self._arena.sig != train_engine.GradArena.signature(self):
                self._arena = train_engine.GradArena(self)
                self._train_weights = train_engine.TrainWeights()
            self._arena_epoch = _Tracked.structure_epoch
        return self._arena

    def autograd_anchor(self, device):
        if self._anchor is None or self._anchor.device != device:
            self._anchor = torch.zeros(1, device=device, requires_grad=True)
        return self._anchor

    def forward(self, speakers, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
                p_targets=None, e_targets=None, d_targets=None, attn_priors=None, spker_embeds=None, p_control=1.0,
                e_control=1.0, d_control=1.0, step=None):
        if self.training:
            from . import train_engine
            return train_engine.forward(self, speakers, texts, src_lens, max_src_len, mels, mel_lens, max_mel_len,
                                        p_targets, e_targets, d_targets, attn_priors, spker_embeds, p_control, e_control,
                                        d_control, step)
        with torch.no_grad():
            return engine.forward(self, speakers, texts, src_lens, max_src_len, mels, mel_lens, max_mel_len,
                                  p_targets, e_targets, d_targets, attn_priors, spker_embeds, p_control, e_control,
                                  d_control, step)
