"""CPU oracle for the RE-SepFormer overlap-separation path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``clearconverse_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED.  The arithmetic of the path lives in the third-party dependency
``speechbrain`` (``/root/reference/back/requirements.txt:16``, bare name, no pinned
version; the import path ``speechbrain.inference`` at ``/root/reference/back/api.py:39``
exists only in speechbrain >= 1.0.0) and in the HF hub repo
``speechbrain/resepformer-wsj02mix`` (``api.py:714``, no revision).  Neither is under
``/root/reference`` nor installable here (no network, no wheel), and the reference holds
no golden vector, known-answer test or fixture for this path.
This file therefore RESTATES the published upstream algorithm module by module, keeping
upstream's module tree and state-dict key names so that a real ``encoder.ckpt`` /
``masknet.ckpt`` / ``decoder.ckpt`` (``api.py:729``) loads unmodified the day one is
reachable.  What stands in for reference-pinned goldens:
  * parameter count 7,955,201 (the model card's "8 M"),
  * an independent hand-decomposed restatement (``oracle/functional_restatement.py``)
    that must agree with this module tree to float rounding,
  * an fp64 run of the same tree as the truth for error budgeting.

Upstream files restated (speechbrain >= 1.0):
  speechbrain/inference/separation.py      SepformerSeparation.separate_batch
  speechbrain/lobes/models/dual_path.py    Encoder, Decoder, GlobalLayerNorm, select_norm
  speechbrain/lobes/models/resepformer.py  SBTransformerBlock_wnormandskip,
                                           ResourceEfficientSeparationPipeline,
                                           ResourceEfficientSeparator
  speechbrain/lobes/models/transformer/Transformer.py
                                           TransformerEncoder, TransformerEncoderLayer,
                                           PositionalEncoding
  speechbrain/nnet/attention.py            MultiheadAttention, PositionalwiseFeedForward
  speechbrain/nnet/normalization.py        LayerNorm
Reference call sites the restatement is anchored on:
  construct  /root/reference/back/api.py:713-717
  weights    /root/reference/back/api.py:729-745
  call       /root/reference/back/api.py:1077
  consume    /root/reference/back/api.py:1080-1092
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

# hyperparams.yaml of speechbrain/resepformer-wsj02mix
N_ENC = 128          # encoder filters == d_model
KERNEL = 16          # encoder / decoder kernel size
STRIDE = 8           # kernel_size // 2
SEGMENT = 150        # masknet segment_size (chunk length K)
N_HEAD = 8
D_FFN = 1024
N_LAYERS = 8         # transformer layers per block
N_BLOCKS = 2         # masknet "layer"
NUM_SPKS = 2
SAMPLE_RATE = 8000
LN_EPS = 1e-6        # sb.nnet.normalization.LayerNorm(d_model, eps=1e-6)
PE_MAX_LEN = 100000  # PositionalEncoding(max_len=100000)
GLN_EPS = torch.finfo(torch.float32).eps   # resepformer.py: EPS = torch.finfo(torch.get_default_dtype()).eps
EXPECTED_PARAMS = 7_955_201


# ----------------------------------------------------------------------------- dual_path.py
class Encoder(nn.Module):
    """dual_path.Encoder: ``relu(conv1d(x.unsqueeze(1)))``; key ``conv1d.weight``."""

    def __init__(self, kernel_size=KERNEL, out_channels=N_ENC, in_channels=1):
        super().__init__()
        self.conv1d = nn.Conv1d(in_channels, out_channels, kernel_size,
                                stride=kernel_size // 2, groups=1, bias=False)
        self.in_channels = in_channels

    def forward(self, x):
        if self.in_channels == 1:
            x = torch.unsqueeze(x, dim=1)          # [B,1,T]
        return F.relu(self.conv1d(x))              # [B,N,L]


class Decoder(nn.ConvTranspose1d):
    """dual_path.Decoder: a ConvTranspose1d subclass; key ``weight``; squeeze quirks kept."""

    def forward(self, x):
        if x.dim() not in [2, 3]:
            raise RuntimeError("{} accept 3/4D tensor as input".format(self.__class__.__name__))
        x = super().forward(x if x.dim() == 3 else torch.unsqueeze(x, 1))
        if torch.squeeze(x).dim() == 1:
            x = torch.squeeze(x, dim=1)
        else:
            x = torch.squeeze(x)
        return x


class GlobalLayerNorm(nn.Module):
    """dual_path.GlobalLayerNorm for shape==3: statistics over (channel, time) per item."""

    def __init__(self, dim, shape, eps=1e-8, elementwise_affine=True):
        super().__init__()
        assert shape == 3
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim, 1))
        self.bias = nn.Parameter(torch.zeros(dim, 1))

    def forward(self, x):                           # x: [N, C, L]
        mean = torch.mean(x, (1, 2), keepdim=True)
        var = torch.mean((x - mean) ** 2, (1, 2), keepdim=True)
        return self.weight * (x - mean) / torch.sqrt(var + self.eps) + self.bias


# ------------------------------------------------------------------------ nnet/normalization
class LayerNorm(nn.Module):
    """sb.nnet.normalization.LayerNorm wrapper; keys ``norm.weight`` / ``norm.bias``."""

    def __init__(self, input_size, eps=LN_EPS):
        super().__init__()
        self.norm = nn.LayerNorm(input_size, eps=eps, elementwise_affine=True)

    def forward(self, x):
        return self.norm(x)


# --------------------------------------------------------------------------- nnet/attention
class MultiheadAttention(nn.Module):
    """sb.nnet.attention.MultiheadAttention: wraps ``nn.MultiheadAttention`` (key ``att.*``),
    batch-first in/out, asks for the head-averaged weights (need_weights=True)."""

    def __init__(self, nhead, d_model, dropout=0.0):
        super().__init__()
        self.att = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead, dropout=dropout, bias=True)

    def forward(self, query, key, value):
        query, key, value = (t.permute(1, 0, 2) for t in (query, key, value))
        output, w = self.att(query, key, value, attn_mask=None, key_padding_mask=None, need_weights=True)
        return output.permute(1, 0, 2), w


class PositionalwiseFeedForward(nn.Module):
    """keys ``ffn.0.*`` (Linear d->ffn) and ``ffn.3.*`` (Linear ffn->d)."""

    def __init__(self, d_ffn, input_size, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        self.ffn = nn.Sequential(nn.Linear(input_size, d_ffn), activation(), nn.Dropout(dropout),
                                 nn.Linear(d_ffn, input_size))

    def forward(self, x):
        return self.ffn(x.permute(1, 0, 2)).permute(1, 0, 2)


# ------------------------------------------------------------------ transformer/Transformer.py
class PositionalEncoding(nn.Module):
    """sin on even / cos on odd channels, base 10000, computed in fp32 exactly as upstream.
    Upstream registers ``pe`` [1,100000,D] as a persistent buffer (51 MB per block); here it
    is built lazily to the needed length with the same fp32 expression and kept out of the
    state dict (``pos_enc.pe`` keys of a real checkpoint are ignored on load)."""

    def __init__(self, input_size, max_len=PE_MAX_LEN):
        super().__init__()
        self.input_size, self.max_len = input_size, max_len

    @staticmethod
    def table(length, input_size, dtype=torch.float32):
        pe = torch.zeros(length, input_size)
        positions = torch.arange(0, length).unsqueeze(1).float()
        denominator = torch.exp(torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        pe[:, 0::2] = torch.sin(positions * denominator)
        pe[:, 1::2] = torch.cos(positions * denominator)
        return pe.to(dtype)

    def forward(self, x):
        if x.size(1) > self.max_len:
            raise RuntimeError("sequence longer than PositionalEncoding max_len")
        return self.table(x.size(1), self.input_size, x.dtype).unsqueeze(0)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_ffn, nhead, d_model, dropout=0.0, normalize_before=True):
        super().__init__()
        self.self_att = MultiheadAttention(nhead=nhead, d_model=d_model, dropout=dropout)
        self.pos_ffn = PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout)
        self.norm1 = LayerNorm(d_model, eps=LN_EPS)
        self.norm2 = LayerNorm(d_model, eps=LN_EPS)
        self.normalize_before = normalize_before

    def forward(self, src):
        src1 = self.norm1(src) if self.normalize_before else src
        output, attn = self.self_att(src1, src1, src1)
        src = src + output
        if not self.normalize_before:
            src = self.norm1(src)
        src1 = self.norm2(src) if self.normalize_before else src
        output = src + self.pos_ffn(src1)
        if not self.normalize_before:
            output = self.norm2(output)
        return output, attn


class TransformerEncoder(nn.Module):
    def __init__(self, num_layers, nhead, d_ffn, d_model, dropout=0.0, normalize_before=True):
        super().__init__()
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(d_ffn=d_ffn, nhead=nhead, d_model=d_model, dropout=dropout,
                                    normalize_before=normalize_before) for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=LN_EPS)

    def forward(self, src):
        output = src
        for layer in self.layers:
            output, _ = layer(output)
        return self.norm(output)


# ------------------------------------------------------------------------------ resepformer.py
class SBTransformerBlock_wnormandskip(nn.Module):
    """x -> mdl(x + PE) -> gLN over (D,len) -> + x  (skip taken BEFORE the PE add)."""

    def __init__(self, num_layers=N_LAYERS, d_model=N_ENC, nhead=N_HEAD, d_ffn=D_FFN, dropout=0.0,
                 use_positional_encoding=True, norm_before=True, use_norm=True, use_skip=True):
        super().__init__()
        self.use_positional_encoding = use_positional_encoding
        self.mdl = TransformerEncoder(num_layers=num_layers, nhead=nhead, d_ffn=d_ffn, d_model=d_model,
                                      dropout=dropout, normalize_before=norm_before)
        self.use_norm, self.use_skip = use_norm, use_skip
        if use_norm:
            self.norm = GlobalLayerNorm(d_model, 3, eps=GLN_EPS)   # select_norm("gln", d_model, 3, EPS)
        if use_positional_encoding:
            self.pos_enc = PositionalEncoding(input_size=d_model, max_len=PE_MAX_LEN)

    def forward(self, x):
        if self.use_positional_encoding:
            out = self.mdl(x + self.pos_enc(x))
        else:
            out = self.mdl(x)
        if self.use_norm:
            out = self.norm(out.permute(0, 2, 1)).permute(0, 2, 1)
        if self.use_skip:
            out = out + x
        return out


class ResourceEfficientSeparationPipeline(nn.Module):
    """Chunk, intra block, memory block over chunk means, intra block, un-chunk, output_fc.

    ``batch_mode``: ``"coupled"`` is upstream's literal behaviour for B>1 (the memory
    transformer and its gLN see the chunks of ALL batch items as one sequence, because of
    ``hc.unsqueeze(0)``); ``"independent"`` runs the memory block per item, which equals
    looping B=1 calls -- the only way the product ever calls it (api.py:1073-1077)."""

    def __init__(self, input_size, output_size, num_blocks, segment_size, seg_model, mem_model,
                 batch_mode="coupled"):
        super().__init__()
        self.segment_size, self.num_blocks, self.batch_mode = segment_size, num_blocks, batch_mode
        self.seg_model = nn.ModuleList([copy.deepcopy(seg_model) for _ in range(num_blocks)])
        self.mem_model = nn.ModuleList([copy.deepcopy(mem_model) for _ in range(num_blocks - 1)])
        self.output_fc = nn.Sequential(nn.PReLU(), nn.Conv1d(input_size, output_size, 1))

    def _padfeature(self, x):
        B, T, D = x.shape
        rest = self.segment_size - T % self.segment_size      # 1..K: a full zero chunk when T % K == 0
        if rest > 0:
            x = F.pad(x, (0, 0, 0, rest))
        return x, rest

    def forward(self, x):                                      # [B, L, D]
        B, T, D = x.shape
        x, _ = self._padfeature(x)
        x = x.view(B, -1, self.segment_size, D)
        B, S, K, D = x.shape
        output = x.reshape(B * S, K, D)
        hc = torch.zeros(output.shape[0], 1, output.shape[-1], dtype=output.dtype)
        for i in range(self.num_blocks):
            output = self.seg_model[i](output + hc)
            if i < self.num_blocks - 1:
                hc = output.mean(1).unsqueeze(0)               # [1, B*S, D]
                if self.batch_mode == "coupled":
                    hc = self.mem_model[i](hc).permute(1, 0, 2)
                else:
                    hc = torch.cat([self.mem_model[i](hc[:, b * S:(b + 1) * S]) for b in range(B)],
                                   dim=1).permute(1, 0, 2)
        output = output.reshape(B, S * K, D)[:, :T]
        return self.output_fc(output.transpose(1, 2)).transpose(1, 2)


class ResourceEfficientSeparator(nn.Module):
    def __init__(self, input_dim=N_ENC, num_spk=NUM_SPKS, layer=N_BLOCKS, segment_size=SEGMENT,
                 seg_model=None, mem_model=None, batch_mode="coupled"):
        super().__init__()
        self.num_spk, self.segment_size = num_spk, segment_size
        self.model = ResourceEfficientSeparationPipeline(
            input_size=input_dim, output_size=input_dim * num_spk, num_blocks=layer,
            segment_size=segment_size, seg_model=seg_model, mem_model=mem_model, batch_mode=batch_mode)
        self.nonlinear = nn.ReLU()

    def forward(self, inpt):                                   # [B, N, L]
        inpt = inpt.permute(0, 2, 1)
        B, T, N = inpt.shape
        processed = self.model(inpt).reshape(B, T, N, self.num_spk)   # channel c = n*num_spk + s
        masks = self.nonlinear(processed).unbind(dim=3)
        return torch.stack([m.permute(0, 2, 1) for m in masks])       # [spk, B, N, L]


# ---------------------------------------------------------------------- inference/separation.py
@dataclass
class _HParams:
    num_spks: int = NUM_SPKS
    sample_rate: int = SAMPLE_RATE


class OracleSepformerSeparation(nn.Module):
    """speechbrain.inference.separation.SepformerSeparation restated (separate_batch only).

    Construction order = upstream's yaml order (encoder, seg template, mem template, masknet
    with deepcopies of the templates, decoder) so that one ``torch.manual_seed`` reproduces
    the same random-init weights every time; seg_model[0] and seg_model[1] start identical.
    """

    def __init__(self, seed: int | None = 0, batch_mode: str = "coupled", num_layers=N_LAYERS,
                 d_model=N_ENC, nhead=N_HEAD, d_ffn=D_FFN, segment_size=SEGMENT, num_spks=NUM_SPKS,
                 num_blocks=N_BLOCKS, distinct_blocks: bool = True):
        super().__init__()
        if seed is not None:
            torch.manual_seed(seed)
        encoder = Encoder(kernel_size=KERNEL, out_channels=d_model)
        seg = SBTransformerBlock_wnormandskip(num_layers, d_model, nhead, d_ffn)
        mem = SBTransformerBlock_wnormandskip(num_layers, d_model, nhead, d_ffn)
        masknet = ResourceEfficientSeparator(d_model, num_spks, num_blocks, segment_size, seg, mem, batch_mode)
        decoder = Decoder(in_channels=d_model, out_channels=1, kernel_size=KERNEL, stride=STRIDE, bias=False)
        self.mods = nn.ModuleDict({"encoder": encoder, "masknet": masknet, "decoder": decoder})
        self.hparams = _HParams(num_spks=num_spks)
        self.device = torch.device("cpu")
        if distinct_blocks and seed is not None:
            # Upstream's deepcopy makes seg_model[0] == seg_model[1] at init, and PyTorch inits
            # LayerNorm/gLN affine to (1,0) and attention biases to 0 -- a trained checkpoint has
            # none of these degeneracies, and a kernel that mixes up the two blocks or drops a
            # bias would still pass.  Perturb every parameter deterministically so each one matters.
            g = torch.Generator().manual_seed(seed + 7919)
            with torch.no_grad():
                for name, p in self.named_parameters():
                    if not name.startswith("mods.masknet"):
                        continue
                    if p.dim() >= 2 and p.numel() > p.shape[0]:      # weight matrices
                        p.add_(0.25 * p.std() * torch.randn(p.shape, generator=g))
                    else:                                            # LN / gLN affine, biases, PReLU slope
                        p.add_(0.05 * torch.randn(p.shape, generator=g))
        self.eval()

    @property
    def batch_mode(self):
        return self.mods["masknet"].model.batch_mode

    @batch_mode.setter
    def batch_mode(self, v):
        assert v in ("coupled", "independent")
        self.mods["masknet"].model.batch_mode = v

    @torch.no_grad()
    def separate_batch(self, mix):
        mix = mix.to(self.device)
        mix_w = self.mods["encoder"](mix)
        est_mask = self.mods["masknet"](mix_w)
        mix_w = torch.stack([mix_w] * self.hparams.num_spks)
        sep_h = mix_w * est_mask
        est_source = torch.cat(
            [self.mods["decoder"](sep_h[i]).unsqueeze(-1) for i in range(self.hparams.num_spks)], dim=-1)
        T_origin, T_est = mix.size(1), est_source.size(1)
        if T_origin > T_est:
            est_source = F.pad(est_source, (0, 0, 0, T_origin - T_est))
        else:
            est_source = est_source[:, :T_origin, :]
        return est_source

    forward = separate_batch

    def component_state_dicts(self):
        """The three dicts upstream saves as encoder.ckpt / masknet.ckpt / decoder.ckpt."""
        return {k: {n: t.detach().clone() for n, t in self.mods[k].state_dict().items()}
                for k in ("encoder", "masknet", "decoder")}


def count_params(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


def frames_of(T: int) -> int:
    """L = floor((T - 16) / 8) + 1 encoder frames (Conv1d k16 s8, no padding)."""
    return (T - KERNEL) // STRIDE + 1


def chunks_of(L: int, K: int = SEGMENT) -> int:
    """S = floor(L / K) + 1 (``rest = K - L % K`` is in [1, K])."""
    return L // K + 1
