"""Second, independent CPU restatement of the RE-SepFormer forward.  TEST INFRASTRUCTURE ONLY.

Written directly from the algorithm statement in SURVEY.md Appendix A, with plain matmuls,
explicit per-head attention and explicit overlap-add -- no ``nn.Conv1d``,
``nn.MultiheadAttention``, ``nn.LayerNorm`` or ``nn.ConvTranspose1d``.  It consumes the three
component state dicts of ``oracle.resepformer_oracle`` (upstream's key names) and exists so
that the module-tree oracle is checked by something that does not share its code: the two
must agree to float rounding in fp32 and to ~1e-12 in fp64.  It also exposes the
intermediate tensors (encoder features, per-block outputs, chunk summaries, masks) that the
per-kernel GPU parity tests compare against.

PARITY UNPINNED (see oracle/resepformer_oracle.py): upstream speechbrain is not available.
"""
from __future__ import annotations

import math

import torch

K_CHUNK, D, H, DH, F_FFN, SPK, KSZ, STRIDE, NLAYERS = 150, 128, 8, 16, 1024, 2, 16, 8, 8
LN_EPS = 1e-6
GLN_EPS = float(torch.finfo(torch.float32).eps)


def pe_table(length: int, dtype=torch.float32) -> torch.Tensor:
    """PE[p,2i]=sin(p*exp(-2i*ln(1e4)/D)), PE[p,2i+1]=cos(same); evaluated in fp32 like upstream
    (the table is a stored fp32 buffer there), then cast."""
    pos = torch.arange(0, length).unsqueeze(1).float()
    den = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
    pe = torch.zeros(length, D)
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.to(dtype)


def _ln(x, w, b, eps=LN_EPS):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def encoder_layer(o, sd, pfx):
    """One pre-norm TransformerEncoderLayer on o [Bx, len, D]."""
    y = _ln(o, sd[pfx + "norm1.norm.weight"], sd[pfx + "norm1.norm.bias"])
    qkv = y @ sd[pfx + "self_att.att.in_proj_weight"].T + sd[pfx + "self_att.att.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    Bx, n, _ = o.shape
    q = q.reshape(Bx, n, H, DH).transpose(1, 2) * (DH ** -0.5)
    k = k.reshape(Bx, n, H, DH).transpose(1, 2)
    v = v.reshape(Bx, n, H, DH).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    c = (a @ v).transpose(1, 2).reshape(Bx, n, D)
    o = o + c @ sd[pfx + "self_att.att.out_proj.weight"].T + sd[pfx + "self_att.att.out_proj.bias"]
    y = _ln(o, sd[pfx + "norm2.norm.weight"], sd[pfx + "norm2.norm.bias"])
    h = torch.relu(y @ sd[pfx + "pos_ffn.ffn.0.weight"].T + sd[pfx + "pos_ffn.ffn.0.bias"])
    return o + h @ sd[pfx + "pos_ffn.ffn.3.weight"].T + sd[pfx + "pos_ffn.ffn.3.bias"]


def block(x, sd, pfx, nlayers=NLAYERS):
    """SBTransformerBlock_wnormandskip on x [Bx, len, D]."""
    o = x + pe_table(x.shape[1], x.dtype)
    for l in range(nlayers):
        o = encoder_layer(o, sd, f"{pfx}mdl.layers.{l}.")
    o = _ln(o, sd[pfx + "mdl.norm.norm.weight"], sd[pfx + "mdl.norm.norm.bias"])
    mu = o.mean(dim=(1, 2), keepdim=True)
    var = ((o - mu) ** 2).mean(dim=(1, 2), keepdim=True)
    gw, gb = sd[pfx + "norm.weight"].reshape(1, 1, D), sd[pfx + "norm.bias"].reshape(1, 1, D)
    return gw * (o - mu) / torch.sqrt(var + GLN_EPS) + gb + x


def encode(mix, w_enc):
    """relu(conv1d k16 s8) as unfold + matmul -> token-major [B, L, D]."""
    frames = mix.unfold(1, KSZ, STRIDE)                         # [B, L, 16]
    return torch.relu(frames @ w_enc.reshape(D, KSZ).T)


def decode(h, w_dec, T):
    """ConvTranspose1d(128,1,16,8) as matmul + explicit overlap-add; zero-pad / crop to T."""
    B, L, _ = h.shape
    fr = h @ w_dec.reshape(D, KSZ)                              # [B, L, 16]
    t_est = STRIDE * (L - 1) + KSZ
    out = torch.zeros(B, t_est + STRIDE, dtype=h.dtype)
    out[:, :STRIDE * L] += fr[:, :, :STRIDE].reshape(B, -1)
    out[:, STRIDE:STRIDE * (L + 1)] += fr[:, :, STRIDE:].reshape(B, -1)
    out = out[:, :t_est]
    if T > t_est:
        out = torch.nn.functional.pad(out, (0, T - t_est))
    return out[:, :T]


@torch.no_grad()
def separate(mix, sds, batch_mode="coupled", dtype=torch.float32, want_intermediates=False):
    """mix [B,T] -> est [B,T,2].  ``sds`` = {'encoder','masknet','decoder'} state dicts."""
    if mix.dim() != 2:
        raise RuntimeError("mix must be [B, T]")
    B, T = mix.shape
    if T < KSZ:
        raise RuntimeError("Kernel size can't be greater than actual input size")
    mix = mix.to(dtype)
    enc = {k: v.to(dtype) for k, v in sds["encoder"].items()}
    mk = {k: v.to(dtype) for k, v in sds["masknet"].items()}
    dec = {k: v.to(dtype) for k, v in sds["decoder"].items()}
    inter = {}

    w = encode(mix, enc["conv1d.weight"])                       # [B, L, D]
    L = w.shape[1]
    rest = K_CHUNK - L % K_CHUNK
    S = (L + rest) // K_CHUNK
    x = torch.nn.functional.pad(w, (0, 0, 0, rest)).reshape(B * S, K_CHUNK, D)
    inter["enc"] = w
    o = block(x, mk, "model.seg_model.0.")
    inter["seg0"] = o
    m = o.mean(dim=1)                                           # [B*S, D]
    inter["chunk_mean"] = m
    if batch_mode == "coupled":
        hc = block(m[None], mk, "model.mem_model.0.")[0]
    else:
        hc = torch.cat([block(m[None, b * S:(b + 1) * S], mk, "model.mem_model.0.")[0] for b in range(B)])
    inter["mem0"] = hc
    o = block(o + hc[:, None, :], mk, "model.seg_model.1.")
    inter["seg1"] = o
    o = o.reshape(B, S * K_CHUNK, D)[:, :L]
    a = mk["model.output_fc.0.weight"]
    o = torch.where(o >= 0, o, a * o)
    o = o @ mk["model.output_fc.1.weight"].reshape(SPK * D, D).T + mk["model.output_fc.1.bias"]
    mask = torch.relu(o.reshape(B, L, D, SPK))
    inter["mask"] = mask
    est = torch.stack([decode(w * mask[..., s], dec["weight"], T) for s in range(SPK)], dim=-1)
    return (est, inter) if want_intermediates else est
