"""CPU emulation of 16-bit operand variants of the tensor-core GEMMs (development probe, test infra).

Rounds GEMM operands the way the CUDA path does (activations to a 16-bit format, weights to a
16-bit format or to a hi+lo pair) inside oracle/functional_restatement's arithmetic and reports
max-abs and the SI-SNR delta of tests/test_gpu_parity.py against the fp32 oracle.
"""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import functional_restatement as fr
from oracle.resepformer_oracle import OracleSepformerSeparation
from clearconverse_b200.synth import synth_batch
from clearconverse_b200.metrics import si_snr_db, si_snr_delta

torch.set_num_threads(os.cpu_count())
D, H, DH = fr.D, fr.H, fr.DH


def rnd(x, fmt):
    if fmt == "f32":
        return x
    dt = torch.bfloat16 if fmt == "bf16" else torch.float16
    return x.to(dt).to(torch.float32)


def wq(w, fmt):
    """weight operand: 'bf16', 'fp16', 'bf16x2' (hi+lo), 'f32'"""
    if fmt == "bf16x2":
        hi = rnd(w, "bf16")
        return hi + rnd(w - hi, "bf16")
    return rnd(w, fmt)


class Cfg:
    def __init__(self, act="bf16", w_in="bf16", w_out="bf16", w_f1="bf16", w_f2="bf16", w_fc="bf16", attn="bf16"):
        self.__dict__.update(locals())


def layer(o, sd, pfx, c):
    y = rnd(fr._ln(o, sd[pfx + "norm1.norm.weight"], sd[pfx + "norm1.norm.bias"]), c.act)
    qkv = y @ wq(sd[pfx + "self_att.att.in_proj_weight"], c.w_in).T + sd[pfx + "self_att.att.in_proj_bias"]
    qkv = rnd(qkv, c.attn)
    q, k, v = qkv.split(D, dim=-1)
    Bx, n, _ = o.shape
    q = q.reshape(Bx, n, H, DH).transpose(1, 2)
    k = k.reshape(Bx, n, H, DH).transpose(1, 2)
    v = v.reshape(Bx, n, H, DH).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (DH ** -0.5)
    p = torch.exp(s - s.amax(-1, keepdim=True))
    l = p.sum(-1, keepdim=True)
    cx = (rnd(p, c.attn) @ v) / l
    cx = rnd(cx.transpose(1, 2).reshape(Bx, n, D), c.act)
    o = o + cx @ wq(sd[pfx + "self_att.att.out_proj.weight"], c.w_out).T + sd[pfx + "self_att.att.out_proj.bias"]
    y = rnd(fr._ln(o, sd[pfx + "norm2.norm.weight"], sd[pfx + "norm2.norm.bias"]), c.act)
    h = rnd(torch.relu(y @ wq(sd[pfx + "pos_ffn.ffn.0.weight"], c.w_f1).T + sd[pfx + "pos_ffn.ffn.0.bias"]), c.act)
    return o + h @ wq(sd[pfx + "pos_ffn.ffn.3.weight"], c.w_f2).T + sd[pfx + "pos_ffn.ffn.3.bias"]


def block(x, sd, pfx, c):
    o = x + fr.pe_table(x.shape[1], x.dtype)
    for l in range(fr.NLAYERS):
        o = layer(o, sd, f"{pfx}mdl.layers.{l}.", c)
    o = fr._ln(o, sd[pfx + "mdl.norm.norm.weight"], sd[pfx + "mdl.norm.norm.bias"])
    mu = o.mean(dim=(1, 2), keepdim=True)
    var = ((o - mu) ** 2).mean(dim=(1, 2), keepdim=True)
    gw, gb = sd[pfx + "norm.weight"].reshape(1, 1, D), sd[pfx + "norm.bias"].reshape(1, 1, D)
    return gw * (o - mu) / torch.sqrt(var + fr.GLN_EPS) + gb + x


@torch.no_grad()
def separate(mix, sds, c):
    B, T = mix.shape
    mk = sds["masknet"]
    w = fr.encode(mix, sds["encoder"]["conv1d.weight"])
    L = w.shape[1]
    rest = 150 - L % 150
    S = (L + rest) // 150
    x = torch.nn.functional.pad(w, (0, 0, 0, rest)).reshape(B * S, 150, D)
    o = block(x, mk, "model.seg_model.0.", c)
    m = o.mean(dim=1)
    hc = block(m[None], mk, "model.mem_model.0.", c)[0]
    o = block(o + hc[:, None, :], mk, "model.seg_model.1.", c)
    o = o.reshape(B, S * 150, D)[:, :L]
    a = mk["model.output_fc.0.weight"]
    o = rnd(torch.where(o >= 0, o, a * o), c.act)
    o = o @ wq(mk["model.output_fc.1.weight"].reshape(2 * D, D), c.w_fc).T + mk["model.output_fc.1.bias"]
    mask = torch.relu(o.reshape(B, L, D, 2))
    return torch.stack([fr.decode(w * mask[..., s], sds["decoder"]["weight"], T) for s in range(2)], dim=-1)


if __name__ == "__main__":
    oracle = OracleSepformerSeparation(seed=0)
    sds = oracle.component_state_dicts()
    cases = [synth_batch(1, 32000, 1), synth_batch(3, 9000, 5), synth_batch(2, 2000, 2)]
    wants = [oracle.separate_batch(x) for x in cases]
    variants = {
        "f32 (sanity)": Cfg("f32", "f32", "f32", "f32", "f32", "f32", "f32"),
        "bf16 single": Cfg(),
        "bf16 act, W hi+lo (current)": Cfg("bf16", "bf16x2", "bf16x2", "bf16x2", "bf16x2", "bf16x2"),
        "fp16 single (attn bf16)": Cfg("fp16", "fp16", "fp16", "fp16", "fp16", "fp16", "bf16"),
        "fp16 single (attn fp16)": Cfg("fp16", "fp16", "fp16", "fp16", "fp16", "fp16", "fp16"),
        "bf16; split in/out/fc only": Cfg("bf16", "bf16x2", "bf16x2", "bf16", "bf16", "bf16x2"),
        "bf16; split f1,f2 only": Cfg("bf16", "bf16", "bf16", "bf16x2", "bf16x2", "bf16"),
        "bf16; split f1 only": Cfg("bf16", "bf16", "bf16", "bf16x2", "bf16", "bf16"),
        "bf16; split f2 only": Cfg("bf16", "bf16", "bf16", "bf16", "bf16x2", "bf16"),
        "bf16 act, W f32": Cfg("bf16", "f32", "f32", "f32", "f32", "f32"),
        "f32 act, W bf16": Cfg("f32", "bf16", "bf16", "bf16", "bf16", "bf16", "f32"),
    }
    sel = sys.argv[1:]
    for name, c in variants.items():
        if sel and not any(s in name for s in sel):
            continue
        row = []
        for x, wnt in zip(cases, wants):
            g = separate(x, sds, c)
            row.append(f"{(g - wnt).abs().max():.1e} {si_snr_db(g.permute(0, 2, 1), wnt.permute(0, 2, 1)).min():.1f}dB d={si_snr_delta(g, wnt, x)[0]:.4f}")
        print(f"{name:32s}", " | ".join(row), flush=True)
