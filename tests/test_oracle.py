"""CPU tests of the oracle (test infrastructure): self-consistency, upstream structural facts,
edge cases, and the committed golden vectors.  PARITY UNPINNED -- see oracle/resepformer_oracle.py."""
import glob
import os

import numpy as np
import pytest
import torch

from clearconverse_b200.synth import synth_batch
from oracle import functional_restatement as fr
from oracle import resepformer_oracle as ro

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_param_count_matches_model_card():
    m = ro.OracleSepformerSeparation(seed=0, distinct_blocks=False)
    assert ro.count_params(m) == ro.EXPECTED_PARAMS == 7_955_201
    assert ro.count_params(m.mods["encoder"]) == 2048 and ro.count_params(m.mods["decoder"]) == 2048


def test_deepcopy_makes_blocks_identical_at_default_init():
    m = ro.OracleSepformerSeparation(seed=0, distinct_blocks=False)
    sd = m.mods["masknet"].state_dict()
    k = "mdl.layers.3.pos_ffn.ffn.0.weight"
    assert torch.equal(sd["model.seg_model.0." + k], sd["model.seg_model.1." + k])
    m2 = ro.OracleSepformerSeparation(seed=0)   # the test weights perturb every parameter
    sd2 = m2.mods["masknet"].state_dict()
    assert not torch.equal(sd2["model.seg_model.0." + k], sd2["model.seg_model.1." + k])


def test_state_dict_keys_are_upstreams(sds):
    from clearconverse_b200.weights import expected_shapes
    exp = expected_shapes()
    for comp in ("encoder", "masknet", "decoder"):
        assert set(sds[comp].keys()) == set(exp[comp].keys())
        for k, shp in exp[comp].items():
            assert tuple(sds[comp][k].shape) == tuple(shp), k
    assert len(sds["masknet"]) == 303


def test_upstream_docstring_shape_example():
    # resepformer.py docstring: seg/mem = SBTransformerBlock_wnormandskip(1, 64, 8); x [10,64,100];
    # ResourceEfficientSeparator(64, num_spk=3, mem_type='av') with defaults layer=3, segment_size=20 -> [3,10,64,100]
    torch.manual_seed(0)
    seg = ro.SBTransformerBlock_wnormandskip(1, 64, 8, d_ffn=2048, use_positional_encoding=False, norm_before=False)
    mem = ro.SBTransformerBlock_wnormandskip(1, 64, 8, d_ffn=2048, use_positional_encoding=False, norm_before=False)
    net = ro.ResourceEfficientSeparator(64, num_spk=3, layer=3, segment_size=20, seg_model=seg, mem_model=mem)
    out = net(torch.randn(10, 64, 100))
    assert out.shape == (3, 10, 64, 100)


@pytest.mark.parametrize("mode", ["coupled", "independent"])
def test_two_restatements_agree(oracle, sds, mode):
    oracle.batch_mode = mode
    mix = synth_batch(2, 4000, seed=1)
    a = oracle.separate_batch(mix)
    b = fr.separate(mix, sds, mode)
    assert a.shape == (2, 4000, 2)
    assert (a - b).abs().max().item() < 2e-5
    b64 = fr.separate(mix, sds, mode, dtype=torch.float64)
    assert (a.double() - b64).abs().max().item() < 2e-5
    oracle.batch_mode = "coupled"


def test_independent_equals_looping_b1_and_differs_from_coupled(oracle):
    mix = synth_batch(2, 4000, seed=1)
    oracle.batch_mode = "independent"
    ind = oracle.separate_batch(mix)
    loop = torch.cat([oracle.separate_batch(mix[i:i + 1]) for i in range(2)])
    oracle.batch_mode = "coupled"
    cpl = oracle.separate_batch(mix)
    assert (ind - loop).abs().max().item() < 1e-5
    assert (ind - cpl).abs().max().item() > 1e-3      # upstream's batched call couples the items


def test_edge_cases(oracle):
    # api.py:858 can hand in zeros(1,100): L=11, S=1, output all-zero
    z = oracle.separate_batch(torch.zeros(1, 100))
    assert z.shape == (1, 100, 2) and z.abs().max().item() == 0.0
    # T < 16: upstream's Conv1d raises
    with pytest.raises(RuntimeError):
        oracle.separate_batch(torch.zeros(1, 10))
    # T not a multiple of 8 -> trailing T - T_est samples are exact zeros; L % 150 == 0 -> S = L/150 + 1
    T = 16 + 8 * 149 + 3
    assert ro.frames_of(T) == 150 and ro.chunks_of(150) == 2
    y = oracle.separate_batch(synth_batch(1, T, 3))
    assert y.shape == (1, T, 2) and y[0, -3:].abs().max().item() == 0.0 and y[0, -4].abs().max().item() > 0
    # minimum length
    assert oracle.separate_batch(synth_batch(1, 16, 4)).shape == (1, 16, 2)


def test_oracle_fp64_error_budget(oracle, sds):
    mix = synth_batch(1, 8000, seed=9)
    a = oracle.separate_batch(mix)
    b = fr.separate(mix, sds, "coupled", dtype=torch.float64)
    assert (a.double() - b).abs().max().item() < 2e-5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_oracle_reproduces_golden(oracle, path):
    g = np.load(path)
    B, T, seed, mode = (int(v) for v in g["meta"])
    cs = {k: float(sum(v.double().abs().sum() for v in sd.values())) for k, sd in oracle.component_state_dicts().items()}
    assert abs(cs["masknet"] - float(g["w_mask"])) < 1e-6 * float(g["w_mask"]), "seeded weights drifted"
    mix = synth_batch(B, T, seed)
    assert abs(mix.double().abs().sum().item() - float(g["mix_checksum"])) < 1e-6 * float(g["mix_checksum"])
    oracle.batch_mode = "coupled" if mode == 0 else "independent"
    est = oracle.separate_batch(mix).numpy()
    oracle.batch_mode = "coupled"
    assert est.shape == g["est"].shape
    assert np.abs(est - g["est"]).max() < 5e-5
