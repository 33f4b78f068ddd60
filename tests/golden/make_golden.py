"""Generates tests/golden/*.npz from the oracle (run here, in the build container):

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference holds no golden vectors for this path and upstream speechbrain
cannot be imported here, so these vectors are outputs of THIS repo's oracle
(oracle/resepformer_oracle.py, seed 0, fp32, torch CPU).  They pin the oracle against drift
(torch version, thread count, host CPU) and give the GPU tests a target that does not need
the oracle at run time.  Inputs are regenerated from clearconverse_b200.synth (seeded); only
outputs and a weight checksum are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from clearconverse_b200.synth import synth_batch  # noqa: E402
from oracle.resepformer_oracle import OracleSepformerSeparation  # noqa: E402

# name -> (batch, samples, synth seed, batch_mode)
CASES = {
    "b1_t32000_cfg1": (1, 32000, 1, "coupled"),       # BASELINE.json configs[0]: one 4 s mixture
    "b2_t2000_coupled": (2, 2000, 2, "coupled"),
    "b2_t2000_independent": (2, 2000, 2, "independent"),
    "b1_t1211_fullchunk": (1, 16 + 8 * 149 + 3, 3, "coupled"),   # L == 150 -> extra all-zero chunk, 3 trailing zeros
    "b3_t9000_coupled": (3, 9000, 5, "coupled"),
}


def weight_checksums(model):
    return {k: float(sum(v.double().abs().sum() for v in sd.values()))
            for k, sd in model.component_state_dicts().items()}


def main():
    torch.set_num_threads(8)
    model = OracleSepformerSeparation(seed=0)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (B, T, seed, mode) in CASES.items():
        model.batch_mode = mode
        mix = synth_batch(B, T, seed)
        est = model.separate_batch(mix).numpy()
        cs = weight_checksums(model)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), est=est.astype(np.float32),
                            mix_checksum=np.float64(mix.double().abs().sum().item()),
                            w_enc=cs["encoder"], w_mask=cs["masknet"], w_dec=cs["decoder"],
                            meta=np.array([B, T, seed, 0 if mode == "coupled" else 1]))
        print(name, est.shape, float(np.abs(est).max()))


if __name__ == "__main__":
    main()
