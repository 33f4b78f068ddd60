"""Development probe: accuracy (max-abs vs the oracle, est-vs-est SI-SNR) and speed of the fp16 mode (full hi + lo
weights, and RESEP_W16F=mixed) next to tf32 and bf16, over several weight seeds and shapes."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import torch
    from clearconverse_b200 import SepformerSeparation, synth
    from clearconverse_b200.metrics import est_vs_est_db
    from oracle.resepformer_oracle import OracleSepformerSeparation
    torch.set_num_threads(os.cpu_count())
    rows = []
    for wseed in (0, 1, 2, 3):
        oracle = OracleSepformerSeparation(seed=wseed)
        sds = oracle.component_state_dicts()
        cases = [synth.synth_batch(2, 2000, 2), synth.synth_batch(1, 32000, 1), synth.synth_batch(3, 9000, 5), synth.synth_batch(8, 32000, 4)]
        wants = [oracle.separate_batch(x) for x in cases]
        for prec in sys.argv[2].split(","):
            with SepformerSeparation(sds, device="cuda:0", precision=prec) as sep:
                for x, w in zip(cases, wants):
                    g = sep.separate_batch(x).cpu()
                    rows.append({"weight_seed": wseed, "precision": prec, "B": x.shape[0], "T": x.shape[1],
                                 "max_abs": (g - w).abs().max().item(), "est_vs_est_db": est_vs_est_db(g, w)})
                    print(rows[-1], flush=True)
    print("WORST", {p: max(r["max_abs"] for r in rows if r["precision"] == p) for p in sys.argv[2].split(",")})
else:
    subprocess.run([sys.executable, __file__, "worker", "fp16,tf32"], check=False)
    subprocess.run([sys.executable, __file__, "worker", "fp16"], check=False, env=dict(os.environ, RESEP_W16F="mixed"))
    for prec, env in (("fp16", {}), ("fp16", {"RESEP_W16F": "mixed"}), ("tf32", {}), ("bf16", {})):
        print("==", prec, env, flush=True)
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_quick.py"), prec], env=dict(os.environ, **env), check=False)
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_mode_profile.py"), "fp16"], check=False)
