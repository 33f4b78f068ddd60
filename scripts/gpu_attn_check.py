"""Development probe for k_attn_tc: per-kernel times of a config-2 forward with the tcgen05 attention, with its
polynomial path off, and with the mma.sync predecessor (each in a subprocess: the knobs are read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for env in ({}, {"RESEP_ATTN_POLY": "0"}, {"RESEP_ATTN_TC": "0"}):
    print("==", env or "default", flush=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_mode_profile.py"), "bf16"], env=dict(os.environ, **env), check=False)
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_quick.py"), "bf16"], env=dict(os.environ, **env), check=False)
