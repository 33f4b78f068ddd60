"""clearconverse_b200 -- B200-native RE-SepFormer overlap separation for ClearConverse.

One hot path of Yotsuei/ClearConverse, re-built from scratch for sm_100a: the
``separate_batch(mix[B,T]) -> est_sources[B,T,n_spk]`` call that ``back/api.py:1077`` makes on
every overlapping-speech segment.  ``SepformerSeparation`` keeps upstream's surface; the work is
done by hand-written CUDA kernels in ``libresep_b200.so`` behind the C ABI of
``include/resep_b200.h``.  No CPU fallback.
"""
from .separation import SepformerSeparation  # noqa: F401
from . import synth, weights  # noqa: F401

__all__ = ["SepformerSeparation", "synth", "weights"]
__version__ = "0.1.0"
