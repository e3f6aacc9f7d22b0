"""circuitmap_b200 -- B200-native (sm_100a) implementation of circuitmap's two data-parallel hot paths.

    from circuitmap_b200 import NeuralDemixer, Model
    psc_dem = NeuralDemixer(path='demixers/nwd_ie_ChroME2f.ckpt')(psc)
    model = Model(N); model.fit(psc_dem, stim_matrix, method='caviar', fit_options={...})

Mirrors circuitmap/__init__.py:1-3 of the reference for the hot-path entry points and the data generator (cavi_sns, mbcs,
cosamp, viz are outside the hot path).  Host code is Python; all arithmetic runs in hand-written CUDA behind
the C ABI of include/circuitmap_b200.h.  No CPU fallback.
"""
from .neural_waveform_demixing import NeuralDemixer
from .model import Model
from .simulation import simulate

__all__ = ["NeuralDemixer", "Model", "simulate"]
