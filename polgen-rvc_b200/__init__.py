"""polgen-rvc_b200: B200-native (sm_100a) RVC synthesizer decode.

Drop-in for the reference ``rvc.lib.algorithm.synthesizers.Synthesizer``
(``rvc/lib/algorithm/synthesizers.py:13``): same constructor, same
``infer(phone, phone_lengths, pitch, nsff0, sid)`` call, same state-dict keys.
The compute is hand-written CUDA behind the C ABI in ``include/polgen_rvc.h``;
there is no CPU or PyTorch fallback -- importing the synthesizer without the
built extension raises.
"""
from .configs import CONFIGS, SynthConfig, config_from_ctor, synth_inputs, synth_noise, synth_weights

__all__ = ["CONFIGS", "SynthConfig", "config_from_ctor", "synth_inputs", "synth_noise",
           "synth_weights"]
from .synthesizer import (Engine, SegmentScheduler, Synthesizer, SynthesizerTrnMs256NSFsid,
                          SynthesizerTrnMs768NSFsid, fold_state_dict)

__all__ += ["Engine", "SegmentScheduler", "Synthesizer", "SynthesizerTrnMs256NSFsid", "SynthesizerTrnMs768NSFsid",
            "fold_state_dict"]
from .pipeline import ClipConverter, coarse_pitch, postprocess, prepare_features

__all__ += ["ClipConverter", "coarse_pitch", "postprocess", "prepare_features"]
from .tiling import TimeTiledDecoder, plan_tiles

__all__ += ["TimeTiledDecoder", "plan_tiles"]
