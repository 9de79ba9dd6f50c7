"""pyctcdecode_b200 -- B200-native CTC prefix beam-search decoder with the pyctcdecode API.

    from pyctcdecode_b200 import build_ctcdecoder
    decoder = build_ctcdecoder(labels, kenlm_model_path="lm.arpa", alpha=0.5, beta=1.0)
    texts = decoder.decode_batch(None, logits_list, beam_width=100)

The per-frame beam update runs in hand-written sm_100a CUDA kernels behind a C ABI
(include/b200ctc.h, pyctcdecode_b200/libb200ctc.so).  There is no CPU fallback.
"""
from .alphabet import Alphabet  # noqa: F401
from .decoder import Beam, BeamSearchDecoderCTC, LMBeam, OutputBeam, build_ctcdecoder  # noqa: F401
from .language_model import HotwordScorer, LanguageModel, MultiLanguageModel, MultiLanguageModelState, NgramModel  # noqa: F401

__version__ = "0.1.0"
