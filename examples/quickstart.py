"""The usage patterns of the reference's README, unchanged except for the import, on synthetic logits.

    python examples/quickstart.py          # needs a CUDA device: there is no CPU path

Every call below exists with the same name and arguments in kensho-technologies/pyctcdecode
(README.md:30-90, decoder.py:730-945, :669-728); `pool` arguments are accepted and ignored -- utterances
run in parallel on the device.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from pyctcdecode_b200 import build_ctcdecoder  # noqa: E402  (reference: from pyctcdecode import build_ctcdecoder)
from pyctcdecode_b200.language_model import HotwordScorer  # noqa: E402
from tests import synth  # noqa: E402  (synthetic vocabulary, n-gram model and CTC-shaped logits)


def main():
    wl = synth.CharWorkload("A", n_words=400, lm_order=3)          # 28 characters + blank, synthetic 3-gram ARPA
    logits_list = [wl.utterance(100 + i, 150 + 10 * i, "peaky") for i in range(4)]
    logits = logits_list[0]

    # shallow fusion with an n-gram model (README.md:37-44)
    decoder = build_ctcdecoder(wl.labels, kenlm_model_path=wl.arpa, unigrams=wl.words, alpha=0.5, beta=1.0)
    text = decoder.decode(logits)
    print("decode            :", text[:60])

    # hotwords (README.md:62-68)
    print("with hotwords     :", decoder.decode(logits, hotwords=[wl.words[3], wl.words[10]], hotword_weight=10.0)[:60])

    # batch (README.md:80-85): the pool argument may be None or a multiprocessing pool, it is not used
    print("decode_batch      :", [t[:20] for t in decoder.decode_batch(None, logits_list)])

    # full beam results: text, LM state, word frames, scores (decoder.py:102-118)
    beams = decoder.decode_beams(logits, beam_width=20)
    top = beams[0]
    print("decode_beams      : %d beams, top %r logit %.3f lm %.3f" % (len(beams), top.text[:30], top.logit_score, top.lm_score))
    print("word frames       :", top.text_frames[:3])

    # stateful decoding: carry the LM state of the best beam into the next utterance (tests/test_decoder.py:447-456)
    second = decoder.decode_beams(logits_list[1], lm_start_state=top.last_lm_state)[0]
    print("next utterance    :", second.text[:40])

    # streaming, chunk by chunk (decoder.py:669-728)
    state_beams, cached_lm_scores, cached_p_lm_scores = decoder.get_starting_state()
    scorer = HotwordScorer.build_scorer([wl.words[3]], weight=10.0)
    for start in range(0, len(logits), 50):
        chunk = logits[start:start + 50]
        state_beams = decoder.partial_decode_beams(chunk, cached_lm_scores, cached_p_lm_scores, state_beams, start,
                                                   hotword_scorer=scorer, is_end=start + 50 >= len(logits))
        print("  after frame %3d : %r + %r" % (start + len(chunk), state_beams[0].text[-25:], state_beams[0].partial_word))

    # a BPE vocabulary is recognised from its labels (README.md:50-58)
    bpe = synth.BpeWorkload(n_words=400, lm_order=0)
    print("BPE               :", build_ctcdecoder(bpe.labels).decode(bpe.utterance(7, 80, "peaky"))[:60])
    return text


if __name__ == "__main__":
    main()
