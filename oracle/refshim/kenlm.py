"""TEST INFRASTRUCTURE ONLY -- ARPA-backed stand-in for the ``kenlm`` python module.

The reference needs ``kenlm`` (third party, kpu/kenlm master, UNPINNED: reference
``.github/workflows/tests_and_lint.yml:32,51``; ``decoder.py:55-61``) and it is not installed
here.  This module restates KenLM's published query algorithm for an ARPA-loaded probing
model so that the unmodified reference can be imported in this container to (a) run its own
tests and (b) generate golden vectors (``oracle/gen_golden.py``).  It is never imported by
the product package and does not travel as part of the measured path.

Call sites served (reference): ``kenlm.Model(path)`` decoder.py:1074, language_model.py:451;
``word in model`` language_model.py:95,352; ``kenlm.State()`` :109; ``.order`` :306;
``.BeginSentenceWrite`` :312; ``.NullContextWrite`` :314; ``.BaseScore(in, word, out)``
:321,347; ``.path`` (bytes) :387,399,402.

Restated KenLM semantics (lm/model.cc GenericModel::FullScore / ScoreExceptBackoff):
  * log10 probabilities and backoffs are stored as float32; score accumulation is float32;
  * p(w | ctx) = prob of the longest n-gram ``ctx[-k:] + w`` present, plus the backoffs of
    every longer context that was skipped, added from the shorter to the longer context;
  * OOV words (including the empty string) map to ``<unk>`` (word index 0);
    ``<unk>`` gets log10 p = -100 if the ARPA file has none;
  * the out state keeps the longest matched n-gram (length <= order-1) that "has an
    extension", i.e. whose backoff is non-zero or which is the context (first n words) of a
    longer n-gram in the model -- KenLM's state minimisation;
  * ``word in model`` is ``vocab.Index(word) != 0`` so ``<unk>`` itself is "not in" the model.
parity unpinned: the reference's tests only exercise a 2-gram toy model whose backoffs are
all zero; order>=3 and non-zero backoffs are pinned by hand-computed cases in
``tests/test_oracle_lm.py`` instead.
"""
import os

import numpy as np

_F = np.float32


class State:
    __slots__ = ("words", "backoffs")

    def __init__(self):
        self.words = ()  # most recent first
        self.backoffs = ()


class Model:
    def __init__(self, path):
        self.path = os.path.abspath(path).encode("utf-8")
        self._vocab = {"<unk>": 0}
        self._uni = {}  # wid -> (prob, backoff)
        self._ngrams = {}  # tuple of wids in sentence order (len>=2) -> [prob, backoff]
        self._ext = set()  # tuples (len>=1) that have an extension
        self.order = 0
        self._load(path)

    def _wid(self, word, create=False):
        wid = self._vocab.get(word)
        if wid is None:
            if not create:
                return 0
            wid = len(self._vocab)
            self._vocab[word] = wid
        return wid

    def _load(self, path):
        cur = 0
        with open(path, encoding="utf-8") as fh:
            for raw in fh:
                line = raw.strip()
                if not line:
                    continue
                if line.startswith("\\"):
                    if line.endswith("-grams:"):
                        cur = int(line[1 : line.index("-")])
                        self.order = max(self.order, cur)
                    elif line == "\\end\\":
                        break
                    continue
                if cur == 0:
                    continue  # header "ngram N=count"
                parts = line.split()
                prob = _F(float(parts[0]))
                words = parts[1 : 1 + cur]
                backoff = _F(float(parts[1 + cur])) if len(parts) > 1 + cur else _F(0.0)
                ids = tuple(self._wid(w, create=True) for w in words)
                if cur == 1:
                    self._uni[ids[0]] = (prob, backoff)
                else:
                    self._ngrams[ids] = (prob, backoff)
                    self._ext.add(ids[:-1])
                if backoff != 0:
                    self._ext.add(ids)
        if 0 not in self._uni:
            self._uni[0] = (_F(-100.0), _F(0.0))

    def __contains__(self, word):
        return self._vocab.get(word, 0) != 0

    def BeginSentenceWrite(self, state):
        bos = self._vocab.get("<s>", 0)
        state.words = (bos,)
        state.backoffs = (self._uni[bos][1],)

    def NullContextWrite(self, state):
        state.words = ()
        state.backoffs = ()

    def BaseScore(self, in_state, word, out_state):
        wid = self._vocab.get(word, 0)
        prob, bo = self._uni[wid]
        out_bo = [bo]
        out_len = 1 if (wid,) in self._ext else 0
        matched = 1
        ctx = in_state.words
        gram = (wid,)
        for k in range(len(ctx)):
            if k + 2 > self.order:
                break
            gram = (ctx[k],) + gram
            ent = self._ngrams.get(gram)
            if ent is None:
                break
            prob = ent[0]
            matched = k + 2
            if matched < self.order:
                out_bo.append(ent[1])
                if gram in self._ext:
                    out_len = matched
        for i in range(matched - 1, len(ctx)):
            prob = _F(prob + in_state.backoffs[i])
        out_state.words = ((wid,) + tuple(ctx))[:out_len]
        out_state.backoffs = tuple(out_bo[:out_len])
        return float(prob)

    def score(self, sentence, bos=True, eos=True):
        st = State()
        if bos:
            self.BeginSentenceWrite(st)
        else:
            self.NullContextWrite(st)
        total = 0.0
        words = sentence.split()
        if eos:
            words = words + ["</s>"]
        for w in words:
            nxt = State()
            total += self.BaseScore(st, w, nxt)
            st = nxt
        return total


LanguageModel = Model
