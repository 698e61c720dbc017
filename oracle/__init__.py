"""CPU oracle for the LASR acoustic front end -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``lighting-asr_b200``) never imports it and has no CPU fallback.

What it restates (file:line of the reference it follows is cited in every function):

* ``kaldi_fbank``   -- torchaudio.compliance.kaldi.fbank (TA = torchaudio 2.11.0+cu128,
  ``torchaudio/compliance/kaldi.py``), the third-party library the reference's
  ``WavToKaldiFbank`` (``lasr/data/datatrans.py:42-104``) delegates all arithmetic to.
  torchaudio is NOT vendored under /root/reference and the reference pins no version;
  the oracle is pinned to the image's torchaudio 2.11.0.
* ``lasr_frontend`` -- ``VoiceNorm``/``WavToKaldiFbank``/``SpecAugment`` masks/``batch_list``
  (``lasr/data/datatrans.py``, ``lasr/utils/specaugment.py``, ``lasr/data/dataset.py:8-22``)
  and the fp64 Kaldi-style CMVN definition adopted in SURVEY.md §8(c).

Pinning: the reference has no tests and no golden vectors.  The oracle is
pinned by (1) fixtures under ``tests/golden/`` produced by ``oracle/gen_golden.py`` from the
UNMODIFIED reference modules imported from /root/reference together with the live
torchaudio, and (2) live comparison against ``torchaudio.compliance.kaldi.fbank`` wherever
torchaudio is importable (it is part of the image, also on the GPU box).
CMVN variance normalisation / global CMVN have no implementation in the reference:
for those rows the oracle is a definition, i.e. "parity unpinned" (DESIGN.md §3).
"""
