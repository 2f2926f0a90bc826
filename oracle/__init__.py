"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's CTC / Gram-CTC loss (musyoku/chainer-speech-recognition,
``asr/loss/gram_ctc.py``) used to *check* the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import anything from here.  The product package (``chainer-speech-recognition_b200``) never does:
it fails loudly when its CUDA library is missing instead of falling back to this code.

Parity status (see DESIGN.md "Oracle"):
  * Gram-CTC: PINNED.  ``oracle.lattice`` / ``oracle/ctc_oracle.c`` are checked against the
    reference file itself, executed unmodified in the build container under a Chainer stub
    (``oracle/ref_stub.py``); the resulting input/output vectors are committed under
    ``tests/golden/`` together with the generating script.
  * plain CTC: the arithmetic lives in upstream Chainer (not vendored, version unpinned,
    not installable offline).  It is pinned through the in-repo derivative: gram_ctc.py with
    every bigram id = -1 *is* plain CTC (asr/loss/gram_ctc.py:95-98 disconnects those nodes),
    plus torch's CPU ``ctc_loss`` in float64 as an independent second opinion and a brute-force
    path enumerator on tiny cases.
"""
