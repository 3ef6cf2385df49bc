"""CPU oracle for the inaSpeechSegmenter hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy / torch-CPU / a little C), the
arithmetic of the reference's per-frame hot path so the CUDA kernels in
``inaspeechsegmenter_b200`` can be checked against it.  It is NOT part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``inaspeechsegmenter_b200``) never imports ``oracle`` and fails loudly when
its CUDA library is missing.

Parity status (see DESIGN.md section "Oracle"):

* ``sidekit_oracle``  -- log-mel + log-energy front-end.  PINNED: checked
  against the real reference module (``/root/reference/inaSpeechSegmenter/
  sidekit_mfcc.py`` imported by file path, ``tests/golden/make_golden.py``)
  bit-for-bit on the media fixtures, and against the ``noEnergy`` rows of the
  reference's golden CSVs.
* ``viterbi_oracle`` (+ ``viterbi_oracle.c``) -- PINNED against the real
  ``pyannote_viterbi.viterbi_decoding`` on seeded inputs (golden .npz).
* ``segmenter_oracle`` -- glue of ``segmenter.py`` (cannot be imported as a
  module here: it pulls TensorFlow); PINNED weights-free through the golden CSV
  ``noEnergy`` rows and silence fixture.
* ``cnn_oracle`` -- Keras layer semantics in torch-CPU fp32.  PARITY UNPINNED:
  the three ``.hdf5`` models and TensorFlow are absent from this container, no
  reference fixture stores per-frame probabilities.
* ``vbx_oracle`` -- VBx front-end PINNED against the real ``features_vbx.py``;
  ResNet101 forward restated from ``resnet.py`` (architecture) -- PARITY
  UNPINNED for real weights (``final.onnx`` absent).
"""
