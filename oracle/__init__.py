"""CPU oracle for the waveform->token encode path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU / numpy, fp32 with optional fp64) of the
algorithm of the reference's hot path, written from the reference's behaviour — every
function cites the reference file:line it follows.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it.  Nothing under ``audiotoken_b200/`` imports it, and the product path raises if the
CUDA library is missing — there is no CPU fallback.

Parity status: **pinned against the reference itself, run in the build container**.  The
reference's own tests hold no golden vectors for this path, so
``tests/golden/make_golden.py`` loads the reference's ``audiotoken/processors.py`` and
``audiotoken/modeling_wav2vec2_bert.py`` from ``/root/reference`` in isolation, runs them
(with HF ``Wav2Vec2BertModel`` / ``EncodecModel`` as the reference does) on seeded inputs and
weights, and commits the outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks this oracle against those fixtures on every CPU test run.  The two third-party
quantisers that are absent from ``/root/reference`` (``vector_quantize_pytorch`` — unpinned in
requirements.txt:10 — and ``encodec`` — unpinned, requirements.txt:5) are restated from their
published algorithm (nearest codeword in Euclidean distance; residual VQ) and anchored on the
reference's call sites (encoder.py:50-52, 100-101, 147-161, 180).

``oracle/hubert.py`` (the reference's own mHuBERT ``semantic_s``) is pinned the same way against HF
``HubertModel`` / ``Wav2Vec2FeatureExtractor`` (``tests/golden/make_golden_hubert.py``); it has no CUDA counterpart yet.
``oracle/quantize.py::vq_ema_train_step`` (codebook training) restates the third-party package's published training
forward and is **parity unpinned** against the package itself.
"""
