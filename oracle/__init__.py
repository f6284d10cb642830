"""CPU oracle for the kotoba-whisper transcription hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.  The product
package (``kotoba_whisper_b200``) never imports this module and fails loudly
when its CUDA library is missing.

What is restated here (numpy float64 / plain torch-CPU fp32), with the reference
lines each function follows (``HF`` = transformers 5.5.0, the third-party package
in which the reference's hot path lives; the reference repo only calls it, from
``run_pseudo_labelling.py:268,338`` and ``run_speed_eval.py:76``):

* ``logmel_ref``  – WhisperFeatureExtractor log-mel
  (HF/models/whisper/feature_extraction_whisper.py:135-164, HF/audio_utils.py:453-544)
* ``whisper_ref`` – encoder, decoder step, logits processors, greedy loop and the
  seek/segment state machine of ``WhisperGenerationMixin.generate``
  (HF/models/whisper/modeling_whisper.py:593-647,734-796,449-506,1081;
   HF/generation/logits_process.py:1855-2043; HF/models/whisper/generation_whisper.py:649-968,1976-2073)
* ``chunking_ref`` – ASR-pipeline chunk iterator and token longest-common-sequence merge
  (HF/pipelines/automatic_speech_recognition.py:61-84; HF/models/whisper/tokenization_whisper.py:1153-1270)

Pinning: the reference ships no tests for this path, so the oracle is
pinned against outputs of the reference's own dependency run in the build container:
``tests/golden/make_golden.py`` imports transformers, runs it on seeded synthetic inputs and
random-init weights, and commits small fixtures; ``tests/test_oracle.py`` checks the
restatement against those fixtures and (when transformers is importable) against HF live.
"""
