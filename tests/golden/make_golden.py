"""Generate the golden fixtures under tests/golden/ by running the REFERENCE implementation
(transformers' Whisper — the code the reference repo calls at run_pseudo_labelling.py:268,338) on CPU fp32.

    python tests/golden/make_golden.py [logmel] [tiny] [tiny_extra] [kotoba] [teacher] [teacher_tf]

Inputs are seeded synthetic audio (tests/_synth.py) and random-init weights (torch.manual_seed(0), tests/_hf.py), so
every consumer can rebuild bit-identical inputs; only the (small) outputs are committed.  Run in the build container;
the fixtures travel to the GPU box with the repo.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _hf import build_hf  # noqa: E402
from _synth import KOTOBA, TEACHER, TINY, TINY80, clips  # noqa: E402
from oracle.logmel_ref import logmel_batch_f64, pad_or_trim  # noqa: E402


def golden_logmel():
    from transformers import WhisperFeatureExtractor
    out = {}
    for nm in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=nm)
        cl = clips("UGSS", 500 + nm)
        cl.append(np.zeros(1000, np.float32))  # all-silence clip: every value is the -1.5 floor
        ht = fe(cl, sampling_rate=16000, return_tensors="np", return_attention_mask=True)
        hn = fe._np_extract_fbank_features(np.stack([pad_or_trim(c, 480000) for c in cl]), "cpu")
        # keep every 37th frame (82 of 3000) of every mel row + the full first/last 8 frames
        out[f"hf_torch_{nm}"] = ht.input_features[:, :, ::37].astype(np.float32)
        out[f"hf_numpy_{nm}"] = hn[:, :, ::37].astype(np.float32)
        out[f"hf_numpy_head_{nm}"] = hn[:, :, :8].astype(np.float32)
        out[f"hf_numpy_tail_{nm}"] = hn[:, :, -8:].astype(np.float32)
        out[f"mask_sum_{nm}"] = ht.attention_mask.sum(-1).astype(np.int64)
        out[f"clip_len_{nm}"] = np.array([len(c) for c in cl], np.int64)
        out[f"hf_numpy_rowsum_{nm}"] = hn.astype(np.float64).sum(-1)  # [B, n_mels] checksum over all frames
    np.savez_compressed(os.path.join(HERE, "logmel.npz"), **out)
    print("logmel.npz written")


def _gen_cases(model, mel, tag, out, cases):
    with torch.no_grad():
        enc = model.model.encoder(mel).last_hidden_state
        out[f"{tag}_enc_sub"] = enc[:, ::97, ::5].numpy().astype(np.float32)
        out[f"{tag}_enc_absmean"] = enc.abs().mean(dim=(1, 2)).numpy()
        for ts, ml, lang, task in cases:
            t0 = time.time()
            ids = model.generate(mel, language=lang, task=task, return_timestamps=ts, max_length=ml, num_beams=1)
            key = f"{tag}_ids_ts{int(ts)}_ml{ml}_{lang}_{task}"
            out[key] = ids.numpy().astype(np.int64)
            print(key, tuple(ids.shape), f"{time.time() - t0:.1f}s", flush=True)
        # raw logits of the first generated position, from a plain forward pass over the 4-token no-timestamp prompt
        # (generate()'s own `logits` output belongs to the LAST seek pass, which differs when >1 pass ran)
        prompt = torch.tensor([[50258, 50266, 50360, 50364]] * mel.shape[0])
        lg = model(input_features=mel, decoder_input_ids=prompt).logits
        out[f"{tag}_logits0_sub"] = lg[:, -1, ::53].float().numpy().astype(np.float32)
        top2 = lg[:, -1].float().topk(2, dim=-1).values
        out[f"{tag}_logits0_margin"] = (top2[:, 0] - top2[:, 1]).numpy().astype(np.float32)


def golden_tiny():
    out = {}
    for name, arch in (("tiny", TINY), ("tiny80", TINY80)):
        model = build_hf(arch)
        mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), arch["num_mel_bins"]))
        _gen_cases(model, mel, name, out,
                   [(True, 40, "ja", "transcribe"), (True, 128, "ja", "transcribe"), (False, 40, "ja", "transcribe"),
                    (False, 128, "ja", "transcribe"), (True, 64, "en", "translate")])
    np.savez_compressed(os.path.join(HERE, "tiny.npz"), **out)
    print("tiny.npz written")


def longform_batch(n_mels=128):
    """3 recordings of 70 / 45 / 33 s, zero-padded to the longest, featurised as one batch + frame attention mask
    (what WhisperFeatureExtractor(..., padding="longest", return_attention_mask=True, truncation=False) produces)."""
    from oracle.logmel_ref import logmel_f64
    rng = np.random.default_rng(21)
    secs = [70, 45, 33]
    n_max = 16000 * max(secs)
    audio = [(rng.standard_normal(16000 * s) * 0.1).astype(np.float32) for s in secs]
    mel = np.stack([logmel_f64(np.pad(a, (0, n_max - len(a))), n_mels, n_samples=n_max) for a in audio])
    mask = np.zeros((3, n_max // 160), np.int64)
    for i, s in enumerate(secs):
        mask[i, : 16000 * s // 160] = 1
    return torch.from_numpy(mel), torch.from_numpy(mask)


def golden_tiny_extra():
    """generate(encoder_outputs=...) with timestamps on / off (HF keeps its seek loop running on the same encoder
    output), batched long-form with attention_mask (batch shrinking), and teacher-forcing logits (labels=...)."""
    from transformers.modeling_outputs import BaseModelOutput
    out = {}
    model = build_hf(TINY)
    mel = torch.from_numpy(logmel_batch_f64(clips("UGS", 7), 128))
    with torch.no_grad():
        enc = model.model.encoder(mel).last_hidden_state
        ids = model.generate(encoder_outputs=BaseModelOutput(last_hidden_state=enc), language="ja",
                             task="transcribe", return_timestamps=False, max_length=40, num_beams=1)
        out["encout_ids_ts0"] = ids.numpy().astype(np.int64)
        # with timestamps HF keeps seeking over the SAME encoder output (generation_whisper.py:785-903 with
        # input_features=None); for B > 1 its batch shrinking crashes on the missing input_features (:1823) as soon as
        # one row finishes early, so the defined behaviour is per row
        for b in range(3):
            ids = model.generate(encoder_outputs=BaseModelOutput(last_hidden_state=enc[b:b + 1]), language="ja",
                                 task="transcribe", return_timestamps=True, max_length=40, num_beams=1)
            out[f"encout_ids_ts1_row{b}"] = ids.numpy().astype(np.int64)
            print("encoder_outputs ts row", b, tuple(ids.shape))
        lmel, lmask = longform_batch()
        ids = model.generate(lmel, attention_mask=lmask, language="ja", task="transcribe", return_timestamps=True,
                             max_length=64, num_beams=1)
        out["longform_b3_ids"] = ids.numpy().astype(np.int64)
        print("long-form B=3", tuple(ids.shape))
        # teacher forcing: labels [3, 24] with -100 padding on the last rows
        g = torch.Generator().manual_seed(5)
        labels = torch.randint(0, 50257, (3, 24), generator=g)
        labels[1, 18:] = -100
        labels[2, 9:] = -100
        res = model(input_features=mel, labels=labels)
        out["tf_labels"] = labels.numpy().astype(np.int64)
        out["tf_logits_sub"] = res.logits[:, :, ::53].numpy().astype(np.float32)
        out["tf_loss"] = np.array(float(res.loss))
        res2 = model(encoder_outputs=BaseModelOutput(last_hidden_state=enc), decoder_input_ids=labels.clamp(min=0))
        out["tf_logits_ids_sub"] = res2.logits[:, :, ::53].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "tiny_extra.npz"), **out)
    print("tiny_extra.npz written")


def golden_teacher_tf():
    """Teacher-forcing logits at T = 128 (max_label_length of the reference's distillation configs) for the kotoba and
    teacher architectures: model(input_features, labels) as run_distillation.py:641-649 calls the frozen teacher."""
    out = {}
    g = torch.Generator().manual_seed(11)
    labels = torch.randint(0, 50257, (2, 128), generator=g)
    labels[1, 90:] = -100
    out["labels"] = labels.numpy().astype(np.int64)
    mel = torch.from_numpy(logmel_batch_f64(clips("GS", 3000), 128))
    for name, arch in (("kotoba", KOTOBA), ("teacher", TEACHER)):
        t0 = time.time()
        model = build_hf(arch)
        with torch.no_grad():
            res = model(input_features=mel, labels=labels)
        out[f"{name}_logits_sub"] = res.logits[:, :, ::212].numpy().astype(np.float32)
        out[f"{name}_loss"] = np.array(float(res.loss))
        print(name, "teacher-forcing", tuple(res.logits.shape), f"{time.time() - t0:.1f}s", flush=True)
        del model
    np.savez_compressed(os.path.join(HERE, "teacher_tf.npz"), **out)
    print("teacher_tf.npz written")


def golden_full(name, arch, spec, seed0, cases):
    out = {}
    t0 = time.time()
    model = build_hf(arch)
    print(name, "built", f"{time.time() - t0:.1f}s", flush=True)
    mel = torch.from_numpy(logmel_batch_f64(clips(spec, seed0), arch["num_mel_bins"]))
    _gen_cases(model, mel, name, out, cases)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}.npz written", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["logmel", "tiny", "kotoba"]
    torch.set_num_threads(os.cpu_count())
    if "logmel" in what:
        golden_logmel()
    if "tiny" in what:
        golden_tiny()
    if "tiny_extra" in what:
        golden_tiny_extra()
    if "teacher_tf" in what:
        golden_teacher_tf()
    if "kotoba" in what:  # BASELINE.json configs[0]: batch 4 x 30 s, ja/transcribe, timestamps, max_length 128
        golden_full("kotoba", KOTOBA, "UGSG", 1000, [(True, 128, "ja", "transcribe"), (False, 128, "ja", "transcribe")])
    if "teacher" in what:  # configs[2] architecture at a CPU-runnable batch of 2
        golden_full("teacher", TEACHER, "GS", 3000, [(True, 128, "ja", "transcribe")])
