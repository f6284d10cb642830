#!/usr/bin/env python
"""Benchmark of the transcription hot path (BASELINE.json metric: RTFx = input audio-seconds per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] — kotoba-whisper-v2.0 architecture (32 enc / 2 dec, d=1280,
128 mels, random init), bf16, greedy short-form (`return_timestamps=False`, ja/transcribe, max_length 128), batch
64 x 30 s synthetic 16 kHz audio per GPU.  One step = log-mel -> encoder -> greedy generate over one batch, plus the
token-id gather across ranks.  `value` times it with the audio already resident in HBM; `e2e` times the public API
(WhisperFeatureExtractorB200.__call__ + model.generate) from host numpy clips to host token ids, copies included.

N > 1: launched by torchrun, one rank per GPU; every rank transcribes its own batch of 64 clips (weak scaling); the
only collective is the all_gather of token ids.  `--impl reference` times the reference's own implementation of the
path (transformers fp32 generate on the host CPU) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

KOTOBA = dict(vocab_size=51866, num_mel_bins=128, d_model=1280, encoder_layers=32, decoder_layers=2,
              encoder_attention_heads=20, decoder_attention_heads=20, encoder_ffn_dim=5120, decoder_ffn_dim=5120)
BATCH = 64
MAX_LENGTH = 128
SR = 16000
CLIP_S = 30
WORKLOAD = "kotoba-whisper-v2.0 arch (32enc/2dec d1280 128mel, random init) greedy short-form ja/transcribe " \
           "max_length 128, batch 64 x 30 s synthetic 16 kHz audio per GPU"


def synth_audio(batch: int, seed: int) -> np.ndarray:
    """The reference's own dummy-audio recipe (run_speed_eval.py:14-17) for half the clips, gaussian for the rest."""
    rng = np.random.default_rng(seed)
    out = np.empty((batch, SR * CLIP_S), np.float32)
    for i in range(batch):
        if i % 2 == 0:
            out[i] = (rng.random(SR * CLIP_S, dtype=np.float32) - 0.5) * 2 * 0.007
        else:
            out[i] = rng.standard_normal(SR * CLIP_S, dtype=np.float32) * 0.1
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.proc, self.lines, self.index, self.t_begin = None, [], index, 0.0

    def start(self):
        """Launch the poller.  Called BEFORE the warm-up: nvidia-smi takes ~1 s to initialise (it enumerates every GPU of
        the box under a driver lock), which must not land inside the timed region; only samples that arrive after
        `mark_begin()` are used."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)

            def reader():
                for l in self.proc.stdout:
                    self.lines.append((time.monotonic(), l))
            threading.Thread(target=reader, daemon=True).start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t_begin = time.monotonic()

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_line, l in self.lines:
            if t_line < self.t_begin:
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank: int, world: int):
    """transformers fp32 greedy generate on the host CPU — the reference's own code path (run_pseudo_labelling.py:268,338
    with the feature extractor's torch STFT) — on a bounded sample: `ref_batch` clips of the same workload per step."""
    if rank != 0:
        return
    from transformers import WhisperFeatureExtractor
    from _hf import build_hf
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_hf(KOTOBA, seed=0, dtype=torch.float32)
    fe = WhisperFeatureExtractor(feature_size=128)
    nb = args.ref_batch
    audio = synth_audio(nb, seed=1000)

    def step():
        feats = fe(list(audio), sampling_rate=SR, return_tensors="pt").input_features
        with torch.no_grad():
            return model.generate(feats, language="ja", task="transcribe", return_timestamps=False,
                                  max_length=MAX_LENGTH, num_beams=1)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    rtfx = nb * CLIP_S * args.steps / dt
    sample = f"{nb} of the {BATCH} clips per step (same arch, fp32, CPU), {args.steps} steps"
    line = {"impl": "reference", "metric": "RTFx", "value": rtfx, "unit": "audio_s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_sample": sample},
            "cpu_baseline": {"value": rtfx, "unit": "audio_s/s", "cores": threads, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": rtfx, "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(clips: int = 1):
    """Bounded CPU sample for the `cpu_baseline` object of our own line: HF fp32 on `clips` clips, one timed pass."""
    from transformers import WhisperFeatureExtractor
    from _hf import build_hf
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_hf(KOTOBA, seed=0, dtype=torch.float32)
    fe = WhisperFeatureExtractor(feature_size=128)
    audio = synth_audio(clips, seed=1000)
    t0 = time.perf_counter()
    feats = fe(list(audio), sampling_rate=SR, return_tensors="pt").input_features
    with torch.no_grad():
        model.generate(feats, language="ja", task="transcribe", return_timestamps=False, max_length=MAX_LENGTH, num_beams=1)
    dt = time.perf_counter() - t0
    return {"value": clips * CLIP_S / dt, "unit": "audio_s/s", "cores": threads, "kind": "reference",
            "sample": f"{clips} clip(s) x 30 s of the same workload through transformers fp32 generate, one pass "
                      f"({dt:.1f} s)"}


# ----------------------------------------------------------------------------------------------------------------------
def ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    from kotoba_whisper_b200 import (WhisperB200Config, WhisperB200ForConditionalGeneration,
                                     WhisperFeatureExtractorB200, _lib)
    from kotoba_whisper_b200.distributed import TokenGather
    from _hf import build_hf

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    cfg = WhisperB200Config(**KOTOBA)
    # ONE state_dict for both arms: HF's own CPU initialisation under torch.manual_seed(0), shipped to
    # the GPU — the reference arm builds the identical model (reference_arm -> build_hf(KOTOBA, seed=0)).
    sd = {k: v.detach() for k, v in build_hf(KOTOBA, seed=0).state_dict().items()}
    co = 1 if args.no_stream else max(1, args.coalesce)   # submitted 64-clip batches per device batch (GenerateStream)
    model = WhisperB200ForConditionalGeneration.from_state_dict(sd, cfg, dtype=torch.bfloat16, max_batch=BATCH * co, device=dev)
    torch.cuda.empty_cache()
    fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
    audio_host = synth_audio(BATCH, seed=1000 * 1 + rank)       # seed = 1000*config + rank
    audio_dev = torch.from_numpy(audio_host).to(dev)
    clips_host = list(audio_host)
    pad = model.generation_config.pad_token_id
    stats = {}
    # the only collective: ONE fixed-shape all_gather of [64, max_length - prompt] int32 per batch, enqueued
    # asynchronously and read one batch later (no size exchange, no host sync between ranks inside a step)
    gather = TokenGather(BATCH, MAX_LENGTH - 4, pad)
    inflight = {"h": None}

    gen_kw = dict(language="ja", task="transcribe", return_timestamps=False, max_length=MAX_LENGTH)

    def step_plain():
        feats = fe.logmel_device(audio_dev)
        ids = model.generate(feats, stats=stats, **gen_kw)
        prev, inflight["h"] = inflight["h"], gather.submit(ids)
        return (prev or inflight["h"]).result()

    # Two device batches in flight (GenerateStream): a step featurises one 64-clip batch and submits it; every `co`-th
    # submit launches the encoder of the `co` buffered batches as ONE device batch, with the decoder positions of the
    # previous device batch slotted between the layer groups, and strips / gathers finished ids one 64-clip batch per
    # step.  Over any `co` consecutive steps the GPU does `co` log-mels, encoder work for `co` x 64 clips, one greedy
    # pass over `co` x 64 rows and `co` gathers; warm-up primes the stream, what is left in flight after the timed
    # steps is flushed outside (its decode replaces the one the first timed launch did for the last warm-up batches).
    stream = model.generate_stream(coalesce=co, stats=stats, **gen_kw)

    def step_stream():
        feats = fe.logmel_device(audio_dev)
        ids = stream.submit(feats)
        if ids is None:
            return None
        prev, inflight["h"] = inflight["h"], gather.submit(ids)
        return (prev or inflight["h"]).result()

    step_resident = step_plain if args.no_stream else step_stream

    def step_e2e():
        feats = fe(clips_host, sampling_rate=SR, return_tensors="pt", keep_on_device=True)["input_features"]
        ids = model.generate(feats, language="ja", task="transcribe", return_timestamps=False, max_length=MAX_LENGTH)
        return gather.submit(ids).result().cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    if not args.no_stream:
        # steady state before the timed region: device batches in flight (so the first timed launch has a pass to
        # interleave and finished ids to hand back) and a buffer fill such that the K timed steps launch ceil(K / co) device batches — never fewer
        # encoder / decoder rows than K x 64 (an odd K at co = 2 does the work of K + 1 batches: counted against us)
        while stream.device_batches < 3 or (stream.buffered + args.steps) % co != 0:
            step_resident()
    # Timed region: only the dominant kernel category (encoder GEMMs) carries CUDA-event pairs; the other categories
    # are timed the same way in two extra, untimed steps right after (event pairs between the decode kernels would
    # break the programmatic-dependent-launch overlap the timed region is supposed to measure).
    for c in range(6):
        _lib.profile_read(c, reset=True)
    lib.kw_profile_enable(1 << _lib.PROF_ENC_GEMM)
    lib.kw_launch_count(1)
    sampler.mark_begin()
    ms, ids = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.kw_launch_count(0))
    prof = {_lib.PROF_ENC_GEMM: _lib.profile_read(_lib.PROF_ENC_GEMM, reset=True)}
    extra_steps = 2
    lib.kw_profile_enable((1 << _lib.PROF_ENC_ATTN) | (1 << _lib.PROF_DEC_CROSS) | (1 << _lib.PROF_LOGMEL))
    ms_extra, _ = timed(step_resident, extra_steps)
    for c in (_lib.PROF_ENC_ATTN, _lib.PROF_DEC_CROSS, _lib.PROF_LOGMEL):
        prof[c] = _lib.profile_read(c, reset=True)
    while (last := stream.flush()) is not None:
        gather.submit(last).result()
    # whole decode pass (all kernels of all positions) under ONE event pair, so programmatic launch chains stay intact;
    # measured batch by batch (plain generate): inside a stream step the pass is interleaved with the next encoder
    step_plain()
    lib.kw_profile_enable(1 << _lib.PROF_DEC_PASS)
    ms_pass_steps, _ = timed(step_plain, extra_steps)
    prof[_lib.PROF_DEC_PASS] = _lib.profile_read(_lib.PROF_DEC_PASS, reset=True)
    prof_co = None
    if co > 1:  # the coalesced pass alone: co x 64 rows through plain generate (same kernels the stream launches)
        feats_co = torch.cat([fe.logmel_device(audio_dev) for _ in range(co)])
        model.generate(feats_co, **gen_kw)
        _lib.profile_read(_lib.PROF_DEC_PASS, reset=True)
        timed(lambda: model.generate(feats_co, **gen_kw), extra_steps)
        prof_co = _lib.profile_read(_lib.PROF_DEC_PASS, reset=True)
        del feats_co
    lib.kw_profile_enable(0)
    passes = stats.get("passes", 0)
    ms_plain, _ = timed(step_plain, args.steps)   # same work batch by batch, for comparison with the stream
    ms_stream1 = None
    if co > 1:  # and the stream schedule without coalescing (one submitted batch per device batch), for the same comparison
        stream1 = model.generate_stream(coalesce=1, **gen_kw)

        def step_stream1():
            ids1 = stream1.submit(fe.logmel_device(audio_dev))
            if ids1 is not None:
                prev, inflight["h"] = inflight["h"], gather.submit(ids1)
                (prev or inflight["h"]).result()

        for _ in range(4):
            step_stream1()
        ms_stream1, _ = timed(step_stream1, args.steps)
        while (last := stream1.flush()) is not None:
            gather.submit(last).result()

    def step_encoder_only():  # log-mel + encoder of one batch, no decode: what a stream step spends outside the decoder
        model.encode(fe.logmel_device(audio_dev), return_hidden=False)

    step_encoder_only()
    ms_enc_only, _ = timed(step_encoder_only, args.steps)

    step_e2e()
    ms_e2e_serial, ids_host = timed(step_e2e, args.steps)

    # Headline e2e: the same public calls with the feature extractor's prefetch handle, the way the reference's
    # DataLoader workers prepare batch i+1 while the model labels batch i.  Every step's host staging, H2D copy and
    # D2H token read happen inside the timed region; only their overlap with the previous step's kernels differs.
    def run_e2e_pipelined(steps):
        out, h = None, None
        pending = fe.prefetch(clips_host, sampling_rate=SR, return_tensors="pt")
        for i in range(steps + 1):
            if i < steps:
                feats = pending.result()["input_features"]
                if i + 1 < steps:
                    pending = fe.prefetch(clips_host, sampling_rate=SR, return_tensors="pt")
                ids = model.generate(feats, **gen_kw) if args.no_stream else stream.submit(feats)
                tail = [ids] if ids is not None else []
            else:  # the stream starts and ends empty inside the timed region
                tail = []
                while not args.no_stream and (ids := stream.flush()) is not None:
                    tail.append(ids)
            for ids in tail:
                prev, h = h, gather.submit(ids)
                if prev is not None:
                    out = prev.result().cpu()  # batch i-1's gathered ids -> host while batch i's gather is in flight
        return h.result().cpu()

    run_e2e_pipelined(2)
    ms_e2e, ids_host = timed(lambda: run_e2e_pipelined(args.steps), 1)

    audio_s = BATCH * CLIP_S * world
    value = audio_s * args.steps / (ms / 1e3)
    e2e = audio_s * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        return
    hbm, tf_sus, tf_burst, peak_src = load_peaks()
    NOMINAL_TF, NOMINAL_GBS = 2250.0, 8000.0  # fractions are reported against nominal peaks too

    def leg(cat, unit_scale, total_ms):
        t, n, work = prof[cat]
        return (work / unit_scale) / (t / 1e3) if t > 0 else 0.0, n, t / total_ms

    g_tf, g_n, g_share = leg(_lib.PROF_ENC_GEMM, 1e12, ms)
    a_tf, a_n, a_share = leg(_lib.PROF_ENC_ATTN, 1e12, ms_extra)
    x_gb, x_n, x_share = leg(_lib.PROF_DEC_CROSS, 1e9, ms_extra)
    m_gb, m_n, m_share = leg(_lib.PROF_LOGMEL, 1e9, ms_extra)
    p_gb, p_n, p_share = leg(_lib.PROF_DEC_PASS, 1e9, ms_pass_steps)
    # ncu-measured DRAM traffic of the dominant kernel category (profiles/r2_traffic.json, written by
    # tools/ncu_traffic.py from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass of this command)
    traffic, traffic_note = None, "no profiles/r2_traffic.json"
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("encoder_gemm_launches"):
            traffic = tj["encoder_gemm_dram_bytes"] / tj["encoder_gemm_launches"]
            traffic_note = f"ncu dram bytes per launch averaged over {tj['encoder_gemm_launches']} encoder GEMM launches; " \
                           f"algorithmic operand bytes per launch {tj.get('encoder_gemm_algorithmic_bytes', 0) / tj['encoder_gemm_launches']:.3e}"
    pass_t, pass_n, pass_work = prof[_lib.PROF_DEC_PASS]
    positions = (MAX_LENGTH - 1) * max(pass_n, 1)
    roofline = {"kernel": "encoder GEMMs (conv stem, QKV, out, fc1, fc2)", "bound": "tensor", "achieved": g_tf,
                "peak": tf_sus, "unit": "TFLOP/s", "frac": g_tf / tf_sus, "frac_burst": g_tf / tf_burst,
                "frac_nominal": g_tf / NOMINAL_TF, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src, "launches": g_n, "share_of_step": g_share}
    # the coalesced pass (co x 64 rows per decoder position): weights and the vocabulary matrix are streamed once for all
    # rows, so its algorithmic bytes per position are NOT co x the 64-row figure (kw_greedy_pass counts them per pass)
    co_t, co_n, co_work = prof_co if prof_co else (pass_t, pass_n, pass_work)
    co_positions = (MAX_LENGTH - 1) * max(co_n, 1)
    co_gb = (co_work / 1e9) / (co_t / 1e3) if co_t > 0 else 0.0
    dec_ms_per_step = max((ms - ms_enc_only) / args.steps, 1e-6)      # decode share of one 64-clip stream step
    dec_work_per_step = co_work / max(co_n, 1) / co                    # algorithmic bytes of that share
    extra = [
        {"kernel": "encoder self-attention", "bound": "tensor", "achieved": a_tf, "peak": tf_sus, "unit": "TFLOP/s",
         "frac": a_tf / tf_sus, "frac_nominal": a_tf / NOMINAL_TF, "launches": a_n, "share_of_step": a_share},
        {"kernel": "decode step, 64 rows (all kernels of a decoder position, whole greedy pass under one event pair, "
                   "batch by batch)",
         "bound": "hbm", "achieved": p_gb, "peak": hbm, "unit": "GB/s", "frac": p_gb / hbm,
         "frac_nominal": p_gb / NOMINAL_GBS, "passes": pass_n, "share_of_step": p_share,
         "ms_per_position": pass_t / positions if positions else None,
         "algorithmic_bytes_per_position": pass_work / positions if positions else None},
        {"kernel": f"decode step, {co * BATCH} rows (the coalesced pass the stream schedule runs: {co} submitted batches "
                   "per device batch; whole greedy pass alone under one event pair)",
         "bound": "hbm", "achieved": co_gb, "peak": hbm, "unit": "GB/s", "frac": co_gb / hbm,
         "frac_nominal": co_gb / NOMINAL_GBS, "passes": co_n,
         "ms_per_position": co_t / co_positions if co_positions else None,
         "ms_per_pass_per_64_clips": co_t / max(co_n, 1) / co,
         "algorithmic_bytes_per_position": co_work / co_positions if co_positions else None} if prof_co else None,
        {"kernel": "decode step inside the stream schedule (DERIVED: timed stream step minus a timed log-mel + encoder-only "
                   "step = the decode share of one 64-clip step, against the algorithmic bytes of the coalesced pass "
                   "divided by the batches it covers; includes the host-side strip / gather of the step)",
         "bound": "hbm", "unit": "GB/s", "peak": hbm,
         "achieved": dec_work_per_step / 1e9 / (dec_ms_per_step / 1e3),
         "frac": dec_work_per_step / 1e9 / (dec_ms_per_step / 1e3) / hbm,
         "ms_decode_share_per_step": dec_ms_per_step,
         "ms_per_position": dec_ms_per_step * co / (MAX_LENGTH - 1),
         "ms_encoder_only_step": ms_enc_only / args.steps} if not args.no_stream else None,
        {"kernel": "decode-step cross-attention", "bound": "hbm", "achieved": x_gb, "peak": hbm, "unit": "GB/s",
         "frac": x_gb / hbm, "frac_nominal": x_gb / NOMINAL_GBS, "launches": x_n, "share_of_step": x_share},
        {"kernel": "log-mel (stft+mel+log, incl. fix-up pass)", "bound": "hbm", "achieved": m_gb, "peak": hbm,
         "unit": "GB/s", "frac": m_gb / hbm, "frac_nominal": m_gb / NOMINAL_GBS, "launches": m_n, "share_of_step": m_share},
    ]
    extra = [e for e in extra if e is not None]
    parity = None
    if not args.no_parity:
        # the benched run parity-checked in place: the same 64 clips and the same weights through the exact-fp32 CUDA path
        try:
            from kotoba_whisper_b200.parity import bf16_token_parity, rounded_state_dict
            m32 = WhisperB200ForConditionalGeneration.from_state_dict(rounded_state_dict(sd), cfg, dtype=torch.float32,
                                                                      max_batch=16, device=dev)
            parity = bf16_token_parity(model, m32, fe.logmel_device(audio_dev), max_length=MAX_LENGTH)
            for k in ("first_divergence_step", "fp32_gap_at_divergence"):
                parity.pop(k, None)
            parity["bar"] = "bf16: token sequences identical on >= 99 % of utterances (near-ties below tau count as identical)"
            parity["meets_bar"] = parity["adjusted_pct"] >= 99.0
            del m32
        except Exception as e:
            parity = {"error": f"{type(e).__name__}: {e}"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_leg(1)
        except Exception as e:  # the CPU leg is a reported baseline, never a reason to lose the GPU number
            cpu = {"value": None, "unit": "audio_s/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"failed: {type(e).__name__}: {e}"}
    line = {"metric": "RTFx", "value": value, "unit": "audio_s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                       "passes_per_step": passes,
                       "schedule": "batch by batch" if args.no_stream else
                       f"GenerateStream(coalesce={co}): every step submits one 64-clip batch; {co} submitted batches run as "
                       f"one {co * BATCH}-row device batch, 2 device batches in flight (encoder of device batch i+1 "
                       "slotted between the decoder positions of device batch i on one stream); per 64-clip step: 1 "
                       f"log-mel + encoder work for 64 clips + 1/{co} of a {co * BATCH}-row greedy pass + 1 token gather",
                       "coalesce": co, "device_batch": co * BATCH,
                       "ms_per_step_batch_by_batch": ms_plain / args.steps,
                       "ms_per_step_stream_coalesce1": ms_stream1 / args.steps if ms_stream1 else None, "l2": "inputs_larger_than_l2 (123 MB audio + 1.5 GB weights per step)",
                       "tokens_out_shape": list(ids.shape)},
            "roofline": roofline, "roofline_extra": extra, "roofline_extra_note": f"timed with CUDA events in {extra_steps} extra steps after the timed region", "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "audio_s/s", "ms_per_step": ms_e2e / args.steps,
                    "mode": "prefetch: step i+1 staged + copied on a side stream while step i runs",
                    "serial_ms_per_step": ms_e2e_serial / args.steps,
                    "serial_value": audio_s * args.steps / (ms_e2e_serial / 1e3),
                    "h2d_bytes_per_step": int(audio_host.nbytes) * world,
                    "d2h_bytes_per_step": int(ids_host.numel() * ids_host.element_size())},
            "gpu_launches": launches, "clocks": clocks, "parity": parity}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=1, help="clips per step of the bounded CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="model.generate batch by batch instead of GenerateStream")
    ap.add_argument("--coalesce", type=int, default=2,
                    help="submitted 64-clip batches run as one device batch by GenerateStream (1 = the round-2 schedule)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-place bf16-vs-fp32 token parity check")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
