"""Drop-in for `WhisperForConditionalGeneration` on the transcription path: `generate(input_features, language, task,
return_timestamps, max_length, ...)` with the reference's semantics (HF/models/whisper/generation_whisper.py:383-968),
driving the CUDA kernels through libkwb200.so.

What lives where:
  * device arithmetic (encoder, cross K/V, decoder steps, logits processors, argmax): C ABI, one call per greedy pass;
  * host state machine restated here: prompt construction (:1455-1608), 30 s seek loop (:785-903), per-pass pad / eos
    stripping (:1063-1086), timestamp segment slicing (:1976-2073), right padding of the result (:126-237).
Beam search, temperature fallback, prompt_ids / previous-text conditioning and token-level timestamps are not on the
reference's path (run_pseudo_labelling.py:307-313 uses num_beams=1, greedy) and raise NotImplementedError.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

# id = 50259 + index (multilingual v3 vocabulary; order of HF tokenization_whisper.LANGUAGES)
LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy "
    "sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo "
    "uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue"
).split()
LANGUAGE_NAMES = {"english": "en", "japanese": "ja", "chinese": "zh", "german": "de", "spanish": "es", "russian": "ru",
                  "korean": "ko", "french": "fr", "portuguese": "pt", "turkish": "tr", "italian": "it"}
TASK_IDS = ("translate", "transcribe")


@dataclass
class WhisperB200Config:
    """The WhisperConfig fields the path reads."""
    vocab_size: int = 51866
    num_mel_bins: int = 128
    d_model: int = 1280
    encoder_layers: int = 32
    decoder_layers: int = 2
    encoder_attention_heads: int = 20
    decoder_attention_heads: int = 20
    encoder_ffn_dim: int = 5120
    decoder_ffn_dim: int = 5120
    max_source_positions: int = 1500
    max_target_positions: int = 448
    decoder_start_token_id: int = 50258
    eos_token_id: int = 50257
    bos_token_id: int = 50257
    pad_token_id: int = 50256

    @classmethod
    def from_any(cls, cfg) -> "WhisperB200Config":
        if isinstance(cfg, cls):
            return cfg
        get = (lambda k, d: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d: getattr(cfg, k, d))
        base = cls()
        return cls(**{k: get(k, getattr(base, k)) for k in base.__dataclass_fields__})


_V3_SUPPRESS = [
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647,
    7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
    18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870,
    50254, 50258, 50359, 50360, 50361, 50362, 50363,
]


@dataclass
class WhisperB200GenerationConfig:
    """GenerationConfig fields the path reads; defaults = public whisper-large-v3 values."""
    decoder_start_token_id: int = 50258
    eos_token_id: int = 50257
    pad_token_id: int = 50257
    bos_token_id: int = 50257
    no_timestamps_token_id: int = 50364
    prev_sot_token_id: int = 50362
    max_initial_timestamp_index: Optional[int] = 50
    max_length: int = 448
    is_multilingual: bool = True
    return_timestamps: Optional[bool] = None
    begin_suppress_tokens: Optional[Sequence[int]] = (220, 50257)
    suppress_tokens: Optional[Sequence[int]] = tuple(_V3_SUPPRESS)
    lang_to_id: Dict[str, int] = field(default_factory=lambda: {f"<|{c}|>": 50259 + i for i, c in enumerate(LANGUAGE_CODES)})
    task_to_id: Dict[str, int] = field(default_factory=lambda: {"translate": 50359, "transcribe": 50360})

    @classmethod
    def from_any(cls, g) -> "WhisperB200GenerationConfig":
        if g is None:
            return cls()
        if isinstance(g, cls):
            return g
        base = cls()
        get = (lambda k, d: g.get(k, d)) if isinstance(g, dict) else (lambda k, d: getattr(g, k, d))
        vals = {k: get(k, getattr(base, k)) for k in base.__dataclass_fields__}
        if isinstance(vals["eos_token_id"], (list, tuple)):
            vals["eos_token_id"] = vals["eos_token_id"][0]
        return cls(**vals)


@dataclass
class EncoderOutput:
    last_hidden_state: torch.Tensor

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


@dataclass
class Seq2SeqOutput:
    """The Seq2SeqLMOutput fields the distillation step reads (run_distillation.py:641-652)."""
    loss: Optional[torch.Tensor]
    logits: torch.Tensor
    encoder_last_hidden_state: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        return getattr(self, k) if isinstance(k, str) else tuple(v for v in (self.loss, self.logits) if v is not None)[k]


def _dev_tensor(t: torch.Tensor, dtype: torch.dtype, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=dtype).contiguous()


class WhisperB200ForConditionalGeneration:
    """`model.generate(...)` drop-in.  Build with `from_state_dict` (HF names) or `from_hf_model`."""

    def __init__(self, config: WhisperB200Config, generation_config: WhisperB200GenerationConfig,
                 state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.bfloat16, max_batch: int = 64,
                 device: Union[str, torch.device] = "cuda"):
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be torch.float32 (exact mode) or torch.bfloat16")
        self.config = config
        self.generation_config = generation_config
        self.dtype = dtype
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.KwError("WhisperB200ForConditionalGeneration needs a CUDA device (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_batch = int(max_batch)
        self._lib = _lib.load()
        self._keep: List[torch.Tensor] = []
        self._handle = C.c_void_p()
        self._extra_handles: List[C.c_void_p] = []
        self._build(state_dict)

    # ---- construction ----------------------------------------------------------------------------------------------
    @classmethod
    def from_state_dict(cls, state_dict, config, generation_config=None, dtype=torch.bfloat16, max_batch=64,
                        device="cuda"):
        return cls(WhisperB200Config.from_any(config), WhisperB200GenerationConfig.from_any(generation_config),
                   state_dict, dtype=dtype, max_batch=max_batch, device=device)

    @classmethod
    def from_hf_model(cls, hf_model, dtype=None, max_batch=64, device="cuda"):
        dtype = dtype or next(hf_model.parameters()).dtype
        return cls.from_state_dict(hf_model.state_dict(), hf_model.config, hf_model.generation_config, dtype=dtype,
                                   max_batch=max_batch, device=device)

    @classmethod
    def from_pretrained(cls, name_or_path, torch_dtype=torch.bfloat16, attn_implementation=None, max_batch=64,
                        device="cuda", **kwargs):
        """Same keywords as the reference's call (run_pseudo_labelling.py:224-232); `attn_implementation` is accepted
        and ignored — the attention kernel is the library's own."""
        from transformers import WhisperForConditionalGeneration  # checkpoint IO only
        hf = WhisperForConditionalGeneration.from_pretrained(name_or_path, torch_dtype=torch.float32, **kwargs)
        return cls.from_hf_model(hf, dtype=torch_dtype, max_batch=max_batch, device=device)

    def _build(self, sd: Dict[str, torch.Tensor]):
        c, dev = self.config, self.device
        if c.encoder_attention_heads != c.decoder_attention_heads or c.encoder_ffn_dim != c.decoder_ffn_dim:
            raise ValueError("encoder/decoder head count and ffn width must match")
        if c.d_model // c.encoder_attention_heads != 64:
            raise ValueError("head dim must be 64")
        wdt = self.dtype
        scale = 64 ** -0.5  # 0.125: folding it into Wq / bq is exact

        def mat(t):
            x = _dev_tensor(t, wdt, dev)
            self._keep.append(x)
            return x.data_ptr()

        def vec(t):
            x = _dev_tensor(t, torch.float32, dev)
            self._keep.append(x)
            return x.data_ptr()

        def f32(name):
            return sd[name].detach().to(torch.float32)

        d = c.d_model
        zeros = torch.zeros(d, dtype=torch.float32, device=next(iter(sd.values())).device)
        enc_layers = (_lib.kw_enc_layer_weights * c.encoder_layers)()
        for l in range(c.encoder_layers):
            p, e = f"model.encoder.layers.{l}.", enc_layers[l]
            e.ln1_w, e.ln1_b = vec(f32(p + "self_attn_layer_norm.weight")), vec(f32(p + "self_attn_layer_norm.bias"))
            e.wqkv = mat(torch.cat([f32(p + "self_attn.q_proj.weight") * scale, f32(p + "self_attn.k_proj.weight"),
                                    f32(p + "self_attn.v_proj.weight")], 0))
            e.bqkv = vec(torch.cat([f32(p + "self_attn.q_proj.bias") * scale, zeros, f32(p + "self_attn.v_proj.bias")]))
            e.wo, e.bo = mat(f32(p + "self_attn.out_proj.weight")), vec(f32(p + "self_attn.out_proj.bias"))
            e.ln2_w, e.ln2_b = vec(f32(p + "final_layer_norm.weight")), vec(f32(p + "final_layer_norm.bias"))
            e.w1, e.b1 = mat(f32(p + "fc1.weight")), vec(f32(p + "fc1.bias"))
            e.w2, e.b2 = mat(f32(p + "fc2.weight")), vec(f32(p + "fc2.bias"))
        dec_layers = (_lib.kw_dec_layer_weights * c.decoder_layers)()
        for l in range(c.decoder_layers):
            p, e = f"model.decoder.layers.{l}.", dec_layers[l]
            e.ln1_w, e.ln1_b = vec(f32(p + "self_attn_layer_norm.weight")), vec(f32(p + "self_attn_layer_norm.bias"))
            e.wqkv = mat(torch.cat([f32(p + "self_attn.q_proj.weight") * scale, f32(p + "self_attn.k_proj.weight"),
                                    f32(p + "self_attn.v_proj.weight")], 0))
            e.bqkv = vec(torch.cat([f32(p + "self_attn.q_proj.bias") * scale, zeros, f32(p + "self_attn.v_proj.bias")]))
            e.wo, e.bo = mat(f32(p + "self_attn.out_proj.weight")), vec(f32(p + "self_attn.out_proj.bias"))
            e.lnx_w, e.lnx_b = vec(f32(p + "encoder_attn_layer_norm.weight")), vec(f32(p + "encoder_attn_layer_norm.bias"))
            e.wq_x = mat(f32(p + "encoder_attn.q_proj.weight") * scale)
            e.bq_x = vec(f32(p + "encoder_attn.q_proj.bias") * scale)
            e.wkv_x = mat(torch.cat([f32(p + "encoder_attn.k_proj.weight"), f32(p + "encoder_attn.v_proj.weight")], 0))
            e.bkv_x = vec(torch.cat([zeros, f32(p + "encoder_attn.v_proj.bias")]))
            e.wo_x, e.bo_x = mat(f32(p + "encoder_attn.out_proj.weight")), vec(f32(p + "encoder_attn.out_proj.bias"))
            e.ln3_w, e.ln3_b = vec(f32(p + "final_layer_norm.weight")), vec(f32(p + "final_layer_norm.bias"))
            e.w1, e.b1 = mat(f32(p + "fc1.weight")), vec(f32(p + "fc1.bias"))
            e.w2, e.b2 = mat(f32(p + "fc2.weight")), vec(f32(p + "fc2.bias"))
        w = _lib.kw_weights()
        w.conv1_w = mat(f32("model.encoder.conv1.weight").permute(0, 2, 1).reshape(d, -1))
        w.conv1_b = vec(f32("model.encoder.conv1.bias"))
        w.conv2_w = mat(f32("model.encoder.conv2.weight").permute(0, 2, 1).reshape(d, -1))
        w.conv2_b = vec(f32("model.encoder.conv2.bias"))
        w.enc_pos = vec(f32("model.encoder.embed_positions.weight"))
        w.enc_ln_w, w.enc_ln_b = vec(f32("model.encoder.layer_norm.weight")), vec(f32("model.encoder.layer_norm.bias"))
        w.tok_embed = mat(f32("model.decoder.embed_tokens.weight"))
        w.dec_pos = vec(f32("model.decoder.embed_positions.weight"))
        w.dec_ln_w, w.dec_ln_b = vec(f32("model.decoder.layer_norm.weight")), vec(f32("model.decoder.layer_norm.bias"))
        w.enc, w.dec = enc_layers, dec_layers
        self._wstructs = (w, enc_layers, dec_layers)

        cfg = _lib.kw_config(c.vocab_size, c.num_mel_bins, d, c.encoder_attention_heads, c.encoder_ffn_dim,
                             c.encoder_layers, c.decoder_layers, c.max_source_positions, c.max_target_positions,
                             _lib.KW_BF16 if wdt == torch.bfloat16 else _lib.KW_F32, self.max_batch)
        g = self.generation_config
        sup = list(g.suppress_tokens or [])
        bsup = list(g.begin_suppress_tokens or [])
        sup_a, bsup_a = (C.c_int32 * max(1, len(sup)))(*sup), (C.c_int32 * max(1, len(bsup)))(*bsup)
        rules = _lib.kw_token_rules(g.eos_token_id, g.pad_token_id, g.no_timestamps_token_id,
                                    -1 if g.max_initial_timestamp_index is None else g.max_initial_timestamp_index,
                                    sup_a, len(sup), bsup_a, len(bsup))
        self._create_args = (cfg, w, rules, sup_a, bsup_a)
        with torch.cuda.device(dev):
            _lib.check(self._lib.kw_model_create(C.byref(cfg), C.byref(w), C.byref(rules), C.byref(self._handle)),
                       "kw_model_create")

    def _new_handle(self) -> C.c_void_p:
        """A second kw_model over the SAME weight tensors: its own workspace / KV pools, no second copy of the weights
        (GenerateStream keeps two batches in flight)."""
        cfg, w, rules, _, _ = self._create_args
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kw_model_create(C.byref(cfg), C.byref(w), C.byref(rules), C.byref(h)), "kw_model_create")
        self._extra_handles.append(h)
        return h

    def __del__(self):
        try:
            for h in getattr(self, "_extra_handles", []):
                if h.value:
                    self._lib.kw_model_destroy(h)
            self._extra_handles = []
            if getattr(self, "_handle", None) and self._handle.value:
                self._lib.kw_model_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ---- nn.Module-ish surface the scripts touch ---------------------------------------------------------------------
    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def get_encoder(self):
        return self.encode_output

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- device stages -----------------------------------------------------------------------------------------------
    def _check_features(self, input_features: torch.Tensor):
        c = self.config
        if input_features.dim() != 3 or input_features.shape[1] != c.num_mel_bins:
            raise ValueError(f"input_features must be [batch, {c.num_mel_bins}, frames], got {tuple(input_features.shape)}")

    def encode(self, input_features: torch.Tensor, return_hidden: bool = True, handle=None) -> Optional[torch.Tensor]:
        """[B, n_mels, 3000] -> fp32 [B, 1500, d] (WhisperEncoder.forward, modeling_whisper.py:593-647)."""
        c = self.config
        self._check_features(input_features)
        expected = 2 * c.max_source_positions
        if input_features.shape[-1] != expected:
            raise ValueError(
                f"Whisper expects the mel input features to be of length {expected}, but found "
                f"{input_features.shape[-1]}. Make sure to pad the input mel features to {expected}.")
        B = input_features.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds max_batch {self.max_batch}")
        # bf16 models: the reference casts features to the model dtype (run_pseudo_labelling.py:338); here the rounding
        # happens where the conv1 im2col kernel stores its bf16 operand — no separate cast kernels
        mel = input_features.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty((B, c.max_source_positions, c.d_model), dtype=torch.float32, device=self.device) \
            if return_hidden else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kw_encode(handle or self._handle, mel.data_ptr(), B,
                                           out.data_ptr() if return_hidden else None, self._stream()), "kw_encode")
        self._last_mel = mel  # keep alive until the stream has consumed it
        return out

    def encode_output(self, input_features, **kwargs) -> EncoderOutput:
        return EncoderOutput(self.encode(input_features))

    def _greedy_pass(self, B: int, prompt: List[int], max_length: int, return_timestamps: bool, handle=None) -> np.ndarray:
        tokens = torch.empty((B, max_length), dtype=torch.int32, device=self.device)
        pr = (C.c_int32 * len(prompt))(*prompt)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kw_greedy_pass(handle or self._handle, B, pr, len(prompt), max_length, int(return_timestamps), 16,
                                                tokens.data_ptr(), self._stream()), "kw_greedy_pass")
        return tokens.cpu().numpy()

    # ---- prompt / segments (host) ----------------------------------------------------------------------------------
    def _init_tokens(self, language, task, return_timestamps: bool) -> List[int]:
        g = self.generation_config
        toks = [g.decoder_start_token_id]
        if language is not None:
            if not isinstance(language, str):
                raise NotImplementedError("per-row language lists are not on the reference path")
            lang = language.lower()
            if lang in g.lang_to_id:
                key = lang
            elif lang in LANGUAGE_NAMES:
                key = f"<|{LANGUAGE_NAMES[lang]}|>"
            elif lang in LANGUAGE_CODES:
                key = f"<|{lang}|>"
            else:
                raise ValueError(f"Unsupported language: {language}. Language should be one of: {LANGUAGE_CODES}.")
            if key not in g.lang_to_id:
                raise ValueError(f"{key} is not supported by this specific model as it is not in the "
                                 "`generation_config.lang_to_id`. (You should just add it to the generation config)")
            toks.append(g.lang_to_id[key])
        elif g.is_multilingual:
            raise NotImplementedError("language detection (language=None) is not on the reference path; pass `language`")
        if task is not None:
            if task not in TASK_IDS:
                raise ValueError(f"The `{task}` task is not supported. The task should be one of `{list(TASK_IDS)}`")
            toks.append(g.task_to_id[task])
        elif language is not None:
            toks.append(g.task_to_id["transcribe"])
        if not return_timestamps and toks[-1] != g.no_timestamps_token_id:
            toks.append(g.no_timestamps_token_id)
        return toks

    @staticmethod
    def _stripped_lengths(arr: np.ndarray, pad_id: int, eos_id: int) -> np.ndarray:
        """Per row of generated ids [n, T]: the length left after HF's pad / eos stripping (generation_whisper.py:1063-1086:
        a row that ends in pad loses as many trailing ids as it holds pads — one fewer when pad == eos — and then a final
        eos), for all rows at once."""
        n_pad = (arr == pad_id).sum(1) - (1 if pad_id == eos_id else 0)
        length = arr.shape[1] - np.where(arr[:, -1] == pad_id, np.maximum(n_pad, 0), 0)
        last = arr[np.arange(arr.shape[0]), np.maximum(length - 1, 0)]
        return length - ((length > 0) & (last == eos_id))

    def _split_segments(self, seq: List[int], seek_num_frames: int):
        """generation_whisper.py:1976-2073 -> (token lists of the completed segments, seek advance in mel frames)."""
        tb = self.generation_config.no_timestamps_token_id + 1
        is_ts = [t >= tb for t in seq]
        single_ending = is_ts[-2:] == [False, True]
        cuts = [i + 1 for i in range(len(seq) - 1) if is_ts[i] and is_ts[i + 1]]
        if not cuts:
            return [seq], seek_num_frames
        if single_ending:
            cuts.append(len(seq))
        else:
            cuts[-1] += 1
        segs, last = [], 0
        for cut in cuts:
            segs.append(seq[last:cut])
            last = cut
        if single_ending:
            return segs, seek_num_frames
        return segs, (seq[last - 2] - tb) * 2  # input_stride = conv1.stride * conv2.stride = 2

    # ---- the reference's entry point --------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_features: Optional[torch.Tensor] = None, generation_config=None, logits_processor=None,
                 stopping_criteria=None, prefix_allowed_tokens_fn=None, synced_gpus: bool = False,
                 return_timestamps: Optional[bool] = None, task: Optional[str] = None,
                 language: Optional[str] = None, is_multilingual: Optional[bool] = None,
                 prompt_ids: Optional[torch.Tensor] = None, prompt_condition_type: Optional[str] = None,
                 condition_on_prev_tokens: Optional[bool] = None, temperature=None,
                 compression_ratio_threshold: Optional[float] = None, logprob_threshold: Optional[float] = None,
                 no_speech_threshold: Optional[float] = None, num_segment_frames: Optional[int] = None,
                 attention_mask: Optional[torch.Tensor] = None, time_precision: float = 0.02,
                 return_token_timestamps: Optional[bool] = None, return_segments: bool = False,
                 return_dict_in_generate: Optional[bool] = None, stats: Optional[dict] = None, **kwargs):
        for name, val in (("logits_processor", logits_processor), ("stopping_criteria", stopping_criteria),
                          ("prefix_allowed_tokens_fn", prefix_allowed_tokens_fn), ("prompt_ids", prompt_ids),
                          ("condition_on_prev_tokens", condition_on_prev_tokens),
                          ("compression_ratio_threshold", compression_ratio_threshold),
                          ("logprob_threshold", logprob_threshold), ("no_speech_threshold", no_speech_threshold),
                          ("return_token_timestamps", return_token_timestamps),
                          ("return_dict_in_generate", return_dict_in_generate)):
            if val is not None and val is not False and not (isinstance(val, (list, tuple)) and len(val) == 0):
                raise NotImplementedError(f"`{name}` is outside the greedy transcription path this library implements")
        if temperature not in (None, 0, 0.0) and temperature != (0.0,):
            raise NotImplementedError("sampling / temperature fallback is outside the greedy transcription path")
        if kwargs.pop("num_beams", 1) not in (None, 1):
            raise NotImplementedError("beam search is outside the greedy transcription path (reference uses num_beams=1)")
        if kwargs.pop("do_sample", False):
            raise NotImplementedError("sampling is outside the greedy transcription path")
        g = WhisperB200GenerationConfig.from_any(generation_config) if generation_config is not None \
            else self.generation_config
        if g is not self.generation_config:
            # the device-side token rules (eos / pad / suppress lists / timestamp ids) were uploaded at construction
            # from self.generation_config; a per-call config may only differ in host-side fields
            baked = ("eos_token_id", "pad_token_id", "no_timestamps_token_id", "max_initial_timestamp_index")
            for k in baked:
                if getattr(g, k) != getattr(self.generation_config, k):
                    raise ValueError(f"per-call generation_config.{k}={getattr(g, k)!r} differs from the value baked "
                                     f"into the device token rules ({getattr(self.generation_config, k)!r}); build "
                                     "the model with that generation_config instead")
            for k in ("suppress_tokens", "begin_suppress_tokens"):
                if list(getattr(g, k) or []) != list(getattr(self.generation_config, k) or []):
                    raise ValueError(f"per-call generation_config.{k} differs from the list baked into the device "
                                     "token rules; build the model with that generation_config instead")
        c = self.config
        # private (GenerateStream): plan only / first pass already done on another workspace handle
        plan_only = kwargs.pop("_plan_only", False)
        first_tokens = kwargs.pop("_first_pass_tokens", None)
        handle = kwargs.pop("_handle", None) or self._handle
        encoder_outputs = kwargs.pop("encoder_outputs", None)
        max_length = kwargs.pop("max_length", None)
        max_new_tokens = kwargs.pop("max_new_tokens", None)
        if kwargs:
            raise TypeError(f"generate() got unsupported keyword arguments: {sorted(kwargs)}")

        seg_frames = 2 * c.max_source_positions
        if encoder_outputs is not None:
            enc = encoder_outputs[0] if not isinstance(encoder_outputs, torch.Tensor) else encoder_outputs
            B, total, in_device = enc.shape[0], seg_frames, enc.device
        else:
            if input_features is None:
                raise ValueError("generate() needs `input_features` or `encoder_outputs`")
            self._check_features(input_features)
            B, total, in_device = input_features.shape[0], input_features.shape[-1], input_features.device
        is_shortform = total <= seg_frames
        if return_timestamps is None:
            return_timestamps = g.return_timestamps
        if not is_shortform:
            if return_timestamps is False:
                raise ValueError(
                    "You have passed more than 3000 mel input features (> 30 seconds) which automatically enables "
                    "long-form generation which requires the model to predict timestamp tokens. Please either pass "
                    "`return_timestamps=True` or make sure to pass no more than 3000 mel input features.")
            return_timestamps = True
        return_timestamps = bool(return_timestamps)
        prompt = self._init_tokens(language, task, return_timestamps)
        if max_new_tokens is not None:
            if max_new_tokens + len(prompt) > c.max_target_positions:
                raise ValueError(
                    f"The length of `decoder_input_ids`, including special start tokens, prompt tokens, and previous "
                    f"tokens, is {len(prompt)},  and `max_new_tokens` is {max_new_tokens}. Thus, the combined length "
                    f"exceeds the `max_target_positions` of the Whisper model: {c.max_target_positions}.")
            max_length = len(prompt) + max_new_tokens
        elif max_length is None:
            max_length = g.max_length
        max_length = min(int(max_length), c.max_target_positions)
        if max_length <= len(prompt):
            raise ValueError(f"max_length={max_length} leaves no room after the {len(prompt)}-token prompt")
        if plan_only:
            return prompt, max_length, return_timestamps

        if not is_shortform and B > 1:
            if attention_mask is None:
                raise ValueError("When doing batched long-form audio transcription, make sure to pass an `attention_mask`.")
            max_frames = [int(v) for v in attention_mask.sum(-1).cpu().tolist()]
        else:
            max_frames = [total] * B
        seek = [0] * B
        out: List[List[int]] = [[] for _ in range(B)]
        n_pass = 0
        feats = None if encoder_outputs is not None else input_features.to(self.device)
        while any(s < m for s, m in zip(seek, max_frames)):
            rows = [b for b in range(B) if seek[b] < max_frames[b]]
            nf = {b: min(max_frames[b] - seek[b], seg_frames) for b in rows}
            for c0 in range(0, len(rows), self.max_batch):
                chunk = rows[c0:c0 + self.max_batch]
                if first_tokens is not None and n_pass == 0:
                    pass  # GenerateStream: this chunk's encoder and first greedy pass already ran on `handle`
                elif encoder_outputs is not None:
                    e = enc[chunk].to(device=self.device, dtype=torch.float32).contiguous()
                    with torch.cuda.device(self.device):
                        _lib.check(self._lib.kw_set_encoder_output(handle, e.data_ptr(), len(chunk),
                                                                   self._stream()), "kw_set_encoder_output")
                    self._last_mel = e
                else:
                    if is_shortform and len(chunk) == B and all(seek[b] == 0 for b in chunk) and total == seg_frames:
                        seg = feats
                    else:  # _get_input_segment: slice [seek, seek + n) and zero-pad to 3000 frames (:1831-1850)
                        seg = torch.zeros((len(chunk), c.num_mel_bins, seg_frames), dtype=feats.dtype, device=self.device)
                        for i, b in enumerate(chunk):
                            seg[i, :, : nf[b]] = feats[b, :, seek[b]: seek[b] + nf[b]]
                    self.encode(seg, return_hidden=False, handle=handle)
                toks = first_tokens if (first_tokens is not None and n_pass == 0) else \
                    self._greedy_pass(len(chunk), prompt, max_length, return_timestamps, handle)
                n_pass += 1
                # strip padding keeping one eos, then the eos itself (:1063-1086) — lengths for the whole chunk at once;
                # rows without timestamp tokens are one segment that consumes the whole window (no per-row Python)
                arr = np.asarray(toks)[:, len(prompt):]
                length = self._stripped_lengths(arr, g.pad_token_id, g.eos_token_id)
                has_ts = (arr >= g.no_timestamps_token_id + 1).any(1)
                for i, b in enumerate(chunk):
                    n = int(length[i])
                    if n <= 0:
                        seek[b] += nf[b]
                        continue
                    seq = arr[i, :n].tolist()
                    if not has_ts[i]:
                        seek[b] += nf[b]
                        out[b].extend(seq)
                        continue
                    segs, adv = self._split_segments(seq, nf[b])
                    seek[b] += adv
                    for s in segs:
                        out[b].extend(s)
            # encoder_outputs: HF keeps the seek loop running over the SAME encoder output (input_features is None, so
            # _get_input_segment has nothing to re-slice; generation_whisper.py:785-903) — so does this loop.
        if stats is not None:
            stats["passes"] = n_pass
        L = max((len(o) for o in out), default=0)
        res_np = np.full((B, L), g.pad_token_id, dtype=np.int64)
        for b, o in enumerate(out):
            if o:
                res_np[b, : len(o)] = o
        res = torch.from_numpy(res_np).to(in_device)
        if return_segments:
            return {"sequences": res, "segments": out}
        return res

    def generate_stream(self, **generate_kwargs) -> "GenerateStream":
        """Throughput mode of the labelling loop (`for batch in loader: ids = model.generate(batch, **kw)`,
        run_pseudo_labelling.py:333-341): `stream.submit(batch)` returns the ids of the oldest finished batch (None while
        the pipeline fills), `stream.flush()` the remaining ones, one batch per call.  Same arguments and results as
        `generate`, plus `coalesce=k` (k submitted batches run as one device batch); see GenerateStream."""
        return GenerateStream(self, **generate_kwargs)

    # ---- teacher-forcing forward (distillation's frozen teacher) -----------------------------------------------------
    @torch.no_grad()
    def forward(self, input_features: Optional[torch.Tensor] = None, attention_mask=None,
                decoder_input_ids: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                encoder_outputs=None, **kwargs) -> "Seq2SeqOutput":
        """`teacher_model(encoder_outputs=..., labels=...)` / `teacher_model(**batch)` of run_distillation.py:641-649
        (WhisperForConditionalGeneration.forward, HF/models/whisper/modeling_whisper.py:964-1100): all target positions
        at once, no cache -> fp32 logits [B, T, vocab] on the device (+ the cross-entropy `loss` when labels are given).
        `labels` are shifted right into decoder inputs exactly as HF does (:68-81, -100 -> pad_token_id)."""
        drop = {k: kwargs.pop(k) for k in ("use_cache", "return_dict", "output_attentions", "output_hidden_states",
                                           "decoder_attention_mask", "past_key_values", "cache_position") if k in kwargs}
        if drop.get("decoder_attention_mask") is not None or drop.get("past_key_values") is not None:
            raise NotImplementedError("decoder_attention_mask / past_key_values are outside the teacher-forcing path")
        if kwargs:
            raise TypeError(f"forward() got unsupported keyword arguments: {sorted(kwargs)}")
        c = self.config
        if labels is not None:
            if labels.shape[1] > c.max_target_positions:
                raise ValueError(f"Labels' sequence length {labels.shape[1]} cannot exceed the maximum allowed length "
                                 f"of {c.max_target_positions} tokens.")
            if decoder_input_ids is None:
                decoder_input_ids = labels.new_zeros(labels.shape)
                decoder_input_ids[:, 1:] = labels[:, :-1]
                decoder_input_ids[:, 0] = c.decoder_start_token_id
                decoder_input_ids = decoder_input_ids.masked_fill(decoder_input_ids == -100, c.pad_token_id)
        if decoder_input_ids is None:
            raise ValueError("forward() needs `decoder_input_ids` or `labels`")
        B, T = decoder_input_ids.shape
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds max_batch {self.max_batch}")
        if encoder_outputs is not None:
            enc = encoder_outputs[0] if not isinstance(encoder_outputs, torch.Tensor) else encoder_outputs
            e = enc.to(device=self.device, dtype=torch.float32).contiguous()
            with torch.cuda.device(self.device):
                _lib.check(self._lib.kw_set_encoder_output(self._handle, e.data_ptr(), B, self._stream()),
                           "kw_set_encoder_output")
            self._last_mel = e
            enc_hidden = enc
        else:
            if input_features is None:
                raise ValueError("forward() needs `input_features` or `encoder_outputs`")
            enc_hidden = self.encode(input_features, return_hidden=True)
        ids = decoder_input_ids.to(device=self.device, dtype=torch.int32).contiguous()
        logits = torch.empty((B, T, c.vocab_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kw_decoder_forward(self._handle, ids.data_ptr(), B, T, logits.data_ptr(),
                                                    self._stream()), "kw_decoder_forward")
        loss = None
        if labels is not None:  # CrossEntropyLoss over all positions, ignore_index = -100 (:1086-1090)
            loss = torch.nn.functional.cross_entropy(logits.view(-1, c.vocab_size),
                                                     labels.to(self.device).reshape(-1), ignore_index=-100)
        return Seq2SeqOutput(loss=loss, logits=logits, encoder_last_hidden_state=enc_hidden)

    __call__ = forward

    # ---- test hooks ------------------------------------------------------------------------------------------------------
    def step_logits(self, tokens: torch.Tensor, pos: int) -> torch.Tensor:
        """Raw fp32 logits of decoder position `pos` given int32 token history [B, >= pos+1] on the device; extends the
        self-KV pool as a side effect (encode + cross_kv must have run)."""
        B = tokens.shape[0]
        logits = torch.empty((B, self.config.vocab_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kw_decode_step(self._handle, tokens.data_ptr(), tokens.shape[1], B, pos, 1, 0, 0, None,
                                                logits.data_ptr(), self._stream()), "kw_decode_step")
        return logits

    def cross_kv(self, B: int):
        with torch.cuda.device(self.device):
            _lib.check(self._lib.kw_cross_kv(self._handle, B, self._stream()), "kw_cross_kv")


class GenerateStream:
    """Two device batches in flight on one GPU: while batch i is decoded, the encoder of batch i + 1 runs in layer groups
    between its decoder positions (kw_encode_decode).  Results are identical to calling `model.generate(batch, **kwargs)`
    batch by batch (same kernels on the same data; only their order on the stream changes); each batch's ids come back a
    few calls later.  Why it is faster: after a ~140 ms block of power-capped encoder GEMMs the clock governor keeps the
    SM clock low through the whole latency-bound decode block that follows; alternating short slices of both avoids that.

    `coalesce=k` (k * batch <= model.max_batch): k submitted batches are run as ONE device batch.  A decoder position
    is a chain of ~25 dependent, latency-bound kernels whose cost hardly depends on the number of rows (the weights are
    streamed once per position either way), so decoding 128 utterances at a time nearly halves the per-utterance cost
    of everything but the cross-attention K/V stream.  Utterances are independent rows of every kernel: the ids of a
    batch do not depend on what it was grouped with (tests/test_surface_gpu.py).

        stream = model.generate_stream(language="ja", task="transcribe", return_timestamps=False, max_length=128)
        for batch in loader:
            ids = stream.submit(batch["input_features"])      # ids of the oldest finished batch (None while priming)
            ...
        while (ids := stream.flush()) is not None:             # one batch per call, in submission order
            ...

    Batches must be short-form ([B <= max_batch, n_mels, 3000]); with return_timestamps=True the first greedy pass of a
    batch is pipelined and any further seek passes run when its result is collected."""

    def __init__(self, model: WhisperB200ForConditionalGeneration, coalesce: int = 1, **generate_kwargs):
        for k in ("encoder_outputs", "attention_mask", "input_features"):
            if generate_kwargs.get(k) is not None:
                raise ValueError(f"GenerateStream takes `{k}` per submit() / not at all")
        if int(coalesce) < 1:
            raise ValueError(f"coalesce={coalesce} must be >= 1")
        self.model = model
        self.coalesce = int(coalesce)
        self.kwargs = dict(generate_kwargs)
        self.stats = self.kwargs.pop("stats", None)
        self._want_segments = bool(self.kwargs.pop("return_segments", False))
        # one extra workspace per model, shared by every stream made from it (use one stream at a time)
        if not model._extra_handles:
            model._new_handle()
        self._handles = [model._handle, model._extra_handles[0]]
        self._slot = 0
        self._pending = None  # (slot, features of the device batch, sizes of the submitted batches inside it)
        self._plan = None
        self._buf: List[torch.Tensor] = []   # submitted batches waiting for their group to fill
        self._ready: List = []               # finished per-batch results, oldest first
        self._late = None                    # (device batch, pinned token buffer, copy-done event): see _launch
        self.device_batches = 0              # device batches launched so far

    @property
    def buffered(self) -> int:
        """Submitted batches whose group (coalesce > 1) has not been launched yet."""
        return len(self._buf)

    def _finish(self, pending, first_tokens):
        """Host side of a device batch whose first greedy pass is done: strip / segment (and seek passes with timestamps),
        then one result per submitted batch, exactly as `generate` packs it."""
        slot, feats, sizes = pending
        m = self.model
        st = {} if self.stats is not None else None
        res = m.generate(feats, _first_pass_tokens=first_tokens, _handle=self._handles[slot], stats=st,
                         return_segments=True, **self.kwargs)
        if st is not None:
            self.stats["passes"] = st.get("passes", 0)
        pad, b0 = m.generation_config.pad_token_id, 0
        for n in sizes:
            rows = res["segments"][b0:b0 + n]
            b0 += n
            L = max((len(o) for o in rows), default=0)
            arr = np.full((n, L), pad, dtype=np.int64)
            for i, o in enumerate(rows):
                if o:
                    arr[i, : len(o)] = o
            ids = torch.from_numpy(arr).to(feats.device)
            self._ready.append({"sequences": ids, "segments": rows} if self._want_segments else ids)

    def _launch(self):
        """Launches the buffered group as one device batch (its encoder, interleaved with the greedy pass of the device
        batch in flight) and finishes the batch that was in flight."""
        m = self.model
        sizes = [int(t.shape[0]) for t in self._buf]
        mel = self._buf[0] if len(self._buf) == 1 else torch.cat(self._buf, 0)
        self._buf = []
        B = mel.shape[0]
        prompt, max_length, ts = self._plan
        prev = self._pending
        pr = (C.c_int32 * len(prompt))(*prompt)
        tokens = None
        with torch.cuda.device(m.device):
            if prev is None:
                _lib.check(m._lib.kw_encode_decode(self._handles[self._slot], mel.data_ptr(), B, None, 0, pr, len(prompt),
                                                   max_length, int(ts), 16, None, m._stream()), "kw_encode_decode")
            else:
                Bp = prev[1].shape[0]
                tokens = torch.empty((Bp, max_length), dtype=torch.int32, device=m.device)
                _lib.check(m._lib.kw_encode_decode(self._handles[self._slot], mel.data_ptr(), B, self._handles[prev[0]], Bp,
                                                   pr, len(prompt), max_length, int(ts), 16, tokens.data_ptr(),
                                                   m._stream()), "kw_encode_decode")
        self.device_batches += 1
        self._pending = (self._slot, mel, sizes)
        self._slot ^= 1
        if prev is None:
            return
        if ts or os.environ.get("KW_STREAM_DEFER", "1") == "0":   # (KW_STREAM_DEFER=0: A/B measurements)
            # a batch may need further seek passes on its own workspace, which the next launch reuses: finish it now
            self._finish(prev, tokens.cpu().numpy())
            return
        # Without timestamps a short-form batch is done after its first pass and only host work is left (strip, pack).
        # Blocking on the tokens here would drain the stream at every device batch: the GPU would idle while the host
        # packs the ids and enqueues the next batch.  The ids are instead copied to pinned memory behind an event and
        # picked up at the NEXT launch, after that launch's kernels are queued (results come one device batch later).
        host = torch.empty(tokens.shape, dtype=tokens.dtype, pin_memory=True)
        with torch.cuda.device(m.device):
            host.copy_(tokens, non_blocking=True)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(m.device))
        late, self._late = self._late, (prev, host, done)
        if late is not None:
            self._finish_late(late)

    def _finish_late(self, late):
        prev, host, done = late
        done.synchronize()
        self._finish(prev, host.numpy())

    def submit(self, input_features: torch.Tensor):
        m, c = self.model, self.model.config
        m._check_features(input_features)
        B = input_features.shape[0]
        held = sum(int(t.shape[0]) for t in self._buf)
        if input_features.shape[-1] != 2 * c.max_source_positions or held + B > m.max_batch:
            raise ValueError(f"GenerateStream batches must be [B, {c.num_mel_bins}, {2 * c.max_source_positions}] "
                             f"short-form features with coalesce x B <= max_batch = {m.max_batch}, got "
                             f"{tuple(input_features.shape)} (coalesce={self.coalesce}, {held} rows already buffered)")
        if self._plan is None:
            self._plan = m.generate(input_features, _plan_only=True, **self.kwargs)
        self._buf.append(input_features.to(device=m.device, dtype=torch.float32).contiguous())
        if len(self._buf) >= self.coalesce:
            self._launch()
        return self._ready.pop(0) if self._ready else None

    def flush(self):
        """Finishes what is still in flight and returns the ids of ONE batch per call, oldest first; None when the stream
        is empty (`while (ids := stream.flush()) is not None`).  With coalesce = 1 a single call returns the last batch."""
        if not self._ready and self._buf:
            self._launch()       # a group that did not fill: its own (smaller) device batch
        if not self._ready and self._late is not None:
            late, self._late = self._late, None
            self._finish_late(late)
        if not self._ready and self._pending is not None:
            prev, self._pending = self._pending, None
            prompt, max_length, ts = self._plan
            toks = self.model._greedy_pass(prev[1].shape[0], prompt, max_length, ts, self._handles[prev[0]])
            self._finish(prev, toks)
        return self._ready.pop(0) if self._ready else None
