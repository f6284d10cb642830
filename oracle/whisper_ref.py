"""Plain torch-CPU fp32 restatement of the Whisper encoder / decoder / greedy `generate` path.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, function by function:
  * encoder            HF/models/whisper/modeling_whisper.py:593-647 (stem :613-626, layer :380-414, attention :284-357)
  * decoder step       HF/models/whisper/modeling_whisper.py:734-796, 449-506; proj_out :1081
  * logits processors  HF/generation/logits_process.py:1855-1862, 1898-1902, 1996-2043
  * greedy loop        HF/generation/utils.py:2743-2809
  * generate           HF/models/whisper/generation_whisper.py:649-968 (init tokens :1455-1608, processors :1774-1812,
                       input segment :1831-1850, pad/eos strip :1063-1086, segments :1976-2073, padding :126-237)

It consumes an HF `state_dict` (names in SURVEY.md §8b) so the CUDA path, HF and this file see identical weights.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# configuration


@dataclass
class ArchConfig:
    """Model dimensions (HF WhisperConfig fields actually used by the path)."""
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


# the first 82 of HF's NON_SPEECH_TOKENS_MULTI (HF/models/whisper/configuration_whisper.py:34-44) are ordinary text
# ids (< 50257); listed literally so that the oracle does not depend on transformers being importable.
_NON_SPEECH_82 = [
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647,
    7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
    18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870,
    50254,
]

# language order of HF/models/whisper/tokenization_whisper.py LANGUAGES (id = 50259 + index)
_LANG_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi ml cy "
    "sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo "
    "uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue"
).split()


@dataclass
class GenConfig:
    """Hand-built generation config = public whisper-large-v3 values."""
    decoder_start_token_id: int = 50258
    eos_token_id: int = 50257
    pad_token_id: int = 50257
    no_timestamps_token_id: int = 50364
    max_initial_timestamp_index: Optional[int] = 50
    max_length: int = 448
    begin_suppress_tokens: Sequence[int] = (220, 50257)
    suppress_tokens: Sequence[int] = tuple(_NON_SPEECH_82 + [50258, 50359, 50360, 50361, 50362, 50363])
    lang_to_id: Dict[str, int] = field(default_factory=lambda: {f"<|{c}|>": 50259 + i for i, c in enumerate(_LANG_CODES)})
    task_to_id: Dict[str, int] = field(default_factory=lambda: {"translate": 50359, "transcribe": 50360})

    @property
    def timestamp_begin(self) -> int:
        return self.no_timestamps_token_id + 1


def sinusoids(length: int, channels: int) -> torch.Tensor:
    """Encoder position table, modeling_whisper.py:55-64 (halves concatenated, not interleaved)."""
    inc = math.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


# --------------------------------------------------------------------------------------------------
# model arithmetic


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _mha(q, k, v, n_heads, causal_offset: Optional[int] = None):
    """softmax(q k^T) v with q already scaled (modeling_whisper.py:310, 342-352). q:[B,Tq,d] k,v:[B,Tk,d]."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    hd = d // n_heads
    qh = q.view(B, Tq, n_heads, hd).transpose(1, 2)
    kh = k.view(B, Tk, n_heads, hd).transpose(1, 2)
    vh = v.view(B, Tk, n_heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if causal_offset is not None:  # query i (absolute position causal_offset+i) sees keys 0..causal_offset+i
        qi = torch.arange(Tq).view(-1, 1) + causal_offset
        kj = torch.arange(Tk).view(1, -1)
        s = s.masked_fill(kj > qi, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Tq, d)


class WhisperRef:
    """Functional fp32 Whisper built from an HF state_dict."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], arch: ArchConfig, gen: Optional[GenConfig] = None):
        self.arch = arch
        self.gen = gen or GenConfig()
        self.w = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}
        if "proj_out.weight" not in self.w:  # tied (modeling_whisper.py:966)
            self.w["proj_out.weight"] = self.w["model.decoder.embed_tokens.weight"]

    # ---- encoder -------------------------------------------------------------------------------
    def _attn_proj(self, pre, x_q, x_kv):
        w = self.w
        hd = self.arch.d_model // self.arch.encoder_attention_heads
        q = F.linear(x_q, w[pre + "q_proj.weight"], w[pre + "q_proj.bias"]) * hd ** -0.5  # scale after bias (:310)
        k = F.linear(x_kv, w[pre + "k_proj.weight"])  # no bias (modeling_whisper.py:279)
        v = F.linear(x_kv, w[pre + "v_proj.weight"], w[pre + "v_proj.bias"])
        return q, k, v

    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, n_mels, 3000] -> [B, 1500, d]."""
        w, a = self.w, self.arch
        if mel.shape[-1] != a.max_source_positions * 2:
            raise ValueError(f"expected mel length {a.max_source_positions * 2}, got {mel.shape[-1]}")
        x = mel.to(torch.float32)
        x = F.gelu(F.conv1d(x, w["model.encoder.conv1.weight"], w["model.encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, w["model.encoder.conv2.weight"], w["model.encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1) + w["model.encoder.embed_positions.weight"]
        for l in range(a.encoder_layers):
            p = f"model.encoder.layers.{l}."
            h = _ln(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"])
            q, k, v = self._attn_proj(p + "self_attn.", h, h)
            o = _mha(q, k, v, a.encoder_attention_heads)
            x = x + F.linear(o, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
            h = _ln(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"])
            h = F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
            x = x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
        return _ln(x, w["model.encoder.layer_norm.weight"], w["model.encoder.layer_norm.bias"])

    # ---- decoder -------------------------------------------------------------------------------
    def cross_kv(self, enc: torch.Tensor):
        """Per decoder layer (K, V) of the encoder output, built once per pass (modeling_whisper.py:326-336)."""
        out = []
        for l in range(self.arch.decoder_layers):
            p = f"model.decoder.layers.{l}.encoder_attn."
            k = F.linear(enc, self.w[p + "k_proj.weight"])
            v = F.linear(enc, self.w[p + "v_proj.weight"], self.w[p + "v_proj.bias"])
            out.append((k, v))
        return out

    def decode(self, ids: torch.Tensor, past_len: int, self_cache: List, cross) -> torch.Tensor:
        """ids [B,T] at positions past_len.. -> hidden [B,T,d] after the final LN; appends to self_cache in place."""
        w, a = self.w, self.arch
        T = ids.shape[1]
        x = w["model.decoder.embed_tokens.weight"][ids] + w["model.decoder.embed_positions.weight"][past_len:past_len + T]
        H = a.decoder_attention_heads
        for l in range(a.decoder_layers):
            p = f"model.decoder.layers.{l}."
            h = _ln(x, w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"])
            q, k, v = self._attn_proj(p + "self_attn.", h, h)
            if self_cache[l] is None:
                self_cache[l] = (k, v)
            else:
                self_cache[l] = (torch.cat([self_cache[l][0], k], 1), torch.cat([self_cache[l][1], v], 1))
            o = _mha(q, self_cache[l][0], self_cache[l][1], H, causal_offset=past_len)
            x = x + F.linear(o, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
            h = _ln(x, w[p + "encoder_attn_layer_norm.weight"], w[p + "encoder_attn_layer_norm.bias"])
            q = F.linear(h, w[p + "encoder_attn.q_proj.weight"], w[p + "encoder_attn.q_proj.bias"]) * (a.d_model // H) ** -0.5
            o = _mha(q, cross[l][0], cross[l][1], H)
            x = x + F.linear(o, w[p + "encoder_attn.out_proj.weight"], w[p + "encoder_attn.out_proj.bias"])
            h = _ln(x, w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"])
            h = F.gelu(F.linear(h, w[p + "fc1.weight"], w[p + "fc1.bias"]))
            x = x + F.linear(h, w[p + "fc2.weight"], w[p + "fc2.bias"])
        return _ln(x, w["model.decoder.layer_norm.weight"], w["model.decoder.layer_norm.bias"])

    def logits(self, hidden_last: torch.Tensor) -> torch.Tensor:
        return F.linear(hidden_last, self.w["proj_out.weight"]).float()

    # ---- logits processors ---------------------------------------------------------------------
    def process_scores(self, scores: torch.Tensor, sampled: List[List[int]], return_timestamps: bool) -> torch.Tensor:
        """scores [B,V] fp32, sampled[b] = tokens generated so far in this pass (after the prompt)."""
        g = self.gen
        s = scores.clone()
        ninf = float("-inf")
        at_begin = len(sampled[0]) == 0
        if at_begin and g.begin_suppress_tokens:  # SuppressTokensAtBegin, logits_process.py:1855-1862
            s[:, list(g.begin_suppress_tokens)] = ninf
        if g.suppress_tokens:  # SuppressTokens, :1898-1902
            s[:, list(g.suppress_tokens)] = ninf
        if not return_timestamps:
            return s
        tb = g.timestamp_begin  # WhisperTimeStamp, :1996-2043
        s[:, g.no_timestamps_token_id] = ninf
        for b, seq in enumerate(sampled):
            last_ts = len(seq) >= 1 and seq[-1] >= tb
            pen_ts = len(seq) < 2 or seq[-2] >= tb
            if last_ts:
                if pen_ts:
                    s[b, tb:] = ninf
                else:
                    s[b, : g.eos_token_id] = ninf
            ts = [t for t in seq if t >= tb]
            if ts:
                last = ts[-1] if (last_ts and not pen_ts) else ts[-1] + 1
                s[b, tb:last] = ninf
        if at_begin:
            s[:, :tb] = ninf
            if g.max_initial_timestamp_index is not None:
                s[:, tb + g.max_initial_timestamp_index + 1:] = ninf
        logp = torch.log_softmax(s.float(), dim=-1)
        for b in range(s.shape[0]):
            if logp[b, tb:].logsumexp(dim=-1) > logp[b, :tb].max():
                s[b, :tb] = ninf
        return s

    # ---- one greedy pass (GenerationMixin._sample with the Whisper processors) ---------------------
    def greedy_pass(self, mel: torch.Tensor, prompt: List[int], max_length: int, return_timestamps: bool,
                    trace: Optional[dict] = None, enc: Optional[torch.Tensor] = None) -> List[List[int]]:
        """Returns, per row, the tokens generated after the prompt (eos kept, pad after eos kept, like HF sequences)."""
        g = self.gen
        B = mel.shape[0] if enc is None else enc.shape[0]
        if enc is None:
            enc = self.encode(mel)
        cross = self.cross_kv(enc)
        cache: List = [None] * self.arch.decoder_layers
        ids = torch.tensor([prompt] * B, dtype=torch.long)
        hidden = self.decode(ids, 0, cache, cross)
        sampled: List[List[int]] = [[] for _ in range(B)]
        unfinished = [True] * B
        cur_len = len(prompt)
        if trace is not None:
            trace.update(enc=enc, logits=[], margins=[])
        while True:
            raw = self.logits(hidden[:, -1])
            sc = self.process_scores(raw, sampled, return_timestamps)
            nxt = sc.argmax(dim=-1).tolist()
            if trace is not None:
                trace["logits"].append(raw)
                top2 = sc.topk(2, dim=-1).values
                trace["margins"].append((top2[:, 0] - top2[:, 1]))
            for b in range(B):
                tok = nxt[b] if unfinished[b] else g.pad_token_id
                sampled[b].append(tok)
                if tok == g.eos_token_id:
                    unfinished[b] = False
            cur_len += 1
            if cur_len >= max_length or not any(unfinished):
                break
            step_ids = torch.tensor([[s[-1]] for s in sampled], dtype=torch.long)
            hidden = self.decode(step_ids, cur_len - 1, cache, cross)
        return sampled

    # ---- WhisperGenerationMixin.generate ---------------------------------------------------------
    def init_tokens(self, language: Optional[str], task: Optional[str], return_timestamps: bool) -> List[int]:
        g = self.gen
        toks = [g.decoder_start_token_id]
        if language is not None:
            key = language if language.startswith("<|") else f"<|{language.lower()}|>"
            if key not in g.lang_to_id:
                raise ValueError(f"Unsupported language: {language}")
            toks.append(g.lang_to_id[key])
        if task is not None:
            if task not in g.task_to_id:
                raise ValueError(f"The `{task}` task is not supported")
            toks.append(g.task_to_id[task])
        elif language is not None:
            toks.append(g.task_to_id["transcribe"])
        if not return_timestamps and toks[-1] != g.no_timestamps_token_id:
            toks.append(g.no_timestamps_token_id)
        return toks

    def retrieve_segment(self, seq: List[int], seek_num_frames: int):
        """-> (list of token lists, seek advance in mel frames); generation_whisper.py:1976-2073, input_stride = 2."""
        tb = self.gen.timestamp_begin
        is_ts = [t >= tb for t in seq]
        single_ending = is_ts[-2:] == [False, True]
        consec = [i + 1 for i in range(len(seq) - 1) if is_ts[i] and is_ts[i + 1]]
        if consec:
            slices = list(consec)
            if single_ending:
                slices.append(len(seq))
            else:
                slices[-1] += 1
            segs, last = [], 0
            for cur in slices:
                segs.append(seq[last:cur])
                last = cur
            if single_ending:
                return segs, seek_num_frames
            return segs, (seq[last - 2] - tb) * 2
        return [list(seq)], seek_num_frames

    def generate(self, mel: torch.Tensor, language: Optional[str] = None, task: Optional[str] = None,
                 return_timestamps: bool = False, max_length: Optional[int] = None,
                 stats: Optional[dict] = None) -> torch.Tensor:
        """mel [B, n_mels, T<=3000 frames... or longer with timestamps] -> LongTensor [B, L] (prompt/eos stripped)."""
        g = self.gen
        B, _, total = mel.shape
        seg_frames = self.arch.max_source_positions * 2
        if total > seg_frames and not return_timestamps:
            raise ValueError("long-form generation requires return_timestamps=True")
        max_length = g.max_length if max_length is None else max_length
        prompt = self.init_tokens(language, task, return_timestamps)
        seek = [0] * B
        max_frames = [total] * B
        out: List[List[int]] = [[] for _ in range(B)]
        n_pass = 0
        while any(s < m for s, m in zip(seek, max_frames)):
            rows = [b for b in range(B) if seek[b] < max_frames[b]]
            nf = [min(max_frames[b] - seek[b], seg_frames) for b in rows]
            seg = torch.zeros(len(rows), mel.shape[1], seg_frames, dtype=mel.dtype)
            for i, b in enumerate(rows):
                seg[i, :, : nf[i]] = mel[b, :, seek[b]: seek[b] + nf[i]]
            sampled = self.greedy_pass(seg, prompt, max_length, return_timestamps)
            n_pass += 1
            for i, b in enumerate(rows):
                seq = list(sampled[i])
                if seq[-1] == g.pad_token_id:  # generation_whisper.py:1063-1076
                    n_pad = sum(1 for t in seq if t == g.pad_token_id)
                    if g.pad_token_id == g.eos_token_id:
                        n_pad -= 1
                    if n_pad:
                        seq = seq[:-n_pad]
                if seq[-1] == g.eos_token_id:
                    seq = seq[:-1]
                if len(seq) == 0:
                    seek[b] += nf[i]
                    continue
                segs, adv = self.retrieve_segment(seq, nf[i])
                seek[b] += adv
                for s in segs:
                    out[b].extend(s)
        if stats is not None:
            stats["passes"] = n_pass
        L = max((len(o) for o in out), default=0)
        res = torch.full((B, L), g.pad_token_id, dtype=torch.long)
        for b, o in enumerate(out):
            res[b, : len(o)] = torch.tensor(o, dtype=torch.long)
        return res
