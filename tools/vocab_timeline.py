"""Timeline of the decode-time vocabulary GEMM (CTA 0: tiles 0, 148, 296) inside a short greedy pass, fused arg-max epilogue
vs plain fp32-logit store: %globaltimer stamps of the LAST skinny GEMM launch of the pass (= the vocabulary projection)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import KOTOBA, synth_audio  # noqa: E402
from kotoba_whisper_b200 import WhisperB200Config, WhisperB200ForConditionalGeneration, WhisperFeatureExtractorB200, _lib  # noqa: E402
from kotoba_whisper_b200.random_init import random_state_dict  # noqa: E402
lib = _lib.load()
lib.kw_set_decode_graph(0)  # the stamp pointer must not be baked into a captured graph
B = 64
dev = torch.device("cuda", 0)
cfg = WhisperB200Config(**dict(KOTOBA, encoder_layers=2))
model = WhisperB200ForConditionalGeneration.from_state_dict(random_state_dict(cfg, 0, dev), cfg, dtype=torch.bfloat16, max_batch=B, device=dev)
fe = WhisperFeatureExtractorB200(feature_size=128, device=dev)
audio = torch.from_numpy(synth_audio(B, 1000)).to(dev)
model.encode(fe.logmel_device(audio), return_hidden=False)
names = {0: "start", 1: "W requested", 2: "dep wait done", 3: "stage0 landed", 8: "acc 0 visible", 11: "tile 0 MMAs committed",
         9: "tile 0 epilogue done", 5: "tile 1 epilogue done", 4: "last MMA commit", 6: "last epilogue done", 7: "all done"}
stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
for ts in (False, True):
    for fused in (0, 1):
        lib.kw_set_sample_fused(fused)
        for ml in (12, 40):
            lib.kw_debug_gemm_stamps(stamps.data_ptr())
            stamps.zero_()
            model._greedy_pass(B, [50258, 50266, 50360] + ([] if ts else [50364]), ml, ts)
            torch.cuda.synchronize()
            lib.kw_debug_gemm_stamps(None)
            s = stamps.cpu().tolist()
            print(f"ts={int(ts)} fused={fused} max_length={ml}:", {names[i]: round((s[i] - s[0]) / 1e3, 2) for i in names if s[i]})
lib.kw_set_sample_fused(1)
