import os, sys, torch
sys.path.insert(0, "/root/repo")
from kotoba_whisper_b200 import _lib
lib = _lib.load(); BF16 = _lib.KW_BF16
B, H, T = 64, 20, 1500; d = H * 64
qkv = (torch.randn(B, T, 3 * d, device="cuda") * 0.5).bfloat16(); out = torch.zeros(B, T, d, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
st = torch.cuda.current_stream().cuda_stream
fn = lambda: _lib.check(lib.kw_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, T, T, T * 3 * d, 3 * d, T * 3 * d, 3 * d, T * d, d, BF16, st))
for _ in range(3): fn()
torch.cuda.synchronize(); lib.kw_debug_attention_desc(-1, 0, 0)
for _ in range(5): fn()
torch.cuda.synchronize(); lib.kw_debug_attention_desc(-1, 0, 0)
