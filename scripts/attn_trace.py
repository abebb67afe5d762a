"""Debug probe (library built with FOHO_B200_EXTRA_NVCC_FLAGS=-DFOHO_ATTN_TRACE): per-block SM-clock stamps of the
attention forward's softmax groups and MMA issuer on CTA 0, printed relative to the first stamp."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from followmyhold_b200 import _lib
from followmyhold_b200.decoder import tc

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = "cuda:0"
H, n_k, n_q = 16, 3072, 65536
q = torch.randn(n_q, H, 64, device=dev).half()
kv = torch.randn(n_k, H, 128, device=dev).half()
out = torch.empty(1, n_q, H * 64, dtype=torch.float16, device=dev)
for _ in range(3):
    tc.attention(q, kv[:, :, :64], kv[:, :, 64:], 1, out=out, q_shared=True, variant=variant)
torch.cuda.synchronize()
buf = np.zeros((9, 32, 8), dtype=np.int64)
lib = _lib.load()
rc = lib.foho_debug_attn_trace(C.c_void_p(buf.ctypes.data))
assert rc == 0, rc
t0 = buf[buf > 0].min()
rel = np.where(buf > 0, buf - t0, -1)
ev = ["enter", "s_full", "ld done", "max done", "p_empty", "exp done", "arrived", "in crit"]
order = [0, 1, 2, 3, 4, 7, 5, 6]
for role in range(8):
    print(f"group {'AB'[role // 4]} warp {role % 4}", [ev[i] for i in order])
    for j in range(3, 9):
        print(f"  j={j:2d}", " ".join(f"{int(rel[role, j, i]):7d}" for i in order))
print("MMA", ["top", "S_A issued", "S_B issued", "p_full A", "PV_A issued", "p_full B", "PV_B issued"])
for j in range(3, 9):
    print(f"  j={j:2d}", " ".join(f"{int(x):7d}" for x in rel[8, j, :7]))
json.dump(rel.tolist(), open(f"gpurun_out/attn_trace_v{variant}.json", "w"))
