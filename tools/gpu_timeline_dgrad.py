"""Per-role %globaltimer timeline of CTA 0 of conv2_dgrad (B=256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributedmnist_b200.ops.lib import check, load, ptr, stream_ptr
lib = load(); B = 256
dy = (torch.randn(B, 14, 14, 64, device="cuda") * 0.1).to(torch.bfloat16)
w = (torch.randn(800, 64, device="cuda") * 0.05).to(torch.bfloat16)
dx = torch.zeros(B, 14, 14, 32, dtype=torch.bfloat16, device="cuda")
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
for _ in range(3):
    check(lib.dm_conv2_dgrad_dbg(ptr(dy), ptr(w), ptr(dx), B, ptr(dbg), stream_ptr()), "dgrad")
torch.cuda.synchronize()
d = dbg.cpu().tolist(); t0 = d[0]
rel = lambda x: (x - t0) if x else None
print("start 0 | setup done", rel(d[1]), "| W resident", rel(d[2]), "| end", rel(d[3]))
print("producer issue times:", [rel(x) for x in d[8:28]])
print("mma data-landed times:", [rel(x) for x in d[32:52]])
print("epilogue acc-complete times:", [rel(x) for x in d[56:64]])
