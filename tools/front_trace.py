#!/usr/bin/env python
"""Timeline of the fused front kernel's pipeline roles in CTA 0 (debug build ab/lib_front_trace.so from tools/ablate_front.sh, -DFF_TRACE):
EIGB200_LIB=ab/lib_front_trace.so python tools/front_trace.py [B]  -- prints, per item, the clock64 stamps of TMA / converter / MMA / B-C prep / scan."""
import ctypes, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
f = tempfile.mktemp(); os.environ["FF_TRACE_PTR_FILE"] = f
import torch
import kbench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
fn, _, _ = kbench.case_front(B)
for _ in range(3):
    fn()
torch.cuda.synchronize()
ptr = int(open(f).read())
n = 8 * 256 * 4
buf = torch.empty(n, dtype=torch.int64, device="cuda")
ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n * 8), 3)
t = buf.cpu().view(8, 256, 4)
t0 = int(t[t > 0].min())
rel = lambda v: (int(v) - t0) if int(v) > 0 else -1
print("item = 4 * step + slot; cycles since the first stamp")
print("%5s | %-15s | %-31s" % ("item", "tma wait,issue", "mma start,op,dempty,done"))
for it in range(16, 48):
    print("%5d | %7d %7d | %7d %7d %7d %7d" % ((it,) + tuple(rel(t[0, it, k]) for k in range(2)) + tuple(rel(t[2, it, k]) for k in range(4))))
print("slot warps: chunk | start, converted next, accumulators ready, chunk end")
for c in range(4, 16):
    print("%3d | " % c + " | ".join("%7d %7d %7d %7d" % tuple(rel(t[4 + s, c, k]) for k in range(4)) for s in range(4)))
