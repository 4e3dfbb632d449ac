import os, sys, torch
sys.path.insert(0, os.getcwd())
import wmb200
from wmb200 import _lib as L, ops
B, T = 4096, 16000
gen = wmb200.Generator(16).to("cuda").eval()
blob = gen.packed(); lib = L.load()
x = ops.to_planar(torch.randn(B, T, 64, device="cuda")); y = torch.empty_like(x)
img = blob[L.G_TC:]; b1 = blob[L.G_RB0 + L.RB_B1:]; b2 = blob[L.G_RB0 + L.RB_B2:]
st = torch.cuda.current_stream().cuda_stream
hb = hasattr(lib, "wm_resblock_tc_hostbias_fwd") and os.environ.get("HB", "1") == "1"
h1, h2 = b1[:64].cpu().contiguous(), b2[:64].cpu().contiguous()
for _ in range(3):
    if hb: L.check(lib.wm_resblock_tc_hostbias_fwd(x.data_ptr(), img.data_ptr(), h1.data_ptr(), h2.data_ptr(), y.data_ptr(), None, B, T, st), "rb")
    else: L.check(lib.wm_resblock_tc_fwd(x.data_ptr(), img.data_ptr(), b1.data_ptr(), b2.data_ptr(), y.data_ptr(), None, B, T, st), "rb")
torch.cuda.synchronize()
