import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/shims")
import torch
from drivescenegen_b200.hostapi import UNet2DModel
REF_CFG = dict(in_channels=3, out_channels=3, layers_per_block=2, block_out_channels=(64, 128, 256, 512),
               down_block_types=("DownBlock2D",) * 4, up_block_types=("UpBlock2D",) * 4)
size=int(sys.argv[1]); b=int(sys.argv[2])
torch.manual_seed(0)
m = UNet2DModel(sample_size=size, **REF_CFG).to("cuda").eval()
x = torch.randn(b, 3, size, size, device="cuda")
eng = m.engine()
prog = eng.program(b, size, size)
tf = torch.full((b,), 10.0, device="cuda")
out = torch.empty_like(x)
import ctypes as C
prog.in_ptr = C.c_void_p(x.data_ptr()); prog.t_ptr = C.c_void_p(tf.data_ptr()); prog.out_ptr = C.c_void_p(out.data_ptr())
st = torch.cuda.current_stream().cuda_stream
for i, (op, (name, meta)) in enumerate(zip(prog.ops, prog.op_info)):
    op(st)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED at op", i, name, meta, str(e)[:200]); sys.exit(1)
print("ok", out.float().abs().mean().item())
