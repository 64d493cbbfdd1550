import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet_refine_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
a = torch.randn(128, 64, device=dev).bfloat16()
w = torch.randn(64, 64, device=dev).bfloat16()
out = torch.zeros(128, 64, device=dev)
_lib.check(_lib.lib.lrn_debug_ts_probe(a.data_ptr(), w.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream), "probe")
torch.cuda.synchronize()
ref = a.float() @ w.float().T
print("TS probe max err", (out - ref).abs().max().item(), "ref max", ref.abs().max().item())
