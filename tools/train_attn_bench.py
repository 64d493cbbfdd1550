"""Train-mode attention kernels alone (lrn_train_attention_forward / _backward): device time per call at the train step's
shape and the achieved share of the HBM bound (K / V read once, dK / dV written once).  usage: train_attn_bench.py [B N p]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet_refine_b200.train_ops import CrossAttnTrainFn, KVGradShare
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
dev = torch.device("cuda:0")
L = 6
q = torch.randn(B, 32, 256, device=dev, requires_grad=True)
kall = torch.randn(B, N, L, 8, 32, device=dev).bfloat16().requires_grad_()
vall = torch.randn(B, N, L, 8, 32, device=dev).bfloat16().requires_grad_()
r = torch.randn(B, 32, 256, device=dev)
share = KVGradShare()
def fwd(layer): return CrossAttnTrainFn.apply(q, kall[:, :, layer].transpose(1, 2), vall[:, :, layer].transpose(1, 2), p, share, layer)
for _ in range(2):
    out = fwd(0); out.backward(r); kall.grad = vall.grad = q.grad = None
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.synchronize()
tf = tb = 0.0
for layer in range(L):
    e[0].record(); out = fwd(layer); e[1].record()
    out.backward(r, inputs=[q]); e[2].record()       # dq only through autograd: dK / dV land in the shared buffers
    torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
kv_bytes = 2 * B * N * 256 * 2
print(json.dumps({"B": B, "N": N, "p_drop": p, "fwd_ms_per_layer": tf / L, "bwd_ms_per_layer": tb / L,
                  "fwd_GBps": kv_bytes / (tf / L * 1e-3) / 1e9, "bwd_GBps": 2 * kv_bytes / (tb / L * 1e-3) / 1e9}))
