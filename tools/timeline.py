"""clock64() timeline of the fusion kernel's first tiles on cluster 0 (tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
_spec = importlib.util.spec_from_file_location("_lrn_build", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pointnet_refine_b200", "build.py"))
_b = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(_b)
os.environ["LRN_B200_LIB"] = _b.build(timeline=True)      # the product library carries no stamps: use the tuning build
import torch
import pointnet_refine_b200 as prb
from pointnet_refine_b200 import _lib
from oracle import synth
dev = torch.device("cuda:0")
m = prb.LineRefineNet().to(dev).eval()
m.load_state_dict(synth.to_torch(synth.make_state_dict(0)))
m.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ctx = torch.randn(148, 4096, 4, device=dev)
with torch.no_grad():
    for _ in range(3): m.context_encoder.run_native(ctx, pool=True)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    _lib.lib.lrn_debug_timeline(buf.data_ptr())
    m.context_encoder.run_native(ctx, pool=True)
    torch.cuda.synchronize()
    _lib.lib.lrn_debug_timeline(None)
t = buf.cpu()[:128].view(16, 8)
t0 = int(t[0, 0])
print("stamps per tile (cycles since first tile start): fusion = [tile start, last MMA issued | G ready, phase A done, F ready, phase B done]; "
      "LRN_DBG_LAYER=k (conv k) = [tile start, last MMA issued | epilogue start, staging free, acc ready, acc drained]")
if os.environ.get("LRN_DBG_LAYER") == "4":
    print("chain kernel: [MMA tile start, conv4 issued | conv2 acc ready, conv2 drained, conv3 acc ready, conv3 drained, next conv1 done, conv4 drained]")
    for i in range(16):
        r = [int(x) - t0 for x in t[i, :8]]
        print(f"{i:3d} " + " ".join(f"{x:8d}" for x in r) + f"   E1={r[3]-r[2]} wait3={r[4]-r[3]} E2={r[5]-r[4]} embed={r[6]-r[5]} E3={r[7]-r[6]} tile={r[7]-r[2]}")
    f = buf.cpu()
    if int(f[128 + 8]) != 0:
        z = int(f[128 + 8])
        print("conv5 of tile 1 (cycles since feat4-in-TMEM): per chunk n: [acc free (MMA), chunk issued (MMA) | staging free, acc ready, drained (epilogue)]")
        for n in range(8):
            print(f"  n={n}: {int(f[128+16+n])-z:7d} {int(f[128+n])-z:7d} | {int(f[128+24+n])-z:7d} {int(f[128+32+n])-z:7d} {int(f[128+40+n])-z:7d}")
    sys.exit(0)
for i in range(16):
    r = [int(x) - t0 for x in t[i, :6]]
    print(f"{i:3d} {r[0]:9d} {r[1]:9d} | {r[2]:9d} {r[3]:9d} {r[4]:9d} {r[5]:9d}   main={r[4]-r[2]:6d} phaseA={r[3]-r[2]:6d} phaseB={r[5]-r[4]:6d} tile={r[5]-r[0]:6d}")
