import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gr-ofdm_tools_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, common as cm, oracle as O
from ofdm_tools import OfdmPhy
dev = torch.device("cuda", 0)
C = dict(bench.config_table()[2], frames=int(sys.argv[1]) if len(sys.argv) > 1 else 8192)
phy = OfdmPhy(device=0, tx_scale=0.01, max_pkt_bytes=1504, **C["cfg"])
x, payload, starts, FS = bench.make_streams(phy, C, 1, 17, dev)
print("n", x.numel(), "rms", float((x.abs() ** 2).mean().sqrt()))
r0 = phy.rx(x)
print("frames without agc", len(r0.frames), "crc ok", int(np.count_nonzero(r0.frames["flags"] & 2)))
y, g = phy.agc2(x)
torch.cuda.synchronize()
ref, gr = O.agc2(x[0].cpu().numpy())
yh = y[0].cpu().numpy()
eq = yh == ref
print("agc equal", bool(eq.all()), "first diff", int(np.argmin(eq)) if not eq.all() else -1, "gain", float(g[0]), gr)
r1 = phy.rx(y)
print("frames with agc", len(r1.frames), "crc ok", int(np.count_nonzero(r1.frames["flags"] & 2)), "triggers", r1.n_triggers)
ro = cm.make_oracle(C["cfg"]).rx(ref[: 40 * FS], want_z=False, byte_stride=1520)
print("oracle frames on agc output (first 40)", len(ro["frames"]), ro["frames"]["trigger"][:5], r1.frames["trigger"][:5])
print("rms out", float(np.sqrt(np.mean(np.abs(ref[100000:200000]) ** 2))))
