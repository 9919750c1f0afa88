"""Kernel-level timing of ofdmx_agc2 on one stream of the headline workload: python tools/prof_agc.py [frames] """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gr-ofdm_tools_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
sys.path.insert(0, ROOT)
import torch
import bench
from ofdm_tools import OfdmPhy
dev = torch.device("cuda", 0)
C = dict(bench.config_table()[2], frames=int(sys.argv[1]) if len(sys.argv) > 1 else 65536)
phy = OfdmPhy(device=0, tx_scale=0.01, max_pkt_bytes=1504, **C["cfg"])
x, payload, starts, FS = bench.make_streams(phy, C, 1, 17, dev)
print("n", x.numel(), "rms", float((x[0, :1 << 20].abs() ** 2).mean().sqrt()))
for span in (os.environ.get("OFDMX_AGC_SPAN", "default"),):
    y, g = phy.agc2(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    y, g = phy.agc2(x)
    b.record()
    torch.cuda.synchronize()
    print("span", span, "agc2 ms", a.elapsed_time(b))
    phy.profile(True)
    y, g = phy.agc2(x)
    torch.cuda.synchronize()
    print(phy.profile_read())
    phy.profile(False)
