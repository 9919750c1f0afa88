import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("gr-ofdm_tools_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch, common as cm
rng = np.random.default_rng(1)
for which, roll in (("c1", 4), ("c3", 18)):
    cfg, plen = {"c3": (cm.cfg_c3(), 1500), "c1": (cm.cfg_c1(2, True, 1), 96)}[which]
    kw = dict(cfg, rolloff=roll, tx_scale=0.01, tx_clip=0.0)
    orc = cm.make_oracle(kw)
    pk = cm.rand_packets(rng, 2, plen)
    so, oo = orc.tx(pk)
    phy = cm.make_phy(kw)
    s, off = phy.tx(pk)
    s = s.cpu().numpy()
    D = cfg["fft_len"] + cfg["cp_len"]
    d = np.abs(s - so)
    print(which, roll, "offsets", off.cpu().numpy(), oo, "peak", np.abs(so).max(), "max diff", d.max(), "at", int(d.argmax()), "sym", int(d.argmax()) // D, "pos", int(d.argmax()) % D)
    for k in range(0, 5):
        print("  sym", k, "maxdiff", d[k * D:(k + 1) * D].max(), "first8", np.round(d[k * D:k * D + 8] / np.abs(so).max(), 6))
