import sys, torch, numpy as np
sys.path.insert(0, 'gr-ofdm_tools_b200'); sys.path.insert(0, 'tests')
import common as cm
from test_next_rows import FORWARD_OOB as B, FEEDBACK_OOB as A
phy = cm.make_phy(cm.cfg_c1())
for n, ns in ((1 << 26, 1), (1 << 16, 4096)):
    x = torch.randn(ns, n, 2, device='cuda').view(torch.complex64) if False else torch.view_as_complex(torch.randn(ns, n, 2, device='cuda'))
    out = torch.empty_like(x)
    for span in (0, 2048, 4096, 8192, 16384, 65536, 1 << 30):
        if span == 1 << 30 and ns == 1: continue
        st = None
        for it in range(2):
            _, st = phy.iir_ccd(x if ns > 1 else x[0], B, A, state=None, span=span, out=out if ns > 1 else out[0])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for it in range(3):
            phy.iir_ccd(x if ns > 1 else x[0], B, A, state=st, span=span, out=out if ns > 1 else out[0])
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print("streams %d n %d span %d: %.3f ms, %.1f Gsamples/s, %.1f GB/s algorithmic" % (ns, n, span, ms, ns * n / ms / 1e6, 16 * ns * n / ms / 1e6), flush=True)
x = torch.view_as_complex(torch.randn(1 << 26, 2, device='cuda'))
for it in range(3): r = phy.papr(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for it in range(5): r = phy.papr(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print("papr n %d: %.3f ms, %.1f GB/s" % (x.numel(), ms, 8 * x.numel() / ms / 1e6), r.tolist())
