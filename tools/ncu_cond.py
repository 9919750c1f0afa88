import sys, torch
sys.path.insert(0, 'gr-ofdm_tools_b200'); sys.path.insert(0, 'tests')
import common as cm
from test_next_rows import FORWARD_OOB as B, FEEDBACK_OOB as A
phy = cm.make_phy(cm.cfg_c1())
x = torch.view_as_complex(torch.randn(1 << 26, 2, device='cuda'))
out = torch.empty_like(x)
phy.iir_ccd(x, B, A, out=out)
phy.papr(x)
torch.cuda.synchronize()
