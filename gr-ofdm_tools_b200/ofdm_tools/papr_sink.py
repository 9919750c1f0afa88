"""B200 drop-in for python/papr_sink.py: `papr_sink(block_len)`.

The reference block keeps the first block_len samples of every work() call (:41-44) and level() returns
max(x * conj(x)) / (vdot(x, x) / len(x)) of that block (:46-54).  Here work() takes a cuda complex64 tensor
and level() runs the reduction on the GPU (ofdmx_papr: one pass over the block, per-CTA partials, a
one-warp final); `phy` is any OfdmPhy (it owns the C-ABI context the kernel is launched through)."""


class papr_sink(object):
    def __init__(self, block_len=512, phy=None):
        self.block_len = block_len
        self.papr = 0
        self.vct_data = None
        self.phy = phy

    def work(self, samples):
        in0 = samples[0:self.block_len]
        self.vct_data = in0
        return len(in0)

    def set_papr(self, measure):
        if self.phy is None:
            raise RuntimeError("papr_sink needs an OfdmPhy (phy=...) to run on")
        self.papr = float(self.phy.papr(measure.contiguous())[0].item())

    def level(self):
        self.set_papr(self.vct_data)
        return self.papr

    def set_block_len(self, block_len):
        self.block_len = block_len

    def get_block_len(self):
        return self.block_len
