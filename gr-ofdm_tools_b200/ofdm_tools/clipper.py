"""B200 drop-in for python/clipper.py: `clipper(clipping_factor)` rails the real and imaginary parts to
+-clipping_factor (complex_to_float -> analog.rail_ff x2 -> float_to_complex, :45-58).

Inside the OFDM transmitter the clipper is fused into the TX kernel (OfdmPhy(tx_clip=...), used by
ofdm_radio_hier(clipper_mode=1)); this class is the stand-alone block for an arbitrary cuda complex64 tensor
(an elementwise clamp: torch is the right tool, there is no hot kernel to write)."""


class clipper(object):
    def __init__(self, clipping_factor):
        self.clipping_factor = clipping_factor

    def work(self, samples):
        import torch
        v = torch.view_as_real(samples)
        return torch.view_as_complex(v.clamp(-self.clipping_factor, self.clipping_factor))

    def get_clipping_factor(self):
        return self.clipping_factor

    def set_clipping_factor(self, clipping_factor):
        self.clipping_factor = clipping_factor
