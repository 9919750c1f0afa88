"""CPU tests of the host side: C-ABI exports, facade constructors / error behaviour, generators,
payload_source / payload_sink, and loud failure without a GPU."""
import os
import re
import threading

import numpy as np
import pytest

import common as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ofdm_tools import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "ofdmx.h")).read()
    declared = set(re.findall(r"\b(ofdmx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ofdmx_ctx"}
    lib = _lib.load()
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.ofdmx_abi_version() == 4 == _lib.ABI_VERSION
    assert lib.ofdmx_profile_slots() >= 10
    assert lib.ofdmx_profile_name(1) == b"plateau_kernel"
    names = {lib.ofdmx_profile_name(i) for i in range(lib.ofdmx_profile_slots())}
    assert {b"rx_framew_kernel", b"sync_metric_fast_kernel", b"sync_metric_tma_kernel"} <= names


def test_struct_layouts():
    import ctypes as C
    from ofdm_tools import _lib, phy
    assert phy.FRAME_DTYPE.itemsize == 32
    assert C.sizeof(_lib.Counts) == 16
    assert _lib.Params.fft_len.offset == 0 and _lib.Params.occ_sizes.offset == 16
    # the binding's structs against the sizes the library itself was compiled with, and against the C header
    # compiled here by gcc (field by field)
    lib = _lib.load()
    assert lib.ofdmx_params_size() == C.sizeof(_lib.Params)
    assert lib.ofdmx_frame_size() == phy.FRAME_DTYPE.itemsize
    import subprocess
    import tempfile
    fields = [n for n, _ in _lib.Params._fields_]
    prog = "#include <stdio.h>\n#include <stddef.h>\n#include \"ofdmx.h\"\nint main(void){printf(\"%zu\", sizeof(ofdmx_params));" \
        + "".join('printf(" %%zu", offsetof(ofdmx_params, %s));' % f for f in fields) \
        + 'printf(" %zu %zu", sizeof(ofdmx_frame), sizeof(ofdmx_counts)); return 0;}'
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "l.c"), os.path.join(td, "l")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    assert vals[0] == C.sizeof(_lib.Params)
    assert vals[1:1 + len(fields)] == [getattr(_lib.Params, f).offset for f in fields]
    assert vals[-2:] == [32, 16]


def test_integration_md_stub_matches_binding():
    """The ctypes stub INTEGRATION.md shows a maintainer is generated from _lib.Params: every field, in order."""
    from ofdm_tools import _lib
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = txt[txt.index("class Params(C.Structure)"):]
    block = block[:block.index("```")]
    names = re.findall(r'\("([a-z0-9_]+)",', block)
    assert names == [n for n, _ in _lib.Params._fields_]
    assert "OFDMX_ABI_VERSION" in txt or "ofdmx_abi_version" in txt
    assert "not built" not in txt


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    phy = cm.make_phy(cm.cfg_c1())
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        phy.header_len()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gr-ofdm_tools_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for bad in ("import oracle", "from oracle", "ofdm_oracle.h", "liboracle", "orc_"):
                    assert bad not in txt, (f, bad)


def test_generators_match_reference_literals():
    from ofdm_tools import ofdm_txrx_modules as m, ofdm_cr_tools as t, ofdm_radio_hier
    c = lambda v: np.array([complex(a, b) for a, b in v])
    g = cm.GOLD["grc_ofdm_rx_hier"]
    assert np.allclose(m._make_sync_word1(64, m._def_occupied_carriers, m._def_pilot_carriers), c(g["sync_word1"]), atol=1e-8)
    assert np.array_equal(m._make_sync_word2(64, m._def_occupied_carriers, m._def_pilot_carriers), c(g["sync_word2"]))
    assert list(m._def_occupied_carriers[0]) == cm.OCC64[0]
    assert [tuple(x) for x in m._def_pilot_symbols] == [tuple(x) for x in cm.PLS64]
    d = cm.GOLD["ofdm_radio_hier_defaults"]
    occ, pil, pls, s1, s2 = t.spectrum_enforcer(128, [], 10)
    assert list(occ[0]) == d["occupied_carriers"][0] and np.array_equal(np.array(s1), c(d["sync_word1"]))
    r = ofdm_radio_hier()
    assert r.fft_len == 128 and r.cp_len == 32
    assert np.array_equal(np.array(r.sync_word1), np.real(c(d["sync_word1"])))
    assert list(r.occupied_carriers[0]) == d["occupied_carriers"][0]
    assert r.phy.params.max_carr_offset == 3 and r.phy.params.scramble_header == 1
    assert abs(r.phy.params.tx_scale - 0.01) < 1e-9


def test_facade_constructor_errors():
    from ofdm_tools import ofdm_tx, ofdm_rx, ofdm_radio_hier, ofdm_tx_rx_hier
    with pytest.raises(ValueError, match="Length of sync sequence"):
        ofdm_tx(sync_word1=[0] * 63)
    with pytest.raises(ValueError, match="Length of sync sequence"):
        ofdm_rx(sync_word2=[0] * 10)
    with pytest.raises(ValueError, match="Modulation not supported"):
        ofdm_tx(bps_payload=5)
    with pytest.raises(ValueError, match="Modulation not supported"):
        ofdm_radio_hier(payload_mod="qam1024")
    h = ofdm_tx_rx_hier(fft_len=64, payload_bps=2)
    assert h.ofdm_tx.cp_len == 16 and h.get_len_tag_key() == "packet_len"
    assert h.ofdm_rx.frame_length_tag_key == "frame_rx_len"
    assert h.ofdm_tx.phy.params.scramble_seed == 0 and h.ofdm_rx.phy.params.bps_payload == 2
    t = ofdm_tx(scramble_bits=True)
    assert t.scramble_seed == 0x7F and t.phy.params.scramble_header == 1


def test_create_rejects_bad_params_without_gpu():
    """Parameter validation happens before any CUDA call, so ValueError surfaces on a CPU box too."""
    from ofdm_tools import OfdmPhy
    cfg = cm.cfg_c1()
    with pytest.raises(ValueError):
        OfdmPhy(**dict(cfg, fft_len=48, sync_word1=[0] * 48, sync_word2=[0] * 48)).header_len()
    with pytest.raises(ValueError):
        OfdmPhy(**dict(cfg, bps_payload=5)).header_len()
    with pytest.raises(ValueError):
        OfdmPhy(**dict(cfg, max_pkt_bytes=5000)).header_len()


def test_payload_source_and_sink():
    from ofdm_tools import payload_source, payload_sink
    src = payload_source(packet_len=10)
    src.send_pkt_s("0123456789abc")
    src.send_pkt_s(b"defghij")
    assert src.pop_packets() == [b"0123456789", b"abcdefghij"]      # cut every packet_len bytes
    assert src.pop_packets() == []
    src.send_pkt_s(eof=True)
    assert src.eof() and src.get_packet_len() == 10
    got, ev = [], threading.Event()

    def cb(p):
        got.append(p)
        if len(got) == 2:
            ev.set()
    snk = payload_sink(cb)
    snk.deliver([b"one", b"two"])
    assert ev.wait(5.0) and got == [b"one", b"two"]


def test_segment_plan_and_stream_shards():
    from ofdm_tools import dist
    assert [list(dist.shard_streams(10, r, 4)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]
    plan = dist.plan_segments(1000000, 4, 1024, 72, 9864)
    assert plan[0][:3] == (0, 0, 250000) and plan[3][2] == 1000000 and plan[3][3] == 1000000
    for (l0, a, b, l1), nxt in zip(plan, plan[1:]):
        assert b == nxt[1] and l1 >= b + 9864 + 1096 and nxt[0] == nxt[1] - (1024 + 2 * 72 + 2)


def test_spectrum_translator_matches_the_reference_formula():
    """spectrum_translator (python/ofdm_cr_tools.py:455-469) against an independent evaluation of the reference's
    formula, and through spectrum_enforcer: the constrained bins disappear from the carrier plan."""
    from ofdm_tools import ofdm_cr_tools as T
    fc, sf, n = 2.4e9, 1.0e6, 128
    hz = [fc + 120e3, fc - 310e3, fc + 499e3]
    got = T.spectrum_translator(hz, fc, sf, n, 8)
    grid = np.linspace(-n / 2, n / 2 - 1, n) * (sf / n) + fc
    want = []
    for f in hz:
        b = (grid[np.argmin(np.abs(grid - f))] - fc) / (sf / n)
        for x in range(4):
            want += [b + x, b - x]
    assert got == want and len(got) == 24
    occ, pil, pls, s1, s2 = T.spectrum_enforcer(n, got, 10)
    assert not (set(int(b) for b in got) & set(occ[0])) and len(s1) == len(s2) == n
    assert T.spectrum_translator([], fc, sf, n, 8) == []
    assert T.find_nearest_l([1, 5, 9], 6) == 5


def test_single_sync_word_mode_oracle_roundtrip():
    """sync_word2=() (python/ofdm_txrx_modules.py:174-183,311-329): one preamble symbol, two OFDM symbols before
    the payload, carrier offset from the |Y[k]-Y[k+2]|^2 correlation, taps from sync word 1 (interpolated).  The
    oracle decodes its own bursts on a channel with a fractional and an integer carrier offset; the facades accept
    the empty word and size their frames accordingly.

    Two properties of the upstream algorithm shape the configurations: the taps of carriers outside sync word 1's
    span stay zero (so the plans below end on odd carriers, which sync word 1 covers), and the difference metric is
    blind when the FFT window sits fft_len/8 inside the prefix (cp_len/2 = 8 at fft_len 64: a quarter turn between
    carriers k and k+2), so the fft_len 64 case pins max_carr_offset to 0 and the integer offsets are exercised at
    fft_len 1024 / cp_len 72."""
    from ofdm_tools import ofdm_tx, ofdm_rx
    rng = np.random.default_rng(8)
    occ64 = [[k for k in cm.OCC64[0] if abs(k) != 26]]
    occ1k = [[k for k in cm.cfg_c3()["occupied_carriers"][0] if abs(k) != 302]]
    for cfg, plen, cfo, kw in ((cm.cfg_c1(2, True, 1, sync_word2=(), occupied_carriers=occ64, max_carr_offset=0), 96, 0.15, {}),
                               (cm.cfg_c3(sync_word2=(), occupied_carriers=occ1k), 700, 2.15, dict(fft_len=1024)),
                               (cm.cfg_c3(sync_word2=(), occupied_carriers=occ1k), 700, -1.9, dict(fft_len=1024))):
        orc = cm.make_oracle(cfg)
        D = cfg["fft_len"] + cfg["cp_len"]
        assert orc.n_sync_words == 1 and orc.frame_samples(plen) % D == 0
        two = cm.make_oracle(dict(cfg, sync_word2=None))
        assert orc.frame_samples(plen) == two.frame_samples(plen) - D          # one preamble symbol less
        pk = cm.rand_packets(rng, 3, plen)
        s, off = orc.tx(pk)
        x = cm.channel(cm.split_frames(s, off), rng, gaps=(100, 500), snr_db=30.0, cfo=cfo, lead=300, tail=3000, **kw)
        ref = orc.rx(x, want_z=False)
        assert orc.payloads(ref) == pk
        assert np.all(ref["frames"]["carr_offset"] == int(round(cfo / 2.0)) * 2)
    t = ofdm_tx(sync_word2=())
    assert len(t.sync_words) == 1 and t.phy.n_sync_words == 1 and not t.phy.params.sync_word2
    r = ofdm_rx(sync_word2=())
    assert r.phy.n_sync_words == 1
    with pytest.raises(ValueError, match="Length of sync sequence"):
        ofdm_rx(sync_word2=[0] * 10)
