#!/usr/bin/env python3
"""Extracts the reference-supplied known-answer data for the OFDM PHY path from /root/reference
into tests/golden/reference_vectors.json.  Run in the authoring container only (the GPU box has
no /root/reference); the JSON it writes is committed.

Sources (relative to /root/reference):
  apps/ofdm_rx_hier.grc, apps/ofdm_tx_hier.grc  parameter blocks sync_word1 / sync_word2
  python/ofdm_radio_hier.py:34-38               default carrier plan + 128-pt sync words
  python/sync_radio_hier.py:50-56               narrow-band 64-pt plan + sync words
  python/ofdm_cr_tools.py:49-54                 same narrow-band literals (_sync_*)
  python/ofdm_txrx_modules.py:54-62             802.11a carrier plan + pilot polarity sequence
"""
import ast
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def cplx(v):
    return [[float(complex(x).real), float(complex(x).imag)] for x in v]


def grc_param(path, ident):
    txt = open(path).read()
    m = re.search(r"<key>id</key>\s*<value>%s</value>.*?<key>value</key>\s*<value>(.*?)</value>"
                  % re.escape(ident), txt, re.S)
    return ast.literal_eval(m.group(1).strip())


def py_literal(path, pattern):
    txt = open(path).read()
    m = re.search(pattern, txt, re.S)
    return ast.literal_eval(m.group(1).strip())


out = {}
for name in ("ofdm_rx_hier", "ofdm_tx_hier"):
    p = os.path.join(REF, "apps", name + ".grc")
    out["grc_%s" % name] = {
        "sync_word1": cplx(grc_param(p, "sync_word1")),
        "sync_word2": cplx(grc_param(p, "sync_word2")),
    }

rh = os.path.join(REF, "python", "ofdm_radio_hier.py")
out["ofdm_radio_hier_defaults"] = {
    "pilot_carriers": py_literal(rh, r"pilot_carriers=(\(\(.*?\),\)), pilot_symbols"),
    "pilot_symbols": py_literal(rh, r"pilot_symbols=(\(\(.*?\),\)),\s*occupied_carriers"),
    "occupied_carriers": py_literal(rh, r"occupied_carriers=(\(\[.*?\],\)),\s*samp_rate"),
    "sync_word1": cplx(py_literal(rh, r"sync_word1=(\[.*?\]),\s*sync_word2")),
    "sync_word2": cplx(py_literal(rh, r"sync_word2=(\[.*?\]),\s*scramble_mode")),
}

sh = os.path.join(REF, "python", "sync_radio_hier.py")
out["sync_radio_hier"] = {
    "sync_word1": cplx(py_literal(sh, r"sync_word1 = sync_word1 = (\[.*?\])\n")),
    "sync_word2": cplx(py_literal(sh, r"sync_word2 = sync_word2 = (\[.*?\])\n")),
    "pilot_symbols": py_literal(sh, r"pilot_symbols = pilot_symbols = (\(.*?\))\n"),
    "pilot_carriers": py_literal(sh, r"pilot_carriers = pilot_carriers = (\(.*?\))\n"),
    "occupied_carriers": py_literal(sh, r"occupied_carriers = occupied_carriers = (\(.*?\))\n"),
}

ct = os.path.join(REF, "python", "ofdm_cr_tools.py")
out["ofdm_cr_tools_sync"] = {
    "sync_word1": cplx(py_literal(ct, r"_sync_sync_word1 = (\[.*?\])\n")),
    "sync_word2": cplx(py_literal(ct, r"_sync_sync_word2 = (\[.*?\])\n")),
    "pilot_carriers_1024": py_literal(ct, r"_1024_pilot_carriers = (\(.*?\))\n"),
    "pilot_carriers_128": py_literal(ct, r"_128_pilot_carriers = (\(.*?\))\n"),
    "pilot_carriers_64": py_literal(ct, r"_64_pilot_carriers = (\(.*?\))\n"),
}

tm = os.path.join(REF, "python", "ofdm_txrx_modules.py")
seq = py_literal(tm, r"_pilot_sym_scramble_seq = (\(.*?\n\))")
out["ofdm_txrx_modules"] = {
    "pilot_sym_scramble_seq": list(seq),
    "def_pilot_carriers": py_literal(tm, r"_def_pilot_carriers=(\(\(.*?\),\))"),
    # python-2 "range(..)+range(..)" expression at :54, evaluated here
    "def_occupied_carriers": [list(range(-26, -21)) + list(range(-20, -7)) + list(range(-6, 0))
                              + list(range(1, 7)) + list(range(8, 21)) + list(range(22, 27))],
}
src54 = open(tm).read().splitlines()[53]
assert "range(-26, -21) + range(-20, -7) + range(-6, 0) + range(1, 7) + range(8, 21) + range(22, 27)" in src54

with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("wrote", os.path.join(HERE, "reference_vectors.json"),
      {k: list(v.keys()) for k, v in out.items()})
