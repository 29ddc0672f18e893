#!/usr/bin/env python
"""Generates tests/golden/golden.json and golden_III_b2.npz FROM THE REFERENCE ITSELF.

Run in the dev container only (needs /root/reference and oracle/_ref/libqtref.so, which
oracle/Makefile builds from the unmodified reference sources):

    python tests/golden/make_golden.py

Sources of truth:
  * table literals: parsed out of /root/reference/constants.h (all eight arrays, host and
    __constant__ copies), sha256 over little-endian u32;
  * transforms / products / Nussbaumer: the reference's own CPU functions (radix2NTTGS, radix2INTT,
    radix2NTT, radix2INTTGS, radix2NTTStock, radix2INTTStock, nussbaumer_fft) called through
    oracle/ref_wrap.cu on the xorshift64 stream of SURVEY.md 8c-3;
  * parameter sets the reference has no code for (I, p-I, p-III): the O(n^2) schoolbook only —
    recorded as "pinned_by": "schoolbook".
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle, Reference, SET_I, SET_III, SET_P_I, SET_P_III, SET_NAMES  # noqa: E402

REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint32).tobytes()).hexdigest()


def parse_constants():
    text = open(os.path.join(REF, "constants.h")).read()
    out = {}
    for m in re.finditer(r"(?:__constant__\s+)?uint32_t\s+(\w+)\s*\[NTTSIZE\]\s*=\s*\{([^}]*)\}", text):
        vals = np.array([int(v) for v in re.findall(r"\d+", m.group(2))], dtype=np.uint32)
        out[m.group(1)] = vals
    return out


def main():
    o, r = Oracle(), Reference()
    g = {"generator": "tests/golden/make_golden.py", "prng": "xorshift64 state0=88172645463325252, x[i]=out%q then y[i]=out%q"}
    consts = parse_constants()
    assert set(consts) == {"bitrev_tbl", "bitrev_tbl_gpu", "Phi", "Phi_gpu", "invPhi", "invPhi_gpu", "tf0_gpu", "ti0_gpu"}
    assert all(v.size == 1024 for v in consts.values())
    g["constants_h_sha256"] = {k: sha(v) for k, v in consts.items()}
    g["constants_h_first4"] = {k: [int(x) for x in v[:4]] for k, v in consts.items()}
    # reference host tables as linked (must equal the literals)
    for i, k in enumerate(("bitrev_tbl", "Phi", "invPhi")):
        assert np.array_equal(r.table(i), consts[k])
    assert np.array_equal(r.table(3), consts["tf0_gpu"]) and np.array_equal(r.table(4), consts["ti0_gpu"])

    q = 8404993
    x, y, _ = o.xorshift_pair(q, 2 * 1024)
    z = [r.polymul(x, y, v, 1) for v in range(4)]
    assert all(np.array_equal(z[0], zz) for zz in z), "reference CPU variants disagree"
    fwd = r.forward(x)
    assert np.array_equal(r.inverse(fwd), x)
    np.savez_compressed(os.path.join(HERE, "golden_III_b2.npz"), x=x, y=y, z=z[0], fwd_x=fwd)
    g["III_random_b2"] = {"x_first4": [int(v) for v in x[:4]], "z_first4": [int(v) for v in z[0][:4]],
                          "z_row1_first4": [int(v) for v in z[0][1024:1028]], "z_sha256": sha(z[0]),
                          "fwd_first4": [int(v) for v in fwd[:4]], "fwd_sha256": sha(fwd),
                          "pinned_by": "reference CPU GS->CT, GS/GS, CT/CT, Stockham (all equal)"}
    ones = np.ones(2048, np.uint32)
    zo = r.polymul(ones, ones, 0, 1)
    g["III_all_ones"] = {"z_first3": [int(v) for v in zo[:3]], "z_sha256": sha(zo), "rule": "z[k]=(2k+2-n) mod q"}
    x1, y1, _ = o.xorshift_pair(q, 1024)
    zn = r.nussbaumer(x1, y1)
    g["III_nussbaumer_b1"] = {"z_first4": [int(v) for v in zn[:4]], "z_sha256": sha(zn), "pinned_by": "nussbaumer_fft NTT.cu:167-277"}
    zno = r.nussbaumer(ones[:1024], ones[:1024])
    g["III_nussbaumer_all_ones"] = {"z_first3": [int(v) for v in zno[:3]], "z_510_512": [int(v) for v in zno[510:513]], "z_sha256": sha(zno)}
    # ramp operand of init_operand (NTT.cu:11,15)
    ramp = np.zeros(2048, np.uint32)
    for b in range(2):
        ramp[b * 1024: b * 1024 + 512] = 512 - np.arange(512)
    g["III_ramp_forward_sha256"] = sha(r.forward(ramp))
    g["III_ramp_square_sha256"] = sha(r.polymul(ramp, ramp, 0, 1))
    for s in (SET_I, SET_P_I, SET_P_III):
        p = o.params(s)
        xx, yy, _ = o.xorshift_pair(p.q, p.n)
        zz = o.schoolbook(s, xx, yy)
        g[SET_NAMES[s] + "_random_b1"] = {"n": int(p.n), "q": int(p.q), "psi": int(p.psi), "z_first4": [int(v) for v in zz[:4]],
                                         "z_sha256": sha(zz), "pinned_by": "schoolbook"}
        zr = o.ring_schoolbook(p.n, xx, yy)
        zr = np.where(zr == 0xFFFFFFFF, 0, zr).astype(np.uint32)
        g[SET_NAMES[s] + "_ring_schoolbook_b1_sha256"] = sha(zr)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote golden.json and golden_III_b2.npz")


if __name__ == "__main__":
    main()
