#!/usr/bin/env python
"""How much of the reference's output depends on heap addresses (ORBextractor.cpp:684)?

Runs the reference build (oracle/_ref, strict: reference + S1 S2) on seeded images twice: with list nodes served in
creation order (fresh monotone heap = the oracle's / GPU's tie-break B2) and with the process heap (glibc malloc),
and prints per (image, level) how many selected keypoints differ as a set and whether the list order is the same.
CPU only; needs /root/reference or a prebuilt oracle/_ref.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O, ref as R          # noqa: E402
from tests._util import textured                  # noqa: E402


def main():
    cases = (((1000, 1000), 2000, 1), ((1800, 1000), 2000, 2), ((600, 900), 1000, 3), ((1000, 1000), 2000, 4),
             ((1200, 1500), 2000, 5), ((1800, 1000), 2000, 6))
    tot = diff = lev = lev_set = lev_order = 0
    for shape, nf, seed in cases:
        img = textured(shape[0], shape[1], seed)
        ko, _ = O.Extractor(nf)(img)
        er = R.Extractor(nf, strict=True)
        R.set_modes(heap_monotone=True, libm_a5=False, strict=True)
        km, _ = er(img)
        assert km.tobytes() == ko.tobytes(), "monotone heap must equal the oracle"
        R.set_modes(heap_monotone=False, libm_a5=False, strict=True)
        kg, _ = er(img)
        R.set_modes(heap_monotone=True, libm_a5=False, strict=True)
        for l in range(6):
            a, b = ko[ko["octave"] == l], kg[kg["octave"] == l]
            sa = {(float(k["x"]), float(k["y"])) for k in a}
            sb = {(float(k["x"]), float(k["y"])) for k in b}
            same_order = len(a) == len(b) and a.tobytes() == b.tobytes()
            tot += len(sa); diff += len(sa - sb); lev += 1; lev_set += sa != sb; lev_order += not same_order
    print(json.dumps(dict(tool="ref_tiebreak_stats", images=len(cases), keypoints=tot, keypoints_differing_as_a_set=diff,
                          levels=lev, levels_with_a_different_set=int(lev_set), levels_with_a_different_order=int(lev_order),
                          note="process heap (glibc) vs creation-order addresses; the process-heap result itself varies from call to call")))


if __name__ == "__main__":
    main()
