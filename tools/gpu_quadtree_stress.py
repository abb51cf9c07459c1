"""Quadtree stress: clustered candidates that force the general (keyed) form, many shapes; CUDA vs oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diasss_b200 import binding as B
from oracle import oracle as O
from tests._util import kps_triples, textured

def clustered(rows, cols, seed, n_blobs=3, rad=40):
    g = np.random.default_rng(seed)
    img = np.full((rows, cols), 90, np.uint8)
    tex = textured(rows, cols, seed)
    yy, xx = np.mgrid[0:rows, 0:cols]
    m = np.zeros((rows, cols), bool)
    for _ in range(n_blobs):
        cy, cx = g.integers(30, rows - 30), g.integers(30, cols - 30)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 < rad * rad
    img[m] = tex[m]
    return img

bad = 0
for (r, c, seed, nf, nb, rad) in [(400, 500, 1, 2000, 2, 30), (600, 300, 2, 2000, 1, 20), (300, 900, 3, 1000, 3, 25), (500, 500, 4, 3000, 1, 60),
                                   (700, 700, 5, 2000, 4, 15), (1000, 400, 6, 500, 2, 50)]:
    img = clustered(r, c, seed, nb, rad)
    ctx = B.Context(nfeatures=nf)
    ex = O.Extractor(nf)
    ok, od = ex(img); gk, gd = ctx.extract(img)
    same = ok.tobytes() == gk.tobytes() and np.array_equal(od, gd)
    lv = [np.array_equal(kps_triples(ex.level_keys(l)), ctx.debug_level_keys(0, l)) for l in range(6)]
    print((r, c, seed, nf), 'n', len(ok), len(gk), 'equal', same, lv, 'cands', [len(ex.candidates(l)) for l in range(6)])
    bad += not same
    ctx.close()
print('FAILURES', bad)
sys.exit(1 if bad else 0)
