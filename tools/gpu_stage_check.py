"""Verbose stage-by-stage comparison of the CUDA path with the oracle (debug aid; the tests assert the same)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from diasss_b200 import binding as B, synth
from diasss_b200.frontend import FrontEnd
from oracle import oracle as O

def check_extract(img, mask=None, nf=2000, tag=""):
    fe = FrontEnd(nfeatures=nf)
    ex = O.Extractor(nf)
    t = time.time(); ok, od = ex(img); to = time.time() - t
    t = time.time(); gk, gd = fe.ctx.extract(img); tg = time.time() - t
    rows, cols = img.shape
    allok = True
    for l in range(6):
        if l >= 1:
            a = ex.level_image(l); b = fe.ctx.debug_level_image(0, l, rows, cols)
            eq = np.array_equal(a, b)
            if not eq: print(tag, "level", l, "pyramid mismatch", (a != b).sum()); allok = False
        ca = ex.candidates(l); cb = fe.ctx.debug_candidates(0, l)
        if not np.array_equal(ca, cb):
            allok = False
            print(tag, "level", l, "candidates differ", len(ca), len(cb))
            n = min(len(ca), len(cb))
            d = np.nonzero((ca[:n] != cb[:n]).any(axis=1))[0]
            if len(d): print("   first diff at", d[0], ca[d[0]], cb[d[0]])
        ka = ex.level_keys(l); kb = fe.ctx.debug_level_keys(0, l)
        ka3 = np.stack([ka["x"], ka["y"], ka["response"]], 1).astype(np.int32)
        if not np.array_equal(ka3, kb):
            allok = False
            print(tag, "level", l, "keys differ", len(ka3), len(kb))
            n = min(len(ka3), len(kb))
            d = np.nonzero((ka3[:n] != kb[:n]).any(axis=1))[0]
            if len(d): print("   first diff at", d[0], ka3[d[0]], kb[d[0]], "ndiff", len(d))
    print(tag, img.shape, "n", len(ok), len(gk), "kps eq", ok.tobytes() == gk.tobytes(), "desc eq", np.array_equal(od, gd),
          "stages ok", allok, "t_oracle %.3f t_gpu %.3f" % (to, tg))
    if len(ok) == len(gk) and ok.tobytes() != gk.tobytes():
        for name in ok.dtype.names:
            d = np.nonzero(ok[name] != gk[name])[0]
            if len(d): print("   field", name, "ndiff", len(d), "first", d[0], ok[name][d[0]], gk[name][d[0]])
    if len(ok) == len(gk) and not np.array_equal(od, gd):
        d = np.nonzero((od != gd).any(axis=1))[0]
        print("   desc rows differing", len(d), d[:5])
    fe.ctx.close()

if __name__ == "__main__":
    print(B.lib().dsx_version())
    for (r, c, seed, nf) in [(300, 260, 1, 2000), (420, 640, 2, 500), (640, 300, 3, 2000), (250, 700, 4, 300), (1000, 500, 5, 2000)]:
        f = synth.make_survey(1, r, c, seed=seed)[0]
        check_extract(f["norm_img"], nf=nf, tag="synth")
    rng = np.random.default_rng(0)
    check_extract(rng.integers(0, 256, (200, 333), dtype=np.uint8), tag="noise")
    check_extract(np.zeros((128, 128), np.uint8), tag="zeros")
    check_extract(np.full((90, 400), 77, np.uint8), tag="flat")
