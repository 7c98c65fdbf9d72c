"""Seeded random parity hunt: CUDA extractor vs the CPU oracle over random frame sizes and parameters.
usage: python tools/fuzz_extract.py [ncases] [seed] [max_w] [max_h]      (needs a GPU; test infrastructure, not product)
Prints one line per failing case (the tuple can be pasted into tests/test_gpu_extract_parity.py) and a summary."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np

from oracle import oracle
from orbslam_in_practice_b200 import _lib as orbx
from orbslam_in_practice_b200.synth import synth_frame
import test_gpu_extract_parity as T


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    max_w = int(sys.argv[3]) if len(sys.argv) > 3 else 1300
    max_h = int(sys.argv[4]) if len(sys.argv) > 4 else 900
    rng = np.random.default_rng(seed)
    fails, done, t0 = 0, 0, time.time()
    while done < n:
        w, h = int(rng.integers(64, max_w + 1)), int(rng.integers(64, max_h + 1))
        sf = float(rng.choice([1.05, 1.1, 1.2, 1.25, 1.3, 1.41, 1.5, 1.7, 2.0, 2.5, 3.0, 3.3]))
        nl = int(rng.integers(1, 13))
        if min(w, h) / sf ** (nl - 1) < 64:
            continue
        ini = int(rng.integers(3, 60)); mn = int(rng.integers(1, 40))
        nf = int(rng.integers(10, 4001))
        case = (w, h, nf, sf, nl, ini, mn, int(rng.integers(0, 1 << 30)))
        done += 1
        try:
            params = dict(nfeatures=nf, scale_factor=sf, nlevels=nl, ini_th=ini, min_th=mn)
            ex = orbx.Extractor(max_width=w, max_height=h, max_batch=1, **params)
            oex = oracle.OracleExtractor(**params)
            if case[7] & 1:
                ex.set_pyramid_border(True)
            kind = case[7] % 5
            img = synth_frame(case[7] % 1000, w, h)
            if kind == 3:
                img = np.random.default_rng(case[7]).integers(0, 256, (h, w), dtype=np.uint8)      # saturated cells
            kps, desc, counts = ex.extract_host(img)
            T._compare_frame(oracle, ex, oex, img, 0, kps, desc, counts)
            ex.close()
        except AssertionError as e:
            fails += 1
            print("FAIL", case, str(e)[:120], flush=True)
        except Exception as e:                                  # unsupported geometry is reported, not hidden
            print("ERR ", case, type(e).__name__, str(e)[:120], flush=True)
    print("cases %d, failures %d, %.1f s" % (done, fails, time.time() - t0))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
