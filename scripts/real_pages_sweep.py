"""Every replaced function on full real pages (a directory of PNGs, e.g. a copy of the reference's test_data/binarize),
GPU vs the OpenCV oracle.  usage: real_pages_sweep.py DIR"""
import sys, os, glob, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import prlib_b200
from oracle import prl_oracle as O
bad = 0; n = 0
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.png"))):
    img = cv2.imread(f)                      # BGR, as the samples read it
    if img is None: continue
    res = []
    for name, gpu, ref in (("Sauvola", prlib_b200.binarizeSauvola, O.binarizeSauvola), ("Niblack", prlib_b200.binarizeNiblack, O.binarizeNiblack),
                           ("WolfJolion", prlib_b200.binarizeWolfJolion, O.binarizeWolfJolion), ("NICK", prlib_b200.binarizeNICK, O.binarizeNICK),
                           ("Feng", prlib_b200.binarizeFeng, O.binarizeFeng)):
        try:
            a = gpu(img.copy()); b = ref(img.copy()); ok = np.array_equal(a, b)
        except Exception as e:
            ok = f"EXC {type(e).__name__}: {e}"
        res.append((name, ok))
    for name, gpu, ref in (("Sauvola15", lambda im: prlib_b200.binarizeSauvola(im, 15, 0.2, 0), lambda im: O.binarizeSauvola(im, 15, 0.2, 0)),
                           ("LocalOtsu", prlib_b200.binarizeLocalOtsu, O.binarizeLocalOtsu),
                           ("LocalOtsuCLAHE", lambda im: prlib_b200.binarizeLocalOtsu(im, 255.0, 2.0), lambda im: O.binarizeLocalOtsu(im, 255.0, 2.0)),
                           ("removeLines", prlib_b200.removeLines, O.removeLines)):
        try:
            a = gpu(img.copy()); b = ref(img.copy()); ok = np.array_equal(a, b)
        except Exception as e:
            ok = f"EXC {type(e).__name__}: {e}"
        res.append((name, ok))
    n += len(res); bad += sum(1 for _, ok in res if ok is not True)
    print(os.path.basename(f), img.shape, " ".join(f"{k}={'ok' if v is True else v}" for k, v in res), flush=True)
print("cases", n, "not identical", bad)
