"""Every replaced function on full real pages (a directory of PNGs, e.g. a copy of the reference's test_data/binarize), GPU vs
the reference's own object code (oracle/_ref) where it travelled to this box, else the OpenCV call sequence (oracle/prl_oracle.py).
usage: real_pages_sweep.py DIR"""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import prlib_b200
from prlib_b200 import PrlCudaError
from oracle import prl_oracle as O
from oracle import ref as R
W = R if R.available() else O
print("checker:", "oracle/_ref (the reference's own C++)" if R.available() else "oracle/prl_oracle.py (cv2 port)")


def outcome(fn):
    try:
        return "ok", fn()
    except ValueError:
        return "invalid_argument", None
    except (cv2.error, PrlCudaError):
        return "cv::Exception", None


CASES = [
    ("Sauvola", lambda m, im: m.binarizeSauvola(im)), ("Niblack", lambda m, im: m.binarizeNiblack(im)),
    ("WolfJolion", lambda m, im: m.binarizeWolfJolion(im)), ("NICK", lambda m, im: m.binarizeNICK(im)), ("Feng", lambda m, im: m.binarizeFeng(im)),
    ("Sauvola15", lambda m, im: m.binarizeSauvola(im, 15, 0.2, 0)), ("Niblack15", lambda m, im: m.binarizeNiblack(im, 15, -0.2, 0)),
    ("NICK15m1", lambda m, im: m.binarizeNICK(im, 15, -0.1, 1)), ("WJ31", lambda m, im: m.binarizeWolfJolion(im, 31, 0.5, 0)),
    ("LocalOtsu", lambda m, im: m.binarizeLocalOtsu(im)), ("LocalOtsuCLAHE", lambda m, im: m.binarizeLocalOtsu(im, 255.0, 2.0)),
    ("removeLines", lambda m, im: m.removeLines(im)),
    ("NativeAdaptive", lambda m, im: m.binarizeNativeAdaptive(im)),
    ("NativeAdaptiveMeanGauss", lambda m, im: m.binarizeNativeAdaptive(im, True, 5, 7, 150.0, False)),
    ("AT", lambda m, im: m.binarizeAT(im, 5, 255, 19, 9)), ("AGT", lambda m, im: m.binarizeAGT(im, 5, 255, 19, 9)),
    ("PureAdaptiveGaussian", lambda m, im: m.binarizePureAdaptiveGaussian(im, 255, 15, 4)),
    ("NativeAdaptiveBilateral5", lambda m, im: m.binarizeNativeAdaptive(im, bilateralFilterBlockSize=5)),
    ("NativeAdaptiveBilateral9", lambda m, im: m.binarizeNativeAdaptive(im, bilateralFilterBlockSize=9, bilateralFilterColorSigma=40.0,
                                                                         bilateralFilterSpaceSigma=3.0)),
]
if len(sys.argv) > 2:                         # optional: only the named cases
    CASES = [c for c in CASES if c[0] in sys.argv[2].split(",")]
bad = 0; n = 0; exc = 0
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.png"))):
    img = cv2.imread(f)                      # BGR, as the samples read it
    if img is None: continue
    res = []
    for name, call in CASES:
        a = outcome(lambda: call(prlib_b200, img.copy()))
        b = outcome(lambda: call(W, img.copy()))
        same = a[0] == b[0] and (a[1] is None or np.array_equal(a[1], b[1]))
        res.append((name, "ok" if same and a[0] == "ok" else (f"both:{a[0]}" if same else f"DIFF gpu={a[0]} ref={b[0]}")))
    n += len(res); bad += sum(1 for _, v in res if v.startswith("DIFF")); exc += sum(1 for _, v in res if v.startswith("both"))
    print(os.path.basename(f), img.shape, " ".join(f"{k}={v}" for k, v in res), flush=True)
print("cases", n, "byte-identical", n - bad - exc, "same exception on both sides", exc, "DIFFERENT", bad)
