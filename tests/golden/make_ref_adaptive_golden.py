"""Regenerates tests/golden/ref_adaptive_golden.json from oracle/_ref: outputs of THE REFERENCE'S OWN
binarize{NativeAdaptive,AT,AGT,PureAdaptiveGaussian}.cpp (compiled unmodified against the cv:: facade, OpenCV primitives
executed by the cv2 wheel) -- the pins of SURVEY.md section 8 row F4.

Run HERE (build container, needs /root/reference):  make -C oracle _ref/_prl_ref.so && python tests/golden/make_ref_adaptive_golden.py
Inputs: seeded noise (BGR / BGRA), synthpage-v2 pages and the colour crops of tests/golden/real_pages.npz."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import prl_oracle as O  # noqa: E402  (page generator only)
from oracle import ref as R  # noqa: E402

import cv2  # noqa: E402

# name -> (function, positional args after the image, keyword args)
CALLS = {
    "at_m5_b19_s9": ("binarizeAT", (5, 255, 19, 9), {}),
    "at_m3_b7_s-2_max180": ("binarizeAT", (3, 180, 7, -2), {}),
    "agt_m5_b19_s9": ("binarizeAGT", (5, 255, 19, 9), {}),
    "agt_m7_b35_s3": ("binarizeAGT", (7, 255, 35, 3), {}),
    "pag_b15_s4": ("binarizePureAdaptiveGaussian", (255, 15, 4), {}),
    "native_defaults": ("binarizeNativeAdaptive", (), {}),
    "native_gaussblur": ("binarizeNativeAdaptive", (), {"isGaussianBlurReqiured": True}),
    "native_meanC_autoblock": ("binarizeNativeAdaptive", (), {"isAdaptiveThresholdCalculatedByGaussian": False, "adaptiveThresholdingBlockSize": 0}),
    "native_m3_shift-2.5_max77.6": ("binarizeNativeAdaptive", (), {"medianBlurKernelSize": 3, "adaptiveThresholdingShift": -2.5,
                                                                     "adaptiveThresholdingMaxValue": 77.6}),
    # the optional last step: cv::bilateralFilter of the mask (binarizeNativeAdaptive.cpp:116-134)
    "native_bilateral5": ("binarizeNativeAdaptive", (), {"bilateralFilterBlockSize": 5}),
    "native_bilateral9_c40_s3_max180": ("binarizeNativeAdaptive", (), {"bilateralFilterBlockSize": 9, "bilateralFilterColorSigma": 40.0,
                                                                         "bilateralFilterSpaceSigma": 3.0, "adaptiveThresholdingMaxValue": 180}),
    "native_bilateral3_meanC": ("binarizeNativeAdaptive", (), {"bilateralFilterBlockSize": 3, "isAdaptiveThresholdCalculatedByGaussian": False,
                                                                 "bilateralFilterColorSigma": 25.0, "bilateralFilterSpaceSigma": 0.9}),
    "native_bilateral_sigma0": ("binarizeNativeAdaptive", (), {"bilateralFilterBlockSize": 7, "bilateralFilterSpaceSigma": 0.0}),
    "native_bilateral_evenblock": ("binarizeNativeAdaptive", (), {"bilateralFilterBlockSize": 7, "bilateralFilterColorSigma": -1.0,
                                                                    "adaptiveThresholdingBlockSize": 20}),
}
GRAY_OK = [k for k in CALLS if k.startswith("native")]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def images():
    rng = np.random.default_rng(11)
    out = {"noise_bgr_120x150": rng.integers(0, 256, (120, 150, 3), dtype=np.uint8),
           "dark_bgr_90x131": (rng.integers(0, 256, (90, 131, 3)) // 6).astype(np.uint8),
           "noise_bgra_64x70": rng.integers(0, 256, (64, 70, 4), dtype=np.uint8),
           "synth_gray_300x421": O.synth_page(3, 300, 421),
           "a4_p2": O.synth_page(2)}
    pages = dict(np.load(os.path.join(HERE, "real_pages.npz")))
    for k, v in pages.items():
        if v.ndim == 3:
            out[k] = v
    for k in list(pages)[:3]:
        if pages[k].ndim == 2:
            out[k] = pages[k]
    return out


def outcome(fn, img, args, kw):
    try:
        return sha(getattr(R, fn)(img, *args, **kw))
    except ValueError:
        return "invalid_argument"
    except cv2.error:
        return "cv::Exception"


def main():
    assert R.available(), "build oracle/_ref first"
    G = {"generator": "oracle/_ref (reference C++ compiled unmodified; OpenCV = cv2 wheel)", "cv2": cv2.__version__,
         "calls": {k: [v[0], list(v[1]), v[2]] for k, v in CALLS.items()}, "images": {}}
    for key, img in images().items():
        names = list(CALLS) if key != "a4_p2" else ["native_defaults", "native_meanC_autoblock", "native_bilateral5"]
        G["images"][key] = {"shape": list(img.shape), "sha1": sha(img),
                            "out": {n: outcome(CALLS[n][0], img, CALLS[n][1], CALLS[n][2]) for n in names}}
    with open(os.path.join(HERE, "ref_adaptive_golden.json"), "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote ref_adaptive_golden.json:", len(G["images"]), "images")


if __name__ == "__main__":
    main()
