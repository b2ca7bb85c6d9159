"""Regenerates tests/golden/ref_golden.json (+ real_pages.npz) from oracle/_ref, i.e. from THE REFERENCE'S OWN C++
(binarize*.cpp, removeLines.cpp, imageLibCommon.cpp compiled unmodified against the cv:: facade, OpenCV primitives
executed by the cv2 wheel).  These are "outputs of the reference itself run here": the pins of the parity claim.

Run HERE (build container, needs /root/reference):  make -C oracle _ref/_prl_ref.so && python tests/golden/make_ref_golden.py

Inputs: the noise page, synthpage-v2 pages (A4 p0/p1, A3-600 p0), degenerate pages, and crops of 24 of the reference's
own test_data/binarize pages (gray through the reference's cvtColor step, plus BGR / BGRA crops fed as they are).
The crops are committed (real_pages.npz) because the GPU box has no /root/reference.
"""
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


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


# name -> (function, args after the image) exactly as passed to prl::<function>
CALLS = {
    "sauvola_w15_k0.2": ("binarizeSauvola", (15, 0.2, 0)),
    "niblack_w15_k-0.2": ("binarizeNiblack", (15, -0.2, 0)),
    "wolfjolion_w15_k0.5": ("binarizeWolfJolion", (15, 0.5, 0)),
    "nick_w15_k-0.1": ("binarizeNICK", (15, -0.1, 0)),
    "feng_w21_default": ("binarizeFeng", (21, 0.75, 0.2, 0.03, 2.0, 0)),
    "sauvola_w101_k0.01": ("binarizeSauvola", (101, 0.01, 0)),
    "nick_w101_k-0.1": ("binarizeNICK", (101, -0.1, 0)),
    "feng_w101_default": ("binarizeFeng", (101, 0.75, 0.2, 0.03, 2.0, 0)),
    "wolfjolion_w101_k0.01": ("binarizeWolfJolion", (101, 0.01, 0)),
    "niblack_w101_k0.01": ("binarizeNiblack", (101, 0.01, 0)),
    # the header defaults, morphology tail included (binarizeSauvola.h:43-47 ...)
    "sauvola_defaults": ("binarizeSauvola", (101, 0.01, 2)),
    "niblack_defaults": ("binarizeNiblack", (101, 0.01, 2)),
    "wolfjolion_defaults": ("binarizeWolfJolion", (101, 0.01, 2)),
    "nick_defaults": ("binarizeNICK", (21, -0.01, 0)),
    "feng_defaults": ("binarizeFeng", (21, 0.75, 0.2, 0.03, 2.0, 2)),
    "sauvola_w15_k0.2_morph2": ("binarizeSauvola", (15, 0.2, 2)),
    "sauvola_w15_k0.2_morph-1": ("binarizeSauvola", (15, 0.2, -1)),
    "nick_w31_k-0.2_morph-2": ("binarizeNICK", (31, -0.2, -2)),
}
SMALL = [k for k in CALLS if "w101" not in k and "defaults" not in k] + ["nick_defaults", "feng_defaults"]


def run(name, img):
    fn, args = CALLS[name]
    out, after = getattr(R, fn)(img, *args, return_input=True)
    return {"shape": list(out.shape), "sha1": sha(out), "white": float((out == 255).mean()),
            "input_after": {"shape": list(after.shape), "sha1": sha(after)}}


def entry(img, names, extras=True):
    e = {"shape": list(img.shape), "sha1": sha(img), "masks": {n: run(n, img) for n in names}}
    if extras and min(img.shape[:2]) >= 50:
        e["removeLines"] = sha(R.removeLines(img))
        try:
            e["localOtsu"] = sha(R.binarizeLocalOtsu(img))
            e["localOtsu_clahe2"] = sha(R.binarizeLocalOtsu(img, 255.0, 2.0))
        except ValueError as ex:  # RemoveChildrenContours: no contours
            e["localOtsu"] = "ValueError"
    return e


def main():
    assert R.available(), "build oracle/_ref first"
    G = {"generator": "oracle/_ref (reference C++ compiled unmodified; OpenCV = cv2 wheel)", "cv2": cv2.__version__,
         "seed": 2024, "calls": {k: [v[0], list(v[1])] for k, v in CALLS.items()}, "images": {}}
    noise = np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
    G["images"]["noise_512x640"] = entry(noise, list(CALLS))
    G["images"]["a4_p0"] = entry(O.synth_page(0), list(CALLS))
    G["images"]["a4_p1"] = entry(O.synth_page(1), ["sauvola_w15_k0.2"], extras=False)
    G["images"]["a3_600_p0"] = entry(O.synth_page(0, 9921, 7016), ["nick_w101_k-0.1", "feng_w101_default"], extras=False)
    # degenerate pages: the NaN / zero-window branches
    deg = {"black_120x130": np.zeros((120, 130), np.uint8), "white_120x130": np.full((120, 130), 255, np.uint8),
           "const77_97x131": np.full((97, 131), 77, np.uint8)}
    half = np.zeros((140, 150), np.uint8); half[:, 75:] = 200; deg["halfblack_140x150"] = half
    sparse = np.zeros((150, 160), np.uint8); sparse[::37, ::41] = 255; deg["sparse_150x160"] = sparse
    for k, v in deg.items():
        G["images"][k] = entry(v, SMALL, extras=False)
    G["degenerate"] = {k: v.tolist() if v.size < 0 else list(v.shape) for k, v in deg.items()}
    # tiny / clamped windows: min(rows, cols) <= windowSize (w is clamped, may become even; WJ/NICK/Feng throw)
    tiny = np.random.default_rng(5).integers(0, 256, (12, 40), dtype=np.uint8)
    e = {"shape": list(tiny.shape), "sha1": sha(tiny), "masks": {}}
    for n in ("sauvola_w15_k0.2", "niblack_w15_k-0.2"):
        e["masks"][n] = run(n, tiny)
    for n in ("wolfjolion_w15_k0.5", "nick_w15_k-0.1", "feng_w21_default"):
        try:
            run(n, tiny)
            e["masks"][n] = "no exception"
        except cv2.error:
            e["masks"][n] = "cv2.error"
    G["images"]["tiny_12x40"] = e

    ref_dir = "/root/reference/test_data/binarize"
    crops = {}
    files = sorted(f for f in os.listdir(ref_dir) if f.endswith(".png") and os.path.getsize(os.path.join(ref_dir, f)) > 0)
    rng = np.random.default_rng(2024)
    picked = 0
    for fn in files[::5]:
        im = cv2.imread(os.path.join(ref_dir, fn), cv2.IMREAD_UNCHANGED)
        if im is None or im.shape[0] < 260 or im.shape[1] < 260 or im.dtype != np.uint8:
            continue
        hh = int(rng.integers(220, min(520, im.shape[0])))
        ww = int(rng.integers(220, min(640, im.shape[1])))
        y0 = int(rng.integers(0, im.shape[0] - hh + 1))
        x0 = int(rng.integers(0, im.shape[1] - ww + 1))
        crop = np.ascontiguousarray(im[y0:y0 + hh, x0:x0 + ww])
        key = "page_" + fn.split(".")[0]
        if crop.ndim == 3 and picked % 4 != 0:      # three of four colour pages go in as gray (the reference's own cvtColor)
            crop = cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY if crop.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)
        if crop.ndim == 3 and crop.shape[2] == 4:
            crop = np.ascontiguousarray(crop[:, :, :3])   # binarizeSauvola.cpp:51 uses COLOR_BGR2GRAY: 3 channels only
        crops[key] = crop
        G["images"][key] = entry(crop, SMALL + ["sauvola_defaults"])
        picked += 1
        if picked >= 24:
            break
    np.savez_compressed(os.path.join(HERE, "real_pages.npz"), **crops)
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote ref_golden.json:", len(G["images"]), "images,", picked, "real-page crops")


if __name__ == "__main__":
    main()
