"""Regenerates tests/golden/golden.json and the real-page crops in tests/golden/real_crops.npz.

Run HERE (build container): python tests/golden/make_golden.py
  * digests come from oracle/prl_oracle.py, i.e. from the real OpenCV primitives (cv2 4.13.0) executed
    in the reference's order -- the closest thing to "outputs of the reference itself" this image
    allows (the C++ library cannot be built: no OpenCV C++ SDK, no Leptonica);
  * the crops are cut from the reference's own test_data/binarize pages (inputs only; the reference
    ships no expected outputs) so that the GPU box, which has no /root/reference, still tests real pages.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import prl_oracle as O  # noqa: E402

import cv2  # noqa: E402


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


CONFIGS = {  # name -> (method, window, params)
    "sauvola_w15_k0.2": (O.SAUVOLA, 15, (0.2,)),
    "niblack_w15_k-0.2": (O.NIBLACK, 15, (-0.2,)),
    "wolfjolion_w15_k0.5": (O.WOLFJOLION, 15, (0.5,)),
    "nick_w15_k-0.1": (O.NICK, 15, (-0.1,)),
    "feng_w21_default": (O.FENG, 21, (0.75, 0.2, 0.03, 2.0)),
    "sauvola_w101_k0.01": (O.SAUVOLA, 101, (0.01,)),
    "nick_w101_k-0.1": (O.NICK, 101, (-0.1,)),
    "feng_w101_default": (O.FENG, 101, (0.75, 0.2, 0.03, 2.0)),
    "wolfjolion_w101_k0.01": (O.WOLFJOLION, 101, (0.01,)),
    "niblack_w101_k0.01": (O.NIBLACK, 101, (0.01,)),
}


def image_entry(img, configs, morphs=((O.SAUVOLA, 15, (0.2,), 2), (O.SAUVOLA, 15, (0.2,), -1))):
    e = {"shape": list(img.shape), "sha1": sha(img), "masks": {}, "t8": {}, "aux": {}}
    for name in configs:
        m, w, p = CONFIGS[name]
        out, aux = O.binarize_local(img, m, w, p, 0, return_aux=True)
        e["masks"][name] = {"shape": list(out.shape), "sha1": sha(out), "white": float((out == 255).mean())}
        e["t8"][name] = sha(aux["T8"])
        if "smax" in aux:
            e["aux"][name] = {"imin": aux["imin"], "smax": aux["smax"]}
    e["morph"] = {}
    for (m, w, p, it) in morphs:
        e["morph"][f"{O.METHOD_NAMES[m]}_w{w}_k{p[0]}_morph{it}"] = sha(O.binarize_local(img, m, w, p, it))
    thr, mask = O.otsu_global(img)
    e["otsu_global"] = {"thr": thr, "sha1": sha(mask)}
    e["otsu_tiles64"] = sha(O.otsu_tiles(img, 64, 64))
    return e


def main():
    G = {"cv2": cv2.__version__, "seed": 2024, "images": {}}
    small = ["sauvola_w15_k0.2", "niblack_w15_k-0.2", "wolfjolion_w15_k0.5", "nick_w15_k-0.1", "feng_w21_default",
             "sauvola_w101_k0.01", "nick_w101_k-0.1", "feng_w101_default", "wolfjolion_w101_k0.01", "niblack_w101_k0.01"]

    a4 = O.synth_page(0)
    e = image_entry(a4, small)
    S, Q = O.integrals_int64(a4, 7)
    e["integral_pad7"] = {"S_sha1": sha(S), "Q_sha1": sha(Q), "S_last": int(S[-1, -1]), "Q_last": int(Q[-1, -1])}
    e["morph"]["sauvola_w101_k0.01_morph2"] = sha(O.binarize_local(a4, O.SAUVOLA, 101, (0.01,), 2))
    G["images"]["a4_p0"] = e

    a4p1 = O.synth_page(1)
    G["images"]["a4_p1"] = image_entry(a4p1, ["sauvola_w15_k0.2"], morphs=())

    a3 = O.synth_page(0, 9921, 7016)
    e = {"shape": list(a3.shape), "sha1": sha(a3), "masks": {}}
    for name in ("nick_w101_k-0.1", "feng_w101_default"):
        m, w, p = CONFIGS[name]
        out = O.binarize_local(a3, m, w, p, 0)
        e["masks"][name] = {"shape": list(out.shape), "sha1": sha(out), "white": float((out == 255).mean())}
    S, Q = O.integrals_int64(a3, 50)
    e["integral_pad50"] = {"S_sha1": sha(S), "Q_sha1": sha(Q), "S_last": int(S[-1, -1]), "Q_last": int(Q[-1, -1])}
    G["images"]["a3_600_p0"] = e

    noise = np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
    e = image_entry(noise, small)
    S, Q = O.integrals_int64(noise, 7)
    e["integral_pad7"] = {"S_sha1": sha(S), "Q_sha1": sha(Q), "S_last": int(S[-1, -1]), "Q_last": int(Q[-1, -1])}
    G["images"]["noise_512x640"] = e

    # real pages: crops of the reference's own inputs (gray via the reference's cvtColor step)
    ref_dir = "/root/reference/test_data/binarize"
    crops = {}
    if os.path.isdir(ref_dir):
        for fn, (y0, x0, hh, ww) in {"0037.png": (400, 300, 700, 900), "0018.png": (500, 200, 640, 801),
                                     "0064.png": (1000, 1500, 555, 777)}.items():
            im = cv2.imread(os.path.join(ref_dir, fn))
            if im is None:
                continue
            gray = O.to_gray(im)
            crop = np.ascontiguousarray(gray[y0:y0 + hh, x0:x0 + ww])
            key = "real_" + fn.split(".")[0]
            crops[key] = crop
            G["images"][key] = image_entry(crop, small[:6])
        bgr = cv2.imread(os.path.join(ref_dir, "0037.png"))[400:656, 300:620].copy()
        crops["bgr_0037"] = bgr
        G["bgr_0037_gray_sha1"] = sha(O.to_gray(bgr))
        np.savez_compressed(os.path.join(HERE, "real_crops.npz"), **crops)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote golden.json with", len(G["images"]), "images")


if __name__ == "__main__":
    main()
