"""Latency of the drop-in single-image path (BASELINE config 1 flow: one page in host memory -> mask in host memory)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import prlib_b200
from oracle import c_oracle as CO, prl_oracle as O
page = CO.synth_page(0)
bgr = np.dstack([page, page, page])
for name, img in (("gray A4", page), ("BGR A4", bgr)):
    for w, k, morph in ((15, 0.2, 0), (101, 0.01, 2)):
        prlib_b200.binarizeSauvola(img, w, k, morph)
        t = []
        for _ in range(10):
            t0 = time.perf_counter(); out = prlib_b200.binarizeSauvola(img, w, k, morph); t.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); ref = O.binarizeSauvola(img, w, k, morph); tc = time.perf_counter() - t0
        print(f"{name} Sauvola w={w} morph={morph}: GPU path {1e3*min(t):.2f} ms (median {1e3*sorted(t)[5]:.2f}), cv2 1 core {1e3*tc:.0f} ms, equal={np.array_equal(out, ref)}")
