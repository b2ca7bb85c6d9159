import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import prlib_b200
from oracle import prl_oracle as O, c_oracle as CO
img = np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
ctx = prlib_b200.Context(0)
tw, th = 7, 5
got = ctx.otsu_tiles(img, tw, th); want = O.otsu_tiles(img, tw, th)
bad = np.argwhere(got != want)
print("mismatching px", len(bad))
seen = set()
for y, x in bad[:2000]:
    t = (y // th, x // tw)
    if t in seen: continue
    seen.add(t)
    if len(seen) > 6: break
    y0, x0 = t[0] * th, t[1] * tw
    tile = img[y0:y0 + th, x0:x0 + tw]
    hist = np.bincount(tile.ravel(), minlength=256)
    g = got[y0:y0 + th, x0:x0 + tw]
    # implied gpu threshold range: max pixel with 0, min pixel with 255
    z = tile[g == 0]; o = tile[g == 255]
    print("tile", t, tile.shape, "cv2", O.otsu_threshold_cv(np.ascontiguousarray(tile)), "py", O.otsu_threshold_from_hist(hist), "C", CO.otsu_from_hist(hist),
          "gpu_rect", ctx.otsu_threshold(np.ascontiguousarray(tile)), "gpu implied: max0", z.max() if z.size else None, "min255", o.min() if o.size else None)
    print(sorted(tile.ravel().tolist()))
