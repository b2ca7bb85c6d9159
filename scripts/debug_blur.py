import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import prlib_b200
from oracle import prl_oracle as O
ctx = prlib_b200.Context(0)
noise = np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
def model(src, kk):
    k = len(kk); r = k // 2
    p = cv2.copyMakeBorder(src, r, r, r, r, cv2.BORDER_REFLECT_101).astype(np.int64)
    h = sum(p[:, i:i + src.shape[1]] * kk[i] for i in range(k))
    v = sum(h[j:j + src.shape[0], :] * kk[j] for j in range(k))
    return np.clip((v + 32768) >> 16, 0, 255).astype(np.uint8)
print("cv2 threads", cv2.getNumThreads(), "cpu features:", [l for l in cv2.getBuildInformation().splitlines() if "Dispatched" in l or "Baseline" in l][:3])
for shape in ((7, 300), (7, 304), (8, 300), (7, 256), (7, 640)):
    img = np.ascontiguousarray(noise[:shape[0], :shape[1]])
    m = model(img, [16, 64, 96, 64, 16])
    for rep in range(2):
        a = ctx.gaussian_blur(img, 5, 0); b = O.gaussian_blur(img, 5, 0)
        cv2.setNumThreads(1); b1 = O.gaussian_blur(img, 5, 0); cv2.setNumThreads(-1)
        print(shape, "gpu!=model", int((a != m).sum()), " cv2!=model", int((b != m).sum()), " cv2(1 thread)!=model", int((b1 != m).sum()), flush=True)
