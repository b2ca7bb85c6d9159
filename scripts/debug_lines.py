import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import prlib_b200
from oracle import prl_oracle as O
from test_edges import _page_with_rules
for shape, seed in (((700, 2000), 3), ((700, 1984), 3), ((700, 2016), 3), ((700, 1600), 3), ((300, 2000), 5)):
    img = _page_with_rules(shape[0], shape[1], seed)
    a = prlib_b200.removeLines(img); b = O.removeLines(img)
    d = np.argwhere(a != b)
    print(shape, "Lh", shape[1] // 50, "Lv", shape[0] // 50, "mismatches", len(d), "rows", sorted(set(d[:, 0].tolist()))[:8], "cols", (int(d[:, 1].min()), int(d[:, 1].max())) if len(d) else None, flush=True)
