import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import prlib_b200
from oracle import c_oracle as CO
what = sys.argv[1]
ctx = prlib_b200.Context(0)
img = CO.synth_page(0, 700, 1000)
if what == "k1_generic":
    ctx.set_option("disable_tma", 1)
    S, Q = ctx.integral(img, 7); Sw, Qw = CO.integrals_int64(img, 7); print(what, np.array_equal(S, Sw), np.array_equal(Q, Qw))
elif what == "k1_tma":
    S, Q = ctx.integral(img, 7); Sw, Qw = CO.integrals_int64(img, 7); print(what, np.array_equal(S, Sw), np.array_equal(Q, Qw))
    if not np.array_equal(S, Sw):
        bad = np.argwhere(S != Sw); print(len(bad), bad[:5], S[tuple(bad[0])], Sw[tuple(bad[0])])
elif what == "k2_exact":
    ctx.set_option("disable_tma", 1); ctx.set_option("exact_threshold", 1)
    m = ctx.binarize_local(img, 0, 15, (0.2,), 0); print(what, np.array_equal(m, CO.binarize_local(img, 0, 15, (0.2,), 0)))
elif what == "k2_fast":
    ctx.set_option("disable_tma", 1)
    m = ctx.binarize_local(img, 0, 15, (0.2,), 0); w = CO.binarize_local(img, 0, 15, (0.2,), 0); print(what, np.array_equal(m, w), (m != w).sum())
