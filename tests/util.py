import hashlib

import numpy as np

CONFIGS = {  # name -> (method, window, params)   (same table as tests/golden/make_golden.py)
    "sauvola_w15_k0.2": (0, 15, (0.2,)),
    "niblack_w15_k-0.2": (1, 15, (-0.2,)),
    "wolfjolion_w15_k0.5": (2, 15, (0.5,)),
    "nick_w15_k-0.1": (3, 15, (-0.1,)),
    "feng_w21_default": (4, 21, (0.75, 0.2, 0.03, 2.0)),
    "sauvola_w101_k0.01": (0, 101, (0.01,)),
    "nick_w101_k-0.1": (3, 101, (-0.1,)),
    "feng_w101_default": (4, 101, (0.75, 0.2, 0.03, 2.0)),
    "wolfjolion_w101_k0.01": (2, 101, (0.01,)),
    "niblack_w101_k0.01": (1, 101, (0.01,)),
}


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
