"""Write openwork_b200/orb_pattern.npy: the 256 x 2 point pairs of ORB's learned rBRIEF test pattern
(OpenCV's `bit_pattern_31_`), as int8[512,2] (x, y).

The reference computes descriptors with `cv2.ORB_create().compute` (MonocularSlam/slam.py:113,119); that
pattern is part of OpenCV (Apache-2.0), not of /root/reference.  There is no network here, so the table is
read out of the installed opencv-python binary: it is located by its well-known first eight values
(8,-3, 9,5, 4,2, 7,-12) stored as little-endian int32.  tests/test_frontend.py re-checks the committed file
against the installed binary and, on the GPU, against real cv2 descriptors."""
import glob
import os
import struct
import sys

import numpy as np


def extract():
    import cv2
    so = glob.glob(os.path.join(os.path.dirname(cv2.__file__), "*.so"))
    for path in so:
        data = open(path, "rb").read()
        i = data.find(struct.pack("<8i", 8, -3, 9, 5, 4, 2, 7, -12))
        if i >= 0:
            tab = np.frombuffer(data[i:i + 256 * 4 * 4], dtype="<i4").reshape(512, 2)
            assert tab.min() >= -15 and tab.max() <= 15
            return tab.astype(np.int8)
    raise RuntimeError("ORB bit pattern not found in the cv2 binary")


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "openwork_b200", "orb_pattern.npy")
    tab = extract()
    np.save(out, tab)
    print("wrote", out, tab.shape, tab[:4].tolist())
