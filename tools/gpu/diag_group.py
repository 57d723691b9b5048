import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import pytest
from fastselect_b200 import _native as native
from oracle import ref_oracle as R
from datasets import mixed
from test_gpu_group import EmulatedGroup

class MP:
    def setenv(self, k, v): os.environ[k] = v
    def delenv(self, k): os.environ.pop(k, None)

x, y = mixed(41, 640, 90, 3)
x32, recip, isd = R.multisurf_prep(x, 10)
yc = np.unique(y, return_inverse=True)[1].astype(np.int32)
cp = (np.bincount(yc) / yc.size).astype(np.float32)
print("isd", isd.astype(int).tolist())
for with_x in (False, True):
    for fs in ("1", "0"):
        os.environ["FS_B200_FEATURE_SHARD"] = fs
        with native.Dataset(x32, yc, 3) as plain:
            plain.set_features(isd, recip, native.FS_ARITH_F32)
            want = plain.score(native.FS_MULTISURF, use_star=True)
        g = EmulatedGroup(native, x32, yc, 3, 2, with_x, MP())
        try:
            g.set_features(isd, recip, native.FS_ARITH_F32)
            got = g.score(native.FS_MULTISURF, use_star=True)
            for r in range(2):
                bad = np.flatnonzero(~np.isclose(got[r], want, rtol=1e-9, atol=1e-9))
                print(f"with_x={with_x} fshard={fs} rank{r}: {bad.size} bad cols", bad.tolist()[:60])
                if bad.size:
                    print("   got ", got[r][bad[:6]], "\n   want", want[bad[:6]])
        finally:
            g.close()
