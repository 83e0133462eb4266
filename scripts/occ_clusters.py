import sys; sys.path.insert(0, '.')
import b200ba, ctypes
from visual_slam_core_b200 import _lib, LocalBASolver
s = LocalBASolver()
lib = _lib.load()
for c in (1, 2, 4, 8):
    print("cluster", c, "max active clusters", lib.ba_debug_max_active_clusters(c))
