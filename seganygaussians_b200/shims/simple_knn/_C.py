"""``simple_knn._C``: ``distCUDA2(points[P,3] cuda float32) -> Tensor[P]`` = mean squared distance to the 3 nearest other
points (ext.cpp / spatial.cu:14-25 / simple_knn.cu:146-219 of the reference's submodule).  Same distance expression and
summation order as the reference kernel; results agree with it to fp32 rounding (measured <= 4e-7 relative)."""
import os
import sys

try:
    from seganygaussians_b200.knn import dist_cuda2 as _dist_cuda2
except ImportError:   # shims/ was put on sys.path without the package root
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
    from seganygaussians_b200.knn import dist_cuda2 as _dist_cuda2


def distCUDA2(points):
    return _dist_cuda2(points)
