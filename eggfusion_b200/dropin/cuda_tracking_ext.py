"""Drop-in module with the import name of the reference's torch extension `cuda_tracking_ext`
(/root/reference/src/utils/cuda/src/tracking.cu:952-962), backed by libeggsplat.so (include/eggtrack.h).

Put `eggfusion_b200/dropin` on sys.path and /root/reference/src/utils/cuda/__init__.py works unchanged: same nine
function names, same out-tensor calling style (Python pre-allocates, native code fills).  Everything is launched on
the current stream with no device synchronisation; solve_block_cuda stays on the device.
"""
import os
import sys

import torch

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from eggfusion_b200 import _lib  # noqa: E402


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _in(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"cuda_tracking_ext: `{name}` must be a CUDA tensor (there is no CPU implementation)")
    return t.contiguous().float()


def _out(t, name):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
        raise RuntimeError(f"cuda_tracking_ext: output `{name}` must be a contiguous float32 CUDA tensor")
    return t


def compute_vertex_and_normal_cuda(depth_image, fx, fy, cx, cy, vertex_map, normal_map):
    d = _in(depth_image, "depth_image")
    ht, wd = d.shape[0], d.shape[1]
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().egt_vertex_normal_map(d.data_ptr(), float(fx), float(fy), float(cx), float(cy),
                                                     _out(vertex_map, "vertex_map").data_ptr(),
                                                     _out(normal_map, "normal_map").data_ptr(), wd, ht, _stream(d)),
                   "vertex_normal_map")


def gaussian_filter_cuda(input_image, output_image, wd, ht, channels, window_size, sigma_s):
    x = _in(input_image, "input_image")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().egt_gaussian_filter(x.data_ptr(), _out(output_image, "output_image").data_ptr(), int(wd),
                                                   int(ht), int(channels), int(window_size), float(sigma_s),
                                                   _stream(x)), "gaussian_filter")


def bilateral_filter_cuda(input_image, output_image, wd, ht, window_size, sigma_c, sigma_s):
    x = _in(input_image, "input_image")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().egt_bilateral_filter(x.data_ptr(), _out(output_image, "output_image").data_ptr(),
                                                    int(wd), int(ht), int(window_size), float(sigma_c), float(sigma_s),
                                                    _stream(x)), "bilateral_filter")


def gaussian_downsample_cuda(input_image, output_image, wd, ht, ch):
    x = _in(input_image, "input_image")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().egt_gaussian_downsample(x.data_ptr(), _out(output_image, "output_image").data_ptr(),
                                                       int(wd), int(ht), int(ch), _stream(x)), "gaussian_downsample")


def compute_gradients_cuda(input_image, grad_x, grad_y, wd, ht):
    x = _in(input_image, "input_image")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().egt_compute_gradients(x.data_ptr(), _out(grad_x, "grad_x").data_ptr(),
                                                     _out(grad_y, "grad_y").data_ptr(), int(wd), int(ht), _stream(x)),
                   "compute_gradients")


def solve_block_cuda(A, b, lm, x):
    """x <- solve((A + lm I) x = b).  The reference copies A and b to the host and runs Eigen's QR there."""
    A_, b_ = _in(A, "A"), _in(b, "b")
    n = A_.shape[0]
    with torch.cuda.device(A_.device):
        _lib.check(_lib.load().egt_solve_block(A_.data_ptr(), b_.data_ptr(), float(lm), _out(x, "x").data_ptr(), int(n),
                                               _stream(A_)), "solve_block")


def _dead(name, why):
    def f(*_a, **_k):
        raise NotImplementedError(f"cuda_tracking_ext.{name}: {why}")
    f.__name__ = name
    return f


# Exports that exist in the reference but are dead or non-functional there: the tracker imports the
# PyTorch implementations from src/core/optimizer.py instead (tracker.py:15-19), the reduction launch of the two
# optimisation kernels is commented out (tracking.cu:335-341, :505-511) so they always return zero Hcc / gc.
projective_transform_cuda = _dead("projective_transform_cuda", "dead code in the reference (tracker.py:15-19)")
rgb_optimization_cuda = _dead("rgb_optimization_cuda", "non-functional in the reference (tracking.cu:335-341)")
icp_optimization_cuda = _dead("icp_optimization_cuda", "non-functional in the reference (tracking.cu:505-511)")
