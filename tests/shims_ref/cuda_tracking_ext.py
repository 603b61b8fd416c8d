"""`cuda_tracking_ext` for the REFERENCE arm of the loop test: the reference's own compiled module (oracle/_ref, built
by oracle/build_ref.sh from /root/reference/src/utils/cuda) re-exported under its import name, with ONE substitution:
`solve_block_cuda`.  The reference solves the 6x6 system on the CPU with Eigen's colPivHouseholderQr
(src/utils/cuda/src/tracking.cu:929-950); Eigen is neither vendored nor installed, so the reference build compiles
that host function against a stub and it cannot be called.  Here it takes the same GPU -> CPU -> GPU round trip with
torch.linalg.lstsq on the damped system.  Test infrastructure only."""
import glob
import importlib.util
import os

import torch

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_so = glob.glob(os.path.join(_root, "oracle", "_ref", "cuda_tracking_ext*.so"))
if not _so:
    raise ImportError("oracle/_ref/cuda_tracking_ext*.so not built (oracle/build_ref.sh)")
_spec = importlib.util.spec_from_file_location("cuda_tracking_ext", _so[0])
_ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ref)
for _n in dir(_ref):
    if not _n.startswith("__"):
        globals()[_n] = getattr(_ref, _n)


def solve_block_cuda(A, b, lm, x):
    A_cpu, b_cpu = A.detach().float().cpu(), b.detach().float().cpu().reshape(-1, 1)
    n = A_cpu.shape[0]
    sol = torch.linalg.lstsq(A_cpu + float(lm) * torch.eye(n), b_cpu).solution
    x.view(-1).copy_(sol.reshape(-1).to(x.device))
