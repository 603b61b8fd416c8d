"""Stand-in for `pytorch_msssim` (absent offline): evaluation only."""


def ms_ssim(*a, **k):
    raise NotImplementedError("pytorch_msssim is not available offline")
