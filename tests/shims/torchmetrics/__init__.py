"""Stand-in for `torchmetrics` (absent offline): the reference builds an LPIPS estimator at import time for its
evaluation code, which the loop test does not call."""
