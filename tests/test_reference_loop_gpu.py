"""The reference's own Python loop, UNMODIFIED (oracle/_ref/egg = byte copy of /root/reference/src + configs), on the
drop-in and on the reference's own native build: EGGFusion.reconstruct = Tracker.tracking -> preprocess ->
Mapping.mapping (Renderer.render, fusion kernels, frame_batch_optimization) -> postprocess, >= 10 frames of a synthetic
RGB-D sequence at the Replica (1200x680) and TUM fr1 (640x480, SH degree 0) calibrations (BASELINE configs 2 and 5).
Both arms must track the sequence, build the same map and render the same model maps (tests/ref_loop.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def _run(arm, config, frames, out):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_loop.py"), "--arm", arm, "--config", config, "--frames",
           str(frames), "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return np.load(out)


@pytest.mark.parametrize("config", ["replica", "tum"])
def test_reference_loop_runs_unchanged_on_the_dropin(config, tmp_path):
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "egg", "src")):
        pytest.skip("oracle/_ref/egg not built (oracle/build_ref.sh needs /root/reference)")
    frames = 12
    ours = _run("ours", config, frames, str(tmp_path / "ours.npz"))
    ref = _run("reference", config, frames, str(tmp_path / "ref.npz"))
    # both arms track the synthetic trajectory (ground truth known exactly) ...
    assert float(ours["ate"]) < 2e-3 and float(ref["ate"]) < 2e-3, (float(ours["ate"]), float(ref["ate"]))
    # ... and agree with each other: poses, map size, the model map rendered at the last frame
    dpos = np.abs(ours["est"][:, :3, 3] - ref["est"][:, :3, 3]).max()
    drot = np.abs(ours["est"][:, :3, :3] - ref["est"][:, :3, :3]).max()
    assert dpos < 1e-3 and drot < 1e-3, (dpos, drot)
    n_o, n_r = int(ours["n_surfels"]), int(ref["n_surfels"])
    assert abs(n_o - n_r) <= 0.01 * n_r, (n_o, n_r)        # surfel sampling is randomised (torch.rand) per run
    both = (ours["opacity"][..., 0] > 0.8) & (ref["opacity"][..., 0] > 0.8)
    assert both.mean() > 0.9
    ddepth = np.abs(ours["depth"][..., 0] - ref["depth"][..., 0])[both]
    assert np.median(ddepth) < 1e-3 and ddepth.mean() < 2e-3, (np.median(ddepth), ddepth.mean())
    mse = np.mean((ours["color"] - ref["color"])[both] ** 2)
    psnr = 10.0 * np.log10(1.0 / max(mse, 1e-12))
    assert psnr > 35.0, psnr      # both arms draw the same sampling sequence (seeded in tests/ref_loop.py) until a count differs
    print("loop %s: ATE ours %.2e ref %.2e m, pose diff %.1e, surfels %d / %d, render PSNR(ours, ref) %.1f dB, "
          "ms/frame ours %.1f ref %.1f" % (config, float(ours["ate"]), float(ref["ate"]), dpos, n_o, n_r, psnr,
                                           1e3 * ours["wall"][1:].mean(), 1e3 * ref["wall"][1:].mean()))
