"""CPU checks of bench.py: the algorithmic-byte formula behind `roofline.achieved` reproduces SURVEY.md 8(d), and the
reference arm's no-GPU fallback (CPU oracle port) prints a line with the contract's keys."""
import json
import os
import subprocess
import sys

from util import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_survey_8d():
    # I = 3.5 P, P_vis = P, M = 16 at C3 -> A_fwd 732 MB + A_bwd 985 MB = 1.717 GB per frame
    P, N_px, M = 1_000_000, 1920 * 1080, 16
    b = bench.algorithmic_bytes(P, P, 3.5 * P, N_px, M)
    assert round(b["A_fwd"] / 1e6) == 732 and round(b["A_bwd"] / 1e6) == 985
    assert abs((b["A_fwd"] + b["A_bwd"]) / 1e9 - 1.717) < 1e-3
    assert b["surfel_forward"] == P * (49 + 12 * M) + 64 * P
    assert b["surfel_backward"] == 124 * P + P * (88 + 24 * M)
    # C4 (4 M surfels): 6.32 GB
    b4 = bench.algorithmic_bytes(4 * P, 4 * P, 14 * P, N_px, M)
    assert abs((b4["A_fwd"] + b4["A_bwd"]) / 1e9 - 6.32) < 0.01


def test_reference_arm_cpu_fallback_prints_contract_line():
    """Without a CUDA device the reference arm times the CPU oracle port on a bounded sample and still prints ONE JSON
    line with the contract's keys (on the GPU box it times the compiled reference instead)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_workload_shapes():
    scene, cams, grads, deg = bench.make_workload("C1")
    P = scene["xyz"].shape[0]
    assert P == 10_000 and deg == 3 and scene["shs"].shape == (P, 16, 3) and len(cams) == 4 == len(grads)
    assert cams[0].width == 256 and grads[0]["color"].shape == (3, 256, 256)


def test_stdout_guard_moves_only_what_is_written_inside_it():
    """bench._stdout_to_stderr (wrapped around NCCL's communicator creation, whose version banner goes to stdout): bytes
    written to file descriptor 1 inside the block land on stderr, everything before and after stays on stdout."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "print('before', flush=True)\n"
            "with bench._stdout_to_stderr():\n"
            "    os.write(1, b'NCCL version banner\\n')\n"
            "bench.emit({'metric': 'x'})\n"
            "os.write(1, b'after\\n')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    assert r.stdout.splitlines() == ["before", '{"metric": "x"}', "after"]
    assert "NCCL version banner" in r.stderr
