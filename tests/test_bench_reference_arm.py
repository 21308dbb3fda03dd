"""`bench.py --impl reference` (the CPU arm the driver times beside the CUDA arm) on the host: the JSON line carries the
keys the contract names, and the arm uses every host core even when the launcher exported ``OMP_NUM_THREADS=1`` (which
`torch.distributed.run` does for N > 1 workers).  Ranks other than 0 exit 0 without printing."""

from __future__ import annotations

import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(env_extra: dict[str, str]) -> subprocess.CompletedProcess:
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_line_and_thread_count() -> None:
    out = _run({"OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "humanoid env-steps/sec" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    # `cores` is omp_get_max_threads() of the oracle library after the arm set it, not a copy of os.cpu_count()
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly() -> None:
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""
