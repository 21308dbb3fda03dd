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


def test_rank0_only_wait_loop_has_no_collective() -> None:
    """The extra warm-up steps bench.py runs while it waits for nvidia-smi's first sample execute on rank 0 ONLY: a collective
    inside them leaves the other ranks with fewer enqueued all-reduces and deadlocks the run (it did, at 8 GPUs).  Static check
    of the source: the loop body calls the collective-free form of the step."""
    import ast
    from pathlib import Path

    src = (Path(__file__).resolve().parent.parent / "bench.py").read_text()
    tree = ast.parse(src)
    loops = [n for n in ast.walk(tree) if isinstance(n, ast.While) and "sampler.samples" in ast.unparse(n.test)]
    assert len(loops) == 1
    calls = [c for c in ast.walk(loops[0]) if isinstance(c, ast.Call) and getattr(c.func, "id", "") == "hot_path"]
    assert calls, "the wait loop should still step the hot path"
    for c in calls:
        kw = {k.arg: ast.unparse(k.value) for k in c.keywords}
        assert kw.get("with_collective") == "False", ast.unparse(c)
    assert "collective()" not in ast.unparse(loops[0])
