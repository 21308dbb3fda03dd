"""Trajectory dataset writer.  Mirrors ``ksim/dataset.py:38-103`` (``TrajectoryDatasetWriter``): a flat fp32 memmap with one
flattened trajectory per row plus a ``.meta.json`` holding the field names and shapes -- the on-disk format of
``collect_dataset`` (ksim/task/rl.py:2315-2411).

The reference writes one (T,)-shaped trajectory per ``write`` call, concatenating ``xquat, xpos, qpos, qvel, ctrl, action,
done, success, timestep, obs.*, command.*, termination_components.*, reward, reward_components.*`` after ``flatten()``.
Here the trajectories of ALL environments of a rollout are packed on the GPU in one launch (``b200sim_dataset_pack``:
record buffer ``rec[T][N][R]`` + reward rows -> ``out[N][total_size]``) and land in the memmap as N consecutive rows.
Dict-valued entries are emitted in sorted key order (the order JAX gives dict pytrees); the meta file names every field, so
a reader never depends on it.
"""

from __future__ import annotations

__all__ = ["TrajectoryDatasetWriter", "dataset_fields"]

import json
from pathlib import Path
from types import TracebackType

import numpy as np
import torch

from ksim_b200 import binding
from ksim_b200.program import TaskProgram


def dataset_fields(program: TaskProgram, n_steps: int) -> tuple[list[str], list[tuple[int, ...]], np.ndarray, int]:
    """names, per-field shapes (leading T), the device field table [n][4] = (source, offset, dim, out_offset), total size."""
    m = program.spec.model
    names: list[str] = []
    shapes: list[tuple[int, ...]] = []
    rows: list[tuple[int, int, int]] = []

    def rec(name: str, slot: str, shape: tuple[int, ...], extra: int = 0) -> None:
        names.append(name)
        shapes.append((n_steps, *shape))
        rows.append((0, program[slot] + extra, int(np.prod(shape)) if shape else 1))

    rec("xquat", "TI_R_XQUAT", (m.nbody, 4))
    rec("xpos", "TI_R_XPOS", (m.nbody, 3))
    rec("qpos", "TI_R_QPOS", (m.nq,))
    rec("qvel", "TI_R_QVEL", (m.nv,))
    rec("ctrl", "TI_R_CTRL", (m.nu,))
    rec("action", "TI_R_ACTION", (program.action_size,))
    rec("done", "TI_R_DONE", ())
    rec("success", "TI_R_SUCCESS", ())
    rec("timestep", "TI_R_TIME", ())
    for name in sorted(program.obs_slices):
        off, dim, shape = program.obs_slices[name]
        assert int(np.prod(shape)) == dim or shape == ()
        rec(f"obs.{name}", "TI_R_OBS", tuple(shape), off)
    for name in sorted(program.cmd_slices):
        off, dim = program.cmd_slices[name]
        rec(f"command.{name}", "TI_R_CMD", (dim,), off)
    for name in sorted(program.termination_names):
        rec(f"termination_components.{name}", "TI_R_TERM", (), program.termination_names.index(name))
    names.append("reward"); shapes.append((n_steps,)); rows.append((1, 0, 1))
    for name in sorted(program.reward_components):
        names.append(f"reward_components.{name}"); shapes.append((n_steps,))
        rows.append((2, program.reward_components.index(name), 1))
    table = np.zeros((len(rows), 4), dtype=np.int32)
    out = 0
    for i, (src, off, dim) in enumerate(rows):
        table[i] = (src, off, dim, out)
        out += n_steps * dim
    return names, shapes, table, out


class TrajectoryDatasetWriter:
    """Same constructor, context-manager protocol, file layout and ``Dataset is full`` error as the reference's writer."""

    def __init__(self, path: str | Path, num_samples: int) -> None:
        self.path = Path(path)
        self.num_samples = num_samples
        self.path.parent.mkdir(parents=True, exist_ok=True)
        self.fp: np.memmap | None = None
        self.index = 0
        self._table: torch.Tensor | None = None
        self._total = 0

    def __enter__(self) -> "TrajectoryDatasetWriter":
        return self

    def __exit__(self, exc_type: type[BaseException] | None, exc_value: BaseException | None, traceback: TracebackType | None) -> None:
        if self.fp is not None:
            self.fp.flush()
            self.fp = None

    def pack(self, engine, rec: torch.Tensor, total: torch.Tensor, comps: torch.Tensor) -> torch.Tensor:  # noqa: ANN001
        """rec[T][N][R] (the T records of a rollout), total[N][T], comps[K][N][T] -> samples[N][total_size] on the GPU."""
        t, n = int(rec.shape[0]), int(rec.shape[1])
        names, shapes, table, size = dataset_fields(engine.program, t)
        if self._table is None or self._total != size:
            self._table = torch.from_numpy(table).to(rec.device)
            self._total, self._names, self._shapes = size, names, shapes
        out = torch.empty((n, size), dtype=torch.float32, device=rec.device)
        ptr = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        with torch.cuda.device(rec.device):
            stream = torch.cuda.current_stream(rec.device).cuda_stream
            binding.check(binding.lib().b200sim_dataset_pack(stream, n, t, int(rec.shape[2]), ptr(rec), ptr(total), ptr(comps),
                                                             ptr(self._table), int(table.shape[0]), size, ptr(out)),
                          "b200sim_dataset_pack")
        engine.launches += 1
        return out

    def write(self, engine, rec: torch.Tensor, total: torch.Tensor, comps: torch.Tensor) -> None:  # noqa: ANN001
        """One row per environment (the reference's ``write`` called once per environment's trajectory)."""
        n = int(rec.shape[1])
        if self.index + n > self.num_samples:
            raise ValueError("Dataset is full")
        samples = self.pack(engine, rec, total, comps)
        if self.fp is None:
            meta = {"names": self._names, "shapes": [list(s) for s in self._shapes], "num_samples": self.num_samples,
                    "total_size": self._total}
            with open(self.path.with_suffix(".meta.json"), "w") as f:
                json.dump(meta, f, indent=2)
            self.fp = np.memmap(self.path, dtype=np.float32, mode="w+", shape=(self.num_samples, self._total))
            self.index = 0
        self.fp[self.index : self.index + n] = samples.cpu().numpy()
        self.index += n
