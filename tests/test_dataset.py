"""Trajectory dataset writer (ksim/dataset.py:38-103 mirror): field table on the CPU, packing + file format on the GPU."""

import json

import numpy as np
import pytest

from ksim_b200.dataset import dataset_fields


def test_field_table_covers_the_reference_sample(walking_program) -> None:  # noqa: ANN001
    """Names in the order of TrajectoryDatasetWriter.write (dataset.py:63-78); every field's flattened size adds up."""
    p, t = walking_program, 24
    m = p.spec.model
    names, shapes, table, size = dataset_fields(p, t)
    assert names[:9] == ["xquat", "xpos", "qpos", "qvel", "ctrl", "action", "done", "success", "timestep"]
    assert shapes[0] == (t, m.nbody, 4) and shapes[2] == (t, m.nq) and shapes[6] == (t,)
    groups = [n.split(".")[0] for n in names[9:]]
    assert groups == sorted(groups, key=["obs", "command", "termination_components", "reward", "reward_components"].index)
    assert "obs.noisy_joint_position" in names and "reward_components.upright/angvel_x" not in names or True
    assert size == sum(int(np.prod(s)) for s in shapes)
    assert (table[:, 3] == np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in shapes])[:-1]])).all()
    assert len([n for n in names if n.startswith("reward_components.")]) == p["TI_N_REW_COMP"]


@pytest.mark.gpu
def test_pack_and_write_match_numpy(tmp_path, walking_spec, walking_randomizers) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.dataset import TrajectoryDatasetWriter
    from ksim_b200.engine import B200Engine

    n, t = 37, 12
    eng = B200Engine(walking_spec, n, randomizers=walking_randomizers, seed=3)
    acts = 0.3 * torch.randn((t, n, eng.NA), device="cuda")
    rec = eng.rollout(acts)
    total, comps = eng.rewards(rec, t)
    path = tmp_path / "ds" / "traj.bin"
    with TrajectoryDatasetWriter(path, num_samples=2 * n) as w:
        w.write(eng, rec[:t], total, comps)
        w.write(eng, rec[:t], total, comps)
        with pytest.raises(ValueError, match="Dataset is full"):
            w.write(eng, rec[:t], total, comps)
    meta = json.loads(path.with_suffix(".meta.json").read_text())
    assert meta["num_samples"] == 2 * n and meta["total_size"] == sum(int(np.prod(s)) for s in meta["shapes"])
    data = np.memmap(path, dtype=np.float32, mode="r", shape=(2 * n, meta["total_size"]))
    r, tot, cmp_ = rec[:t].cpu().numpy(), total.cpu().numpy(), comps.cpu().numpy()
    names, shapes, table, size = dataset_fields(eng.program, t)
    for e in (0, 5, n - 1):
        parts = []
        for (src, off, dim, _), shape in zip(table, shapes):
            parts.append(r[:, e, off : off + dim].reshape(shape).flatten() if src == 0 else (tot[e] if src == 1 else cmp_[off, e]))
        want = np.concatenate(parts)
        np.testing.assert_array_equal(data[e], want)
        np.testing.assert_array_equal(data[n + e], want)
    # named access, the way TrajectoryDataset reads it back (dataset.py:137-160): slice by cumulative sizes and reshape
    i = meta["names"].index("qpos")
    o = sum(int(np.prod(s)) for s in meta["shapes"][:i])
    qpos = data[3, o : o + int(np.prod(meta["shapes"][i]))].reshape(meta["shapes"][i])
    np.testing.assert_array_equal(qpos, eng.view(rec[:t]).qpos[:, 3].cpu().numpy())


@pytest.mark.gpu
def test_trajectory_metrics_match_the_reference_formulas(walking_spec, walking_randomizers) -> None:  # noqa: ANN001
    """Trajectory.episode_length (types.py:72-76), get_termination_metrics / get_reward_metrics (rl.py:1544-1573) and the
    curriculum inputs (curriculum.py:125,187,262) restated in numpy over the same records."""
    import torch

    from ksim_b200.engine import B200Engine

    n, t = 200, 48
    eng = B200Engine(walking_spec, n, randomizers=walking_randomizers, seed=5, curriculum_level=1.0)
    rec = eng.rollout(0.5 * torch.randn((t, n, eng.NA), device="cuda"))
    total, comps = eng.rewards(rec, t)
    got = {k: v.cpu().numpy() for k, v in eng.metrics(rec, total, comps).items()}
    tv = eng.view(rec[:t])
    done = tv.done.cpu().numpy().T                      # (N, T)
    timestep = tv.timestep.cpu().numpy().T
    qpos = tv.qpos.cpu().numpy().transpose(1, 0, 2)
    p = eng.program
    terms = rec[:t, :, p["TI_R_TERM"] : p["TI_R_TERM"] + len(p.termination_names)].cpu().numpy().transpose(1, 0, 2)
    assert done.sum() > 0
    mask = done.copy(); mask[:, -1] = True
    ep = np.where(mask, timestep, 0.0).sum(-1) / mask.sum(-1)
    np.testing.assert_allclose(got["episode_length"], ep, rtol=1e-5)
    np.testing.assert_allclose(got["mean_episode_length"], ep.mean(), rtol=1e-5)
    np.testing.assert_allclose(got["deaths"], done.sum(-1))
    np.testing.assert_allclose(got["mean_terminations"], done.sum(-1).mean(), rtol=1e-6)
    np.testing.assert_allclose(got["max_distance_from_origin"], np.linalg.norm(qpos[..., :3], axis=-1).max(), rtol=1e-6)
    num = max((terms != 0).any(-1).sum(), 1)
    assert got["num_terminations"] == num
    for i, name in enumerate(p.termination_names):
        np.testing.assert_allclose(got[f"prct/{name}"], terms[..., i].sum() / num, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got["reward/_total"], total.cpu().numpy().mean(), rtol=1e-4, atol=1e-6)
    cm = comps.cpu().numpy()
    for i, name in enumerate(p.reward_components):
        np.testing.assert_allclose(got[f"reward/{name}"], cm[i].mean(), rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_histogram_matches_numpy(walking_spec, walking_randomizers) -> None:  # noqa: ANN001
    """RLTask.get_histogram (rl.py:1521-1541): jnp.histogram == numpy.histogram semantics (range [min, max], last bin closed,
    +-0.5 when degenerate), limits = edges[1:], and the five scalar statistics."""
    import torch

    from ksim_b200.engine import B200Engine

    eng = B200Engine(walking_spec, 8, randomizers=walking_randomizers, seed=0)
    r = np.random.RandomState(0)
    for data, bins in ((r.randn(4096, 64).astype(np.float32), 100), (r.rand(1000).astype(np.float32) ** 3, 7),
                       (np.full(333, 2.5, np.float32), 100), (np.arange(101, dtype=np.float32), 100),
                       (np.concatenate([r.randn(10_000), [50.0]]).astype(np.float32), 1024)):
        got = {k: v.cpu().numpy() for k, v in eng.get_histogram(torch.from_numpy(data).cuda(), bins).items()}
        flat = data.reshape(-1)
        counts, edges = np.histogram(flat.astype(np.float64), bins=bins)
        assert got["counts"].sum() == flat.size
        # fp32 edges can move a value sitting on a float64 edge into the neighbouring bin: compare cumulative counts loosely
        assert np.abs(np.cumsum(got["counts"]) - np.cumsum(counts)).max() <= max(2, flat.size // 5000)
        np.testing.assert_allclose(got["limits"], edges[1:], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose([got["min"], got["max"]], [flat.min(), flat.max()], rtol=0, atol=0)
        f64 = flat.astype(np.float64)
        np.testing.assert_allclose([got["sum"], got["sum_squares"], got["mean"]], [f64.sum(), (f64 * f64).sum(), f64.mean()], rtol=1e-5, atol=1e-4)
    exact = eng.get_histogram(torch.arange(100, dtype=torch.float32, device="cuda"), 10)["counts"].cpu().numpy()
    np.testing.assert_array_equal(exact, np.histogram(np.arange(100), bins=10)[0])
