"""B200 rollout engine: the host-side mirror of ksim's ``PhysicsEngine`` for the batched CUDA path.

Reference boundary (ksim/engine.py:171-199): ``PhysicsEngine`` owns ``actuators``, ``resets``, ``events`` and the
engine config, and exposes ``reset(physics_model, curriculum_level, rng) -> PhysicsState`` and
``step(action, physics_model, physics_state, curriculum_level, rng) -> PhysicsState``; the rollout driver
(``RLTask._single_unroll``, ksim/task/rl.py:1604-1665) vmaps it over environments and scans it over time,
wrapping it with observations, terminations, reset-on-done, commands and finally rewards + GAE.

``B200Engine`` keeps those names and meanings but is *batched*: state for all N environments of this rank
lives in HBM as flat per-env blocks, and one kernel launch advances every environment by a control step (or a
whole rollout).  PyTorch is used only for device memory and streams.  No CPU fallback exists: construction
fails loudly when the CUDA library is missing (ksim_b200/binding.py).
"""

from __future__ import annotations

__all__ = ["B200Engine", "TrajectoryView", "PhysicsData", "ContactData", "get_physics_engine"]

import ctypes
from dataclasses import dataclass
from typing import Mapping

import numpy as np
import torch

from ksim_b200 import binding
from ksim_b200 import codes as C
from ksim_b200.blob import pack_blob
from ksim_b200.envinit import EnvInit, make_env_init
from ksim_b200.program import TaskProgram, TaskSpec, build_program
from ksim_b200.randomization import PhysicsRandomizer


def _ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(None if t is None else t.data_ptr())


@dataclass
class TrajectoryView:
    """Named views into a record buffer ``rec[T][N][R]`` -- the fields of ksim's ``Trajectory`` (types.py:48-76).

    ``obs`` / ``command`` / ``curriculum_level`` are pre-step, everything else post-step (types.py:51-54).
    """

    rec: torch.Tensor
    program: TaskProgram

    def _blk(self, slot: str, n: int) -> torch.Tensor:
        off = self.program[slot]
        return self.rec[..., off : off + n]

    @property
    def qpos(self) -> torch.Tensor:
        return self._blk("TI_R_QPOS", self.program.spec.model.nq)

    @property
    def qvel(self) -> torch.Tensor:
        return self._blk("TI_R_QVEL", self.program.spec.model.nv)

    @property
    def xpos(self) -> torch.Tensor:
        nb = self.program.spec.model.nbody
        return self._blk("TI_R_XPOS", 3 * nb).unflatten(-1, (nb, 3))

    @property
    def xquat(self) -> torch.Tensor:
        nb = self.program.spec.model.nbody
        return self._blk("TI_R_XQUAT", 4 * nb).unflatten(-1, (nb, 4))

    @property
    def ctrl(self) -> torch.Tensor:
        return self._blk("TI_R_CTRL", self.program.spec.model.nu)

    @property
    def action(self) -> torch.Tensor:
        return self._blk("TI_R_ACTION", self.program.action_size)

    @property
    def done(self) -> torch.Tensor:
        return self._blk("TI_R_DONE", 1)[..., 0] != 0

    @property
    def success(self) -> torch.Tensor:
        return self._blk("TI_R_SUCCESS", 1)[..., 0] != 0

    @property
    def timestep(self) -> torch.Tensor:
        return self._blk("TI_R_TIME", 1)[..., 0]

    @property
    def curriculum_level(self) -> torch.Tensor:
        return self._blk("TI_R_LEVEL", 1)[..., 0]

    @property
    def obs_flat(self) -> torch.Tensor:
        return self._blk("TI_R_OBS", self.program.obs_size)

    @property
    def obs(self) -> dict[str, torch.Tensor]:
        flat = self.obs_flat
        return {k: flat[..., off : off + dim].unflatten(-1, shape) for k, (off, dim, shape) in self.program.obs_slices.items()}

    @property
    def command(self) -> dict[str, torch.Tensor]:
        flat = self._blk("TI_R_CMD", self.program["TI_CMD_SIZE"])
        return {k: flat[..., off : off + dim] for k, (off, dim) in self.program.cmd_slices.items()}

    @property
    def event_state(self) -> dict[str, torch.Tensor]:
        flat = self._blk("TI_R_EVENT", self.program["TI_EVENT_SIZE"])
        return {k: flat[..., off : off + dim] for k, (off, dim) in self.program.event_slices.items()}

    @property
    def aux_outputs(self) -> dict[str, torch.Tensor]:
        """``Trajectory.aux_outputs``: writable views ``[..., dim]`` of the record's aux region, one per declared output."""
        off = self.program["TI_R_AUX"]
        return {k: self.rec[..., off + o : off + o + d] for k, (o, d) in self.program.aux_slices.items()}

    @property
    def termination_components(self) -> dict[str, torch.Tensor]:
        names = self.program.termination_names
        flat = self._blk("TI_R_TERM", len(names))
        return {k: flat[..., i].to(torch.int32) for i, k in enumerate(names)}


@dataclass
class ContactData:
    """``physics_state.data.contact`` as the plugins read it (utils/mujoco.py:175-217, terminations.py:84-99): fixed candidate
    slots, a contact is active when ``dist < 0``."""

    geom1: torch.Tensor   # [N, ncon] int32
    geom2: torch.Tensor
    dist: torch.Tensor    # [N, ncon]
    pos: torch.Tensor     # [N, ncon, 3]

    @property
    def geom(self) -> torch.Tensor:
        return torch.stack([self.geom1, self.geom2], dim=-1)


@dataclass
class PhysicsData:
    """Named views into the data blocks ``b200sim_engine_step`` / ``_reset`` write -- the fields of ``physics_state.data`` the
    reference's Observation / Reward / Termination / Command / Event plugins read, batched over environments."""

    block: torch.Tensor           # [N, DATA] fp32
    offsets: tuple[int, ...]      # B2S_DATA_N_FIELDS + 1 float offsets
    model: object

    def _f(self, field: int, *shape: int) -> torch.Tensor:
        v = self.block[:, self.offsets[field] : self.offsets[field + 1]]
        return v.unflatten(-1, shape) if shape else v

    @property
    def qpos(self) -> torch.Tensor: return self._f(C.B2S_DATA_QPOS)                                    # noqa: E704
    @property
    def qvel(self) -> torch.Tensor: return self._f(C.B2S_DATA_QVEL)                                    # noqa: E704
    @property
    def qacc(self) -> torch.Tensor: return self._f(C.B2S_DATA_QACC)                                    # noqa: E704
    @property
    def ctrl(self) -> torch.Tensor: return self._f(C.B2S_DATA_CTRL)                                    # noqa: E704
    @property
    def time(self) -> torch.Tensor: return self._f(C.B2S_DATA_TIME)[:, 0]                              # noqa: E704
    @property
    def xpos(self) -> torch.Tensor: return self._f(C.B2S_DATA_XPOS, self.model.nbody, 3)               # noqa: E704
    @property
    def xquat(self) -> torch.Tensor: return self._f(C.B2S_DATA_XQUAT, self.model.nbody, 4)             # noqa: E704
    @property
    def cinert(self) -> torch.Tensor: return self._f(C.B2S_DATA_CINERT, self.model.nbody, 10)          # noqa: E704
    @property
    def cvel(self) -> torch.Tensor: return self._f(C.B2S_DATA_CVEL, self.model.nbody, 6)               # noqa: E704
    @property
    def actuator_force(self) -> torch.Tensor: return self._f(C.B2S_DATA_ACTUATOR_FORCE)                # noqa: E704
    @property
    def sensordata(self) -> torch.Tensor: return self._f(C.B2S_DATA_SENSORDATA)                        # noqa: E704
    @property
    def cfrc_ext(self) -> torch.Tensor: return self._f(C.B2S_DATA_CFRC_EXT, self.model.nbody, 6)       # noqa: E704
    @property
    def site_xpos(self) -> torch.Tensor: return self._f(C.B2S_DATA_SITE_XPOS, self.model.nsite, 3)     # noqa: E704
    @property
    def site_xmat(self) -> torch.Tensor: return self._f(C.B2S_DATA_SITE_XMAT, self.model.nsite, 3, 3)  # noqa: E704
    @property
    def subtree_com(self) -> torch.Tensor: return self._f(C.B2S_DATA_SUBTREE_COM, self.model.nbody, 3)  # noqa: E704
    @property
    def xfrc_applied(self) -> torch.Tensor: return self._f(C.B2S_DATA_XFRC_APPLIED, self.model.nbody, 6)  # noqa: E704
    @property
    def ncon(self) -> torch.Tensor: return self._f(C.B2S_DATA_NCON)[:, 0].to(torch.int32)              # noqa: E704

    @property
    def contact(self) -> ContactData:
        nc = self.offsets[C.B2S_DATA_CONTACT_GEOM2] - self.offsets[C.B2S_DATA_CONTACT_GEOM1]
        return ContactData(self._f(C.B2S_DATA_CONTACT_GEOM1).to(torch.int32), self._f(C.B2S_DATA_CONTACT_GEOM2).to(torch.int32),
                           self._f(C.B2S_DATA_CONTACT_DIST), self._f(C.B2S_DATA_CONTACT_POS, nc, 3))


class B200Engine:
    """Batched drop-in for ``MjxEngine`` + the rollout functions around it (see module docstring)."""

    def __init__(self, spec: TaskSpec, num_envs: int, device: torch.device | str = "cuda:0",
                 randomizers: Mapping[str, PhysicsRandomizer] | None = None, seed: int = 0, curriculum_level: float = 0.0,
                 env_offset: int = 0, total_envs: int | None = None, program: TaskProgram | None = None) -> None:
        self.lib = binding.lib()  # raises if the CUDA library is not built
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise binding.B200SimError("B200Engine runs on CUDA devices only; there is no CPU fallback")
        # mirror of PhysicsEngine's fields (ksim/engine.py:171-178)
        self.actuators, self.resets, self.events, self.data = spec.actuators, tuple(spec.resets), dict(spec.events), spec.config
        self.spec = spec
        self.program = build_program(spec) if program is None else program
        self.num_envs = int(num_envs)
        self.randomizers = dict(randomizers or {})
        self.blob = pack_blob(spec.model, self.program.ti, self.program.tf)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            buf = (ctypes.c_char * len(self.blob)).from_buffer_copy(self.blob)
            binding.check(self.lib.b200sim_model_create(buf, len(self.blob), ctypes.byref(self._handle)), "b200sim_model_create")
        binding.check(self.lib.b200sim_model_set_option(self._handle, C.B2S_OPT_SOLVER,
                                                        C.B2S_SOLVER_REFERENCE if spec.config.solver == "reference" else C.B2S_SOLVER_FAST),
                      "b200sim_model_set_option")
        p = self.program
        self.S, self.R, self.K, self.RAND, self.NA = p.state_size, p.rec_size, p.key_size, p.rand_size, p.action_size
        for name, want in (("STATE_SIZE", self.S), ("REC_SIZE", self.R), ("KEY_SIZE", self.K), ("RAND_SIZE", self.RAND),
                           ("ACTION_SIZE", self.NA)):
            got = self.info(name)
            if got != want:
                raise binding.B200SimError(f"layout mismatch for {name}: library {got}, host {want}")
        n, dev = self.num_envs, self.device
        self.state = torch.zeros((n, self.S), dtype=torch.float32, device=dev)
        self.keys = torch.zeros((n, self.K), dtype=torch.int32, device=dev)  # uint32 bit patterns
        self.rand = torch.zeros((n, max(self.RAND, 1)), dtype=torch.float32, device=dev)
        self.levels = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.reward_carry = torch.zeros((n, max(p["TI_REW_CARRY_SIZE"], 1)), dtype=torch.float32, device=dev)
        self.first_rec = torch.zeros((n, self.R), dtype=torch.float32, device=dev)
        self.act_keys = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        self.env_init: EnvInit | None = None
        self.launches = 0
        self.reset(seed=seed, curriculum_level=curriculum_level, env_offset=env_offset, total_envs=total_envs)

    def __del__(self) -> None:
        try:
            if self._handle:
                self.lib.b200sim_model_destroy(self._handle)
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------------------------------------ helpers
    def info(self, name: str) -> int:
        return int(self.lib.b200sim_model_info(self._handle, binding.INFO[name]))

    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t: torch.Tensor, shape: tuple[int, ...], dtype: torch.dtype, name: str) -> None:
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != shape or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous {dtype} {shape} on {self.device}, got {t.dtype} {tuple(t.shape)} on {t.device}")

    def new_record_buffer(self, n_steps: int) -> torch.Tensor:
        """rec[T+1][N][R]; slot 0 receives the current pre-step block (observations of the next policy call).  Slot T of a
        rollout holds only the NEXT step's pre-step block -- scoring passes work on the first T slots."""
        rec = torch.zeros((n_steps + 1, self.num_envs, self.R), dtype=torch.float32, device=self.device)
        rec[0].copy_(self.first_rec)
        return rec

    @staticmethod
    def _scored_steps(rec: torch.Tensor, n_steps: int | None) -> int:
        """Record buffers are ``[T+1][N][R]`` (``rollout`` / ``new_record_buffer``): the reference's ``get_rewards`` works on
        exactly the T transitions (rl.py:189-245), so the default is ``rec.shape[0] - 1``; slot T is never a transition."""
        t = int(rec.shape[0]) - 1 if n_steps is None else int(n_steps)
        if not 0 <= t <= int(rec.shape[0]):
            raise ValueError(f"n_steps {t} outside a record buffer of {int(rec.shape[0])} slots")
        return t

    def view(self, rec: torch.Tensor) -> TrajectoryView:
        return TrajectoryView(rec, self.program)

    # ----------------------------------------------------------------------- the engine alone (PhysicsEngine)
    def data_layout(self) -> tuple[int, ...]:
        offs = (ctypes.c_int32 * (C.B2S_DATA_N_FIELDS + 1))()
        binding.check(self.lib.b200sim_data_layout(self._handle, offs), "b200sim_data_layout")
        return tuple(int(x) for x in offs)

    def new_data(self) -> PhysicsData:
        offs = self.data_layout()
        return PhysicsData(torch.zeros((self.num_envs, offs[-1]), dtype=torch.float32, device=self.device), offs, self.spec.model)

    def _api_check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.b200sim_engine_last_error()
            raise binding.B200SimError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")

    def engine_reset(self, rng: torch.Tensor, curriculum_level: torch.Tensor | float = 0.0, data: PhysicsData | None = None) -> PhysicsData:
        """``PhysicsEngine.reset(physics_model, curriculum_level, rng)`` (ksim/engine.py:205-246) for every environment, without the
        task layer: ``rng [N, 2]`` int32 (uint32 bit patterns) is the key each environment's reset consumes.  The physics state
        lands in ``self.state`` (the ``PhysicsState``), its ``data`` fields in the returned ``PhysicsData``."""
        n = self.num_envs
        self._chk(rng, (n, 2), torch.int32, "rng")
        lv = curriculum_level if torch.is_tensor(curriculum_level) else torch.full((n,), float(curriculum_level), device=self.device)
        self._chk(lv, (n,), torch.float32, "curriculum_level")
        data = self.new_data() if data is None else data
        with torch.cuda.device(self.device):
            self._api_check(self.lib.b200sim_engine_reset(self._handle, self._stream(), n, _ptr(rng), _ptr(lv), _ptr(self.rand), _ptr(self.state),
                                                          _ptr(self.keys), _ptr(data.block)), "b200sim_engine_reset")
        self.launches += 1
        self.levels.copy_(lv)
        return data

    def engine_step(self, action: torch.Tensor, rng: torch.Tensor, data: PhysicsData | None = None) -> PhysicsData:
        """``PhysicsEngine.step(action, physics_model, physics_state, curriculum_level, rng)`` (ksim/engine.py:259-361) for every
        environment: actuators + events + the physics sub-steps of one control step, nothing else (no observation, termination,
        reset-on-done or command update -- ``RLTask._step_physics``, rl.py:1143-1157, is exactly this call)."""
        n = self.num_envs
        self._chk(action, (n, self.NA), torch.float32, "action")
        self._chk(rng, (n, 2), torch.int32, "rng")
        data = self.new_data() if data is None else data
        with torch.cuda.device(self.device):
            self._api_check(self.lib.b200sim_engine_step(self._handle, self._stream(), n, _ptr(action), _ptr(rng), _ptr(self.rand),
                                                         _ptr(self.state), _ptr(self.keys), _ptr(data.block)), "b200sim_engine_step")
        self.launches += 1
        return data

    def get_terminations(self, data: PhysicsData, curriculum_level: torch.Tensor | None = None,
                         ) -> tuple[dict[str, torch.Tensor], torch.Tensor, torch.Tensor]:
        """``get_terminations`` (rl.py:288-299) of the task's Termination plugins on a ``PhysicsData``: ``({name: code [N] int32},
        done [N] bool, success [N] bool)`` with ``done = any(code != 0)``, ``success = all(code != -1) & done`` (rl.py:1049-1050)."""
        n = int(data.block.shape[0])
        names = self.program.termination_names
        codes = torch.zeros((n, max(len(names), 1)), dtype=torch.int32, device=self.device)
        done = torch.zeros((n,), dtype=torch.uint8, device=self.device)
        succ = torch.zeros((n,), dtype=torch.uint8, device=self.device)
        lv = self.levels[:n].contiguous() if curriculum_level is None else curriculum_level
        self._chk(lv, (n,), torch.float32, "curriculum_level")
        with torch.cuda.device(self.device):
            self._api_check(self.lib.b200sim_terminations(self._handle, self._stream(), n, _ptr(data.block), _ptr(lv), _ptr(codes), _ptr(done),
                                                          _ptr(succ)), "b200sim_terminations")
        self.launches += 1
        return {k: codes[:, i] for i, k in enumerate(names)}, done != 0, succ != 0

    # ------------------------------------------------------------------------------------------------- reset
    def reset(self, seed: int = 0, curriculum_level: float = 0.0, env_offset: int = 0, total_envs: int | None = None) -> None:
        """(Re)initialises every environment: the batched ``_get_env_state`` + ``engine.reset`` (rl.py:2175-2272)."""
        init = make_env_init(self.program, self.randomizers, self.num_envs, seed=seed, level=curriculum_level,
                             env_offset=env_offset, total_envs=total_envs)
        self.env_init = init
        dev = self.device
        if self.RAND > 0:
            self.rand.copy_(torch.from_numpy(init.rand).to(dev))
        self.levels.copy_(torch.from_numpy(init.levels).to(dev))
        ik = torch.from_numpy(init.init_keys.view(np.int32)).to(dev).contiguous()
        self.reward_carry.zero_()
        with torch.cuda.device(dev):
            binding.check(self.lib.b200sim_init_envs(self._handle, self._stream(), self.num_envs, _ptr(ik), _ptr(self.levels),
                                                     _ptr(self.rand), _ptr(self.state), _ptr(self.keys), _ptr(self.first_rec),
                                                     _ptr(self.act_keys)), "b200sim_init_envs")
        self.launches += 1
        torch.cuda.current_stream(dev).synchronize()  # ik must outlive the launch

    # -------------------------------------------------------------------------------------------------- step
    def step(self, action: torch.Tensor, rec_cur: torch.Tensor, rec_next: torch.Tensor) -> None:
        """One control step for every env (``RLTask.step_engine``, rl.py:1164-1211)."""
        n = self.num_envs
        self._chk(action, (n, self.NA), torch.float32, "action")
        self._chk(rec_cur, (n, self.R), torch.float32, "rec_cur")
        self._chk(rec_next, (n, self.R), torch.float32, "rec_next")
        with torch.cuda.device(self.device):
            binding.check(self.lib.b200sim_step_ctrl(self._handle, self._stream(), n, _ptr(action), _ptr(self.rand), _ptr(self.state),
                                                     _ptr(self.keys), _ptr(rec_cur), _ptr(rec_next), _ptr(self.act_keys)),
                          "b200sim_step_ctrl")
        self.launches += 1

    def rollout(self, actions: torch.Tensor, rec: torch.Tensor | None = None) -> torch.Tensor:
        """T control steps with given ``actions[T][N][NA]`` in one launch (``_single_unroll``'s scan, rl.py:1620).  ``rec`` may
        be a caller-owned ``[T+1][N][R]`` buffer that is reused between calls: its slot 0 is refreshed here from the
        engine-owned pre-step block (``first_rec``), which is in turn refreshed from slot T after the launch."""
        t, n = int(actions.shape[0]), self.num_envs
        self._chk(actions, (t, n, self.NA), torch.float32, "actions")
        if rec is None:
            rec = self.new_record_buffer(t)
        else:
            self._chk(rec, (t + 1, n, self.R), torch.float32, "rec")
            rec[0].copy_(self.first_rec)
        if t == 0 or n == 0:
            return rec
        with torch.cuda.device(self.device):
            binding.check(self.lib.b200sim_rollout(self._handle, self._stream(), n, t, _ptr(actions), _ptr(self.rand), _ptr(self.state),
                                                   _ptr(self.keys), _ptr(rec)), "b200sim_rollout")
        self.launches += 1
        self.first_rec.copy_(rec[t])  # pre-step block of the next rollout (engine-owned: set_curriculum_level patches it)
        return rec

    # ----------------------------------------------------------------------------------------------- scoring
    def rewards(self, rec: torch.Tensor, n_steps: int | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """``get_rewards`` for every env (rl.py:189-245): returns total[N][T], components[K][N][T]; updates the carry."""
        t = self._scored_steps(rec, n_steps)
        n, k = self.num_envs, self.program["TI_N_REW_COMP"]
        total = torch.empty((n, t), dtype=torch.float32, device=self.device)
        comps = torch.empty((max(k, 1), n, t), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            binding.check(self.lib.b200sim_rewards(self._handle, self._stream(), n, t, _ptr(rec), _ptr(self.levels),
                                                   _ptr(self.reward_carry), _ptr(total), _ptr(comps)), "b200sim_rewards")
        self.launches += 1
        return total, comps

    def postprocess_and_rewards(self, rec: torch.Tensor, n_steps: int | None = None,
                                postprocess_trajectory=None) -> tuple[torch.Tensor, torch.Tensor]:  # noqa: ANN001
        """What ``_single_unroll`` does after the scan (rl.py:1636-1660): ``postprocess_trajectory(trajectory)`` -- the hook AMP /
        teacher tasks override (amp.py:327-358, teacher.py:113-136) -- then ``get_rewards``.  The hook receives the
        ``TrajectoryView`` of the T transitions and returns ``{aux output name: tensor [T, N] or [T, N, dim]}``; the values are
        stored in the records' aux region (``TaskSpec.aux_outputs`` declares the slots) where the rewards pass reads them."""
        t = self._scored_steps(rec, n_steps)
        hook = postprocess_trajectory if postprocess_trajectory is not None else getattr(self, "postprocess_trajectory", None)
        if hook is not None:
            view = self.view(rec[:t])
            slots = view.aux_outputs
            for name, value in hook(view).items():
                if name not in slots:
                    raise KeyError(f"postprocess_trajectory returned '{name}', the task declares aux outputs {sorted(slots)}")
                dst = slots[name]
                dst.copy_(value.reshape(dst.shape).to(dst.dtype))
        return self.rewards(rec, t)

    def metrics(self, rec: torch.Tensor, total: torch.Tensor, comps: torch.Tensor, n_steps: int | None = None) -> dict[str, torch.Tensor]:
        """The per-iteration trajectory reductions the curricula and the logger consume: ``Trajectory.episode_length``
        (types.py:72-76), ``get_termination_metrics`` / ``get_reward_metrics`` (rl.py:1544-1573), the inputs of
        ``EpisodeLengthCurriculum`` / ``DistanceFromOriginCurriculum`` / ``StepWhenSaturated`` (curriculum.py:125,187,262).
        Returns 0-d / 1-d device tensors; ``episode_length`` and ``deaths`` are per environment."""
        t = self._scored_steps(rec, n_steps)
        n, k = self.num_envs, self.program["TI_N_REW_COMP"]
        width = 13 + k
        per_env = torch.empty((n, width), dtype=torch.float32, device=self.device)
        out = torch.empty((width,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            binding.check(self.lib.b200sim_metrics(self._handle, self._stream(), n, t, _ptr(rec), _ptr(total), _ptr(comps),
                                                   _ptr(per_env), _ptr(out)), "b200sim_metrics")
        self.launches += 2
        res = {"episode_length": per_env[:, 0], "deaths": per_env[:, 1], "mean_episode_length": out[0], "mean_terminations": out[1],
               "max_distance_from_origin": out[2], "num_terminations": out[3], "reward/_total": out[12]}
        for i, name in enumerate(self.program.termination_names):
            res[f"prct/{name}"] = out[4 + i]
        for i, name in enumerate(self.program.reward_components):
            res[f"reward/{name}"] = out[13 + i]
        return res

    def get_histogram(self, arr: torch.Tensor, bins: int = 100) -> dict[str, torch.Tensor]:
        """``RLTask.get_histogram`` (rl.py:1521-1541), the reduction ``_rl_train_loop_step`` applies to every logged metric with
        more than one element (rl.py:1815): ``counts`` (int32 ``[bins]``), ``limits`` (upper bin edges), ``min``, ``max``, ``sum``,
        ``sum_squares``, ``mean`` -- the fields of ``xax.Histogram``."""
        x = arr.detach().to(torch.float32).contiguous().reshape(-1)
        if x.device != self.device or x.numel() == 0:
            raise ValueError("get_histogram needs a non-empty tensor on the engine's device")
        counts = torch.empty((bins,), dtype=torch.int32, device=self.device)
        limits = torch.empty((bins,), dtype=torch.float32, device=self.device)
        stats = torch.empty((5,), dtype=torch.float32, device=self.device)
        scratch = torch.empty(int(self.lib.b200sim_histogram_scratch_bytes()), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.b200sim_histogram(self._stream(), x.numel(), _ptr(x), bins, _ptr(counts), _ptr(limits), _ptr(stats), _ptr(scratch))
        if rc != 0:
            msg = self.lib.b200sim_policy_last_error()
            raise binding.B200SimError(f"b200sim_histogram failed with code {rc}: {msg.decode() if msg else ''}")
        self.launches += 2
        return {"counts": counts, "limits": limits, "min": stats[0], "max": stats[1], "sum": stats[2], "sum_squares": stats[3], "mean": stats[4]}

    def set_curriculum_level(self, level: float) -> None:
        """The level the next rollout runs at.  The reference passes ``curriculum_state.level`` to every ``step_engine`` call
        of an iteration (rl.py:1615-1634); here it travels in each environment's state block, so the once-per-iteration update
        (``ksim_b200.curriculum``) is written into that column and into the per-env levels the rewards pass scales with."""
        self.state[:, self.program["TI_S_LEVEL"]].fill_(float(level))
        self.levels.fill_(float(level))
        # the pre-step block of the next rollout's first step was written at the end of the last one: its recorded level is
        # patched here; its observation noise was drawn at the old level (the fused step evaluates the observation of step
        # t + 1 at the end of step t), a one-step lag at the iteration boundary that the reference does not have
        self.first_rec[:, self.program["TI_R_LEVEL"]].fill_(float(level))

    def score(self, rec: torch.Tensor, n_steps: int | None = None, values: torch.Tensor | None = None, gamma: float = 0.0,
              lam: float = 0.0, normalize_advantages: bool = False, monte_carlo_returns: bool = False) -> dict[str, torch.Tensor]:
        """The scoring pass after a rollout as ONE launch (``b200sim_score``): ``get_rewards`` (rl.py:189-245), the done / success
        flags and -- with ``values`` [N][T] -- ``compute_ppo_inputs`` (ppo.py:54-121).  Same results as ``rewards`` + ``flags`` +
        ``gae``; updates the reward carry.  Returns total, comps, dones, successes (+ advantages, targets, returns)."""
        t = self._scored_steps(rec, n_steps)
        n, k = self.num_envs, self.program["TI_N_REW_COMP"]
        dev = self.device
        out = {"total": torch.empty((n, t), dtype=torch.float32, device=dev),
               "comps": torch.empty((max(k, 1), n, t), dtype=torch.float32, device=dev),
               "dones": torch.empty((n, t), dtype=torch.uint8, device=dev),
               "successes": torch.empty((n, t), dtype=torch.uint8, device=dev)}
        gae_out = [None, None, None]
        if values is not None:
            self._chk(values, (n, t), torch.float32, "values")
            for name in ("advantages", "targets", "returns"):
                out[name] = torch.empty((n, t), dtype=torch.float32, device=dev)
            gae_out = [_ptr(out["advantages"]), _ptr(out["targets"]), _ptr(out["returns"])]
        flags = (C.GAE_NORMALIZE_ADVANTAGES if normalize_advantages else 0) | (C.GAE_MONTE_CARLO_RETURNS if monte_carlo_returns else 0)
        with torch.cuda.device(dev):
            binding.check(self.lib.b200sim_score(self._handle, self._stream(), n, t, _ptr(rec), _ptr(self.levels), _ptr(self.reward_carry),
                                                 _ptr(values) if values is not None else None, ctypes.c_float(gamma), ctypes.c_float(lam),
                                                 flags, _ptr(out["total"]), _ptr(out["comps"]), _ptr(out["dones"]), _ptr(out["successes"]),
                                                 *gae_out), "b200sim_score")
        self.launches += 2 if (normalize_advantages and values is not None) else 1
        return out

    def flags(self, rec: torch.Tensor, n_steps: int) -> tuple[torch.Tensor, torch.Tensor]:
        n = self.num_envs
        dones = torch.empty((n, n_steps), dtype=torch.uint8, device=self.device)
        succ = torch.empty((n, n_steps), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            binding.check(self.lib.b200sim_extract_flags(self._handle, self._stream(), n, n_steps, _ptr(rec), _ptr(dones), _ptr(succ)),
                          "b200sim_extract_flags")
        self.launches += 1
        return dones, succ

    def gae(self, values: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor, successes: torch.Tensor, gamma: float,
            lam: float, normalize_advantages: bool = False, monte_carlo_returns: bool = False,
            ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``compute_ppo_inputs`` (ksim/task/ppo.py:54-121): advantages, value targets, returns, all [B][T]."""
        b, t = int(values.shape[0]), int(values.shape[1])
        self._chk(values, (b, t), torch.float32, "values")
        self._chk(rewards, (b, t), torch.float32, "rewards")
        self._chk(dones, (b, t), torch.uint8, "dones")
        self._chk(successes, (b, t), torch.uint8, "successes")
        adv, tgt, ret = (torch.empty((b, t), dtype=torch.float32, device=self.device) for _ in range(3))
        flags = (C.GAE_NORMALIZE_ADVANTAGES if normalize_advantages else 0) | (C.GAE_MONTE_CARLO_RETURNS if monte_carlo_returns else 0)
        with torch.cuda.device(self.device):
            binding.check(self.lib.b200sim_gae(self._stream(), b, t, _ptr(values), _ptr(rewards), _ptr(dones), _ptr(successes),
                                               ctypes.c_float(gamma), ctypes.c_float(lam), flags, _ptr(adv), _ptr(tgt), _ptr(ret)),
                          "b200sim_gae")
        self.launches += 2 if normalize_advantages else 1
        return adv, tgt, ret


def get_physics_engine(spec: TaskSpec, num_envs: int, **kwargs: object) -> B200Engine:
    """Counterpart of ``ksim.engine.get_physics_engine`` (engine.py:488-515) for the ``"b200"`` engine type."""
    return B200Engine(spec, num_envs, **kwargs)  # type: ignore[arg-type]
