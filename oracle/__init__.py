"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package (see ``oracle/oracle.h``).  The product (``ksim_b200``) never does.
"""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path
from typing import Any

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIBS: dict[str, ctypes.CDLL] = {}


def build(force: bool = False) -> None:
    """Compiles liboracle_f64.so / liboracle_f32.so with the committed Makefile."""
    if force:
        subprocess.run(["make", "-C", str(_DIR), "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)


_PERF_FLAGS = "-O3 -march=native -fPIC -std=c11 -fopenmp -ffp-contract=fast -w"
_PERF: Path | None = None


def build_perf() -> str:
    """A second f32 build tuned for THIS host (``-O3 -march=native``, contraction allowed) for the timing legs of ``bench.py``
    only: the CPU baseline should not be handicapped by the checker's conservative flags (``-O2 -ffp-contract=off``, which keep
    the f32 oracle's rounding reproducible).  Compiled where it runs (``-march=native`` must match the host), into
    ``oracle/_perf/``; every later ``lib("f32")`` of this process loads it.  Returns the flags used, or raises."""
    global _PERF
    out = _DIR / "_perf"
    out.mkdir(exist_ok=True)
    target = out / "liboracle_f32.so"
    cmd = ["gcc", *_PERF_FLAGS.split(), "-DREAL=float", "-DORACLE_F32", "-shared", "-o", str(target), "physics.c", "model_io.c", "engine.c", "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, cwd=str(_DIR))
    if "f32" in _LIBS:
        raise RuntimeError("build_perf() must run before the f32 oracle library is first loaded")
    _PERF = target
    return _PERF_FLAGS


def lib(precision: str = "f64") -> ctypes.CDLL:
    if precision not in ("f64", "f32"):
        raise ValueError(precision)
    if precision not in _LIBS:
        path = _DIR / f"liboracle_{precision}.so"
        if precision == "f32" and _PERF is not None:
            path = _PERF
        if not path.exists():
            build()
        cdll = ctypes.CDLL(str(path))
        cdll.o_model_new.restype = ctypes.c_void_p
        cdll.o_model_clone.restype = ctypes.c_void_p
        cdll.o_model_clone.argtypes = [ctypes.c_void_p]
        cdll.o_model_free.argtypes = [ctypes.c_void_p]
        cdll.o_data_new.restype = ctypes.c_void_p
        cdll.o_data_new.argtypes = [ctypes.c_void_p]
        cdll.o_data_free.argtypes = [ctypes.c_void_p]
        for fn in (cdll.o_model_set, cdll.o_model_get, cdll.o_data_set, cdll.o_data_get):
            fn.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
            fn.restype = ctypes.c_int
        for fn in (cdll.o_forward, cdll.o_step, cdll.o_kinematics_only, cdll.o_make_data):
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            fn.restype = None
        _LIBS[precision] = cdll
    return _LIBS[precision]


_MODEL_SCALARS = ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite", "nsensor", "nsensordata", "npair")


class OracleModel:
    """An ``OModel`` filled from a compiled ``ksim_b200.mjcf.Model``."""

    def __init__(self, model: Any, precision: str = "f64", _ptr: int | None = None) -> None:  # noqa: ANN401
        self.precision = precision
        self.lib = lib(precision)
        self.src = model
        if _ptr is not None:
            self.ptr = _ptr
            return
        self.ptr = self.lib.o_model_new()
        for name in _MODEL_SCALARS:
            self.set(name, [getattr(model, name)])
        o = model.opt
        self.set("timestep", [o.timestep])
        self.set("gravity", o.gravity)
        self.set("tolerance", [o.tolerance])
        self.set("ls_tolerance", [o.ls_tolerance])
        self.set("impratio", [o.impratio])
        self.set("meaninertia", [model.meaninertia])
        self.set("iterations", [o.iterations])
        self.set("ls_iterations", [o.ls_iterations])
        self.set("disable_refsafe", [int(o.disable_refsafe)])
        self.set("disable_warmstart", [int(o.disable_warmstart)])
        for name, arr in model.arrays.items():
            rc = self.set(name, arr, strict=False)
            if rc == 2:
                raise ValueError(f"model field {name} too large for the oracle")

    def set(self, name: str, values: Any, strict: bool = True) -> int:  # noqa: ANN401
        a = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
        rc = self.lib.o_model_set(self.ptr, name.encode(), a.ctypes.data, a.size)
        if strict and rc != 0:
            raise KeyError(f"o_model_set({name}) -> {rc}")
        return rc

    def get(self, name: str, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float64)
        rc = self.lib.o_model_get(self.ptr, name.encode(), out.ctypes.data, n)
        if rc != 0:
            raise KeyError(f"o_model_get({name}) -> {rc}")
        return out

    def clone(self) -> "OracleModel":
        return OracleModel(self.src, self.precision, _ptr=self.lib.o_model_clone(self.ptr))

    def __del__(self) -> None:
        try:
            self.lib.o_model_free(self.ptr)
        except Exception:  # noqa: BLE001
            pass


class OracleData:
    def __init__(self, model: OracleModel) -> None:
        self.model = model
        self.lib = model.lib
        self.ptr = self.lib.o_data_new(model.ptr)

    def set(self, name: str, values: Any) -> None:  # noqa: ANN401
        a = np.ascontiguousarray(np.asarray(values, dtype=np.float64).ravel())
        rc = self.lib.o_data_set(self.ptr, name.encode(), a.ctypes.data, a.size)
        if rc != 0:
            raise KeyError(f"o_data_set({name}) -> {rc}")

    def get(self, name: str, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float64)
        rc = self.lib.o_data_get(self.ptr, name.encode(), out.ctypes.data, n)
        if rc != 0:
            raise KeyError(f"o_data_get({name}) -> {rc}")
        return out

    def mat(self, name: str, rows: int, cols: int, stride: int | None = None) -> np.ndarray:
        """Reads a 2-D field whose C row stride is ``stride`` (defaults to ``cols``)."""
        stride = cols if stride is None else stride
        flat = self.get(name, rows * stride)
        return flat.reshape(rows, stride)[:, :cols].copy()

    def forward(self) -> None:
        self.lib.o_forward(self.model.ptr, self.ptr)

    def step(self) -> None:
        self.lib.o_step(self.model.ptr, self.ptr)

    def kinematics(self) -> None:
        self.lib.o_kinematics_only(self.model.ptr, self.ptr)

    def __del__(self) -> None:
        try:
            self.lib.o_data_free(self.ptr)
        except Exception:  # noqa: BLE001
            pass


# ---------------------------------------------------------------------------------------------------------
# engine-level entry points (oracle/engine.c): env-major buffers, layouts from the task program ti[]
# ---------------------------------------------------------------------------------------------------------


class OracleEngine:
    """Batch front-end of o_init_envs / o_step_ctrl / o_rewards / o_gae for a ``ksim_b200.program.TaskProgram``."""

    def __init__(self, program: Any, precision: str = "f32") -> None:  # noqa: ANN401
        self.program = program
        self.precision = precision
        self.real = np.float32 if precision == "f32" else np.float64
        self.lib = lib(precision)
        self.model = OracleModel(program.spec.model, precision)
        self.ti = np.ascontiguousarray(program.ti, dtype=np.int32)
        self.tf = np.ascontiguousarray(program.tf.astype(self.real))
        self.S, self.R, self.K = program.state_size, program.rec_size, program.key_size
        self.RAND, self.NA = program.rand_size, program.action_size
        for fn in (self.lib.o_init_envs, self.lib.o_step_ctrl, self.lib.o_rewards, self.lib.o_gae, self.lib.o_ppo_loss):
            fn.restype = ctypes.c_int

    @staticmethod
    def _p(a: np.ndarray | None) -> ctypes.c_void_p:
        return ctypes.c_void_p(None if a is None else a.ctypes.data)

    def init_envs(self, init_keys: np.ndarray, levels: np.ndarray, rand: np.ndarray) -> tuple[np.ndarray, ...]:
        n = init_keys.shape[0]
        init_keys = np.ascontiguousarray(init_keys, dtype=np.uint32).reshape(n, 10)
        levels = np.ascontiguousarray(levels, dtype=self.real)
        rand = np.ascontiguousarray(rand, dtype=self.real).reshape(n, self.RAND)
        state = np.zeros((n, self.S), dtype=self.real)
        keys = np.zeros((n, self.K), dtype=np.uint32)
        rec0 = np.zeros((n, self.R), dtype=self.real)
        act_keys = np.zeros((n, 2), dtype=np.uint32)
        rc = self.lib.o_init_envs(ctypes.c_void_p(self.model.ptr), self._p(self.ti), self._p(self.tf), ctypes.c_int(n),
                                  self._p(init_keys), self._p(levels), self._p(rand), self._p(state), self._p(keys),
                                  self._p(rec0), self._p(act_keys))
        if rc != 0:
            raise RuntimeError(f"o_init_envs -> {rc}")
        return state, keys, rec0, act_keys

    def step_ctrl(self, action: np.ndarray, rand: np.ndarray, state: np.ndarray, keys: np.ndarray,
                  rec_cur: np.ndarray, rec_next: np.ndarray, act_keys: np.ndarray | None = None) -> None:
        n = state.shape[0]
        action = np.ascontiguousarray(action, dtype=self.real).reshape(n, self.NA)
        rand = np.ascontiguousarray(rand, dtype=self.real).reshape(n, self.RAND)
        for a in (state, keys, rec_cur, rec_next):
            assert a.flags["C_CONTIGUOUS"]
        rc = self.lib.o_step_ctrl(ctypes.c_void_p(self.model.ptr), self._p(self.ti), self._p(self.tf), ctypes.c_int(n),
                                  self._p(action), self._p(rand), self._p(state), self._p(keys), self._p(rec_cur),
                                  self._p(rec_next), self._p(act_keys))
        if rc != 0:
            raise RuntimeError(f"o_step_ctrl -> {rc}")

    # ---- the engine alone (PhysicsEngine.reset / .step): data blocks in the B2S_DATA_* layout of include/b200sim_codes.h
    def data_size(self, ncon: int) -> int:
        return int(self.lib.o_data_size(ctypes.c_void_p(self.model.ptr), ctypes.c_int(ncon)))

    def engine_reset(self, rng: np.ndarray, levels: np.ndarray, rand: np.ndarray, ncon: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """rng [N][2] uint32 -> (state [N][S], keys [N][K], data [N][DATA])."""
        n = rng.shape[0]
        rng = np.ascontiguousarray(rng, dtype=np.uint32).reshape(n, 2)
        levels = np.ascontiguousarray(levels, dtype=self.real)
        rand = np.ascontiguousarray(rand, dtype=self.real).reshape(n, self.RAND)
        state = np.zeros((n, self.S), dtype=self.real)
        keys = np.zeros((n, self.K), dtype=np.uint32)
        data = np.zeros((n, self.data_size(ncon)), dtype=self.real)
        rc = self.lib.o_engine_reset(ctypes.c_void_p(self.model.ptr), self._p(self.ti), self._p(self.tf), ctypes.c_int(n), self._p(rng),
                                     self._p(levels), self._p(rand), self._p(state), self._p(keys), ctypes.c_int(ncon), self._p(data))
        if rc != 0:
            raise RuntimeError(f"o_engine_reset -> {rc}")
        return state, keys, data

    def engine_step(self, action: np.ndarray, rng: np.ndarray, rand: np.ndarray, state: np.ndarray, keys: np.ndarray, ncon: int) -> np.ndarray:
        """Advances ``state`` / ``keys`` in place by one control step of the engine alone; returns data [N][DATA]."""
        n = state.shape[0]
        action = np.ascontiguousarray(action, dtype=self.real).reshape(n, self.NA)
        rng = np.ascontiguousarray(rng, dtype=np.uint32).reshape(n, 2)
        rand = np.ascontiguousarray(rand, dtype=self.real).reshape(n, self.RAND)
        assert state.flags["C_CONTIGUOUS"] and keys.flags["C_CONTIGUOUS"] and state.dtype == self.real
        data = np.zeros((n, self.data_size(ncon)), dtype=self.real)
        rc = self.lib.o_engine_step(ctypes.c_void_p(self.model.ptr), self._p(self.ti), self._p(self.tf), ctypes.c_int(n), self._p(action),
                                    self._p(rng), self._p(rand), self._p(state), self._p(keys), ctypes.c_int(ncon), self._p(data))
        if rc != 0:
            raise RuntimeError(f"o_engine_step -> {rc}")
        return data

    def rollout(self, actions: np.ndarray, rand: np.ndarray, state: np.ndarray, keys: np.ndarray, rec: np.ndarray) -> None:
        """actions [T][N][NA]; rec [T+1][N][R] with slot 0's pre-step block valid."""
        for t in range(actions.shape[0]):
            self.step_ctrl(actions[t], rand, state, keys, rec[t], rec[t + 1])

    def rewards(self, rec: np.ndarray, levels: np.ndarray, carry: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        t, n = rec.shape[0], rec.shape[1]
        k = int(self.ti[31])  # TI_N_REW_COMP
        rec = np.ascontiguousarray(rec, dtype=self.real)
        levels = np.ascontiguousarray(levels, dtype=self.real)
        assert carry.dtype == self.real and carry.flags["C_CONTIGUOUS"]
        total = np.zeros((n, t), dtype=self.real)
        comps = np.zeros((k, n, t), dtype=self.real)
        rc = self.lib.o_rewards(ctypes.c_void_p(self.model.ptr), self._p(self.ti), self._p(self.tf), ctypes.c_int(n),
                                ctypes.c_int(t), self._p(rec), self._p(levels), self._p(carry), self._p(total), self._p(comps))
        if rc != 0:
            raise RuntimeError(f"o_rewards -> {rc}")
        return total, comps

    def gae(self, values: np.ndarray, rewards: np.ndarray, dones: np.ndarray, successes: np.ndarray, gamma: float,
            lam: float, flags: int) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        b, t = values.shape
        values = np.ascontiguousarray(values, dtype=self.real)
        rewards = np.ascontiguousarray(rewards, dtype=self.real)
        dones = np.ascontiguousarray(dones, dtype=np.uint8)
        successes = np.ascontiguousarray(successes, dtype=np.uint8)
        adv, tgt, ret = (np.zeros((b, t), dtype=self.real) for _ in range(3))
        real_c = ctypes.c_float if self.precision == "f32" else ctypes.c_double
        rc = self.lib.o_gae(ctypes.c_int(b), ctypes.c_int(t), self._p(values), self._p(rewards), self._p(dones),
                            self._p(successes), real_c(gamma), real_c(lam), ctypes.c_int(flags), self._p(adv),
                            self._p(tgt), self._p(ret))
        if rc != 0:
            raise RuntimeError(f"o_gae -> {rc}")
        return adv, tgt, ret

    def ppo_loss(self, logp_on: np.ndarray, logp_off: np.ndarray, values_on: np.ndarray, values_off: np.ndarray, adv: np.ndarray,
                 targets: np.ndarray, entropy: np.ndarray | None, clip_param: float, value_loss_coef: float, entropy_coef: float,
                 kl_coef: float, log_clip_value: float, kl_scale: float, flags: int,
                 ) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """o_ppo_loss: (losses[4][B][T], out[5], d_logp[B][T], d_values[B][T])."""
        b, t, a = logp_on.shape
        arrs = [np.ascontiguousarray(x, dtype=self.real) for x in (logp_on, logp_off, values_on, values_off, adv, targets)]
        ent = None if entropy is None else np.ascontiguousarray(entropy, dtype=self.real)
        losses = np.zeros((4, b, t), dtype=self.real)
        out = np.zeros(5, dtype=self.real)
        d_logp, d_values = (np.zeros((b, t), dtype=self.real) for _ in range(2))
        real_c = ctypes.c_float if self.precision == "f32" else ctypes.c_double
        rc = self.lib.o_ppo_loss(ctypes.c_int(b), ctypes.c_int(t), ctypes.c_int(a), *(self._p(x) for x in arrs),
                                 None if ent is None else self._p(ent), real_c(clip_param), real_c(value_loss_coef),
                                 real_c(entropy_coef), real_c(kl_coef), real_c(log_clip_value), real_c(kl_scale),
                                 ctypes.c_int(flags), self._p(losses), self._p(out), self._p(d_logp), self._p(d_values))
        if rc != 0:
            raise RuntimeError(f"o_ppo_loss -> {rc}")
        return losses, out, d_logp, d_values
