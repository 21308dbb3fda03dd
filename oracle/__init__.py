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


def lib(precision: str = "f64") -> ctypes.CDLL:
    if precision not in ("f64", "f32"):
        raise ValueError(precision)
    if precision not in _LIBS:
        path = _DIR / f"liboracle_{precision}.so"
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
