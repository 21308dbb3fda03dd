"""ctypes binding of the C-ABI in ``include/b200sim.h`` (``ksim_b200/csrc/libb200sim.so``).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""

from __future__ import annotations

__all__ = ["B200SimError", "lib", "library_path", "check", "EXPORTS", "INFO"]

import ctypes
from pathlib import Path

_CSRC = Path(__file__).resolve().parent / "csrc"
_LIB: ctypes.CDLL | None = None

# every symbol include/b200sim.h declares (tests assert the library exports exactly these)
EXPORTS = (
    "b200sim_model_create", "b200sim_model_destroy", "b200sim_model_info", "b200sim_model_set_option", "b200sim_last_error", "b200sim_init_envs",
    "b200sim_step_ctrl", "b200sim_rollout", "b200sim_rewards", "b200sim_gae", "b200sim_extract_flags", "b200sim_score", "b200sim_dataset_pack",
    "b200sim_metrics", "b200sim_ppo_loss",
    "b200sim_gru_workspace_bytes", "b200sim_gru_saved_floats", "b200sim_gru_forward", "b200sim_gru_backward", "b200sim_gru_last_error",
    "b200sim_mixture_logprob", "b200sim_mixture_logprob_backward", "b200sim_mixture_sample", "b200sim_policy_last_error",
    "b200sim_data_layout", "b200sim_engine_reset", "b200sim_engine_step", "b200sim_terminations", "b200sim_engine_last_error",
    "b200sim_histogram_scratch_bytes", "b200sim_histogram", "b200sim_row_sums",
    "b200sim_optimizer_scratch_bytes", "b200sim_optimizer_step", "b200sim_optimizer_last_error",
)

INFO = dict(STATE_SIZE=0, REC_SIZE=1, KEY_SIZE=2, RAND_SIZE=3, ACTION_SIZE=4, N_REW_COMP=5, REW_CARRY_SIZE=6,
            SMEM_PER_ENV=7, WARPS_PER_CTA=8, VARIANT=9, STAGED_ARRAYS=10)


class B200SimError(RuntimeError):
    pass


def library_path() -> Path:
    return _CSRC / "libb200sim.so"


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = library_path()
        if not path.exists():
            raise B200SimError(
                f"{path} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback."
            )
        cdll = ctypes.CDLL(str(path))
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        cdll.b200sim_model_create.argtypes = [vp, ctypes.c_size_t, ctypes.POINTER(vp)]
        cdll.b200sim_model_create.restype = ci
        cdll.b200sim_model_destroy.argtypes = [vp]
        cdll.b200sim_model_destroy.restype = None
        cdll.b200sim_model_info.argtypes = [vp, ci]
        cdll.b200sim_model_info.restype = ci
        cdll.b200sim_model_set_option.argtypes, cdll.b200sim_model_set_option.restype = [vp, ci, ci], ci
        cdll.b200sim_last_error.argtypes = []
        cdll.b200sim_last_error.restype = ctypes.c_char_p
        cdll.b200sim_init_envs.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp]
        cdll.b200sim_init_envs.restype = ci
        cdll.b200sim_step_ctrl.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, vp, vp]
        cdll.b200sim_step_ctrl.restype = ci
        cdll.b200sim_rollout.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
        cdll.b200sim_rollout.restype = ci
        cdll.b200sim_rewards.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
        cdll.b200sim_rewards.restype = ci
        cdll.b200sim_score.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, cf, cf, ci, vp, vp, vp, vp, vp, vp, vp]
        cdll.b200sim_score.restype = ci
        cdll.b200sim_gae.argtypes = [vp, ci, ci, vp, vp, vp, vp, cf, cf, ci, vp, vp, vp]
        cdll.b200sim_gae.restype = ci
        cdll.b200sim_extract_flags.argtypes = [vp, vp, ci, ci, vp, vp, vp]
        cdll.b200sim_extract_flags.restype = ci
        cdll.b200sim_dataset_pack.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, ci, ci, vp]
        cdll.b200sim_dataset_pack.restype = ci
        cdll.b200sim_metrics.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
        cdll.b200sim_metrics.restype = ci
        cdll.b200sim_ppo_loss.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, cf, cf, cf, cf, cf, cf, ci, vp, vp, vp, vp, vp]
        cdll.b200sim_ppo_loss.restype = ci
        sz, u32p = ctypes.c_size_t, vp
        cdll.b200sim_gru_workspace_bytes.argtypes, cdll.b200sim_gru_workspace_bytes.restype = [ci, ci, ci, ci], sz
        cdll.b200sim_gru_saved_floats.argtypes, cdll.b200sim_gru_saved_floats.restype = [ci, ci], sz
        cdll.b200sim_gru_forward.argtypes, cdll.b200sim_gru_forward.restype = [vp, ci, vp, ci, ci, vp, vp, sz], ci
        cdll.b200sim_gru_backward.argtypes, cdll.b200sim_gru_backward.restype = [vp, ci, vp, ci, ci, vp, vp, sz], ci
        cdll.b200sim_gru_last_error.argtypes, cdll.b200sim_gru_last_error.restype = [], ctypes.c_char_p
        cdll.b200sim_mixture_logprob.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp, cf, cf, cf, vp]
        cdll.b200sim_mixture_logprob.restype = ci
        cdll.b200sim_mixture_logprob_backward.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp, cf, cf, cf, vp, ci, vp, ci]
        cdll.b200sim_mixture_logprob_backward.restype = ci
        cdll.b200sim_mixture_sample.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, cf, cf, cf, u32p, vp, ci, vp, vp]
        cdll.b200sim_mixture_sample.restype = ci
        cdll.b200sim_policy_last_error.argtypes, cdll.b200sim_policy_last_error.restype = [], ctypes.c_char_p
        cdll.b200sim_data_layout.argtypes, cdll.b200sim_data_layout.restype = [vp, vp], ci
        cdll.b200sim_engine_reset.argtypes, cdll.b200sim_engine_reset.restype = [vp, vp, ci, vp, vp, vp, vp, vp, vp], ci
        cdll.b200sim_engine_step.argtypes, cdll.b200sim_engine_step.restype = [vp, vp, ci, vp, vp, vp, vp, vp, vp], ci
        cdll.b200sim_terminations.argtypes, cdll.b200sim_terminations.restype = [vp, vp, ci, vp, vp, vp, vp, vp], ci
        cdll.b200sim_engine_last_error.argtypes, cdll.b200sim_engine_last_error.restype = [], ctypes.c_char_p
        cdll.b200sim_histogram_scratch_bytes.argtypes, cdll.b200sim_histogram_scratch_bytes.restype = [], sz
        cdll.b200sim_histogram.argtypes, cdll.b200sim_histogram.restype = [vp, sz, vp, ci, vp, vp, vp, vp], ci
        cdll.b200sim_row_sums.argtypes, cdll.b200sim_row_sums.restype = [vp, ci, ci, vp, vp], ci
        cdll.b200sim_optimizer_scratch_bytes.argtypes, cdll.b200sim_optimizer_scratch_bytes.restype = [], sz
        cdll.b200sim_optimizer_step.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, cf, cf, cf, ci, cf, cf, cf, cf, cf, vp, vp]
        cdll.b200sim_optimizer_step.restype = ci
        cdll.b200sim_optimizer_last_error.argtypes, cdll.b200sim_optimizer_last_error.restype = [], ctypes.c_char_p
        _LIB = cdll
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b200sim_last_error()
        raise B200SimError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
