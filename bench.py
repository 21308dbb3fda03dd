"""Benchmark of the batched-rollout hot path (BASELINE.json metric: humanoid env-steps/sec).

A "step" = one PPO-iteration's worth of the hot path over one batch of synthetic input:
rollout of `envs` humanoid environments x `rollout` control steps (5 physics sub-steps each, with actuators,
events, observations, terminations, reset-on-done, commands) -> rewards -> done/success flags -> GAE.
Workload = BASELINE.json configs[1]: humanoid walking 4096 envs x rollout 64 on 1 B200; with --gpus N each rank owns
4096 envs (weak scaling, configs[3] at N=8).  Policy = random-init surrogate: a_t = 0.3 N(0,1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

ENVS_PER_GPU = 4096
ROLLOUT = 64
L2_NOTE = ""  # set by main(): inputs larger than L2, or an explicit flush between timed iterations
METRIC = "humanoid env-steps/sec"
UNIT = "env-steps/s"
GAMMA, LAM = 0.97, 0.95
KBOT_SOLVER: str | None = None
WORKLOAD = "humanoid"  # --workload kbot: BASELINE.json configs[2] (K-Bot, 8192 envs x 100), a secondary measurement


def _task():  # noqa: ANN202
    """(make_spec, RANDOMIZERS) of the selected workload."""
    if WORKLOAD == "kbot":
        from ksim_b200.examples import kbot as task

        if KBOT_SOLVER is not None:  # --kbot-solver fast: the opt-in solver (the task ships "reference", examples/kbot.py)
            return (lambda: task.make_spec(task.KbotConfig(solver=KBOT_SOLVER))), task.RANDOMIZERS
    else:
        from ksim_b200.examples import walking as task
    return task.make_spec, task.RANDOMIZERS


def _peaks() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.samples, self.proc, self.first = index, [], None, 0

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def mark(self) -> None:
        """The timed region starts here: samples taken earlier (while nvidia-smi was starting up under the warm-up load,
        where its NVML initialisation cannot perturb a timed step) are dropped from the report."""
        self.first = len(self.samples)

    def _read(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in (self.samples[self.first:] or self.samples):
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


PPO_UPDATES_PER_STEP = 32   # num_passes 4 x (4096 envs / batch 512): gradient all-reduces of one PPO iteration (examples/walking.py:772-776)
PPO_GRAD_FLOATS = 7_809_024  # actor 43->512->2xGRU(512)->MLP->510 + critic 340->512->2xGRU(512)->MLP->1


def _config(n_gpus: int) -> dict:
    """The same dict for both arms (the reference arm times a bounded SAMPLE of this workload and says so in ``cpu_baseline``)."""
    name = ("K-Bot walking (examples/kbot/train.py task incl. its own commands / observations / rewards, level 1: actuator noise, latency, "
            "drop, force pushes, seven randomizers)"
            if WORKLOAD == "kbot" else "humanoid walking (examples/walking.py plugin set)")
    return {"workload": f"{name} {ENVS_PER_GPU} envs x rollout {ROLLOUT} per GPU: physics+obs+termination+"
                        "reset+reward+GAE", "envs_per_gpu": ENVS_PER_GPU, "rollout": ROLLOUT, "total_envs": ENVS_PER_GPU * n_gpus,
            "physics_substeps_per_env_step": 5, "policy": "synthetic: a_t = 0.3*N(0,1), values = 0", "parallelism": f"env-sharded x{n_gpus}",
            "collective": (f"{PPO_UPDATES_PER_STEP} x NCCL all-reduce of the {PPO_GRAD_FLOATS * 4 / 1e6:.1f} MB flat PPO gradient inside every timed "
                           "step when n_gpus > 1 (BASELINE.json configs[3]); none at n_gpus = 1")}


def _cpu_arm(seconds: float | None, steps: int = 0, warmup: int = 0) -> dict:
    """The CPU implementation of the same path (the C oracle, f32, OpenMP over envs) timed on the host cores: either `steps`
    timed steps after `warmup`, or as many steps as fit in `seconds`.  One step is a bounded SAMPLE of the workload:
    64 envs per host core x a 64-step rollout -> rewards -> GAE.

    The reference's own CPU path is MJX on the JAX CPU backend; neither jax nor mujoco exists in this image, so
    `kind` is "port".  The timing legs use a host-tuned build of the port (-O3 -march=native, oracle.build_perf) when it
    compiles on this box, else the checker's build; `build` says which.
    """
    import ctypes

    import oracle
    from ksim_b200.envinit import make_env_init
    from ksim_b200.program import build_program
    from oracle import OracleEngine

    make_spec, RANDOMIZERS = _task()
    oracle.build()
    try:
        build = "perf: gcc " + oracle.build_perf()
    except Exception as e:  # noqa: BLE001  (no compiler on the box, or the f32 library is already loaded)
        build = f"checker build (-O2 -ffp-contract=off); perf build unavailable: {type(e).__name__}"
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to its workers (N > 1); the arm runs on rank 0 alone and is meant to use every
    # host core, so the OpenMP runtime the oracle library is linked to is told so directly.
    os.environ["OMP_NUM_THREADS"] = str(cores)
    omp_set = oracle.lib("f32").omp_set_num_threads
    omp_set.argtypes, omp_set.restype = [ctypes.c_int], None
    omp_set(cores)
    cores = int(oracle.lib("f32").omp_get_max_threads())  # what the library will actually fan out to
    spec = make_spec()
    prog = build_program(spec)
    rz = {k: f(spec.model) for k, f in RANDOMIZERS.items()}
    n, t = max(cores * 64, 256), 64  # one step = 64 envs per core x a 64-step rollout: 0.5-2 s of work on all cores
    ora = OracleEngine(prog, "f32")
    init = make_env_init(prog, rz, n, seed=0, level=1.0 if WORKLOAD == "kbot" else 0.0)
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size), np.float32)
    rec[0] = rec0
    actions = (0.3 * np.random.RandomState(1).randn(t, n, prog.action_size)).astype(np.float32)
    carry = np.zeros((n, prog["TI_REW_CARRY_SIZE"]), np.float32)

    def one() -> None:
        ora.rollout(actions, init.rand, st, keys, rec)
        total, _ = ora.rewards(rec[:t], init.levels, carry)
        d, s = prog["TI_R_DONE"], prog["TI_R_SUCCESS"]
        ora.gae(np.zeros_like(total), total, (rec[:t, :, d] != 0).T, (rec[:t, :, s] != 0).T, GAMMA, LAM, 0)
        rec[0] = rec[t]

    for _ in range(max(warmup, 1)):
        one()
    reps, t0 = 0, time.perf_counter()
    while (reps < steps) if seconds is None else (time.perf_counter() - t0 < seconds or reps == 0):
        one()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    value = n * t / dt
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "per_core": value / cores, "build": build,
            "sample_envs": n, "sample_steps": t, "sample_reps": reps, "ms_per_sample": dt * 1e3,
            "sample": f"{reps} x ({n} envs x {t} control steps: rollout + rewards + GAE, same plugin set) of the {ENVS_PER_GPU} x {ROLLOUT} "
                      "workload, f32 C port of the path, OpenMP over envs"}


def run_reference(args: argparse.Namespace) -> None:
    """`--impl reference`: rank 0 alone times the CPU arm and prints its line; the other ranks exit quietly."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cb = _cpu_arm(None, steps=args.steps, warmup=args.warmup)
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cb["ms_per_sample"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": _config(args.gpus), "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_leg() -> dict:
    """Bounded CPU sample (~10 s) of the same workload for the `cpu_baseline` object (rank 0, N=1 only)."""
    return _cpu_arm(10.0)


_JSON_FD = 1


def _emit(line: str) -> None:
    """The ONE stdout line of the contract.  Everything else a library prints to fd 1 (NCCL announces its version there when
    the first communicator is built) has been routed to stderr by main()."""
    os.write(_JSON_FD, (line + "\n").encode())


def _extras(eng, actions, rec, device, rank: int, world: int, args: argparse.Namespace) -> dict:  # noqa: ANN001
    """Measurements beside the headline, every one device-timed after its own warm-up, max over ranks:

    * ``closed_loop``: the rollout as a real PPO loop drives it -- ONE ``b200sim_step_ctrl`` launch per control step (the policy
      would run between launches; here the next action block is already resident), same environments, same workload;
    * ``ppo_iteration`` (humanoid only): ``ksim_b200.trainer.PPOTrainer.iteration`` -- closed-loop rollout with the random-init
      GRU policy, rewards, on-policy replay, 32 minibatch steps (replay + GAE + loss + backward + optimiser; at N > 1 the NCCL
      all-reduce of the flat gradient is inside every captured step), closing replay -- at the reference's default rollout of
      24 frames and at this benchmark's rollout;
    * ``secondary`` (N = 1, humanoid runs only): the K-Bot workload (BASELINE.json configs[2]) through the same hot path."""
    import torch

    from ksim_b200 import distributed as D

    out: dict = {}
    t, n = int(actions.shape[0]), eng.num_envs
    stream = torch.cuda.current_stream(device)

    def timed(fn, reps: int) -> float:  # noqa: ANN001
        fn()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize(device)
        return D.max_over_ranks(a.elapsed_time(b) / reps, device)

    def closed_loop() -> None:
        rec[0].copy_(eng.first_rec)
        for s in range(t):
            eng.step(actions[s], rec[s], rec[s + 1])
        eng.first_rec.copy_(rec[t])

    l0 = eng.launches
    ms = timed(closed_loop, max(args.steps // 2, 2))
    out["closed_loop"] = {"value": n * t * world / (ms * 1e-3), "unit": UNIT, "ms_per_rollout": ms, "launches_per_rollout": t,
                          "gpu_launches": eng.launches - l0, "what": "one b200sim_step_ctrl launch per control step, actions resident"}
    if WORKLOAD == "humanoid" and ENVS_PER_GPU % 512 == 0:
        from ksim_b200.examples import walking
        from ksim_b200.trainer import PPOConfig, PPOTrainer

        ppo_out = {}
        for frames in sorted({24, t}):
            torch.manual_seed(0)
            model = walking.make_model(eng.spec).to(device)
            tr = PPOTrainer(eng, model, walking.network_columns(eng.program, walking.ACTOR_INPUTS),
                            walking.network_columns(eng.program, walking.CRITIC_INPUTS),
                            PPOConfig(batch_size=512, num_passes=4, rollout_length_frames=frames), use_graphs=True, seed=rank)
            for _ in range(2):
                tr.iteration()
            if world > 1:
                torch.distributed.barrier()
            runs = [tr.iteration() for _ in range(3)]
            mean = {k: D.max_over_ranks(sum(r[k] for r in runs) / len(runs), device) for k in runs[0]}
            ppo_out[f"rollout_{frames}"] = {"iteration_ms": mean["iteration_ms"], "rollout_ms": mean["rollout_ms"], "scoring_ms": mean["scoring_ms"],
                                            "update_ms": mean["update_ms"], "env_steps_per_s": n * frames * world / (mean["iteration_ms"] * 1e-3),
                                            "final_loss": float(tr.loss)}
            del tr, model
            torch.cuda.empty_cache()
        # the replay kernels alone: one eager update with CUDA events around every GRU launch (ksim_b200.replay.KERNEL_TIMES)
        gru = None
        try:
            if world > 1:   # a per-kernel figure: measured at N = 1 only, which also keeps the N > 1 run free of eager collectives
                raise _SkipGru
            from ksim_b200 import replay

            torch.manual_seed(0)
            model = walking.make_model(eng.spec).to(device)
            tr = PPOTrainer(eng, model, walking.network_columns(eng.program, walking.ACTOR_INPUTS),
                            walking.network_columns(eng.program, walking.CRITIC_INPUTS),
                            PPOConfig(batch_size=512, num_passes=4, rollout_length_frames=24), use_graphs=False, seed=rank)
            tr.iteration()
            tr.rollout(); tr.score()
            torch.cuda.synchronize(device)
            replay.KERNEL_TIMES = []
            tr.update_model()
            torch.cuda.synchronize(device)
            times, replay.KERNEL_TIMES = replay.KERNEL_TIMES, None
            peak_tf = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("bf16_tflops_sustained", 1385.4)) if (ROOT / "MEASURED_PEAKS.json").exists() else 1385.4
            gru = {"peak": peak_tf, "unit": "TFLOP/s", "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)",
                   "shape": "B = 512 rows x T = 24 steps x 2 networks per launch, GRU(512): 2 T B 3H H flops per network (the sequential h W_hh^T / dgh W_hh products only)"}
            for d in ("fwd", "bwd"):
                ms_ = [a.elapsed_time(b) for (dd, t_, b_, ns, a, b) in times if dd == d and b_ == 512]
                if ms_:
                    flops = 2.0 * 24 * 512 * 2 * (3 * 512) * 512
                    m_ = float(np.median(ms_))
                    gru[d] = {"launches": len(ms_), "ms_per_launch": m_, "achieved": flops / (m_ * 1e-3) / 1e12, "frac": flops / (m_ * 1e-3) / 1e12 / peak_tf}
            del tr, model
            torch.cuda.empty_cache()
        except _SkipGru:
            gru = {"note": "measured at n_gpus = 1 only"}
        except Exception as e:  # noqa: BLE001
            gru = {"error": f"{type(e).__name__}: {e}"[:200]}
        out["ppo_iteration"] = {
            **ppo_out, "gru_kernels": gru, "unit": "ms", "envs_per_gpu": n, "batch_size": 512, "num_passes": 4, "updates_per_iteration": 4 * (n // 512),
            "gradient_allreduce": (f"in-place NCCL all-reduce of the {PPO_GRAD_FLOATS * 4 / 1e6:.1f} MB flat gradient inside every captured "
                                   "minibatch step") if world > 1 else "none (single rank)",
            "ours": "engine step, rewards, flags, GAE, PPO loss fwd+bwd, GRU sequence kernels (tcgen05), mixture head, optimiser step",
            "library": "batched affine maps (cuBLAS bf16 through torch), gathers, NCCL"}
    return out


class _SkipGru(Exception):
    pass


def _mjx_baseline() -> dict:
    """The reference's own MjxEngine on the JAX CPU backend (baseline/run_mjx_cpu.py) when jax + mujoco-mjx + a reference
    checkout exist on this box; otherwise the one-line reason (they do not exist in this repository's image)."""
    try:
        res = subprocess.run([sys.executable, str(ROOT / "baseline" / "run_mjx_cpu.py"), "--envs", "256", "--steps", "64", "--reps", "1"],
                             capture_output=True, text=True, timeout=900, cwd=str(ROOT))
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"impl": "mjx", "unavailable": f"{type(e).__name__}: {e}"[:200]}


def _secondary_kbot(args: argparse.Namespace, solver: str | None = None) -> dict | None:
    """The K-Bot line (BASELINE.json configs[2]) from a child process of the same script, attached to the humanoid line.
    ``solver`` None: the task as it ships (reference-shaped capped Newton iteration); "fast": the opt-in solver."""
    cmd = [sys.executable, str(ROOT / "bench.py"), "--workload", "kbot", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-extras"]
    if solver is not None:
        cmd += ["--kbot-solver", solver]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
        line = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}
    keep = ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "roofline", "config")
    return {k: line[k] for k in keep if k in line}


def main() -> None:
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="humanoid", choices=["humanoid", "kbot"])
    ap.add_argument("--envs", type=int, default=None, help="environments per GPU (env-count sweeps, BASELINE.json configs[4]); "
                    "the default is the configuration the metric is quoted on")
    ap.add_argument("--rollout", type=int, default=None, help="control steps per rollout")
    ap.add_argument("--kbot-solver", default=None, choices=["fast", "reference"], help="K-Bot workload: override the task's solver option")
    ap.add_argument("--no-extras", action="store_true", help="skip the closed-loop, PPO-iteration and K-Bot secondary measurements")
    args = ap.parse_args()
    global WORKLOAD, ENVS_PER_GPU, ROLLOUT, METRIC, GAMMA, LAM, KBOT_SOLVER
    KBOT_SOLVER = args.kbot_solver
    if args.workload == "kbot":
        WORKLOAD, ENVS_PER_GPU, ROLLOUT, METRIC, GAMMA, LAM = "kbot", 8192, 100, "kbot env-steps/sec", 0.94, 0.94
    if args.envs is not None:
        ENVS_PER_GPU = args.envs
    if args.rollout is not None:
        ROLLOUT = args.rollout
    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    from ksim_b200 import distributed as D
    from ksim_b200.engine import B200Engine

    make_spec, RANDOMIZERS = _task()
    rank, local, world = D.init_from_env(timeout_s=240.0)  # a collective mismatch fails within minutes instead of the default 10
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    spec = make_spec()
    rz = {k: f(spec.model) for k, f in RANDOMIZERS.items()}
    n, t = ENVS_PER_GPU, ROLLOUT
    lo, _ = D.shard_range(n * world, rank, world)
    level = 1.0 if WORKLOAD == "kbot" else 0.0  # the K-Bot task pins its curriculum at 1 (train.py:1268-1273)
    eng = B200Engine(spec, n, device=device, randomizers=rz, seed=0, curriculum_level=level, env_offset=lo, total_envs=n * world)
    na = eng.NA
    # synthetic policy outputs for every step of the run, resident in HBM before timing starts
    gen = torch.Generator(device=device).manual_seed(1 + rank)
    actions = 0.3 * torch.randn((t, n, na), generator=gen, device=device)
    values = torch.zeros((n, t), device=device)
    rec = eng.new_record_buffer(t)
    stream = torch.cuda.current_stream(device)
    # BASELINE.json configs[3]: "per-GPU env slices + NCCL PPO grad allreduce" -- at N > 1 every timed step carries the gradient
    # all-reduces of one PPO iteration on the flat gradient buffer; nothing at N = 1
    grad = torch.zeros(PPO_GRAD_FLOATS, device=device) if world > 1 else None

    def collective() -> None:
        if grad is not None:
            for _ in range(PPO_UPDATES_PER_STEP):
                torch.distributed.all_reduce(grad)

    def hot_path(acts: torch.Tensor, with_collective: bool = True) -> torch.Tensor:
        eng.rollout(acts, rec)
        adv = eng.score(rec, t, values=values, gamma=GAMMA, lam=LAM)["advantages"]  # rewards + flags + GAE: one launch
        if with_collective:
            collective()
        return adv  # (the engine hands this rollout's final pre-step block to the next one itself)

    def barrier() -> None:
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # spawned before the warm-up so that its start-up cost is not paid inside the timed region
    for _ in range(max(args.warmup, 3)):
        hot_path(actions)
    t_wait = time.perf_counter()
    forced = int(os.environ.get("B2S_BENCH_FORCE_WAIT_STEPS", "0"))  # test hook: make the rank-0-only wait loop run
    while sampler.proc is not None and ((not sampler.samples and time.perf_counter() - t_wait < 3.0) or forced > 0):
        # extra untimed warm-up until nvidia-smi has printed its first sample.  RANK 0 ONLY, so it must not contain a collective:
        # the other ranks are already waiting at the barrier below (with the all-reduces in it this loop deadlocked an 8-GPU run
        # whenever nvidia-smi was slow to start)
        hot_path(actions, with_collective=False)
        torch.cuda.synchronize(device)
        forced -= 1
    # ---- device-resident timing (value) + dominant-kernel timing (roofline) ----
    barrier()
    sampler.mark()
    l0 = eng.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # Timing hygiene: the per-step working set (actions + record buffer) is far larger than the 126 MB L2 at the quoted
    # configuration; a small env-count sweep point that is not gets an explicit L2 flush (untimed) between iterations.
    global L2_NOTE
    step_bytes = 4 * (actions.numel() + rec.numel())
    flush = step_bytes < 2 * 126 * 2**20
    flush_buf = torch.empty(256 * 2**20, dtype=torch.uint8, device=device) if flush else None
    L2_NOTE = (f"per-step inputs+outputs {step_bytes / 2**20:.0f} MiB " +
               ("do not exceed the 126 MB L2: a 256 MiB buffer is written between timed iterations (outside the timed spans)" if flush
                else "exceed the 126 MB L2"))
    sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record(stream)
    for i in range(args.steps):
        if flush:
            flush_buf.fill_(i)
        sev[i][0].record(stream)
        kev[i][0].record(stream)
        eng.rollout(actions, rec)
        kev[i][1].record(stream)
        eng.score(rec, t, values=values, gamma=GAMMA, lam=LAM)
        collective()
        sev[i][1].record(stream)
    ev[1].record(stream)
    barrier()
    launches = eng.launches - l0  # this repo's kernels only (the two record hand-over copies per step are torch's)
    ms = (sum(a.elapsed_time(b) for a, b in sev) if flush else ev[0].elapsed_time(ev[1])) / args.steps
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    ms = D.max_over_ranks(ms, device)
    # ---- end-to-end through the public API with HOST buffers (pinned actions in, rewards+advantages out) ----
    host_actions = torch.empty((t, n, na), dtype=torch.float32).pin_memory()
    host_actions.copy_(actions.cpu())
    host_adv = torch.empty((n, t), dtype=torch.float32).pin_memory()
    host_total = torch.empty((n, t), dtype=torch.float32).pin_memory()
    dev_actions = torch.empty_like(actions)

    def e2e_step() -> None:
        dev_actions.copy_(host_actions, non_blocking=True)
        eng.rollout(dev_actions, rec)
        scored = eng.score(rec, t, values=values, gamma=GAMMA, lam=LAM)
        collective()
        host_adv.copy_(scored["advantages"], non_blocking=True)
        host_total.copy_(scored["total"], non_blocking=True)
        torch.cuda.synchronize(device)  # the caller reads the result

    e2e_step()
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record(stream)
    for _ in range(args.steps):
        e2e_step()
    ev2[1].record(stream)
    barrier()
    e2e_ms = D.max_over_ranks(ev2[0].elapsed_time(ev2[1]) / args.steps, device)
    extras = {} if args.no_extras else _extras(eng, actions, rec, device, rank, world, args)
    clocks = sampler.stop() if rank == 0 else {}
    nonfinite = float(eng.state[:, eng.program["TI_S_NONFINITE"]].sum())
    if nonfinite:
        raise SystemExit(f"{nonfinite} environments produced non-finite states")
    if rank != 0:
        D.shutdown()
        return
    env_steps = n * t * world
    value = env_steps / (ms * 1e-3)
    # algorithmic HBM bytes per env-step of the rollout kernel: state+keys read & written once per launch
    # (amortised over the 64 steps), action read, full record written
    p = eng.program
    per_launch_env = 4 * (2 * p.state_size + 2 * p.key_size + p.rand_size)
    bytes_per_env_step = 4 * (p.action_size + p.rec_size) + per_launch_env / t
    peak, peak_src = _peaks()
    achieved = n * t * bytes_per_env_step / (kernel_ms * 1e-3) / 1e9
    # measured DRAM bytes of one launch of this configuration (one ncu capture per round, profiles/r02_traffic.json)
    traffic = None
    tpath = ROOT / "profiles" / "r02_traffic.json"
    if tpath.exists():
        rec_t = json.loads(tpath.read_text()).get(WORKLOAD)
        if rec_t and rec_t["envs"] == n and rec_t["rollout"] == t:
            traffic = rec_t["dram_bytes_read"] + rec_t["dram_bytes_write"]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(world), "l2_note": L2_NOTE,
        "e2e": {"value": env_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_actions.numel() * 4),
                "d2h_bytes_per_step": int((host_adv.numel() + host_total.numel()) * 4), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/r02_traffic.json (one ncu capture of this configuration, committed; not measured in this run)" if traffic else None,
                     "algorithmic_bytes_per_launch": n * t * bytes_per_env_step,
                     "kernel": "k_rollout<DimsKbot>" if WORKLOAD == "kbot" else "k_rollout<DimsHumanoid>", "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms,
                     "algorithmic_bytes_per_env_step": bytes_per_env_step, "peak_source": peak_src,
                     "note": "latency/issue-bound by design (tiny per-env matrices); see DESIGN.md and profiles/"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_leg()
    out.update(extras)
    if world == 1 and WORKLOAD == "humanoid" and not args.no_extras:
        out["mjx_baseline"] = _mjx_baseline()
        del eng
        torch.cuda.empty_cache()
        out["secondary"] = _secondary_kbot(args)
        fast = _secondary_kbot(args, "fast")
        out["secondary_fast_solver"] = {k: fast[k] for k in ("value", "unit", "ms_per_step", "error") if k in fast} | {
            "what": "same K-Bot workload with EngineConfig.solver='fast' (opt-in; departs from the reference's capped iterate, DESIGN.md 2)"}
    _emit(json.dumps(out))
    D.shutdown()


if __name__ == "__main__":
    main()
