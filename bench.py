"""Benchmark of the PPO hot path: env transitions / s for ``collect()`` + ``step()``.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): CartPole, ``num_envs=65536`` PER GPU, ``horizon=32``,
Categorical policy, default ``AlgorithmConfig`` update (4 SGD epochs, full batch, ``enable_amp=False``:
fp32 results, computed on tcgen05 through split operands (two fp16 pieces per fp32 value in the update, three bf16
pieces in the rollout) -- the mode that reproduces the reference-recorded golden vectors at 1e-5).  ``amp_mode`` in the same line is ``enable_amp=True`` (plain
bf16 operands); with more than one GPU ``c5`` is BASELINE.json configs[4] exactly.  A "step"
is one ``collect()`` + one ``step()``.  ``value`` is timed with CUDA events around K steps
(max over ranks); ``e2e`` repeats it through ``Trainer.step()`` with the sampling noise
supplied from pinned host memory every step (H2D inside the timed region) and the stats read
back (D2H).  Working set per step (buffer 112 MB + activations) exceeds nothing by itself,
so an L2 flush (write of a 256 MB buffer) runs between timed iterations.

``--impl reference`` times the UNMODIFIED reference (``oracle/_ref``, staged by ``tools/stage_ref.py`` from the
upstream tree, run behind ``oracle/refshim`` with ``device="cpu"`` on all host cores) on the same workload; only
when ``oracle/_ref`` is absent does it fall back to the CPU restatement ``oracle/ppo_oracle.py``
(``cpu_baseline.kind`` says which).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env transitions/sec (collect+PPO step)"
UNIT = "transitions/s"


#: workload -> (description, env, distribution, recurrent, num_envs per GPU, horizon, obs dim D, head width P)
WORKLOADS = {
    "cartpole": ("CartPole feedforward PPO (BASELINE.json configs[1])", "CartPole", "Categorical", False,
                 65536, 32, 5, 3),
    "dummy": ("DiscreteDummyEnv feedforward PPO (BASELINE.json configs[0])", "DiscreteDummyEnv", "Categorical",
              False, 8192, 32, 1, 2),
    "pendulum": ("Pendulum feedforward PPO, SquashedNormal (BASELINE.json configs[2])", "Pendulum",
                 "SquashedNormal", False, 262144, 64, 3, 2),
    "cartpole_lstm": ("CartPole RecurrentAlgorithm, LSTM policy (BASELINE.json configs[3])", "CartPole",
                      "Categorical", True, 65536, 32, 5, 3),
}


def parse() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cartpole", choices=sorted(WORKLOADS),
                    help="cartpole = BASELINE.json configs[1] (the bench line); the others are extra"
                         " measurements of configs[0], [2] and [3]")
    ap.add_argument("--num-envs-per-gpu", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--horizon", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--sgd-iters", type=int, default=4)
    ap.add_argument("--minibatch", type=int, default=0, help="0 = whole buffer (reference default)")
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    a.num_envs_per_gpu = a.num_envs_per_gpu or w[4]
    a.horizon = a.horizon or w[5]
    return a


def workload_config(a: argparse.Namespace, world: int) -> dict:
    w = WORKLOADS[a.workload]
    return {
        "workload": w[0],
        "env": w[1],
        "num_envs_per_gpu": a.num_envs_per_gpu,
        "num_envs_total": a.num_envs_per_gpu * world,
        "horizon": a.horizon,
        "distribution": w[2],
        "num_sgd_iters": a.sgd_iters,
        "sgd_minibatch_size": a.minibatch or a.num_envs_per_gpu * a.horizon,
        "parallelism": f"env-sharded dp{world}",
        "l2": "flushed between timed iterations (256 MB write)",
    }


# ---------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference)
# ---------------------------------------------------------------------------------------


_ORACLE_ENV = {"CartPole": "cartpole", "DiscreteDummyEnv": "discrete_dummy", "Pendulum": "pendulum"}
_ORACLE_DIST = {"Categorical": "categorical", "Normal": "normal", "SquashedNormal": "squashed_normal"}


def cpu_step_fn(workload: str, N: int, T: int, sgd_iters: int, minibatch: int):
    import torch

    from oracle import ppo_oracle as O
    from oracle import recurrent_oracle as R

    _, env_name, dist_name, recurrent, _, _, D, P = WORKLOADS[workload]
    torch.manual_seed(0)
    env = O.OracleEnv(_ORACLE_ENV[env_name], N)
    dist = O.Dist(_ORACLE_DIST[dist_name])
    discrete = dist_name == "Categorical"
    kind, n_act = ("discrete", P) if discrete else ("continuous", 1)
    opt: dict = {}
    noise_shape = (T, N, 1, P) if discrete else (T, N, 1)

    def draw() -> "torch.Tensor":
        z = torch.empty(noise_shape)
        return z.exponential_(1) if discrete else z.normal_()

    if recurrent:
        params = R.init_recurrent_params(D, kind, n_act)
        buf = R.new_recurrent_buffer(N, T, D, kind)
        seqs = [0]

        def step() -> None:
            stats, seqs[0] = R.collect_recurrent(params, env, buf, dist, draw(), seqs=seqs[0])
            R.step_recurrent(params, buf, dist, opt, reward_scale=stats["reward_scale"],
                             num_sgd_iters=sgd_iters, sgd_minibatch_size=minibatch or None,
                             shuffle=bool(minibatch))
            for v in buf.values():
                v.zero_()

        return step

    params = O.init_params(D, kind, n_act)
    buf = O.new_buffer(N, T, D, kind)

    def step() -> None:
        stats = O.collect(params, env, buf, dist, draw())
        O.step(params, buf, dist, opt, reward_scale=stats["reward_scale"], num_sgd_iters=sgd_iters,
               sgd_minibatch_size=minibatch or None, shuffle=bool(minibatch))

    return step


REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_available(workload: str) -> bool:
    """The unmodified upstream source staged by tools/stage_ref.py (git-ignored oracle/_ref/)."""
    return os.path.isdir(os.path.join(REF_DIR, "src", "rl8"))


def reference_step_fn(workload: str, N: int, T: int, sgd_iters: int, minibatch: int, device: str = "cpu"):
    """``collect()`` + ``step()`` of the UNMODIFIED reference (its own AlgorithmConfig / Algorithm classes, eager)
    behind the arithmetic-free stand-ins of oracle/refshim.

    The reference picks ``"cuda"`` whenever ``torch.cuda.is_available()`` -- whatever ``config.device`` says
    (src/rl8/algorithms/_feedforward.py:210-216: the conditional binds as ``"cuda" if available else (...)``) -- so
    for its CPU path CUDA is hidden from it while it builds (``--impl reference`` also hides the GPUs from the whole
    process before torch is imported)."""
    os.environ["TORCHDYNAMO_DISABLE"] = "1"  # Inductor's C++ backend does not build in this image
    for p_ in (REF_DIR, os.path.join(REF_DIR, "src"), os.path.join(ROOT, "oracle", "refshim")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    import torch
    from rl8 import AlgorithmConfig as RefConfig  # noqa: E402  (upstream)
    from rl8 import RecurrentAlgorithmConfig as RefRecurrentConfig  # noqa: E402
    from rl8.distributions import SquashedNormal as RefSquashed  # noqa: E402
    from rl8.env import DiscreteDummyEnv as RefDummy  # noqa: E402

    _, env_name, dist_name, recurrent, _, _, _, _ = WORKLOADS[workload]
    if env_name == "CartPole":
        from examples.cartpole.env import CartPole as env_cls  # noqa: E402, N813
    elif env_name == "Pendulum":
        from examples.pendulum.env import Pendulum as env_cls  # noqa: E402, N813
    else:
        env_cls = RefDummy
    torch.manual_seed(0)
    kw = dict(num_envs=N, horizon=T, num_sgd_iters=sgd_iters, sgd_minibatch_size=minibatch or None, device=device)
    if dist_name == "SquashedNormal":
        kw["distribution_cls"] = RefSquashed
    seen = torch.cuda.is_available
    if device == "cpu":
        torch.cuda.is_available = lambda: False
    try:
        algo = (RefRecurrentConfig if recurrent else RefConfig)(**kw).build(env_cls)
    finally:
        torch.cuda.is_available = seen
    assert str(algo.hparams.device).startswith(device), algo.hparams.device

    def step() -> None:
        algo.collect()
        algo.step()

    return step


def cpu_leg(workload: str, N: int, T: int, sgd_iters: int, minibatch: int):
    """(step function, kind): the staged reference when present, else the CPU port of it."""
    if reference_available(workload):
        try:
            return reference_step_fn(workload, N, T, sgd_iters, minibatch), "reference"
        except Exception as e:  # noqa: BLE001  (a broken staging must not take the bench down)
            print(f"reference import failed ({type(e).__name__}: {e}); timing the CPU port", file=sys.stderr)
    return cpu_step_fn(workload, N, T, sgd_iters, minibatch), "port"


def cpu_probe(a: argparse.Namespace, budget_s: float, iters: int) -> tuple[int, float]:
    """Pick the largest power-of-two num_envs (<= the workload's) whose `iters` steps fit the
    budget; returns (num_envs, seconds per transition estimate)."""
    N0 = 1024
    fn, _ = cpu_leg(a.workload, N0, a.horizon, a.sgd_iters, 0)
    fn()
    t0 = time.perf_counter()
    fn()
    per_tr = (time.perf_counter() - t0) / (N0 * a.horizon)
    N = a.num_envs_per_gpu
    while N > 1024 and per_tr * N * a.horizon * iters > budget_s:
        N //= 2
    return N, per_tr


def run_reference(a: argparse.Namespace) -> None:
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # the full workload when warm-up + timed steps fit ~10 minutes of host time, else the largest power-of-two
    # share of its environments that does; the line's config carries the size that actually ran
    N, _ = cpu_probe(a, 600.0, a.steps + a.warmup)
    fn, kind = cpu_leg(a.workload, N, a.horizon, a.sgd_iters, 0)
    for _ in range(a.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    value = N * a.horizon * a.steps / dt
    what = ("the unmodified reference (oracle/_ref, staged by tools/stage_ref.py) behind oracle/refshim, device=cpu, eager"
            if kind == "reference" else "oracle/ppo_oracle.py (CPU port of the reference, torch fp32 eager)")
    sample = (f"{WORKLOADS[a.workload][1]} num_envs={N} (of {a.num_envs_per_gpu}), horizon={a.horizon}, "
              f"{a.warmup} warm-up + {a.steps} timed collect+step of {what}")
    cfg = workload_config(a, 1)  # same keys as the GPU arm; the environment count is the one that actually ran
    cfg["num_envs_per_gpu"] = cfg["num_envs_total"] = N
    cfg["sgd_minibatch_size"] = a.minibatch or N * a.horizon
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": a.gpus,
        "steps": a.steps,
        "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": what,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self) -> "ClockSampler":
        self._t.start()
        return self

    def __exit__(self, *exc) -> None:  # noqa: ANN002
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        p["source"] = "measured (MEASURED_PEAKS.json)"
        return p
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md)"}


def run_ours(a: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist

    import rl8_b200.distributions as dists
    import rl8_b200.env as envs
    from rl8_b200 import AlgorithmConfig, RecurrentAlgorithmConfig, RecurrentTrainer, Trainer, _lib

    _, env_name, dist_name, recurrent, _, _, D, P = WORKLOADS[a.workload]
    env_cls, base_dist = getattr(envs, env_name), getattr(dists, dist_name)
    config_cls = RecurrentAlgorithmConfig if recurrent else AlgorithmConfig
    trainer_cls = RecurrentTrainer if recurrent else Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    N, T = a.num_envs_per_gpu, a.horizon
    lib = _lib.load()

    def make(dist_cls=None, precision=a.precision):  # noqa: ANN001, ANN202
        # auto = the reference's default enable_amp=False: fp32 results (on tcgen05 through split operands for
        # the feedforward models), the mode the golden vectors are reproduced in; bf16 = enable_amp=True
        amp = precision == "bf16"
        dist_cls = dist_cls or base_dist
        return config_cls(
            num_envs=N, horizon=T, num_sgd_iters=a.sgd_iters, sgd_minibatch_size=a.minibatch or None,
            enable_amp=amp, distribution_cls=dist_cls,
        ).build(env_cls), ("bf16" if amp else "f32")

    # Identical initial weights on every rank (same seed before build(); Algorithm.__init__ also broadcasts rank 0's
    # parameters), then per-rank streams for the env resets and the sampling noise.
    torch.manual_seed(0)
    algo, dtype = make()
    torch.manual_seed(1000 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_step: list[float] = []  # ms of every timed step of the last timed() call (this rank)

    def timed(step_fn, steps: int, warmup: int) -> tuple[float, int]:  # noqa: ANN001
        # one-time work (workspace allocation, kernel attribute set-up, the CUDA-graph capture of the update on the
        # second eligible call) must not land in the timed steps however small the caller's --warmup is
        for _ in range(max(warmup, 3)):
            step_fn()
        barrier()
        total_ms, launches = 0.0, 0
        per_step.clear()
        for _ in range(steps):
            flush.zero_()  # L2 flush, outside the timed bracket
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launches += step_fn()
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            per_step.append(round(e0.elapsed_time(e1), 3))
        barrier()
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    def core_step() -> int:
        algo.collect()
        algo.step()
        return algo.last_launches["collect"] + algo.last_launches["step"]

    # clocks / throttle reasons are sampled while BOTH timed legs run (one nvidia-smi query takes ~0.2 s)
    clocks = ClockSampler(local)
    clocks.__enter__()
    ms, launches = timed(core_step, a.steps, a.warmup)
    value = N * T * world * a.steps / (ms / 1e3)
    core_per_step = list(per_step)

    # ---- e2e: Trainer.step() with host-supplied noise (pinned, H2D every step) + stats D2H --------
    discrete = dist_name == "Categorical"
    width = P if discrete else 1
    host_noise = torch.empty(T, N, width)
    host_noise = (host_noise.exponential_(1) if discrete else host_noise.normal_()).pin_memory()
    # The input pipeline a host-fed deployment would use: two device buffers, the copy of the NEXT step's
    # noise is enqueued on a side stream as soon as the current step has taken its buffer (one 25 MB H2D
    # per step, inside the timed region, overlapped with the kernels of the step before).
    dev_noise = [torch.empty(T, N, width, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    turn = [0]

    def enqueue_copy(i: int) -> None:
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i])  # the rollout that read this buffer has finished
            dev_noise[i].copy_(host_noise, non_blocking=True)
            ready[i].record(copy_stream)

    for i in range(2):
        consumed[i].record(torch.cuda.current_stream())
    enqueue_copy(0)

    class HostNoise(base_dist):  # type: ignore[misc, valid-type]
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            if steps != T:  # build() -> validate()
                return super().draw_noise(steps, num, width, device)
            i = turn[0]
            turn[0] = i ^ 1
            torch.cuda.current_stream().wait_event(ready[i])
            enqueue_copy(i ^ 1)
            return dev_noise[i]

    torch.manual_seed(0)
    algo2, _ = make(HostNoise, "bf16" if dtype == "bf16" else "fp32")
    torch.manual_seed(2000 + rank)
    trainer = trainer_cls(algo2)

    def e2e_step() -> int:
        stats = trainer.step()
        consumed[turn[0] ^ 1].record(torch.cuda.current_stream())  # this step's noise buffer is free again
        assert stats["losses/total"] == stats["losses/total"]
        return algo2.last_launches["collect"] + algo2.last_launches["step"]

    ms2, _ = timed(e2e_step, a.steps, 3)
    clocks.__exit__(None, None, None)
    e2e_value = N * T * world * a.steps / (ms2 / 1e3)
    e2e_per_step = list(per_step)
    h2d = host_noise.numel() * 4
    d2h = 16 * 8 + algo2._loss_sums.numel() * 8

    # ---- roofline of the update: every rank runs it (its collect() holds collectives) -------------
    peaks = measured_peaks()
    upd = update_roofline(a, algo, _lib, peaks, dtype, flush)
    barrier()

    # ---- beside the headline (enable_amp=False, parity with the reference at 1e-5): the reference's AMP switch,
    #      enable_amp=True = plain bf16 operands on tcgen05 (checked at bf16 tolerances, tests/test_gpu_tc.py) -------
    amp_mode = None
    if dtype == "f32" and a.precision == "auto":
        torch.manual_seed(0)
        algo16, _ = make(None, "bf16")
        torch.manual_seed(3000 + rank)

        def amp_step() -> int:
            algo16.collect()
            algo16.step()
            return algo16.last_launches["collect"] + algo16.last_launches["step"]

        k16 = max(1, min(a.steps, 10))
        ms16, launches16 = timed(amp_step, k16, 3)
        amp_mode = {
            "value": N * T * world * k16 / (ms16 / 1e3), "unit": UNIT, "ms_per_step": ms16 / k16, "steps": k16,
            "warmup": 3, "dtype": "bf16 (enable_amp=True: bf16 operands on tcgen05, fp32 accumulate)",
            "gpu_launches": launches16, "ms_each_step": list(per_step),
            "parity": "bf16 tolerances against a bf16-emulating oracle (tests/test_gpu_tc.py); NOT the 1e-5 bar",
            "roofline": update_roofline(a, algo16, _lib, peaks, "bf16", flush),
        }
        del algo16
        barrier()

    # ---- BASELINE.json configs[4] exactly: CartPole, 1 048 576 envs in total, horizon 32, one PPO update per collect
    c5 = None
    if world > 1 and a.workload == "cartpole" and a.precision == "auto" and (1 << 20) % world == 0:
        n5 = (1 << 20) // world
        c5 = {"config": {"workload": "CartPole env-sharded scaling sweep (BASELINE.json configs[4])",
                         "num_envs_total": 1 << 20, "num_envs_per_gpu": n5, "horizon": 32, "num_sgd_iters": 1,
                         "sgd_minibatch_size": n5 * 32, "parallelism": f"env-sharded dp{world}"}}
        for key, amp in (("f32", False), ("bf16", True)):
            torch.manual_seed(0)
            algo5 = config_cls(num_envs=n5, horizon=32, num_sgd_iters=1, enable_amp=amp).build(env_cls)
            torch.manual_seed(4000 + rank)

            def c5_step() -> int:
                algo5.collect()
                algo5.step()
                return algo5.last_launches["collect"] + algo5.last_launches["step"]

            k5 = max(1, min(a.steps, 10))
            ms5, _ = timed(c5_step, k5, 3)
            c5[key] = {"value": (1 << 20) * 32 * k5 / (ms5 / 1e3), "unit": UNIT, "ms_per_step": ms5 / k5, "steps": k5,
                       "warmup": 3}
            del algo5
            barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines of the stream kernels (rank 0; no collectives) ----------------------------------
    roofs = stream_rooflines(lib, _lib, dev, peaks)

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": a.steps,
        "warmup": a.warmup,
        "ms_per_step": ms / a.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": dtype,
        "data": "synthetic",
        "config": workload_config(a, world),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "ms_each_step": {"value": core_per_step, "e2e": e2e_per_step},
        "roofline": upd,
        "amp_mode": amp_mode,
        "c5": c5,
        "stream_rooflines": roofs,
        "peaks": {k: peaks.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained", "source")},
    }
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(a)
        line["reference_on_gpu"] = reference_on_gpu(a)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a: argparse.Namespace) -> dict:
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    n, _ = cpu_probe(a, a.cpu_budget_s, 2)
    fn, kind = cpu_leg(a.workload, n, a.horizon, a.sgd_iters, 0)
    fn()
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    what = ("the unmodified reference (oracle/_ref) behind oracle/refshim, device=cpu, eager" if kind == "reference"
            else "oracle/ppo_oracle.py (torch fp32 eager)")
    return {
        "value": n * a.horizon / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
        "sample": f"{WORKLOADS[a.workload][1]} num_envs={n} (of {a.num_envs_per_gpu}), horizon={a.horizon}, 1 warm-up + 1 timed"
                  f" collect+step of {what}",
    }


def reference_on_gpu(a: argparse.Namespace) -> None | dict:
    """Context, not the baseline: the unmodified reference on ITS OWN CUDA path (eager PyTorch kernels, cuBLAS) on
    the same B200, same workload -- what switching the hot path to this library buys a user of the reference."""
    import torch

    if not reference_available(a.workload):
        return None
    try:
        fn = reference_step_fn(a.workload, a.num_envs_per_gpu, a.horizon, a.sgd_iters, a.minibatch, device="cuda")
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return {"value": a.num_envs_per_gpu * a.horizon / dt, "unit": UNIT, "ms_per_step": 1e3 * dt,
            "note": "unmodified reference (oracle/_ref), device=cuda, eager, fp32; 1 warm-up + 3 timed collect+step,"
                    " wall clock with synchronize"}


def _time_kernel(fn, flush, iters: int = 10) -> float:  # noqa: ANN001
    """Average ms of `fn` with an L2 flush before every timed launch."""
    import torch

    for _ in range(3):
        fn()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters


def stream_rooflines(lib, L, dev, peaks: dict) -> list[dict]:  # noqa: ANN001
    """Achieved GB/s of the HBM-bound kernels (algorithmic bytes of SURVEY.md §8d / DESIGN.md)
    at sizes well beyond L2: GAE at N=2^20, T=32 and the CartPole env step at N=2^24."""
    import torch

    from rl8_b200.env import CartPole

    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm = float(peaks["hbm_gbs"])
    # GAE scan (horizon-major): 8 B read + 12 B written per transition, + 12 B per env
    N, T = 1 << 20, 32
    r = torch.randn(T + 1, N, device=dev)
    v = torch.randn(T + 1, N, device=dev)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    mom = torch.zeros(3, dtype=torch.float64, device=dev)
    st = L.stream()
    ms = _time_kernel(lambda: lib.rl8_gae_scan(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N,
                                               0.95, 0.95, 1.0, L.ptr(mom), st), flush)
    alg = 20.0 * N * T + 20.0 * N
    out.append({"kernel": "gae_scan_hm_kernel (functional form: scaled rewards written back, 20 B per transition)",
                "bound": "hbm", "bytes": alg, "ms": ms,
                "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / hbm,
                "shape": f"N={N} T={T}"})
    # the form Algorithm.step() launches: divisor read from device memory, no scaled-reward write-back (the buffer is
    # dropped right after): SURVEY.md §8d's 16 B per transition + 8 B per env (+ 8 B per env for the slot-T rewards read)
    scale_dev = torch.ones(2, device=dev)
    ms = _time_kernel(lambda: lib.rl8_gae_scan_dev(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N, 0.95, 0.95,
                                                   L.ptr(scale_dev), 0, L.ptr(mom), st), flush)
    alg = 16.0 * N * T + 16.0 * N
    out.append({"kernel": "gae_scan_hm_kernel (Algorithm.step form: no reward write-back)", "bound": "hbm", "bytes": alg,
                "ms": ms, "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / hbm,
                "shape": f"N={N} T={T}"})
    ms = _time_kernel(lambda: lib.rl8_gae_normalize(L.ptr(adv), N, T, 1, N, L.ptr(mom), st), flush)
    alg = 8.0 * N * T
    out.append({"kernel": "gae_normalize_kernel", "bound": "hbm", "bytes": alg, "ms": ms,
                "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / hbm,
                "shape": f"N={N} T={T}"})
    del r, v, adv, ret
    # CartPole env step: 64 B per env-step (16 R state + 8 R action + 16 W state + 20 W obs + 4 W reward)
    N = 1 << 24
    env = CartPole(N, 32, device=dev)
    env.reset()
    act = torch.randint(0, 3, (N,), device=dev)
    cfg = env.rl8_cfg()
    ms = _time_kernel(lambda: lib.rl8_env_step(env.rl8_kind, cfg, L.ptr(env.state), L.ptr(act), L.ptr(env._obs),
                                               1, N, L.ptr(env._reward), N, st), flush)
    alg = 64.0 * N
    out.append({"kernel": "env_step_kernel<cartpole>", "bound": "hbm", "bytes": alg, "ms": ms,
                "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / hbm,
                "shape": f"N={N}"})
    del env, act
    # Padded rolling windows (shift 3) over a horizon-major obs field: item read once, windows + mask written once
    N, T1, D, size = 1 << 19, 33, 5, 4
    x = torch.randn(T1, D, N, device=dev).permute(2, 0, 1)
    o = torch.empty(N * T1, size, D, device=dev)
    mk = torch.empty(N * T1, size, dtype=torch.bool, device=dev)
    ms = _time_kernel(lambda: lib.rl8_view_windows(L.ptr(x), 4, N, T1, D, x.stride(0), x.stride(1), x.stride(2),
                                                   size, -(size - 1), T1, L.ptr(o), L.ptr(mk), st), flush)
    alg = 4.0 * N * T1 * D * (1 + size) + 1.0 * N * T1 * size
    out.append({"kernel": "view_windows_kernel<u32> (PaddedRollingWindow.apply_all, shift 3)", "bound": "hbm",
                "bytes": alg, "ms": ms, "achieved": alg / ms / 1e6, "peak": hbm, "unit": "GB/s",
                "frac": alg / ms / 1e6 / hbm, "shape": f"N={N} T+1={T1} D={D} size={size}"})
    return out


#: dram__bytes_read.sum + dram__bytes_write.sum of the four kernel launches of one CartPole
#: N=65536 T=32 full-batch rl8_ppo_minibatch (tc_update_h + tc_update_w over 2^21 rows), from the
#: `ncu --set full` captures summarised in profiles/r01_update_v8_ncu_summary.md.
NCU_TRAFFIC = {("cartpole", 2097152): (84.31e6 + 2089.0e6) + (2190.0e6 + 7.0e6)}
#: the same call in RL8_PREC_FP32_TC (x3_update_f + x3_update_b + x3_update_w), dram read + write of each kernel from
#: profiles/r02_x3_fp16_ncu_summary.md: no dZ2 scratch, 80 B per row and network of masks / dOut instead
NCU_TRAFFIC_X3 = {("cartpole", 2097152): (84.8e6 + 280.3e6) + (378.2e6 + 5.1e6) + (244.8e6 + 4.7e6)}


def update_roofline(a, algo, L, peaks: dict, dtype: str, flush) -> dict:  # noqa: ANN001
    """The update (forward + loss + backward over one minibatch) dominates a step: its GEMM
    FLOPs are 3 x the forward's (forward, input-gradient and weight-gradient contractions)."""
    import torch

    _, _, _, recurrent, _, _, D, P = WORKLOADS[a.workload]
    hp = algo.hparams
    N, T = hp.num_envs, hp.horizon
    peak = float(peaks["bf16_tflops_sustained"])  # recurrent: a whole multi-kernel step() is timed
    H = 256
    if recurrent:
        # one full step() = GAE + num_sgd_iters passes over every sequence (fp32 CUDA-core path)
        def fn() -> None:
            algo.collect()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            e0.record()
            algo.step()
            e1.record()
            e1.synchronize()
            fn.ms += e0.elapsed_time(e1)

        fn.ms = 0.0
        fn()
        fn.ms = 0.0
        for _ in range(3):
            fn()
        ms = fn.ms / 3
        flops = hp.num_sgd_iters * N * T * 3.0 * 2.0 * ((D + H) * 4 * H + (P + 1) * H)
        return {
            "kernel": "RecurrentAlgorithm.step (rl8_lstm_ppo_minibatch x num_sgd_iters; LSTM GEMMs: "
                      + ("tcgen05 bf16: tc_lstm_cell / lstm_cell_bwd_tc / lstm_dh_tc / lstm_wgrad_tc kernels)"
                         if dtype == "bf16" else "fp32 CUDA cores)"),
            "bound": "tensor", "achieved": flops / ms / 1e9, "peak": peak, "unit": "TFLOP/s",
            "frac": flops / ms / 1e9 / peak, "traffic": None, "ms": ms, "rows": N * T, "dtype": dtype,
            "peak_source": peaks["source"] + " bf16 sustained (whole step() incl. its elementwise kernels)",
        }
    algo.collect()
    model = algo.policy.model
    m, g = model.struct_for(model.flat_params), model.struct_for(algo._grads)
    M = hp.sgd_minibatch_size
    prec = algo.policy.precision
    ws = algo._workspace("ppo", int(algo._lib.rl8_ppo_workspace(m, M, prec)))
    batch = algo._batch_struct()
    ppo = L.PpoHparams(0.2, 0.0, 0.0, 5.0, 1.0, 1.0)
    sums = torch.zeros(5, dtype=torch.float64, device=algo.device)

    def fn() -> None:
        rc = algo._lib.rl8_ppo_minibatch(m, g, batch, None, 0, M, float(M), ppo, L.ptr(sums), prec, L.ptr(ws),
                                         ws.numel(), L.stream())
        assert rc == 0, rc

    ms = _time_kernel(fn, flush, iters=5)
    flops = M * 3.0 * 2.0 * (2 * H * H + 2 * D * H + (P + 1) * H)
    burst = float(peaks["bf16_tflops"])  # the call is timed alone (L2 flushed before it): the burst figure applies
    if prec == L.PREC_FP32_TC:
        # every fp32 product of the 256 x 256 contractions is three fp16 piece products on the tensor pipe (kind::f16
        # multiplies fp16 operands at the bf16 rate the peak was measured at)
        return {
            "kernel": "rl8_ppo_minibatch, RL8_PREC_FP32_TC (x3_update_f + x3_update_b + x3_update_w per 2^21-row chunk:"
                      " pair MMAs on two fp16 pieces per fp32 operand, operands recomputed from 80 B/row of scratch)",
            "bound": "tensor", "achieved": flops / ms / 1e9, "peak": burst / 3.0, "unit": "TFLOP/s",
            "frac": flops / ms / 1e9 / (burst / 3.0), "traffic": NCU_TRAFFIC_X3.get((a.workload, M)), "ms": ms,
            "rows": M, "dtype": "f32",
            # piece products actually issued per row: policy network 3 + 3 + 3, value network 3 + 2 + 2 (its
            # gradient contractions take the ReLU mask as a one-piece operand)
            "mma_tflops": M * 16.0 * 2.0 * H * H / ms / 1e9, "mma_frac": M * 16.0 * 2.0 * H * H / ms / 1e9 / burst,
            "traffic_source": ("ncu --set full, profiles/r02_x3_fp16_ncu_summary.md"
                               if (a.workload, M) in NCU_TRAFFIC_X3 else None),
            "peak_source": peaks["source"] + " bf16 burst / 3 (three fp16 piece products per fp32 product at the bf16"
                                             " rate; `achieved` counts the algorithm's fp32 FLOPs; `mma_tflops` /"
                                             " `mma_frac` count the piece products actually issued -- 16 per row, not"
                                             " 18: two of the value network's contractions need two -- against the"
                                             " plain burst peak).  The kernels are bound by the CUDA-core operand"
                                             " producers and epilogues, not by the tensor pipe (profiles/)",
        }
    traffic = NCU_TRAFFIC.get((a.workload, M)) if dtype == "bf16" else None
    return {
        "kernel": "rl8_ppo_minibatch (forward + losses + backward, one minibatch: tc_update_h_kernel +"
                  " tc_update_w_kernel per 2^21-row chunk)",
        "bound": "tensor", "achieved": flops / ms / 1e9, "peak": burst, "unit": "TFLOP/s",
        "frac": flops / ms / 1e9 / burst, "traffic": traffic, "ms": ms, "rows": M, "dtype": dtype,
        "peak_source": peaks["source"] + " bf16 burst (the call is timed alone, L2 flushed before each launch)",
        "traffic_source": "ncu --set full, profiles/r01_update_v8_ncu_summary.md" if traffic else None,
    }


def main() -> None:
    a = parse()
    if a.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the reference's CPU path (see reference_step_fn); before torch loads
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
