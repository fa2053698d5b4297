#!/usr/bin/env python
"""bench.py -- env steps/sec of the hot path (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]

Workloads (BASELINE.json `configs`; `--config`, default c2 = the configuration the metric is quoted on):

  c2  cologne8 (8 signals) / MaxPressure / 4096 lock-step instances per GPU; states.mplight + rewards.wait
  c3  ingolstadt21 (21 signals) / IDQN / 8192 instances: the kernel emits states.drq_norm + rewards.wait_norm and
      rewards.pressure every step (agent_config.py:83-94 and BASELINE's wording); IDQN's own network is pfrl code on the
      caller's side, so actions are its epsilon = 1 exploration: uniform random green phases drawn on the device
  c4  ingolstadt21 / MPLight shared controller / 8192 instances per GPU (65536 on 8): every rank evaluates FRAP
      (agents/mplight.py, random-init weights: no checkpoints here; the shared weights are replicated) for its own
      instances, and states.mplight + the pressure reward of every rank are all-gathered over NCCL each step so that
      the learner rank holds the full batch; two half-batches per rank on two streams, so the collectives of one half
      overlap the env-step kernel of the other.  --c4-policy rank0: the literal single-rank policy (gather, FRAP on
      rank 0 for all instances, broadcast of the actions)
  c5  synthetic 4x4 grid, Poisson (Bernoulli per tick) demand swept over 300 / 600 / 900 / 1200 veh/h per entry lane,
      16384 instances per GPU, MaxPressure; `value` is the sweep total, per-rate lines in config.sweep

One "step" = one MultiSignal.step() of every instance = batched policy + step_length one-second simulation ticks +
Signal.observe + states + rewards (multi_signal.py:164-197), issued as ONE CUDA-graph launch (rs_env_step_policy).
Per-instance driver randomness (speedFactor, dawdling) is Philox keyed by the global instance id; "data": "synthetic".

`value`  : device-timed (CUDA events on the launching stream), actions / observations resident in HBM.
`e2e`    : the same metric through the host-buffer C-ABI call (rs_env_step_host[_async]): pinned H2D of the actions +
           D2H of the observation and reward inside the timed region, host-side agent.
`--impl reference`: the CPU arm.  The reference's own CPU path (MultiSignal over libsumo) cannot run anywhere in this
           environment (SUMO is not installed), so this arm times the CPU oracle port of the same
           algorithm on all host cores (cpu_baseline.kind = "port").
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c2": dict(map="cologne8", n_env=4096, policy="maxpressure", tile=128, vcap=1024, rates=[0.0], preroll=90,
               host_obs="mplight", reward_kind=0, outputs=(),
               what="cologne8 (8 signals) / MaxPressure / {n} lock-step instances per GPU"),
    "c3": dict(map="ingolstadt21", n_env=8192, policy="random", tile=0, vcap=0, rates=[0.0], preroll=60,
               host_obs="drq_norm", reward_kind=1, outputs=("drq_norm",),
               what="ingolstadt21 (21 signals) / IDQN (epsilon = 1: uniform random actions on the device) / {n} instances "
                    "per GPU / kernel emits drq_norm + wait_norm + pressure"),
    "c4": dict(map="ingolstadt21", n_env=8192, policy="frap", tile=0, vcap=0, rates=[0.0], preroll=60,
               host_obs="mplight", reward_kind=2, outputs=(),
               what="ingolstadt21 / MPLight shared controller (FRAP forward kernel, replicated weights; states.mplight + "
                    "pressure reward NCCL all-gathered for the learner rank every step) / {n} instances per GPU"),
    "c5": dict(map="grid4x4", n_env=16384, policy="maxpressure", tile=4096, vcap=4096, rates=[300.0, 600.0, 900.0, 1200.0],
               preroll=60, host_obs="mplight", reward_kind=0, outputs=(),
               what="synthetic 4x4 grid / Bernoulli-per-tick demand sweep 300-1200 veh/h per entry lane / MaxPressure / "
                    "{n} instances per GPU"),
}
MAP = CONFIGS["c2"]["map"]


def _marshal(map_name, vcap=0, tile=0, synthetic_rate=0.0):
    from resco_b200.abi import marshal
    from resco_b200.scenario import Scenario
    sc = Scenario.load(os.path.join(ROOT, "resco_b200", "data", map_name + ".npz"))
    mc = sc.meta["map_config"]
    synth = None
    if synthetic_rate > 0:
        from resco_b200.scenario.synth import synth_demand
        synth = synth_demand(sc, synthetic_rate)
    m = marshal(sc, step_length=mc["step_length"], yellow_length=mc["yellow_length"], max_distance=200.0, vcap=vcap,
                tile_vcap=tile, synthetic=synth)
    return sc, m


def frap_state_dict(n_pairs, seed=0):
    """Random-init FRAP parameters (no checkpoints in this environment), PyTorch default initialisers, fixed seed."""
    import torch
    from resco_b200.agents import BatchedFRAP
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = BatchedFRAP([[0, 1]] * n_pairs)
    del g
    return {k: v.detach().clone() for k, v in model.state_dict().items()}, model


def numpy_wave_agent(sc, m):
    """MAXPRESSURE in numpy (CPU-baseline workers: the oracle arm must not depend on the CUDA library)."""
    pairs = np.asarray(sc.meta["phase_pairs"], np.int64)
    va = sc.meta["valid_acts"]
    tables = []
    for s in m.info["signal_ids"]:
        idxs = list(range(len(pairs))) if va is None else [int(k) for k in va[s].keys()]
        acts = idxs if va is None else [va[s][str(k)] for k in idxs]
        tables.append((pairs[idxs, 0] + 1, pairs[idxs, 1] + 1, np.asarray(acts, np.int32)))

    def act(obs):
        mplight = obs["mplight"]
        out = np.empty(mplight.shape[:2], np.int32)
        for i, (p0, p1, acts) in enumerate(tables):
            out[:, i] = acts[np.argmax(mplight[:, i, p0] + mplight[:, i, p1], 1)]
        return out
    return act


def host_agent(cfg, sc, m, n_env, seed):
    """The caller's act() for the CPU arm and for the e2e loop: observation batch in host memory -> [n, S] actions."""
    sig = m.info["signal_ids"]
    if cfg["policy"] == "maxpressure":
        return numpy_wave_agent(sc, m)
    ng = np.asarray([len(m.info["green_states"][s]) for s in sig], np.int64)
    if cfg["policy"] == "random":
        rng = np.random.default_rng(seed)
        return lambda obs: (rng.integers(0, 1 << 30, (n_env, len(sig))) % ng[None, :]).astype(np.int32)
    import torch
    torch.set_num_threads(1)
    sd, _ = frap_state_dict(len(sc.meta["phase_pairs"]))
    from resco_b200.agents import BatchedFRAP
    model = BatchedFRAP(sc.meta["phase_pairs"])
    model.load_state_dict(sd)
    va = sc.meta["valid_acts"]
    return lambda obs: model.act(torch.from_numpy(obs["mplight"]), va, sig).numpy().astype(np.int32)


# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process = one core: `n_inst` oracle instances, `warm` untimed + `n_steps` timed env steps of the config."""
    key, n_inst, n_steps, seed, first, warm, rate = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import OracleSim
    cfg = CONFIGS[key]
    sc, m = _marshal(cfg["map"], cfg["vcap"], 0, rate)
    agent = host_agent(cfg, sc, m, n_inst, seed + first)
    sim = OracleSim(m, n_inst, seed=seed)
    sim.reset(seed, first)
    sim.observe()
    for _ in range(warm):
        sim.env_step(agent(sim.obs()))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        sim.env_step(agent(sim.obs()))
    return time.perf_counter() - t0, n_inst * n_steps


def cpu_baseline(key="c2", n_steps=120, n_inst=None, cores=None, warm=None):
    """Oracle port on the host cores, bounded sample (about 10-30 s of CPU work): every core steps its own `n_inst`
    instances `n_steps` env steps after `warm` untimed ones (the GPU arm's pre-roll: both arms time the loaded network)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    pyoracle.lib()          # dlopen liboracle.so in THIS process too: the forked workers inherit the mapping
    cfg = CONFIGS[key]
    cores = cores or os.cpu_count() or 1
    warm = cfg["preroll"] if warm is None else warm
    n_inst = n_inst or {"c2": 32, "c3": 4, "c4": 4, "c5": 4}[key]
    n_steps = max(1, min(int(n_steps), 355 - warm))      # one episode is 360 env steps
    rate = cfg["rates"][len(cfg["rates"]) // 2]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(key, n_inst, n_steps, 1, c * n_inst, warm, rate) for c in range(cores)])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return dict(value=total / busy, unit="env steps/s", cores=cores, kind="port",
                sample=f"{cores} procs x {n_inst} {cfg['map']} instances x {n_steps} env steps ({cfg['policy']}"
                       + (f", synthetic {rate:g} veh/h/lane" if rate > 0 else "") + f", after {warm} untimed steps), "
                       f"oracle/microsim.c, max-over-procs busy time {busy:.2f}s (wall {wall:.2f}s)")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions: an NVML polling thread (2 ms period; a 20-step window of the
    headline config lasts ~50 ms, too short for an `nvidia-smi -lms` child to report once), gated by window(): samples
    are kept only while a timed region is open.  Falls back to the nvidia-smi child if NVML cannot be opened."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, gpu_index):
        self.rows = []            # nvidia-smi fallback rows
        self.sm, self.mask = [], 0
        self.max_mhz = None
        self.proc = None
        self.gpu = gpu_index
        self.open = False         # a timed region is running
        self.alive = False
        self.source = None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip() != ""]
        if ids and all(x.strip().isdigit() for x in ids) and self.gpu < len(ids):
            return int(ids[self.gpu])
        return self.gpu

    def start(self):
        if self.alive or self.proc is not None:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)

            def poll():
                while self.alive:
                    if self.open:
                        try:
                            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                            self.mask |= int(reasons(h))
                        except Exception:
                            pass
                    time.sleep(0.002)
            self.alive, self.source = True, "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.alive = False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100 (whole run)"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def window(self, is_open):
        self.open = bool(is_open)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.alive:
            self.alive = False
            self.thread.join(timeout=1)
            return dict(sm_mhz=float(np.median(self.sm)) if self.sm else None,
                        sm_min_mhz=float(np.min(self.sm)) if self.sm else None, sm_max_mhz=self.max_mhz,
                        reasons=sorted(n for bit, n in self.REASONS if self.mask & bit), samples=len(self.sm),
                        source="NVML polled every 2 ms inside the timed regions (device-timed steps and e2e steps)")
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), source=self.source)


def algorithmic_bytes_per_step(m, vbar, obs_floats_per_signal=13):
    """B_step = T*B_tick + L_in*20 + S*(obs_dim*4+4),
    B_tick = V*(24R+16W) + L*(8R+8W) + S*(8R+8W) + K*1R."""
    st = m.struct
    n_tl_links = int(sum(len(p[0][1]) for p in m.info["programs_installed"].values()))
    b_tick = vbar * 40.0 + st.n_lanes * 16.0 + st.n_signals * 16.0 + n_tl_links
    return st.step_length * b_tick + st.n_sig_lanes * 20.0 + st.n_signals * (obs_floats_per_signal * 4 + 4)


def kernel_counters(key):
    """ncu-derived per-launch counters of the config's dominant kernel (profiles/kernel_counters.json, written from the
    committed ncu captures by tools/ncu_counters.py): DRAM bytes and executed warp instructions per env step."""
    p = os.path.join(ROOT, "profiles", "kernel_counters.json")
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from resco_b200.sim import HostWaveAgent, VecSim, build_library
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if not os.path.exists(os.path.join(ROOT, "resco_b200", "csrc", "libresco_b200.so")):
        build_library()
    n_env = args.n_env or cfg["n_env"]
    policy = cfg["policy"]
    shared = args.config == "c4"
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_sim(sc, m, count, first):
        sim = VecSim(m, count, seed=args.seed, device=local)
        if cfg["outputs"]:
            sim.select_outputs(*cfg["outputs"])
        sim.reset(args.seed, first)
        sim.observe()
        pairs, va, sig = sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"]
        if policy == "maxpressure":
            sim.policy_maxpressure(pairs, va, sig)         # uploads the action tables
        elif policy == "frap":
            sim.load_frap(frap_state_dict(len(pairs))[0], pairs, va, sig)
        return sim

    sweep = []
    clocks = None
    tot_steps, tot_ms = 0, 0.0
    sampler = ClockSampler(local)
    last = None
    for rate in cfg["rates"]:
        sc, m = _marshal(cfg["map"], cfg["vcap"], cfg["tile"] if args.tile < 0 else args.tile, rate)
        episode_steps = m.struct.end_tick // m.struct.step_length
        S = m.struct.n_signals
        if shared:       # two half-batches per rank on two streams: the collectives of one overlap the kernel of the other
            n_a = n_env // 2
            parts = [(0, n_a), (n_a, n_env - n_a)]
            streams = [torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)]
        else:
            parts = [(0, n_env)]
            streams = [stream]
        sims = [make_sim(sc, m, cnt, rank * n_env + first) for first, cnt in parts]
        gather = [torch.empty((world * cnt, S, 13), device=dev) for _, cnt in parts] if shared and world > 1 else None
        acts_all = [torch.empty((world * cnt, S), dtype=torch.int32, device=dev) for _, cnt in parts] if shared else None
        gather_rew = [torch.empty((world * cnt, S), device=dev) for _, cnt in parts] if shared and world > 1 else None
        state = {"step": 0}

        def one_step():
            if not shared:
                sims[0].env_step_policy(policy, seed=args.seed)
                return
            for h, (sim, st) in enumerate(zip(sims, streams)):
                with torch.cuda.stream(st):
                    cnt = parts[h][1]
                    if world > 1 and args.c4_policy == "rank0":
                        # the literal reading of configs[3]: obs all-gather, the shared policy evaluated by rank 0 for ALL
                        # instances, actions broadcast -- rank 0's FRAP time grows with the number of GPUs
                        dist.all_gather_into_tensor(gather[h], sim.obs_view()["mplight"])
                        if rank == 0:
                            acts_all[h].copy_(sim.policy_frap(gather[h]))
                        dist.broadcast(acts_all[h], src=0)
                        sim.env_step(acts_all[h][rank * cnt:(rank + 1) * cnt])
                    else:
                        # acting: the shared controller's weights are replicated, every rank evaluates FRAP for its own
                        # instances (one graph launch: policy + env step); learning side: the step's observations and
                        # rewards are all-gathered over NVLink so that the learner rank holds the full batch
                        # (SharedDQN.observe, agents/pfrl_dqn.py) -- on this half's stream, overlapping the other half's kernel
                        sim.env_step_policy("frap")
                        if world > 1:
                            dist.all_gather_into_tensor(gather[h], sim.obs_view()["mplight"])
                            dist.all_gather_into_tensor(gather_rew[h], sim.obs_view()["reward_pressure"])

        def preroll():
            # untimed set-up: reset and run the episode up to a loaded network before anything is timed
            for h, sim in enumerate(sims):
                sim.reset(args.seed, rank * n_env + parts[h][0])
                sim.observe()
            for _ in range(cfg["preroll"] if args.preroll < 0 else args.preroll):
                one_step()
            torch.cuda.synchronize()
            state["step"] = cfg["preroll"] if args.preroll < 0 else args.preroll

        def ensure_room():
            if state["step"] + 1 > episode_steps:
                preroll()

        preroll()
        if rank == 0:
            sampler.start()                     # (idempotent) the polling thread is up before the warm-up steps
        for _ in range(args.warmup):
            ensure_room(); one_step(); state["step"] += 1
        barrier()
        st0 = [s.stats() for s in sims]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kern_ms, deferred, redone = [], 0, 0
        barrier()
        if args.profile:
            torch.cuda.profiler.start()         # ncu --profile-from-start off: captures start at the timed steps
        sampler.window(True)
        wall0 = time.perf_counter()
        for i in range(args.steps):
            ensure_room()
            flush.zero_()                       # evict the instance tiles from L2 between timed steps (untimed)
            ev[i][0].record(stream)
            for st in streams:
                if st is not stream:
                    st.wait_event(ev[i][0])
            one_step()
            for st in streams:
                if st is not stream:
                    stream.wait_stream(st)
            ev[i][1].record(stream)
            state["step"] += 1
            if not shared and (i % 4 == 3 or i == args.steps - 1):
                torch.cuda.synchronize()        # every 4th step: kernel time of the graph launch + deferred-instance count
                kern_ms.append(sims[0].last_step_ms())
                ti = sims[0].tile_info()
                deferred, redone = max(deferred, ti["last_deferred"]), max(redone, ti["last_redone"])
        barrier()
        sampler.window(False)
        if args.profile:
            torch.cuda.profiler.stop()
        wall = time.perf_counter() - wall0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        st1 = [s.stats() for s in sims]
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max = float(t.item())
        ticks = int(st1[0]["tick"][0] - st0[0]["tick"][0])
        act_ticks = sum(float((b["sum_active_ticks"].astype(np.float64) - a["sum_active_ticks"]).sum()) for a, b in zip(st0, st1))
        vbar = act_ticks / max(ticks, 1) / n_env
        refused = int(sum(int((b["n_cap_refused"] - a["n_cap_refused"]).sum()) for a, b in zip(st0, st1)))
        sweep.append(dict(rate_veh_h_lane=rate, value=n_env * world * args.steps / (dev_ms_max / 1e3),
                          ms_per_step=dev_ms_max / args.steps, vbar_active_vehicles=vbar, n_cap_refused=refused,
                          max_redone_in_cta_per_step=redone, max_deferred_to_overflow_pass_per_step=deferred, backlog_mean=float(np.mean(np.concatenate([b["n_backlog"] for b in st1]))),
                          kernel_ms=float(np.mean(kern_ms)) if kern_ms else None, wall_s=wall))
        tot_steps += args.steps
        tot_ms += dev_ms_max
        last = (sc, m, sims, streams, parts, vbar, kern_ms, S)
        if rate != cfg["rates"][-1]:
            for s in sims:
                s.close()
    sc, m, sims, streams, parts, vbar, kern_ms, S = last
    total_env = n_env * world
    value = total_env * tot_steps / (tot_ms / 1e3)
    refused_total = sum(x["n_cap_refused"] for x in sweep)
    shape = sims[0].launch_shape()
    tile = sims[0].tile_info(with_deferred=False)

    # ---- end-to-end through the host-buffer C-ABI call + host agent: two sims holding half of the rank's instances
    #      each, stepped alternately (rs_env_step_host_async / rs_wait), so the host agent of one half overlaps the
    #      device step of the other.  Same instances (global ids), same bytes over PCIe per env step. ----
    for s in sims:
        s.close()
    n_a = n_env // 2
    hparts = [(0, n_a), (n_a, n_env - n_a)]
    hstreams = [torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)]
    halves = []
    if policy == "maxpressure":
        hagent = HostWaveAgent(sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"])
        hact = [lambda obs: hagent(obs)] * 2
    elif policy == "random":
        rngs = [np.random.default_rng(args.seed + rank * 2 + k) for k in range(2)]
        ng = np.asarray([len(m.info["green_states"][s]) for s in m.info["signal_ids"]], np.int64)
        hact = [(lambda obs, r=r, c=c: (r.integers(0, 1 << 30, (c, S)) % ng[None, :]).astype(np.int32)) for r, (_, c) in zip(rngs, hparts)]
    else:
        hact = None     # MPLight: the shared policy runs on the device; the host only reads the reward back
    for first, cnt in hparts:
        h = make_sim(sc, m, cnt, rank * n_env + first)
        h.set_host_obs(cfg["host_obs"])
        pre = cfg["preroll"] if args.preroll < 0 else args.preroll
        for _ in range(pre):
            h.env_step_policy(policy, seed=args.seed) if policy != "frap" else h.env_step(h.policy_frap())
        halves.append(h)
    torch.cuda.synchronize()
    episode_steps = m.struct.end_tick // m.struct.step_length
    pre = cfg["preroll"] if args.preroll < 0 else args.preroll
    pipe_warm = 3
    pipe_steps = max(4, min(args.steps, 100, episode_steps - pre - pipe_warm - 2))
    rk = cfg["reward_kind"]
    obs_bytes = int(np.prod(halves[0]._host_obs_shape[1:])) * 4 * n_env
    if hact is not None:
        obs_half = [None, None]
        for hi, (h, st) in enumerate(zip(halves, hstreams)):
            h.env_step_host_async(hact[hi](h.obs()[cfg["host_obs"]]), reward_kind=rk, stream=st)
        obs_half = [h.wait()[0] for h in halves]
        for _ in range(pipe_warm - 1):
            for hi, (h, st) in enumerate(zip(halves, hstreams)):
                h.env_step_host_async(hact[hi](obs_half[hi]), reward_kind=rk, stream=st)
            obs_half = [h.wait()[0] for h in halves]
        barrier()
        sampler.window(True)
        t0 = time.perf_counter()
        for hi, (h, st) in enumerate(zip(halves, hstreams)):        # prime: one step in flight per half
            h.env_step_host_async(hact[hi](obs_half[hi]), reward_kind=rk, stream=st)
        for _ in range(pipe_steps - 1):
            with torch.cuda.stream(hstreams[0]):
                flush.zero_()                                       # L2 eviction once per iteration, INSIDE the timed region
                fl = hstreams[0].record_event()
            hstreams[1].wait_event(fl)
            for hi, (h, st) in enumerate(zip(halves, hstreams)):
                o, _ = h.wait()
                h.env_step_host_async(hact[hi](o), reward_kind=rk, stream=st)
        for h in halves:
            h.wait()
        e2e_t = time.perf_counter() - t0
        e2e = dict(h2d_bytes_per_step=n_env * S * 4, d2h_bytes_per_step=obs_bytes + n_env * S * 4,
                   mode="2 half-batch sims double-buffered on 2 streams (rs_env_step_host_async / rs_wait), host agent "
                        f"({policy}) on the returned {cfg['host_obs']} observation, L2 flush inside the timed region")
    else:
        rew_host = [torch.empty((cnt, S), dtype=torch.float32).pin_memory() for _, cnt in hparts]
        key_r = "reward_pressure"

        def mp_step():
            for hi, (h, st) in enumerate(zip(halves, hstreams)):
                with torch.cuda.stream(st):
                    h.env_step(h.policy_frap())
                    rew_host[hi].copy_(h.obs_view()[key_r], non_blocking=True)
        for _ in range(pipe_warm):
            mp_step()
        barrier()
        sampler.window(True)
        t0 = time.perf_counter()
        for _ in range(pipe_steps):
            mp_step()
            with torch.cuda.stream(hstreams[0]):
                flush.zero_()
        torch.cuda.synchronize()
        e2e_t = time.perf_counter() - t0
        e2e = dict(h2d_bytes_per_step=0, d2h_bytes_per_step=n_env * S * 4,
                   mode="MPLight's shared policy runs on the device (rs_policy_frap), so a step's inputs never leave HBM; "
                        "per step: FRAP + fused env step per half-batch on 2 streams, D2H of the pressure reward into pinned "
                        "memory, L2 flush inside the timed region")
    sampler.window(False)
    clocks = sampler.stop() if rank == 0 else None
    te = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e.update(value=total_env * pipe_steps / float(te.item()), unit="env steps/s", steps=pipe_steps)
    for h in halves:
        h.close()

    if rank == 0:
        obs_fl = {"mplight": 13, "drq_norm": 13}[cfg["host_obs"]]
        bstep = algorithmic_bytes_per_step(m, vbar, obs_fl)
        ms_step = tot_ms / tot_steps
        k_ms = float(np.mean(kern_ms)) if kern_ms else ms_step
        k_ms = min(k_ms, ms_step)          # a kernel cannot take longer than the step that contains it
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak = 6650.0; peak_src = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        achieved = bstep * n_env / (k_ms / 1e3) / 1e9
        kc = kernel_counters(args.config)
        traffic = kc["dram_bytes_per_env_step"] * n_env if kc else None
        issue = None
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6 if clocks else 1965e6
        if kc and kc.get("warp_insts_per_env_step"):
            t_issue = kc["warp_insts_per_env_step"] * n_env / (148 * 4 * sm_hz) * 1e3
            issue = dict(warp_insts_per_env_step=kc["warp_insts_per_env_step"], issue_slots_per_s=148 * 4 * sm_hz,
                         ms_at_full_issue_rate=t_issue, frac=t_issue / k_ms, source=kc.get("source"),
                         what="executed warp instructions of one launch / (148 SMs x 4 schedulers x SM clock) over the measured kernel time: the "
                              "fraction of the issue-slot ceiling this instruction stream reaches")
        cb = cpu_baseline(args.config) if world == 1 and not args.no_cpu else None
        out = {
            "metric": "env steps/sec (summed instances)", "value": value, "unit": "env steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["what"].format(n=n_env), "name": args.config,
                       "n_env_per_gpu": n_env, "n_env_total": total_env, "sim_ticks_per_env_step": m.struct.step_length,
                       "vcap": m.struct.vcap, **tile, **shape,
                       "n_cap_refused_in_timed_window": refused_total,
                       "l2": "flushed between timed steps (160 MiB memset > 126 MB L2, untimed)",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                       "allgather_obs": bool(shared and world > 1), "c4_policy": args.c4_policy if shared else None,
                       "avg_delay_parity_vs_sumo": "not measurable here: SUMO/libsumo is installed neither in the build container nor on the GPU box; statistical anchors against utils/avg_timeLoss.py: tests/test_anchors.py, DESIGN.md section 7",
                       "preroll_env_steps": cfg["preroll"] if args.preroll < 0 else args.preroll,
                       "episode_window": "timed steps start after an untimed pre-roll of the episode (loaded network); the episode restarts (reset + pre-roll, untimed) when its steps are used up",
                       "max_redone_in_cta_per_step": max(x["max_redone_in_cta_per_step"] for x in sweep),
                       "max_deferred_to_overflow_pass_per_step": max(x["max_deferred_to_overflow_pass_per_step"] for x in sweep),
                       "sweep": sweep if len(sweep) > 1 else None},
            "sim_ticks_per_s": value * m.struct.step_length,
            "e2e": e2e,
            "gpu_launches": int(tot_steps * (2 + (1 if tile["overflow_pass"] else 0)) * len(parts)),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "rs::k_run (fused env step)", "kernel_ms": k_ms,
                         "kernel_ms_source": "mean of rs_last_step_ms (CUDA events around the graph launch) over every 4th timed step, capped by ms_per_step",
                         "algorithmic_bytes_per_env_step": bstep, "vbar_active_vehicles": vbar, "peak_source": peak_src,
                         "issue": issue,
                         "note": "algorithmic bytes follow SURVEY 8(d) (one tile round trip PER TICK); the fused kernel moves the tile once per env step, so DRAM traffic is far below the algorithmic count and the kernel is latency/issue bound, not HBM bound: roofline.issue is the ceiling that binds"},
            "cpu_baseline": cb, "clocks": clocks,
        }
        if refused_total:
            out["invalid"] = f"{refused_total} insertions were refused by the vehicle store in the timed window: raise vcap"
        print(json.dumps(out))
        if refused_total:
            sys.exit(3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    t0 = time.perf_counter()
    # one "step" of this arm = one env step of every core's instances; W warm-up steps untimed, K timed (K is capped
    # by the episode length: the sample stays bounded whatever K the caller passes)
    pre = cfg["preroll"] if args.preroll < 0 else args.preroll
    cb = cpu_baseline(args.config, n_steps=max(args.steps, 1), warm=pre + max(args.warmup, 0))
    wall = time.perf_counter() - t0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_inst = {"c2": 32, "c3": 4, "c4": 4, "c5": 4}[args.config]
    out = {"impl": "reference", "metric": "env steps/sec (summed instances)", "value": cb["value"],
           "unit": "env steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * n_inst * cb["cores"] / cb["value"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["what"].format(n=f"bounded sample: {n_inst} per host core;"), "name": args.config,
                      "note": "reference CPU path (MultiSignal over libsumo) unavailable: SUMO is not installed and its source is not in the reference tree; this arm is the CPU oracle port of the same algorithm (kind=port)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--n-env", type=int, default=0, help="instances per GPU (0: the config's)")
    ap.add_argument("--tile", type=int, default=-1, help="vehicles in the shared-memory tile (-1: the config's)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--preroll", type=int, default=-1, help="untimed env steps after reset before warm-up (-1: the config's)")
    ap.add_argument("--c4-policy", default="replicated", choices=["replicated", "rank0"],
                    help="c4: every rank evaluates the shared FRAP for its own instances and the observations are all-gathered for "
                         "the learner (default), or rank 0 evaluates the gathered batch and broadcasts the actions")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
