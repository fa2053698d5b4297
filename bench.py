#!/usr/bin/env python
"""bench.py -- env steps/sec of the hot path (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- cologne8 (8 signals), MaxPressure agent,
4096 lock-step environment instances per GPU ("weak" scaling: per-GPU work is fixed).  One "step" =
one MultiSignal.step() of every instance = batched MaxPressure action selection + step_length (10)
one-second simulation ticks + Signal.observe + states.mplight + rewards (multi_signal.py:164-197).
Demand is the map's own 2046-trip table replicated to every instance (synthetic in the sense that
per-instance driver randomness -- speedFactor, dawdling -- is drawn from Philox keyed by the
global instance id); "data": "synthetic".

`value`  : device-timed (CUDA events on the launching stream), actions/obs resident in HBM.
`e2e`    : the same metric through the host-buffer C-ABI call (rs_env_step_host) with the batched
           host agent: H2D of actions + D2H of obs/reward inside the timed region.
`--impl reference`: the CPU arm.  The reference's own CPU path (MultiSignal over libsumo) cannot
           run anywhere in this environment (SUMO is not installed), so this arm times
           the CPU oracle port of the same algorithm on all host cores (cpu_baseline.kind = "port").
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

MAP = "cologne8"
N_ENV_PER_GPU = 4096
VCAP = 128   # cologne8/MaxPressure peaks at ~110 concurrent vehicles per instance (32-seed CPU check); a full tile only delays insertions


def _marshal(map_name=MAP, vcap=VCAP, synthetic_rate=0.0):
    from resco_b200.abi import marshal
    from resco_b200.scenario import Scenario
    sc = Scenario.load(os.path.join(ROOT, "resco_b200", "data", map_name + ".npz"))
    mc = sc.meta["map_config"]
    synth = None
    if synthetic_rate > 0:
        from resco_b200.scenario.synth import synth_demand
        synth = synth_demand(sc, synthetic_rate)
    m = marshal(sc, step_length=mc["step_length"], yellow_length=mc["yellow_length"], max_distance=200.0, vcap=vcap,
                synthetic=synth)
    return sc, m


def host_maxpressure(sc, m):
    """Batched MAXPRESSURE agent over host observation buffers (agents/maxpressure.py + maxwave.py:18-38): the
    library's host-side agent front-end (rs_host_agent_wave)."""
    from resco_b200.sim import HostWaveAgent
    return HostWaveAgent(sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"])


def numpy_maxpressure(sc, m):
    """The same agent in numpy (CPU-baseline workers: the oracle arm must not depend on the CUDA library)."""
    pairs = np.asarray(sc.meta["phase_pairs"], np.int64)
    va = sc.meta["valid_acts"]
    sig = m.info["signal_ids"]
    tables = []
    for s in sig:
        idxs = list(range(len(pairs))) if va is None else [int(k) for k in va[s].keys()]
        acts = idxs if va is None else [va[s][str(k)] for k in idxs]
        tables.append((pairs[idxs, 0] + 1, pairs[idxs, 1] + 1, np.asarray(acts, np.int32)))

    def act(mplight):
        out = np.empty(mplight.shape[:2], np.int32)
        for i, (p0, p1, acts) in enumerate(tables):
            press = mplight[:, i, p0] + mplight[:, i, p1]
            out[:, i] = acts[np.argmax(press, 1)]
        return out
    return act


# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process = one core: `n_inst` oracle instances, `warm` untimed + `n_steps` timed env steps with MaxPressure."""
    n_inst, n_steps, seed, first, warm = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import OracleSim
    sc, m = _marshal()
    agent = numpy_maxpressure(sc, m)
    sim = OracleSim(m, n_inst, seed=seed)
    sim.reset(seed, first)
    sim.observe()
    for _ in range(warm):
        sim.env_step(agent(sim.obs()["mplight"]))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        sim.env_step(agent(sim.obs()["mplight"]))
    return time.perf_counter() - t0, n_inst * n_steps


def cpu_baseline(n_steps=120, n_inst=32, cores=None, warm=90):
    """Oracle port on the host cores, bounded sample (about 10-30 s of CPU work in total): every core steps its own
    `n_inst` instances `n_steps` env steps after `warm` untimed ones (the GPU arm's pre-roll: both arms time the loaded network)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    cores = cores or os.cpu_count() or 1
    n_steps = max(1, min(int(n_steps), 355 - warm))      # one episode is 360 env steps
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(n_inst, n_steps, 1, c * n_inst, warm) for c in range(cores)])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return dict(value=total / busy, unit="env steps/s", cores=cores, kind="port",
                sample=f"{cores} procs x {n_inst} cologne8 instances x {n_steps} env steps (MaxPressure, after {warm} untimed "
                       f"steps), oracle/microsim.c, max-over-procs busy time {busy:.2f}s (wall {wall:.2f}s)")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
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
                    reasons=sorted(reasons), samples=len(sm))


def algorithmic_bytes_per_step(m, vbar):
    """B_step = T*B_tick + L_in*20 + S*(obs_dim*4+4),
    B_tick = V*(24R+16W) + L*(8R+8W) + S*(8R+8W) + K*1R."""
    st = m.struct
    n_tl_links = int(sum(len(p[0][1]) for p in m.info["programs_installed"].values()))
    b_tick = vbar * 40.0 + st.n_lanes * 16.0 + st.n_signals * 16.0 + n_tl_links
    return st.step_length * b_tick + st.n_sig_lanes * 20.0 + st.n_signals * (13 * 4 + 4)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from resco_b200.sim import VecSim, build_library
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not os.path.exists(os.path.join(ROOT, "resco_b200", "csrc", "libresco_b200.so")):
        build_library()
    sc, m = _marshal(args.map, args.vcap, args.synthetic_rate)
    n_env = args.n_env
    sim = VecSim(m, n_env, seed=args.seed, device=local)
    sim.reset(args.seed, rank * n_env)
    sim.observe()
    pairs, va, sig = sc.meta["phase_pairs"], sc.meta["valid_acts"], m.info["signal_ids"]
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    stream = torch.cuda.current_stream()
    gather_buf = None
    if args.allgather and world > 1:
        gather_buf = torch.empty((world,) + tuple(sim.obs_view()["mplight"].shape), device=f"cuda:{local}")

    def one_step():
        act = sim.policy_maxpressure(pairs, va, sig)
        sim.env_step(act)
        if gather_buf is not None:      # shared-policy configs (C4): obs all-gather over NVLink
            dist.all_gather_into_tensor(gather_buf, sim.obs_view()["mplight"])

    episode_steps = m.struct.end_tick // m.struct.step_length
    state = {"step": 0}

    def preroll():
        # untimed set-up: reset and run the episode up to a loaded network before anything is timed
        sim.reset(args.seed, rank * n_env)
        sim.observe()
        for _ in range(args.preroll):
            act = sim.policy_maxpressure(pairs, va, sig)
            sim.env_step(act)
        state["step"] = args.preroll

    def ensure_room():
        if state["step"] + 1 > episode_steps:
            preroll()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    preroll()
    for _ in range(args.warmup):
        ensure_room(); one_step(); state["step"] += 1
    barrier()
    st0 = sim.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        ensure_room()
        flush.zero_()                       # evict the instance tiles from L2 between timed steps (untimed)
        ev[i][0].record(stream)
        one_step()
        ev[i][1].record(stream)
        state["step"] += 1
        if i % 16 == 15:
            torch.cuda.synchronize()
            kern_ms.append(sim.last_step_ms())
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    launches = args.steps * 2   # per timed step: k_policy + k_run (pre-roll/reset launches are untimed set-up)
    st1 = sim.stats()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    total_env = n_env * world
    value = total_env * args.steps / (dev_ms_max / 1e3)

    # ---- end-to-end through the host-buffer C-ABI call + host agent ----
    agent = host_maxpressure(sc, m)
    obs_h = sim.obs()["mplight"]
    e2e_steps = max(8, min(args.steps, 100))
    if state["step"] + e2e_steps + 3 > episode_steps:
        preroll()
        obs_h = sim.obs()["mplight"]
    for _ in range(3):
        obs_h, _ = sim.env_step_host(agent(obs_h), reward_kind=0)
    barrier()
    e2e_t = 0.0
    for _ in range(e2e_steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        obs_h, rew_h = sim.env_step_host(agent(obs_h), reward_kind=0)
        e2e_t += time.perf_counter() - t0
    te = torch.tensor([e2e_t], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sync_value = total_env * e2e_steps / float(te.item())

    # ---- the same loop double-buffered: two sims hold half of the rank's instances each and are stepped
    #      alternately (rs_env_step_host_async / rs_wait), so the host agent of one half overlaps the device
    #      step of the other.  Same instances (global ids), same agent, same bytes over PCIe per env step. ----
    n_a = n_env // 2
    halves, streams = [], [torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)]
    for first, cnt in ((0, n_a), (n_a, n_env - n_a)):
        h = VecSim(m, cnt, seed=args.seed, device=local)
        h.reset(args.seed, rank * n_env + first)
        h.observe()
        for _ in range(args.preroll):
            h.env_step(h.policy_maxpressure(pairs, va, sig))
        halves.append(h)
    pipe_warm = 3
    pipe_steps = max(8, min(args.steps, episode_steps - args.preroll - pipe_warm - 2))
    obs_half = [h.obs()["mplight"] for h in halves]
    for _ in range(pipe_warm):           # untimed: first calls allocate the page-locked host buffers of the two halves
        for hi, (h, st) in enumerate(zip(halves, streams)):
            h.env_step_host_async(agent(obs_half[hi]), reward_kind=0, stream=st)
        obs_half = [h.wait()[0] for h in halves]
    barrier()
    t0 = time.perf_counter()
    for h, st, o in zip(halves, streams, obs_half):            # prime: one step in flight per half
        h.env_step_host_async(agent(o), reward_kind=0, stream=st)
    for _ in range(pipe_steps - 1):
        with torch.cuda.stream(streams[0]):
            flush.zero_()                                       # L2 eviction once per iteration, INSIDE the timed region
            fl = streams[0].record_event()
        streams[1].wait_event(fl)
        for h, st in zip(halves, streams):
            o, _ = h.wait()
            h.env_step_host_async(agent(o), reward_kind=0, stream=st)
    for h in halves:
        h.wait()
    e2e_t = time.perf_counter() - t0
    te = torch.tensor([e2e_t], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_env * pipe_steps / float(te.item())
    for h in halves:
        h.close()
    S = sim.S

    if rank == 0:
        ticks = int(st1["tick"][0] - st0["tick"][0])
        vbar = float((st1["sum_active_ticks"].astype(np.float64) - st0["sum_active_ticks"]).mean() / max(ticks, 1))
        bstep = algorithmic_bytes_per_step(m, vbar)
        k_ms = float(np.mean(kern_ms)) if kern_ms else dev_ms / args.steps
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak = 6650.0; peak_src = "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        achieved = bstep * n_env / (k_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("k_run_env_step")
        cb = cpu_baseline() if world == 1 and not args.no_cpu else None
        out = {
            "metric": "env steps/sec (summed instances)", "value": value, "unit": "env steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.map} ({sim.S} signals) / MaxPressure / {n_env} lock-step instances per GPU"
                                   + (f" / synthetic Bernoulli demand {args.synthetic_rate:g} veh/h/entry-lane" if args.synthetic_rate > 0 else ""),
                       "n_env_per_gpu": n_env, "n_env_total": total_env, "sim_ticks_per_env_step": m.struct.step_length,
                       "vcap": m.struct.vcap, **sim.launch_shape(), "persistent_grid": os.environ.get("RESCO_B200_PERSIST", "1") != "0",
                       "l2": "flushed between timed steps (160 MiB memset > 126 MB L2, untimed)",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                       "allgather_obs": bool(gather_buf is not None),
                       "avg_delay_parity_vs_sumo": "not measurable here: SUMO/libsumo is installed neither in the build container nor on the GPU box; statistical anchors against utils/avg_timeLoss.py are in DESIGN.md section 7",
                       "preroll_env_steps": args.preroll,
                       "episode_window": "timed steps start after an untimed pre-roll of the episode (loaded network); the episode restarts (reset + pre-roll, untimed) when its 360 steps are used up"},
            "sim_ticks_per_s": value * m.struct.step_length,
            "e2e": {"value": e2e_value, "unit": "env steps/s", "h2d_bytes_per_step": n_env * S * 4,
                    "d2h_bytes_per_step": n_env * S * 13 * 4 + n_env * S * 4, "steps": pipe_steps,
                    "mode": "2 half-batch sims double-buffered on 2 streams (rs_env_step_host_async / rs_wait), host MaxPressure agent (rs_host_agent_wave), L2 flush inside the timed region",
                    "sync_value": e2e_sync_value,
                    "what": "per env step: pinned H2D of the actions, fused env step, D2H of mplight obs + reward, host MaxPressure agent (rs_host_agent_wave) on the returned obs; wall clock; value = double-buffered (mode), sync_value = one rs_env_step_host call per step over the whole batch"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "rs::k_run<BLOCK> (fused env step)", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_env_step": bstep, "vbar_active_vehicles": vbar, "peak_source": peak_src,
                         "note": "algorithmic bytes follow SURVEY §8(d) (one tile round trip PER TICK); the fused kernel moves the tile once per env step, so DRAM traffic is far below the algorithmic count and the kernel is latency/issue bound, not HBM bound"},
            "cpu_baseline": cb, "clocks": clocks, "wall_s": wall,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    # one "step" of this arm = one env step of every core's 32 instances; W warm-up steps untimed, K timed (K is capped
    # by the episode length: the sample stays bounded whatever K the caller passes)
    cb = cpu_baseline(n_steps=max(args.steps, 1), n_inst=32, warm=args.preroll + max(args.warmup, 0))
    wall = time.perf_counter() - t0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "env steps/sec (summed instances)", "value": cb["value"],
           "unit": "env steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * 32 * cb["cores"] / cb["value"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{MAP} (8 signals) / MaxPressure / bounded sample: 32 instances per host core",
                      "note": "reference CPU path (MultiSignal over libsumo) unavailable: SUMO is not installed and its source is not in the reference tree; this arm is the CPU oracle port of the same algorithm (kind=port)"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--map", default=MAP)
    ap.add_argument("--n-env", type=int, default=N_ENV_PER_GPU)
    ap.add_argument("--vcap", type=int, default=VCAP)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--preroll", type=int, default=90, help="untimed env steps after reset before warm-up")
    ap.add_argument("--synthetic-rate", type=float, default=0.0, help="veh/h per entry lane (configs[4]); 0 = map's trip table")
    ap.add_argument("--allgather", action="store_true", help="NCCL all-gather of the mplight obs every step (C4)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
