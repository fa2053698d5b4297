"""VecSim: Python handle on the C-ABI library (``include/resco_b200.h``).

PyTorch is used only as plumbing: it owns the CUDA context/streams and wraps the borrowed device
pointers of the observation view as tensors (zero copy).  All simulation work happens in
``resco_b200/csrc/libresco_b200.so``; there is no CPU fallback -- constructing a VecSim without the
library or without a Blackwell GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from .abi import HOSTOBS, Marshalled, RsObsView, RsScenario, RsStats, STATS_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("RESCO_B200_LIB", os.path.join(_CSRC, "libresco_b200.so"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

_LIB = None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, "sim.cu")]
    deps = srcs + [os.path.join(_CSRC, "sim_kernels.cuh"), os.path.join(_CSRC, "agents.cuh"),
                   os.path.join(_HERE, "..", "include", "resco_b200.h")]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    subprocess.check_call(cmd)
    return LIB_PATH


EXPORTS = ["rs_last_error", "rs_abi_version", "rs_create", "rs_destroy", "rs_reset", "rs_set_phase", "rs_tick",
           "rs_observe", "rs_env_step", "rs_env_step_host", "rs_env_step_host_async", "rs_wait", "rs_policy_maxpressure", "rs_host_agent_wave", "rs_get_obs", "rs_get_stats",
           "rs_dump_vehicles", "rs_get_phases", "rs_get_trip_records", "rs_kernel_launches", "rs_last_step_ms", "rs_get_launch_shape",
           "rs_select_outputs", "rs_set_host_obs", "rs_frap_load", "rs_policy_frap", "rs_policy_random", "rs_env_step_policy",
           "rs_get_tile_info", "rs_set_demand_window"]


def load_library():
    """dlopen the in-tree library (fails loudly if it has not been built)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the B200 backend has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.rs_last_error.restype = C.c_char_p
    lib.rs_abi_version.restype = C.c_int
    lib.rs_create.argtypes = [C.POINTER(RsScenario), C.c_int32, C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.rs_destroy.argtypes = [C.c_void_p]
    lib.rs_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p]
    lib.rs_set_phase.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rs_set_demand_window.argtypes = [C.c_void_p, C.c_void_p]
    lib.rs_tick.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.rs_observe.argtypes = [C.c_void_p, C.c_void_p]
    lib.rs_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rs_env_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.rs_env_step_host_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.rs_wait.argtypes = [C.c_void_p]
    lib.rs_policy_maxpressure.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p]
    lib.rs_host_agent_wave.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p]
    lib.rs_get_obs.argtypes = [C.c_void_p, C.POINTER(RsObsView)]
    lib.rs_get_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.rs_dump_vehicles.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)] + [C.c_void_p] * 14
    lib.rs_get_phases.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.rs_get_trip_records.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
    lib.rs_select_outputs.argtypes = [C.c_void_p, C.c_int32]
    lib.rs_set_host_obs.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
    lib.rs_frap_load.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.rs_policy_frap.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.rs_policy_random.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.rs_env_step_policy.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p]
    lib.rs_get_tile_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 6
    lib.rs_kernel_launches.restype = C.c_int64
    lib.rs_kernel_launches.argtypes = [C.c_void_p]
    lib.rs_last_step_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.rs_get_launch_shape.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 5
    _LIB = lib
    return lib


class RsError(RuntimeError):
    pass


def _check(lib, rc: int):
    if rc != 0:
        raise RsError(f"resco_b200 error {rc}: {lib.rs_last_error().decode()}")


_VEH_FIELDS = [("lane", np.int32), ("pos", np.float32), ("speed", np.float32), ("accel", np.float32),
               ("wait", np.float32), ("rwait", np.float32), ("tloss", np.float32), ("vid", np.int32),
               ("vtype", np.int32), ("route", np.int32), ("cursor", np.int32), ("sf", np.float32),
               ("depart", np.int32), ("acc_wait", np.float32)]

_OBS_FIELDS = [("lane_queue", "f", "L"), ("lane_approach", "f", "L"), ("lane_total_wait", "f", "L"),
               ("lane_max_wait", "f", "L"), ("lane_speed_sum", "f", "L"), ("phase", "i", "S"),
               ("mplight", "f", "S13"), ("wave", "f", "S12"), ("reward_wait", "f", "S"),
               ("reward_wait_norm", "f", "S"), ("reward_pressure", "f", "S"), ("sig_queue_len", "i", "S"),
               ("sig_max_queue", "i", "S"), ("lane_arrivals", "f", "L"),
               ("drq", "f", "L5"), ("drq_norm", "f", "L5"), ("mplight_full", "f", "S49")]
_OPTIONAL_OUT = {"drq": 1, "drq_norm": 2, "mplight_full": 4}          # RS_OUT_* bit of the optional tensors


def policy_tables(pairs, valid_acts, signal_ids):
    """(pairs [n_pairs*2], order [S, n_pairs, 2], n_pairs) in the layout rs_policy_maxpressure / rs_host_agent_wave read:
    order[s, k] = (pair index, action) in the reference's evaluation order (iteration order of valid_acts[signal],
    agents/maxwave.py:28-36), pair index -1 terminates a row."""
    npairs = len(pairs)
    pr = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1))
    va = np.full((len(signal_ids), npairs, 2), -1, np.int32)
    for i, s in enumerate(signal_ids):
        if valid_acts is None:
            va[i, :, 0] = np.arange(npairs)
            va[i, :, 1] = np.arange(npairs)
        else:
            for k, (pair_idx, action) in enumerate(valid_acts[s].items()):   # dict order == evaluation order
                va[i, k] = (int(pair_idx), int(action))
    return pr, np.ascontiguousarray(va), npairs


class HostWaveAgent:
    """Batched MAXPRESSURE / MAXWAVE for observation batches in host memory (agents/maxpressure.py, maxwave.py:18-38):
    the act() of a caller that drives env_step_host; runs in the C library (rs_host_agent_wave), no device work."""

    def __init__(self, pairs, valid_acts, signal_ids, use_wave: bool = False):
        self.lib = load_library()
        self.pairs, self.order, self.n_pairs = policy_tables(pairs, valid_acts, signal_ids)
        self.S = len(signal_ids)
        self.skip = 0 if use_wave else 1     # MaxAgent drops the phase entry of states.mplight (maxpressure.py:15-17)

    def act(self, obs: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        obs = np.ascontiguousarray(obs, np.float32)
        n, S, D = obs.shape
        if S != self.S:
            raise ValueError(f"observation batch has {S} signals, agent was built for {self.S}")
        if out is None:
            out = np.empty((n, S), np.int32)
        _check(self.lib, self.lib.rs_host_agent_wave(obs.ctypes.data, n, S, D, self.skip, self.pairs.ctypes.data,
                                                     self.n_pairs, self.order.ctypes.data, out.ctypes.data))
        return out

    __call__ = act


class VecSim:
    """N lock-step environment instances on one GPU."""

    def __init__(self, m: Marshalled, n_env: int, seed: int = 0, device: int = 0):
        import torch  # plumbing only (context, streams, tensor views)
        self._torch = torch
        self.lib = load_library()
        if self.lib.rs_abi_version() != m.struct.abi_version:
            raise RsError("ABI version mismatch between resco_b200.abi and libresco_b200.so")
        self.m = m
        self.n_env = n_env
        self.S = m.struct.n_signals
        self.SL = m.struct.n_sig_lanes
        self.vcap = m.struct.vcap
        self.device = device
        h = C.c_void_p()
        _check(self.lib, self.lib.rs_create(C.byref(m.struct), n_env, device, seed, C.byref(h)))
        self._h = h
        self._views: Optional[Dict[str, object]] = None
        self._policy_tables = None

    # lifecycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.rs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    # simulation --------------------------------------------------------------------------------
    def reset(self, seed: int = 0, first_env_id: int = 0):
        _check(self.lib, self.lib.rs_reset(self._h, seed, first_env_id, self._stream()))

    def set_demand_window(self, origin_off):
        """Per origin lane, the range of the trip table that may depart in the next episode (call before reset)."""
        w = np.ascontiguousarray(origin_off, np.int32)
        assert w.shape == (self.m.struct.n_origins + 1,)
        _check(self.lib, self.lib.rs_set_demand_window(self._h, w.ctypes.data))

    def set_phase(self, phase, mask=None):
        t = self._torch
        p = t.as_tensor(np.asarray(phase), dtype=t.int32, device=f"cuda:{self.device}").contiguous() \
            if not t.is_tensor(phase) else phase.to(dtype=t.int32).contiguous()
        mk = None
        if mask is not None:
            mk = t.as_tensor(np.asarray(mask), dtype=t.uint8, device=f"cuda:{self.device}").contiguous()
        _check(self.lib, self.lib.rs_set_phase(self._h, p.data_ptr(), mk.data_ptr() if mk is not None else None,
                                               self._stream()))

    def tick(self, n: int = 1):
        _check(self.lib, self.lib.rs_tick(self._h, n, self._stream()))

    def observe(self):
        _check(self.lib, self.lib.rs_observe(self._h, self._stream()))

    def env_step(self, actions):
        """actions: [N, S] int32 torch CUDA tensor (or array-like, copied H2D)."""
        t = self._torch
        if not t.is_tensor(actions):
            actions = t.as_tensor(np.ascontiguousarray(actions, np.int32), device=f"cuda:{self.device}")
        actions = actions.to(dtype=t.int32).contiguous()
        assert actions.numel() == self.n_env * self.S
        self._keep_actions = actions
        _check(self.lib, self.lib.rs_env_step(self._h, actions.data_ptr(), self._stream()))

    def select_outputs(self, *names: str):
        """Switch on optional observation tensors ("drq", "drq_norm", "mplight_full"); cumulative."""
        mask = getattr(self, "_out_mask", 0)
        for n in names:
            mask |= _OPTIONAL_OUT[n]
        _check(self.lib, self.lib.rs_select_outputs(self._h, mask))
        self._out_mask = mask
        self._views = None

    def set_host_obs(self, kind: str = "mplight"):
        """Choose the tensor env_step_host / env_step_host_async return as observation: "mplight" [N, S, 13],
        "wave" [N, S, 12], "drq_norm" / "drq" [N, n_sig_lanes, 5], "mplight_full" [N, S, 49]."""
        fl = C.c_int32(0)
        _check(self.lib, self.lib.rs_set_host_obs(self._h, HOSTOBS[kind], C.byref(fl)))
        if kind in _OPTIONAL_OUT:
            self._out_mask = getattr(self, "_out_mask", 0) | _OPTIONAL_OUT[kind]
        self._host_obs_shape = {"mplight": (self.n_env, self.S, 13), "wave": (self.n_env, self.S, 12),
                                "mplight_full": (self.n_env, self.S, 49)}.get(kind, (self.n_env, self.SL, 5))
        self._host_bufs = None
        self._views = None

    def _host_buffers(self):
        if getattr(self, "_host_bufs", None) is None:
            t = self._torch
            shape = getattr(self, "_host_obs_shape", (self.n_env, self.S, 13))
            self._host_bufs = (t.empty((self.n_env, self.S), dtype=t.int32).pin_memory(),
                               t.empty(shape, dtype=t.float32).pin_memory(),
                               t.empty((self.n_env, self.S), dtype=t.float32).pin_memory())
            self._host_np = tuple(b.numpy() for b in self._host_bufs)
        return self._host_bufs, self._host_np

    def env_step_host(self, actions: np.ndarray, reward_kind: int = 0):
        """End-to-end call with HOST buffers (H2D actions, D2H mplight obs + reward).  The returned arrays
        are page-locked buffers owned by this object and are overwritten by the next call."""
        (act_t, obs_t, rew_t), (act_np, obs_np, rew_np) = self._host_buffers()
        np.copyto(act_np, np.asarray(actions).reshape(self.n_env, self.S), casting="unsafe")
        _check(self.lib, self.lib.rs_env_step_host(self._h, act_t.data_ptr(), obs_t.data_ptr(), rew_t.data_ptr(),
                                                   reward_kind))
        return obs_np, rew_np

    def env_step_host_async(self, actions: np.ndarray, reward_kind: int = 0, stream=None):
        """Enqueue one end-to-end step (H2D actions -> env step -> D2H obs + reward) on `stream` (a raw
        cudaStream_t / torch stream; default: the current torch stream) and return immediately; `wait()`
        delivers the results.  Two sims holding half a batch each, stepped alternately, overlap the host
        agent of one half with the device step of the other."""
        (act_t, obs_t, rew_t), (act_np, _, _) = self._host_buffers()
        np.copyto(act_np, np.asarray(actions).reshape(self.n_env, self.S), casting="unsafe")
        st = self._stream() if stream is None else int(getattr(stream, "cuda_stream", stream))
        _check(self.lib, self.lib.rs_env_step_host_async(self._h, act_t.data_ptr(), obs_t.data_ptr(),
                                                         rew_t.data_ptr(), reward_kind, st))

    def wait(self):
        """Block until the step enqueued by `env_step_host_async` has finished -> (obs, reward) host arrays."""
        _check(self.lib, self.lib.rs_wait(self._h))
        return self._host_np[1], self._host_np[2]

    def policy_maxpressure(self, pairs, valid_acts, signal_ids, use_wave: bool = False):
        """Batched MAXPRESSURE / MAXWAVE on the device -> [N, S] int32 CUDA tensor of actions."""
        t = self._torch
        if self._policy_tables is None:
            self._policy_tables = policy_tables(pairs, valid_acts, signal_ids)
            self._policy_out = t.zeros((self.n_env, self.S), dtype=t.int32, device=f"cuda:{self.device}")
        pr, va, npairs = self._policy_tables
        _check(self.lib, self.lib.rs_policy_maxpressure(self._h, pr.ctypes.data, npairs, va.ctypes.data,
                                                        1 if use_wave else 0, self._policy_out.data_ptr(),
                                                        self._stream()))
        return self._policy_out

    def load_frap(self, state_dict, pairs, valid_acts, signal_ids):
        """Upload the FRAP Q-network of MPLight (agents/mplight.py:43-131): `state_dict` maps the reference module's
        parameter names (p.weight, d.weight, ..., before_merge.bias) to arrays / tensors."""
        names = ["p.weight", "d.weight", "d.bias", "lane_embedding.weight", "lane_embedding.bias", "lane_conv.weight",
                 "lane_conv.bias", "relation_embedding.weight", "relation_conv.weight", "relation_conv.bias",
                 "hidden_layer.weight", "hidden_layer.bias", "before_merge.weight", "before_merge.bias"]
        keep = [np.ascontiguousarray(np.asarray(state_dict[n].detach().cpu() if hasattr(state_dict[n], "detach") else state_dict[n],
                                                np.float32).reshape(-1)) for n in names]
        params = (C.c_void_p * len(names))(*[k.ctypes.data for k in keep])
        pr, order, npairs = policy_tables(pairs, valid_acts, signal_ids)
        _check(self.lib, self.lib.rs_frap_load(self._h, params, pr.ctypes.data, npairs, order.ctypes.data))
        self._frap_pairs = npairs

    def policy_frap(self, obs=None, want_q: bool = False):
        """Greedy MPLight actions: FRAP forward over `obs` ([n, S, 13] CUDA float tensor; default: this sim's own
        states.mplight) -> actions [n, S] int32 (and Q-values [n, S, n_pairs] with want_q)."""
        t = self._torch
        n = self.n_env if obs is None else int(obs.shape[0])
        dev = f"cuda:{self.device}"
        if obs is not None:
            obs = obs.to(dtype=t.float32).contiguous()
        acts = t.empty((n, self.S), dtype=t.int32, device=dev)
        q = t.empty((n, self.S, self._frap_pairs), dtype=t.float32, device=dev) if want_q else None
        _check(self.lib, self.lib.rs_policy_frap(self._h, obs.data_ptr() if obs is not None else None, n, acts.data_ptr(),
                                                 q.data_ptr() if want_q else None, self._stream()))
        return (acts, q) if want_q else acts

    def policy_random(self, seed: int = 0):
        """Uniform random green phase per (instance, signal) -> [N, S] int32 CUDA tensor."""
        t = self._torch
        if getattr(self, "_rand_out", None) is None:
            self._rand_out = t.zeros((self.n_env, self.S), dtype=t.int32, device=f"cuda:{self.device}")
        _check(self.lib, self.lib.rs_policy_random(self._h, seed, self._rand_out.data_ptr(), self._stream()))
        return self._rand_out

    POLICIES = {"maxpressure": 1, "maxwave": 2, "frap": 3, "random": 4}

    def env_step_policy(self, policy: str, seed: int = 0):
        """policy kernel + fused env step as ONE CUDA-graph launch (the tables of the policy must have been uploaded by
        one ordinary policy_* / load_frap call)."""
        _check(self.lib, self.lib.rs_env_step_policy(self._h, self.POLICIES[policy], seed, self._stream()))

    def tile_info(self, with_deferred: bool = True) -> dict:
        """tile / store / in-CTA redo capacities; with_deferred: how many instances the last launch stepped again in
        their CTA (`last_redone`) or deferred to the overflow pass (`last_deferred`) -- synchronises the device."""
        v = [C.c_int32(0) for _ in range(6)]
        _check(self.lib, self.lib.rs_get_tile_info(self._h, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]),
                                                   C.byref(v[4]) if with_deferred else None,
                                                   C.byref(v[5]) if with_deferred else None))
        out = dict(tile_vcap=v[0].value, store_vcap=v[1].value, redo_vcap=v[2].value, overflow_pass=bool(v[3].value))
        if with_deferred:
            out.update(last_redone=v[4].value, last_deferred=v[5].value)
        return out

    # results -----------------------------------------------------------------------------------
    def obs_view(self) -> Dict[str, object]:
        """Zero-copy torch views on the device observation buffers (borrowed: valid until close())."""
        if self._views is None:
            t = self._torch
            v = RsObsView()
            _check(self.lib, self.lib.rs_get_obs(self._h, C.byref(v)))
            shapes = {"L": (self.n_env, self.SL), "S": (self.n_env, self.S), "S13": (self.n_env, self.S, 13),
                      "S12": (self.n_env, self.S, 12), "L5": (self.n_env, self.SL, 5), "S49": (self.n_env, self.S, 49)}
            out = {}
            for name, kind, shp in _OBS_FIELDS:
                ptr = getattr(v, name)
                if not ptr:          # optional tensor that was not selected (select_outputs)
                    continue
                shape = shapes[shp]
                n = int(np.prod(shape))
                out[name] = _wrap_device(t, ptr, n, kind, self.device).view(*shape) if n > 0 else \
                    t.zeros(shape, dtype=t.float32 if kind == "f" else t.int32, device=f"cuda:{self.device}")
            self._views = out
        return self._views

    def obs(self) -> Dict[str, np.ndarray]:
        self._torch.cuda.synchronize(self.device)
        return {k: v.cpu().numpy() for k, v in self.obs_view().items()}

    def stats(self) -> np.ndarray:
        st = np.zeros(self.n_env, STATS_DTYPE)
        _check(self.lib, self.lib.rs_get_stats(self._h, st.ctypes.data))
        return st

    def vehicles(self, env: int = 0) -> Dict[str, np.ndarray]:
        arrs = {n: np.zeros(self.vcap, dt) for n, dt in _VEH_FIELDS}
        n = C.c_int32(0)
        _check(self.lib, self.lib.rs_dump_vehicles(self._h, env, C.byref(n),
                                                   *[arrs[k].ctypes.data for k, _ in _VEH_FIELDS]))
        return {k: v[:n.value] for k, v in arrs.items()}

    def phases(self, env: int = 0) -> np.ndarray:
        p = np.zeros(self.m.struct.n_tls, np.int32)
        _check(self.lib, self.lib.rs_get_phases(self._h, env, p.ctypes.data))
        return p

    def trip_records(self, env: int = 0) -> Dict[str, np.ndarray]:
        n = self.m.struct.n_trips
        out = dict(arrival=np.zeros(n, np.int32), depart=np.zeros(n, np.int32), time_loss=np.zeros(n, np.float32),
                   depart_delay=np.zeros(n, np.int32), waiting_time=np.zeros(n, np.float32))
        _check(self.lib, self.lib.rs_get_trip_records(self._h, env, out["arrival"].ctypes.data, out["depart"].ctypes.data,
                                                      out["time_loss"].ctypes.data, out["depart_delay"].ctypes.data,
                                                      out["waiting_time"].ctypes.data))
        return out

    def kernel_launches(self) -> int:
        return int(self.lib.rs_kernel_launches(self._h))

    def launch_shape(self) -> dict:
        v = [C.c_int32(0) for _ in range(5)]
        _check(self.lib, self.lib.rs_get_launch_shape(self._h, *[C.byref(x) for x in v]))
        return dict(threads_per_instance=v[0].value, instances_per_cta=v[1].value, grid_ctas=v[2].value,
                    smem_bytes_per_cta=v[3].value, tile_buffers=v[4].value)

    def last_step_ms(self) -> float:
        ms = C.c_float(0)
        _check(self.lib, self.lib.rs_last_step_ms(self._h, C.byref(ms)))
        return float(ms.value)


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap a borrowed device pointer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def _wrap_device(torch, ptr: int, n: int, kind: str, device: int):
    return torch.as_tensor(_DevArray(ptr, n, "<f4" if kind == "f" else "<i4"), device=f"cuda:{device}")
