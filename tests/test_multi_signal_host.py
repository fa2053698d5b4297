"""Host-side contracts of MultiSignal that the golden episodes do not exercise (CPU, oracle backend): the `gymma` list
forms EPyMARL uses (multi_signal.py:148-155,164-168,190-196), the metrics CSV (:218-226), the arguments we refuse."""
import os

import numpy as np
import pytest

import resco_b200.rewards as rewards
import resco_b200.states as states
from resco_b200.multi_signal import MultiSignal


def _env(gymma, log_dir=None, **kw):
    from pyoracle import OracleSim
    return MultiSignal("host", "cologne3", None, states.drq_norm, rewards.wait_norm, step_length=10, yellow_length=3,
                       log_dir=log_dir, gymma=gymma, backend=lambda m: OracleSim(m, 1, seed=0), seed=11, **kw)


def test_gymma_lists_are_the_dict_view_in_ts_order():
    a, b = _env(False), _env(True)
    oa, ob = a.reset(), b.reset()
    assert isinstance(ob, list) and len(ob) == len(a.ts_order) == b.n_agents
    for i, ts in enumerate(a.ts_order):
        np.testing.assert_array_equal(oa[ts], ob[i])
    ng = [len(a.phases[ts]) for ts in a.ts_order]
    for step in range(25):
        acts = [(step // 2 + i) % n for i, n in enumerate(ng)]
        ra = a.step({ts: acts[i] for i, ts in enumerate(a.ts_order)})
        rb = b.step(acts)                                   # EPyMARL passes a list in ts_order
        assert isinstance(rb[0], list) and isinstance(rb[1], list) and rb[2] == [ra[2]] and rb[3] == ra[3]
        for i, ts in enumerate(a.ts_order):
            np.testing.assert_array_equal(ra[0][ts], rb[0][i])
            assert float(ra[1][ts]) == float(rb[1][i])
    assert [sp.n for sp in b.action_space] == ng
    assert [tuple(sp.shape) for sp in b.observation_space] == [tuple(a.obs_shape[ts]) for ts in a.ts_order]
    a.close(); b.close()


def test_metrics_csv_lines_parse_like_the_reference_reader(tmp_path):
    env = _env(False, log_dir=str(tmp_path))
    env.reset()
    ng = [len(env.phases[ts]) for ts in env.ts_order]
    for step in range(6):
        env.step({ts: step % ng[i] for i, ts in enumerate(env.ts_order)})
    env.metrics_saved = list(env.metrics)
    env.reset()                                             # the reference saves metrics_<run>.csv on the next reset
    path = os.path.join(str(tmp_path), env.connection_name, "metrics_1.csv")
    lines = open(path).read().splitlines()
    assert len(lines) == 6
    for n, line in enumerate(lines):
        assert float(line.split(",")[0]) == env._begin + 10 * (n + 1)
        # the reference's reader (utils/readCSV.py:31-40): third '}'-separated field = queue_lengths, ':'-split per signal
        queues = line.split('}')[2]
        signals = queues.split(':')
        got = [int(sig.split(',')[0]) for sig in signals[1:]]
        assert got == [env.metrics_saved[n]['queue_lengths'][ts] for ts in env.ts_order]
    env.close()


def test_unsupported_arguments_are_refused_loudly():
    with pytest.raises(NotImplementedError):
        _env(False, warmup=5)
    with pytest.raises(NotImplementedError):
        _env(False, step_ratio=4)
    with pytest.raises(ValueError):
        _env(False, lights=["not-a-signal"])
    with pytest.raises(FileNotFoundError):
        MultiSignal("host", "no_such_map", None, states.mplight, rewards.wait, log_dir=None)
