"""resco_b200 -- B200-native vectorised traffic-microsimulation backend behind RESCO's MultiSignal surface."""
__version__ = "0.1.0"

EPYMARL_MAPS = ['grid4x4', 'arterial4x4', 'cologne1', 'cologne3', 'cologne8', 'ingolstadt1', 'ingolstadt7', 'ingolstadt21']
EPYMARL_ALGS = ['ia2c', 'ippo', 'maa2c', 'mappo', 'coma', 'iql', 'maddpg', 'qmix', 'vdn', 'ia2c_ns', 'ippo_ns', 'maa2c_ns',
                'mappo_ns', 'coma_ns', 'iql_ns', 'maddpg_ns', 'qmix_ns', 'vdn_ns']


def register_epymarl(register=None, log_dir=None):
    """The reference's EPyMARL registration (resco_benchmark/__init__.py:16-61) with this backend as the entry point:
    the same ids (``<map>-<alg>-v<trial>``) and the same constructor kwargs (drq_norm / wait_norm, 10 s steps, 4 s
    yellow, ``gymma=True``), so EPyMARL's ``gymma`` wrapper finds the environments it expects.  `register` defaults to
    ``gym.envs.registration.register`` (gym is optional: pass a callable to use another registry).  Returns the ids."""
    import os
    from . import rewards, states
    from .multi_signal import load_scenario
    if register is None:
        from gym.envs.registration import register          # noqa: F811  (raises if gym is not installed)
    log_dir = os.getcwd() if log_dir is None else log_dir
    ids = []
    for m in EPYMARL_MAPS:
        mc = load_scenario(m).meta["map_config"]
        for alg in EPYMARL_ALGS:
            for trial in range(1, 30):
                env_id = m + "-" + alg + "-v" + str(trial)
                register(id=env_id, entry_point="resco_b200.multi_signal:MultiSignal",
                         kwargs={'run_name': alg + '-tr' + str(trial), 'map_name': m, 'net': mc.get('net'),
                                 'state_fn': states.drq_norm, 'reward_fn': rewards.wait_norm, 'route': mc.get('route'),
                                 'gui': False, 'end_time': mc['end_time'], 'step_length': 10, 'yellow_length': 4,
                                 'step_ratio': 1, 'max_distance': 200, 'lights': (), 'log_dir': log_dir,
                                 'libsumo': False, 'warmup': 0, 'gymma': True})
                ids.append(env_id)
    return ids
