from .compiler import Scenario, compile_scenario, green_phase_indices, MOVEMENTS  # noqa: F401
