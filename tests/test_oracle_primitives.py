"""Known-answer and property tests pinning the oracle's building blocks."""
import ctypes as C

import numpy as np

import util


def test_philox_known_answers(oracle_lib):
    """Philox4x32-10 KATs from the Random123 distribution (kat_vectors)."""
    def run(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        oracle_lib.orc_philox(c, key[0], key[1])
        return [int(x) for x in c]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_stop_speed_never_overshoots(oracle_lib):
    """maximumSafeStopSpeed: braking at `decel` from the returned speed (after the reaction step) stops
    within the gap -- iterate the Euler update and check the travelled distance."""
    for decel in (4.5, 4.0, 10.0):
        for gap in np.linspace(0.0, 120.0, 241):
            v = oracle_lib.orc_max_safe_stop_speed(float(gap), decel, 1.0)
            assert v >= 0
            x, g = 0.0, float(gap)
            for _ in range(200):           # keep following the rule: each step re-evaluates the remaining gap
                x += v
                if v <= 0:
                    break
                v_next = oracle_lib.orc_max_safe_stop_speed(g - x, decel, 1.0)
                assert v_next >= v - decel - 1e-3, (gap, v, v_next)    # never needs more than `decel`
                v = v_next
            assert x <= gap + 1e-3, (gap, x)


def test_follow_speed_monotone(oracle_lib):
    last = -1.0
    for gap in np.linspace(0, 80, 161):
        v = oracle_lib.orc_follow_speed(float(gap), 5.0, 4.5, 4.5, 1.0)
        assert v >= last - 1e-5
        last = v
    # standing leader, zero gap -> must stop
    assert oracle_lib.orc_follow_speed(0.0, 0.0, 4.5, 4.5, 1.0) == 0.0


def test_brake_gap_matches_closed_form(oracle_lib):
    for v in np.linspace(0, 30, 61):
        steps = int(v / 4.5)
        expect = steps * v - 4.5 * steps * (steps + 1) / 2 + v * 1.0
        assert abs(oracle_lib.orc_brake_gap(float(v), 4.5, 1.0) - expect) < 1e-3


def test_free_speed_reaches_target(oracle_lib):
    # far away: unconstrained (large); at the line: exactly the target speed
    assert oracle_lib.orc_free_speed(4.5, 0.0, 8.0) == 8.0
    assert oracle_lib.orc_free_speed(4.5, 500.0, 8.0) > 30.0
    last = 0
    for d in np.linspace(0, 100, 101):
        v = oracle_lib.orc_free_speed(4.5, float(d), 8.0)
        assert v >= 8.0 and v >= last - 1e-4
        last = v


def test_create_yellows_examples():
    from resco_b200.abi import create_yellows
    ys, yd = create_yellows(["GGrr", "rrGG"])
    assert ys == ["yyrr", "rryy"] and yd == {"0_1": 2, "1_0": 3}
    # 'g' that stays green needs no yellow; identical phases need none at all
    ys, yd = create_yellows(["Ggrr", "Ggrr"])
    assert ys == [] and yd == {}
    ys, yd = create_yellows(["GgGr", "Ggrs"])
    assert ys == ["Ggyr"] and yd == {"0_1": 2}
