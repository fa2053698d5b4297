"""Episode outputs the reference's post-hoc utilities read.

* ``write_tripinfo`` -- ``tripinfo_<run>.xml`` as SUMO writes it with ``--tripinfo-output
  --tripinfo-output.write-unfinished`` (multi_signal.py:127-129): one ``<tripinfo>`` per departed vehicle,
  unfinished trips with ``arrival="-1.00"`` and their running duration / timeLoss.
* ``avg_delay_from_tripinfo`` -- the per-episode average the reference publishes in
  ``utils/avg_timeLoss.py``, restating ``utils/readXML.py:27-77``: mean over tripinfo entries of
  ``timeLoss + departDelay``; for ``<vehicle>``-type demand (grid4x4 / arterial4x4 route files) vehicles
  whose scheduled departure lies after that of the last vehicle that did depart are charged
  ``end_time - depart`` and counted as trips (``<trip>``-type demand contributes none: readXML.py:62
  skips every child whose tag is not 'vehicle').
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from typing import Dict, Optional

import numpy as np


def write_tripinfo(path: str, scenario, records: Dict[str, np.ndarray], running: Dict[str, np.ndarray], now_tick: int) -> int:
    """records: VecSim.trip_records(env); running: VecSim.vehicles(env).  Returns the number of entries."""
    begin = float(scenario.meta["begin"])
    ids = scenario.meta["trip_ids"]
    vt_ids = scenario.meta["vtype_ids"]
    vt_of = scenario.arrays["trip_vtype"]
    n = 0
    with open(path, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<tripinfos>\n')
        arrived = np.nonzero(records["arrival"] >= 0)[0]
        order = arrived[np.argsort(records["arrival"][arrived], kind="stable")]
        for i in order:
            dep = begin + float(records["depart"][i])
            arr = begin + float(records["arrival"][i])
            f.write(f'    <tripinfo id="{ids[i]}" depart="{dep:.2f}" departDelay="{float(records["depart_delay"][i]):.2f}" '
                    f'arrival="{arr:.2f}" duration="{arr - dep:.2f}" timeLoss="{float(records["time_loss"][i]):.2f}" '
                    f'waitingTime="{float(records["waiting_time"][i]):.2f}" vType="{vt_ids[int(vt_of[i])]}"/>\n')
            n += 1
        for k in range(len(running["vid"])):          # --tripinfo-output.write-unfinished
            i = int(running["vid"][k])
            dep = begin + float(running["depart"][k])
            ddelay = float(running["depart"][k]) - float(scenario.arrays["trip_depart"][i])
            f.write(f'    <tripinfo id="{ids[i]}" depart="{dep:.2f}" departDelay="{ddelay:.2f}" arrival="-1.00" '
                    f'duration="{begin + now_tick - dep:.2f}" timeLoss="{float(running["tloss"][k]):.2f}" '
                    f'waitingTime="{float(running["acc_wait"][k]):.2f}" vType="{vt_ids[int(vt_of[i])]}"/>\n')
            n += 1
        f.write('</tripinfos>\n')
    return n


def episode_trip_range(scenario, episode: int = 0):
    """(first, last + 1) of the trip table rows that are the demand of `episode` (the route file of that run)."""
    bank = scenario.arrays.get("bank_origin_off")
    if bank is None:
        return 0, len(scenario.arrays["trip_depart"])
    e = episode % bank.shape[0]
    return int(bank[e, 0]), int(bank[e, -1])


def avg_delay_from_tripinfo(path: str, scenario=None, end_time: Optional[float] = None,
                            vehicle_demand: bool = False, metric: str = "timeLoss", episode: int = 0) -> float:
    """readXML.py:27-77 for one tripinfo file.  `vehicle_demand`: the route file holds <vehicle> elements
    (grid4x4 / arterial4x4), so never-departed vehicles are charged; needs `scenario` + `end_time` (+ `episode`: which
    of the compiled route files the run used)."""
    root = ET.parse(path).getroot()
    num_trips, total = 0, 0.0
    last_departure_time, last_depart_id = 0.0, ''
    for child in root:
        num_trips += 1
        total += float(child.attrib[metric])
        if metric == 'timeLoss':
            total += float(child.attrib['departDelay'])
            depart_time = float(child.attrib['depart'])
            if depart_time > last_departure_time:
                last_departure_time = depart_time
                last_depart_id = child.attrib['id']
    if metric == 'timeLoss' and vehicle_demand:
        begin = float(scenario.meta["begin"])
        t0, t1 = episode_trip_range(scenario, episode)
        ids = scenario.meta["trip_ids"][t0:t1]
        sched = begin + scenario.arrays["trip_depart"][t0:t1].astype(np.float64)
        if last_depart_id not in ids:
            raise ValueError('Wrong trip file')
        last_sched = float(sched[ids.index(last_depart_id)])
        never = sched[sched > last_sched]
        total += float(np.sum(float(end_time) - never))
        num_trips += int(len(never))
    return total / num_trips
