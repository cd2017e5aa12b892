#!/usr/bin/env python
"""Where the adaptation transient (BASELINE configs[3]) spends its time: host wall-clock split of adaptTopology()
(dcg_get_info adapt_*_ms) over the first 20 steps after reset(), geometric and flow-driven score."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, make_ext, scene_params
d, M = 512, 524288
p = scene_params(d, solids=True)
for name, kw, pre in (("geometric", {}, 0), ("geometric_device", dict(selection=1), 0), ("flow", dict(score_mode=1), 0), ("flow_device", dict(score_mode=1, selection=1), 0),
                      ("flow_developed", dict(score_mode=1), 60), ("flow_developed_device", dict(score_mode=1, selection=1), 60)):
    for sync in (0,):
        sim = FluidSimulationDCGrid((d, d, d), M, p)
        if kw:
            sim.setExt(make_ext(**kw))
        sim.reset()
        if pre:
            sim.step(pre)
        if sync:
            sim.info("timing_sync")
        keys = ("adapt_move_ms", "adapt_refine_ms", "adapt_apron_ms", "adapt_layout_ms", "adapt_propagate_ms", "select_scores_ms", "select_d2h_ms", "select_host_ms")
        t0 = {k: sim.info(k) for k in keys}
        c0 = sim.counters().copy()
        w = time.time()
        sim.step(20)
        wall = (time.time() - w) * 1e3 / 20
        c = sim.counters()
        print(name, "sync" if sync else "async", "ms/step dev %.3f wall %.3f" % (sim.lastStepMs() / 20, wall), {k.split('_', 1)[1]: round((sim.info(k) - t0[k]) / 20, 3) for k in keys},
              "changed", int(c[1] - c0[1]), "moved", int(c[2] - c0[2]), "refined", int(c[3] - c0[3]), "host_sel", sim.info("host_selections"), "dev_sel", sim.info("device_selections"), "shortcut", sim.info("levels_shortcut"), flush=True)
        del sim
