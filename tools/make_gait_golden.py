#!/usr/bin/env python
"""Golden joint-target sequences from the reference's OWN scripted gait engine.

`nikengine.engine.EngineNode.update(lin, ang, 'awake', 'walk')` (reference nikengine/engine.py:679-701) is pure
numpy, so unlike MuJoCo it CAN be imported in the build container.  This script runs it exactly the way the reference's
keyboard player does (custom_play.py:49-74: ENGINE_FPS = 1/(timestep*decimation), STAND_HEIGHT = 0.2, engine clock set
from simulation time, per-step joint-target rate limit of 0.08 rad) over a fixed command schedule and stores the
resulting joint targets as a fixture.  The reference tree is only read here; tests and the GPU box use the committed
`tests/golden/nikengine_gait_targets.npz`.

    python tools/make_gait_golden.py [/root/reference]
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 0.008 * 2                       # mjmodel.xml timestep * decimation (custom_play.py:17,52)
RATE = 0.08                          # custom_play.py:18
# (steps, lin_speed, ang_speed): wake up + stand (the engine needs ~200 steps to rise), walk forward, turn left while walking, walk backward, stand
SCHEDULE = ((300, 0.0, 0.0), (200, 0.05, 0.0), (160, 0.05, 0.2), (160, -0.05, 0.0), (40, 0.0, 0.0))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    sys.path.insert(0, ref)
    with contextlib.redirect_stdout(io.StringIO()):
        from nikengine.engine import EngineNode, config, set_time_s
        eng = EngineNode()
        eng.update(0.0, 0.0, "idle")
        config.ENGINE_FPS = 1.0 / DT
        config.STAND_HEIGHT = 0.2
        raw, lim, cmd = [], [], []
        prev = np.zeros(18)
        t = 0
        for steps, lin, ang in SCHEDULE:
            for _ in range(steps):
                set_time_s(t * DT)
                a = np.asarray(eng.update(lin, ang, "awake", "walk"), dtype=np.float64)
                prev = prev + np.clip(a - prev, -RATE, RATE)
                raw.append(a)
                lim.append(prev.copy())
                cmd.append((lin, 0.0, ang))
                t += 1
    out = os.path.join(ROOT, "tests", "golden", "nikengine_gait_targets.npz")
    np.savez_compressed(out, raw=np.array(raw), targets=np.array(lim), commands=np.array(cmd), dt=DT, rate=RATE)
    print(f"wrote {out}: {len(raw)} steps, |target| max {np.abs(lim).max():.3f}")


if __name__ == "__main__":
    main()
