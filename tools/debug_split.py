import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import astronomix_b200 as ax
from astronomix_b200 import problems as pr, stepping as ti, option_classes as oc
from oracle import fv_oracle as fo
np.set_printoptions(linewidth=200, precision=6)
for fof in (True, False):
    s = pr.smooth_random(1, 16, seed=8, dtype=np.float64)
    config, params, rv = pr.make(1, 16, dtype=np.float64, split=oc.SPLIT, time_integrator=oc.MUSCL,
                                 first_order_fallback=fof, params=ax.SimulationParams(dt_max=1.0))
    o = fo.Opts.from_config(config, params)
    ref, info = fo.time_integration(s, o, max_steps=1, return_info=True)
    out = ti.integrate_steps(s, config, params, rv, 1)
    print("first_order", fof, "dt ref", info["dts"], "gpu", ti.last_info.last_dt)
    print("in ", s)
    print("ref", ref)
    print("out", out)
