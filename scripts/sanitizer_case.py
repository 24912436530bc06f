import sys, os
sys.path[:0]=['.','tests','oracle']
import numpy as np, torch
from helpers import stream_config
from stream_sim_b200.engine import StarEngine
from stream_sim_b200.frames import GreatCircleFrame
from stream_sim_b200.model import StreamModel
from stream_sim_b200.surveys import Survey
from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
for name in ("pal5","toy1"):
    eng=StarEngine(stream=StreamModel(stream_config(name)), survey=Survey.synthetic(64,128), frame=GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS), hist=True, perfect_galstarsep=True, grid_table=(name=="pal5"))
    c=eng.new_counters()
    out=eng.generate_observe(200003, seed=3, counters=c, pix=True)
    cat={k:out[k] for k in ("ra","dec","mag_g","mag_r")}
    o2=eng.observe(cat, seed=3)
    o3=eng.rotate(out["phi1"], out["phi2"])
    torch.cuda.synchronize()
    print(name, eng.summarize(c)["n_observed"], int(o2["flag_observed"].sum()))
h=eng.allocate_host(100000, list(out.keys()))
eng.generate_observe_host(100000, h, eng.host_workspace(1<<15), 1<<15, seed=3)
print("host ok", int(h["counters"][0]))
