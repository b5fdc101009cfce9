import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from sph_sm_monodomain_b200 import Sim
from tests.common import load_golden, setup_from_golden
g, kw = load_golden("cfg1_4944")
sim = Sim(**kw)
setup_from_golden(sim, g, False)
sim.Animation(1)
p = sim.particles()
print("ok", np.abs(p["pos"] - g["step1.pos"]).max())
