import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, bronko_b200
from bronko_b200 import sim
c = bronko_b200.Bronko(0)
c.build_index(21, [sim.genome_path(sim.HPV16)])
r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.HPV16), 60, sim.SEED0)
print("calling", flush=True)
s = c.call_sample([(r1, o1), (r2, o2)], bronko_b200.CallArgs())
print("best", s.best_genome, "variants", len(s.variants), flush=True)
