"""tools/noise_probe.py — score-stage time and chain statistics (BK_NOISE_DEBUG=1) at several depths (developer tool)."""
import os, sys
os.environ["BK_NOISE_DEBUG"] = "1"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bronko_b200
from bronko_b200 import sim
c = bronko_b200.Bronko(0)
c.build_index(21, [sim.genome_path(n) for n in sim.SARS4])
for depth in (400, 2000, 2500, 5000, 10000):
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), depth, sim.SEED0 + depth)
    for rep in range(2):
        s = c.call_sample([(r1, o1), (r2, o2)])
    t = c.stage_times()
    print("SARS %6dx: score %.3f ms, total %.3f ms, variants %d" % (depth, t["score_ms"], t["total_ms"], len(s.variants)), flush=True)
h = bronko_b200.Bronko(0)
h.build_index(21, [sim.genome_path(sim.HPV16)])
r1, o1, r2, o2, _ = sim.config_reads("C1")
for rep in range(2):
    s = h.call_sample([(r1, o1), (r2, o2)])
t = h.stage_times()
print("C1 HPV16 5000x: score %.3f ms, total %.3f ms, variants %d" % (t["score_ms"], t["total_ms"], len(s.variants)))
