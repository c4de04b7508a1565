"""tools/deep_repro.py DEPTH CHUNKS — one deep sample pushed in chunks through an unsharded context (developer tool)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import bronko_b200
from bronko_b200 import sim
depth, n_chunks = float(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
g = sim.load_genome(sim.SARS4[0])
pairs = int(round(depth / n_chunks * len(g) / 300))
plan = sim.plant_for(g, 1)
c = bronko_b200.Bronko(0)
c.build_index(21, [sim.genome_path(n) for n in sim.SARS4])
for rep in range(2):
    c.begin(bronko_b200.CallArgs())
    for k in range(n_chunks):
        r1, r2, off = sim.simulate_pairs_torch(g, pairs, 1000 + k, dev, plan)
        torch.cuda.current_stream().synchronize()
        c.push_device(0, r1.data_ptr(), off.data_ptr(), pairs, pairs * 150, 150)
        c.push_device(1, r2.data_ptr(), off.data_ptr(), pairs, pairs * 150, 150)
        for slot in (0, 1):
            torch.cuda.ExternalStream(int(c._lib.bk_stream_slot(c.h, slot))).synchronize()
        del r1, r2, off
    s = c.finish()
    print("rep", rep, "variants", len(s.variants), s.kmc_stats(0), c.stage_times()["total_ms"])
