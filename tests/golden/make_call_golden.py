#!/usr/bin/env python
"""Regenerates tests/golden/hpv16_call_1500x.json: the observables of `bronko call` for one seeded HPV16 sample
(bundled hpv.bkdb, 150 bp PE at 1,500x, sim.SEED0 + 5) as the ORACLE computes them.  The reference has no expected
outputs for `call` (parity unpinned), so this file pins the oracle against drift — not against the
reference; the GPU parity tests compare the CUDA path with the oracle on the same inputs.

    python tests/golden/make_call_golden.py
"""
import hashlib
import json
import lzma
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

OUT = os.path.join(HERE, "hpv16_call_1500x.json")
DEPTH, SEED_OFFSET, READS_NAME = 1500, 5, "reads/rep1_R1.fastq.gz"


def sha(b):
    return hashlib.sha256(b).hexdigest()


def observables():
    import bronko_b200
    from bronko_b200 import sim
    from oracle import oracle as O
    from util import oracle_sample
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "hpv.bkdb")
        with lzma.open(os.path.join(HERE, "hpv.bkdb.xz")) as f, open(p, "wb") as g:
            g.write(f.read())
        oi = O.Index.load(p)
        r1, o1, r2, o2, truth = sim.simulate_pairs(sim.load_genome(sim.HPV16), DEPTH, sim.SEED0 + SEED_OFFSET)
        counts, s = oracle_sample(oi, [(r1, o1), (r2, o2)], bronko_b200.CallArgs())
        obs = {
            "workload": "HPV16.fa, hpv.bkdb (k=21), 150 bp PE at %dx, seed SEED0+%d, default CallArgs" % (DEPTH, SEED_OFFSET),
            "reads_sha256": [sha(r1.tobytes()), sha(r2.tobytes())],
            "kmc_stats": [list(map(int, c.stats())) for c in counts],
            "kmers_sha256": [sha(c.get()[0].tobytes() + c.get()[1].tobytes()) for c in counts],
            "tallies": [np.asarray(s.stats(f)).astype(np.int64).tolist() for f in range(2)],
            "best_genome": int(s.best),
            "pileup_sha256": sha(np.ascontiguousarray(s.pileup()).tobytes()),
            "noise_max_sha256": sha(np.ascontiguousarray(s.noise_max()).tobytes()),
            "summary": [int(s.summary()[0]), int(s.summary()[1]), repr(float(s.summary()[2])), repr(float(s.summary()[3]))],
            "unmapped": int(s.unmapped()),
            "pileup_tsv_sha256": sha(s.pileup_text().encode()),
            "vcf": s.vcf_text(READS_NAME),
            "planted": {"pos": truth["pos"].tolist(), "alt": truth["alt"].tolist(), "af": truth["af"].tolist()},
        }
    return obs


if __name__ == "__main__":
    with open(OUT, "w") as f:
        json.dump(observables(), f, indent=1)
        f.write("\n")
    print("wrote", OUT)
