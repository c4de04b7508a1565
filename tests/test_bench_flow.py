"""bench.py's control flow on a CPU box: the CUDA context and torch.cuda are replaced by stand-ins, so that a slip in
the timing / JSON plumbing is caught here and not by the one GPU run at the end of a round.  (The numbers mean
nothing; the GPU run is the measurement.)"""
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeSample:
    best_genome = 0
    variants = np.zeros(30, dtype=[("pos", "<u4")])


class FakeBronko:
    def __init__(self, device=0):
        self.n = 0

    def build_index(self, k, paths): pass
    def share_index(self, other): pass
    def begin(self, args=None): pass
    def push_device(self, *a): pass
    def push_ptr(self, *a): pass
    def push_packed_ptr(self, *a): pass
    def close(self): pass
    def set_stage_timing(self, on): pass

    def finish(self):
        self.n += 1
        return FakeSample()

    def stage_times(self):
        return {"scan_ms": 0.2, "leftover_ms": 0.1, "finalize_ms": 0.5, "map_ms": 0.4, "score_ms": 0.5, "total_ms": 1.7,
                "launches": 33, "scan_launches": 2, "coll_ms": 0.0, "coll_calls": 0, "decode_ms": 0.0}


class FakeStream:
    def __init__(self, *a, **k): pass
    def synchronize(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


class FakeEvent:
    def __init__(self, enable_timing=True): pass
    def record(self): pass
    def elapsed_time(self, other): return 12.5


def test_bench_prints_one_json_line_with_every_contract_key(monkeypatch, capfd):
    import torch
    import bronko_b200
    from bronko_b200 import sim
    monkeypatch.setattr(bronko_b200, "Bronko", FakeBronko)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: s)
    tiny = sim.simulate_pairs(sim.load_genome(sim.HPV16), 20, sim.SEED0)
    monkeypatch.setattr(sim, "simulate_pairs", lambda *a, **k: tiny)
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    monkeypatch.setattr(b, "oracle_step", lambda *a, **k: None)
    # (the FASTQ and sharded legs need the real library: the GPU run covers them)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "8", "--warmup", "3", "--no-fastq", "--no-sharded"])
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    saved = os.dup(1)
    try:
        b.main()
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    out = capfd.readouterr().out.strip().splitlines()
    lines = [ln for ln in out if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks",
                "h2d_ceiling", "roofline_path", "fastq", "sharded"):
        assert key in d, key
    assert d["steps"] == 8 and d["n_gpus"] == 1 and d["gpu_launches"] == 8 * 33
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert "workload" in d["config"] and "window" in d["clocks"]
