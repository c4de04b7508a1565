import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def hpv_bkdb_bytes():
    """test_data/hpv.bkdb of the reference (bundled build output, k=21), stored xz-compressed."""
    with lzma.open(os.path.join(ROOT, "tests", "golden", "hpv.bkdb.xz")) as f:
        return f.read()


@pytest.fixture(scope="session")
def hpv_bkdb_path(hpv_bkdb_bytes, tmp_path_factory):
    p = tmp_path_factory.mktemp("golden") / "hpv.bkdb"
    p.write_bytes(hpv_bkdb_bytes)
    return str(p)


@pytest.fixture(scope="session")
def sars_paths():
    from bronko_b200 import sim
    return [sim.genome_path(n) for n in sim.SARS4]


@pytest.fixture(scope="session")
def hpv_fasta():
    from bronko_b200 import sim
    return sim.genome_path(sim.HPV16)
