import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # a fresh checkout has no built artefacts (*.so are git-ignored): build them once (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "tfhe-omr_b200", "lib", "libomr_b200.so")
    orc = os.path.join(ROOT, "oracle", "libomr_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def keypack():
    """Real keys from the oracle keygen (recipient) — test infrastructure."""
    import oracle
    return oracle.KeyPack(seed=0x4F4D520001)


@pytest.fixture(scope="session")
def decoy():
    import oracle
    return oracle.KeyPack(seed=0x4F4D520002, sender_only=True)


@pytest.fixture(scope="session")
def detector(keypack):
    import tfhe_omr_b200 as omr
    dk = omr.DetectionKey(keypack.bsk1, keypack.ksk, keypack.bsk2, keypack.trk)
    return omr.Detector(dk, device=0)
