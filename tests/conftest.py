import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
FONT = os.path.join(GOLDEN, "Aileron-Regular.ttf")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def sha16(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()[:16]


@pytest.fixture(scope="session")
def built():
    """Build the product library and the CPU checker once (idempotent, seconds)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def O(built):
    from oracle import oracle

    return oracle


@pytest.fixture(scope="session")
def port(O):
    return O.Port()


@pytest.fixture(scope="session")
def ref(O):
    r = O.Ref()
    if not (r.have_sws and r.have_ft):
        pytest.skip("bundled libswscale / FreeType not present in this image")
    return r


@pytest.fixture(scope="session")
def glyphs(O):
    return O.GlyphTable.load(os.path.join(GOLDEN, "glyphs_aileron20.npz"))


@pytest.fixture(scope="session")
def N(built):
    import ngp_encode_server_b200 as n

    return n


@pytest.fixture(scope="session")
def session(N, glyphs):
    """One GPU session with the golden glyph atlas loaded (gpu tests only)."""
    s = N.Session(device=0, max_width=7680, max_height=4320, max_sources=4)
    s.atlas_set(glyphs.metrics, glyphs.bitmaps)
    yield s
    s.close()
