import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = ["cornell", "cornellSphere", "monkeys", "terrain64", "tiny", "soup"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """name -> (FlatScene, dict of golden arrays) for the committed fixtures."""
    from cudart_b200.scenefile import HIT_DTYPE, RAY_DTYPE, TRIANGLE_DTYPE, load_crts

    cache = {}

    def get(name):
        if name not in cache:
            fs = load_crts(os.path.join(GOLDEN, name + ".crts.gz"))
            z = np.load(os.path.join(GOLDEN, name + ".golden.npz"))
            g = {
                "pretris": z["pretris"].view(TRIANGLE_DTYPE),
                "rays": z["rays"].view(RAY_DTYPE),
                "hits": z["hits"].view(HIT_DTYPE),
                "positions": z["positions"], "premats": z["premats"],
                "accum": z["accum"],
                "rgba": z["rgba"],
                "accum_hi": z["accum_hi"], "hi": tuple(int(v) for v in z["meta"][4:7]),
                "W": int(z["meta"][0]), "H": int(z["meta"][1]), "spp": int(z["meta"][2]), "seed": int(z["meta"][3]),
            }
            cache[name] = (fs, g)
        return cache[name]

    return get
