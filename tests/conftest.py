import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The pinned CPU oracle: the reference's own classes when oracle/_ref was built, else the C port."""
    import oracle_lib
    kind = "ref" if oracle_lib.available("ref") else "port"
    if not oracle_lib.available(kind):
        pytest.skip("no oracle library built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
    return oracle_lib.RefLib(kind)


@pytest.fixture(scope="session")
def port():
    import oracle_lib
    if not oracle_lib.available("port"):
        pytest.skip("oracle/liboracle_port.so not built")
    return oracle_lib.RefLib("port")
