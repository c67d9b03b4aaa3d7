"""CPU check of the row-sweep kernel's barrier protocol (conv_rs.cu) on the randomised model in sim_rs_protocol.py:
no deadlock, no parity aliasing, for single- and multi-unit CTAs, 3 and 4 lane quarters, ring depths 3 and 4."""
import pytest

from tests import sim_rs_protocol as sim


@pytest.mark.parametrize("cfg", [(6, 3, 4, 3, 1), (6, 3, 4, 3, 2), (8, 9, 4, 3, 3), (5, 6, 3, 4, 2), (7, 12, 4, 3, 2)])
def test_protocol_has_no_deadlock_or_aliasing(cfg):
    for seed in range(6):
        status, detail, errors = sim.run(*cfg, seed, ni=1 + seed % 2)  # one issuer warp (default) and two
        assert status == "OK", (cfg, seed, status, detail)
        assert not errors
