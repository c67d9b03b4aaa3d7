"""CPU check of the barrier protocol of the channel-streamed persistent kernel (conv_cs.cu) on the randomised model in
sim_cs_protocol.py: no deadlock, no parity aliasing, no ring stage overwritten under a pending MMA, no TMEM half reused
before the epilogue drained it -- for the launch shapes of the bench (one or two TMEM halves, 1..8 row tiles over four
issuers, single-chunk units) -- and the model does catch deliberately broken protocols."""
import pytest

from tests import sim_cs_protocol as sim


@pytest.mark.parametrize("cfg", sim.CONFIGS)
def test_protocol_has_no_deadlock_or_hazard(cfg):
    for seed in range(5):
        status, detail, errors = sim.run(*cfg, seed)
        assert status == "OK", (cfg, seed, status, detail)
        assert not errors


@pytest.mark.parametrize("bug", ["no_aempty", "no_tempty"])
def test_model_catches_broken_protocols(bug):
    caught = 0
    for seed in range(12):
        status, detail, _ = sim.run(6, 6, 3, 2, 4, 2, 3, seed, bug=bug)
        caught += status == "HAZARD"
    assert caught >= 4, f"{bug}: only {caught} of 12 seeds flagged a hazard"
