"""CPU check of the NVLink seam-exchange protocol of the multi-GPU driver (seg/p2p.py) on the randomised model in
sim_seam_protocol.py: DATA / ACK counters, the tail and push-done events, buffers reused from volume to volume."""
import pytest

from tests import sim_seam_protocol as sim


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_seam_protocol_has_no_deadlock_or_hazard(world):
    for seed in range(20):
        status, detail = sim.run(world, 12, seed)
        assert status == "OK", (world, seed, status, detail)


@pytest.mark.parametrize("bug", ["no_ack", "no_begin"])
def test_model_catches_broken_protocols(bug):
    caught = sum(sim.run(4, 12, seed, bug=bug)[0] == "HAZARD" for seed in range(12))
    assert caught >= 6, f"{bug}: only {caught} of 12 seeds flagged a hazard"
