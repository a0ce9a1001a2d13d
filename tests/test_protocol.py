"""Model-check the fused-unit kernel's synchronisation protocol (tests/protocol_sim.py): random interleavings of the
loader / producer / issuer / epiA / epiB roles for every plan family tc3_plan emits — no deadlock, no mbarrier phase
aliasing, no read of stale or overwritten tiles, ring slots or accumulator sets.  No GPU needed."""
import pytest

from protocol_sim import PLANS, ProtocolError, Sim


@pytest.mark.parametrize("plan", list(PLANS))
def test_fused_unit_protocol_is_hazard_free(plan):
    kw = PLANS[plan]
    for n_tiles in (1, 2, 3, 4, 5, 8):
        for K in (1, 3, 5):
            for n_issuers in ((1, 2, 3) if not kw["resident"] else (1, 2, 4)):
                for seed in range(6):
                    Sim(n_tiles=n_tiles, K=K, n_issuers=n_issuers, seed=seed, **kw).run()


def _loader_without_buffer_free_wait(s):
    """A broken loader: refills its A buffer without waiting for the issuers' buffer-free commit."""
    for it in range(s.n):
        b = it % s.a1
        if s.A1_readers[b] != s.ni:
            raise ProtocolError("loader overwrites a busy buffer")
        s.A1[b], s.A1_readers[b] = ("x", it), 0
        s.a1_full[b].arrive()
        yield


def test_the_model_catches_a_broken_protocol():
    """Sanity of the checker itself: a loader that ignores the buffer-free barrier must be reported (in ping-pong mode the
    buffer holds h until conv2 of the tile has read it)."""
    with pytest.raises(ProtocolError):
        for seed in range(20):
            s = Sim(n_tiles=6, K=3, n_issuers=2, seed=seed, **PLANS["ping-pong tiles, streamed weights"])
            s.loader = lambda s=s: _loader_without_buffer_free_wait(s)
            s.run()
