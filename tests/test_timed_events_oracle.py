"""ctx_set_event_block on the CPU oracle: a store stamped with block b of a K-block call gives exactly what the reference's host
gets by calling process for b blocks, storing, and calling again (processor.rs:214 polls per block) — checked by doing both."""
import pytest

import timed_scenarios as ts
from helpers import assert_bit_exact


@pytest.mark.parametrize("scenario,kw", [(ts.gain_pan_timed, dict(bus=False)), (ts.gain_pan_timed, dict(bus=True)), (ts.sampler_timed, {}), (ts.filters_timed, {})])
def test_stamped_stores_equal_hand_split_calls(oracle, scenario, kw):
    a, b = scenario(oracle, timed=True, **kw), scenario(oracle, timed=False, **kw)
    assert len(a) == len(b)
    for i, ((ya, ma), (yb, mb)) in enumerate(zip(a, b)):
        assert_bit_exact(ya, yb, f"{scenario.__name__} call {i}")
        assert ma == mb
