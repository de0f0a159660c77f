"""The counter-based random stream (DESIGN.md "Random stream"): Philox4x32-10 known answers (Random123
kat_vectors), agreement of the three independent implementations that exist on the CPU side (pure Python in
tests/golden/ref_recorder.py, C in the oracle, host C++ in libecmc_b200.so) and CPython's derived draws."""
import random

import numpy as np

import ref_recorder as rr
from jellyfysh_b200 import abi


def test_philox_known_answers():
    assert rr.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert rr.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert rr.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_oracle_stream_matches_python(oracle):
    for seed, stream, event, slot in [(0, 0, 0, 0), (7, 3, 12345, abi.slot(abi.SLOT_PAIR_TIME, 77)),
                                      (0xdeadbeef, 4095, (1 << 40) + 17, abi.slot(abi.SLOT_VETO_CHOICE))]:
        words = oracle.random_words(seed, stream, event, slot, 0, 13)
        doubles = oracle.random_doubles(seed, stream, event, slot, 0, 9)
        assert [int(w) for w in words] == [rr.stream_word(seed, stream, event, slot, i) for i in range(13)]
        assert [float(d) for d in doubles] == [rr.stream_double(seed, stream, event, slot, i) for i in range(9)]
        assert np.all((doubles >= 0.0) & (doubles < 1.0))


def test_slot_random_uses_cpython_derivations():
    """expovariate / uniform / choice / randint of the recorder are CPython's own algorithms on the stream."""
    rng = rr.SlotRandom(5, 9)
    rng.set_context(3, rr.make_slot(rr.SLOT_VETO_TIME), rr.make_slot(rr.SLOT_VETO_CHOICE))
    u0 = rr.stream_double(5, 9, 3, rr.make_slot(rr.SLOT_VETO_TIME), 0)
    u1 = rr.stream_double(5, 9, 3, rr.make_slot(rr.SLOT_VETO_TIME), 1)
    import math
    assert rng.uniform(0.0, 2.5) == 0.0 + 2.5 * u0
    assert rng.expovariate(2.0) == -math.log(1.0 - u1) / 2.0
    n, k, i = 37, 6, 0
    while True:
        r = rr.stream_word(5, 9, 3, rr.make_slot(rr.SLOT_VETO_CHOICE), i) >> (32 - k)
        i += 1
        if r < n:
            break
    assert rng.choice(list(range(n))) == r
    assert isinstance(rng, random.Random)
