"""Pins the Philox4x32-10 generator behind the action stream against the Random123
known-answer vectors (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
Random123 kat_vectors, philox4x32 10 rounds), and the mask -> action rule."""
from oracle import lle_oracle as lo


def test_philox_known_answers():
    assert lo.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert lo.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert lo.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_action_rule():
    # only STAY available -> always STAY
    for e in range(50):
        assert lo.sample_action(1, e, 7, e % 5, 0b10000) == 4
    # the k-th set bit, k = mulhi(word, popcount)
    for seed, env, t, agent, mask in [(3, 11, 5, 0, 0b10101), (3, 11, 5, 5, 0b11111), (2**40 + 1, 99, 2**33, 2, 0b10011)]:
        out = lo.philox4x32_10([env, t & 0xFFFFFFFF, agent // 4, t >> 32], [seed & 0xFFFFFFFF, seed >> 32])
        bits = [b for b in range(5) if mask >> b & 1]
        k = (out[agent % 4] * len(bits)) >> 32
        assert lo.sample_action(seed, env, t, agent, mask) == bits[k]
    # roughly uniform over the available set
    counts = [0] * 5
    for e in range(4000):
        counts[lo.sample_action(0, e, 0, 0, 0b10110)] += 1
    assert counts[0] == counts[3] == 0
    assert all(1150 < counts[b] < 1520 for b in (1, 2, 4))
