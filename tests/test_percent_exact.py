"""The default bedGraph line prints (int)(100.0 * nm / (nm + nu)) (extract.c:50).  host/format.hpp computes it in integers; this
checks the arithmetic claim behind that on IEEE doubles: random pairs, pairs around exact multiples — where a
rounding error would flip the truncation — and the largest 32-bit counts."""
import random


def _ref(nm, nu):
    return int(100.0 * float(nm) / float(nm + nu))            # C: 100.0 * ((double) nm) / (nm + nu), then (int)


def _int(nm, nu):
    return (100 * nm) // (nm + nu)


def test_integer_percentage_equals_the_double_expression():
    rng = random.Random(12345)
    pairs = []
    for _ in range(200000):
        tot = rng.randrange(1, 1 << rng.randrange(1, 33))
        pairs.append((rng.randrange(0, tot + 1), tot))
    for tot in [1, 2, 3, 7, 100, 101, 997, 65535, 65536, (1 << 31) - 1, (1 << 32) - 1] + [rng.randrange(1, 1 << 32) for _ in range(2000)]:
        for pct in range(0, 101):
            base = pct * tot // 100                            # the counts on either side of every percentage boundary
            for nm in (base - 1, base, base + 1):
                if 0 <= nm <= tot:
                    pairs.append((nm, tot))
    assert len(pairs) > 500000
    for nm, tot in pairs:
        assert _ref(nm, tot - nm) == _int(nm, tot - nm), (nm, tot)
