"""bench.py's pure helpers (no GPU): the JSON blocks the driver reads must come out well-formed whatever the timers say."""
import json
import os
import sys

import cases

sys.path.insert(0, cases.ROOT)
import bench  # noqa: E402


def test_inflate_summary_numbers():
    d = bench.inflate_summary(192816128, 2, 7.606, 553800000 * 13, 182900000 * 13, 1.98, 2.37, 6558.1, 21.6)
    assert d["kernel"] == "inflate_kernel" and d["segments_per_step"] == 2
    assert abs(d["compressed_GBps"] - 12.68) < 0.01 and abs(d["inflated_GBps"] - 38.38) < 0.05
    assert 0.005 < d["hbm_frac"] < 0.01 and 0.6 < d["share_of_e2e_step"] < 0.8
    json.dumps(d)


def test_inflate_summary_without_timers_or_peak():
    assert bench.inflate_summary(1000, 2, 0.0, 0, 0, 0.0, 0.0, None, 0.0) == {"kernel": "inflate_kernel", "ms_per_segment": None}
    d = bench.inflate_summary(1000, 1, 1.0, 3000, 1000, 0.1, 0.1, None, 0.0)
    assert "hbm_frac" not in d and "share_of_e2e_step" not in d and d["inflated_GBps"] == round(d["compressed_GBps"] * 3, 2)


def test_traffic_lookup_reads_the_committed_capture():
    t, src = bench.traffic_of("count_warp")
    assert t and t > 100e6 and "ncu --set full" in src
    assert bench.traffic_of("count_warp_c3") == (None, None)            # no capture for that configuration: null, not a stand-in


def test_algorithmic_bytes_formula():
    # SURVEY 8d / DESIGN 3.1: n x (ceil(L/2) + L + 20) + 4 x cigar ops + contig + 8 x calls
    assert bench.algorithmic_bytes(10, 150, 12, 1000, 7) == 10 * (75 + 150 + 20) + 4 * 12 + 1000 + 8 * 7
