"""Randomised differential tests on hand-built SoA tiles (tests/random_tiles.py)."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import random_tiles as rt
from methyldackel_b200 import _abi as A


def _oracle(cfg, ref, beg, end, ce, soa):
    cap = (end - beg) + 16
    exp = (A.MdCall * cap)(); st = A.MdTileStats()
    rc = ob.lib().mdo_extract_tile_ce(C.byref(cfg), ref, len(ref), beg, end, ce[0], ce[1], C.byref(soa), exp, cap, C.byref(st))
    assert rc == 0
    return bytes(C.string_at(exp, st.n_calls * 16)), st


def test_oracle_accepts_adversarial_tiles(built):
    """CPU tier: the generator and the oracle agree on the struct layout (no crash, plausible output) for every encoding"""
    rng = np.random.default_rng(7)
    ref = rt.random_reference(rng, 5000)
    for bits in (8, 4, 2):
        t = rt.Tile(rng, len(ref), 400, qual_bits=bits)
        cfg = A.default_config(keepCHG=1, keepCHH=1, minMapq=0, minPhred=1, ignoreFlags=0, keepDupes=1, keepSingleton=1, keepDiscordant=1)
        calls, st = _oracle(cfg, ref, 0, len(ref), (0, 0), t.soa)
        assert st.n_admitted > 50 and st.n_calls > 100


@pytest.mark.gpu
@pytest.mark.parametrize("split", [0, 3, 7], ids=["one_cta_per_window", "split3", "split7"])
@pytest.mark.parametrize("seed", list(range(24)))
def test_gpu_equals_oracle_on_random_tiles(built, seed, split, monkeypatch):
    """split > 0 forces the deep-tile path: several CTAs per window, counters summed in HBM, the last CTA writes the calls"""
    from methyldackel_b200 import api
    if split:
        if seed % 2:
            pytest.skip("forced split runs on every other seed")
        monkeypatch.setenv("MD_FORCE_SPLIT", str(split))
    rng = np.random.default_rng(1000 + seed)
    reflen = int(rng.choice([300, 4096, 4097, 9000, 20000]))
    ref = rt.random_reference(rng, reflen)
    n = int(rng.choice([0, 1, 31, 32, 33, 500, 3000]))
    t = rt.Tile(rng, reflen, n, qual_bits=int(rng.choice([8, 4, 2])), maxlen=int(rng.choice([40, 180, 400])), depth_hot=bool(rng.random() < 0.3))
    cfg = rt.random_config(rng)
    beg = int(rng.integers(0, reflen // 3)); end = int(rng.integers(beg, reflen + 50))
    ce = (0, 0)
    if cfg.minConversionEfficiency > 0:
        ce = (max(0, beg - 2), min(reflen, end + 11))
    exp, est = _oracle(cfg, ref, beg, min(end, reflen), ce, t.soa)
    with api.GpuContext(cfg) as g:
        g.load_contig(0, ref)
        cap = max(end, reflen) + 64
        calls = (A.MdCall * cap)(); st = A.MdTileStats()
        td = A.MdTileDesc(0, beg, end, ce[0], ce[1])
        rc = g.g.md_extract_tile(g.h, C.byref(td), C.byref(t.soa), calls, cap, C.byref(st))
        assert rc == 0, g.g.md_last_error()
    assert st.n_admitted == est.n_admitted
    assert bytes(C.string_at(calls, st.n_calls * 16)) == exp


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(6)))
def test_gpu_mbias_equals_oracle_on_random_tiles(built, seed, monkeypatch):
    from methyldackel_b200 import api
    if seed % 2:
        monkeypatch.setenv("MD_FORCE_SPLIT", "5")
    rng = np.random.default_rng(5000 + seed)
    reflen = int(rng.choice([4096, 9000, 20000]))
    ref = rt.random_reference(rng, reflen)
    t = rt.Tile(rng, reflen, int(rng.choice([33, 500, 3000])), qual_bits=int(rng.choice([8, 4, 2])), maxlen=int(rng.choice([180, 400])))
    cfg = rt.random_config(rng); cfg.noOverlapMerge = 1; cfg.minOppositeDepth = 0; cfg.minConversionEfficiency = 0.0
    for k in range(16):
        cfg.bounds[k] = 0
    h = A.load_host()
    bounds = (C.c_uint32 * 256)()
    n_chunks = h.mdh_chunk_bounds(ref, reflen, int(rng.choice([700, 3000, 100000])), 0, 0, bounds, 255)
    ehist = (C.c_uint32 * (4 * 2 * A.MD_MBIAS_MAXLEN * 2))(); elens = (C.c_int32 * 4)()
    assert ob.lib().mdo_mbias_tile(C.byref(cfg), ref, reflen, 0, reflen, bounds, n_chunks, C.byref(t.soa), ehist, elens, None) == 0
    with api.GpuContext(cfg) as g:
        g.load_contig(0, ref)
        g.set_mbias_chunks(0, [bounds[i] for i in range(n_chunks + 1)])
        g.mbias_tile(0, 0, reflen, t.soa)
        ghist, glens = g.mbias_hist()
    assert glens == list(elens) and bytes(ghist) == bytes(ehist)
