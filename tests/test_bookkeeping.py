"""Bit-exact integer bookkeeping: goldens produced by executing the reference's own functions
(tests/golden/make_bookkeeping_golden.py), plus property tests of the stitcher."""
import ast
import json
from pathlib import Path

import pytest
import torch
from hypothesis import given, settings, strategies as st

from dove_b200 import bookkeeping as bk
from dove_b200.runner import StitchError, super_resolve

GOLD = json.loads((Path(__file__).parent / "golden" / "bookkeeping.json").read_text())
REF = Path("/root/reference/inference_script.py")


@pytest.mark.parametrize("case", GOLD["chunks"], ids=lambda c: str(c["args"]))
def test_chunks_golden(case):
    if "raises" in case:
        with pytest.raises(Exception) as ei:
            bk.make_temporal_chunks(*case["args"])
        assert type(ei.value).__name__ == case["raises"]
    else:
        assert [list(x) for x in bk.make_temporal_chunks(*case["args"])] == case["out"]


@pytest.mark.parametrize("case", GOLD["tiles"], ids=lambda c: str(c["args"]))
def test_tiles_golden(case):
    H, W, ts, ov = case["args"]
    assert [list(x) for x in bk.make_spatial_tiles(H, W, tuple(ts), tuple(ov))] == case["out"]


def test_regions_golden():
    for case in GOLD["regions"]:
        a = case["args"]
        assert bk.get_valid_tile_region(*a[:6], tuple(a[6]), *a[7:]) == case["out"]


def test_survey_goldens():
    """SURVEY.md section 8a-2 values."""
    assert bk.make_temporal_chunks(129, 33, 8) == [(0, 33), (25, 58), (50, 83), (75, 129)]
    assert bk.make_temporal_chunks(33, 17, 8) == [(0, 17), (9, 33)]
    assert bk.make_spatial_tiles(768, 1280, (512, 512), (32, 32)) == [(0, 768, 0, 512), (0, 768, 480, 1280)]
    assert len(bk.make_spatial_tiles(768, 1280, (416, 352), (64, 64))) == 8
    assert bk.frame_padding(32) == 1 and bk.frame_padding(33) == 0 and bk.spatial_padding(180) == 12


def _ref_functions():
    tree = ast.parse(REF.read_text())
    ns = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("make_temporal_chunks", "make_spatial_tiles",
                                                               "get_valid_tile_region"):
            exec(compile(ast.Module([node], []), str(REF), "exec"), ns)
    return ns


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
@settings(max_examples=300, deadline=None)
@given(F=st.integers(9, 200), cl=st.integers(0, 64), ov=st.integers(0, 32), H=st.integers(32, 400), W=st.integers(32, 400),
       th=st.integers(0, 300), tw=st.integers(0, 300), oh=st.integers(0, 64), ow=st.integers(0, 64))
def test_matches_reference_source(F, cl, ov, H, W, th, tw, oh, ow):
    ref = _ref_functions()

    def both(fn_ref, fn_ours, *a):
        try:
            r = fn_ref(*a)
        except Exception as e:
            with pytest.raises(type(e)):
                fn_ours(*a)
            return None
        o = fn_ours(*a)
        assert [tuple(x) for x in o] == [tuple(x) for x in r]
        return o
    both(ref["make_temporal_chunks"], bk.make_temporal_chunks, F, cl, ov)
    both(ref["make_spatial_tiles"], bk.make_spatial_tiles, H, W, (th, tw), (oh, ow))


def _identity_fn(unit, k, seed):
    return (unit * 0.5 + 0.5)


@pytest.mark.parametrize("F,H,W,cl,ovt,ts,ovhw", [(129, 136, 240, 33, 8, (68, 120), (32, 32)),
                                                  (33, 96, 160, 0, 8, (52, 44), (8, 8)), (25, 64, 64, 17, 8, (0, 0), (32, 32))])
def test_stitch_write_count(F, H, W, cl, ovt, ts, ovhw):
    v = torch.rand(1, 3, F, H, W) * 2 - 1
    out = super_resolve(v, _identity_fn, chunk_len=cl, overlap_t=ovt, tile_size_hw=ts, overlap_hw=ovhw)
    assert torch.equal(out, v * 0.5 + 0.5)


def test_stitch_odd_overlap_aborts():
    v = torch.rand(1, 3, 9, 100, 100)
    with pytest.raises(StitchError):
        super_resolve(v, _identity_fn, tile_size_hw=(40, 40), overlap_hw=(9, 9))
