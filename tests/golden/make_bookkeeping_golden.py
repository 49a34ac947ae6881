"""Generates tests/golden/bookkeeping.json by EXECUTING the reference's own pure-Python bookkeeping functions
(extracted with `ast` from /root/reference/inference_script.py, which cannot be imported whole because
diffusers/decord are absent).  Run in the build container only (the GPU box has no /root/reference)."""
import ast
import json
import itertools
from pathlib import Path

SRC = Path("/root/reference/inference_script.py")
NAMES = {"make_temporal_chunks", "make_spatial_tiles", "get_valid_tile_region"}


def load_reference_functions():
    tree = ast.parse(SRC.read_text())
    ns = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in NAMES:
            exec(compile(ast.Module([node], []), str(SRC), "exec"), ns)
    return ns


def main():
    ref = load_reference_functions()
    chunks = []
    for F, cl, ov in [(33, 0, 8), (129, 33, 8), (129, 41, 8), (33, 17, 8), (25, 17, 8), (129, 25, 12), (100, 33, 8),
                      (57, 33, 8), (41, 25, 12), (17, 17, 8), (49, 33, 16), (161, 41, 16), (9, 33, 8)]:
        try:
            chunks.append(dict(args=[F, cl, ov], out=ref["make_temporal_chunks"](F, cl, ov)))
        except Exception as e:   # the reference raises IndexError for F <= overlap; record the type
            chunks.append(dict(args=[F, cl, ov], raises=type(e).__name__))
    tiles = []
    for H, W, ts, ov in [(768, 1280, (512, 512), (32, 32)), (1088, 1920, (544, 960), (32, 32)),
                         (256, 256, (128, 128), (32, 32)), (768, 1280, (416, 352), (64, 64)), (768, 1280, (0, 0), (32, 32)),
                         (720, 1280, (384, 640), (64, 64)), (272, 480, (68, 120), (32, 32)), (512, 512, (512, 512), (32, 32)),
                         (640, 640, (256, 320), (32, 64)), (192, 320, (96, 160), (16, 16)), (100, 100, (64, 64), (32, 32))]:
        tiles.append(dict(args=[H, W, list(ts), list(ov)], out=ref["make_spatial_tiles"](H, W, ts, ov)))
    regions = []
    for (F, H, W, cl, ovt, ts, ovhw) in [(129, 136, 240, 33, 8, (68, 120), (32, 32)), (33, 768, 1280, 0, 8, (416, 352), (64, 64)),
                                         (129, 1088, 1920, 25, 12, (544, 960), (32, 32))]:
        shape = (1, 3, F, H, W)
        cs = ref["make_temporal_chunks"](F, cl, ovt)
        tl = ref["make_spatial_tiles"](H, W, ts, ovhw)
        for (t0, t1), (h0, h1, w0, w1) in itertools.product(cs, tl):
            r = ref["get_valid_tile_region"](t0, t1, h0, h1, w0, w1, shape, ovt, ovhw[0], ovhw[1])
            regions.append(dict(args=[t0, t1, h0, h1, w0, w1, list(shape), ovt, ovhw[0], ovhw[1]], out=r))
    out = dict(source=str(SRC), chunks=chunks, tiles=tiles, regions=regions)
    Path(__file__).with_name("bookkeeping.json").write_text(json.dumps(out, indent=0))
    print(len(chunks), len(tiles), len(regions))


if __name__ == "__main__":
    main()
