"""dove_b200.cli accepts exactly the reference's command-line flags with the reference's defaults
(ref inference_script.py:507-554; golden extracted by tests/golden/make_reference_source_fixture.py)."""
import json
from pathlib import Path

from dove_b200.cli import build_parser, effective_overlaps

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "cli_flags.json").read_text())


def test_flags_match_reference():
    parser = build_parser()
    acts = {a.option_strings[0]: a for a in parser._actions if a.option_strings and a.option_strings[0] != "-h"}
    assert len(GOLDEN) == 22
    for g in GOLDEN:
        a = acts.pop(g["flag"])
        if g.get("action") == "store_true":
            assert a.nargs == 0 and a.default is False and a.const is True, g
            continue
        assert (a.type.__name__ if a.type else None) == g.get("type"), g
        d = g.get("default")
        assert (tuple(a.default) if isinstance(a.default, (tuple, list)) else a.default) == \
            (tuple(d) if isinstance(d, (tuple, list)) else d), g
        assert a.nargs == g.get("nargs"), g
    assert set(acts) == {"--random_init"}                 # the only extra flag


def test_reference_command_lines_parse():
    """The README / inference.sh invocations of the reference parse unchanged."""
    p = build_parser()
    a = p.parse_args("--input_dir datasets/demo --model_path pretrained_models/DOVE --output_path results/DOVE/demo "
                     "--is_vae_st --save_format yuv420p".split())
    assert a.is_vae_st and a.save_format == "yuv420p" and a.sr_noise_step == 399 and a.upscale == 4 and a.seed == 42
    assert effective_overlaps(a) == (0, (0, 0))           # overlaps only apply when chunking / tiling is on (ref :565-576)
    b = p.parse_args("--input_dir d --model_path m --tile_size_hw 416 368 --overlap_hw 64 64 --chunk_len 25 --overlap_t 12".split())
    assert effective_overlaps(b) == (12, (64, 64)) and tuple(b.tile_size_hw) == (416, 368)
