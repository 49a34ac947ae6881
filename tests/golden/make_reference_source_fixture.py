"""Regenerates the two reference-held fixtures on the hot path (run in the build container, where /root/reference
exists; the GPU box only sees the committed outputs):

  reference_process_video.py.txt   the UNMODIFIED source text of `no_grad`, `prepare_rotary_positional_embeddings`
                                   and `process_video` from /root/reference/inference_script.py (:44-48, :364-503),
                                   ast-extracted.  tests/test_pipeline_gpu.py exec()s it against the dove_b200 pipe on
                                   the GPU: the drop-in claim is checked with the reference's own caller code.
  empty_prompt_embedding.safetensors   copy of pretrained_models/prompt_embeddings/e3b0...b855.safetensors, the
                                   pre-computed T5 embedding of "" ([226, 4096] bf16) that the reference feeds to every
                                   unit (ref :580-590, :423-428) — the only numeric fixture the reference ships.

  cli_flags.json                   every `parser.add_argument` of the reference's __main__ (:507-554): flag, type, default,
                                   nargs, action — the contract `dove_b200.cli.build_parser()` is tested against.

Test fixtures only: nothing under dove_b200/ reads them.
"""
import ast
import hashlib
import shutil
from pathlib import Path

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
WANTED = ("no_grad", "prepare_rotary_positional_embeddings", "process_video")


def cli_flags(tree):
    flags = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" \
                and getattr(node.func.value, "id", "") == "parser":
            kw = {}
            for k in node.keywords:
                if k.arg == "help":
                    continue
                kw[k.arg] = k.value.id if isinstance(k.value, ast.Name) else ast.literal_eval(k.value)
            flags.append({"flag": ast.literal_eval(node.args[0]), **kw})
    return flags


def main():
    src = (REF / "inference_script.py").read_text()
    tree = ast.parse(src)
    import json
    (HERE / "cli_flags.json").write_text(json.dumps(cli_flags(tree), indent=1) + "\n")
    chunks = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            start = min([node.lineno] + [d.lineno for d in node.decorator_list])
            seg = "\n".join(src.splitlines()[start - 1:node.end_lineno])
            chunks.append(f"# ---- inference_script.py:{start}-{node.end_lineno}\n{seg}\n")
    assert len(chunks) == len(WANTED)
    (HERE / "reference_process_video.py.txt").write_text(
        "# FIXTURE: verbatim functions of zhengchen1999/DOVE inference_script.py (see make_reference_source_fixture.py)\n"
        + "\n".join(chunks))
    emb = next((REF / "pretrained_models" / "prompt_embeddings").glob("*.safetensors"))
    assert emb.stem == hashlib.sha256(b"").hexdigest()
    shutil.copyfile(emb, HERE / "empty_prompt_embedding.safetensors")
    print("wrote", [p.name for p in HERE.iterdir()])


if __name__ == "__main__":
    main()
