"""Extracts the `pipeline` blocks (the JSON `_class_name` schema the reference's entry script
instantiates, examples/ctsd_generation_example.py:41-69) of every CTSD example config under
/root/reference/examples into tests/golden/example_pipeline_blocks.json, so the drop-in test can
run where the reference tree is absent.  Checkpoint paths are site-specific and dropped.

    python tests/golden/make_example_blocks.py
"""
import glob
import json
import os

REF = os.environ.get("DWM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DROP = ("pretrained_model_name_or_path", "model_checkpoint_path")


def blocks():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "examples", "ctsd_*.json"))):
        with open(path) as f:
            cfg = json.load(f)
        out[os.path.basename(path)] = {
            "generator_seed": cfg.get("generator_seed"),
            "pipeline": {k: v for k, v in cfg["pipeline"].items() if k not in DROP}}
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "example_pipeline_blocks.json"), "w") as f:
        json.dump(blocks(), f, indent=1, sort_keys=True)
    print(sorted(blocks()))
