"""Helpers shared by CPU and GPU tests: load a golden case and regenerate its
weights/inputs from the recorded seeds (oracle/synth.py)."""
import os

import numpy as np

from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EVAL_CASES = ["b2_n1024", "b3_n1000_ragged", "b2_n1031_ragged", "b5_n37_tiny",
              "b2_n1_single", "b2_n2048_realistic", "b2_n256_defaultbn"]


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    B, N, wseed, iseed, rbn, sub_c, sub_n = (int(v) for v in g["meta"])
    dist = str(g["dist"]) if "dist" in g else "parity"
    sd = synth.make_state_dict(wseed, bool(rbn))
    ctx, line = synth.make_inputs(B, N, seed=iseed, dist=dist)
    return g, sd, ctx, line, (sub_c, sub_n)


def report(**kv):
    """Append one measured-error record to gpurun_out/parity_report.jsonl (created on GPU runs; ignored elsewhere):
    the numbers behind the assertions, for DESIGN.md / profiles/."""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(kv) + "\n")
