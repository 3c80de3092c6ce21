"""Regenerate tests/golden/ from the read-only reference checkout (/root/reference).

Run in the build container only (the GPU box has no /root/reference).  Copies the tutorial INPUT
fixtures the reference's own golden run log was produced from, and extracts the per-iteration
likelihood trajectories / composition sizes / trained weights that log and the *.trained files pin
 into golden.json.
"""
import json
import os
import re
import shutil

REF = "/root/reference/carmel"
HERE = os.path.dirname(os.path.abspath(__file__))
TUT = os.path.join(REF, "carmel-tutorial")

INPUTS = ["epron-jpron.fst", "epron-jpron.data", "cipher.wfsa", "cipher.fst", "cipher.data", "tagging.fsa",
          "tagging.fst", "tagging.data", "cluster.fsa", "cluster.data", "cat.fsa", "spellout.fst"]
TRAINED = ["cipher.fst.trained", "cipher.wfsa.trained"]


def trajectory(lines):
    out = []
    for ln in lines:
        m = re.search(r"i=(\d+) \(rate=1\): probability=2\^(-?[0-9.e+-]+)", ln)
        if m:
            out.append([int(m.group(1)), float(m.group(2))])
    return out


def main():
    for f in INPUTS + TRAINED:
        shutil.copyfile(os.path.join(TUT, f), os.path.join(HERE, f))
        os.chmod(os.path.join(HERE, f), 0o644)
    for f in ["span.spell.corpus", "span.spell.wfst"]:
        shutil.copyfile(os.path.join(REF, "test", f), os.path.join(HERE, f))
        os.chmod(os.path.join(HERE, f), 0o644)
    # forest-em: sample INPUTS (forest-em ships no expected outputs) + the parse/print round-trip vectors
    # of its unit test (forest-em/forest.hpp:1041-1048 test_forests[])
    fdir = os.path.join(HERE, "forest")
    os.makedirs(fdir, exist_ok=True)
    for f in ["forest", "forests", "norm", "norm_and_forests", "best_forest", "best_norm", "best_weights"]:
        shutil.copyfile(os.path.join(os.path.dirname(REF), "forest-em", "sample", f), os.path.join(fdir, f))
        os.chmod(os.path.join(fdir, f), 0o644)
    src = open(os.path.join(os.path.dirname(REF), "forest-em", "forest.hpp")).read()
    m = re.search(r"test_forests\[\]\s*=\s*\{(.*?)\};", src, re.S)
    json.dump(re.findall(r'"([^"]*)"', m.group(1)), open(os.path.join(fdir, "test_forests.json"), "w"), indent=1)
    trace = open(os.path.join(TUT, "commands.trace"), errors="replace").read().split("\n")
    runs = []  # every EM run in the log: starts at an "i=1" line
    for ln in trace:
        for it, v in trajectory([ln]):
            if it == 1:
                runs.append([])
            runs[-1].append([it, v])

    def run_starting(v0, nth=0):
        for r in runs:
            if abs(r[0][1] - v0) < 1e-9 * abs(v0):
                if nth == 0:
                    return r
                nth -= 1
        raise KeyError(v0)
    g = {
        "source": "carmel/carmel-tutorial/commands.trace (line numbers 1-based)",
        "epron_jpron": {"lines": "9-19", "trajectory_log2": run_starting(-43.6883), "states": 57, "arcs": 154,
                        "converged_at": 5,
                        "final_model_excerpt": {'(S22 "AY" "A"': 0.999916773012262, '(S "L" "R"': 0.999986128835377,
                                                '(S26 "L" "R"': 1.38711646229817e-05}},
        "tagging": {"lines": "5869-5889", "trajectory_log2": run_starting(-293197),
                    "composed_states": 46, "composed_arcs": 400994, "converged_at": 9},
        "cipher": {"lines": "6905-6950", "trajectory_log2": run_starting(-2245.63),
                   "composed_states": 57, "composed_arcs": 11511, "converged_at": 22},
        "cluster_first_start": {"lines": "112-114", "trajectory_log2": run_starting(-258374)[:2]},
        # `--train-cascade -HJ -! 100 cluster.data cat.fsa spellout.fst`: the RNG-free first start
        "cat_spellout_first_start": {"lines": "641-649", "trajectory_log2": run_starting(-258374, 1)[:3],
                                     "composed_states": 4, "composed_arcs": 316, "converged_at": 3},
    }
    json.dump(g, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print({k: (len(v["trajectory_log2"]) if isinstance(v, dict) else v) for k, v in g.items()})


if __name__ == "__main__":
    main()
