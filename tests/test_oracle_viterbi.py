"""The oracle's best-derivation pass (oracle_cli.cpp --dump-viterbi: the restatement behind the --viterbi parity tests,
fst.h:769-800 bestPaths on the composed string x transducer x string machine) against brute-force enumeration of EVERY
derivation of the dumped lattices: the reported weight must be the maximum over all paths, and the reported arc ids
must be a path of the lattice with exactly that weight.  CPU only."""
import math
import os

import numpy as np
import pytest

from helpers import random_wfst, read_estimate_dump, read_trellis_dump, run, sample_pairs


def _paths(off, dst, aid, fin, ln_w, limit=200000):
    """all start -> goal paths of an acyclic lattice as (weight, [arc-table ids])"""
    out, stack, n = [], [(0, 0.0, [])], 0
    while stack:
        s, w, ids = stack.pop()
        if s == fin:
            out.append((w, ids))  # (the goal of a pruned lattice may have outgoing arcs only if they lead back to it)
        for k in range(off[s], off[s + 1]):
            n += 1
            assert n < limit, "lattice too large for enumeration"
            stack.append((int(dst[k]), w + ln_w[aid[k]], ids + [int(aid[k])]))
    return out


@pytest.mark.parametrize("seed", range(8))
def test_oracle_viterbi_is_the_maximum_over_all_derivations(oracle_bin, tmp_path, seed):
    rng = np.random.default_rng(20270301 + seed)
    ns = int(rng.integers(2, 6))
    fst, ins, outs, arcs = random_wfst(rng, n_states=ns, eps_rate=float(rng.uniform(0, 0.3)))
    corpus = sample_pairs(rng, arcs, ns, n_pairs=8, max_steps=6, ins=ins, outs=outs)
    d = str(tmp_path)
    f, c = os.path.join(d, "r.fst"), os.path.join(d, "r.data")
    open(f, "w").write(fst)
    open(c, "w").write(corpus)
    rc, _, err = run(oracle_bin, ["-t", f"--dump-viterbi={d}/v", c, f])
    assert rc == 0, err
    rc, _, err = run(oracle_bin, ["-t", f"--dump-trellis={d}/t", f"--dump-estimate={d}/e", c, f])
    assert rc == 0, err
    t, est = read_trellis_dump(f"{d}/t"), read_estimate_dump(f"{d}/e")
    vit = [(float(x[0]), [int(y) for y in x[2:2 + int(x[1])]]) for x in (ln.split() for ln in open(f"{d}/v"))]
    assert len(vit) == t["n_ex"] >= 4
    row = arc = 0
    checked = 0
    for e in range(t["n_ex"]):
        n = int(t["ex_states"][e])
        off = t["arc_off"][row:row + n + 1].astype(np.int64) + arc
        paths = _paths(off, t["arc_dst"], t["arc_id"], int(t["ex_fin"][e]), est["ln_w"])
        row += n + 1
        arc = int(off[-1])
        best = max(w for w, _ in paths)
        w, ids = vit[e]
        if best == -math.inf:
            assert w == -math.inf
            continue
        assert abs(w - best) <= 1e-12 * max(1.0, abs(best)), (e, w, best)
        assert any(ids == p and abs(pw - best) <= 1e-12 * max(1.0, abs(best)) for pw, p in paths), (e, ids)
        checked += 1
    assert checked >= 3
