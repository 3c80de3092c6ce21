"""Shared helpers for the parity tests: run the CPU oracle, parse its dumps."""
import json
import os
import re
import shutil
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden():
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


def stage(tmp_path, *names):
    """copy golden input fixtures into a scratch dir (trained files are written next to inputs)"""
    out = []
    for n in names:
        dst = os.path.join(str(tmp_path), n)
        shutil.copyfile(os.path.join(GOLDEN, n), dst)
        out.append(dst)
    return out


def run(binary, args, cwd=None, timeout=120, env=None):
    p = subprocess.run([binary, *args], cwd=cwd, capture_output=True, text=True, timeout=timeout,
                       env={**os.environ, **env} if env else None)
    return p.returncode, p.stdout, p.stderr


def trajectory_log2(stderr_text):
    """[(iter, log2 corpus probability)] from carmel-format 'i=N (rate=r): probability=2^x' lines"""
    return [(int(m.group(1)), float(m.group(2)))
            for m in re.finditer(r"i=(\d+) \(rate=[0-9.e+-]+\): probability=2\^(-?[0-9.e+-]+)", stderr_text)]


def read_history(path):
    """oracle / product --history file: iter ln_prob ln_weighted_prob max_change"""
    rows = [ln.split() for ln in open(path) if ln.strip()]
    return [(int(r[0]), float(r[1]), float(r[2]), float(r[3])) for r in rows]


def read_trellis_dump(path):
    """Binary trellis dump (oracle_cli.cpp dump_trellis / carmel-b200 --dump-trellis).
    Returns dict with per-example arrays in the C-ABI batch layout (reference state order)."""
    buf = open(path, "rb").read()
    n_ex, n_arcs_table = struct.unpack_from("<II", buf, 0)
    pos = 8
    ex_states, ex_fin, ex_weight = [], [], []
    arc_off, arc_dst, arc_id = [], [], []
    for _ in range(n_ex):
        ns, na, fin, w = struct.unpack_from("<IIId", buf, pos)
        pos += 20
        ex_states.append(ns)
        ex_fin.append(fin)
        ex_weight.append(w)
        off = [0]
        for _s in range(ns):
            (k,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            pairs = np.frombuffer(buf, dtype="<u4", count=2 * k, offset=pos).reshape(k, 2)
            pos += 8 * k
            arc_dst.append(pairs[:, 0])
            arc_id.append(pairs[:, 1])
            off.append(off[-1] + k)
        assert off[-1] == na
        arc_off.append(np.asarray(off, np.uint32))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint32)
    return dict(n_ex=n_ex, n_arcs_table=n_arcs_table, ex_states=np.asarray(ex_states, np.uint32),
                ex_fin=np.asarray(ex_fin, np.uint32), ex_weight=np.asarray(ex_weight, np.float64),
                arc_off=cat(arc_off).astype(np.uint32), arc_dst=cat(arc_dst).astype(np.uint32),
                arc_id=cat(arc_id).astype(np.uint32))


def read_estimate_dump(path):
    buf = open(path, "rb").read()
    na, ne = struct.unpack_from("<II", buf, 0)
    a = np.frombuffer(buf, dtype="<f8", offset=8)
    return dict(ln_w=a[:na].copy(), ln_counts=a[na:2 * na].copy(), ln_p=a[2 * na:2 * na + ne].copy())


def rel_close(a, b, rel, abs_=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)) + abs_)


# ---------------------------------------------------------------------------------------------
# random transducers / corpora in carmel text format (seeded) for property tests
# ---------------------------------------------------------------------------------------------
def random_wfst(rng, n_states=4, n_in=3, n_out=3, density=0.5, eps_rate=0.25, lock_rate=0.1, tie_rate=0.1,
                in_prefix="i", out_prefix="o"):
    """returns (text, ins, outs, arcs) with arcs = [(src, dst, in or None, out or None)]; state 0 is the
    start, the last state is final; *e*:*e* arcs only go forward so lattices stay acyclic."""
    ins = [f'"{in_prefix}{k}"' for k in range(n_in)]
    outs = [f'"{out_prefix}{k}"' for k in range(n_out)]
    names = [f"q{k}" for k in range(n_states)]
    arcs = [(k, k + 1, ins[k % n_in], outs[k % n_out], 0.5, "") for k in range(n_states - 1)]
    for s_ in range(n_states):
        for d in range(n_states):
            if rng.random() > density:
                continue
            for _ in range(1 + int(rng.integers(0, 3))):
                i = None if rng.random() < eps_rate else ins[int(rng.integers(0, n_in))]
                o = None if rng.random() < eps_rate else outs[int(rng.integers(0, n_out))]
                if i is None and o is None and d <= s_:
                    o = outs[0]
                suffix = ""
                r = rng.random()
                if r < lock_rate:
                    suffix = "!"
                elif r < lock_rate + tie_rate:
                    suffix = f"!{1 + int(rng.integers(0, 2))}"
                arcs.append((s_, d, i, o, float(rng.uniform(0.05, 1.0)), suffix))
    lines = [names[-1]]
    for (s_, d, i, o, w, suf) in arcs:
        lines.append(f"({names[s_]} ({names[d]} {i or '*e*'} {o or '*e*'} {w:.6g}{suf}))")
    return "\n".join(lines) + "\n", ins, outs, [(a[0], a[1], a[2], a[3]) for a in arcs]


def sample_pairs(rng, arcs, n_states, n_pairs=6, max_steps=12, weighted=False, noise=0.15, ins=None, outs=None):
    """corpus text whose pairs are mostly random accepting walks (so derivations exist), with a few
    random corruptions (so some examples have no derivation)."""
    by_src = {}
    for a in arcs:
        by_src.setdefault(a[0], []).append(a)
    out = []
    made = 0
    tries = 0
    while made < n_pairs and tries < 200 * n_pairs:
        tries += 1
        s_, xi, xo = 0, [], []
        for _ in range(max_steps):
            if s_ == n_states - 1 and rng.random() < 0.5:
                break
            cand = by_src.get(s_, [])
            if not cand:
                break
            a = cand[int(rng.integers(0, len(cand)))]
            if a[2]:
                xi.append(a[2])
            if a[3]:
                xo.append(a[3])
            s_ = a[1]
        if s_ != n_states - 1:
            continue
        if rng.random() < noise and ins and outs:
            (xi if rng.random() < 0.5 else xo).append((ins if rng.random() < 0.5 else outs)[0])
        if weighted:
            out.append(f"{rng.uniform(0.5, 3.0):.4g}")
        out.append(" ".join(xi))
        out.append(" ".join(xo))
        made += 1
    return "\n".join(out) + "\n"


def _ln_weight(tok):
    """ln of a carmel weight token, or None if tok is not a weight"""
    import math
    t = tok.rstrip(")")
    t = re.sub(r"!\d*$", "", t)
    try:
        if t.startswith("e^"):
            return float(t[2:])
        if t.endswith("ln"):
            return float(t[:-2])
        v = float(t)
        return math.log(v) if v > 0 else float("-inf")
    except ValueError:
        return None


def compare_wfst_text(got, want, rel, ln_floor=-690.0):
    """token-wise comparison of two carmel WFST listings: identical structure, weights equal within
    rel (relative, on the probability) -- weights below e^ln_floor are only required to be tiny in both."""
    import math
    g, w = got.split(), want.split()
    assert len(g) == len(w), (len(g), len(w))
    worst = 0.0
    for a, b in zip(g, w):
        if a == b:
            continue
        la, lb = _ln_weight(a), _ln_weight(b)
        assert la is not None and lb is not None, (a, b)
        assert re.sub(r"^[^!)]*", "", a) == re.sub(r"^[^!)]*", "", b), (a, b)  # same lock/tie suffix and parens
        if la < ln_floor and lb < ln_floor:
            continue
        d = abs(math.expm1(la - lb)) if abs(la - lb) < 1 else float("inf")
        worst = max(worst, d)
        assert d <= rel, (a, b, d)
    return worst
