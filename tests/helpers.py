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


def run(binary, args, cwd=None, timeout=600):
    p = subprocess.run([binary, *args], cwd=cwd, capture_output=True, text=True, timeout=timeout)
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
