"""Helpers for the forest-em tests: a tiny independent forest parser with brute-force enumeration of all
derivations (the pin for the inside/outside numerics), and seeded random forest generators."""
import math
import re
from collections import Counter

import numpy as np

TOKEN = re.compile(r"#\d+\(|#\d+|\(|\)|OR|\d+")


def parse_forest(text):
    """-> nested nodes: ('and', rule, [children]) | ('or', [children]) | ('ref', id); defs = {id: node}"""
    toks = TOKEN.findall(text)
    pos = 0
    defs = {}

    def node():
        nonlocal pos
        t = toks[pos]
        if t.startswith("#") and not t.endswith("("):
            pos += 1
            return ("ref", int(t[1:]))
        did = None
        if t.startswith("#"):
            did = int(t[1:-1])
            t = "("
        if t == "(":
            pos += 1
            head = toks[pos]
            pos += 1
            kids = []
            while toks[pos] != ")":
                kids.append(node())
            pos += 1
            n = ("or", kids) if head == "OR" else ("and", int(head), kids)
        else:
            pos += 1
            n = ("and", int(t), [])
        if did is not None:
            defs[did] = n
        return n

    root = node()
    assert pos == len(toks), (pos, len(toks))
    return root, defs


def enumerate_derivations(root, defs, w):
    """all derivations as (probability, Counter(rule -> uses)); w[rule] = probability"""
    def rec(n):
        if n[0] == "ref":
            return rec(defs[n[1]])
        if n[0] == "or":
            out = []
            for k in n[1]:
                out.extend(rec(k))
            return out
        out = [(w[n[1]], Counter({n[1]: 1}))]
        for k in n[2]:
            sub = rec(k)
            out = [(p * q, c + d) for p, c in out for q, d in sub]
        return out
    return rec(root)


def brute_force(text, w):
    """(total probability, {rule: expected count under the posterior})"""
    root, defs = parse_forest(text)
    ds = enumerate_derivations(root, defs, w)
    total = sum(p for p, _ in ds)
    counts = Counter()
    if total > 0:
        for p, c in ds:
            for r, k in c.items():
                counts[r] += p * k / total
    return total, counts


def split_forests(text):
    """top-level forests of a file (balanced parens or a bare leaf)"""
    out, depth, cur = [], 0, ""
    for t in TOKEN.findall(text):
        cur += (" " if cur and not cur.endswith("(") and t != ")" else "") + t
        depth += t.endswith("(")
        depth -= t == ")"
        if depth == 0:
            out.append(cur)
            cur = ""
    return out


def parse_ln(tok):
    tok = tok.strip()
    if tok.startswith("e^"):
        return float(tok[2:])
    v = float(tok)
    return math.log(v) if v > 0 else -math.inf


def read_weights(path):
    """forest-em params/counts file -> [ln weight], index 0 = parameter #1"""
    return [parse_ln(t) for t in open(path).read().split()]


def random_forest(rng, n_rules, depth=4, share=0.15, zipf=1.2):
    """random AND/OR forest text in forest-em syntax (SURVEY 8d config C5 shape): OR fan-out U[2,4], AND arity
    {1:.3, 2:.6, 3:.1} below the leaves' depth, `share` of the children are #k back references"""
    defined = []
    next_id = [1]
    ranks = np.arange(1, n_rules + 1) ** -zipf
    ranks /= ranks.sum()

    def rule():
        return 1 + int(rng.choice(n_rules, p=ranks))

    def and_node(d):
        if d <= 0 or rng.random() < 0.15:
            return str(rule())
        ar = int(rng.choice([1, 2, 3], p=[0.3, 0.6, 0.1]))
        return "(" + str(rule()) + "".join(" " + child(d - 1) for _ in range(ar)) + ")"

    def or_node(d):
        k = int(rng.integers(2, 5))
        return "(OR " + " ".join(and_node(d) for _ in range(k)) + ")"

    def child(d):
        if defined and rng.random() < share:
            return "#" + str(defined[int(rng.integers(0, len(defined)))])
        body = or_node(d) if rng.random() < 0.7 else and_node(d)
        if body.startswith("(") and rng.random() < 0.3:
            i = next_id[0]
            next_id[0] += 1
            defined.append(i)
            return "#" + str(i) + body
        return body

    return or_node(depth)


def random_normgroups(rng, n_rules, lo=2, hi=6, leave_out=0.1):
    """partition of (most of) 1..n_rules into groups; returns the file text"""
    ids = [i for i in range(1, n_rules + 1) if rng.random() >= leave_out]
    rng.shuffle(ids)
    groups, i = [], 0
    while i < len(ids):
        k = int(rng.integers(lo, hi + 1))
        groups.append(ids[i:i + k])
        i += k
    return "(" + " ".join("(" + " ".join(map(str, g)) + ")" for g in groups) + ")\n"
