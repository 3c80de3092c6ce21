"""Synthetic workloads of the shapes BASELINE.json names, written as carmel text files (WFST format +
pair corpus) so that the product, the CPU oracle and (if it could be built) the reference binary all
consume byte-identical inputs.  numpy PCG64, seeds 20260101 + config index.

  cipher : 27-state locked letter-bigram LM (*e*:letter FSA)  o  1-state 27x27 substitution channel,
           ciphertext lines of `line_len` letters (configs[1]; 729 lattice arcs per letter)
  hmm    : Q-tag bigram FSA (*e*:tag)  o  1-state tag->word channel, V words, k tags allowed per word,
           sentences of U[10,40] words (configs[2]; about k*k lattice arcs per word)
"""
from __future__ import annotations

import os

import numpy as np

LETTERS = ["_"] + [chr(ord("A") + i) for i in range(26)]
CIPHER = ["_"] + [chr(ord("a") + i) for i in range(26)]


def _q(s: str) -> str:
    return '"' + s + '"'


def write_cipher(outdir: str, n_lines: int = 2000, line_len: int = 50, seed: int = 20260102) -> dict:
    rng = np.random.default_rng(seed)
    os.makedirs(outdir, exist_ok=True)
    n = 27
    lm = rng.dirichlet(np.full(n, 0.35), size=n) * 0.98 + 0.02 / n  # every bigram possible
    lm /= lm.sum(1, keepdims=True)
    key = np.concatenate([[0], 1 + rng.permutation(26)])  # space stays space
    wfsa, fst, data = (os.path.join(outdir, f) for f in ("lm.wfsa", "channel.fst", "cipher.data"))
    with open(wfsa, "w") as f:
        f.write("_\n")
        for a in range(n):
            for b in range(n):
                f.write(f"({LETTERS[a]} ({LETTERS[b]} *e* {_q(LETTERS[b])} {lm[a, b]:.12g}!))\n")
    with open(fst, "w") as f:
        f.write("0\n")
        for a in range(n):
            for b in range(n):
                f.write(f"(0 (0 {_q(LETTERS[a])} {_q(CIPHER[b])}))\n")
    cdf = np.cumsum(lm, axis=1)
    with open(data, "w") as f:
        for _ in range(n_lines):
            s = 0
            out = []
            u = rng.random(line_len)
            for t in range(line_len - 1):
                s = int(min(n - 1, np.searchsorted(cdf[s], u[t])))
                out.append(CIPHER[key[s]])
            out.append("_")  # lines end in the final (space) state
            f.write("\n" + " ".join(_q(c) for c in out) + "\n")
    return dict(workload="cipher", files=[data, wfsa, fst], argv=["--train-cascade", data, wfsa, fst],
                n_lines=n_lines, line_len=line_len, letters=n_lines * line_len, seed=seed)


def write_hmm(outdir: str, n_sent: int = 125000, n_tags: int = 32, vocab: int = 5000, tags_per_word: int = 4,
              seed: int = 20260103, len_range: tuple = (10, 41)) -> dict:
    rng = np.random.default_rng(seed)
    os.makedirs(outdir, exist_ok=True)
    fsa, fst, data = (os.path.join(outdir, f) for f in ("tags.fsa", "lexicon.fst", "sentences.data"))
    tags = [f"T{i}" for i in range(n_tags)]
    words = [f"w{i}" for i in range(vocab)]
    with open(fsa, "w") as f:  # like carmel-tutorial/tagging.fsa: *e*:tag bigram, trainable
        f.write("F\n")
        for t in tags:
            f.write(f"(S ({t} *e* {_q(t)} 1))\n")
        for a in tags:
            for b in tags:
                f.write(f"({a} ({b} *e* {_q(b)} 1))\n")
            f.write(f"({a} (F *e* *e* 1))\n")
    tag_w = 1.0 / np.arange(1, n_tags + 1)
    tag_w /= tag_w.sum()
    allowed = [np.sort(rng.choice(n_tags, size=tags_per_word, replace=False, p=tag_w)) for _ in range(vocab)]
    with open(fst, "w") as f:  # like tagging.fst: dictionary of allowed (tag, word) pairs
        f.write("0\n")
        for w in range(vocab):
            for t in allowed[w]:
                f.write(f"(0 (0 {_q(tags[t])} {_q(words[w])} 1))\n")
    zipf = 1.0 / np.arange(1, vocab + 1) ** 1.1
    zipf /= zipf.sum()
    lens = rng.integers(len_range[0], len_range[1], size=n_sent)
    toks = rng.choice(vocab, size=int(lens.sum()), p=zipf)
    with open(data, "w") as f:
        pos = 0
        buf = []
        for ln in lens:
            buf.append("\n" + " ".join(_q(words[w]) for w in toks[pos:pos + ln]) + "\n")
            pos += ln
            if len(buf) >= 4096:
                f.write("".join(buf))
                buf = []
        f.write("".join(buf))
    return dict(workload="hmm", files=[data, fsa, fst], argv=["--train-cascade", data, fsa, fst], n_sent=n_sent,
                n_tags=n_tags, vocab=vocab, tags_per_word=tags_per_word, words=int(lens.sum()), seed=seed)


def head_corpus(src: str, dst: str, n_pairs: int) -> int:
    """first n_pairs (input line, output line) pairs of an unweighted corpus file -> dst"""
    k = 0
    with open(src) as f, open(dst, "w") as g:
        while k < n_pairs:
            a = f.readline()
            b = f.readline()
            if not b:
                break
            g.write(a)
            g.write(b)
            k += 1
    return k


# ---------------------------------------------------------------------------------------------------
# forest-em workload (configs[4]: random binary-branching AND/OR derivation forests)
# ---------------------------------------------------------------------------------------------------
def _forest_template(rng, depth: int, share: float):
    """one random forest shape as pre-order arrays (next, kind, backref target); kind 0 = OR, 1 = AND, 2 = backref.
    OR fan-out U[2,4]; AND arity {1:.3, 2:.6, 3:.1}; `share` of the children are references to an earlier subforest."""
    nxt, kind, target = [], [], []
    shareable = []

    def node_and(d):
        i = len(nxt)
        nxt.append(0)
        kind.append(1)
        target.append(0)
        if d > 0 and rng.random() >= 0.12:
            for _ in range(int(rng.choice([1, 2, 3], p=[0.3, 0.6, 0.1]))):
                child(d - 1)
        nxt[i] = len(nxt)
        return i

    def node_or(d):
        i = len(nxt)
        nxt.append(0)
        kind.append(0)
        target.append(0)
        for _ in range(int(rng.integers(2, 5))):
            node_and(d)
        nxt[i] = len(nxt)
        return i

    def child(d):
        if shareable and rng.random() < share:
            i = len(nxt)
            nxt.append(i + 1)
            kind.append(2)
            target.append(shareable[int(rng.integers(0, len(shareable)))])
            return
        i = node_or(d) if rng.random() < 0.7 else node_and(d)
        if nxt[i] > i + 1:
            shareable.append(i)

    node_or(depth)
    return np.array(nxt, np.uint32), np.array(kind, np.uint8), np.array(target, np.uint32)


FOREST_CHUNKS = 64  # the distinct-shape corpus is defined as 64 independently seeded chunks (so that rank r of N can
                    # generate exactly its block of the same corpus)


def _distinct_shapes(n_forests: int, seed: int, target_hyperedges: int, share: float = 0.15, threads: int | None = None,
                     part: tuple = (0, 1)):
    """every forest its own random shape: the C generator (csrc/tools/forest_synth.c, same shape law as
    _forest_template), chunks of forests generated on host threads (ctypes releases the GIL).  part = (r, N): only
    the chunks of block r of N are generated; returns the arrays of that block and its chunk ids."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from . import build as _b
    if not os.path.exists(_b.SYNTH_LIB):
        _b.build()
    lib = C.CDLL(_b.SYNTH_LIB)
    lib.cb200_synth_forests.restype = C.c_uint64
    lib.cb200_synth_forests.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_uint64] + [C.c_void_p] * 4
    threads = threads or max(1, min(32, (os.cpu_count() or 1) // max(1, part[1])))
    n_chunks = max(1, min(FOREST_CHUNKS, n_forests // 64 or 1))
    bounds = np.linspace(0, n_forests, n_chunks + 1).astype(np.int64)
    c_lo, c_hi = part[0] * n_chunks // part[1], (part[0] + 1) * n_chunks // part[1]

    def gen(c):
        n = int(bounds[c + 1] - bounds[c])
        cap = n * (4 * target_hyperedges) + 4096
        off = np.zeros(n + 1, np.uint64)
        nx, lab, br = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.uint8)
        tot = int(lib.cb200_synth_forests(n, seed * 1000003 + c, target_hyperedges, share, cap, off.ctypes.data,
                                          nx.ctypes.data, lab.ctypes.data, br.ctypes.data))
        if n and not tot:
            raise RuntimeError("forest generator: output capacity too small")
        return off, nx[:tot], lab[:tot], br[:tot]

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(gen, range(c_lo, c_hi)))
    n_local = int(bounds[c_hi] - bounds[c_lo])
    node_off = np.zeros(n_local + 1, np.uint64)
    base = 0
    for k, (off, _, _, _) in enumerate(parts):
        c = c_lo + k
        node_off[bounds[c] - bounds[c_lo] + 1:bounds[c + 1] - bounds[c_lo] + 1] = off[1:] + np.uint64(base)
        base += int(off[-1])
    if not parts:
        return node_off, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    return (node_off, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]),
            np.concatenate([p[3] for p in parts]))


def make_forests(n_forests: int = 100000, n_rules: int = 1000000, seed: int = 20260105, templates: int = 512,
                 target_hyperedges: int = 500, zipf: float = 1.0, group_seed: int = 20260105, part: tuple = (0, 1)) -> dict:
    """Synthetic forest corpus in the C ABI's layout (cml_forest_batch) plus normalization groups of size U[2,50]
    covering every rule.  Shapes come from `templates` random forests (depth 6-12, about `target_hyperedges` AND nodes
    each); every forest draws its own rule ids, Zipf(`zipf`) over `n_rules`."""
    rng = np.random.default_rng(seed)
    if templates <= 0:  # every forest its own shape (BASELINE configs[4] as written: "100k random forests")
        node_off, nxt, lab0, backref = _distinct_shapes(n_forests, seed, target_hyperedges, part=part)
        rng = np.random.default_rng([seed, part[0], part[1]])  # rule ids of this block
        total = int(node_off[-1])
        label = lab0.copy()
        and_pos = np.nonzero((lab0 == 1) & (backref == 0))[0]
        label[(backref == 0)] = 0
        ranks = np.arange(1, n_rules + 1, dtype=np.float64) ** -zipf
        cdf = np.cumsum(ranks / ranks.sum())
        ids = np.searchsorted(cdf, rng.random(len(and_pos)), side="right").astype(np.uint32)
        perm = rng.permutation(n_rules).astype(np.uint32)
        label[and_pos] = 1 + perm[np.minimum(ids, n_rules - 1)]
        grng = np.random.default_rng(group_seed)
        order = 1 + grng.permutation(n_rules).astype(np.uint64)
        gsz = grng.integers(2, 51, size=n_rules // 2 + 1)
        goff = np.concatenate([[0], np.cumsum(gsz)])
        goff = goff[goff < n_rules]
        goff = np.concatenate([goff, [n_rules]]).astype(np.uint64)
        return {"node_off": node_off, "next": nxt, "label": label, "backref": backref, "n_rules": n_rules,
                "rulespace": n_rules + 1, "group_off": goff, "group_members": order, "hyperedges": int(len(and_pos)),
                "nodes": total, "templates": len(node_off) - 1}
    shapes = []
    tries = 0
    while len(shapes) < templates and tries < 40 * templates:
        tries += 1
        t = _forest_template(rng, depth=int(rng.integers(4, 8)), share=0.15)
        he = int((t[1] == 1).sum())
        if target_hyperedges * 0.4 <= he <= target_hyperedges * 1.8:
            shapes.append(t)
    if not shapes:
        raise RuntimeError("no forest template of the requested size")
    pick = rng.integers(0, len(shapes), n_forests)
    sizes = np.array([len(s[0]) for s in shapes], np.uint64)
    node_off = np.zeros(n_forests + 1, np.uint64)
    np.cumsum(sizes[pick], out=node_off[1:])
    total = int(node_off[-1])
    nxt = np.empty(total, np.uint32)
    label = np.empty(total, np.uint32)
    backref = np.empty(total, np.uint8)
    # fill template by template (vectorised over the forests that use it)
    for s, (t_next, t_kind, t_target) in enumerate(shapes):
        fs = np.nonzero(pick == s)[0]
        if not len(fs):
            continue
        idx = (node_off[fs][:, None] + np.arange(len(t_next), dtype=np.uint64)[None, :]).astype(np.int64)
        nxt[idx] = t_next[None, :]
        backref[idx] = (t_kind == 2)[None, :]
        lab = np.where(t_kind == 2, t_target, 0).astype(np.uint32)
        label[idx] = lab[None, :]
    # AND nodes: recover from the templates
    is_and = np.empty(total, bool)
    for s, (t_next, t_kind, t_target) in enumerate(shapes):
        fs = np.nonzero(pick == s)[0]
        if not len(fs):
            continue
        idx = (node_off[fs][:, None] + np.arange(len(t_next), dtype=np.uint64)[None, :]).astype(np.int64)
        is_and[idx] = (t_kind == 1)[None, :]
    and_pos = np.nonzero(is_and)[0]
    ranks = np.arange(1, n_rules + 1, dtype=np.float64) ** -zipf
    cdf = np.cumsum(ranks / ranks.sum())
    ids = np.searchsorted(cdf, rng.random(len(and_pos)), side="right").astype(np.uint32)
    perm = rng.permutation(n_rules).astype(np.uint32)  # frequent rules are scattered over the id space
    label[and_pos] = 1 + perm[np.minimum(ids, n_rules - 1)]
    # normalization groups: consecutive runs of a random permutation of the rule ids (own seed: every rank of a
    # multi-GPU run must see the same groups whatever forests it generated)
    grng = np.random.default_rng(group_seed)
    order = 1 + grng.permutation(n_rules).astype(np.uint64)
    gsz = grng.integers(2, 51, size=n_rules // 2 + 1)
    goff = np.concatenate([[0], np.cumsum(gsz)])
    goff = goff[goff < n_rules]
    goff = np.concatenate([goff, [n_rules]]).astype(np.uint64)
    return {"node_off": node_off, "next": nxt, "label": label, "backref": backref, "n_rules": n_rules, "rulespace": n_rules + 1,
            "group_off": goff, "group_members": order, "hyperedges": int(len(and_pos)), "nodes": total,
            "templates": len(shapes)}


def forest_text(fs: dict, f: int) -> str:
    """forest f of make_forests() in forest-em's text syntax"""
    b, e = int(fs["node_off"][f]), int(fs["node_off"][f + 1])
    nxt, label, backref = fs["next"][b:e], fs["label"][b:e], fs["backref"][b:e]
    n = e - b
    ids = {}
    for p in range(n):
        if backref[p] and int(label[p]) not in ids:
            ids[int(label[p])] = len(ids) + 1
    out, ends = [], [n]
    for p in range(n):
        while p == ends[-1] and len(ends) > 1:
            out.append(")")
            ends.pop()
        if p:
            out.append(" ")
        if p in ids:
            out.append(f"#{ids[p]}")
        if backref[p]:
            out.append(f"#{ids[int(label[p])]}")
        elif nxt[p] == p + 1:
            out.append(f"({label[p]})" if p in ids else str(label[p]))
        else:
            out.append("(" + (str(label[p]) if label[p] else "OR"))
            ends.append(int(nxt[p]))
    out.append(")" * (len(ends) - 1))
    return "".join(out)


def write_forests(fs: dict, outdir: str, n_forests: int | None = None) -> dict:
    """write the first n forests + the normalization groups as forest-em input files"""
    os.makedirs(outdir, exist_ok=True)
    n = len(fs["node_off"]) - 1 if n_forests is None else min(n_forests, len(fs["node_off"]) - 1)
    with open(os.path.join(outdir, "forests"), "w") as o:
        for f in range(n):
            o.write(forest_text(fs, f))
            o.write("\n")
    go, gm = fs["group_off"], fs["group_members"]
    with open(os.path.join(outdir, "norm"), "w") as o:
        o.write("(")
        for g in range(len(go) - 1):
            o.write("(" + " ".join(str(int(x)) for x in gm[int(go[g]):int(go[g + 1])]) + ")\n")
        o.write(")\n")
    return {"forests": os.path.join(outdir, "forests"), "norm": os.path.join(outdir, "norm"), "n_forests": n}
