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
              seed: int = 20260103) -> dict:
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
    lens = rng.integers(10, 41, size=n_sent)
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
