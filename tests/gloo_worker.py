"""Worker of tests/test_sharding_gloo.py: one rank of a world_size-N `gloo` job on CPU.

Exercises the N>1 data path of the product without a GPU: the product's shard partitioner + host lattice
builder (carmel-b200 --trellis-only --shard=r/N) or forest reader (forest-em-b200 --parse-only --shard=r/N)
produce this rank's share; a small numpy forward-backward / brute-force inside-outside stands in for the CUDA
E-step; the packed buffer [counts | sum ln p | n] is all-reduced exactly like the library's reduce buffer
(include/carmel_b200.h: cml_reduce_buffer / cml_forests_reduce_buffer); rank 0 writes the result."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from forest_helpers import brute_force  # noqa: E402
from helpers import read_trellis_dump  # noqa: E402


def lse(a, b):
    if a == -math.inf:
        return b
    if b == -math.inf:
        return a
    m = max(a, b)
    return m + math.log1p(math.exp(-abs(a - b)))


def carmel_rank(cfg, rank, world):
    d = cfg["dir"]
    dump = os.path.join(d, f"tr{rank}")
    r = subprocess.run([cfg["cli"], *cfg["args"], "--trellis-only", f"--shard={rank}/{world}", f"--dump-trellis={dump}",
                        *cfg["files"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = read_trellis_dump(dump)
    ln_w = np.asarray(cfg["ln_w"], np.float64)
    buf = np.zeros(len(ln_w) + 2)
    off_pos, arc_pos = 0, 0
    for e in range(t["n_ex"]):
        ns, fin, wt = int(t["ex_states"][e]), int(t["ex_fin"][e]), float(t["ex_weight"][e])
        off = t["arc_off"][off_pos:off_pos + ns + 1].astype(np.int64)
        na = int(off[-1])
        dst = t["arc_dst"][arc_pos:arc_pos + na].astype(np.int64)
        aid = t["arc_id"][arc_pos:arc_pos + na].astype(np.int64)
        off_pos += ns + 1
        arc_pos += na
        # topological order by DFS post-order from state 0
        order, seen, stack = [], [False] * ns, [(0, 0)]
        seen[0] = True
        while stack:
            s, k = stack.pop()
            if k < off[s + 1] - off[s]:
                stack.append((s, k + 1))
                c = int(dst[off[s] + k])
                if not seen[c]:
                    seen[c] = True
                    stack.append((c, 0))
            else:
                order.append(s)
        topo = order[::-1]
        alpha = [-math.inf] * ns
        beta = [-math.inf] * ns
        alpha[0] = 0.0
        for s in topo:
            for k in range(off[s], off[s + 1]):
                alpha[dst[k]] = lse(alpha[dst[k]], alpha[s] + ln_w[aid[k]])
        beta[fin] = 0.0
        for s in reversed(topo):
            for k in range(off[s], off[s + 1]):
                beta[s] = lse(beta[s], ln_w[aid[k]] + beta[dst[k]])
        lnp = alpha[fin]
        for s in topo:
            for k in range(off[s], off[s + 1]):
                v = alpha[s] + ln_w[aid[k]] + beta[dst[k]] - lnp
                if v > -700:
                    buf[aid[k]] += wt * math.exp(v)
        buf[-2] += lnp
        buf[-1] += 1
    return buf


def forest_rank(cfg, rank, world):
    d = cfg["dir"]
    out = os.path.join(d, f"pf{rank}")
    r = subprocess.run([cfg["cli"], "-f", cfg["forests"], "-n", cfg["norm"], "--parse-only", f"--shard={rank}/{world}",
                        f"--print-forests={out}", "-i", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w = {i: math.exp(v) if v > -math.inf else 0.0 for i, v in enumerate(cfg["ln_w"])}
    buf = np.zeros(len(cfg["ln_w"]) + 2)
    for line in open(out):
        if not line.strip():
            continue
        total, counts = brute_force(line, w)
        if total > 0:
            for rule, c in counts.items():
                buf[rule] += c
            buf[-2] += math.log(total)
            buf[-1] += 1
    return buf


def gibbs_rank(cfg, rank, world):
    """one batched --crp sweep's count update on this rank's blocks (cml_gibbs_sweep with a communicator): the
    (new - old) sample counts of the rank's blocks go to a zeroed delta table [n_params | n_norms]; the summed deltas
    are added to the replicated counts.  Blocks are split like the product's contiguous shards."""
    n_params, n_norms = cfg["n_params"], cfg["n_norms"]
    norm = cfg["param_norm"]
    blocks = cfg["blocks"]
    b0, b1 = rank * len(blocks) // world, (rank + 1) * len(blocks) // world
    delta = np.zeros(n_params + n_norms + 2)
    for old, new, wt in blocks[b0:b1]:
        for arcs, sign in ((old, -wt), (new, wt)):
            for a in arcs:
                for p in cfg["chains"][a]:
                    if norm[p] >= 0:
                        delta[p] += sign
                        delta[n_params + norm[p]] += sign
    delta[-1] = b1 - b0
    return delta


def main():
    cfg = json.load(open(sys.argv[1]))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)  # MASTER_ADDR=127.0.0.1 from the environment
    buf = (carmel_rank if cfg["case"] == "carmel" else gibbs_rank if cfg["case"] == "gibbs" else forest_rank)(cfg, rank, world)
    n_local = buf[-1]
    t = torch.from_numpy(buf)
    dist.all_reduce(t)  # the one collective of an EM iteration: fp64 sum of [counts | sum ln p | n]
    sizes = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.float64))
    if rank == 0:
        json.dump({"reduced": t.tolist(), "per_rank": [float(s.item()) for s in sizes]}, open(os.path.join(cfg["dir"], "result.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
