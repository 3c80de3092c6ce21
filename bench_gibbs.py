"""bench_gibbs.py -- `carmel --crp` Gibbs sampling throughput (samples/s); run as `python bench.py --workload gibbs`.

Workload = BASELINE.json configs[3]: the synthetic cipher model (27x27 channel o locked 27-state
bigram LM) with a Dirichlet prior alpha=0.01 on the channel, ciphertext in lines of 50 letters; one block = one
line; one sample = one block resampled once (graehl/shared/gibbs.hpp:844-872).  A step = one sweep over all
resident blocks in batched mode (all blocks in parallel against the previous sweep's counts, then the count
deltas are applied); the exact sequential sampler's rate is reported next to it (`sequential`).  The default
keeps 5000 lines (250k letters) per GPU so that the host lattice build stays short; --scale 4 is C4's 1M letters."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(ROOT, "oracle", "_build", "carmel_oracle")
METRIC, UNIT = "gibbs_samples_per_sec", "samples/s"


def config_of(a, world):
    return {"workload": "configs[3] --crp Gibbs on the synthetic cipher (27x27 channel o locked bigram LM, alpha=0.01), "
                        f"{5000 * a.scale} lines x 50 letters IN TOTAL, batched sweeps",
            "blocks_total": 5000 * a.scale, "l2": "lattices 2.9 GB per 20,000 lines > 126 MB L2",
            "parallelism": ("one GPU" if world == 1 else
                            f"blocks sharded over {world} GPUs (strong scaling); every rank samples its blocks against the counts "
                            "of the previous sweep, one NCCL all-reduce of the count deltas per sweep (SURVEY 8e); the exact "
                            "sampler is sequential over the corpus and stays on one GPU")}


def sample_parity(files, n_lines=40, sweeps=3, seed=5, gpu=0):
    """north_star: identical sampled derivations in sequential mode given the same uniform draws.  The exact sampler of
    the product (carmel-b200 --crp, a separate process) and the CPU oracle run `sweeps` sweeps over the first n_lines of
    the SAME corpus with the shared counter-based uniforms; the derivations of the last sweep and the per-sweep
    cache-model probabilities must agree."""
    from carmel_b200 import CLI_PATH
    if not os.path.exists(ORACLE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    d = tempfile.mkdtemp(prefix="cb200_gpar_")
    try:
        head = os.path.join(d, "head.data")
        with open(files[0]) as f, open(head, "w") as g:
            for _ in range(2 * n_lines):
                g.write(f.readline())
        models = [shutil.copy(m, d) for m in files[1:]]
        common = ["--crp", "-M", str(sweeps), "--priors=0,1e-2", f"--seed={seed}", "-q"]
        out = {}
        for name, exe, extra in (("gpu", CLI_PATH, [f"--gpu={gpu}"]), ("cpu", ORACLE, [])):
            r = subprocess.run([exe, *common, *extra, f"--history={d}/h.{name}", f"--dump-samples={d}/s.{name}", head, *models],
                               capture_output=True, text=True, cwd=d, timeout=300)
            if r.returncode != 0:
                return {"n": 0, "ok": False, "error": f"{name}: rc {r.returncode}: {r.stderr[-200:]}"}
            out[name] = ([ln.split() for ln in open(f"{d}/s.{name}")],
                         [float(ln.split()[1]) for ln in open(f"{d}/h.{name}") if ln.strip()])
        same = out["gpu"][0] == out["cpu"][0]
        hg, hc = out["gpu"][1], out["cpu"][1]
        rel = max((abs(a - b) / max(1.0, abs(b)) for a, b in zip(hg, hc)), default=0.0) if len(hg) == len(hc) else float("inf")
        return {"n": len(out["cpu"][0]), "sweeps": sweeps, "identical_derivations": bool(same), "max_rel": rel, "tol": 1e-9,
                "ok": bool(same and rel <= 1e-9),
                "what": "exact sequential sampler, GPU command line vs CPU oracle on the first n lines with shared uniforms: "
                        "sampled derivations of the last sweep and the per-sweep ln cache-model probabilities"}
    except Exception as ex:  # noqa: BLE001
        return {"n": 0, "ok": False, "error": f"{type(ex).__name__}: {ex}"[:300]}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_oracle_gibbs(files, n_lines, procs, sweeps=3):
    """samples/s of the CPU oracle's sequential sampler on the first n_lines, `procs` independent processes"""
    if not os.path.exists(ORACLE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    d = tempfile.mkdtemp(prefix="cb200_gcpu_")
    try:
        per = max(1, n_lines // procs)
        shards = []
        with open(files[0]) as f:
            for i in range(procs):
                path = os.path.join(d, f"s{i}.data")
                k = 0
                with open(path, "w") as g:
                    while k < per:
                        x, y = f.readline(), f.readline()
                        if not y:
                            break
                        g.write(x)
                        g.write(y)
                        k += 1
                if k:
                    shards.append(path)
        t0 = time.time()
        ps = [subprocess.Popen([ORACLE, "--crp", "-M", str(sweeps), "--priors=0,1e-2", "--seed=1", s, *files[1:]],
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d) for s in shards]
        for p in ps:
            p.wait()
        t_all = time.time() - t0
        # subtract the one-time part (read, compose, lattice build, sweep 0) measured with -M 0
        t0 = time.time()
        ps = [subprocess.Popen([ORACLE, "--crp", "-M", "0", "--priors=0,1e-2", "--seed=1", s, *files[1:]],
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d) for s in shards]
        for p in ps:
            p.wait()
        t_base = time.time() - t0
        secs = max(1e-6, t_all - t_base)
        n = per * len(shards) * sweeps
        return {"value": n / secs, "unit": UNIT, "cores": len(shards), "kind": "port",
                "sample": f"first {per * len(shards)} lines, {sweeps} sequential sweeps (lattices cached), {len(shards)} independent "
                          f"single-threaded oracle processes; {t_all:.1f}s total, {t_base:.1f}s one-time part subtracted"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_arm(a):
    from carmel_b200 import synth
    t0 = time.time()
    d = tempfile.mkdtemp(prefix="cb200_gref_")
    try:
        w = synth.write_cipher(d, n_lines=400, line_len=50)
        procs = max(1, os.cpu_count() or 1)
        cb = cpu_oracle_gibbs(w["files"], 25 * procs, procs, sweeps=max(2, min(a.steps, 4)))
    finally:
        shutil.rmtree(d, ignore_errors=True)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(a, a.gpus), "impl": "reference", "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "carmel needs Boost (absent): this arm times the CPU oracle restatement", "wall_s": time.time() - t0}
    print(json.dumps(line))


def run(a, rank, world, local, as_leg=False, with_cpu=True, token=None):
    """as_leg: called by bench.py for its c4 leg (returns the line instead of printing it; no process-group handling)"""
    import torch
    import torch.distributed as dist
    import carmel_b200 as cb
    from bench import ClockSampler, FALLBACK_HBM_GBS, measured_peaks
    from carmel_b200 import synth

    assert torch.cuda.is_available(), "bench.py needs a GPU (carmel_b200 has no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1 and not as_leg:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sharded = world > 1
    if sharded:  # one corpus for all ranks (rank 0 writes it), every rank keeps its block: --shard + the library's communicator
        d = os.path.join(tempfile.gettempdir(), f"cb200_gibbs_{os.environ.get('MASTER_PORT', 'x')}_{os.getppid()}")
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)
            w = synth.write_cipher(d, n_lines=5000 * a.scale, line_len=50, seed=20260104)
            json.dump(w, open(os.path.join(d, "workload.json"), "w"))
        dist.barrier()
        w = json.load(open(os.path.join(d, "workload.json")))
        if token is None:
            box = [cb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            token = box[0]
    else:
        d = tempfile.mkdtemp(prefix=f"cb200_gibbs_{rank}_")
        w = synth.write_cipher(d, n_lines=5000 * a.scale, line_len=50, seed=20260104 + rank)
    shard = [f"--shard={rank}/{world}"] if sharded else []
    my_seed = 1 + 7919 * rank  # (draws are keyed by the local block number)
    stream = torch.cuda.Stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    no_dense = bool(getattr(a, "no_dense", False))

    def timed(fn, steps, flush=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = 0.0
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
                for _ in range(steps):
                    fn()
                e1.record()
            torch.cuda.synchronize()
            total = e0.elapsed_time(e1)
        else:  # working set fits in L2: flush it between steps, every step has its own event pair
            for _ in range(steps):
                with torch.cuda.stream(stream):
                    flush_buf.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
        ms = torch.tensor([total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # the same corpus on the lattice sampler (k_gibbs_batched), for the record beside the dense-state sampler
    lattice_leg = None
    if not no_dense and not sharded:
        job0 = cb.Job(["-q", f"--gpu={local}", "--crp", "-M", "1000", "--crp-batched", "--no-dense", "--priors=0,1e-2",
                       "--seed=1", *w["argv"][1:]])
        ctx0 = job0.prepare()
        ctx0.set_stream(stream.cuda_stream)
        k0 = [0]

        def step0():
            ctx0.gibbs_sweep(1, k0[0], seed=1, power=1.0, accumulate_dt=1.0)
            k0[0] += 1

        for _ in range(a.warmup):
            step0()
        n0 = max(2, min(a.steps, 5))
        ms0 = timed(step0, n0)
        lattice_leg = {"value": world * job0.stats()["examples"] * n0 / (ms0 / 1e3), "unit": UNIT, "ms_per_step": ms0 / n0,
                       "note": "same corpus with --no-dense: k_gibbs_batched walks the materialised lattices"}
        job0.close()

    t_build = time.time()
    job = cb.Job(["-q", f"--gpu={local}", "--crp", "-M", "1000", "--crp-batched", *(["--no-dense"] if no_dense else []), *shard,
                  "--priors=0,1e-2", "--seed=1", *w["argv"][1:]], comm_token=token if sharded else None)
    ctx = job.prepare()
    ctx.set_stream(stream.cuda_stream)
    t_build = time.time() - t_build
    info = job.stats()
    blocks, arcs, states = info["examples"], info["trellis_arcs"], info["trellis_states"]
    is_dense = bool(info.get("dense", 0))
    tot = torch.tensor([blocks, arcs, states], dtype=torch.float64, device="cuda")
    if sharded:
        dist.all_reduce(tot)
    blocks_total = float(tot[0].item())
    c0 = ctx.collective_count()

    sweep = [0]

    def step():
        ctx.gibbs_sweep(1, sweep[0], seed=my_seed, power=1.0, accumulate_dt=1.0)
        sweep[0] += 1

    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms = timed(step, a.steps, flush=is_dense)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = blocks_total * a.steps / (ms / 1e3)
    collectives = ctx.collective_count() - c0

    # e2e: what the host loop of `carmel --crp` does every sweep: read the sampled derivations back (for the cache-model
    # probability, gibbs.hpp:712-742)
    cap = ctx.gibbs_sample_capacity()
    h_len = torch.empty(blocks, dtype=torch.int32).pin_memory()
    h_arcs = torch.empty(cap, dtype=torch.int32).pin_memory()

    def e2e_step():
        step()
        ctx.gibbs_get_samples_ptr(h_len.data_ptr(), h_arcs.data_ptr(), cap)

    e2e_step()
    ms_e2e = timed(e2e_step, a.steps, flush=is_dense)
    e2e = {"value": blocks_total * a.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": 0,
           "d2h_bytes_per_step": 4 * (blocks + cap), "ms_per_step": ms_e2e / a.steps,
           "lattices": f"resident; one-time host build + upload took {t_build:.2f}s on this rank"}
    # the exact sequential sampler on the same blocks (one CTA walks the corpus)
    n_seq = max(1, min(2, a.steps))
    ms_seq = None if sharded else timed(lambda: ctx.gibbs_sweep(0, 100000 + sweep[0], seed=1, power=1.0, accumulate_dt=0.0), n_seq)
    if rank == 0:
        peaks, which = measured_peaks()
        # per sample one backward sweep over the block's lattice (8 B/arc record + the state scores written and
        # read once) + the sampled path (16 B per visited state)
        lattice_bytes = 8.0 * arcs + 2.0 * 8 * states + 16.0 * 51 * blocks
        peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
        if is_dense:
            # dense-state sampler: per position the symbol (twice), one beta row of 32 doubles written and read, the
            # sampled arc id; the 187 KB probability table and the own-sample tables live in L2 / shared memory
            positions = 50.0 * blocks
            bytes_step = positions * (2 * 2 + 2 * 32 * 8 + 4)
            achieved = bytes_step / (ms / a.steps / 1e3) / 1e9
            from bench import measured_traffic
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": measured_traffic("k_gibbs_dense") if a.scale == 1 else None, "peak_source": which,
                        "kernel": "k_gibbs_dense_table + k_gibbs_dense (dense-state backward filter + forward sample, one "
                                  "warp per block) + k_gibbs_apply",
                        "kernel_ms": ms / a.steps, "algorithmic_bytes_per_sample": bytes_step / max(1, blocks),
                        "lattice_equivalent": {"algorithmic_bytes": lattice_bytes,
                                               "achieved_gbs": lattice_bytes / (ms / a.steps / 1e3) / 1e9,
                                               "frac_of_hbm_peak": lattice_bytes / (ms / a.steps / 1e3) / 1e9 / peak,
                                               "note": "what the same samples cost as a sweep over 8-byte lattice records "
                                                       "(SURVEY 8d); this kernel does not read them"}}
        else:
            bytes_step = lattice_bytes
            achieved = bytes_step / (ms / a.steps / 1e3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "peak_source": which,
                        "kernel": "k_gibbs (backward filter + forward sample per block) + k_gibbs_apply",
                        "kernel_ms": ms / a.steps, "algorithmic_bytes_per_sample": bytes_step / max(1, blocks)}
        cpu = None
        if with_cpu:
            try:
                procs = max(1, os.cpu_count() or 1)
                cpu = cpu_oracle_gibbs(w["files"], 25 * procs, procs, sweeps=2 if as_leg else 3)
            except Exception as ex:
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        parity = sample_parity(w["files"], gpu=local) if not sharded else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "parity": parity,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config_of(a, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "collectives": int(collectives), "clocks": clocks,
                "sequential": None if ms_seq is None else
                {"value": blocks * n_seq / (ms_seq / 1e3), "unit": UNIT, "ms_per_sweep": ms_seq / n_seq,
                 "note": "exact collapsed sampler, identical derivations to the CPU oracle (tests/test_gibbs_gpu.py)"},
                "totals": {"blocks": blocks_total, "trellis_arcs": float(tot[1].item()), "trellis_states": float(tot[2].item())},
                "sampler": "dense-state (cml_gibbs_attach_dense)" if is_dense else "lattice"}
        if lattice_leg is not None:
            line["lattice_path"] = lattice_leg
        if is_dense:
            line["config"]["l2"] = ("dense-state working set (symbols, beta rows, probability table) fits in L2: L2 flushed "
                                    "(256 MB write) between timed sweeps, each sweep timed with its own CUDA-event pair")
        if not as_leg:
            print(json.dumps(line))
    if world > 1 and not as_leg:
        dist.barrier()
    job.close()
    if not sharded or rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    if world > 1 and not as_leg:
        dist.destroy_process_group()
    return line if rank == 0 else None
