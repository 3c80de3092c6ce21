"""bench_forest.py -- forest-em inside-outside EM throughput (hyperedges/s); run as `python bench.py --workload forest`.

Workload = BASELINE.json configs[4]: 100k random AND/OR derivation forests IN TOTAL (strong scaling:
rank r of N keeps block r), EVERY forest its own random shape (about 500 hyperedges each, 15 % shared sub-forests), rule
ids Zipf over 10^6 rules, normalization groups of 2..50 rules.
A step = one EM iteration: inside + outside + expected counts over the resident forests, (all-reduce of the rule
count table when N>1), NormalizeGroups M-step.  Same JSON contract as bench.py; the unit of work is one hyperedge
(an AND node with its tail list)."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
FOREST_ORACLE = os.path.join(ROOT, "oracle", "_build", "forest_oracle")
METRIC, UNIT = "forest_em_iteration_hyperedges_per_sec", "hyperedges/s"


def cpu_forest_oracle(fs, n_sample, procs, precision, budget_s=15.0):
    from carmel_b200 import synth
    if not os.path.exists(FOREST_ORACLE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    d = tempfile.mkdtemp(prefix="cb200_fcpu_")
    try:
        per = max(1, n_sample // procs)
        t0 = time.time()
        dirs = []
        for i in range(procs):
            sub = dict(fs)
            lo, hi = i * per, min((i + 1) * per, len(fs["node_off"]) - 1)
            if hi <= lo:
                break
            b, e = int(fs["node_off"][lo]), int(fs["node_off"][hi])
            sub["node_off"] = fs["node_off"][lo:hi + 1] - fs["node_off"][lo]
            for k in ("next", "label", "backref"):
                sub[k] = fs[k][b:e]
            dirs.append(synth.write_forests(sub, os.path.join(d, f"s{i}")))
        flag = ["-U"] if precision == 64 else []
        r = subprocess.run([FOREST_ORACLE, *flag, "-f", dirs[0]["forests"], "-n", dirs[0]["norm"], "--time-estimate=1"],
                           capture_output=True, text=True)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        iters = int(max(1, min(200, budget_s / max(j["seconds"], 1e-4) / 2)))
        ps = [subprocess.Popen([FOREST_ORACLE, *flag, "-f", x["forests"], "-n", x["norm"], f"--time-estimate={iters}"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for x in dirs]
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
        he = sum(o["hyperedges"] for o in outs)
        secs = max(o["seconds"] for o in outs)
        return {"value": he * iters / secs, "unit": UNIT, "cores": len(dirs), "kind": "port",
                "sample": f"first {per * len(dirs)} forests ({he} hyperedges), {iters} E-steps (inside + outside + counts), "
                          f"{len(dirs)} independent single-threaded forest-oracle processes; prep {time.time() - t0:.1f}s"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def config_of(a, world):
    return {"workload": f"configs[4] forest-em inside-outside: {100000 * a.scale} random AND/OR forests in total, every forest "
                        "its own shape (~500 hyperedges each, 15% shared sub-forests), rule ids Zipf(1.0) over 1e6 rules, "
                        "normgroups of 2..50 rules",
            "forests_total": 100000 * a.scale, "l2": "forest topology ~2 GB > 126 MB L2",
            "parallelism": f"forests sharded over {world} GPU(s) (strong scaling), one NCCL all-reduce of the rule count "
                           "table per iteration, issued by the library on its own stream"}


def inside_parity(fs, n, precision, local):
    """ln inside of the first n forests at uniform initial weights: GPU (a small separate context) against the CPU
    oracle (forest-em -i 0 -S).  Returns {"n", "max_rel", "tol", "ok"}."""
    import numpy as np
    from carmel_b200 import synth
    from carmel_b200.forest_api import Forests
    n = min(n, len(fs["node_off"]) - 1)
    d = tempfile.mkdtemp(prefix="cb200_fpar_")
    try:
        files = synth.write_forests(fs, d, n_forests=n)
        flag = ["-U"] if precision == 64 else []
        subprocess.run([FOREST_ORACLE, *flag, "-f", files["forests"], "-n", files["norm"], "-i", "0", "-S", f"{d}/s.out"],
                       capture_output=True, text=True, check=True)
        want = np.array([float(t[2:]) if t.startswith("e^") else (np.log(float(t)) if float(t) > 0 else -np.inf)
                         for t in open(f"{d}/s.out").read().split()][:n])
        G = Forests(device=local, precision=precision)
        G.set_rules(fs["rulespace"], fs["group_off"], fs["group_members"])
        w0 = np.full(fs["rulespace"], -np.inf)
        go, gm = fs["group_off"].astype(np.int64), fs["group_members"].astype(np.int64)
        w0[gm] = -np.log(np.repeat(np.diff(go), np.diff(go)).astype(np.float64))
        G.set_params(w0)
        e = int(fs["node_off"][n])
        G.add(fs["node_off"][:n + 1], fs["next"][:e], fs["label"][:e], fs["backref"][:e])
        G.estimate()
        got = G.inside(n)
        G.close()
        rel = float(np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))))
        tol = 1e-6 if precision == 64 else 1e-4
        return {"n": int(n), "max_rel": rel, "tol": tol, "ok": bool(rel <= tol),
                "what": "ln inside[root] per forest at uniform initial weights, GPU vs CPU oracle (forest-em -i 0 -S)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_arm(a):
    from carmel_b200 import synth
    t0 = time.time()
    procs = max(1, os.cpu_count() or 1)
    fs = synth.make_forests(n_forests=100 * procs, n_rules=1000000, templates=0)
    cb = cpu_forest_oracle(fs, 100 * procs, procs, a.precision, budget_s=8.0 * max(1, min(a.steps, 4)))
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if a.precision == 64 else "f32", "data": "synthetic", "config": config_of(a, a.gpus),
            "impl": "reference", "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "forest-em needs Boost (absent): this arm times the CPU oracle restatement", "wall_s": time.time() - t0}
    print(json.dumps(line))


def run(a, rank, world, local, as_leg=False, token=None, with_cpu=True):
    """as_leg: called by bench.py for its c5 leg (process group already up; returns the line on rank 0 instead of
    printing it).  token: NCCL rendezvous token for the library's own all-reduce (world > 1)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, FALLBACK_HBM_GBS, measured_peaks
    from carmel_b200 import synth
    from carmel_b200.forest_api import UNIFORM, Forests

    assert torch.cuda.is_available(), "bench.py needs a GPU (carmel_b200 has no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1 and not as_leg:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1 and token is None:
        import carmel_b200 as cb
        box = [cb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        token = box[0]
    t_build = time.time()
    fs = synth.make_forests(n_forests=100000 * a.scale, n_rules=1000000, seed=20260105, templates=0, part=(rank, world))
    stream = torch.cuda.Stream()
    if os.environ.get("CB200_DIRTY"):  # debugging aid: the library's cudaMalloc calls then see non-zero memory
        x = torch.full((int(os.environ["CB200_DIRTY"]) << 30,), 0xAB, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        del x
        torch.cuda.empty_cache()
    F = Forests(device=local, precision=a.precision)
    if os.environ.get("CB200_FOREST_LAYOUT"):  # development runs: 1 group, 2 thread, 3 level (default: the library's choice)
        F.set_layout(int(os.environ["CB200_FOREST_LAYOUT"]))
    F.set_stream(stream.cuda_stream)
    if world > 1:
        F.comm_init_rank(world, rank, token)
    F.set_rules(fs["rulespace"], fs["group_off"], fs["group_members"])
    rulespace = fs["rulespace"]
    w0 = np.full(rulespace, -np.inf)
    go, gm = fs["group_off"].astype(np.int64), fs["group_members"].astype(np.int64)
    w0[gm] = -np.log(np.repeat(np.diff(go), np.diff(go)).astype(np.float64))
    F.set_params(w0)
    F.add(fs["node_off"], fs["next"], fs["label"], fs["backref"])
    t_build = time.time() - t_build
    tot_local = F.totals()
    tot = torch.tensor([tot_local["hyperedges"], tot_local["nodes"], tot_local["links"], tot_local["forests"]], dtype=torch.float64,
                       device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    he_total, nodes_total, links_total, forests_total = (float(x) for x in tot.tolist())
    def step():
        F.estimate_launch()
        F.allreduce_counts()  # NCCL on the library's stream (no-op on one GPU)
        r = F.estimate_finish()
        F.maximize(0.0, 0.0, UNIFORM)
        return r

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(a.warmup):
        last = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = F.launch_count()
    k_ms = []

    def step_and_sample():
        step()
        k_ms.append(F.last_time_ms())

    ms = timed(step_and_sample, a.steps)
    launches = F.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = he_total * a.steps / (ms / 1e3)

    h_params = torch.empty(rulespace, dtype=torch.float64).pin_memory()
    h_counts = torch.empty(rulespace, dtype=torch.float64).pin_memory()
    import ctypes as C
    f64p = C.POINTER(C.c_double)
    lib = F.lib
    lib.cml_forests_get_params(F.h, C.cast(h_params.data_ptr(), f64p))

    def e2e_step():
        lib.cml_forests_set_params(F.h, C.cast(h_params.data_ptr(), f64p))            # H2D: rule weights
        F.estimate_launch()
        F.allreduce_counts()
        F.estimate_finish()                                                           # D2H: likelihood scalars
        lib.cml_forests_get_counts(F.h, C.cast(h_counts.data_ptr(), f64p), rulespace)  # D2H: expected rule counts
        F.maximize(0.0, 0.0, UNIFORM)
        lib.cml_forests_get_params(F.h, C.cast(h_params.data_ptr(), f64p))            # D2H: new rule weights

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, a.steps)
    e2e = {"value": he_total * a.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8 * rulespace,
           "d2h_bytes_per_step": 16 * rulespace + 24, "ms_per_step": ms_e2e / a.steps,
           "forests": f"resident in HBM; one-time generation + levelisation + upload took {t_build:.2f}s on this rank"}
    if rank == 0:
        peaks, which = measured_peaks()
        # algorithmic bytes of one E-step: per node label + child_off (inside) and label + par_off (outside), 4 B each;
        # per child/parent link one u32 in each pass; inside/posterior values stay in shared memory
        bytes_step = 16.0 * tot_local["nodes"] + 8.0 * tot_local["links"]
        lay = {**F.layout_stats(), **F.level_stats()}
        if lay["tile_forests"]:
            # thread-per-forest tiles: one u32 word per node header and per link in each pass (rule ids ride in the
            # headers), inside[] and gamma[] written once (tree children / tree parents are read back from the
            # per-lane shared-memory stacks, the rest are L1/L2 hits by post-order locality); padding is not counted
            bytes_step = 4.0 * lay["steps"] + 2.0 * (a.precision // 8) * tot_local["nodes"]
        kms = sum(m for m, _ in k_ms) / len(k_ms)
        achieved = bytes_step / (kms / 1e3) / 1e9
        peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
        from bench import measured_traffic
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (measured_traffic("k_forest_level", tot_local["hyperedges"]) if lay["level_forests"] else
                                measured_traffic("k_forest_thread", tot_local["hyperedges"]) if lay["tile_forests"] else None) if a.precision == 32 else None,
                    "peak_source": which, "kernel": ("k_forest_level (a 128-thread CTA per run of forests, 16 CTAs per SM, nodes height-major, values in "
                                                     "shared memory, 8 B/node + 2 B/link streamed per pass)" if lay["level_forests"] else
                                                     "k_forest_thread" if lay["tile_forests"] else "k_forest_warp/k_forest_cta") +
                    f" (inside + outside + counts, {k_ms[0][1]} launch(es) per iteration)",
                    "kernel_ms": kms, "algorithmic_bytes_per_hyperedge": bytes_step / max(1, tot_local["hyperedges"]),
                    "hyperedges_per_launch_set": tot_local["hyperedges"], "kernel_share_of_step": kms / (ms / a.steps)}
        cpu = None
        if with_cpu and not os.environ.get("CB200_NO_CPU"):  # (development runs skip the CPU baseline)
            try:
                procs = max(1, os.cpu_count() or 1)
                cpu = cpu_forest_oracle(fs, 100 * procs, procs, a.precision, budget_s=10.0 if as_leg else 15.0)
            except Exception as ex:
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        try:
            parity = inside_parity(fs, 256, a.precision, local)
        except Exception as ex:
            parity = {"n": 0, "max_rel": None, "ok": False, "error": str(ex)[:200]}
        roofline["padded_steps_over_steps"] = (lay["padded_steps"] / lay["steps"]) if lay.get("steps") else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "parity": parity,
                "dtype": "f64" if a.precision == 64 else "f32", "data": "synthetic", "config": config_of(a, world),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "totals": {"forests": forests_total, "hyperedges": he_total, "nodes": nodes_total, "links": links_total,
                           "rulespace": rulespace, "avg_ln_p": last[0] / max(1.0, last[2] - last[1])},
                "layout": lay, "templates": "every forest its own shape (csrc/tools/forest_synth.c)"}
        if not as_leg:
            print(json.dumps(line))
    if world > 1 and not as_leg:  # (as a leg: bench.py's run_leg synchronises the ranks, over a channel that survives a CUDA error)
        dist.barrier()
    F.close()
    if world > 1 and not as_leg:
        dist.destroy_process_group()
    return line if rank == 0 else None
