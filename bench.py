#!/usr/bin/env python
"""bench.py -- EM-iteration throughput of the B200-native carmel training path (trellis arcs/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cipher|hmm] [--precision 64|32]
                  [--space scaled|log] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one EM iteration over the resident derivation lattices: arc weights from the parameter
table, forward, fused backward + expected counts, (all-reduce of the count table when N>1),
normalisation M-step.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".

  value      whole-job trellis arcs/s, lattices resident in HBM, timed with CUDA events on the
             library's stream, max over ranks
  e2e        same metric through the C ABI with HOST buffers: every step copies the parameter
             vector host->device from pinned memory and reads parameters, expected counts and the
             likelihood back (lattices stay resident, like carmel's in-memory derivation cache -:)
  roofline   the forward/backward/count kernel alone: algorithmic bytes (SURVEY.md 8d:
             16 B/arc + 2*sizeof(real)*states/arcs) / its CUDA-event time, against the measured HBM peak
  cpu_baseline / --impl reference
             the CPU oracle (restatement of the reference algorithm; the reference binary needs Boost
             and cannot be built here) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ORACLE = os.path.join(ROOT, "oracle", "_build", "carmel_oracle")
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_traffic(kernel):
    """dram bytes per launch of a kernel from the committed ncu captures (profiles/traffic.json), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
        return float(t["bytes"]) if t else None  # bytes per launch (source file named in profiles/traffic.json)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": FALLBACK_HBM_GBS}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.p = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.p:
            time.sleep(0.15)
            self.p.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name, scale, outdir):
    from carmel_b200 import synth
    if name == "cipher":
        return synth.write_cipher(outdir, n_lines=2000 * scale, line_len=50)
    if name == "hmm":
        return synth.write_hmm(outdir, n_sent=125000 * scale)
    raise SystemExit(f"unknown workload {name}")


def cpu_oracle_throughput(w, sample_pairs, procs, budget_s=20.0):
    """arcs/s of the CPU oracle on the first `sample_pairs` examples, split over `procs` independent
    single-threaded processes (the reference is single threaded; examples are independent)."""
    from carmel_b200 import synth
    if not os.path.exists(ORACLE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    d = tempfile.mkdtemp(prefix="cb200_cpu_")
    try:
        per = max(1, sample_pairs // procs)
        shards = []
        with open(w["files"][0]) as f:
            for i in range(procs):
                path = os.path.join(d, f"s{i}.data")
                k = 0
                with open(path, "w") as g:
                    while k < per:
                        a, b = f.readline(), f.readline()
                        if not b:
                            break
                        g.write(a)
                        g.write(b)
                        k += 1
                if k:
                    shards.append(path)
        # calibrate the iteration count on shard 0 so the whole run stays near the budget
        t0 = time.time()
        r = subprocess.run([ORACLE, "--train-cascade", "-:", "--time-estimate=1", shards[0], *w["files"][1:]],
                           capture_output=True, text=True)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        iters = int(max(1, min(50, budget_s / max(j["seconds"], 1e-3) / 2)))
        ps = [subprocess.Popen([ORACLE, "--train-cascade", "-:", f"--time-estimate={iters}", s, *w["files"][1:]],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for s in shards]
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
        arcs = sum(o["trellis_arcs"] for o in outs)
        secs = max(o["seconds"] for o in outs)
        return {"value": arcs * iters / secs, "unit": "trellis arcs/s", "cores": len(shards), "kind": "port",
                "sample": f"first {per * len(shards)} examples ({arcs} lattice arcs), {iters} cached EM iterations "
                          f"(-: semantics), {len(shards)} independent single-threaded oracle processes; "
                          f"calibration {time.time() - t0:.1f}s"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cipher", choices=["cipher", "hmm", "forest", "gibbs"])
    ap.add_argument("--precision", type=int, default=None, choices=[32, 64],
                    help="score precision (default 64 for cipher/hmm like carmel, 32 for forest like forest-em)")
    ap.add_argument("--space", default="scaled", choices=["scaled", "log"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="per-GPU corpus multiplier")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="cipher: force the lattice (sparse) path")
    ap.add_argument("--no-sparse-leg", action="store_true", help="skip the extra lattice-path measurement")
    ap.add_argument("--unlock-lm", action="store_true",
                    help="cipher: make the bigram LM trainable too (transition counts xi are then part of the E-step)")
    a = ap.parse_args()
    a.warmup = max(3, a.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.precision is None:
        a.precision = 32 if a.workload == "forest" else 64
    if a.workload == "forest":  # forest-em inside-outside (configs[4]): see bench_forest.py
        import bench_forest
        if a.impl == "reference":
            if rank == 0:
                bench_forest.reference_arm(a)
            return
        return bench_forest.run(a, rank, world, local)
    if a.workload == "gibbs":  # --crp Gibbs sampling (configs[3]): see bench_gibbs.py
        import bench_gibbs
        if a.impl == "reference":
            if rank == 0:
                bench_gibbs.reference_arm(a)
            return
        return bench_gibbs.run(a, rank, world, local)
    metric = "em_iteration_trellis_arcs_per_sec"
    unit = "trellis arcs/s"
    wl_name = {"cipher": "configs[1] cipher decipherment: 27x27 channel o 27-state locked bigram LM, "
                         "100k-letter synthetic ciphertext per GPU (2000 lines x 50), EM, sparse layered-CSR path",
               "hmm": "configs[2] HMM tagging-style FST: 32 tags, 5k vocab, 4 tags/word, 125k sentences per GPU"}[a.workload]
    config = {"workload": wl_name, "examples_per_gpu": (2000 if a.workload == "cipher" else 125000) * a.scale,
              "space": a.space, "l2": "inputs larger than L2 per iteration for hmm; cipher arcs 1.2 GB > 126 MB L2",
              "parallelism": f"examples sharded over {world} GPU(s), one NCCL all-reduce of the count table per iteration"}

    # ---------------------------------------------------------------- reference arm (CPU oracle)
    if a.impl == "reference":
        if rank != 0:
            return
        d = tempfile.mkdtemp(prefix="cb200_ref_")
        try:
            w = make_workload(a.workload, 1, d)
            procs = max(1, os.cpu_count() or 1)
            n_pairs = 160 if a.workload == "cipher" else 16000
            t0 = time.time()
            cb = cpu_oracle_throughput(w, n_pairs, procs, budget_s=8.0 * max(1, min(a.steps, 4)))
            line = {"metric": metric, "value": cb["value"], "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
                    "warmup": a.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference",
                    "cpu_baseline": cb,
                    "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "note": "the reference binary needs Boost (absent, no network): this arm times the CPU oracle, a "
                            "restatement of the reference algorithm with the same class of data structures",
                    "wall_s": time.time() - t0}
            print(json.dumps(line))
        finally:
            shutil.rmtree(d, ignore_errors=True)
        return

    # ---------------------------------------------------------------- B200 arm
    import numpy as np
    import torch
    import torch.distributed as dist
    import carmel_b200 as cb

    assert torch.cuda.is_available(), "bench.py needs a GPU (carmel_b200 has no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shared = os.path.join(tempfile.gettempdir(), f"cb200_bench_{os.environ.get('MASTER_PORT', 'single')}_{os.getppid() if world > 1 else os.getpid()}")
    if rank == 0:
        shutil.rmtree(shared, ignore_errors=True)
        w = make_workload(a.workload, a.scale * world, shared)
        if a.unlock_lm and a.workload == "cipher":
            lm = w["files"][1]
            txt = open(lm).read().replace("!))", "))")
            open(lm, "w").write(txt)
            config["workload"] += " (LM unlocked: transitions trainable)"
        json.dump(w, open(os.path.join(shared, "workload.json"), "w"))
    if world > 1:
        dist.barrier()
    w = json.load(open(os.path.join(shared, "workload.json")))

    stream = torch.cuda.Stream()
    reduce_tensor = {}

    def allreduce(ptr, n):  # fp64 sum over ranks, in place, on the library's stream
        t = reduce_tensor.get((ptr, n))
        if t is None:
            # wrap the library's device buffer as a torch tensor (no copy)
            class _Arr:
                __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
            t = torch.as_tensor(_Arr(), device=torch.device("cuda", local))
            reduce_tensor[(ptr, n)] = t
        with torch.cuda.stream(stream):
            dist.all_reduce(t)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def measure(no_dense: bool) -> dict:
        """build the job, run warm-up + timed steps + the e2e loop; returns everything rank 0 prints"""
        argv = list(w["argv"])
        extra = ["-q", f"--gpu={local}"]
        if a.precision == 32:
            extra.append("--float")
        if a.space == "scaled":
            extra.append("--scaled")
        if no_dense:
            extra.append("--no-dense")
        if world > 1:
            extra.append(f"--shard={rank}/{world}")
        t_build = time.time()
        job = cb.Job(extra + argv, allreduce=allreduce if world > 1 else None)
        # the job's context must run on our stream before lattices are uploaded
        ctx = job.prepare()
        ctx.set_stream(stream.cuda_stream)
        if os.environ.get("CML_BENCH_NO_COUNTS"):  # profiling experiment only: the sweep without its count REDs
            ctx.set_option(cb.OPT_NO_COUNTS, 1)
        t_build = time.time() - t_build
        info = job.stats()
        dense = ctx.dense_stats()
        is_dense = dense["sequences"] > 0
        arcs_local, states_local = info["trellis_arcs"], info["trellis_states"]
        tot = torch.tensor([arcs_local, states_local, info["examples"], dense["positions"]], dtype=torch.float64,
                           device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        arcs_total, states_total, ex_total, pos_total = (float(x) for x in tot.tolist())
        # the lattice path streams > L2 of topology per iteration; the dense-state path's whole working set
        # (symbols + alpha rows) fits in L2, so L2 is flushed between its timed steps (outside the event pairs)
        flush = is_dense

        def step():
            ctx.estimate_launch()
            if world > 1:
                p, n = ctx.reduce_buffer()
                allreduce(p, n)
            r = ctx.estimate_finish()
            ctx.maximize(1.0)
            return r

        def timed(fn, steps):
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
            else:
                for _ in range(steps):
                    with torch.cuda.stream(stream):
                        flush_buf.fill_(1)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        fn()
                        e1.record()
                    torch.cuda.synchronize()
                    total += e0.elapsed_time(e1)
            if world > 1:
                dist.barrier()
            ms = torch.tensor([total], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        for _ in range(a.warmup):
            step()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = ctx.launch_count()
        fb_ms = []

        def step_and_sample():
            step()
            fb_ms.append(ctx.last_fb_time_ms())

        ms = timed(step_and_sample, a.steps)
        launches = ctx.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        value = arcs_total * a.steps / (ms / 1e3)

        # ---- e2e: host buffers through the C ABI every step
        n_params, n_arcs = info["n_params"], info["n_arcs"]
        h_params = torch.empty(n_params, dtype=torch.float64).pin_memory()
        n_slots = ctx.count_slots()
        h_counts = torch.empty(n_slots, dtype=torch.float64).pin_memory()
        ctx.get_params_ptr(h_params.data_ptr())

        def e2e_step():
            ctx.set_params_ptr(h_params.data_ptr())          # H2D: parameter vector
            ctx.estimate_launch()
            if world > 1:
                p, n = ctx.reduce_buffer()
                allreduce(p, n)
            ctx.estimate_finish()                             # D2H: likelihood scalars
            ctx.get_counts_ptr(h_counts.data_ptr(), n_slots)  # D2H: expected counts (one per count slot)
            ctx.maximize(1.0)
            ctx.get_params_ptr(h_params.data_ptr())           # D2H: new parameters

        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, a.steps)
        e2e = {"value": arcs_total * a.steps / (ms_e2e / 1e3), "unit": unit, "h2d_bytes_per_step": 8 * n_params,
               "d2h_bytes_per_step": 8 * n_params + 8 * n_slots + 24, "ms_per_step": ms_e2e / a.steps,
               "lattices": ("never materialised (dense-state view): symbol sequences resident; one-time host prep + upload "
                            if is_dense else "resident (carmel -: derivation-cache semantics); one-time host build + "
                            "flatten + upload ") + f"took {t_build:.2f}s on this rank"}
        res = {"is_dense": is_dense, "value": value, "ms": ms, "e2e": e2e, "launches": int(launches), "clocks": clocks,
               "totals": {"examples": ex_total, "trellis_arcs": arcs_total, "trellis_states": states_total,
                          "n_params": n_params, "n_arcs": n_arcs}, "count_slots": n_slots, "l2_flush": flush}
        if rank == 0:
            peaks, which = measured_peaks()
            rs = a.precision // 8
            k_ms = sum(m for m, _ in fb_ms) / len(fb_ms)
            n_k = fb_ms[0][1]
            if is_dense and dense["kernel"] == "sparse":  # (dense / dense_tc: the elif below)
                # sparse-emission kernel: per position one symbol, K alpha values written + read, one exponent written +
                # read; emission rows / transition matrix come from L2 / shared memory
                res["totals"]["positions"] = pos_total
                K = dense["k"]
                bytes_pos = 2.0 + 2.0 * K * rs + 8.0
                ach = bytes_pos * dense["positions"] / (k_ms / 1e3) / 1e9
                peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
                lattice_bytes = (16.0 + 2.0 * rs * (states_local / max(1, arcs_local))) * arcs_local
                res["roofline"] = {
                    "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": measured_traffic("k_fb_sparse") if a.workload == "hmm" and a.precision == 64 and a.scale == 1 else None,
                    "peak_source": which,
                    "kernel": "k_fb_sparse (forward + backward + counts, one sequence per lane, lattices never "
                              "materialised; 1 launch per iteration)",
                    "kernel_ms": k_ms, "algorithmic_bytes_per_position": bytes_pos,
                    "positions_per_launch": dense["positions"], "kernel_share_of_step": k_ms / (ms / a.steps),
                    "lattice_equivalent": {
                        "note": "what the same arcs would cost as streamed lattice records (SURVEY 8d: 16 B/arc + alpha): "
                                "this kernel does not read them, so the figure can exceed the HBM peak",
                        "algorithmic_bytes": lattice_bytes, "achieved_gbs": lattice_bytes / (k_ms / 1e3) / 1e9,
                        "frac_of_hbm_peak": lattice_bytes / (k_ms / 1e3) / 1e9 / peak}}
                res["layout"] = dense
            elif is_dense:
                res["totals"]["positions"] = pos_total
                S = dense["n_states"]  # useful flops (S x S mat-vecs); the kernel computes on 32 padded lanes
                products = 3 if dense["t_slots"] else 2  # alpha, beta (+ xi when transitions are trainable)
                flops = 2.0 * products * S * S * dense["positions"]
                tf32_peak = float(peaks.get("bf16_tflops", 1665.0)) / 2.0
                ach = flops / (k_ms / 1e3) / 1e12
                hbm_bytes = dense["positions"] * (2.0 * 32 * rs + 8 + 2 + 2)  # alpha row written + read, exponents, symbol twice
                res["roofline"] = {
                    "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                    "traffic": measured_traffic("k_fb_dense") if a.workload == "cipher" and a.precision == 64 and a.scale == 1 else None,
                    "peak_source": f"{which} bf16 dense peak / 2 (TF32 rate; no TF32 entry in MEASURED_PEAKS.json)",
                    "kernel": ("k_dense_tc<fwd> + k_dense_tc<bwd> + k_dense_tc_counts (3xTF32 mma.sync sweeps, 16 sequences per "
                               "warp; 3 launches per iteration)" if dense["kernel"] == "dense_tc" else
                               "k_fb_dense (forward + backward + counts over never-materialised lattices, 1 launch per iteration)"),
                    "kernel_ms": k_ms, "flops_per_position": 2.0 * products * S * S, "positions_per_launch": dense["positions"],
                    "kernel_share_of_step": k_ms / (ms / a.steps),
                    "note": ("tensor-core path: each position step of 16 sequences is a [16x32].[32x32] product in 3xTF32"
                             if dense["kernel"] == "dense_tc" else
                             "CUDA-core FMA kernel (fp32/fp64), one warp per sequence: the step is a serial chain of "
                             "line_len dependent 32x32 products, latency bound at this corpus size (0.3 GFLOP per iteration); "
                             "reported against the tensor peak as SURVEY 8(d) asks for the dense case; fp32 corpora of >= 16,384 "
                             "sequences take the 3xTF32 tensor-core kernels (k_dense_tc)"),
                    "hbm_equivalent": {"algorithmic_bytes": hbm_bytes, "achieved_gbs": hbm_bytes / (k_ms / 1e3) / 1e9,
                                       "peak_gbs": float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))}}
                res["layout"] = dense
            else:
                bytes_per_arc = 16.0 + 2.0 * rs * (states_local / max(1, arcs_local))
                achieved = bytes_per_arc * arcs_local / (k_ms / 1e3) / 1e9
                peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
                res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                   "traffic": (measured_traffic("k_fb_lane" if a.workload == "hmm" else "k_fb_ell_cipher")
                                               if a.precision == 64 and a.scale == 1 else None),
                                   "peak_source": which, "kernel": "k_fb_* (forward + backward + counts, "
                                   f"{n_k} launch(es) per iteration)", "kernel_ms": k_ms,
                                   "algorithmic_bytes_per_arc": bytes_per_arc, "arcs_per_launch_set": arcs_local,
                                   "kernel_share_of_step": k_ms / (ms / a.steps)}
                res["layout"] = {**ctx.layout_stats(), **ctx.lane_stats()}
        job.close()
        return res

    main_res = measure(a.no_dense)
    sparse_res = None
    if main_res["is_dense"] and not a.no_sparse_leg:
        # the same corpus on the lattice path (the sparse-trellis kernels the HBM roofline target is about)
        flush_buf.fill_(0)
        sparse_res = measure(True)

    if rank == 0:
        if main_res["is_dense"]:
            config["workload"] = config["workload"].replace("sparse layered-CSR path", "dense-state path")
            if a.workload == "hmm":
                config["workload"] += ", dense-state path (sparse emission rows, one sequence per lane)"
            config["l2"] = ("dense-state working set (symbols + alpha rows) may fit in L2: L2 flushed (256 MB write) between "
                            "timed steps, each step timed with its own CUDA-event pair; the sparse_path leg streams "
                            "~1 GB of lattice records per iteration (> 126 MB L2)")
        try:
            n_sample = 160 if a.workload == "cipher" else 16000
            cpu = cpu_oracle_throughput(w, n_sample, max(1, os.cpu_count() or 1), budget_s=15.0)
        except Exception as ex:  # the bench line must still be printed
            cpu = {"value": None, "unit": unit, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        line = {"metric": metric, "value": main_res["value"], "unit": unit, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": main_res["ms"] / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64" if a.precision == 64 else "f32", "data": "synthetic", "config": config,
                "roofline": main_res["roofline"], "cpu_baseline": cpu, "e2e": main_res["e2e"],
                "gpu_launches": main_res["launches"], "clocks": main_res["clocks"], "totals": main_res["totals"],
                "layout": main_res["layout"], "count_slots": main_res["count_slots"]}
        if sparse_res is not None:
            line["sparse_path"] = {"value": sparse_res["value"], "unit": unit, "ms_per_step": sparse_res["ms"] / a.steps,
                                   "roofline": sparse_res["roofline"], "e2e": sparse_res["e2e"],
                                   "layout": sparse_res["layout"], "count_slots": sparse_res["count_slots"],
                                   "note": "same corpus with --no-dense: lattices materialised and streamed (level-sliced ELL "
                                           "kernel for the cipher, lane-per-lattice kernel for hmm)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
    if rank == 0 and not a.keep:
        shutil.rmtree(shared, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
