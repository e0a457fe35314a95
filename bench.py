#!/usr/bin/env python
"""bench.py -- leapfrog gradient evaluations per second of the MAGI hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # own arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle C port, all host cores)

One "step" = one logdensity_and_gradient pass over one batch of chains (BASELINE configs[1]: FitzHugh-Nagumo,
n=201, band 20, 4096 chains per GPU).  Chains shard across ranks with no data-path collective (weak scaling).
The timed block of --steps launches is repeated --repeats times (median reported, min / max beside it): a single block of
20 launches is under a millisecond.  The own arm also measures, as extra keys of the same JSON line, BASELINE configs 3
(`dense`: Lotka-Volterra n=1281, band n-1, 2048 chains), 4 (`setup`: Lorenz-96 D=64 n=2001 device GP setup, both modes)
and 5 (`cfg5`: FN 65 536 chains split over the ranks, on-device HMC, one all-gather of the draws, R-hat / ESS);
--sections none skips them.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "fp64_tflops": 37.1, "fp64_src": "profiles/fp64_peaks_r01.json"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        peaks["hbm_gbs"] = float(m["hbm_gbs"]); peaks["hbm_src"] = "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")) as f:
            m = json.load(f)
        peaks["fp64_tflops"] = float(m["fp64_dmma_tflops"])
        peaks["fp64_src"] = "measured FP64 DMMA.8x8x4 peak (profiles/fp64_peaks_r01.json; MEASURED_PEAKS.json has no FP64 figure)"
    except Exception:
        pass
    return peaks


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML during the timed region."""
    def __init__(self, index):
        self.samples, self.reasons, self.maxmhz, self._stop = [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxmhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and not nm.endswith("None") and not nm.endswith("All"):
                        self.reasons.add(nm.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.maxmhz, "reasons": sorted(x for x in self.reasons if x not in ("GpuIdle",)),
                "samples": len(s)}


def cpu_arm(work, tables, steps, warmup, chains, nthreads=0, min_seconds=0.0):
    """Times the oracle's C restatement of the reference path (all host threads) on the same inputs and band tables."""
    from oracle import c_oracle, magi_oracle as mo
    covs = []
    for d in range(work["D"]):
        g = mo.GPCov(bandsize=work["bandsize"], tvec=work["tvec"])
        g.CinvBand, g.mphiBand, g.KinvBand = tables[d]
        covs.append(g)
    mid = {"fn": mo.MODEL_FN, "lv": mo.MODEL_LV}[work["model"]]
    tgt = mo.make_target(work["yobs"], covs, mid, work["sigma_init"], work["beta"], False)
    params = work["params"][:chains]
    for _ in range(max(1, warmup)):
        ll, g = c_oracle.batched(tgt, params, nthreads)
    t0 = time.perf_counter()
    done = 0
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        ll, g = c_oracle.batched(tgt, params, nthreads)
        done += 1
    dt = time.perf_counter() - t0
    return dict(evals_per_s=chains * done / dt, seconds=dt, passes=done, cores=(nthreads or c_oracle.num_threads()), ll=ll, grad=g)


def host_cores():
    """Every core this process may run on: torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2] if len(xs) % 2 else 0.5 * (xs[len(xs) // 2 - 1] + xs[len(xs) // 2])


def section_dense(pkg, synthetic, torch, dev, local, peaks, reps=10):
    """BASELINE config 3, dense mode: Lotka-Volterra n=1281, band n-1 (the FP64 DMMA GEMM path), 2048 chains, one GPU."""
    w = synthetic.make_workload("lv1281", 2048)
    n, D, nch = w["n"], w["D"], 2048
    tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.lv_system(), w["sigma_init"], bandsize=n - 1, jitter=1e-6,
                                    setup_mode="stable", device=local, max_chains=nch)
    p = torch.from_numpy(w["params"]).to(dev); g = torch.empty_like(p); ll = torch.empty(nch, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st)
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    l0 = tg.launch_count()
    ts = []
    for _ in range(reps):
        flush.zero_()                                       # 256 MB written: nothing of the previous pass is left in the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    launches = (tg.launch_count() - l0) // reps
    clocks = sampler.stop()
    ms = _median(ts)
    flops = synthetic.algorithmic_flops_per_eval(n, D, n - 1)
    tf = nch * flops / (ms * 1e-3) * 1e-12
    finite = bool(torch.isfinite(ll).all().item())
    tg.close()
    return {"config": {"workload": "lv1281 dense: lv n=%d D=%d band=%d (= n-1) matern52 jitter=1e-6, %d chains, one GPU" % (n, D, n - 1, nch)},
            "metric": "leapfrog grad evals/sec (all chains)", "value": nch / (ms * 1e-3), "unit": "evals/s", "ms_per_step": ms,
            "ms_min": min(ts), "ms_max": max(ts), "repeats": reps, "gpu_launches": int(launches), "ll_finite": finite, "clocks": clocks,
            "l2": "256 MB flush buffer written between timed passes",
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["fp64_tflops"],
                         "kernel": "gemm_f64_dmma_streamk_kernel (4 GEMMs per evaluation) + 2 pointwise kernels",
                         "algorithmic_flops_per_eval": flops, "peak_source": peaks["fp64_src"]}}


def section_smallband(pkg, synthetic, torch, dev, local, peaks, reps=10):
    """The north-star HBM regime: FN n=201 with band half-widths 1, 2 and 4, 65 536 chains on one GPU (427 MB of state + gradient per
    pass: nothing fits the 126 MB L2).  Large batches of these bands run on K1-narrow (csrc/narrow_kernel.cuh)."""
    nch = 65536
    out = {"config": {"workload": "fn201 small bands: fn n=201 D=2 k=3 band in (1, 2, 4) matern52 jitter=1e-6, %d chains, one GPU" % nch},
           "l2": "427 MB of parameters + gradient per pass > 126 MB L2, no explicit flush", "bands": {}}
    w = synthetic.make_workload("fn201", nch)
    n, D = w["n"], w["D"]
    p = torch.from_numpy(w["params"]).to(dev); g = torch.empty_like(p); ll = torch.empty(nch, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    P = p.shape[1]
    sampler = ClockSampler(local); sampler.start()
    for b in (1, 2, 4):
        tg = pkg.MagiTarget.from_config(w["yobs"], w["tvec"], w["phi"], pkg.fn_system(), w["sigma_init"], bandsize=b, jitter=1e-6,
                                        setup_mode="stable", device=local, max_chains=8)
        for _ in range(3):
            tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st)
        torch.cuda.synchronize()
        l0 = tg.launch_count(); ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tg.logdensity_and_gradient_batched_dev(nch, p.data_ptr(), ll.data_ptr(), g.data_ptr(), st); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = _median(ts)
        flops, abytes = synthetic.algorithmic_flops_per_eval(n, D, b), synthetic.algorithmic_bytes_per_eval(P)
        gbs, tf = nch * abytes / (ms * 1e-3) * 1e-9, nch * flops / (ms * 1e-3) * 1e-12
        out["bands"][str(b)] = {"ms_per_step": ms, "ms_min": min(ts), "ms_max": max(ts), "value": nch / (ms * 1e-3), "unit": "evals/s",
                                "gpu_launches": int((tg.launch_count() - l0) // reps), "ll_finite": bool(torch.isfinite(ll).all().item()),
                                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                             "kernel": "narrow_logpost_kernel", "algorithmic_bytes_per_eval": abytes, "peak_source": peaks["hbm_src"]},
                                "roofline_fp64": {"bound": "fp64", "achieved": tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["fp64_tflops"],
                                                  "algorithmic_flops_per_eval": flops}}
        tg.close()
    out["clocks"] = sampler.stop()
    return out


def section_setup(pkg, synthetic, torch, local, peaks):
    """BASELINE config 4: Lorenz-96 D=64, n=2001: device GP setup (covariance build, blocked Cholesky, inverses, GEMMs, bands)."""
    rng = np.random.default_rng(20251018 + 3)
    n, D = 2001, 64
    tvec = np.linspace(0.0, 20.0, n)
    phi = np.stack([rng.uniform(10, 20, D), rng.uniform(0.2, 0.4, D)])
    Y = np.full((n, D), np.nan); Y[::10] = 8.0 + rng.normal(size=(len(tvec[::10]), D))
    out = {"config": {"workload": "lorenz96 setup: D=%d n=%d band=20 matern52 jitter=1e-6, 64 distinct (variance, lengthscale) pairs, one GPU" % (D, n)},
           "modes": {}}
    sampler = ClockSampler(local); sampler.start()
    for mode in ("stable", "reference_order"):
        best = None
        for rep in range(2):                                # the first create of a process also pays the allocator's first 20 GB
            torch.cuda.synchronize(); t0 = time.perf_counter()
            tg = pkg.MagiTarget.from_config(Y, tvec, phi, pkg.get_ode_system("lorenz96", D), np.full(D, 0.5), bandsize=20, jitter=1e-6,
                                            setup_mode=mode, device=local)
            wall = time.perf_counter() - t0
            kernel_ms, alloc_ms = tg.setup_timing()
            rep_piv = [tg.setup_status(d) for d in range(D)]
            if mode == "stable" and rep == 1:               # the 64-chain evaluation of config 4 (band products on DMMA tiles + pointwise kernels)
                nch = 64
                prm = np.concatenate([8.0 + rng.normal(size=(nch, n * D)), 8.0 + 0.1 * rng.normal(size=(nch, 1)), np.log(0.5) + 0.1 * rng.normal(size=(nch, D))], axis=1)
                dev = torch.device("cuda", local)
                pt = torch.from_numpy(prm).to(dev); gt = torch.empty_like(pt); lt = torch.empty(nch, dtype=torch.float64, device=dev)
                stc = torch.cuda.current_stream().cuda_stream
                flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
                for _ in range(2):
                    tg.logdensity_and_gradient_batched_dev(nch, pt.data_ptr(), lt.data_ptr(), gt.data_ptr(), stc)
                torch.cuda.synchronize()
                l0 = tg.launch_count(); tsv = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); tg.logdensity_and_gradient_batched_dev(nch, pt.data_ptr(), lt.data_ptr(), gt.data_ptr(), stc); e1.record()
                    torch.cuda.synchronize(); tsv.append(e0.elapsed_time(e1))
                ms_ev = _median(tsv)
                fl_ev = float(synthetic.algorithmic_flops_per_eval(n, D, 20))      # 4 D 2 nnz + pointwise
                tf_ev = nch * fl_ev / (ms_ev * 1e-3) * 1e-12
                out["evaluation"] = {"config": {"workload": "lorenz96 D=%d n=%d band=20, %d chains, one GPU" % (D, n, nch)}, "ms_per_step": ms_ev,
                                     "value": nch / (ms_ev * 1e-3), "unit": "evals/s", "gpu_launches": int((tg.launch_count() - l0) // 5),
                                     "ll_finite": bool(torch.isfinite(lt).all().item()),
                                     "roofline": {"bound": "tensor", "achieved": tf_ev, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": tf_ev / peaks["fp64_tflops"],
                                                  "kernel": "band_product_kernel (4 launches) + 2 pointwise kernels", "algorithmic_flops_per_eval": fl_ev}}
                del pt, gt, lt, flush
            tg.close()
            if best is None or kernel_ms < best["kernel_seconds"] * 1e3:
                flop = (5 if mode == "stable" else 6) * float(n) ** 3 * D
                tf = flop / (kernel_ms * 1e-3) * 1e-12
                best = {"kernel_seconds": kernel_ms * 1e-3, "cudaMalloc_seconds": alloc_ms * 1e-3, "create_wall_seconds": wall,
                        "nominal_TFLOP": flop * 1e-12, "repaired_pivots": [int(sum(r[0] for r in rep_piv)), int(sum(r[1] for r in rep_piv))],
                        "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["fp64_tflops"],
                                     "kernel": "gemm_f64_dmma_big_kernel (Cholesky trailing updates, triangular products, m and K products)",
                                     "peak_source": peaks["fp64_src"],
                                     "note": "nominal %dn^3 flop per dimension; device time from the covariance build to the band tables (CUDA events), allocations excluded" % (5 if mode == "stable" else 6)}}
        out["modes"][mode] = best
    out["clocks"] = sampler.stop()
    return out


def section_cfg5(pkg, synthetic, torch, dist, dev, local, rank, world, chains_total, iters, leapfrog=10):
    """BASELINE config 5: FN n=201, chains_total chains split over the ranks, on-device HMC (no hot-path collective; the
    warm-up's pooled statistics and the final all-gather of the draws run in-library over NCCL), R-hat / ESS on rank 0."""
    from manifold_constrained_gaussian_process_inference_b200 import distributed as Dm
    first, n_local = Dm.shard_chains(chains_total, rank, world)
    work = synthetic.make_workload("fn201", chains_total, rank=0)     # the same global population on every rank; each keeps its slice
    tg = pkg.MagiTarget.from_config(work["yobs"], work["tvec"], work["phi"], pkg.fn_system(), work["sigma_init"], bandsize=20, jitter=1e-6,
                                    setup_mode="stable", device=local, max_chains=n_local)
    if world > 1:
        Dm.init_device_comm(tg)                              # NCCL communicator inside the library, warmed
    params = work["params"][first:first + n_local]
    n_adapt = iters // 2
    st = torch.cuda.current_stream().cuda_stream
    sampler = ClockSampler(local) if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.start()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    _, stt = pkg.run_hmc_sampler(tg, params, n_samples=iters, n_adapts=n_adapt, initial_step_size=0.002, n_leapfrog=leapfrog,
                                 seed=20251018 + 5, chain_id_offset=first, keep_on_device=True, n_chains_total=chains_total, stream=st)
    if world > 1:                                            # receive buffer of the all-gather: allocated outside the timed region
        from manifold_constrained_gaussian_process_inference_b200.samplers import hmc_draws_device_view
        _, ns_, nc_, ncol_ = hmc_draws_device_view(tg)
        gbuf = torch.empty((world, ns_, nc_, ncol_), dtype=torch.float64, device=dev)
    e1.record()
    if world > 1:
        full = Dm.allgather_draws_device(tg, stream=st, out=gbuf)      # first gather into these buffers: pays NCCL's per-buffer set-up
        ew0, ew1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        torch.cuda.synchronize(); dist.barrier()
        full = Dm.allgather_draws_device(tg, stream=st, out=gbuf)      # second call, untimed: NCCL settles at a message size within two
        torch.cuda.synchronize(); dist.barrier()                       # calls (tools/allgather_probe.py, 4 GPUs: 21.8, 3.9, 0.32, 0.27, 0.27 ms)
        ew0.record()
        full = Dm.allgather_draws_device(tg, stream=st, out=gbuf)      # the same gather again, warm
        ew1.record()
    else:
        full = Dm.device_draws_as_tensor(tg)
        e2.record()
    torch.cuda.synchronize()
    warm_gather_ms = ew0.elapsed_time(ew1) if world > 1 else 0.0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), float(stt["grad_evals"]), warm_gather_ms], dtype=torch.float64, device=dev)
    tmax, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    res = None
    if rank == 0:
        d = full[:, :: max(1, full.shape[1] // 512)][:, :512].cpu().numpy()      # R-hat / ESS on 512 chains spread over all ranks' shards
        names = ["theta_a", "theta_b", "theta_c", "sigma_1", "sigma_2", "lp"]
        summ = pkg.diagnostics.summarize(d, names=names)
        sample_s, gather_first_ms, evals, gather_ms = float(tmax[0]) * 1e-3, float(tmax[1]), float(tsum[2]), float(tmax[3])
        gathered_bytes = int(full.numel() * 8)
        res = {"config": {"workload": "fn201 cfg5: fn n=201 D=2 k=3 band=20, %d chains over %d GPU(s), HMC %d iterations (%d warm-up) x %d leapfrog steps" % (chains_total, world, iters, n_adapt, leapfrog),
                          "chains_total": chains_total, "chains_per_gpu": n_local, "scaling": "strong"},
               "metric": "leapfrog grad evals/sec (all chains)", "value": evals / sample_s, "unit": "evals/s", "n_gpus": world,
               "sample_seconds": sample_s, "grad_evals": evals, "allgather_ms": gather_ms, "allgather_first_call_ms": gather_first_ms,
               "allgather_bytes": gathered_bytes,
               "allgather_GBps": (gathered_bytes / (gather_ms * 1e-3) * 1e-9) if world > 1 and gather_ms > 0 else None,
               "draws": list(full.shape), "accept_rate_median": float(np.median(stt["accept_rate"])),
               "posterior_mean": {nm: r["mean"] for nm, r in zip(names, summ)}, "rhat": {nm: r["rhat"] for nm, r in zip(names, summ)},
               "converged": bool(max(r["rhat"] for r in summ) < 1.05),
               "convergence_note": "a run of this length is a throughput measurement (theta mixes along a narrow ridge of the (X, theta) posterior); the converged "
                                   "run of the same sampler -- 1024 chains, 100 000 iterations x 100 leapfrog steps, 248 s on one B200, split R-hat <= 1.03 on every "
                                   "parameter -- is profiles/cfg5_convergence_r02.jsonl (tools/cfg5_convergence.py)",
               "ess_bulk_512_chains": {nm: r["ess_bulk"] for nm, r in zip(names, summ)}, "theta_true": [0.2, 0.2, 3.0], "clocks": clocks,
               "how": "CUDA events on the sampler's stream around the whole run (warm-up included) and around the all-gather; max over ranks"}
    tg.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fn201")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the workload's)")
    ap.add_argument("--bandsize", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--repeats", type=int, default=20, help="repetitions of the timed block of --steps launches")
    ap.add_argument("--sections", default="dense,setup,smallband,cfg5", help="extra measurements of the own arm (comma list or 'none')")
    ap.add_argument("--cfg5-chains", type=int, default=65536)
    ap.add_argument("--cfg5-iters", type=int, default=240, help="HMC iterations of the cfg5 section (half of them warm-up)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    import manifold_constrained_gaussian_process_inference_b200 as pkg
    from manifold_constrained_gaussian_process_inference_b200 import synthetic

    chains = args.chains or synthetic.CONFIGS[args.workload]["chains"]
    work = synthetic.make_workload(args.workload, chains, rank=rank, bandsize=(args.bandsize if args.bandsize >= 0 else None))
    n, D, k, b = work["n"], work["D"], work["k"], work["bandsize"]
    P = n * D + k + D
    config = {"workload": "%s: %s n=%d D=%d k=%d band=%d matern52 jitter=1e-6, %d chains/GPU, sigma sampled" % (args.workload, work["model"], n, D, k, b, chains),
              "chains_per_gpu": chains, "n_times": n, "bandsize": b, "sharding": "chains across ranks, no data-path collective"}     # identical in both arms

    if args.impl == "reference":
        if rank != 0:
            return 0
        # band tables for the CPU arm: built by the oracle itself (this arm runs none of our kernels)
        from oracle import magi_oracle as mo
        tables = []
        for d in range(D):
            g = mo.calculate_gp_covariances(mo.MATERN52, work["phi"][:, d], work["tvec"], b, jitter=1e-6, setup_mode="stable")
            tables.append((g.CinvBand, g.mphiBand, g.KinvBand))
        res = cpu_arm(work, tables, args.steps, warmup, chains, nthreads=host_cores())
        val = res["evals_per_s"]
        out = {"impl": "reference", "metric": "leapfrog grad evals/sec (all chains)", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
               "steps": res["passes"], "warmup": warmup, "ms_per_step": res["seconds"] / res["passes"] * 1e3, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": val, "unit": "evals/s", "cores": res["cores"], "kind": "port",
                                "sample": "%d passes over the %d-chain batch, C restatement of likelihoods.jl + interface.jl, OpenMP over chains" % (res["passes"], chains)},
               "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    tg = pkg.MagiTarget.from_config(work["yobs"], work["tvec"], work["phi"], pkg.get_ode_system(work["model"]), work["sigma_init"],
                                    prior_temperature=work["beta"], sigma_is_fixed=False, kernel="matern52", bandsize=b, jitter=1e-6,
                                    setup_mode="stable", device=local, max_chains=chains)
    assert tg.dimension() == P
    # rotating input/output sets so that consecutive steps never find their operands in the 126 MB L2
    bytes_per_set = chains * (2 * P + 1) * 8
    nsets = max(2, int(np.ceil(2.5 * 126e6 / bytes_per_set)))
    host = torch.from_numpy(work["params"])
    psets = [(host.to(dev) + (1e-6 * s)).contiguous() for s in range(nsets)]
    gsets = [torch.empty_like(psets[0]) for _ in range(nsets)]
    lsets = [torch.empty(chains, dtype=torch.float64, device=dev) for _ in range(nsets)]
    l2_note = "rotating %d input/output sets (%.0f MB) > 126 MB L2, no explicit flush" % (nsets, nsets * bytes_per_set / 1e6)
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        s = i % nsets
        tg.logdensity_and_gradient_batched_dev(chains, psets[s].data_ptr(), lsets[s].data_ptr(), gsets[s].data_ptr(), stream)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = tg.launch_count()
    reps = max(1, args.repeats)
    rep_ms = []
    for r in range(reps):                                   # each repetition: barrier + sync, EXACTLY --steps launches between two events
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(warmup + r * args.steps + i)
        e1.record()
        torch.cuda.synchronize()
        rep_ms.append(e0.elapsed_time(e1))
    launches = (tg.launch_count() - l0) // reps
    t = torch.tensor(rep_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # per repetition: the slowest rank
    rep_ms = sorted(float(x) for x in t.cpu())
    ms = rep_ms[len(rep_ms) // 2] if len(rep_ms) % 2 else 0.5 * (rep_ms[len(rep_ms) // 2 - 1] + rep_ms[len(rep_ms) // 2])

    # ---- e2e: the public host API (HOST buffers in, HOST buffers out; copies inside the timed region) ----
    # The pinned buffers are allocated, and the calls are made, from the CPUs NVML reports as local to this GPU (what
    # `numactl` would do): on a multi-socket box a process that lands on the far socket sees half the PCIe bandwidth.
    old_affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (ncpu + 63) // 64)
        cpus = {64 * wi + bit for wi, wd in enumerate(words) for bit in range(64) if (wd >> bit) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            old_affinity = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
    except Exception:
        old_affinity = None
    hp = torch.from_numpy(work["params"]).pin_memory()
    hg = torch.empty_like(hp).pin_memory()
    hl = torch.empty(chains, dtype=torch.float64).pin_memory()
    hp_np, hg_np, hl_np = hp.numpy(), hg.numpy(), hl.numpy()
    from manifold_constrained_gaussian_process_inference_b200 import _lib
    L = _lib.lib()

    def e2e_step():
        _lib.check(L.magi_logdensity_and_gradient_batched(tg._h, chains, _lib.as_dp(hp_np), _lib.as_dp(hl_np), _lib.as_dp(hg_np)))

    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    clocks = sampler.stop() if sampler else None
    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)          # the CPU baseline below uses every host core

    wanted = [x for x in args.sections.split(",") if x and x != "none"]
    extra = {}
    peaks = load_peaks()
    del psets, gsets, lsets
    torch.cuda.empty_cache()
    if rank == 0 and "dense" in wanted:
        extra["dense"] = section_dense(pkg, synthetic, torch, dev, local, peaks)
    if rank == 0 and "setup" in wanted:
        extra["setup"] = section_setup(pkg, synthetic, torch, local, peaks)
        torch.cuda.empty_cache()
    if rank == 0 and "smallband" in wanted:
        extra["smallband"] = section_smallband(pkg, synthetic, torch, dev, local, peaks)
        torch.cuda.empty_cache()
    if "cfg5" in wanted:
        if world > 1:
            dist.barrier()
        r5 = section_cfg5(pkg, synthetic, torch, dist, dev, local, rank, world, args.cfg5_chains, args.cfg5_iters)
        # the same flow with a chain count that fills whole waves of the machine on 1, 2, 4 and 8 GPUs (a block of the banded kernel holds
        # 32 chains, a wave 148 blocks: 8 x 2 x 4736 chains); 65 536 / 8 = 8192 chains per GPU are 1.73 waves, i.e. cost two
        r5a = section_cfg5(pkg, synthetic, torch, dist, dev, local, rank, world, 8 * 2 * 148 * 32, args.cfg5_iters)
        if rank == 0:
            extra["cfg5"] = r5
            extra["cfg5_wave_aligned"] = r5a
    if rank == 0:
        value = world * chains * args.steps / (ms * 1e-3)
        per_launch_s = ms * 1e-3 / args.steps
        flops = synthetic.algorithmic_flops_per_eval(n, D, b)
        abytes = synthetic.algorithmic_bytes_per_eval(P)
        tf = chains * flops / per_launch_s * 1e-12
        gbs = chains * abytes / per_launch_s * 1e-9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("workload") == args.workload and tj.get("chains") == chains:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        out = {"metric": "leapfrog grad evals/sec (all chains)", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
               "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": config, "gpu_launches": int(launches), "clocks": clocks,
               "timing": {"repeats": reps, "steps_per_repeat": args.steps, "ms_per_step_median": ms / args.steps, "ms_per_step_min": rep_ms[0] / args.steps,
                          "ms_per_step_max": rep_ms[-1] / args.steps, "l2": l2_note, "how": "CUDA events around each block of --steps launches, max over ranks per block, median over blocks"},
               "e2e": {"value": world * chains * e2e_steps / e2e_dt, "unit": "evals/s", "h2d_bytes_per_step": chains * P * 8,
                       "d2h_bytes_per_step": chains * (P + 1) * 8, "steps": e2e_steps, "api": "magi_logdensity_and_gradient_batched (pinned host buffers)"},
               "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["fp64_tflops"],
                            "traffic": traffic, "kernel": "banded_logpost_kernel", "peak_source": peaks["fp64_src"],
                            "algorithmic_flops_per_eval": flops,
                            "note": "band 20 is FP64-pipe bound (AI %.1f flop/B vs ridge %.1f): the binding roof is the FP64 tensor (DMMA) pipe" % (flops / abytes, peaks["fp64_tflops"] * 1e3 / peaks["hbm_gbs"])},
               "roofline_hbm": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                "algorithmic_bytes_per_eval": abytes, "peak_source": peaks["hbm_src"]}}
        if world == 1 and not args.no_cpu_baseline:
            tables = [(tg.get_band_table(d, "CinvBand"), tg.get_band_table(d, "mphiBand"), tg.get_band_table(d, "KinvBand")) for d in range(D)]
            res = cpu_arm(work, tables, 1, 1, chains, nthreads=host_cores(), min_seconds=args.cpu_seconds)
            out["cpu_baseline"] = {"value": res["evals_per_s"], "unit": "evals/s", "cores": res["cores"], "kind": "port",
                                   "sample": "%d passes over the %d-chain batch in %.1f s (C restatement of the reference loop, OpenMP over chains, same band tables)" % (res["passes"], chains, res["seconds"])}
            # parity spot check of the measured kernel against the CPU arm (not timed)
            ll0, _ = tg.logdensity_and_gradient_batched(work["params"][:64])
            out["parity_ll_max_rel_err_vs_cpu"] = float(np.max(np.abs(ll0 - res["ll"][:64]) / np.abs(res["ll"][:64])))
        out.update(extra)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
