#!/usr/bin/env python
"""bench.py -- natural-gradient CAVI iterations/second of the SVGP AnalyticSVI hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config auto|C2|C3|C4|C5]

Workloads (BASELINE.json `configs`, sizes of SURVEY 8d; one "step" = one update_parameters! call,
training/training.jl:140-158 of the reference; synthetic seeded data; K_mm fixed = optimiser=false semantics):
  C2  SVGP Logistic SqExp            n=1e6 D=32  m=512  B=8192   1 latent          (N = 1 default: the metric's configuration)
  C3  SVGP StudentT(3) Matern-3/2    n=1e7 D=64  m=1024 B=16384  1 latent          (extra evidence run, n cut to 2e6 rows)
  C4  LogisticSoftMax 8 classes      n=1e6 D=128 m=256  B=8192   8 latents (one per GPU at N = 8)
  C5  multi-output SVGP 64 x 64      n=1e6 D=32  m=512  B=8192   64 latents, 64/N per GPU   (N > 1 default)
`value` = latent-GP iterations per second over all ranks (= plain iterations/s for a single-latent model).  Multi-latent
configurations shard BY LATENT: the total work is fixed, so N > 1 lines say "scaling": "strong".

--impl reference: the fp64 NumPy/OpenBLAS restatement of the reference path (oracle/) on the host cores, all threads
(the Julia reference itself cannot run here: no Julia in the image); for C4 / C5 each step is a bounded sample of the
workload (a sub-model of the same shape with fewer latents), stated in `cpu_baseline.sample`.
"""
from __future__ import annotations

import os

# the CPU legs (reference arm, cpu_baseline, ELBO parity replay) use every host thread: torchrun exports OMP_NUM_THREADS=1,
# and OpenBLAS reads its environment when NumPy is first imported -- so this must precede `import numpy`
_NCPU = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
if int(os.environ.get("RANK", "0")) == 0:
    os.environ["OPENBLAS_NUM_THREADS"] = str(min(_NCPU, 64))
    os.environ["OMP_NUM_THREADS"] = str(min(_NCPU, 64))
    os.environ["MKL_NUM_THREADS"] = str(min(_NCPU, 64))

import argparse
import json
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "natural-gradient CAVI iters/sec, SVGP m=512 bs=8192"
CONFIGS = {
    "C2": dict(n=1_000_000, D=32, m=512, B=8192, Q=1, T=1),
    "C3": dict(n=2_000_000, D=64, m=1024, B=16384, Q=1, T=1),
    "C4": dict(n=1_000_000, D=128, m=256, B=8192, Q=8, T=1),
    "C5": dict(n=1_000_000, D=32, m=512, B=8192, Q=64, T=64),
}
WORKLOAD = {
    "C2": "C2: SVGP Logistic SqExp n=1e6 D=32 m=512 minibatch=8192",
    "C3": "C3 (extra evidence run): SVGP StudentT(3) Matern-3/2 n=2e6 (of 1e7) D=64 m=1024 minibatch=16384",
    "C4": "C4: LogisticSoftMax 8 classes SqExp n=1e6 D=128 m=256 per class minibatch=8192, latents sharded over the ranks",
    "C5": "C5: multi-output SVGP 64 Logistic tasks x 64 latent GPs SqExp n=1e6 D=32 m=512 minibatch=8192, latents sharded over the ranks",
}
REF_SAMPLE_LATENTS = {"C4": 8, "C5": 4}   # latents in the reference arm's bounded sample step


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        v = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(v)) if v else _NCPU
    except Exception:
        return _NCPU


def make_problem(cfg_name, n_lists, seed=0, q_limit=None):
    """seeded synthetic problem of one configuration; q_limit: only the first q latents / tasks (reference-arm sample)"""
    c = CONFIGS[cfg_name]
    n, D, m, B, Q, T = c["n"], c["D"], c["m"], c["B"], c["Q"], c["T"]
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D), dtype=np.float32)
    if cfg_name == "C3":
        y = (np.sin(X[:, 0]) + 0.5 * X[:, 1] + 0.1 * rng.standard_t(3.0, n)).astype(np.float64)
    elif cfg_name == "C4":
        K = Q if q_limit is None else q_limit
        y = (np.argmax(X @ rng.standard_normal((D, Q)).astype(np.float32)[:, :K] + 0.1 * rng.standard_normal((n, K), dtype=np.float32), axis=1) + 1).astype(np.int64)
    else:
        Tn = T if q_limit is None else q_limit
        W = rng.standard_normal((D, T)).astype(np.float32)
        ys = [np.where(X @ W[:, t] + 0.1 * rng.standard_normal(n, dtype=np.float32) >= 0, 1.0, -1.0) for t in range(Tn)]
        y = ys[0] if cfg_name == "C2" else ys
    if cfg_name == "C5":
        Qn = Q if q_limit is None else q_limit
        Z = [X[np.random.default_rng(100 + q).permutation(n)[:m]].astype(np.float64) for q in range(Qn)]
        A = np.random.default_rng(7).standard_normal((T, Q))[:Qn, :Qn]
        A = A / np.linalg.norm(A, axis=1, keepdims=True)
    else:
        Z = X[rng.permutation(n)[:m]].astype(np.float64)
        A = None
    mbs = np.stack([rng.choice(n, B, replace=False) for _ in range(n_lists)]).astype(np.int64)
    return X, y, Z, A, mbs


def flops_per_iter(B, m, D):
    """algorithmic FLOPs of one step for one latent (SURVEY 8d / BASELINE.md section 3)."""
    return 2.0 * B * m * D + 6.0 * B * m * m + 8.0 * B * m + (5.0 / 3.0) * m**3


def oracle_model(O, cfg_name, Z, A, q_limit=None):
    c = CONFIGS[cfg_name]
    D, B = c["D"], c["B"]
    sc = 1.0 / np.sqrt(D)
    if cfg_name == "C2":
        return O.SVGP(O.Kernel("sqexp", scale=sc), O.LogisticLikelihood(), O.AnalyticSVI(B), Z)
    if cfg_name == "C3":
        return O.SVGP(O.Kernel("matern32", scale=sc), O.StudentTLikelihood(3.0, 1.0), O.AnalyticSVI(B), Z)
    if cfg_name == "C4":
        return O.SVGP(O.Kernel("sqexp", scale=sc), O.LogisticSoftMaxLikelihood(q_limit or c["Q"]), O.AnalyticSVI(B), Z)
    Qn = q_limit or c["Q"]
    return O.MOSVGP(O.Kernel("sqexp", scale=sc), [O.LogisticLikelihood() for _ in range(Qn)], O.AnalyticSVI(B), Z, A)


def engine_model(agp, cfg_name, Z, A, precision, device, stream, shard):
    c = CONFIGS[cfg_name]
    D, B = c["D"], c["B"]
    sc = 1.0 / np.sqrt(D)
    kw = dict(precision=precision, device=device, stream=stream)
    if shard is not None and shard[1] > 1:
        kw["shard"] = shard
    if cfg_name == "C2":
        return agp.SVGP(agp.SqExponentialKernel() @ agp.ScaleTransform(sc), agp.LogisticLikelihood(), agp.AnalyticSVI(B), Z, **kw)
    if cfg_name == "C3":
        return agp.SVGP(agp.Matern32Kernel() @ agp.ScaleTransform(sc), agp.StudentTLikelihood(3.0, 1.0), agp.AnalyticSVI(B), Z, **kw)
    if cfg_name == "C4":
        return agp.SVGP(agp.SqExponentialKernel() @ agp.ScaleTransform(sc), agp.LogisticSoftMaxLikelihood(c["Q"]), agp.AnalyticSVI(B), Z, **kw)
    return agp.MOSVGP(agp.SqExponentialKernel() @ agp.ScaleTransform(sc), [agp.LogisticLikelihood() for _ in range(c["T"])], agp.AnalyticSVI(B), Z, A=A, **kw)


def config_dict(cfg_name):
    """identical in both arms (the driver compares the two `config` objects)"""
    return dict(workload=WORKLOAD[cfg_name], **CONFIGS[cfg_name])


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank, world, cfg_name):
    """CPU restatement of the reference path (oracle) on the host cores, rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import agp_oracle as O

    c = CONFIGS[cfg_name]
    K, W = args.steps, args.warmup
    qs = REF_SAMPLE_LATENTS.get(cfg_name)          # None: the whole model is one latent
    X, y, Z, A, mbs = make_problem(cfg_name, K + W, q_limit=qs)
    X64 = X.astype(np.float64)
    model = oracle_model(O, cfg_name, Z, A, q_limit=qs)
    state = None
    if W > 0:
        model, state = O.train(model, X64, y, W, minibatches=list(mbs[:W]))
    t0 = time.perf_counter()
    model, state = O.train(model, X64, y, K, minibatches=list(mbs[W:]), state=state)
    dt = time.perf_counter() - t0
    nlat = qs or 1
    v = nlat * K / dt
    sample = (f"{K} full iterations of the same workload" if qs is None else
              f"{K} iterations of a {qs}-latent sub-model of the same shape (the full model has {c['Q']} latents; the per-latent cost is "
              "identical and the O(T Q^2 B) mixing loops of the reference only grow with Q, so the sample favours the reference)")
    line = dict(metric=METRIC, value=v, unit="iters/s", n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=1e3 * dt / K,
                higher_is_better=True, scaling="weak" if c["Q"] == 1 else "strong", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=config_dict(cfg_name),
                cpu_baseline=dict(value=v, unit="iters/s", cores=blas_threads(), kind="port",
                                  sample=sample + " (NumPy/SciPy fp64 on OpenBLAS, all host threads)"),
                e2e=dict(value=v, unit="iters/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank, cfg_name):
    import ctypes as C

    import torch

    import agp_b200 as agp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    c = CONFIGS[cfg_name]
    n, D, m, B, Q = c["n"], c["D"], c["m"], c["B"], c["Q"]
    if Q % world:
        raise SystemExit(f"{cfg_name} has {Q} latents: not divisible by {world} ranks")
    K, W = args.steps, max(args.warmup, 3)
    n_lists = K + W
    X, y, Z, A, mbs = make_problem(cfg_name, n_lists)
    # a dedicated non-default stream: the engine launches on it, torch events / NCCL are ordered on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    model = engine_model(agp, cfg_name, Z, A, args.precision, local_rank, stream, (rank, world))
    # one API-level step initialises everything (upload, compute_K) and checks the error path
    agp.train(model, X, y, 1, minibatches=[mbs[0]])
    eng = model._eng
    lib = eng.lib
    L = agp._lib
    rho = n / B
    eng.ck(lib.agp_minibatches_upload(eng.model, mbs.ctypes.data_as(L.c_int64_p), n_lists, B, 0))

    def step_async():
        if world == 1:
            eng.ck(lib.agp_step_async(eng.model, None, B, 0, rho))
        else:
            agp.api._sharded_step(model, eng, None, B, rho)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peer = bool(getattr(model, "_peer", False))
    if args.graph and (world == 1 or peer):   # peer mode: the whole sharded step (incl. the NVLink exchange) is one CUDA graph
        eng.ck(lib.agp_use_graph(eng.model, 1))
    for _ in range(W):
        step_async()
    eng.ck(lib.agp_sync(eng.model))
    # ---------------- device-resident timing ----------------
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_async()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    eng.ck(lib.agp_sync(eng.model))
    launches = model.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = Q * K / (ms * 1e-3)
    elbo = agp.ELBO(model)
    post_timed = model.posterior(0) if rank == 0 else None       # (mu, Sigma, ..) of this rank's first latent after the timed run

    if args.timed_only:  # ncu launch-list runs: nothing but the timed loop (no per-kernel re-timing, no e2e, no CPU leg)
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, value=value, unit="iters/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                                  gpu_launches=int(launches), config=config_dict(cfg_name), note="--timed-only run (profiling aid, not a bench line)")), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- end-to-end through the host-buffer API (before the profiling passes disturb the pipeline) ----------------
    e2e = None
    Ke = min(K, 50)
    if world == 1 and Q == 1:
        x_dt = np.float32 if args.e2e_dtype == "f32" else np.float64
        xb = [torch.empty((B, D), dtype=torch.float32 if args.e2e_dtype == "f32" else torch.float64).pin_memory() for _ in range(Ke)]
        yb = [torch.empty((B,), dtype=torch.float64).pin_memory() for _ in range(Ke)]
        for i in range(Ke):
            idx = mbs[(W + i) % n_lists]
            xb[i].numpy()[:] = X[idx].astype(x_dt)
            yb[i].numpy()[:] = y[idx]
        mu = np.empty(m)
        eng.ck(lib.agp_use_graph(eng.model, 1 if args.graph else 0))   # public switch: the compute part of a host-batch step replays one CUDA graph
        pending = []   # lagged mode: ticket of the step whose result has not been read yet
        xd = L.DTYPE_F32 if args.e2e_dtype == "f32" else L.DTYPE_F64

        def e2e_step(i):
            arr = (C.c_void_p * 1)(yb[i].data_ptr())
            if args.e2e_mode == "lagged":
                # no per-step synchronisation: the result of step i-1 is read while step i runs, the H2D copy of step i overlaps step i-1
                tk = C.c_int64(0)
                eng.ck(lib.agp_step_batch_async(eng.model, C.c_void_p(xb[i].data_ptr()), xd, L.LAYOUT_ROWMAJOR, arr, L.Y_REAL, B, rho, C.byref(tk)))
                if pending:
                    eng.ck(lib.agp_result_wait(eng.model, pending.pop(), L.dptr(mu)))
                pending.append(tk.value)
                return
            eng.ck(lib.agp_step_batch(eng.model, C.c_void_p(xb[i].data_ptr()), xd, L.LAYOUT_ROWMAJOR, arr, L.Y_REAL, B, rho))
            eng.ck(lib.agp_get_posterior(eng.model, 0, L.dptr(mu), None, None, None))

        def e2e_drain():
            while pending:
                eng.ck(lib.agp_result_wait(eng.model, pending.pop(), L.dptr(mu)))
        for i in range(3):
            e2e_step(i)
        e2e_drain()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(Ke):
            e2e_step(i)
        e2e_drain()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        es = 4 if args.e2e_dtype == "f32" else 8
        e2e = dict(value=Ke / dt, unit="iters/s", h2d_bytes_per_step=B * D * es + B * 8, d2h_bytes_per_step=m * 8 + 4, steps=Ke,
                   call=(f"agp_step_batch_async(host x[B,D] {args.e2e_dtype}, host y[B]) + agp_result_wait(mu of the previous step): every step's result is read, one step late"
                         if args.e2e_mode == "lagged" else f"agp_step_batch(host x[B,D] {args.e2e_dtype}, host y[B]) + agp_get_posterior(mu)"))
    else:
        # latent-sharded / multi-latent models keep X resident (replicated); the public call is train(model, X, y, minibatches=...):
        # per step the HOST index list goes up (pinned) and the posterior mean of this rank's first latent comes back
        idx_host = [torch.from_numpy(mbs[(W + i) % n_lists].copy()).pin_memory() for i in range(Ke)]
        mu = np.empty(m)

        def e2e_step(i):
            ip = C.cast(idx_host[i].data_ptr(), L.c_int64_p)
            if world == 1:
                eng.ck(lib.agp_step_async(eng.model, ip, B, 0, rho))
            else:
                agp.api._sharded_step(model, eng, ip, B, rho)
            eng.ck(lib.agp_get_posterior(eng.model, 0, L.dptr(mu), None, None, None))
        eng.ck(lib.agp_use_graph(eng.model, 0))
        for i in range(2):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            e2e_step(i)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = dict(value=Q * Ke / float(tt.item()), unit="iters/s", h2d_bytes_per_step=B * 8, d2h_bytes_per_step=m * 8 + 4, steps=Ke,
                   call="step with a HOST minibatch index list (X, y resident and replicated) + agp_get_posterior(mu of the rank's first latent), per rank")
        if args.graph and (world == 1 or peer):
            eng.ck(lib.agp_use_graph(eng.model, 1))

    # ---------------- per-kernel phase timers (second pass; roofline) ----------------
    roof = None
    phases = {}
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    bf16_b, bf16_s = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    peak_burst = (bf16_b / 2.0) if bf16_b else 1590.0 / 2.0
    peak_sust = (bf16_s / 2.0) if bf16_s else peak_burst
    psrc = ("MEASURED_PEAKS.json bf16 / 2 (TF32 = half the bf16 rate): burst figure for kernels timed alone, sustained for the whole step; of measured"
            if bf16_b else "fallback 1.59 PFLOP/s bf16 / 2, of fallback")
    ql = Q // world
    step_s = ms / K * 1e-3
    fl_step = ql * flops_per_iter(B, m, D)
    whole = dict(seconds=step_s, algorithmic_flops=fl_step, achieved=fl_step / step_s / 1e12, peak=peak_sust, unit="TFLOP/s",
                 frac=fl_step / step_s / 1e12 / peak_sust, latents_per_rank=ql,
                 note="algorithmic FLOPs of SURVEY 8d per latent x latents on this rank / device-timed step")
    if world == 1:
        eng.ck(lib.agp_use_graph(eng.model, 0))
        eng.ck(lib.agp_profile_enable(eng.model, 1))
        Kp = min(K, 30)
        for _ in range(Kp):
            step_async()
        names = (C.c_char_p * 32)()
        msv = (C.c_double * 32)()
        lv = (C.c_int64 * 32)()
        nph = lib.agp_profile_read(eng.model, 32, names, msv, lv)
        eng.ck(lib.agp_profile_enable(eng.model, 0))
        tot = sum(msv[i] for i in range(nph)) / Kp
        for i in range(nph):
            phases[names[i].decode()] = dict(ms_per_step=msv[i] / Kp, launches_per_step=lv[i] / Kp, share_of_phase_sum=(msv[i] / Kp / tot) if tot > 0 else None)
        kern = {}
        traffic = {}
        for tf in (os.path.join(ROOT, "profiles", "r2", "ncu_traffic.json"), os.path.join(ROOT, "profiles", "r1", "ncu_traffic.json")):
            if os.path.exists(tf) and cfg_name == "C2":
                try:
                    traffic = json.load(open(tf)); traffic["_source"] = os.path.relpath(tf, ROOT)
                    break
                except Exception:
                    traffic = {}
        if Q == 1 and args.precision == "tf32x3":
            # per-kernel durations: `reps` back-to-back launches between one CUDA-event pair on the engine's stream (agp_time_kernel)
            reps = 50
            tk = {}
            for which, name in ((0, "kmat_knm"), (1, "gemm_v"), (2, "gemm_v_sigma"), (3, "gemm_gram")):
                v = C.c_double(0.0)
                eng.ck(lib.agp_time_kernel(eng.model, which, reps, C.byref(v)))
                tk[name] = v.value * 1e-3
            # algorithmic FLOPs: V = Knm L^-T and V X^T have a lower-triangular right operand (B m^2 each); the Gram product
            # U^T U is 2 B m^2 (SURVEY 8d counts it in full; only the upper tiles are executed).  3xTF32 executes 3x these.
            fl = {"gemm_v": 1.0 * B * m * m, "gemm_v_sigma": 1.0 * B * m * m, "gemm_gram": 2.0 * B * m * m}
            for k_ in fl:
                kern[k_] = dict(bound="tensor", seconds_per_launch=tk[k_], algorithmic_flops_per_launch=fl[k_], achieved=fl[k_] / tk[k_] / 1e12,
                                peak=peak_burst, unit="TFLOP/s", frac=fl[k_] / tk[k_] / 1e12 / peak_burst, share_of_step=tk[k_] / step_s,
                                tensor_pipe_frac_3xtf32=3 * fl[k_] / tk[k_] / 1e12 / peak_burst * ((m // 128 + 1) / (2.0 * (m // 128)) if k_ == "gemm_gram" else 1.0),
                                traffic=traffic.get(k_))
            byts = 4.0 * (B * D + m * D + B * m) + 8.0 * B
            hbm = peaks.get("hbm_gbs", 6650.0)
            kern["knm"] = dict(bound="hbm", kernel="knm_umma_kernel", seconds_per_launch=tk["kmat_knm"], achieved=byts / tk["kmat_knm"] / 1e9, peak=hbm,
                               unit="GB/s", frac=byts / tk["kmat_knm"] / 1e9 / hbm, algorithmic_bytes_per_launch=byts, share_of_step=tk["kmat_knm"] / step_s,
                               peak_source="MEASURED_PEAKS.json hbm_gbs, of measured" if "hbm_gbs" in peaks else "fallback 6.65 TB/s, of fallback",
                               traffic=traffic.get("kmat_knm"),
                               note="side stream (overlaps the tail); at C2 the 16.8 MB K_nm output stays in the 126 MB L2, so DRAM traffic is far below the algorithmic bytes")
        # the fp64 m x m tail, timed by the phase timers (graph off)
        if "chol_blocked" in phases and phases["chol_blocked"]["ms_per_step"] > 0:
            t_tail = phases["chol_blocked"]["ms_per_step"] * 1e-3 / ql
            fl_tail = 2.0 * (2.0 / 3.0) * m**3          # Cholesky (m^3/3 FMA) + inverse factor (m^3/3 FMA), 2 flops per FMA
            kern["tail"] = dict(bound="tensor", pipe="fp64 DMMA (latency-bound pivot chain)", kernels="agp_tail3.cuh / agp_tail2.cuh: fp64 Cholesky + inverse factor of P_v",
                                seconds_per_launch=t_tail, algorithmic_flops_per_launch=fl_tail, achieved=fl_tail / t_tail / 1e12, unit="TFLOP/s",
                                peak=36.0, peak_source="measured DFMA/DMMA rate on this part: 62-64 FMA/clk/SM x 148 SMs x 1.965 GHz (profiles/r1/microbench_out); no fp64 entry in MEASURED_PEAKS.json",
                                frac=fl_tail / t_tail / 1e12 / 36.0, share_of_step=min(1.0, t_tail * ql / step_s), traffic=None,
                                note="per latent; m sequential pivots x ~114 cycles (measured) = %.0f us is the floor of any Cholesky-based update at this m" % (m * 114 / 1.965e3))
        if kern:
            top = max((k_ for k_ in kern if k_ != "knm"), key=lambda k_: kern[k_]["share_of_step"])
            roof = dict(bound=kern[top]["bound"], kernel=top, achieved=kern[top]["achieved"], peak=kern[top]["peak"], unit=kern[top]["unit"], frac=kern[top]["frac"],
                        traffic=kern[top].get("traffic"), share_of_step=kern[top]["share_of_step"],
                        traffic_source=(traffic.get("_source", None) and traffic["_source"] + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"),
                        peak_source=kern[top].get("peak_source", psrc), dominant_by="largest share of the device-timed step",
                        timing="GEMM / K_nm kernels: 50 back-to-back launches between one CUDA-event pair on the launching stream; tail: per-phase CUDA events, graph off",
                        kernels=kern, whole_step=whole)
    if roof is None:
        roof = dict(bound="tensor", kernel="whole_step", achieved=whole["achieved"], peak=whole["peak"], unit="TFLOP/s", frac=whole["frac"], traffic=None,
                    peak_source=psrc, whole_step=whole, note="multi-rank / multi-latent run: per-kernel figures are in the N=1 C2 line")

    if rank == 0:
        cpu = None
        parity = None
        if not args.no_cpu_baseline and cfg_name == "C2" and world == 1:
            cpu, parity = cpu_replay(cfg_name, X, y, Z, A, mbs, W, K, elbo, post_timed)
        # the driver's N = 1 point is C2 (the metric's single-latent configuration, which does not shard); a multi-latent line carries the
        # one-GPU figure of ITS workload (measured with `--gpus 1 --config C5|C4`, committed) so that strong-scaling efficiency can be read
        # as value / (N x same_workload_one_gpu.value)
        one_gpu = None
        if Q > 1:
            try:
                one_gpu = json.load(open(os.path.join(ROOT, "profiles", "r2", "one_gpu_points.json"))).get(cfg_name)
            except Exception:
                one_gpu = None
        line = dict(metric=METRIC, value=value, unit="iters/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                    higher_is_better=True, scaling="weak" if Q == 1 else "strong", vs_baseline=None,
                    dtype={"f32": "f32 (fp32 SIMT contractions, f64 m x m tail)", "tf32x3": "tf32x3 (tcgen05, f64 m x m tail)", "f64": "f64"}[args.precision],
                    data="synthetic", config=config_dict(cfg_name),
                    run=dict(l2=f"inputs larger than L2: every step gathers a new random minibatch from the resident {int(n * (4 * D + 12) / 1e6)} MB (X, |x|^2, y) arrays; no flush",
                             graph=bool(args.graph and (world == 1 or peer)), precision=args.precision, latents_per_rank=ql,
                             exchange=("none (single rank)" if world == 1 else "per-sample moments over NVLink peer memory inside the step" if peer else "NCCL all-gather of the per-sample moments"),
                             env={k: os.environ[k] for k in sorted(os.environ) if k.startswith("AGP_")}),
                    gpu_launches=int(launches), elbo_last=elbo, elbo_parity=parity, roofline=roof, cpu_baseline=cpu, e2e=e2e, clocks=clocks, phases=phases)
        if one_gpu is not None:
            line["same_workload_one_gpu"] = one_gpu
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            sys.stderr.write("ELBO / posterior parity against the CPU oracle FAILED: %r\n" % (parity,))
            sys.exit(3)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_predict(args, local_rank):
    """Extra evidence leg (`--predict`, not the driver's bench line): _predict_f (training/predictions.jl:25-50) of the C2 model on
    `--predict-rows` test points.  The engine predicts in chunks of its batch capacity: with the default chunk = all rows the
    test-point kernel matrix K_* (rows x m fp32, 2.1 GB at 2^20 x 512) is written to and read back from HBM -- the one
    configuration in which the K_nm construction kernel is HBM-bound -- while `--predict-chunk 8192` keeps every
    chunk's K_* in the 126 MB L2 like a training step.  Prints one JSON line: end-to-end rows/s through agp_predict_f (host f32 rows
    in, host f64 mean / variance out) and the per-kernel rooflines at that chunk size (agp_time_kernel)."""
    import ctypes as C

    import torch

    import agp_b200 as agp

    torch.cuda.set_device(local_rank)
    L = agp._lib
    c = CONFIGS["C2"]
    D, m, B0 = c["D"], c["m"], c["B"]
    n = int(args.predict_rows)
    chunk = int(args.predict_chunk or n)
    if n % 128 or chunk % 128:
        raise SystemExit("--predict-rows / --predict-chunk must be multiples of 128")
    rng = np.random.default_rng(0)
    X = rng.standard_normal((max(n, 1 << 20), D), dtype=np.float32)            # training inputs (also the pool the kernel timings gather from)
    w = rng.standard_normal(D).astype(np.float32)
    y = np.where(X @ w + 0.1 * rng.standard_normal(X.shape[0], dtype=np.float32) >= 0, 1.0, -1.0)
    Z = X[rng.permutation(X.shape[0])[:m]].astype(np.float64)
    mbs = np.stack([rng.choice(X.shape[0], B0, replace=False) for _ in range(3)]).astype(np.int64)
    sc = 1.0 / np.sqrt(D)
    model = agp.SVGP(agp.SqExponentialKernel() @ agp.ScaleTransform(sc), agp.LogisticLikelihood(), agp.AnalyticSVI(B0), Z, precision="tf32x3", device=local_rank)
    model._engine(chunk)                      # batch capacity = prediction chunk
    agp.train(model, X, y, 3, minibatches=list(mbs))
    eng = model._eng
    lib = eng.lib
    Xt = rng.standard_normal((n, D), dtype=np.float32)
    # parity of the prediction against the oracle on the first rows (outside every timed region)
    par = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import agp_oracle as O

        mo = O.SVGP(O.Kernel("sqexp", scale=sc), O.LogisticLikelihood(), O.AnalyticSVI(B0), Z)
        mo, so = O.train(mo, X.astype(np.float64), y, 3, minibatches=list(mbs))
        mu_o, var_o = O.predict_f(mo, Xt[:2048].astype(np.float64), cov=True)
        mu_e, var_e = agp.predict_f(model, Xt[:2048], cov=True)
        par = dict(rows=2048, mu_rel=float(np.linalg.norm(mu_e - mu_o[0]) / np.linalg.norm(mu_o[0])),
                   var_rel=float(np.linalg.norm(var_e - var_o[0]) / np.linalg.norm(var_o[0])))
        par["ok"] = bool(par["mu_rel"] < 5e-4 and par["var_rel"] < 5e-4)
    # end to end: agp_predict_f on host rows
    mu = np.empty(n); var = np.empty(n)
    Xk, xp, dt, layout, nt, _ = agp.api._x_args(Xt)
    reps = max(3, min(args.steps, 10))
    for _ in range(2):
        eng.ck(lib.agp_predict_f(eng.model, xp, dt, layout, nt, 1, L.dptr(mu), L.dptr(var)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.ck(lib.agp_predict_f(eng.model, xp, dt, layout, nt, 1, L.dptr(mu), L.dptr(var)))
    torch.cuda.synchronize()
    dt_s = (time.perf_counter() - t0) / reps
    # kernel level at this chunk size: one resident-list step with B = chunk, then agp_time_kernel
    kern = {}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    peak_tf32 = (peaks.get("bf16_tflops", 1590.0)) / 2.0
    Bk = chunk
    lists = np.stack([rng.permutation(X.shape[0])[:Bk] for _ in range(3)]).astype(np.int64)
    eng.ck(lib.agp_minibatches_upload(eng.model, lists.ctypes.data_as(L.c_int64_p), 3, Bk, 0))
    eng.ck(lib.agp_step(eng.model, None, Bk, 0, X.shape[0] / Bk))
    tk = {}
    for which, name in ((0, "knm"), (1, "gemm_v"), (2, "gemm_v_sigma")):
        v = C.c_double(0.0)
        eng.ck(lib.agp_time_kernel(eng.model, which, 5, C.byref(v)))
        tk[name] = v.value * 1e-3
    byts = 4.0 * (Bk * D + m * D + Bk * m) + 8.0 * Bk
    traffic = {}
    tf = os.path.join(ROOT, "profiles", "r2", "ncu_traffic_predict.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf))
    kern["knm"] = dict(bound="hbm", kernel="knm_umma_kernel", rows=Bk, seconds_per_launch=tk["knm"], algorithmic_bytes_per_launch=byts,
                       achieved=byts / tk["knm"] / 1e9, peak=hbm, unit="GB/s", frac=byts / tk["knm"] / 1e9 / hbm, traffic=traffic.get("knm"),
                       peak_source="MEASURED_PEAKS.json hbm_gbs, of measured" if "hbm_gbs" in peaks else "fallback 6.65 TB/s, of fallback")
    for k_ in ("gemm_v", "gemm_v_sigma"):
        fl = 1.0 * Bk * m * m
        kern[k_] = dict(bound="tensor", rows=Bk, seconds_per_launch=tk[k_], algorithmic_flops_per_launch=fl, achieved=fl / tk[k_] / 1e12, peak=peak_tf32,
                        unit="TFLOP/s", frac=fl / tk[k_] / 1e12 / peak_tf32, tensor_pipe_frac_3xtf32=3 * fl / tk[k_] / 1e12 / peak_tf32, traffic=traffic.get(k_))
    line = dict(leg="predict", metric="predict_f rows/sec, SVGP Logistic SqExp m=512 D=32 (mean + variance)", value=n / dt_s, unit="rows/s", n_gpus=1,
                ms_per_call=1e3 * dt_s, higher_is_better=True, dtype="tf32x3 (tcgen05), f64 statistics", data="synthetic",
                config=dict(workload="C2 model, _predict_f on test rows", rows=n, chunk=chunk, D=D, m=m),
                e2e=dict(value=n / dt_s, unit="rows/s", h2d_bytes_per_call=n * D * 4, d2h_bytes_per_call=n * 16,
                         call="agp_predict_f(host x[n,D] f32) -> host mean[n], var[n] f64 (pageable host arrays)"),
                roofline=dict(bound="hbm", kernel="knm", achieved=kern["knm"]["achieved"], peak=hbm, unit="GB/s", frac=kern["knm"]["frac"],
                              traffic=kern["knm"]["traffic"], kernels=kern,
                              note="chunk = rows: K_* is materialised in HBM; chunk = 8192: K_* stays in L2 and the DRAM traffic is the x rows in, 16 B/row out"),
                predict_parity=par)
    print(json.dumps(line), flush=True)
    if par is not None and not par["ok"]:
        sys.exit(3)


def cpu_replay(cfg_name, X, y, Z, A, mbs, W, K, elbo_engine, post_engine, max_iters=130):
    """The fp64 oracle replays the trajectory the engine just ran (1 initial + W warm-up + K timed iterations on the same lists) when
    that is at most `max_iters` iterations, otherwise a fresh pair is not available and only the first iterations are timed.
    Returns (cpu_baseline, elbo_parity): the CPU baseline is the oracle's throughput over this replay, the parity object compares
    the ELBO / mu / Sigma of the TIMED engine run with the oracle's at the same iteration (outside every timed region)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import agp_oracle as O

    total = 1 + W + K
    X64 = X.astype(np.float64)
    model = oracle_model(O, cfg_name, Z, A)
    lists = [mbs[0]] + [mbs[i] for i in range(W + K)]
    n_run = min(total, max_iters)
    model, state = O.train(model, X64, y, 2, minibatches=lists[:2])
    t0 = time.perf_counter()
    model, state = O.train(model, X64, y, n_run - 2, minibatches=lists[2:n_run], state=state)
    dt = time.perf_counter() - t0
    cpu = dict(value=(n_run - 2) / dt, unit="iters/s", cores=blas_threads(), kind="port",
               sample=f"{n_run - 2} iterations of this workload (the first {n_run} minibatches of the GPU run replayed) with the fp64 NumPy/OpenBLAS oracle on all host threads")
    parity = None
    if n_run == total:
        eo = model.ELBO(state, state["y_batch"])
        mu, S = post_engine[0], post_engine[1]
        gp = model.f[0]
        rel = abs(elbo_engine - eo) / max(1.0, abs(eo))
        r_mu = float(np.linalg.norm(mu - gp.mu) / np.linalg.norm(gp.mu))
        r_S = float(np.linalg.norm(S - gp.Sigma) / np.linalg.norm(gp.Sigma))
        parity = dict(ours=elbo_engine, oracle=eo, rel=rel, mu_rel_fro=r_mu, Sigma_rel_fro=r_S, iterations=total, tol=5e-4,
                      ok=bool(rel <= 5e-4 and r_mu <= 5e-4 and r_S <= 5e-4),
                      what="ELBO, mu, Sigma of the TIMED engine run after its last step vs the fp64 oracle replaying the same minibatches")
    return cpu, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("AGP_BENCH_PRECISION", "tf32x3"), choices=["f32", "tf32x3", "f64"])
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--e2e-mode", default=os.environ.get("AGP_BENCH_E2E", "lagged"), choices=["sync", "lagged"],
                    help="lagged: agp_step_batch_async + agp_result_wait (every result read, one step late); sync: agp_step_batch + agp_get_posterior")
    ap.add_argument("--e2e-dtype", default="f64", choices=["f64", "f32"], help="dtype of the host x rows of the end-to-end leg (Julia arrays are f64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="auto", choices=["auto", "C2", "C3", "C4", "C5"],
                    help="auto = C2 on one GPU (the metric's configuration), C5 (64 latents sharded over the ranks) on N > 1")
    ap.add_argument("--timed-only", action="store_true", help="profiling aid: run only warm-up + the timed loop")
    ap.add_argument("--predict", action="store_true", help="extra evidence leg: _predict_f of the C2 model on --predict-rows test points")
    ap.add_argument("--predict-rows", type=int, default=1 << 20)
    ap.add_argument("--predict-chunk", type=int, default=0, help="prediction chunk = batch capacity of the engine (0 = all rows: K_* goes through HBM)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name = args.config if args.config != "auto" else ("C2" if args.gpus == 1 else "C5")
    if args.impl == "reference":
        run_reference(args, rank, world, cfg_name)
        return
    if args.predict:
        if rank == 0:
            run_predict(args, local_rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank, cfg_name)


if __name__ == "__main__":
    main()
