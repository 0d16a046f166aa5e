#!/usr/bin/env python
"""bench.py -- natural-gradient CAVI iterations/second of the SVGP AnalyticSVI hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f32|tf32x3|f64]

Workload (BASELINE.json configs[1], "C2"): SVGP, LogisticLikelihood, SqExponentialKernel with lengthscale
sqrt(D), n = 1e6, D = 32, m = 512 inducing points, minibatch 8192, RobbinsMonro(0.51, 1), K_mm fixed
(optimiser=false semantics).  Synthetic data, seeded.  One "step" = one update_parameters! call
(training/training.jl:140-144 of the reference).

N > 1 (torchrun, one rank per GPU): multi-output SVGP with Q = T = N Logistic tasks of the same shape, one
latent GP per rank, the per-sample moments all-gathered over NCCL every step (weak scaling: per-GPU work is
fixed); value = latent-GP iterations per second summed over ranks (at N = 1 this is plain iterations/s).

--impl reference: the fp64 NumPy/OpenBLAS restatement of the reference path (oracle/) timed on the host
cores (the Julia reference itself cannot run here: no Julia in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n=1_000_000, D=32, m=512, B=8192)
METRIC = "natural-gradient CAVI iters/sec, SVGP m=512 bs=8192"
# --config C3 (BASELINE.json configs[2], an extra evidence run, NOT the contract line): SVGP StudentT(nu=3) Matern-3/2,
# n=1e7 D=64 m=1024 minibatch=16384.  n is cut to 2e6 rows (512 MB resident, still >> L2; the step cost is n-independent)
# to keep host-side data generation short.
CFG_C3 = dict(n=2_000_000, D=64, m=1024, B=16384)


def make_problem(n, D, m, B, n_lists, seed=0, n_task=1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, D), dtype=np.float32)
    W = rng.standard_normal((D, n_task)).astype(np.float32)
    ys = [np.sign(X @ W[:, t] + 0.1 * rng.standard_normal(n, dtype=np.float32)).astype(np.float64) for t in range(n_task)]
    for y in ys:
        y[y == 0] = 1.0
    Z = X[rng.permutation(n)[:m]].astype(np.float64)
    mbs = np.stack([rng.choice(n, B, replace=False) for _ in range(n_lists)]).astype(np.int64)
    return X, ys, Z, mbs, rng


def flops_per_iter(B, m, D):
    """algorithmic FLOPs of one step for one latent (SURVEY 8d / BASELINE.md section 3)."""
    return 2.0 * B * m * D + 6.0 * B * m * m + 8.0 * B * m + (5.0 / 3.0) * m**3


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
def run_reference(args, rank, world):
    """CPU restatement of the reference path (oracle) on the host cores, rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import agp_oracle as O

    n, D, m, B = CFG["n"], CFG["D"], CFG["m"], CFG["B"]
    K, W = args.steps, args.warmup
    X, ys, Z, mbs, _ = make_problem(n, D, m, B, K + W)
    X64 = X.astype(np.float64)
    model = O.SVGP(O.Kernel("sqexp", scale=1.0 / np.sqrt(D)), O.LogisticLikelihood(), O.AnalyticSVI(B), Z)
    state = None
    if W > 0:
        model, state = O.train(model, X64, ys[0], W, minibatches=list(mbs[:W]))
    t0 = time.perf_counter()
    model, state = O.train(model, X64, ys[0], K, minibatches=list(mbs[W:]), state=state)
    dt = time.perf_counter() - t0
    v = K / dt
    cores = os.cpu_count()
    line = dict(metric=METRIC, value=v, unit="iters/s", n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=1e3 * dt / K,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="C2: SVGP Logistic SqExp n=1e6 D=32 m=512 minibatch=8192", **CFG),
                cpu_baseline=dict(value=v, unit="iters/s", cores=cores, kind="port",
                                  sample=f"{K} full iterations of the same workload (NumPy/SciPy fp64 on OpenBLAS, all host threads)"),
                e2e=dict(value=v, unit="iters/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(n_iter=16, warm=2):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import agp_oracle as O

    n, D, m, B = 200_000, CFG["D"], CFG["m"], CFG["B"]  # the step cost does not depend on n (only the gather does)
    X, ys, Z, mbs, _ = make_problem(n, D, m, B, n_iter + warm, seed=1)
    X64 = X.astype(np.float64)
    model = O.SVGP(O.Kernel("sqexp", scale=1.0 / np.sqrt(D)), O.LogisticLikelihood(), O.AnalyticSVI(B), Z)
    model, state = O.train(model, X64, ys[0], warm, minibatches=list(mbs[:warm]))
    t0 = time.perf_counter()
    O.train(model, X64, ys[0], n_iter, minibatches=list(mbs[warm:]), state=state)
    dt = time.perf_counter() - t0
    return dict(value=n_iter / dt, unit="iters/s", cores=os.cpu_count(), kind="port",
                sample=f"{n_iter} iterations of the C2 step (D=32 m=512 B=8192; n=2e5 rows, the step cost is n-independent) "
                       "with the fp64 NumPy/OpenBLAS oracle on all host threads")


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch

    import agp_b200 as agp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    c3 = args.config == "C3"
    cfg = CFG_C3 if c3 else CFG
    n, D, m, B = cfg["n"], cfg["D"], cfg["m"], cfg["B"]
    K, W = args.steps, max(args.warmup, 3)
    n_lists = K + W
    X, ys, Z, mbs, rng = make_problem(n, D, m, B, n_lists, n_task=world)
    if c3:   # regression targets with Student-t noise
        ys = [(np.sin(X[:, 0]) + 0.5 * X[:, 1] + 0.1 * rng.standard_t(3.0, n)).astype(np.float64)]
    kern = (agp.Matern32Kernel() if c3 else agp.SqExponentialKernel()) @ agp.ScaleTransform(1.0 / np.sqrt(D))
    # a dedicated non-default stream: the engine launches on it, torch events / NCCL are ordered on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    if world == 1:
        model = agp.SVGP(kern, agp.StudentTLikelihood(3.0, 1.0) if c3 else agp.LogisticLikelihood(), agp.AnalyticSVI(B), Z, precision=args.precision,
                         device=local_rank, stream=stream)
        y_arg = ys[0]
    else:
        A = rng.standard_normal((world, world))
        A /= np.linalg.norm(A, axis=1, keepdims=True)
        Zs = [X[np.random.default_rng(100 + q).permutation(n)[:m]].astype(np.float64) for q in range(world)]
        model = agp.MOSVGP(kern, [agp.LogisticLikelihood() for _ in range(world)], agp.AnalyticSVI(B), Zs, A=A, precision=args.precision,
                           device=local_rank, stream=stream, shard=(rank, world))
        y_arg = ys
    # one API-level step initialises everything (upload, compute_K) and checks the error path
    agp.train(model, X, y_arg, 1, minibatches=[mbs[0]])
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
    value = world * K / (ms * 1e-3)
    elbo = agp.ELBO(model)

    # ---------------- per-kernel phase timers (second pass; roofline) ----------------
    roof = None
    phases = {}
    if args.timed_only:  # ncu launch-list runs: nothing but the timed loop (no per-kernel re-timing, no e2e, no CPU leg)
        if rank == 0:
            print(json.dumps(dict(metric=METRIC, value=value, unit="iters/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                                  gpu_launches=int(launches), note="--timed-only run (profiling aid, not a bench line)")), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    if world == 1:
        eng.ck(lib.agp_use_graph(eng.model, 0))
        eng.ck(lib.agp_profile_enable(eng.model, 1))
        for _ in range(K):
            step_async()
        import ctypes as C

        names = (C.c_char_p * 32)()
        msv = (C.c_double * 32)()
        lv = (C.c_int64 * 32)()
        nph = lib.agp_profile_read(eng.model, 32, names, msv, lv)
        eng.ck(lib.agp_profile_enable(eng.model, 0))
        for i in range(nph):
            phases[names[i].decode()] = dict(ms_per_step=msv[i] / K, launches_per_step=lv[i] / K)
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        # per-kernel durations: `reps` back-to-back launches between one CUDA-event pair on the engine's stream
        # (agp_time_kernel), so the ~4 us event overhead seen by the per-phase timers is amortised
        reps = 50
        tk = {}
        for which, name in ((0, "kmat_knm"), (1, "gemm_v"), (2, "gemm_v_sigma"), (3, "gemm_gram")):
            v = C.c_double(0.0)
            eng.ck(lib.agp_time_kernel(eng.model, which, reps, C.byref(v)))
            tk[name] = v.value * 1e-3
        bf16 = peaks.get("bf16_tflops_sustained")
        peak = (bf16 / 2.0) if bf16 else 1590.0 / 2.0
        psrc = ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 = half the bf16 rate), of measured" if bf16
                else "fallback 1.59 PFLOP/s bf16 / 2, of fallback")
        # algorithmic FLOPs: V = Knm L^-T and V X^T have a lower-triangular right operand (B m^2 each); the Gram product
        # U^T U is 2 B m^2 (SURVEY 8d counts it in full; only the upper tiles are executed).  3xTF32 executes 3x these.
        fl = {"gemm_v": 1.0 * B * m * m, "gemm_v_sigma": 1.0 * B * m * m, "gemm_gram": 2.0 * B * m * m}
        kern = {k: dict(seconds_per_launch=tk[k], algorithmic_flops_per_launch=fl[k], achieved=fl[k] / tk[k] / 1e12, peak=peak,
                        unit="TFLOP/s", frac=fl[k] / tk[k] / 1e12 / peak, tensor_pipe_frac_3xtf32=3 * fl[k] / tk[k] / 1e12 / peak *
                        ((m // 128 + 1) / (2.0 * (m // 128)) if k == "gemm_gram" else 1.0)) for k in fl}
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this same command
        # (profiles/r1/ncu_traffic.json, written by profiles/extract_ncu.py); null when the file is absent
        traffic = {}
        tf = os.path.join(ROOT, "profiles", "r1", "ncu_traffic.json")
        if os.path.exists(tf) and not c3:
            try:
                traffic = json.load(open(tf))
            except Exception:
                traffic = {}
        for k_ in kern:
            kern[k_]["traffic"] = traffic.get(k_)
        top = max(fl, key=lambda k: tk[k])
        roof = dict(bound="tensor", kernel=top, achieved=kern[top]["achieved"], peak=peak, unit="TFLOP/s", frac=kern[top]["frac"],
                    traffic=traffic.get(top), traffic_source="profiles/r1/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" if traffic else None,
                    peak_source=psrc, algorithmic_flops_per_launch=fl[top],
                    timing=f"{reps} back-to-back launches between one CUDA-event pair on the launching stream", kernels=kern)
        # the fp64 m x m tail (largest share of the step): latency-bound by the 512-pivot chain, reported for completeness
        if "chol_blocked" in phases and phases["chol_blocked"]["ms_per_step"] > 0:
            t_tail = phases["chol_blocked"]["ms_per_step"] * 1e-3
            fl_tail = 2.0 * (2.0 / 3.0) * m**3          # Cholesky (m^3/3 FMA) + inverse factor (m^3/3 FMA), 2 flops per FMA
            roof["tail"] = dict(bound="latency (fp64 pivot chain)", kernels="tail2_potf2_first_kernel + tail2_step_kernel x m/64",
                                seconds_per_step=t_tail, algorithmic_flops=fl_tail, achieved=fl_tail / t_tail / 1e12, unit="TFLOP/s",
                                peak=36.0, peak_source="measured DFMA/DMMA rate on this part: 62-64 FMA/clk/SM x 148 SMs x 1.965 GHz (profiles/r1/microbench_out)",
                                frac=fl_tail / t_tail / 1e12 / 36.0,
                                note="m sequential pivots x ~114 cycles (measured) = %.0f us is the floor of any Cholesky-based update at this m; "
                                     "timed by the phase timers (graph off)" % (m * 114 / 1.965e3))
        byts = 4.0 * (B * D + m * D + B * m) + 8.0 * B
        hbm = peaks.get("hbm_gbs", 6650.0)
        roof["knm"] = dict(bound="hbm", kernel="knm_umma_kernel", seconds_per_launch=tk["kmat_knm"], achieved=byts / tk["kmat_knm"] / 1e9, peak=hbm,
                           unit="GB/s", frac=byts / tk["kmat_knm"] / 1e9 / hbm, algorithmic_bytes_per_launch=byts,
                           peak_source="MEASURED_PEAKS.json hbm_gbs, of measured" if "hbm_gbs" in peaks else "fallback 6.65 TB/s, of fallback",
                           traffic=traffic.get("kmat_knm"),
                           note="at C2 the 16.8 MB K_nm output stays in the 126 MB L2 (write-back), so DRAM traffic is far below the algorithmic bytes")

    # ---------------- end-to-end through the host-buffer API ----------------
    e2e = None
    if world == 1:
        Ke = min(K, 50)
        xb = [torch.empty((B, D), dtype=torch.float64).pin_memory() for _ in range(Ke)]
        yb = [torch.empty((B,), dtype=torch.float64).pin_memory() for _ in range(Ke)]
        for i in range(Ke):
            idx = mbs[(W + i) % n_lists]
            xb[i].numpy()[:] = X[idx]
            yb[i].numpy()[:] = ys[0][idx]
        import ctypes as C

        mu = np.empty(m)
        eng.ck(lib.agp_use_graph(eng.model, 1 if args.graph else 0))   # public switch: the compute part of a host-batch step replays one CUDA graph
        pending = []   # --e2e-async: ticket of the step whose result has not been read yet

        def e2e_step(i):
            arr = (C.c_void_p * 1)(yb[i].data_ptr())
            if args.e2e_async:
                # opt-in (untested on a GPU at the time of writing): no per-step synchronisation; the result of step i-1 is
                # read while step i runs, the H2D copy of step i overlaps the computation of step i-1
                tk = C.c_int64(0)
                eng.ck(lib.agp_step_batch_async(eng.model, C.c_void_p(xb[i].data_ptr()), L.DTYPE_F64, L.LAYOUT_ROWMAJOR, arr, L.Y_REAL, B, rho,
                                                C.byref(tk)))
                if pending:
                    eng.ck(lib.agp_result_wait(eng.model, pending.pop(), L.dptr(mu)))
                pending.append(tk.value)
                return
            eng.ck(lib.agp_step_batch(eng.model, C.c_void_p(xb[i].data_ptr()), L.DTYPE_F64, L.LAYOUT_ROWMAJOR, arr, L.Y_REAL, B, rho))
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
        e2e = dict(value=Ke / dt, unit="iters/s", h2d_bytes_per_step=B * D * 8 + B * 8, d2h_bytes_per_step=m * 8 + 4,
                   steps=Ke, call=("agp_step_batch_async(host x[B,D] f64, host y[B]) + agp_result_wait(previous step's mu)" if args.e2e_async
                                   else "agp_step_batch(host x[B,D] f64, host y[B]) + agp_get_posterior(mu)"))
    else:
        e2e = dict(value=None, unit="iters/s", h2d_bytes_per_step=B * 8, d2h_bytes_per_step=0,
                   note="latent-sharded run: measured at N=1 only")

    if rank == 0:
        cpu = cpu_baseline_sample() if (world == 1 and not args.no_cpu_baseline and not c3) else None
        line = dict(metric=METRIC, value=value, unit="iters/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                    higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"f32": "f32 (fp32 SIMT contractions, f64 m x m tail)", "tf32x3": "tf32x3 (tcgen05, f64 m x m tail)", "f64": "f64"}[args.precision],
                    data="synthetic",
                    config=dict(workload=("C3 (extra evidence run): SVGP StudentT(3) Matern-3/2 n=2e6 (of 1e7) D=64 m=1024 minibatch=16384" if c3 else
                                          "C2: SVGP Logistic SqExp n=1e6 D=32 m=512 minibatch=8192" if world == 1 else
                                          f"C2-shaped multi-output SVGP: {world} Logistic tasks x {world} latent GPs, one latent per GPU, "
                                          + ("per-sample moments exchanged over NVLink peer memory inside the step" if peer else "moments NCCL all-gather per step")),
                                l2=f"inputs larger than L2: every step gathers a new random minibatch from the resident {int(n * (4 * D + 12) / 1e6)} MB (X, |x|^2, y) arrays; no flush",
                                graph=bool(args.graph and (world == 1 or peer)), precision=args.precision,
                                **({"experimental_env": {k: os.environ[k] for k in ("AGP_UMMA_V2", "AGP_TAIL_NS", "AGP_TAIL_NS_AFTER", "AGP_TAIL_NS_TOL") if k in os.environ}}
                                   if any(k in os.environ for k in ("AGP_UMMA_V2", "AGP_TAIL_NS")) else {}), **cfg),
                    gpu_launches=int(launches), elbo_last=elbo, roofline=roof, cpu_baseline=cpu, e2e=e2e, clocks=clocks, phases=phases)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("AGP_BENCH_PRECISION", "tf32x3"), choices=["f32", "tf32x3", "f64"])
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--e2e-async", action="store_true", help="end-to-end leg through agp_step_batch_async / agp_result_wait (opt-in)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="C2", choices=["C2", "C3"], help="C2 = the contract workload (default); C3 = extra evidence run at the larger configuration")
    ap.add_argument("--timed-only", action="store_true", help="profiling aid: run only warm-up + the timed loop")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
