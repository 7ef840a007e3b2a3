#!/usr/bin/env python
"""Benchmark of the mock-library hot path (BASELINE.json metric: galaxies/s synthesised, SED -> photometry).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement on the host cores

One "step" = one pass of the hot path (group by metallicity bracket and redshift, weight and IGM kernels,
tcgen05 contraction + fused epilogue, finalize) over one batch of synthetic galaxies.  At N=1 the workload is BASELINE configs[1]
(1M galaxies, LogNormal SFH + Calzetti dust screen, z 0-10 with IGM, 20 NIRCam+MIRI filters).
For N>1 every rank processes its own batch of the same size (weak scaling, no data-path collective);
time is the max over ranks of the CUDA-event time of the K steps.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "galaxies_per_second_synthesised_sed_to_photometry"
UNIT = "galaxies/s"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            d = json.load(fh)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_inputs(w):
    """Plain arrays the CPU restatement needs (bench.py is allowed to drive oracle/)."""
    from oracle import oracle as O
    from synference_b200 import igm as I
    lam = np.asarray(w.grid.lam)
    ga, gu = O.emission_parts(w.grid.spectra, lam, w.emission_key, float(w.emission_model.fesc),
                              float(w.emission_model.fesc_ly_alpha))
    kap = O.dust_kappa(lam) if w.emission_model.dust_curve is not None else None
    filt = [(f.lam, f.t) for f in w.filters]
    return dict(log10ages=w.grid.log10ages, metallicities=w.grid.metallicity, lam=lam, g_att=ga, g_un=gu,
                filters=filt, kappa=kap, igm=(I.INOUE14_LAF, I.INOUE14_DLA))


def time_cpu(w, n_sample, threads):
    from oracle import c_oracle as CO
    CO.build()
    inp = oracle_inputs(w)
    p = w.params.slice(slice(0, n_sample))
    t0 = time.perf_counter()
    CO.synthesize(p, inp["log10ages"], inp["metallicities"], inp["lam"], inp["g_att"], inp["g_un"], inp["filters"],
                  kappa=inp["kappa"], igm=inp["igm"], nthreads=threads)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """CPU arm: the C/OpenMP restatement of the reference path on all host threads."""
    if rank != 0:
        return
    from synference_b200.configs import make_workload
    threads = os.cpu_count() or 1
    # a bounded sample per step: about two minutes of CPU work for the whole run at ~10 k galaxies/s on 16 threads
    n_step = args.ref_sample if args.ref_sample > 0 else int(np.clip(1_200_000 // max(1, args.steps + 1), 10_000, 150_000))
    w = make_workload(args.workload, n_step)
    from synference_b200.engine import build_tables
    tables = build_tables(w.grid, w.emission_model, w.emission_key, w.filters)
    for _ in range(max(1, min(args.warmup, 1))):
        time_cpu(w, min(n_step, 2000), threads)
    t = [time_cpu(w, n_step, threads) for _ in range(args.steps)]
    total = float(np.sum(t))
    value = n_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the workload the metric is quoted on (what the CUDA arm runs); each CPU step is a bounded sample of it
        "config": bench_config(args.workload, args.galaxies, tables, w.params),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_step} galaxies of {workload_name(args.workload)} per step; C/OpenMP float64 restatement "
                                   "of the reference path (oracle/oracle_c.c) -- the reference's own Synthesizer extensions are "
                                   "not installable offline"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(workload, n, tables=None, params=None):
    """The `config` object of the JSON line -- the same keys in the CUDA arm and in the reference arm."""
    cfg = {"workload": workload_name(workload), "galaxies_per_gpu_per_step": int(n)}
    if tables is not None:
        t = tables
        k_alg = t["n_age"] * t["n_z"]
        delta = params is not None and params.zd_type in (0, 1) and t["n_z"] >= 2
        cfg.update({"n_lam": t["n_lam"], "n_filt": t["n_filt"], "k": k_alg, "k_exec": 2 * t["n_age"] if delta else k_alg,
                    "n_comp": t["n_comp"],
                    "l2": "per-step working set (SFH bin masses / weights %.2f GB + parameters + IGM rows %.2f GB) >> 126 MB L2; "
                          "no explicit flush" % ((8.0 * t["n_age"] if delta else 8.0 * t["k_pad"]) * n / 1e9,
                                                 4.0 * (t["igm"]["n_blue"] if t["igm"] else 0) * n / 1e9)})
    return cfg


def workload_name(key):
    return {"cfg1": "cfg1: README quickstart (LogNormal, delta Z, intrinsic, 8 NIRCam wide)",
            "cfg2": "cfg2: LogNormal SFH + Calzetti screen, z 0-10 + IGM, 20 NIRCam+MIRI filters",
            "cfg3": "cfg3: continuity SFH + Normal Z distribution, z 0-15, 20 filters"}[key]


def measure_tf32_peak(dev, seconds=1.0):
    """Dense TF32 throughput of this GPU the way MEASURED_PEAKS.json was made for bf16: torch.matmul 8192^3 with TF32
    tensor cores (cuBLAS), best of 10 (burst) and back-to-back for ~1 s (sustained)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        nn = 8192
        a = torch.randn((nn, nn), device=dev, dtype=torch.float32)
        b = torch.randn((nn, nn), device=dev, dtype=torch.float32)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * nn ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        reps = max(4, int(seconds * best * 1e12 / (2.0 * nn ** 3)))
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); torch.cuda.synchronize()
        sustained = reps * 2.0 * nn ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        return {"tf32_tflops": best, "tf32_tflops_sustained": sustained,
                "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32 tensor cores): best of 10, and %d back to back" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def raw_draw_params(w):
    """The workload's parameters as the float32 draws of draw_from_hypercube (library.py:1098) with max_age derived on the
    device (library.py:1206, :1287-1289) -- what a library build sends over PCIe (sb2_params.host_f32).  LogNormal workloads."""
    from synference_b200.cosmology import Planck18
    from synference_b200.engine import GalaxyParams
    s, p = w.samples, w.params
    rows = np.zeros((len(p), 4), dtype=np.float32)
    rows[:, 2], rows[:, 3] = s["tau"], s["peak_age_norm"]
    f32 = lambda a: None if a is None else np.asarray(a, dtype=np.float32)  # noqa: E731
    return GalaxyParams(f32(s["redshift"]), p.sfh_type, rows, p.zd_type, f32(s["log_zmet"]), None, f32(s["log_stellar_mass"]),
                        f32(s.get("tau_v")), max_age_from_z=True, norm_mask=0b10, age_zmax_gyr=float(Planck18.age(20.0).value))


def api_flow(args, rank, world, local, n):
    """The README flow through the reference-facing API (README.md:95-134): draw_from_hypercube -> generate_sfh_basis ->
    GalaxyBasis -> create_mock_library (pipeline files + compiled library on disk, library in memory).  With N > 1 ranks:
    create_mock_library(multi_node=True) -- every rank synthesises its contiguous slice of world*n galaxies and writes its own
    shard (library.py:3127-3138) -- then ONE gather of the in-memory photometry over NCCL.  Returns seconds per stage."""
    import shutil
    import tempfile
    import torch
    import synference_b200 as S
    from synference_b200 import distributed as D
    from synference_b200.configs import make_workload
    total = n * world
    t = {}
    w0 = make_workload(args.workload, 64)           # model objects (grid, instrument, emission model)
    t0 = time.perf_counter()
    d = S.draw_from_hypercube({"log_stellar_mass": (8.0, 12.0), "redshift": (0.01, 10.0), "log_zmet": (-4.0, -1.4),
                               "peak_age_norm": (0.0, 0.99), "tau": (0.2, 2.0), "tau_v": (0.0, 3.0)}, N=total, rng=42)
    t["draw_from_hypercube"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    sfhs, _ = S.generate_sfh_basis(S.SFH.LogNormal, ["tau", "peak_age_norm"], np.vstack((d["tau"], d["peak_age_norm"])).T,
                                   redshifts=np.array(d["redshift"]), max_redshift=20)
    zds = S.generate_metallicity_distribution(S.ZDist.DeltaConstant, log10metallicity=np.asarray(d["log_zmet"], dtype=float))
    t["generate_sfh_basis+metallicity"] = time.perf_counter() - t0
    out_dir = tempfile.mkdtemp(prefix="sb2_api_") if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        box = [out_dir]
        dist.broadcast_object_list(box, src=0)
        out_dir = box[0]
    basis = S.GalaxyBasis("bench_basis", d["redshift"], w0.grid, w0.emission_model, sfhs, zds, galaxy_params={"tau_v": d["tau_v"]},
                          instrument=w0.instrument, redshift_dependent_sfh=True, build_library=False)
    logm = np.asarray(d["log_stellar_mass"], dtype=float)
    basis._engine(w0.emission_key, max_batch=250_000)        # model creation (CUDA context, tables) is a one-off, not per galaxy
    if world > 1:
        D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cb = basis.create_mock_library("bench_lib", log_stellar_masses=logm, emission_model_key=w0.emission_key, out_dir=out_dir,
                                   overwrite=True, batch_size=250_000, multi_node=world > 1)
    t["create_mock_library"] = time.perf_counter() - t0
    files = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)) if rank == 0 else 0
    if rank == 0:
        # what this box's file system takes for the same number of bytes: ONE plain write of a resident buffer, no format,
        # no kernels (the limiter of the call above when the two are close)
        blob = np.zeros(files // 8, dtype=np.float64)
        t0 = time.perf_counter()
        with open(os.path.join(out_dir, "plain_write.bin"), "wb", buffering=0) as fh:
            fh.write(blob.view(np.uint8))
        t["plain_write_of_the_same_bytes"] = time.perf_counter() - t0
        del blob
    if world > 1:
        t0 = time.perf_counter()
        local_rows = torch.as_tensor(np.ascontiguousarray(cb.library_photometry.T)).to(torch.device("cuda", local))
        full = D.gather_rows(local_rows, total)
        torch.cuda.synchronize()
        t["gather_rows_nccl"] = time.perf_counter() - t0
        assert full.shape[0] == total
        D.barrier()
    if rank == 0:
        shutil.rmtree(out_dir, ignore_errors=True)
    return t, files


def run_cfg5(args, rank, world, local):
    """BASELINE configs[4]: cfg2 physics with photometry AND ~1000-pixel PRISM-like spectra out, full-wavelength write path.
    value: the device chain (contraction kernel with its full-wavelength output kept in HBM + resample kernel), device-resident;
    e2e: write_spectral_library -- the same chain with the pixels and fluxes copied to pinned host buffers (double-buffered, on
    a side stream) and written as uncompressed shards by a writer thread.  Weak scaling: every rank its own batch and shards."""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from synference_b200.configs import make_workload
    from synference_b200.engine import SynthEngine
    from synference_b200.spectral import SpectrumResampler, write_spectral_library
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = min(args.galaxies, 262144)                      # spectra of one batch stay in HBM: 262144 x 3712 x 4 B = 3.9 GB
    w = make_workload("cfg2", n, seed=42 + rank)
    eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n, device=local)
    ow = np.linspace(0.6, 5.3, 1000)
    rw = np.linspace(0.55, 5.4, 80)
    rr = 30.0 + 270.0 * ((rw - 0.55) / 4.85) ** 1.3
    plan = SpectrumResampler(np.asarray(w.grid.lam) * 1e-4, ow, rw, rr, device=local)
    dpar = eng.to_device(w.params)
    spec = torch.empty((n, eng.n_lam), dtype=torch.float32, device=dev)
    flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device=dev)
    pix = torch.empty((n, plan.n_px), dtype=torch.float32, device=dev)

    def step():
        eng.photometry_device(dpar, flux_base=flux, spectra=spec)
        plan.transform_into(spec, dpar.tensors["redshift"], pix)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = int(eng.lib.sb2_kernel_launches())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = int(eng.lib.sb2_kernel_launches()) - launches0
    resample_ms = plan.last_ms()
    out_dir = tempfile.mkdtemp(prefix=f"sb2_cfg5_r{rank}_")
    write_spectral_library(eng, plan, w.params, out_dir=out_dir, batch_size=65536)          # warm-up (pinned buffers, files)
    barrier()
    e2e_steps = max(2, min(args.steps, 10))
    t0 = time.perf_counter()
    bytes_out = 0
    for _ in range(e2e_steps):
        r = write_spectral_library(eng, plan, w.params, out_dir=out_dir, batch_size=65536)
        bytes_out += r["bytes"]
    e2e_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        write_spectral_library(eng, plan, w.params, out_dir=None, batch_size=65536)
    e2e_nofile_s = time.perf_counter() - t0
    shutil.rmtree(out_dir, ignore_errors=True)
    clocks = sampler.stop() if rank == 0 else None
    times = torch.tensor([ms_total, e2e_s * 1e3, e2e_nofile_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_nofile_ms = (float(x) for x in times)
    if rank == 0:
        peaks, peak_src = read_peaks()
        hbm = float(peaks["hbm_gbs"])
        alg = 4.0 * (eng.n_lam + plan.n_px)          # resample kernel: reads the spectrum once, writes the pixels once
        line = {
            "metric": METRIC, "value": world * n * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32x3 (fp32 accumulate; launches that write spectra keep three TF32 passes); weights/IGM in f64; spectra f32",
            "data": "synthetic",
            "config": {"workload": "cfg5: cfg2 physics, photometry + 1000-pixel PRISM-like spectra out (full-wavelength path)",
                       "galaxies_per_gpu_per_step": n, "n_lam": eng.n_lam, "n_px": plan.n_px, "n_filt": eng.n_filt,
                       "l2": "one batch's spectra (%.1f GB) >> 126 MB L2; no explicit flush" % (4.0 * eng.n_lam * n / 1e9)},
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(8 * 8 * n),
                    "d2h_bytes_per_step": int(4 * (plan.n_px + eng.n_filt) * n), "steps": e2e_steps,
                    "api": "synference_b200.spectral.write_spectral_library: device chain per 65536-galaxy batch, pixels + fluxes "
                           "to pinned double buffers on a side stream, uncompressed .npy shards by a writer thread",
                    "without_files_value": world * n * e2e_steps / (e2e_nofile_ms * 1e-3),
                    "bytes_written_per_step": int(bytes_out / e2e_steps)},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "kernel": "resample_kernel (variable-width Gaussian + flux-conserving rebin)",
                         "kernel_ms": resample_ms, "achieved": alg * n / (resample_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg * n / (resample_ms * 1e-3) / 1e9 / hbm, "traffic": None,
                         "peak_source": f"{peak_src} HBM copy bandwidth (MEASURED_PEAKS.json)",
                         "note": "algorithmic bytes per galaxy: 4*n_lam read + 4*n_px written; the contraction kernel of this "
                                 "chain additionally writes the 4*n_lam bytes the resample kernel reads"},
            "cpu_baseline": None, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--galaxies", type=int, default=1_000_000, help="galaxies per GPU per step")
    ap.add_argument("--ref-sample", type=int, default=0, help="galaxies per CPU step (0: sized so that K steps take ~2 min)")
    ap.add_argument("--cpu-sample", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the create_mock_library leg (api_e2e)")
    ap.add_argument("--traffic", type=float, default=None, help="ncu dram bytes per contraction launch (overrides profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "cfg5":
        run_cfg5(args, rank, world, local)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from synference_b200.configs import make_workload
    from synference_b200.engine import SynthEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.galaxies
    # every rank draws its own slice of the same Latin hypercube family (different seed per rank)
    w = make_workload(args.workload, n, seed=42 + rank)
    eng = SynthEngine(w.grid, w.emission_model, w.emission_key, w.filters, max_batch=n, device=local)
    dpar = eng.to_device(w.params)
    flux = torch.empty((n, eng.n_filt), dtype=torch.float32, device=dev)
    stage = np.zeros(3, dtype=np.float32)

    def step_device():
        eng.photometry_device(dpar, flux_base=flux)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stages = []
    launches0 = int(eng.lib.sb2_kernel_launches())
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = int(eng.lib.sb2_kernel_launches()) - launches0       # counted by the library, kernel by kernel
    # dominant kernel timed live over several launches, from the events the library records on the stream
    synth_ms = []
    for _ in range(min(args.steps, 5)):
        step_device()
        eng.lib.sb2_last_stage_ms(eng._h, stage.ctypes.data_as(C.POINTER(C.c_float)))
        synth_ms.append(float(stage[2]))
        stages.append([float(x) for x in stage])
    synth_ms_avg = float(np.mean(synth_ms))

    # ---- end-to-end through the host-buffer API (pinned host memory, H2D + kernels + D2H per step).
    # Headline: the streaming form a library build uses (GalaxyBasis.create_mock_library walks the population batch by
    # batch): step k is submitted while step k-1 is still in flight, so the PCIe copies of one batch overlap the kernels
    # of the next; every step's copy-in and copy-out lie inside the timed region.  Parameters travel as the float32 draws
    # draw_from_hypercube produces (sb2_params.host_f32, widened on the device; max_age derived on the device) whenever the
    # workload is built from such draws; the blocking single-call form is reported beside it.
    host_outs = [torch.empty((n, eng.n_filt), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
    host_out = host_outs[0]
    f32_transport = w.params.sfh_type == 5 and "peak_age_norm" in w.samples       # LogNormal workloads (cfg1, cfg2)
    pinned = raw_draw_params(w) if f32_transport else w.params
    transport = "f32" if f32_transport else "f64"
    dt = np.float32 if f32_transport else np.float64
    for k in ("redshift", "log_mass", "tau_v", "zd_value", "zd_sigma", "sfh_rows"):
        a = getattr(pinned, k)
        if a is not None:
            setattr(pinned, k, torch.as_tensor(np.ascontiguousarray(a, dtype=dt)).pin_memory().numpy())
    h2d = sum(getattr(pinned, k).nbytes for k in ("redshift", "log_mass", "tau_v", "zd_value", "zd_sigma", "sfh_rows")
              if getattr(pinned, k) is not None)
    d2h = host_out.nbytes
    for _ in range(2):
        eng.photometry(pinned, scaled=False, out=host_out, transport=transport)
    barrier()
    e2e_steps = max(2, min(args.steps, 20))     # the same K as the device-resident loop (bounded: host time)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.photometry(pinned, scaled=False, out=host_out, transport=transport)
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    tickets = []
    for k in range(e2e_steps):
        if len(tickets) == 2:
            eng.wait(tickets.pop(0))
        tickets.append(eng.submit(pinned, host_outs[k & 1], scaled=False, slot=k & 1, transport=transport))
    for tk in tickets:
        eng.wait(tk)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert np.isfinite(host_outs[0]).all() and np.array_equal(host_outs[0], host_outs[1])
    # what the host link alone sustains with every rank copying at once: the same bytes per step, both directions on their
    # own streams, no kernels -- if this equals the end-to-end number, the limiter is the host side (PCIe root / memory)
    barrier()
    h_in = torch.empty(h2d, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h, dtype=torch.uint8, device=dev)
    h_out = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    copy_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([ms_total, e2e_s * 1e3, e2e_sync_s * 1e3, copy_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_sync_ms, copy_ms = (float(x) for x in times)
    value = world * n * args.steps / (ms_total * 1e-3)
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)

    api = None
    if not args.no_api and args.workload == "cfg2":
        t_api, files = api_flow(args, rank, world, local, n)
        plain_s = t_api.pop("plain_write_of_the_same_bytes", None)
        tt = torch.tensor([t_api.get("create_mock_library", 0.0), sum(t_api.values())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        api = {"value": world * n / float(tt[0]), "unit": UNIT, "galaxies": world * n,
               "call": "GalaxyBasis.create_mock_library(%s): pipeline files + compiled library written (async, uncompressed), "
                       "library in memory" % ("multi_node=True, one shard per rank" if world > 1 else "batch_size=250000"),
               "value_whole_readme_flow": world * n / float(tt[1]), "seconds": t_api, "bytes_written_rank0_dir": int(files),
               "plain_write_of_the_same_bytes_s": plain_s,
               "note": "model creation (CUDA context, device tables) happens once before the timed call; "
                       "draw_from_hypercube and generate_sfh_basis are host numpy / scipy as in the reference"}

    if rank == 0:
        peaks, peak_src = read_peaks()
        t = eng.tables
        k_alg = t["n_age"] * t["n_z"]
        # DeltaConstant batches are grouped by metallicity bracket and multiply only the two grid
        # metallicities they can touch: K_exec = 2 n_age (report K_exec, count executed flops)
        delta = w.params.zd_type in (0, 1) and t["n_z"] >= 2
        k_exec = 2 * t["n_age"] if delta else k_alg
        k_mma = 2 * t["n_age_pad"] if delta else t["k_pad"]
        flops_alg = 2.0 * k_exec * t["n_lam"] * t["n_comp"] * n          # 2 K_exec N_lam C per galaxy
        # executed on the TF32 pipe: 3 passes, padded K, whole chunks -- but only the chunks some filter of the tile reaches
        # (tile_range).  The chunk count is estimated here from the redshifts (a tile's galaxies are redshift-neighbours).
        s3 = delta and 2 * t["n_age_pad"] <= 112 and t["n_age"] <= 64        # synth3_kernel (weights in tensor memory)
        cols = (96 if t["n_comp"] == 1 else 128) if s3 else (160 if t["n_comp"] == 1 else 256)
        lch = cols // t["n_comp"]
        mm = np.floor(np.log1p(np.asarray(w.params.redshift, dtype=np.float64)) / np.log(t["q"])).astype(np.int64)
        i_lo = np.maximum(0, int(t["filt_lo"].min()) - 1 - mm)
        i_hi = np.minimum(t["n_lam"] - 1, int(t["filt_hi"].max()) - mm)
        chunks = np.where(i_hi >= i_lo, i_hi // lch - i_lo // lch + 1, 0)
        n_chunk_all = -(-t["n_lam"] // lch)
        frac_lam = float(chunks.mean()) / n_chunk_all
        flops_exec = 3.0 * 2.0 * k_mma * (lch * t["n_comp"]) * float(chunks.sum())
        # synth3_kernel runs the two small products of the split as ONE bfloat16 MMA (twice the TF32 rate): in units of TF32
        # tensor time a chunk costs 2 passes, not 3 (SB2_TF32X3=1 restores three TF32 passes)
        env_x3 = os.environ.get("SB2_TF32X3", "")
        cross = s3 and not (env_x3 and env_x3[0] != "0")
        tf32_passes = 2.0 if cross else 3.0
        achieved = flops_alg / (synth_ms_avg * 1e-3) / 1e12
        executed = flops_exec / (synth_ms_avg * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of the contraction kernel: read from the committed ncu capture of
        # this kernel (profiles/, one `ncu --set full` launch at 1M galaxies), scaled by the batch size; null without it
        traffic, traffic_src = None, None
        if args.traffic is not None:
            traffic, traffic_src = args.traffic, "--traffic"
        else:
            prof = os.path.join(ROOT, "profiles", "r02_synth3_bf16_summary.txt")       # the final kernel (bfloat16 small terms)
            if not cross or not os.path.isfile(prof):
                prof = os.path.join(ROOT, "profiles", "r02_synth3_final_summary.txt")  # 3 x TF32, fused output
            if s3 and args.workload == "cfg2" and os.path.isfile(prof):
                import re
                txt = open(prof).read()
                rd = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", txt)
                wr = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", txt)
                if rd and wr:
                    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                    traffic = (float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)]) * n / 1e6
                    traffic_src = "ncu --set full capture in profiles/%s (one launch, 1M galaxies), scaled by batch size" % os.path.basename(prof)
        tf32 = measure_tf32_peak(dev)
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        kname = ("synth3_kernel (tcgen05: TF32 hi*hi + one bfloat16 MMA for the two small terms of the split product, weights "
                 "as the TMEM operand, fused epilogue)" if cross else
                 "synth3_kernel (3xTF32 tcgen05, weights as the TMEM operand, fused epilogue)" if s3 else
                 "synth_kernel (3xTF32 tcgen05 contraction + fused epilogue)")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": kname,
                    "kernel_ms": synth_ms_avg, "peak_source": f"{peak_src} bf16 sustained (MEASURED_PEAKS.json)",
                    "executed_tflops": executed,
                    "tf32_peak_measured": tf32,
                    "executed_tf32_time_equivalent_tflops": executed * tf32_passes / 3.0,
                    "executed_frac_of_tf32_sustained": executed * tf32_passes / 3.0 / tf32["tf32_tflops_sustained"],
                    "k_exec": k_exec, "k_dense": k_alg, "wavelength_chunks_computed_frac": frac_lam,
                    "note": "achieved counts ALGORITHMIC flops 2*K_exec*N_lam*C per galaxy (K_exec = 2*n_age for "
                            "DeltaConstant batches grouped by metallicity bracket, n_age*n_z otherwise); the kernel "
                            "executes 3x that (split operands for the 1e-5 tolerance: hi*hi as TF32, whose dense peak is "
                            "half the bf16 peak, and -- synth3_kernel -- the two small products as one bfloat16 MMA, i.e. "
                            "2 TF32 passes of tensor time per chunk; 3 TF32 passes in the other kernels), so frac <= 1/4 "
                            "(1/6) if every wavelength were multiplied; chunks of the axis "
                            "that no filter of a tile reaches are skipped (wavelength_chunks_computed_frac), which is "
                            "how frac can exceed 1/6; executed_tflops counts only the chunks actually multiplied "
                            "(estimated from the redshifts), executed_tf32_time_equivalent_tflops weighs the bfloat16 "
                            "products by half and executed_frac_of_tf32_sustained compares that with the TF32 GEMM rate "
                            "measured in this run",
                    "stage_ms": {"sort": float(np.mean([s[0] for s in stages])),
                                 "weights_igm": float(np.mean([s[1] for s in stages])),
                                 "contraction_epilogue": synth_ms_avg}}
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ns = min(args.cpu_sample, n)
            time_cpu(w, min(ns, 2000), threads)
            dtc = time_cpu(w, ns, threads)
            cpu = {"value": ns / dtc, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {ns} galaxies of the same workload, C/OpenMP float64 restatement "
                             f"(oracle/oracle_c.c), {dtc:.1f} s"}
        copy_value = world * n * e2e_steps / (copy_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": ("tf32 + bf16 small terms of the split product (fp32 accumulate); weights/IGM in f64" if cross else "tf32x3 (fp32 accumulate); weights/IGM in f64"),
            "data": "synthetic",
            "config": bench_config(args.workload, n, t, w.params),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps,
                    "api": "SynthEngine.submit/wait (sb2_synth_photometry_host_submit/_wait): one batch per step, two staging "
                           "slots, pinned host buffers; parameters as %s" % (
                               "the float32 draws of draw_from_hypercube (host_f32 transport, max_age derived on the device)"
                               if f32_transport else "float64 arrays"),
                    "blocking_call_value": world * n * e2e_steps / (e2e_sync_ms * 1e-3),
                    "blocking_call_api": "SynthEngine.photometry (sb2_synth_photometry_host), one blocking call per step",
                    "copy_only_value": copy_value,
                    "copy_only_gbs_per_gpu": (h2d + d2h) * e2e_steps / (copy_ms * 1e-3) / 1e9,
                    "limiter": ("host link: copying the same bytes with no kernels gives %.0f M galaxies/s" % (copy_value / 1e6))
                    if copy_value < 1.25 * e2e_value else "kernels / launch overhead (the host link alone is faster)"},
            "api_e2e": api,
            "gpu_launches": gpu_launches,
            "gpu_launches_note": "kernels of libsynference_b200.so launched inside the timed region, counted by the library "
                                 "(sb2_kernel_launches); CUB's radix-sort kernels (6 per step) are not included",
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
