#!/usr/bin/env python
"""Benchmark of the two B200 hot paths (BASELINE.json metric: CAVIaR fits/s & iters/s at N=1k,K=10k; NWD traces/s).

    python bench.py --gpus N --steps K --warmup W            # own arm (torchrun launches N>1)
    python bench.py --impl reference ...                     # reference arm: the CPU oracle on the host cores

A step = one pass of the CAVIaR hot path over one batch of B independent synthetic maps of the C3 shape
(N=1000 neurons, K=10000 trials x 900 samples, 10-target holograms, 3 powers, 50 iterations) per GPU, inputs
resident in HBM.  Prints ONE JSON line (rank 0).  The NWD path (C2: 20000 traces) is measured in the same run
and reported under "nwd".  Synthetic maps come from the device generator cm_simulate (every fit of a step is a
distinct map); `e2e` streams them from pinned host memory through circuitmap_b200.streaming.FitPipeline.
    python bench.py --traffic-probe                          # one fit launch and exit (run under ncu by the own arm)
"""
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
GOLD = os.path.join(ROOT, "tests", "golden")

ALGO_NOTE = "iters*(16NK+8N^2+12K)+3600K+8NK bytes per fit (SURVEY.md 8(d), dense fp32 accounting)"


def algorithmic_bytes_per_fit(N, K, iters):
    return iters * (16.0 * N * K + 8.0 * N * N + 12.0 * K) + 3600.0 * K + 8.0 * N * K


# ----------------------------------------------------------------------------------------- synthetic data
def synth_map(N, K, H, seed, T=900, powers=(45.0, 55.0, 65.0), connection_prob=0.1):
    """Synthetic compressive-mapping experiment with circuitmap.simulation's distributions (blockwise design):
    stim (N,K) f64 with exactly H targets per trial, psc (K,T) f64 = evoked + spontaneous PSCs + GP + iid noise."""
    rng = np.random.default_rng(seed)
    pw = np.sort(np.asarray(powers))[::-1]
    nh = int(np.ceil(N / H))
    tars = np.zeros((K, H), dtype=np.int64)
    pcol = np.zeros(K)
    k = 0
    while k < K:
        order = rng.permutation(N)
        for p in pw:
            for h in range(nh):
                if k >= K:
                    break
                sl = order[h * H:(h + 1) * H]
                tars[k, :len(sl)] = sl
                tars[k, len(sl):] = sl[0]
                pcol[k] = p
                k += 1
    perm = rng.permutation(K)
    tars, pcol = tars[perm], pcol[perm]
    stim = np.zeros((N, K))
    stim[tars, np.arange(K)[:, None]] = pcol[:, None]
    tau_r = rng.uniform(25, 60, N)
    tau_d = tau_r + rng.uniform(75, 250, N)
    phi0, phi1 = rng.uniform(0.2, 0.25, N), rng.uniform(10, 15, N)
    ncon = int(connection_prob * N)
    conn = rng.choice(N, ncon, replace=False)
    ns = int(np.ceil(0.2 * ncon))
    w = np.zeros(N)
    w[conn[:ns]] = rng.uniform(20, 40, ns)
    w[conn[ns:]] = rng.exponential(4, ncon - ns) + 9
    t = np.arange(T)
    psc = np.zeros((K, T))
    for n in conn:
        ks = np.nonzero(stim[n])[0]
        fr = 1 / (1 + np.exp(-(phi0[n] * stim[n, ks] - phi1[n])))
        sp = rng.random(ks.size) <= fr
        top = stim[n, ks] == pw[0]
        if top.any() and sp[top].mean() < 0.4:
            idx = np.nonzero(top & ~sp)[0]
            need = int(np.ceil((0.4 - sp[top].mean()) * top.sum()))
            sp[rng.choice(idx, min(need, idx.size), replace=False)] = True
        kern = np.exp(-t / tau_d[n]) - np.exp(-t / tau_r[n])
        kern = kern / (kern.sum() - 0.5 * (kern[0] + kern[-1]) + 1e-5)
        for kk in ks[sp]:
            s = int(160 + rng.gamma(1e4 / stim[n, kk] ** 2, 15.0))
            if s < T:
                ke = np.zeros(T)
                ke[s:] = kern[:T - s]
                psc[kk] += ke / (ke.sum() + 1e-5) * rng.lognormal(0, 0.01) * w[n]
    for kk in np.nonzero(rng.random(K) <= 0.05)[0]:
        tr = rng.uniform(25, 60)
        td = tr + rng.uniform(75, 250)
        st = rng.integers(1, T)
        ke = (np.exp(-(t - st) / td) - np.exp(-(t - st) / tr)) * (t > st)
        psc[kk] += rng.uniform(w[conn].min(), w[conn].max()) * ke / (ke.sum() + 1e-5)
    D = t[None, :] - t[:, None]
    Lc = np.linalg.cholesky(np.exp(-D ** 2 / (2 * 50.0 ** 2)) + 1e-6 * np.eye(T))
    psc += 4e-3 * (rng.standard_normal((K, T)) @ Lc.T)
    psc += rng.normal(0, 6e-4, (K, T))
    return stim, psc, w


def _gen_map_compact(a):
    """Pool worker: one synthetic map, returned compactly (sparse design, fp32 traces) to keep the pipe traffic small."""
    N, K, H, seed = a
    from threadpoolctl import threadpool_limits
    with threadpool_limits(1):                 # one BLAS thread per worker: the pool already fills the cores
        stim, psc, _ = synth_map(N, K, H, seed)
    nz = np.nonzero(stim)
    return (nz[0].astype(np.int32), nz[1].astype(np.int32), stim[nz].astype(np.float32), psc.astype(np.float32))


def synth_maps_parallel(specs, procs):
    """Distinct synthetic maps generated on the host cores in parallel (fork pool; call BEFORE CUDA is initialised)."""
    import multiprocessing as mp
    procs = max(1, min(procs, len(specs)))
    if procs == 1:
        return [_gen_map_compact(a) for a in specs]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_map_compact, specs, chunksize=1)


def dense_stim(m, N, K):
    s = np.zeros((N, K))
    s[m[0], m[1]] = m[2]
    return s


def synth_traces(K, seed, T=900):
    rng = np.random.default_rng(seed)
    t = np.arange(T)[None, :]
    tr = rng.uniform(25, 60, (K, 1))
    td = tr + rng.uniform(75, 250, (K, 1))
    d = rng.uniform(100, 400, (K, 1))
    with np.errstate(over="ignore"):
        ev = (np.exp(-(t - d) / td) - np.exp(-(t - d) / tr)) * (t > d)
    ev = ev / (ev.sum(1, keepdims=True) + 1e-5) * rng.uniform(5, 40, (K, 1))
    return ev + np.cumsum(rng.normal(0, 2e-4, (K, T)), axis=1) + rng.normal(0, 6e-4, (K, T))


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def measure_traffic(args, B):
    """DRAM bytes of ONE launch of the persistent fit kernel at this run's batch size, measured now: the own arm re-runs
    itself as `bench.py --traffic-probe` (one cm_caviar_fit call on B device-generated maps) under
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` and parses the CSV.  Returns (bytes, source) or (None, why)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:caviar_fit_kernel", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-probe",
           "--N", str(args.N), "--K", str(args.K), "--H", str(args.H), "--iters", str(args.iters), "--fits-per-gpu", str(B)]
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
    except Exception as e:                                   # noqa: BLE001
        return None, "ncu probe failed: %r" % (e,)
    total, seen = 0.0, 0
    import csv
    rows = list(csv.reader(out.stdout.splitlines()))
    hdr = next((r for r in rows if "Metric Name" in r and "Metric Value" in r), None)
    if hdr is None:
        return None, "ncu probe printed no metrics (rc %d): %s" % (out.returncode, (out.stderr or out.stdout)[-300:])
    iname, iunit, ival = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    for r in rows:
        if len(r) == len(hdr) and r[iname] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(r[iunit], None)
            if mult is None:
                return None, "unknown ncu unit %r" % r[iunit]
            total += float(r[ival].replace(",", "")) * mult
            seen += 1
    if seen != 2:
        return None, "ncu probe: expected 2 metric rows, got %d" % seen
    return total, "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, measured in this run"


def read_traffic(name, units=None):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json), scaled linearly to
    the number of fits / traces of this launch when the capture used a different batch size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    e = json.load(open(p)).get(name)
    if not e:
        return None
    per = e.get("fits_per_launch") or e.get("traces_per_launch")
    return e["bytes_per_launch"] * (units / per if (units and per) else 1.0)


# ----------------------------------------------------------------------------------------- CPU legs (oracle)
def cpu_caviar_sample(N, K, H, iters_total, sample_iters, seed=0):
    """The reference CPU path: NumPy fp64 restatement (oracle, reduced O(nnz) form) of caviar.py, timed on
    `sample_iters` iterations of one map and scaled to a full `iters_total`-iteration fit."""
    from oracle import caviar as oc
    stim, psc, _ = synth_map(N, K, H, seed)
    pr = oc.default_priors(N)
    t0 = time.time()
    oc.caviar(psc, stim, pr["mu"], pr["beta"], pr["shape"], pr["rate"], pr["phi"], pr["phi_cov"], iters=sample_iters,
              seed=1, msrmp=0.4, fn_scan=False)
    dt = time.time() - t0
    return dt, 1.0 / (dt / sample_iters * iters_total)


def cpu_nwd_sample(Ktr, seed=0):
    """The reference CPU path of the demixer: the torch module restated from the reference (bit-identical to it on
    the golden vectors), whole batch at once as nwd.py:44-46, all host threads."""
    import torch
    from oracle import nwd as onwd
    net = onwd.TorchNWD(dict(np.load(os.path.join(GOLD, "nwd_ie_ChroME2f_weights.npz"))))
    tr = synth_traces(Ktr, seed)
    net.demix(tr[:64].copy())
    t0 = time.time()
    net.demix(tr.copy())
    dt = time.time() - t0
    return dt, Ktr / dt, torch.get_num_threads()


def reference_arm(args, rank):
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    times = []
    for i in range(args.warmup + args.steps):
        dt, fps = cpu_caviar_sample(args.N, args.K, args.H, args.iters, args.ref_iters, seed=i % 2)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    fits_per_s = 1.0 / (ms / 1e3 / args.ref_iters * args.iters)
    ndt, ntps, nthr = cpu_nwd_sample(2000)
    sample = "%d of %d iterations of one N=%d,K=%d map per step, scaled linearly to a full fit" % (
        args.ref_iters, args.iters, args.N, args.K)
    line = {"impl": "reference", "metric": "caviar_fits_per_s", "value": fits_per_s, "unit": "fits/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "iters_per_s": fits_per_s * args.iters,
            "cpu_baseline": {"value": fits_per_s, "unit": "fits/s", "cores": int(torch.get_num_threads()),
                             "host_cores": cores, "kind": "port", "sample": sample,
                             "note": "JAX is not installable here; NumPy fp64 restatement of caviar.py (oracle, reduced form)"},
            "e2e": {"value": fits_per_s, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "nwd": {"metric": "nwd_traces_per_s", "value": ntps, "unit": "traces/s", "kind": "port",
                    "cores": nthr, "sample": "2000 traces in one batch (torch CPU restatement of NWDUNet)"}}
    print(json.dumps(line))


def workload_config(args, B=None):
    """The same dict in both arms (the driver compares them): the unit of work is ONE C3 fit."""
    return {"workload": "C3 compressive ensemble map: N=%d neurons, K=%d trials x 900 samples, H=%d targets, P=3 powers, "
                        "caviar %d iters; throughput of independent maps" % (args.N, args.K, args.H, args.iters),
            "l2_hygiene": "inputs larger than L2 (every fit of a step reads its own map: %.0f MB of traces + design)"
                          % ((args.N * args.K + args.K * 3600) / 1e6),
            "parallelism": "independent fits sharded over GPUs, no data-path collective"}


# ----------------------------------------------------------------------------------------- own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--K", type=int, default=10000)
    ap.add_argument("--H", type=int, default=10)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--fits-per-gpu", type=int, default=0, help="B; default = 2 x number of SMs (two fit CTAs per SM)")
    ap.add_argument("--maps", type=int, default=16, help="distinct maps held in pinned host memory for the e2e leg (tiled to B)")
    ap.add_argument("--nwd-traces", type=int, default=20000)
    ap.add_argument("--ref-iters", type=int, default=8, help="CPU oracle iterations per step (scaled to a full fit)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nwd", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (1024 maps of N=500, K=5000) secondary measurement")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 (one map of N=5000, K=100000) secondary measurement")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="fits per chunk of the e2e pipeline (default: a quarter wave)")
    ap.add_argument("--e2e-depth", type=int, default=0, help="chunks in flight in the e2e pipeline")
    ap.add_argument("--no-single", action="store_true", help="skip the single-fit latency measurement")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu DRAM-traffic probe of the fit kernel")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from circuitmap_b200 import NeuralDemixer, optimise, streaming, _lib
    from circuitmap_b200.simulation import simulate_batch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (own arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout on the first collective; keep stdout to the single JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    B = args.fits_per_gpu or 2 * sms
    N, K, H, iters = args.N, args.K, args.H, args.iters
    hbm_peak, bf16_burst, bf16_sust, peak_kind = measured_peaks()
    f64 = dict(dtype=torch.float64, device=dev)
    powers = np.array([45.0, 55.0, 65.0])
    opts = dict(iters=iters, msrmp=0.4)

    def default_priors(b, n):
        cov = torch.zeros(b, n, 2, 2, **f64)
        cov[..., 0, 0] = 0.1
        cov[..., 1, 1] = 1.0
        phi = torch.stack([0.1 * torch.ones(b, n, **f64), 5 * torch.ones(b, n, **f64)], -1).contiguous()
        return (torch.zeros(b, n, **f64), 10 * torch.ones(b, n, **f64), 1.0, 0.1, phi, cov)

    def gen_maps(count, n, k, seed0, chunk=32):
        """`count` DISTINCT synthetic maps from the device generator (csrc/simulate.cu, the distributions of
        circuitmap.simulation.simulate): uint8 power codes (count, n, k) + float32 traces (count, k, 900), resident in HBM."""
        codes = torch.empty((count, n, k), dtype=torch.uint8, device=dev)
        traces = torch.empty((count, k, 900), dtype=torch.float32, device=dev)
        wsim = None
        for lo in range(0, count, chunk):
            hi = min(lo + chunk, count)
            r = simulate_batch([seed0 + i for i in range(lo, hi)], device=dev, workspace=wsim, N=n, trials=k, H=H,
                               connection_prob=0.1, out=dict(codes=codes[lo:hi], psc=traces[lo:hi]))
            wsim = r["_workspace"]
            assert int(r["status"].sum().item()) == 0
            if r["codes"].data_ptr() != codes[lo:hi].data_ptr():
                codes[lo:hi].copy_(r["codes"]); traces[lo:hi].copy_(r["psc"])
        return codes, traces

    if args.traffic_probe:                       # one launch of the timed configuration, for `ncu` (see measure_traffic)
        codes, traces = gen_maps(B, N, K, 1)
        out = optimise.caviar_batched(codes, powers, *default_priors(B, N), psc=traces, seeds=list(range(1, B + 1)),
                                      nnz_cap=K * H, want_lam=False, **opts)
        torch.cuda.synchronize()
        return

    # ---- synthetic inputs: B distinct maps per rank, generated on the device (seeded per rank) ----
    t_gen = time.time()
    stim, psc = gen_maps(B, N, K, 1 + 100000 * rank)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    nnz = K * H
    pri = default_priors(B, N)
    seeds = [1 + b + 100000 * rank for b in range(B)]
    ws = [None]
    outbuf = [None]

    def step():
        out = optimise.caviar_batched(stim, powers, *pri, psc=psc, seeds=seeds, nnz_cap=nnz, want_lam=False, lam_csr=True,
                                      workspace=ws[0], out=outbuf[0], **opts)
        ws[0] = out["_workspace"]
        outbuf[0] = out
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    assert int(out["status"].sum().item()) == 0, "device status reported an error"
    launches_per_step = out["launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    e0.record()
    for _ in range(args.steps):
        out = step()
        kern_ms.append(lib.cm_last_main_kernel_ms())        # waits for this step's fit kernel (same stream)
    e1.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    fits_per_s = world * B / (ms_step / 1e3)
    kms = float(np.mean(kern_ms))
    algo = algorithmic_bytes_per_fit(N, K, iters) * B
    achieved = algo / (kms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "caviar_fit_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None,
                "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                "kernel_ms_per_launch": kms, "kernel_share_of_step": kms / ms_step,
                "algorithmic_bytes_per_launch": algo, "algorithmic_model": ALGO_NOTE,
                "note": "kernel works on a CSR/CSC index of the design (supp(lam) within supp(stim)); achieved is the "
                        "DENSE algorithmic byte count over time, traffic is what DRAM actually moved, dram_frac = traffic / "
                        "kernel time / peak is the bandwidth the kernel really uses"}
    connected = int((out["mu"][0] != 0).sum().item())
    nnz_lam = int(out["lam_csr_ptr"][0, -1].item())

    # ---- ONE C3 fit alone on the GPU: the latency north_star sets the roofline target on ----
    single = None
    if not args.no_single:
        pri1 = default_priors(1, N)
        ws1, o1b, t1 = None, None, []
        for rep in range(4):
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            o1 = optimise.caviar_batched(stim[:1], powers, *pri1, psc=psc[:1], seeds=[seeds[0]], nnz_cap=nnz, want_lam=False,
                                         lam_csr=True, workspace=ws1, out=o1b, **opts)
            ws1, o1b = o1["_workspace"], o1
            k1 = lib.cm_last_main_kernel_ms()
            q1.record()
            torch.cuda.synchronize()
            t1.append((q0.elapsed_time(q1), k1))
        ms1, k1 = min(t1[1:])
        algo1 = algorithmic_bytes_per_fit(N, K, iters)
        single = {"metric": "caviar_single_fit_latency_ms", "value": ms1, "unit": "ms", "higher_is_better": False,
                  "fits_per_s": 1e3 / ms1, "iters_per_s": iters * 1e3 / ms1, "kernel_ms": k1,
                  "roofline": {"bound": "hbm", "achieved": algo1 / (ms1 / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": algo1 / (ms1 / 1e3) / 1e9 / hbm_peak, "roofline_time_ms": 1e3 * algo1 / (hbm_peak * 1e9),
                               "algorithmic_model": ALGO_NOTE},
                  "note": "one map, inputs resident in HBM (L2-resident working set), whole cm_caviar_fit call incl. prologue"}
        del o1, o1b, ws1

    # ---- end to end through host buffers (circuitmap_b200.streaming.FitPipeline): pinned float32 traces + uint8 design codes
    # -> H2D -> [NeuralDemixer -> y / sum-of-squares hand-off] -> cm_caviar_fit -> D2H of the state with lam as CSR ----
    e2e = None
    if not args.no_e2e:
        npin = max(1, min(B, args.maps))
        hs_pin = [stim[i].cpu().pin_memory() for i in range(npin)]
        hp_pin = [psc[i].cpu().pin_memory() for i in range(npin)]
        hs_coo = [optimise.codes_to_coo(stim[i]) for i in range(npin)]        # the same designs as sparse triples
        hs = [hs_coo[b % npin] for b in range(B)]
        hp = [hp_pin[b % npin] for b in range(B)]
        # what this box's host -> device path delivers from pinned memory (a pure copy of the e2e buffers, no kernels)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        probe = torch.empty((min(B, 64),) + tuple(hp_pin[0].shape), dtype=torch.float32, device=dev)
        for rep in range(2):
            torch.cuda.synchronize()
            h0.record()
            for i in range(probe.shape[0]):
                probe[i].copy_(hp[i], non_blocking=True)
            h1.record()
            torch.cuda.synchronize()
        h2d_gbs = probe.numel() * 4 / (h0.elapsed_time(h1) / 1e3) / 1e9
        del probe
        # four chunks of half a wave in flight: their fit CTAs share the SMs, the first kernels start after 1/4 of the upload
        e2e_chunk, e2e_depth = args.e2e_chunk or max(1, sms // 2), args.e2e_depth or 4
        dem_e2e = NeuralDemixer(path=os.path.join(GOLD, "nwd_ie_ChroME2f_weights.npz"), device=dev, precision="fp16")
        res = {}
        for label, demixer in (("fit", None), ("pipeline", dem_e2e)):
            pipe = streaming.FitPipeline(N, K, powers, chunk=e2e_chunk, nnz_cap=nnz, device=dev, demixer=demixer, design="coo",
                                         depth=e2e_depth, **opts)
            got = [0]

            def on_result(lo, hi, views, got=got):
                got[0] += int((views["mu"] != 0).sum().item())      # the host reads every result (connected counts)

            assert pipe.run(hs, hp, seeds, on_result) == 0
            sync_all()
            # K steps of B fits as ONE continuous stream (what a sweep over many maps is): every step's uploads and
            # downloads are inside the timed region, the pipeline is filled and drained once, not once per step
            n_e2e = max(1, min(args.steps, 8))
            t0 = time.time()
            assert pipe.run(hs * n_e2e, hp * n_e2e, list(seeds) * n_e2e, on_result) == 0
            torch.cuda.synchronize()
            sync_all()
            dt = torch.tensor([(time.time() - t0) / n_e2e], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            res[label] = {"value": world * B / float(dt.item()), "unit": "fits/s", "ms_per_step": 1e3 * float(dt.item()),
                          "h2d_bytes_per_step": B * pipe.h2d_bytes_per_fit, "d2h_bytes_per_step": B * pipe.d2h_bytes_per_fit,
                          "h2d_gbs_measured": h2d_gbs,
                          "h2d_bound_fits_per_s": world * h2d_gbs * 1e9 / pipe.h2d_bytes_per_fit}
            del pipe
            torch.cuda.empty_cache()
        e2e = dict(res["fit"])
        e2e["note"] = ("circuitmap_b200.streaming.FitPipeline: pinned host buffers in the formats the data has (float32 traces, the "
                       "design as sparse (neuron, trial, power code) triples; %d distinct maps tiled to %d fits) -> H2D -> "
                       "cm_expand_stim_coo -> cm_caviar_fit -> D2H of mu, beta, shape, rate, phi, "
                       "phi_cov, z and lam as CSR, read by the host; %d chunks of %d fits in flight (upload, kernels and "
                       "download overlap; the chunks' fit kernels share the SMs); %d steps streamed back to back as one run "
                       "(pipeline filled and drained once)" % (npin, B, e2e_depth, e2e_chunk, n_e2e))
        e2e["pipeline_with_demixer"] = dict(res["pipeline"], note="the same with RAW traces in and the fp16 tensor-core "
                                            "NeuralDemixer in front of the fit (README.md:28-51 of the reference: demix -> fit); "
                                            "the demixed traces never leave the device, only y = trapz and sum x^2 reach the fit")
        del hs, hp, hs_pin, hp_pin, hs_coo, dem_e2e
    # ---- the reference-facing call itself: Model(N).fit(psc, stim) with NumPy arrays in and out, one map (rank 0) ----
    if e2e is not None and rank == 0:
        import contextlib
        import io
        from circuitmap_b200 import Model
        s_np = (powers[(stim[0].long() - 1).clamp(min=0).cpu().numpy()] * (stim[0].cpu().numpy() > 0)).astype(np.float64)
        p_np = psc[0].double().cpu().numpy()
        tcall = []
        for _ in range(3):
            mdl = Model(N)
            torch.cuda.synchronize()
            t0 = time.time()
            with contextlib.redirect_stdout(io.StringIO()):
                mdl.fit(p_np, s_np, method="caviar", fit_options=dict(opts, seed=1))
            tcall.append(time.time() - t0)
        e2e["drop_in_call"] = {"fits_per_s": 1.0 / min(tcall), "ms_per_fit": 1e3 * min(tcall),
                               "note": "Model(N).fit(psc, stim, 'caviar') on pageable float64 NumPy arrays, full state incl. the dense "
                                       "N x K lam back as NumPy (one map at a time: latency of the drop-in call, not batch throughput)"}
        del s_np, p_np, mdl
    del stim, psc, out
    ws[0] = None
    outbuf[0] = None
    torch.cuda.empty_cache()

    # ---- measured DRAM traffic of the dominant kernel (rank 0, N=1): a second process under ncu, outside every timed region ----
    if rank == 0 and world == 1 and not args.no_traffic:
        tb, src = measure_traffic(args, B)
        if tb is None:
            tb, src = read_traffic("caviar_fit_kernel", B), "profiles/traffic.json (round-1 capture, scaled) -- live probe failed: " + src
        roofline["traffic"] = tb
        roofline["traffic_source"] = src
        if tb:
            roofline["dram_frac"] = tb / (kms / 1e3) / 1e9 / hbm_peak
            roofline["traffic_per_fit_gb"] = tb / B / 1e9

    # ---- C4 (BASELINE.json configs[3]): sweep of 1024 independent maps N=500, K=5000, sharded 1024 / world per GPU ----
    c4 = None
    if not args.no_c4:
        N4, K4, Btot = 500, 5000, 1024
        lo4, hi4 = (rank * Btot) // world, ((rank + 1) * Btot) // world
        B4 = hi4 - lo4
        nnz4 = K4 * H
        stim4, psc4 = gen_maps(B4, N4, K4, 7000000 + lo4, chunk=128)
        pri4 = default_priors(B4, N4)
        seeds4 = [1 + lo4 + b for b in range(B4)]
        ws4 = None
        for _ in range(2):
            o4 = optimise.caviar_batched(stim4, powers, *pri4, psc=psc4, seeds=seeds4, nnz_cap=nnz4, want_lam=False,
                                         workspace=ws4, **opts)
            ws4 = o4["_workspace"]
        sync_all()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        o4 = optimise.caviar_batched(stim4, powers, *pri4, psc=psc4, seeds=seeds4, nnz_cap=nnz4, want_lam=False,
                                     workspace=ws4, **opts)
        k4 = lib.cm_last_main_kernel_ms()
        c1.record()
        sync_all()
        t4 = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        assert int(o4["status"].sum().item()) == 0
        algo4 = algorithmic_bytes_per_fit(N4, K4, iters) * B4
        c4 = {"metric": "caviar_fits_per_s", "value": Btot / (float(t4.item()) / 1e3), "unit": "fits/s", "scaling": "strong",
              "config": {"workload": "C4 batched sweep: 1024 independent maps N=500, K=5000, H=%d, %d iters, sharded %d per GPU "
                                     "(all maps distinct, device generator; traces fp32, design as uint8 codes, posteriors "
                                     "without the dense lam)" % (H, iters, B4)},
              "ms_per_step": float(t4.item()), "connected_in_fit0": int((o4["mu"][0] != 0).sum().item()),
              "roofline": {"bound": "hbm", "kernel": "caviar_fit_kernel", "achieved": algo4 / (k4 / 1e3) / 1e9,
                           "peak": hbm_peak, "unit": "GB/s", "frac": algo4 / (k4 / 1e3) / 1e9 / hbm_peak, "traffic": None,
                           "kernel_ms_per_launch": k4, "algorithmic_model": ALGO_NOTE}}
        del stim4, psc4, o4, ws4
        torch.cuda.empty_cache()

    # ---- C5 (BASELINE.json configs[4]): ONE large map N=5000, K=100000 on one GPU (rank 0).  The K-sharded multi-GPU fit
    # is not built (measured 2-GPU exchange prototype); this is the single-GPU latency of the same map. ----
    c5 = None
    if not args.no_c5 and rank == 0:
        N5, K5 = 5000, 100000
        stim5, psc5 = gen_maps(1, N5, K5, 9000)
        pri5 = default_priors(1, N5)
        ws5 = None
        t5 = []
        for rep in range(2):
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            o5 = optimise.caviar_batched(stim5, powers, *pri5, psc=psc5, seeds=[1], nnz_cap=K5 * H, want_lam=False,
                                         workspace=ws5, **opts)
            ws5 = o5["_workspace"]
            q1.record()
            torch.cuda.synchronize()
            t5.append(q0.elapsed_time(q1))
        assert int(o5["status"].sum().item()) == 0
        algo5 = algorithmic_bytes_per_fit(N5, K5, iters)
        c5 = {"metric": "caviar_fits_per_s", "value": 1e3 / t5[-1], "unit": "fits/s", "ms_per_fit": t5[-1],
              "iters_per_s": iters * 1e3 / t5[-1],
              "config": {"workload": "C5 large single map: N=5000, K=100000, H=%d, %d iters, ONE B200 (one persistent fit CTA + 127 "
                                     "helper CTAs; the K-sharded 8-GPU variant is not built: DESIGN.md 5)" % (H, iters)},
              "connected": int((o5["mu"][0] != 0).sum().item()),
              "roofline": {"bound": "hbm", "achieved": algo5 / (t5[-1] / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": algo5 / (t5[-1] / 1e3) / 1e9 / hbm_peak, "algorithmic_model": ALGO_NOTE}}
        del stim5, psc5, o5, ws5
        torch.cuda.empty_cache()

    # ---- NWD (C2): K traces through cm_nwd_forward ----
    nwd = None
    if not args.no_nwd:
        Kt = args.nwd_traces
        dem = NeuralDemixer(path=os.path.join(GOLD, "nwd_ie_ChroME2f_weights.npz"), device=dev, precision="fp16")
        htr = torch.from_numpy(synth_traces(Kt, seed=rank)).pin_memory()
        x32 = htr.to(dev).float()
        o32 = torch.empty_like(x32)

        def time_nwd(n_it=5):
            for _ in range(3):
                dem.forward_device(x32, out=o32)
            sync_all()
            kms_ = []
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0.record()
            for _ in range(n_it):
                dem.forward_device(x32, out=o32)
                kms_.append(lib.cm_last_main_kernel_ms())
            n1.record()
            sync_all()
            t_ = torch.tensor([n0.elapsed_time(n1) / n_it], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            return float(t_.item()), float(np.mean(kms_))

        dem.set_precision("fp32")
        ms_fp32, _ = time_nwd()
        dem.set_precision("fp16")
        ms_tc, kms_tc = time_nwd()
        nms = torch.tensor([ms_tc], dtype=torch.float64, device=dev)
        kms_n = [kms_tc]
        tps = world * Kt / (ms_tc / 1e3)
        flops = 2 * 8435200.0 * Kt
        ach = flops / (kms_tc / 1e3) / 1e12
        hout = torch.empty((Kt, 900), dtype=torch.float64).pin_memory()
        x64 = torch.empty((Kt, 900), **f64)
        o64 = torch.empty((Kt, 900), **f64)

        _streaming = streaming
        nstreams = _streaming._Streams(dev)

        def nwd_e2e():
            _streaming.demix_pinned(dem, htr, hout, x64, o64, chunk=max(1, Kt // 8), streams=nstreams)
            torch.cuda.synchronize()

        nwd_e2e()
        sync_all()
        t0 = time.time()
        for _ in range(3):
            nwd_e2e()
        sync_all()
        edt = torch.tensor([(time.time() - t0) / 3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(edt, op=dist.ReduceOp.MAX)
        nwd = {"metric": "nwd_traces_per_s", "value": tps, "unit": "traces/s", "dtype": "fp16 operands (fp32 accumulate)",
               "fp32_cuda_core_path_traces_per_s": world * Kt / (ms_fp32 / 1e3),
               "error_bound": "fp16 tensor-core path: max-abs <= 2e-2, relative L2 <= 3e-3 on unit-normalised traces (tests/test_nwd_gpu.py)",
               "config": {"workload": "C2: NeuralDemixer nwd_ie_ChroME2f forward on %d x 900 traces per GPU" % Kt,
                          "l2_hygiene": "in+out = %.0f MB > L2" % (2 * Kt * 3600 / 1e6)},
               "ms_per_step": ms_tc,
               "roofline": {"bound": "tensor", "kernel": "nwd_forward_mt_kernel", "achieved": ach, "peak": bf16_burst,
                            "unit": "TFLOP/s", "frac": ach / bf16_burst, "traffic": read_traffic("nwd_forward_mt_kernel", Kt),
                            "kernel_ms_per_launch": kms_tc,
                            "peak_kind": peak_kind + " bf16 burst (MEASURED_PEAKS.json); operands are fp16 (tcgen05 "
                                                     "kind::f16, same dense peak as bf16); achieved counts the network's "
                                                     "16.87 MFLOP/trace (SURVEY.md App. C), not the zero taps of the "
                                                     "widened implicit GEMMs"},
               "e2e": {"value": world * Kt / float(edt.item()), "unit": "traces/s", "h2d_bytes_per_step": Kt * 7200,
                       "d2h_bytes_per_step": Kt * 7200}}
        if rank == 0:       # the reference-facing call itself: NeuralDemixer(...)(traces) on a pageable float64 NumPy array
            tr_np = htr.numpy().copy()
            dem(tr_np[:256], verbose=False)
            tcall = []
            for _ in range(3):
                t0 = time.time()
                dem(tr_np, verbose=False)
                tcall.append(time.time() - t0)
            nwd["e2e"]["drop_in_call"] = {"traces_per_s": Kt / min(tcall), "ms": 1e3 * min(tcall),
                                          "note": "NeuralDemixer(path, precision='fp16')(traces) with a pageable float64 NumPy "
                                                  "array in and out"}

    # ---- CPU baseline on the host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dt, fps = cpu_caviar_sample(N, K, H, iters, args.ref_iters)
        cpu = {"value": fps, "unit": "fits/s", "cores": int(torch.get_num_threads()), "host_cores": os.cpu_count(),
               "kind": "port",
               "sample": "%d of %d iterations of one N=%d,K=%d map (%.1f s), scaled linearly to a full fit" % (
                   args.ref_iters, iters, N, K, dt)}
        if nwd is not None:
            ndt, ntps, nthr = cpu_nwd_sample(2000)
            nwd["cpu_baseline"] = {"value": ntps, "unit": "traces/s", "cores": nthr, "kind": "port",
                                   "sample": "2000 traces in one batch (%.2f s)" % ndt}

    if rank == 0:
        line = {"metric": "caviar_fits_per_s", "value": fits_per_s, "unit": "fits/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (device generator cm_simulate with circuitmap.simulation's distributions; every fit a distinct map)",
                "config": workload_config(args, B), "iters_per_s": fits_per_s * iters,
                "fits_per_gpu_per_step": B, "data_generation_s": t_gen,
                "connected_in_fit0": connected, "lam_nnz_in_fit0": nnz_lam, "roofline": roofline, "single_fit": single,
                "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks, "nwd": nwd, "c4": c4, "c5": c5}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
