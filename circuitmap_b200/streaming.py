"""Host-buffer entry points with copy/compute overlap.

`NeuralDemixer.__call__` and `Model.fit` keep the reference's NumPy-in / NumPy-out surface (one synchronous round
trip).  Sweeps that keep their inputs in *pinned* host memory use the two functions below instead: the batch is cut
into chunks, and chunk i+1's host->device copy, chunk i's kernels and chunk i-1's device->host copy run on three
streams, so a step costs about max(H2D, kernel, D2H) instead of their sum.  Results are bit-identical to the
one-shot calls (every trace / fit is independent of its batch mates).
"""
from . import optimise

_STATE_KEYS = ("mu", "beta", "shape", "rate", "phi", "phi_cov", "z", "lam")


class _Streams:
    def __init__(self, device):
        import torch
        self.copy_in, self.compute, self.copy_out = (torch.cuda.Stream(device=device) for _ in range(3))


def demix_pinned(demixer, src, dst, dev_in, dev_out, chunk=5000, monotone_filter_start=500, streams=None):
    """src / dst: pinned host tensors (K, 900); dev_in / dev_out: device tensors of the same shape (reused across
    calls).  Returns when dst is complete."""
    import torch
    dev = dev_in.device
    st = streams or _Streams(dev)
    K = src.shape[0]
    cur = torch.cuda.current_stream(dev)
    for s in (st.copy_in, st.compute, st.copy_out):
        s.wait_stream(cur)
    for lo in range(0, K, chunk):
        hi = min(lo + chunk, K)
        with torch.cuda.stream(st.copy_in):
            dev_in[lo:hi].copy_(src[lo:hi], non_blocking=True)
            ready = st.copy_in.record_event()
        with torch.cuda.stream(st.compute):
            st.compute.wait_event(ready)
            demixer.forward_device(dev_in[lo:hi], monotone_filter_start, out=dev_out[lo:hi])
            done = st.compute.record_event()
        with torch.cuda.stream(st.copy_out):
            st.copy_out.wait_event(done)
            dst[lo:hi].copy_(dev_out[lo:hi], non_blocking=True)
    st.copy_out.synchronize()
    return dst


def fit_pinned(host_stim, host_psc, dev_stim, dev_psc, powers, priors, seeds, host_out, chunk, nnz_cap=None,
               workspaces=None, streams=None, on_chunk=None, **fit_options):
    """B independent fits from pinned host inputs to pinned host outputs.
       host_stim[b] (N, K), host_psc[b] (K, T): pinned tensors (sequences of length B; entries may repeat);
       dev_stim (B, N, K), dev_psc (B, K, T): device staging, reused across calls;
       priors: tuple (mu0, beta0, shape0, rate0, phi0, phi_cov0) of device tensors with leading dimension B (scalars
       for shape0 / rate0);  host_out: dict of pinned tensors for mu, beta, shape, rate, phi, phi_cov, z, lam
       ('lam' optional).  Every tensor of host_out has either leading dimension >= B (fit b lands in row b) or is a RING
       of R >= chunk slabs (dense lam is 8 N K bytes per fit): ring tensors are filled chunk by chunk, slot
       (chunk index mod R // chunk), and `on_chunk(lo, hi, views)` is called -- after the chunk's copies have completed
       and before its slabs are reused -- with views[k] = the (hi - lo) rows holding fits lo..hi-1 (all keys, ring or not).
       A ring without on_chunk would silently drop results and is refused.
       Returns the per-fit status (device tensor, all zero on success)."""
    import torch
    dev = dev_stim.device
    st = streams or _Streams(dev)
    B = dev_stim.shape[0]
    cur = torch.cuda.current_stream(dev)
    for s in (st.copy_in, st.compute, st.copy_out):
        s.wait_stream(cur)
    want_lam = "lam" in host_out
    keys = [k for k in _STATE_KEYS if k in host_out]
    rings = {k: host_out[k].shape[0] for k in keys if host_out[k].shape[0] < B}
    if rings:
        if on_chunk is None:
            raise ValueError("fit_pinned: host_out[%r] has %d < B = %d slabs: pass on_chunk to drain the ring "
                             "(results would be overwritten)" % (next(iter(rings)), next(iter(rings.values())), B))
        if min(rings.values()) < min(chunk, B):
            raise ValueError("fit_pinned: a ring needs at least `chunk` slabs")
    nslots = min([r // chunk for r in rings.values()] + [2]) if rings else 2
    mu0, beta0, shape0, rate0, phi0, cov0 = priors
    workspaces = workspaces if workspaces is not None else {}
    status, keep = [], []
    pending = None

    def drain(item):
        lo, hi, slot, ev = item
        ev.synchronize()
        if on_chunk is not None:
            views = {k: (host_out[k][slot * chunk:slot * chunk + hi - lo] if k in rings else host_out[k][lo:hi])
                     for k in keys}
            on_chunk(lo, hi, views)

    for ci, lo in enumerate(range(0, B, chunk)):
        hi = min(lo + chunk, B)
        with torch.cuda.stream(st.copy_in):
            for b in range(lo, hi):
                dev_stim[b].copy_(host_stim[b], non_blocking=True)
                dev_psc[b].copy_(host_psc[b], non_blocking=True)
            ready = st.copy_in.record_event()
        with torch.cuda.stream(st.compute):
            st.compute.wait_event(ready)
            out = optimise.caviar_batched(dev_stim[lo:hi], powers, mu0[lo:hi], beta0[lo:hi], shape0, rate0, phi0[lo:hi],
                                          cov0[lo:hi], psc=dev_psc[lo:hi], seeds=list(seeds[lo:hi]), nnz_cap=nnz_cap,
                                          want_lam=want_lam, workspace=workspaces.get(0), **fit_options)
            workspaces[0] = out["_workspace"]           # chunks run back to back on one stream: one workspace
            done = st.compute.record_event()
        keep.append(out)                      # outputs stay alive until the copies below have run
        status.append(out["status"])
        slot = ci % nslots
        if pending is not None and rings and pending[2] == slot:
            drain(pending)                    # the slabs this chunk lands in still hold an undelivered chunk
            pending = None
        with torch.cuda.stream(st.copy_out):
            st.copy_out.wait_event(done)
            for k in keys:
                if k in rings:
                    host_out[k][slot * chunk:slot * chunk + hi - lo].copy_(out[k], non_blocking=True)
                else:
                    host_out[k][lo:hi].copy_(out[k], non_blocking=True)
            ev = st.copy_out.record_event()
        if pending is not None:
            drain(pending)                    # deliver chunk ci-1 while chunk ci is in flight
        pending = (lo, hi, slot, ev)
    if pending is not None:
        drain(pending)
    st.copy_out.synchronize()
    st.compute.synchronize()
    return torch.cat(status)


class FitPipeline:
    """The pipeline users run (README.md:28-51 of the reference: demix -> fit), streamed from pinned host memory in the
    compact formats the data really has:

        host (pinned):  traces (K, T) float32   +   design (N, K) uint8 power codes (cm_pack_stim_u8)
                                                     or its nnz (neuron, trial, code) triples (design='coo': 9 nnz bytes)
          -- H2D -->  [NeuralDemixer forward: y = trapz, sum x^2 stay on the device]  -->  cm_caviar_fit
          -- D2H -->  mu, beta, shape, rate, phi, phi_cov, z  +  lam as CSR (8 nnz bytes instead of 8 N K)

    `depth` chunks are in flight: one upload stream, one download stream and `depth` compute streams whose fit kernels
    (forced to the two-CTAs-per-SM variant) share the SMs, so the GPU runs a full complement of fits while the next chunk's
    traces are still on the bus.  `on_result(lo, hi, views)` receives pinned host views of fits lo..hi-1 in order (valid until
    it returns); `views['lam']` is a list of optimise.CsrLam.  Bit-identical to calling NeuralDemixer / caviar_batched on
    each map (every fit is independent of its batch mates).
    """

    OUT_KEYS = ("mu", "beta", "shape", "rate", "phi", "phi_cov", "z", "status", "lam_csr_val", "lam_csr_col",
                "lam_csr_ptr")

    def __init__(self, N, K, powers, chunk, nnz_cap, device=None, demixer=None, T=900, priors=None, design="codes",
                 depth=2, **fit_options):
        import numpy as np
        import torch
        self.N, self.K, self.T, self.chunk, self.nnz_cap = int(N), int(K), int(T), int(chunk), int(nnz_cap)
        self.depth = max(1, int(depth))
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.powers = np.ascontiguousarray(powers, dtype=np.float64)
        self.demixer = demixer
        if design not in ("codes", "coo"):
            raise ValueError("design must be 'codes' (dense uint8 matrix) or 'coo' (sparse triples)")
        self.design = design           # 'coo': host_stim[b] = (neuron int32, trial int32, code uint8) pinned triples of length nnz_b
        self.fit_options = dict(fit_options)
        d, c, D = self.dev, self.chunk, self.depth
        self.dev_stim = [torch.empty((c, N, K), dtype=torch.uint8, device=d) for _ in range(D)]
        self.dev_psc = [torch.empty((c, K, T), dtype=torch.float32, device=d) for _ in range(D)]
        self.dev_dem = [torch.empty((c * K, T), dtype=torch.float32, device=d) if demixer is not None else None for _ in range(D)]
        if design == "coo":
            self.dev_coo = [(torch.empty((c, self.nnz_cap), dtype=torch.int32, device=d),
                             torch.empty((c, self.nnz_cap), dtype=torch.int32, device=d),
                             torch.empty((c, self.nnz_cap), dtype=torch.uint8, device=d)) for _ in range(D)]
        self.dev_out = [None] * D
        self.workspace = [None] * D
        f64 = dict(dtype=torch.float64, device=d)
        if priors is None:                                                       # model.py:24-31
            cov = torch.zeros(c, N, 2, 2, **f64)
            cov[..., 0, 0] = 1e-1
            cov[..., 1, 1] = 1e0
            phi = torch.stack([1e-1 * torch.ones(c, N, **f64), 5e0 * torch.ones(c, N, **f64)], -1).contiguous()
            priors = (torch.zeros(c, N, **f64), 1e1 * torch.ones(c, N, **f64), 1.0, 1e-1, phi, cov)
        self.priors = priors
        shapes = dict(mu=(c, N), beta=(c, N), shape=(c,), rate=(c,), phi=(c, N, 2), phi_cov=(c, N, 2, 2), z=(c, K),
                      lam_csr_val=(c, self.nnz_cap))
        self.host_out = [dict({k: torch.empty(v, dtype=torch.float64).pin_memory() for k, v in shapes.items()},
                              status=torch.empty((c,), dtype=torch.int32).pin_memory(),
                              lam_csr_col=torch.empty((c, self.nnz_cap), dtype=torch.int32).pin_memory(),
                              lam_csr_ptr=torch.empty((c, N + 1), dtype=torch.int32).pin_memory()) for _ in range(D)]
        self.copy_in, self.copy_out = torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)
        self.compute = [torch.cuda.Stream(device=d) for _ in range(D)]
        self.h2d_bytes_per_fit = (N * K if design == "codes" else 9 * self.nnz_cap) + K * T * 4
        self.d2h_bytes_per_fit = sum(t[0].numel() * t.element_size() for t in self.host_out[0].values())

    def run(self, host_stim, host_psc, seeds, on_result=None):
        """host_stim[b]: pinned uint8 (N, K) -- or, with design='coo', pinned triples (neuron, trial, code) of the non-zero
        entries (optimise.codes_to_coo); host_psc[b]: pinned float32 (K, T) -- raw traces when the pipeline has a demixer,
        demixed ones otherwise.  Returns the number of fits that reported a non-zero status."""
        import torch
        c, K, T, D = self.chunk, self.K, self.T, self.depth
        B = len(host_stim)
        cur = torch.cuda.current_stream(self.dev)
        for s in [self.copy_in, self.copy_out] + self.compute:
            s.wait_stream(cur)
        mu0, beta0, shape0, rate0, phi0, cov0 = self.priors
        ev_compute = [None] * D            # compute of the chunk that last used set s
        ev_out = [None] * D                # D2H of the chunk that last used set s
        pending, failed = [], 0

        def drain(item):
            nonlocal failed
            lo, hi, s_, ev = item
            ev.synchronize()
            ho = self.host_out[s_]
            n = hi - lo
            failed += int((ho["status"][:n] != 0).sum().item())
            if on_result is not None:
                views = {k: ho[k][:n] for k in ("mu", "beta", "shape", "rate", "phi", "phi_cov", "z", "status")}
                views["lam"] = [optimise.CsrLam(ho["lam_csr_val"][i].numpy(), ho["lam_csr_col"][i].numpy(),
                                                ho["lam_csr_ptr"][i].numpy(), K) for i in range(n)]
                on_result(lo, hi, views)

        for ci, lo in enumerate(range(0, B, c)):
            hi = min(lo + c, B)
            n, s_ = hi - lo, ci % D
            while pending and pending[0][2] == s_:
                drain(pending.pop(0))                           # the pinned output slabs of this set are handed over first
            with torch.cuda.stream(self.copy_in):
                if ev_compute[s_] is not None:
                    self.copy_in.wait_event(ev_compute[s_])     # the kernels that read this staging set are done
                for i, b in enumerate(range(lo, hi)):
                    if self.design == "coo":
                        hn, ht, hc = host_stim[b]
                        m = hc.numel()
                        dn, dt_, dc = self.dev_coo[s_]
                        dn[i, :m].copy_(hn, non_blocking=True)
                        dt_[i, :m].copy_(ht, non_blocking=True)
                        dc[i, :m].copy_(hc, non_blocking=True)
                    else:
                        self.dev_stim[s_][i].copy_(host_stim[b], non_blocking=True)
                    self.dev_psc[s_][i].copy_(host_psc[b], non_blocking=True)
                ready = self.copy_in.record_event()
            cs = self.compute[s_]
            with torch.cuda.stream(cs):
                cs.wait_event(ready)
                if ev_out[s_] is not None:
                    cs.wait_event(ev_out[s_])                   # this output set has been copied out
                if self.design == "coo":                        # scatter the triples into the dense code matrices (HBM only)
                    dn, dt_, dc = self.dev_coo[s_]
                    for i, b in enumerate(range(lo, hi)):
                        m = host_stim[b][2].numel()
                        optimise.expand_coo(dn[i, :m], dt_[i, :m], dc[i, :m], self.N, self.K, out=self.dev_stim[s_][i])
                if self.demixer is not None:
                    _, y, ss = self.demixer.forward_device(self.dev_psc[s_][:n].reshape(n * K, T), out=self.dev_dem[s_][:n * K],
                                                           stats=True)
                    src = dict(y=y.view(n, K), ss=ss.view(n, K))
                else:
                    src = dict(psc=self.dev_psc[s_][:n])
                out = optimise.caviar_batched(self.dev_stim[s_][:n], self.powers, mu0[:n], beta0[:n], shape0, rate0,
                                              phi0[:n], cov0[:n], seeds=list(seeds[lo:hi]), nnz_cap=self.nnz_cap,
                                              want_lam=False, lam_csr=True, workspace=self.workspace[s_],
                                              out=self.dev_out[s_], cta_variant=256 if D > 1 else 0, **src, **self.fit_options)
                self.workspace[s_] = out["_workspace"]
                self.dev_out[s_] = out
                ev_compute[s_] = cs.record_event()
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(ev_compute[s_])
                ho = self.host_out[s_]
                for k in self.OUT_KEYS:
                    ho[k][:n].copy_(out[k], non_blocking=True)
                ev_out[s_] = self.copy_out.record_event()
            pending.append((lo, hi, s_, ev_out[s_]))
            while pending and pending[0][3].query():
                drain(pending.pop(0))                           # deliver whatever has already arrived
        while pending:
            drain(pending.pop(0))
        for s in [self.copy_out] + self.compute:
            cur.wait_stream(s)
        return failed
