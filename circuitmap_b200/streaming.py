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
               workspaces=None, streams=None, **fit_options):
    """B independent fits from pinned host inputs to pinned host outputs.
       host_stim[b] (N, K), host_psc[b] (K, T): pinned tensors (sequences of length B; entries may repeat);
       dev_stim (B, N, K), dev_psc (B, K, T): device staging, reused across calls;
       priors: tuple (mu0, beta0, shape0, rate0, phi0, phi_cov0) of device tensors with leading dimension B (scalars
       for shape0 / rate0);  host_out: dict of pinned tensors for mu, beta, shape, rate, phi, phi_cov, z, lam
       (leading dimension B; 'lam' optional).  Returns the per-fit status (device tensor, all zero on success)."""
    import torch
    dev = dev_stim.device
    st = streams or _Streams(dev)
    B = dev_stim.shape[0]
    cur = torch.cuda.current_stream(dev)
    for s in (st.copy_in, st.compute, st.copy_out):
        s.wait_stream(cur)
    want_lam = "lam" in host_out
    mu0, beta0, shape0, rate0, phi0, cov0 = priors
    workspaces = workspaces if workspaces is not None else {}
    status, keep = [], []
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
        with torch.cuda.stream(st.copy_out):
            st.copy_out.wait_event(done)
            for k in _STATE_KEYS:
                if k not in host_out:
                    continue
                ring = host_out[k].shape[0]
                if ring >= B:
                    host_out[k][lo:hi].copy_(out[k], non_blocking=True)
                else:                                     # a ring of pinned slabs the caller drains (dense lam is 8 N K bytes per fit)
                    for b0 in range(0, hi - lo, ring):
                        n = min(ring, hi - lo - b0)
                        host_out[k][:n].copy_(out[k][b0:b0 + n], non_blocking=True)
    st.copy_out.synchronize()
    st.compute.synchronize()
    return torch.cat(status)
