// The persistent CAVIaR fit kernel and its device functions, compiled twice by caviar.cu:
//   CM_NT = 512 : one CTA (16 warps) per SM  -- lowest latency of a single fit
//   CM_NT = 256 : two CTAs (8 warps each) per SM -- the latency-bound phases of one fit (sequential sweep, block
//                 Cholesky, Newton) overlap with the other fit's work; used when the batch has >= 2 fits per SM
// Everything that depends on the CTA size lives here, inside namespace CM_FITNS.
namespace CM_FITNS {

constexpr int NT = CM_NT;          // threads of a fit CTA
constexpr int NW = NT / 32;
constexpr bool HELPERS = (NT == 512);   // helper CTAs exist for the 16-warp variant only: the 8-warp one compiles them away
constexpr int GCT = 16 * NW;       // panel GEMM column tile: 16 columns per warp
constexpr int GCT_LOG2 = (GCT == 256) ? 8 : 7;
constexpr int STAGE_GEMM_DOUBLES = GK * NB + GK * GCT;            // A chunk [GK][32] + X chunk [GK][GCT]
constexpr int GEMM_SMEM_DOUBLES = NST * STAGE_GEMM_DOUBLES;       // 144 KB (NT=512) / 80 KB (NT=256)
constexpr int SDXD_DOUBLES = NB * (NB + 1) + NB * XD_LD;          // diagonal block + its inverse, behind the ring
constexpr int FIT_SMEM_BYTES = (NT == 512) ? 200 * 1024 : (GEMM_SMEM_DOUBLES + SDXD_DOUBLES + 96) * 8;
constexpr int RC = (NT == 512) ? 512 : 256;    // staged row capacity (entries) of the chain warp's prefetch buffers

// X = L^-1 (lower) with X^T mirrored in the upper half is stored in GCT-column tiles, each tile row-major with 256
// doubles per row, so that a GK x 256 chunk is ONE contiguous block (one bulk copy).  Odd rows have bit 3 of the
// in-tile column flipped: with this swizzle the DMMA B-fragment loads from the dense shared-memory copy hit every
// 8-byte bank exactly twice (the minimum for 32 lanes).
__device__ __forceinline__ size_t xidx(int kk, int cc, int ldr) {
    return ((size_t)(cc >> GCT_LOG2) * ldr + kk) * GCT + ((cc & (GCT - 1)) ^ ((kk & 1) << 3));
}
struct Ctx {
    long long* tlast;   // shared: last phase timestamp (block 0 / thread 0 only)
    int N, K, P, nnz, it;
    double *X, *XI, *Dinv, *PA, *PB, *PP, *lam, *cst, *y, *ss, *pred, *resid, *z, *mu, *beta, *bvec, *dvec, *wvec, *slam, *slam2, *sp,
        *phibar, *phi, *phicov, *phiz, *phicovz, *lamhist, *lamT, *growbuf, *rcnt, *mce;
    double2* cscq;    // per CSC entry: (active index of the row as int bits, lam) -- rebuilt every a2
    int *row_ptr, *col_ptr, *col_k, *csc_row, *csc_pos, *cntp, *n0p, *n1p, *act, *ainv, *order, *order2, *pos, *rownz,
        *phizok, *dcnt, *dlist, *colpw, *nmask;
    // The by-trial (CSC) index IN USE: the full one built by the prologue, or -- once most rows have been pruned -- its
    // restriction to the rows that still carry a non-zero posterior (compact_csc).  lamT / cscq are indexed by it.
    int *ucol_ptr, *ucsc_row, *ucsc_pos, *ccol_ptr, *ccsc_row, *ccsc_pos, *member;
    int2* rowcb;      // per CSR entry: (begin, length) of its trial's list in the by-trial index in use (Gram expansion)
    int unnz;
    int4* chinfo;
    uint32_t *sortkeys, *keys;
    unsigned char *pw, *mask, *blocked;
    const double *mu0, *beta0, *phi0, *phicov0;
    double* red;      // shared: NW doubles scratch
    double* sm;       // shared: dynamic region
    int smd;          // its size in doubles
    int ct, role;     // CTAs working on this fit (1 = none but this one), role of this CTA (0 = the fit itself, > 0 = helper)
    int* job;         // global: job board between the fit CTA and its helpers (see panel_gemm_dist)
    int* status;      // global: status word of this fit (wait_helpers reports a helper that never answered)
};

// phase accounting for cm_caviar_debug_phase_cycles (block 0, thread 0; enabled on request only)
__device__ __forceinline__ void phase_mark(const Ctx& c, int id) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && (g_phase_enable & 1)) {   // one thread reads the switch, not every thread of every CTA
        const long long t = clock64();
        g_phase_cycles[id] += t - *c.tlast;
        *c.tlast = t;
    }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) r += red[i];
    return r;
}
__device__ __forceinline__ int block_sum_int(int v, double* red) {
    v = warp_sum(v);
    int* ri = reinterpret_cast<int*>(red);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) ri[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) r += ri[i];
    return r;
}

// ordered compaction of {i < n : flag(i)} into out[]; returns the count (block-wide, deterministic)
template <typename F>
__device__ int block_compact(int n, F flag, int* out, int* inv, double* red) {
    int* wcnt = reinterpret_cast<int*>(red);     // NW ints
    __shared__ int base_s;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i0 = 0; i0 < n; i0 += NT) {
        const int i = i0 + threadIdx.x;
        const bool f = i < n && flag(i);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wcnt[wid] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; ++w) off += wcnt[w];
        if (f) {
            const int idx = off + __popc(bal & ((1u << lane) - 1u));
            out[idx] = i;
            if (inv) inv[i] = idx;
        } else if (i < n && inv) {
            inv[i] = -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < NW; ++w) t += wcnt[w];
            base_s += t;
        }
        __syncthreads();
    }
    return base_s;
}

// fresh prediction pred[k] = sum_n mu[n] lam[n,k] over the (neuron-sorted) column lists
// sum_i mu[row_i] lam_i over one trial's list, entries added in list order; four entries' loads are issued together
// (the loop is bound by the latency of its dependent loads -- row index, then mu -- not by arithmetic)
__device__ __forceinline__ double trial_pred(const Ctx& c, int beg, int end) {
    double s = 0.0;
    for (int i = beg; i < end; i += 4) {
        int r[4];
        double l[4], m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = i + u < end;
            r[u] = ok ? c.ucsc_row[i + u] : -1;
            l[u] = ok ? c.lamT[i + u] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) m[u] = r[u] >= 0 ? c.mu[r[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (r[u] >= 0) s += m[u] * l[u];
    }
    return s;
}
__device__ __forceinline__ void compute_pred(const Ctx& c, double* dst) {
    for (int k = threadIdx.x; k < c.K; k += NT) dst[k] = trial_pred(c, c.ucol_ptr[k], c.ucol_ptr[k + 1]);
}

// D = sum lam (1 - lam) and b = sigma sum lam y + mu0 / beta0^2 of active row ia (caviar.py:167-171), one warp.
// (Four groups of 32 entries per round, as in row_cst below, was measured here and is slower.)
__device__ __forceinline__ void row_dvec_bvec(const Ctx& c, int ia, double sigma) {
    const int lane = threadIdx.x & 31;
    const int n = c.act[ia];
    double d = 0.0, by = 0.0;
    for (int j = c.row_ptr[n] + lane; j < c.row_ptr[n + 1]; j += 32) {
        const double l = c.lam[j];
        d += l * (1.0 - l);
        by += l * c.y[c.col_k[j]];
    }
    d = warp_sum(d); by = warp_sum(by);
    if (lane == 0) {
        const double b0 = c.beta0[n];
        c.dvec[ia] = d;
        c.bvec[ia] = sigma * by + c.mu0[n] / (b0 * b0);
    }
}
// One CSR row by one warp, four groups of 32 entries per round so that the dependent gathers (trial index, then y) of
// a whole ~100-entry row are two round trips instead of eight:
// per-entry constant part of the sigmoid argument of row n (caviar.py:216-218 with the Monte-Carlo term of mc_means)
__device__ __forceinline__ void row_cst(const Ctx& c, int n, double sigma) {
    const int lane = threadIdx.x & 31;
    const double mu_n = c.mu[n], be = c.beta[n];
    const double cterm = 0.5 * sigma * (mu_n * mu_n + be * be);
    const double* mce = c.mce + n * PMAX;                    // Monte-Carlo term per power (mc_means)
    const int beg = c.row_ptr[n], end = c.row_ptr[n + 1];
    for (int j0 = beg + lane; j0 < end; j0 += 128) {
        double l[4], yv[4], mc[4];
        int k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u;
            const bool ok = j < end;
            k[u] = ok ? c.col_k[j] : -1;
            l[u] = ok ? c.lam[j] : 0.0;
            mc[u] = ok ? mce[c.pw[j]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) yv[u] = k[u] >= 0 ? c.y[k[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (k[u] >= 0) c.cst[j0 + 32 * u] = (mc[u] - cterm) + sigma * mu_n * yv[u] + sigma * mu_n * mu_n * l[u];
    }
}

// ------------------------------------------------------------------------------------------------ a2
// block_update_mu (caviar.py:166-172) on the active set A = {n : lam[n,:] != 0}:
// M = sigma (diag(sum lam(1-lam)) + lam_A lam_A^T) + diag(1/beta0^2);  C = M^-1;  mu = C b;  beta = diag C.
// C = X^T X with X = L^-1 built by a bordered (block-row) recursion that keeps X (lower) and X^T (upper) in one
// na x na array, so both panel GEMMs read it with unit stride across threads.
// ---- async copy helpers (LDGSTS) ----
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned sdst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned sdst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N_) : "memory"); }

// ---- mbarrier / bulk-copy helpers (UBLKCP + SYNCS in SASS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Gram rows of one 32-row block of M into PA (lower part incl. the diagonal block).
// One warp per row; 32 entries of the row are expanded at once (each lane walks the column list of its own
// trial, 8 list entries prefetched per round), lanes that hit the same target column in the same step are
// combined in lane order -> deterministic.
constexpr int GCH = 8;      // list entries per lane loaded in one round (10 -- a whole list of the 10-target designs -- was measured: slower)
__device__ void gram_rows(const Ctx& c, int i0, int nb, double sigma, int part = 0, int nparts = 1) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cap = GEMM_SMEM_DOUBLES / NW;       // row-buffer doubles per warp
    const int tagcap = ((c.smd - GEMM_SMEM_DOUBLES) * 8) / NW;   // conflict-tag bytes per warp (behind the GEMM ring)
    const double2* __restrict__ cscq = c.cscq;
    for (int r = part + nparts * wid; r < nb; r += nparts * NW) {     // row r of the block: CTA r % nparts, warp r / nparts
        const int ia = i0 + r;
        const int n = c.act[ia];
        double* acc = (ia + 1 <= cap) ? (c.sm + (size_t)wid * cap) : (c.growbuf + (size_t)r * (c.N + 2));
        unsigned char* tags = reinterpret_cast<unsigned char*>(c.sm + GEMM_SMEM_DOUBLES) + (size_t)wid * tagcap;
        const bool use_tags = ia + 1 <= tagcap;
        for (int q = lane; q <= ia; q += 32) acc[q] = 0.0;
        __syncwarp();
        const int beg = c.row_ptr[n], end = c.row_ptr[n + 1];
        const int2* __restrict__ rowcb = c.rowcb;
        double la_n = 0.0;                                   // the next 32 entries are fetched while these are expanded
        int2 rc_n = make_int2(0, 0);
        if (beg + lane < end) { la_n = c.lam[beg + lane]; rc_n = rowcb[beg + lane]; }
        for (int jb = beg; jb < end; jb += 32) {
            const double la = la_n;
            const int cb = rc_n.x, len = (la != 0.0) ? rc_n.y : 0;
            la_n = 0.0; rc_n = make_int2(0, 0);
            if (jb + 32 + lane < end) { la_n = c.lam[jb + 32 + lane]; rc_n = rowcb[jb + 32 + lane]; }
            int maxlen = len;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
            for (int t0 = 0; t0 < maxlen; t0 += GCH) {
                int ibs[GCH];
                double vs[GCH];
#pragma unroll
                for (int u = 0; u < GCH; ++u) {                  // independent 16-byte loads (memory-level parallelism)
                    const int t = t0 + u;
                    int ib = -1;
                    double lv = 0.0;
                    if (t < len) {
                        const double2 rec = cscq[cb + t];
                        ib = (int)__double_as_longlong(rec.x);
                        lv = rec.y;
                    }
                    if (ib > ia) ib = -1;
                    ibs[u] = ib;
                    vs[u] = la * lv;
                }
#pragma unroll
                for (int u = 0; u < GCH; ++u) {
                    if (t0 + u >= maxlen) break;               // warp-uniform
                    const int ib = ibs[u];
                    const double v = vs[u];
                    // conflict check: every lane tags its target; a lane that reads back another id shares the target
                    bool lost = false;
                    if (use_tags) {
                        if (ib >= 0) tags[ib] = (unsigned char)lane;
                        __syncwarp();
                        lost = (ib >= 0) && (tags[ib] != (unsigned char)lane);
                    }
                    if (use_tags && !__any_sync(0xffffffffu, lost)) {
                        if (ib >= 0) acc[ib] += v;             // all targets distinct: plain read-modify-write
                    } else {
                        // same target hit by several lanes: combine them in lane order (deterministic)
                        const unsigned amask = __ballot_sync(0xffffffffu, ib >= 0);
                        if (ib >= 0) {
                            const unsigned grp = __match_any_sync(amask, ib);
                            const int leader = __ffs(grp) - 1;
                            unsigned rest = grp & ~(1u << leader);
                            double ssum = __shfl_sync(amask, v, leader);
                            while (__any_sync(amask, rest != 0)) {
                                const int src = rest ? (__ffs(rest) - 1) : lane;
                                const double ov = __shfl_sync(amask, v, src);
                                if (rest) { ssum += ov; rest &= rest - 1; }
                            }
                            if (lane == leader) acc[ib] += ssum;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        const double b0 = c.beta0[n];
        const double dd = c.dvec[ia];
        for (int q = lane; q <= ia; q += 32) {
            double v = acc[q];
            if (q == ia) v = sigma * (dd + v) + 1.0 / (b0 * b0);
            else v = sigma * v;
            c.PA[pidx(r, q)] = v;
        }
        __syncwarp();
    }
}

// D(8x8) += A(8x4, row) * B(4x8, col): the fp64 tensor-core path (SASS DMMA.8x8x4); operands in registers.
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Pipeline state of the panel GEMMs: NST-stage ring of (A chunk, X chunk) filled by bulk copies, one "full" and one
// "empty" mbarrier per stage.  `seq` counts chunks since kernel start (stage = seq % NST, parity = seq / NST & 1).
struct GemmPipe {
    uint64_t* full;      // [NST]
    uint64_t* empty;     // [NST]
    unsigned seq;
};

// OUT[r][cc] = sum_kk IN[r][kk] * X[kk][cc] ; UPPER: kk <= cc (the X^T half), else kk >= cc (the X half).
// 32 x GK chunks of IN and GK x 256 chunks of X are streamed through shared memory by cp.async.bulk (issued by warp
// 0, completion on mbarriers, 4 stages in flight); each warp owns the 32 x 16 slice of the 32 x 256 output tile as
// 4 x 2 DMMA tiles.  Row strides = 4 (mod 16) doubles keep the fragment loads bank-conflict free.
// Split-k: the k range of every column tile is cut into `nseg` segments (a function of i0 only, see kseg_for); a job is
// (tile, segment), jobs part, part + nparts, ... are taken by this CTA; with nseg > 1 a job writes its partial panel to
// PP[segment] and the caller adds the segments in fixed order (reduce_partials) -- the result does not depend on how many
// CTAs shared the jobs.
__device__ __forceinline__ int kseg_for(int i0) {
    if (!HELPERS) return 1;                        // the 8-warp (batched) variant has no helper CTAs to share segments with
    return i0 >= 768 ? 4 : (i0 >= 576 ? 3 : (i0 >= 384 ? 2 : 1));
}
template <bool UPPER>
__device__ void panel_gemm(const Ctx& c, int ldr, int i0, int nb, const double* IN, double* OUT, GemmPipe& gp,
                           int part = 0, int nparts = 1, int nseg = 1) {
    const int dbg = g_phase_enable;                // diagnostics switches, read once (not once per chunk in the loops below)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int lq = lane >> 2, lr = lane & 3;
    const int sw = (lr & 1) << 3;                 // swizzle term of this lane's k rows (k = 4*ks + lr)
    const int wc = wid * 16;
    double* stage0 = c.sm;
    const double* Xg = c.X;
    // the staging ring was last written through the generic proxy (row buffers, pred): order those writes before
    // the async-proxy bulk copies that reuse the same shared memory
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    const int ntiles = (i0 + GCT - 1) / GCT;
    for (int job = part; job < ntiles * nseg; job += nparts) {       // jobs (column tile, k segment) part, part + nparts, ...
        const int tile = job / nseg, seg = job - tile * nseg;
        const int ct0 = tile * GCT;
        const int tile_end = min(i0, ct0 + GCT);
        const int klast = UPPER ? tile_end : i0;
        const int nc_all = (klast - (UPPER ? 0 : ct0) + GK - 1) / GK;
        const int c_lo = (nc_all * seg) / nseg, c_hi = (nc_all * (seg + 1)) / nseg;
        const int kfirst = (UPPER ? 0 : ct0) + c_lo * GK;             // first k row of this segment
        const int nc = c_hi - c_lo;
        double* OUTj = nseg > 1 ? c.PP + (size_t)seg * NB * (c.N + ROWPAD) : OUT;
        const unsigned seq0 = gp.seq;
        const double* Xtile = Xg + (size_t)(ct0 >> GCT_LOG2) * ldr * GCT;
        auto fill = [&](int ci) {                             // warp 0 only: two bulk copies per chunk
            const unsigned g = seq0 + ci;
            const int st = g % NST;
            if (g >= NST) mbar_wait(&gp.empty[st], ((g / NST) - 1) & 1);
            if (lane == 0) {
                double* As = stage0 + (size_t)st * STAGE_GEMM_DOUBLES;
                double* Xs = As + GK * NB;
                const int kk0 = kfirst + ci * GK;
                if (dbg & 4) mbar_arrive(&gp.full[st]);                     // debug: no copies
                else {
                    mbar_expect_tx(&gp.full[st], (uint32_t)(GK * NB * 8 + GK * GCT * 8));
                    bulk_g2s(As, IN + (size_t)kk0 * NB, GK * NB * 8, &gp.full[st]);
                    bulk_g2s(Xs, Xtile + (size_t)kk0 * GCT, GK * GCT * 8, &gp.full[st]);
                }
            }
            __syncwarp();
        };
        if (wid == 0)
            for (int ci = 0; ci < min(nc, NST); ++ci) fill(ci);
        double acc[4][2][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        const int cmin = ct0 + wc, cmax = cmin + 15;
        for (int ci = 0; ci < nc; ++ci) {
            const unsigned g = seq0 + ci;
            const int st = g % NST;
            mbar_wait(&gp.full[st], (g / NST) & 1);
            const int kk0 = kfirst + ci * GK;
            if (cmin < i0 && !(dbg & 2)) {                                  // debug bit 2: no math
                const double* ab = stage0 + (size_t)st * STAGE_GEMM_DOUBLES;
                const double* xb = ab + GK * NB;
                // interior chunk of the triangle: every (kk, cc) pair of this warp is valid -> no masking at all
                const bool interior = (nb == NB) && (cmax < i0) && (kk0 + GK <= i0) &&
                                      (UPPER ? (kk0 + GK - 1 <= cmin) : (kk0 >= cmax));
                if (interior) {
                    double af[GK / 4][4], bf[GK / 4][2];
#pragma unroll
                    for (int ks = 0; ks < GK / 4; ++ks) {
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) af[ks][mt] = ab[(4 * ks + lr) * NB + ((8 * mt + lq) ^ sw)];
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) bf[ks][nt] = xb[(4 * ks + lr) * GCT + ((wc + 8 * nt + lq) ^ sw)];
                    }
#pragma unroll
                    for (int ks = 0; ks < GK / 4; ++ks)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int mt = 0; mt < 4; ++mt) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[ks][mt], bf[ks][nt]);
                } else {
#pragma unroll
                    for (int ks = 0; ks < GK / 4; ++ks) {
                        const int kbase = kk0 + 4 * ks;
                        const bool skip = UPPER ? (kbase > cmax) : (kbase + 3 < cmin);
                        if (skip || kbase >= i0) continue;                   // warp-uniform
                        const int kk = kbase + lr;
                        double af[4];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            const double av = ab[(4 * ks + lr) * NB + ((8 * mt + lq) ^ sw)];
                            af[mt] = (8 * mt + lq < nb && kk < i0) ? av : 0.0;
                        }
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int cl = wc + 8 * nt + lq;
                            const int cc = ct0 + cl;
                            const bool in = (kk < i0) && (cc < i0) && (UPPER ? (kk <= cc) : (kk >= cc));
                            const double xv = xb[(4 * ks + lr) * GCT + (cl ^ sw)];
                            const double bv = in ? xv : 0.0;
#pragma unroll
                            for (int mt = 0; mt < 4; ++mt) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bv);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&gp.empty[st]);
            if (wid == 0 && ci + NST < nc) fill(ci + NST);
        }
        gp.seq = seq0 + nc;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int r = 8 * mt + lq;
            if (r < nb) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int cc = cmin + 8 * nt + 2 * lr;
                    if (cc < i0) OUTj[pidx(r, cc)] = acc[mt][nt][0];
                    if (cc + 1 < i0) OUTj[pidx(r, cc + 1)] = acc[mt][nt][1];
                }
            }
        }
    }
    __syncthreads();
}

// ---- helper CTAs: the column tiles of a panel GEMM are independent, so for a single large fit the launch adds
// `ct - 1` helper CTAs (on otherwise idle SMs) that take tiles part, part + ct, ... of every panel GEMM with more than
// one tile.  All operands (IN panel, X, OUT panel) live in global memory; every output element is still computed by
// one warp in the same order, so the result is bitwise identical to the single-CTA run.
// Job board (ints, global, zeroed by the host): [0] sequence number (release/acquire), [1] type (0 quit, 1 upper,
// 2 lower), [2] i0, [3] nb, [16] number of helper completions, [20] watchdog flag.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// A helper only ever leaves on a quit job (type 0): it may idle for as long as the fit CTA runs its serial phases.
// The fit CTA bounds its own wait for the helpers of ONE job (they are co-resident by construction -- cooperative
// launch -- so this only fires on a genuine fault); it then flags the fit at once (status CM_EHELPER) and stops
// waiting for helpers for the rest of the fit, so a broken launch ends quickly and loudly instead of hanging.
constexpr long long WATCHDOG_CYCLES = 1ll << 36;      // ~35 s at 1.9 GHz for a single job
constexpr int CM_EHELPER = 9;

__device__ __noinline__ void newton_rows(const Ctx& c, const double* powers, const int* list, int nlist, int part, int nparts);
// Monte-Carlo term of update_lam (caviar.py:209-215,233-235) per (neuron, power): mcE = mean_s log(f_s / (1 - f_s)) with
// f_s = sigmoid(phi0_s I - phi1_s) over the S truncated-normal samples.  log(f / (1 - f)) IS its argument, so the mean
// collapses to mean(phi0) I - mean(phi1) (App. A.2) -- as long as no sample saturates the float64 sigmoid: 1 - f loses its
// bits as x grows (the reference's own value is rounding noise of 2^-53 e^x per sample: 2e-4 at x = 28, 1 at x = 36) and
// from x ~ 36.7 on it is +inf (f rounds to 1).  So the linear form is used when the LARGEST possible argument
// max(phi0) I - min(phi1) is <= 28 (every reference experiment: powers <= 70 give <= ~23 even with the prior's wide phi),
// and the reference's own expression is evaluated sample by sample otherwise (large powers / wide phi
// posteriors).  One warp per neuron; even lanes carry phi0 samples, odd lanes phi1.
constexpr double MC_LITERAL_X = 28.0;
__device__ void mc_means(const Ctx& c, const uint32_t* keys_cur, int S, const double* powers, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int n = part * NW + wid; n < c.N; n += nparts * NW) {
        if (c.dcnt[n]) continue;
        const int m = c.pos[n];
        const uint32_t k0 = keys_cur[2 * m], k1 = keys_cur[2 * m + 1];
        const int cc = lane & 1;                                   // flat index e = 2 s + component
        const double mean = c.phi[2 * n + cc];
        const double sd = c.phicov[4 * n + 3 * cc];                // diag(phi_cov): a variance used as sd
        const double cdf0 = normcdf(-mean / sd);
        double acc = 0.0, ext = -CUDART_INF;                       // ext: max phi0 (even lanes) / max of -phi1 (odd lanes)
        for (int e = lane; e < 2 * S; e += 32) {
            uint32_t x0 = (uint32_t)e, x1 = (uint32_t)(2 * S + e);
            threefry2x32(k0, k1, x0, x1);
            const double u = bits_to_unit_double(x0, x1);
            const double smp = normcdfinv(cdf0 + u * (1.0 - cdf0)) * sd + mean;
            acc += smp;
            ext = fmax(ext, cc ? -smp : smp);
        }
#pragma unroll
        for (int off = 16; off > 1; off >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, off);
            ext = fmax(ext, __shfl_xor_sync(0xffffffffu, ext, off));
        }
        if (lane < 2) c.phibar[2 * n + lane] = acc / (double)S;
        const double pb0 = __shfl_sync(0xffffffffu, acc, 0) / (double)S, pb1 = __shfl_sync(0xffffffffu, acc, 1) / (double)S;
        const double p0max = __shfl_sync(0xffffffffu, ext, 0), p1min = -__shfl_sync(0xffffffffu, ext, 1);
        bool lit = false;
        double pw = 0.0;
        if (lane < c.P) {
            pw = powers[lane];
            lit = !(p0max * pw - p1min <= MC_LITERAL_X);           // also taken for NaN / inf samples, as the reference would propagate them
            c.mce[n * PMAX + lane] = pb0 * pw - pb1;
        }
        unsigned need = __ballot_sync(0xffffffffu, lit);
        if (need) {                                                // rare: the reference's expression, sample by sample
            double lacc[PMAX];
#pragma unroll
            for (int p = 0; p < PMAX; ++p) lacc[p] = 0.0;
            for (int e0 = 0; e0 < 2 * S; e0 += 32) {               // uniform trip count: the shuffle below needs the whole warp
                const int e = e0 + lane;
                double smp = 0.0;
                if (e < 2 * S) {
                    uint32_t x0 = (uint32_t)e, x1 = (uint32_t)(2 * S + e);
                    threefry2x32(k0, k1, x0, x1);
                    const double u = bits_to_unit_double(x0, x1);
                    smp = normcdfinv(cdf0 + u * (1.0 - cdf0)) * sd + mean;
                }
                const double other = __shfl_xor_sync(0xffffffffu, smp, 1);     // even lanes: phi1 of the same sample
                if (cc == 0 && e < 2 * S) {
#pragma unroll
                    for (int p = 0; p < PMAX; ++p)
                        if (need & (1u << p)) {
                            const double f = sigmoid_d(smp * powers[p] - other);
                            lacc[p] += log(f / (1.0 - f));
                        }
                }
            }
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (need & (1u << p)) {
                    double v = lacc[p];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (lane == 0) c.mce[n * PMAX + p] = v / (double)S;
                }
        }
    }
}

// w = X b (rows) and mu = X^T w, beta = column sums of squares of X (columns): the tail of block_update_mu
__device__ void a2_wvec(const Ctx& c, int na, int ldr, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = part * NW + wid; i < na; i += nparts * NW) {
        double s = 0.0;
        for (int q = lane; q <= i; q += 32) s += c.X[xidx(i, q, ldr)] * c.bvec[q];
        s = warp_sum(s);
        if (lane == 0) c.wvec[i] = s;
    }
}
__device__ void a2_mubeta(const Ctx& c, int na, int ldr, int part, int nparts) {
    for (int cc = part * NT + threadIdx.x; cc < na; cc += nparts * NT) {
        double m = 0.0, v = 0.0;
        for (int i = cc; i < na; ++i) {
            const double x = c.X[xidx(i, cc, ldr)];
            m += x * c.wvec[i];
            v += x * x;
        }
        const int n = c.act[cc];
        c.mu[n] = m;
        c.beta[n] = v;
    }
}

// post a job for the helpers (all threads of the fit CTA call this; data written before the call is published)
__device__ void post_job(const Ctx& c, int type, int a, int b, int d) {
    __syncthreads();
    if (threadIdx.x == 0) {
        c.job[1] = type; c.job[2] = a; c.job[3] = b; c.job[4] = d;
        __threadfence();
        st_release_gpu(&c.job[0], c.job[0] + 1);
    }
}
// wait until every helper has finished the job posted last; their global writes are visible afterwards
__device__ void wait_helpers(const Ctx& c) {
    __syncthreads();
    if (threadIdx.x == 0 && !c.job[20]) {
        const long long t0 = clock64();
        const int want = c.job[0] * (c.ct - 1);
        while (ld_acquire_gpu(&c.job[16]) < want)
            if (clock64() - t0 > WATCHDOG_CYCLES) { c.job[20] = 1; atomicExch(c.status, CM_EHELPER); break; }
    }
    __syncthreads();
    __threadfence();
}

// OUT = sum of the nseg partial panels, segments added in order; columns < i0, all 32 rows of the panel layout
__device__ void reduce_partials(const Ctx& c, int i0, int nseg, double* OUT) {
    const size_t stride = (size_t)NB * (c.N + ROWPAD);
    const int total = i0 * NB;
    for (int e = threadIdx.x; e < total; e += NT) {
        double v = c.PP[e];
        for (int s = 1; s < nseg; ++s) v += c.PP[(size_t)s * stride + e];
        OUT[e] = v;
    }
    __syncthreads();
}
template <bool UPPER>
__device__ void panel_gemm_dist(const Ctx& c, int ldr, int i0, int nb, const double* IN, double* OUT, GemmPipe& gp) {
    const int nseg = kseg_for(i0);
    const int njobs = ((i0 + GCT - 1) / GCT) * nseg;
    const bool dist = HELPERS && c.ct > 1 && njobs > 1;
    if (dist) post_job(c, UPPER ? 1 : 2, i0, nb, nseg);   // IN and X are complete (written by this CTA)
    panel_gemm<UPPER>(c, ldr, i0, nb, IN, OUT, gp, 0, dist ? c.ct : 1, nseg);
    if (dist) wait_helpers(c);
    if (nseg > 1) reduce_partials(c, i0, nseg, OUT);
}

// Per CSR entry (n, k): (begin, length) of the part of trial k's list in the by-trial index in use that the Gram
// expansion of row n needs -- saves it one dependent memory round trip per 32 entries (col_k -> col_ptr).  The lists are
// sorted by neuron and the active order is the neuron order, so the entries with an active index <= that of row n
// (the lower triangle) are exactly the list's prefix up to the row's own entry: the length is the own position + 1 and
// the upper-triangle half of every list is never loaded.  Rebuilt whenever the index in use changes; entries of rows
// outside it keep stale values and are never read (only active rows are expanded).
__device__ __forceinline__ void build_rowcb(const Ctx& c) {
#pragma unroll 4
    for (int i = threadIdx.x; i < c.unnz; i += NT) {
        const int j = c.ucsc_pos[i];
        const int cb = c.ucol_ptr[c.col_k[j]];
        c.rowcb[j] = make_int2(cb, i - cb + 1);
    }
}

// Restriction of the by-trial index to the rows with a non-zero posterior (see the call site).  Three block-wide passes
// over the FULL index: per-trial counts, exclusive scan over the K trials, ordered fill (neuron order within a trial is
// preserved, so every sum keeps its order).
__device__ __noinline__ void compact_csc(Ctx& c, double* red) {
    const int K = c.K, N = c.N;
    __shared__ int s_carry;
    int* wsum = reinterpret_cast<int*>(red);                 // NW ints
    for (int n = threadIdx.x; n < N; n += NT) c.member[n] = c.rownz[n] > 0 ? 1 : 0;
    if (threadIdx.x == 0) { s_carry = 0; c.ccol_ptr[0] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k0 = 0; k0 < K; k0 += NT) {
        const int k = k0 + threadIdx.x;
        int cnt = 0;
        if (k < K)
            for (int i = c.col_ptr[k]; i < c.col_ptr[k + 1]; ++i) cnt += c.member[c.csc_row[i]];
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        int off = s_carry;
        for (int w = 0; w < wid; ++w) off += wsum[w];
        const int start = off + inc - cnt;                   // exclusive prefix of this trial
        if (k < K) {
            c.ccol_ptr[k + 1] = off + inc;
            int o2 = start;
            for (int i = c.col_ptr[k]; i < c.col_ptr[k + 1]; ++i) {
                const int r = c.csc_row[i];
                if (c.member[r]) { c.ccsc_row[o2] = r; c.ccsc_pos[o2] = c.csc_pos[i]; ++o2; }
            }
        }
        __syncthreads();
        if (threadIdx.x == NT - 1) s_carry = off + inc;
        __syncthreads();
    }
    c.ucol_ptr = c.ccol_ptr; c.ucsc_row = c.ccsc_row; c.ucsc_pos = c.ccsc_pos;
    c.unnz = s_carry;
    __syncthreads();
    build_rowcb(c);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ a2, small systems
// block_update_mu (caviar.py:166-172) when the active set is small enough for the whole system to live in shared memory
// (packed lower triangle: na (na + 1) / 2 doubles -- na <= 222 in the 16-warp variant, 155 in the 8-warp one).  That is
// the normal state of a fit after its first gated sweep (C3: ~250 -> 100 active rows for 47 of 50 iterations), where the
// bordered 32-row recursion over global memory spends its time in per-block pipeline fill, barriers and 32 x 32 factor
// steps rather than in arithmetic.  Here: Gram rows straight into the packed matrix (one warp per row, same deterministic
// expansion as gram_rows), right-looking Cholesky in place, X = L^-1 in place row by row, then w = X b, mu = X^T w,
// beta = column sums of squares of X.  No global-memory matrix traffic, no helper jobs.
__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }
__device__ __forceinline__ int a2_small_capacity(int smd) {
    int na = 0;
    while (tri(na + 1) + (na + 1) + ((NW * (na + 1) + 7) >> 3) + 64 <= smd) ++na;
    return na;
}
__device__ __noinline__ void a2_small(const Ctx& c, double sigma, int na) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* Mp = c.sm;                                       // packed lower triangle, row-major: (r, q) at tri(r) + q
    double* vec = Mp + tri(na);                              // na doubles: current column / row
    unsigned char* tagbase = reinterpret_cast<unsigned char*>(vec + na);
    const double2* __restrict__ cscq = c.cscq;
    // ---- Gram rows (cf. gram_rows) ----
    for (int ia = wid; ia < na; ia += NW) {
        const int n = c.act[ia];
        double* acc = Mp + tri(ia);
        unsigned char* tags = tagbase + (size_t)wid * na;
        for (int q = lane; q <= ia; q += 32) acc[q] = 0.0;
        __syncwarp();
        const int beg = c.row_ptr[n], end = c.row_ptr[n + 1];
        for (int jb = beg; jb < end; jb += 32) {
            const int j = jb + lane;
            double la = 0.0;
            int cb = 0, len = 0;
            if (j < end) {
                la = c.lam[j];
                if (la != 0.0) {
                    const int k = c.col_k[j];
                    cb = c.ucol_ptr[k];
                    len = c.ucol_ptr[k + 1] - cb;
                }
            }
            int maxlen = len;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
            for (int t = 0; t < maxlen; ++t) {
                int ib = -1;
                double v = 0.0;
                if (t < len) {
                    const double2 rec = cscq[cb + t];
                    ib = (int)__double_as_longlong(rec.x);
                    v = la * rec.y;
                }
                if (ib > ia) ib = -1;
                if (ib >= 0) tags[ib] = (unsigned char)lane;
                __syncwarp();
                const bool lost = (ib >= 0) && (tags[ib] != (unsigned char)lane);
                if (!__any_sync(0xffffffffu, lost)) {
                    if (ib >= 0) acc[ib] += v;                 // all targets distinct
                } else {                                       // same target hit by several lanes: combine in lane order
                    const unsigned amask = __ballot_sync(0xffffffffu, ib >= 0);
                    if (ib >= 0) {
                        const unsigned grp = __match_any_sync(amask, ib);
                        const int leader = __ffs(grp) - 1;
                        unsigned rest = grp & ~(1u << leader);
                        double ssum = __shfl_sync(amask, v, leader);
                        while (__any_sync(amask, rest != 0)) {
                            const int src = rest ? (__ffs(rest) - 1) : lane;
                            const double ov = __shfl_sync(amask, v, src);
                            if (rest) { ssum += ov; rest &= rest - 1; }
                        }
                        if (lane == leader) acc[ib] += ssum;
                    }
                }
                __syncwarp();
            }
        }
        const double b0 = c.beta0[n];
        const double dd = c.dvec[ia];
        for (int q = lane; q <= ia; q += 32) {
            const double v = acc[q];
            acc[q] = (q == ia) ? sigma * (dd + v) + 1.0 / (b0 * b0) : sigma * v;
        }
    }
    __syncthreads();
    phase_mark(c, 1);
    // ---- Cholesky, right-looking, in place (two barriers per column) ----
    for (int j = 0; j < na; ++j) {
        __syncthreads();                                      // the trailing update of column j - 1 is complete
        const double djj = sqrt(Mp[tri(j) + j]);
        for (int r = j + threadIdx.x; r < na; r += NT) {
            if (r == j) vec[j] = djj;                         // (the pivot itself is overwritten after the next barrier)
            else {
                const double v = Mp[tri(r) + j] / djj;
                vec[r] = v;
                Mp[tri(r) + j] = v;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) Mp[tri(j) + j] = djj;
        // trailing update: row r by warp, columns j < q <= r by lanes
        for (int r = j + 1 + wid; r < na; r += NW) {
            const double lr = vec[r];
            double* row = Mp + tri(r);
            for (int q = j + 1 + lane; q <= r; q += 32) row[q] -= lr * vec[q];
        }
    }
    __syncthreads();
    phase_mark(c, 3);
    // ---- X = L^-1 in place, row by row: X[r][q] = -(sum_{t=q}^{r-1} L[r][t] X[t][q]) / L[r][r], X[r][r] = 1 / L[r][r] ----
    {
        int tpc = 1;                                         // threads per column (power of two, <= 4)
        while (tpc < 4 && 2 * tpc * na <= NT) tpc *= 2;
        const int col = threadIdx.x / tpc, part = threadIdx.x % tpc;
        for (int r = 0; r < na; ++r) {
            __syncthreads();
            for (int t = threadIdx.x; t <= r; t += NT) vec[t] = Mp[tri(r) + t];       // L row r
            __syncthreads();
            const double d = vec[r];
            double sacc = 0.0;
            if (col < r)
                for (int t = col + part; t < r; t += tpc) sacc += vec[t] * Mp[tri(t) + col];
            if (tpc >= 2) sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);              // every lane takes part
            if (tpc >= 4) sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            if (part == 0) {
                if (col < r) Mp[tri(r) + col] = -sacc / d;
                else if (col == r) Mp[tri(r) + r] = 1.0 / d;
            }
        }
    }
    __syncthreads();
    phase_mark(c, 5);
    // ---- w = X b ; mu = X^T w ; beta = column sums of squares ----
    for (int r = threadIdx.x; r < na; r += NT) {
        double sacc = 0.0;
        const double* row = Mp + tri(r);
        for (int q = 0; q <= r; ++q) sacc += row[q] * c.bvec[q];
        vec[r] = sacc;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < na; q += NT) {
        double m = 0.0, v = 0.0;
        for (int r = q; r < na; ++r) {
            const double x = Mp[tri(r) + q];
            m += x * vec[r];
            v += x * x;
        }
        const int n = c.act[q];
        c.mu[n] = m;
        c.beta[n] = v;
    }
    __syncthreads();
    phase_mark(c, 6);
}

// ------------------------------------------------------------------------------------------------ a2, tile solve
// block_update_mu for LARGE active sets in the 16-warp variant (single large fits and their helper CTAs).  The bordered
// 32-row recursion above is a chain of ~6 dependent phases per 32 rows (Gram job, two panel-GEMM jobs, CTA-wide 32 x 32
// factor, X update), 20-45 us each whatever the number of CTAs -- a C3 fit spends ~50 ms in ~200 such block steps.  The
// tile solve keeps the same mathematics (M = L L^T, X = L^-1, mu = X^T X b, beta = colsumsq X) in a form whose parallel
// width grows with the matrix:
//   * M (lower triangle) is a plain row-major square of 32 x 32 tiles in global memory (L2-resident), padded to a multiple
//     of 32 with an identity block; all Gram rows are one job;
//   * right-looking Cholesky over tile columns kb: the fit CTA factors the diagonal tile (and inverts the factor), then ONE
//     job scales the panel below it and ONE job applies the rank-32 update to all (nt-kb)(nt-kb-1)/2 trailing tiles, a warp
//     per tile (DMMA, operands straight from L2);
//   * X = L^-1 by tile columns: a column is a task of one CTA (its 16 warps split the k range of every tile and add
//     their partial tiles in fixed order through shared memory), columns are independent;
//   * every tile is produced by one warp (or one CTA) with a fixed k order, so the result does not depend on how many
//     CTAs share the jobs (bitwise neutral in the helper count, like the rest).
struct Tiles { double* A; double* XI; double* Dinv; int lda, nt, na; };
__device__ __forceinline__ Tiles make_tiles(const Ctx& c, int na) {
    Tiles t;
    t.na = na; t.nt = (na + 31) >> 5; t.lda = t.nt * 32; t.A = c.X; t.XI = c.XI; t.Dinv = c.Dinv;
    return t;
}

// one warp: acc(32 x 32) += A(32 x 32, row-major) * B^T  [NT: B row-major (n, k)]  or  A * B  [NN: B row-major (k, n)]
// accumulator layout: acc[mt][nt][e] = C[8 mt + lane / 4][8 nt + 2 (lane % 4) + e]
template <bool NN>
__device__ __forceinline__ void tile_mma(double (&acc)[4][4][2], const double* __restrict__ A, int lda,
                                         const double* __restrict__ B, int ldb) {
    const int lane = threadIdx.x & 31, lq = lane >> 2, lr = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        double a[4], b[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) a[mt] = A[(size_t)(8 * mt + lq) * lda + 4 * ks + lr];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[nt] = NN ? B[(size_t)(4 * ks + lr) * ldb + 8 * nt + lq] : B[(size_t)(8 * nt + lq) * ldb + 4 * ks + lr];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
}
__device__ __forceinline__ void tile_zero(double (&acc)[4][4][2]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
}
// C = sgn * acc (+ C if ADD)
template <bool ADD>
__device__ __forceinline__ void tile_store(const double (&acc)[4][4][2], double* C, int ldc, double sgn) {
    const int lane = threadIdx.x & 31, lq = lane >> 2, lr = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            double2* p = reinterpret_cast<double2*>(C + (size_t)(8 * mt + lq) * ldc + 8 * nt + 2 * lr);
            double2 v = ADD ? *p : make_double2(0.0, 0.0);
            v.x += sgn * acc[mt][nt][0]; v.y += sgn * acc[mt][nt][1];
            *p = v;
        }
}

// Strip form of the same products for jobs with FEW tiles: four consecutive warps share one output tile, warp `strip`
// computes its rows 8 strip .. 8 strip + 7.  All operands of a 32 x 32 x 32 product (8 + 32 doubles per lane) are loaded
// before the first DMMA, so a product costs one L2 round trip instead of eight (the full-tile form keeps 32 accumulators
// per lane and can only prefetch one k step ahead: ~3.3 us per product, latency-bound), and a tile is finished four
// times sooner.  Every output element accumulates its k steps in the same order as in the full-tile form: bitwise equal.
template <bool NN>
__device__ __forceinline__ void tile_mma_strip(double (&acc)[4][2], const double* __restrict__ A, int lda,
                                               const double* __restrict__ B, int ldb, int strip) {
    const int lane = threadIdx.x & 31, lq = lane >> 2, lr = lane & 3;
    double a[8], b[8][4];
    const double* Ar = A + (size_t)(8 * strip + lq) * lda + lr;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) a[ks] = Ar[4 * ks];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[ks][nt] = NN ? B[(size_t)(4 * ks + lr) * ldb + 8 * nt + lq] : B[(size_t)(8 * nt + lq) * ldb + 4 * ks + lr];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma8x8x4(acc[nt][0], acc[nt][1], a[ks], b[ks][nt]);
}
__device__ __forceinline__ void strip_zero(double (&acc)[4][2]) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
}
template <bool ADD>
__device__ __forceinline__ void strip_store(const double (&acc)[4][2], double* C, int ldc, double sgn, int strip) {
    const int lane = threadIdx.x & 31, lq = lane >> 2, lr = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        double2* p = reinterpret_cast<double2*>(C + (size_t)(8 * strip + lq) * ldc + 8 * nt + 2 * lr);
        double2 v = ADD ? *p : make_double2(0.0, 0.0);
        v.x += sgn * acc[nt][0]; v.y += sgn * acc[nt][1];
        *p = v;
    }
}
// strips when every quad of warps gets at most two tiles of the job
__device__ __forceinline__ bool use_strips(long long items, int nparts) { return items * 2 <= (long long)nparts * NW; }

// Gram rows part, part + W, ... of the whole active set (W = warps of all CTAs of the fit) into the square A, incl. the
// identity padding up to a multiple of 32 (cf. gram_rows: same expansion, same deterministic combination of lanes)
__device__ void gram_rows_A(const Ctx& c, const Tiles T, double sigma, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int cap = GEMM_SMEM_DOUBLES / NW;
    const int tagcap = ((c.smd - GEMM_SMEM_DOUBLES) * 8) / NW;
    const double2* __restrict__ cscq = c.cscq;
    const int2* __restrict__ rowcb = c.rowcb;
    for (int ia = part * NW + wid; ia < T.lda; ia += nparts * NW) {
        double* dst = T.A + (size_t)ia * T.lda;
        if (ia >= T.na) {                                       // identity padding
            for (int q = lane; q <= ia; q += 32) dst[q] = (q == ia) ? 1.0 : 0.0;
            continue;
        }
        const int n = c.act[ia];
        double* acc = (ia + 1 <= cap) ? (c.sm + (size_t)wid * cap) : dst;
        unsigned char* tags = reinterpret_cast<unsigned char*>(c.sm + GEMM_SMEM_DOUBLES) + (size_t)wid * tagcap;
        const bool use_tags = ia + 1 <= tagcap;
        for (int q = lane; q <= ia; q += 32) acc[q] = 0.0;
        __syncwarp();
        const int beg = c.row_ptr[n], end = c.row_ptr[n + 1];
        double la_n = 0.0;
        int2 rc_n = make_int2(0, 0);
        if (beg + lane < end) { la_n = c.lam[beg + lane]; rc_n = rowcb[beg + lane]; }
        for (int jb = beg; jb < end; jb += 32) {
            const double la = la_n;
            const int cb = rc_n.x, len = (la != 0.0) ? rc_n.y : 0;
            la_n = 0.0; rc_n = make_int2(0, 0);
            if (jb + 32 + lane < end) { la_n = c.lam[jb + 32 + lane]; rc_n = rowcb[jb + 32 + lane]; }
            int maxlen = len;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
            for (int t0 = 0; t0 < maxlen; t0 += GCH) {
                int ibs[GCH];
                double vs[GCH];
#pragma unroll
                for (int u = 0; u < GCH; ++u) {
                    const int t = t0 + u;
                    int ib = -1;
                    double lv = 0.0;
                    if (t < len) {
                        const double2 rec = cscq[cb + t];
                        ib = (int)__double_as_longlong(rec.x);
                        lv = rec.y;
                    }
                    if (ib > ia) ib = -1;
                    ibs[u] = ib;
                    vs[u] = la * lv;
                }
#pragma unroll
                for (int u = 0; u < GCH; ++u) {
                    if (t0 + u >= maxlen) break;               // warp-uniform
                    const int ib = ibs[u];
                    const double v = vs[u];
                    bool lost = false;
                    if (use_tags) {
                        if (ib >= 0) tags[ib] = (unsigned char)lane;
                        __syncwarp();
                        lost = (ib >= 0) && (tags[ib] != (unsigned char)lane);
                    }
                    if (use_tags && !__any_sync(0xffffffffu, lost)) {
                        if (ib >= 0) acc[ib] += v;
                    } else {
                        const unsigned amask = __ballot_sync(0xffffffffu, ib >= 0);
                        if (ib >= 0) {
                            const unsigned grp = __match_any_sync(amask, ib);
                            const int leader = __ffs(grp) - 1;
                            unsigned rest = grp & ~(1u << leader);
                            double ssum = __shfl_sync(amask, v, leader);
                            while (__any_sync(amask, rest != 0)) {
                                const int src = rest ? (__ffs(rest) - 1) : lane;
                                const double ov = __shfl_sync(amask, v, src);
                                if (rest) { ssum += ov; rest &= rest - 1; }
                            }
                            if (lane == leader) acc[ib] += ssum;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        const double b0 = c.beta0[n];
        const double dd = c.dvec[ia];
        for (int q = lane; q <= ia; q += 32) {
            const double v = acc[q];
            dst[q] = (q == ia) ? sigma * (dd + v) + 1.0 / (b0 * b0) : sigma * v;
        }
        __syncwarp();
    }
}

// panel below diagonal tile kb: A[ib][kb] <- A[ib][kb] * Dinv[kb]^T (= L[ib][kb]); a warp per tile, in place
__device__ void tiles_trsm(const Tiles T, int kb, int part, int nparts) {
    const int wid = threadIdx.x >> 5;
    const double* Di = T.Dinv + (size_t)kb * 1024;
    if (use_strips(T.nt - kb - 1, nparts)) {
        const int strip = wid & 3;
        for (int ib = kb + 1 + ((part * NW + wid) >> 2); ib < T.nt; ib += (nparts * NW) >> 2) {
            double* C = T.A + (size_t)(32 * ib) * T.lda + 32 * kb;
            double acc[4][2];
            strip_zero(acc);
            tile_mma_strip<false>(acc, C, T.lda, Di, 32, strip);
            __syncwarp();                                       // the strip has been read before it is overwritten
            strip_store<false>(acc, C, T.lda, 1.0, strip);
        }
        return;
    }
    for (int ib = kb + 1 + part * NW + wid; ib < T.nt; ib += nparts * NW) {
        double* C = T.A + (size_t)(32 * ib) * T.lda + 32 * kb;
        double acc[4][4][2];
        tile_zero(acc);
        tile_mma<false>(acc, C, T.lda, Di, 32);
        __syncwarp();                                           // the whole tile has been read before it is overwritten
        tile_store<false>(acc, C, T.lda, 1.0);
    }
}
// trailing update after tile column kb: A[ib][jb] -= L[ib][kb] L[jb][kb]^T for kb < jb <= ib; a warp per tile
__device__ void tiles_update(const Tiles T, int kb, int part, int nparts, int t0 = 0, int t1 = 0x7fffffff) {
    const int wid = threadIdx.x >> 5;
    const int m = T.nt - kb - 1;
    const int items = min((m * (m + 1)) >> 1, t1);
    if (use_strips(items - t0, nparts)) {
        const int strip = wid & 3;
        for (int t = t0 + ((part * NW + wid) >> 2); t < items; t += (nparts * NW) >> 2) {
            int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
            while (((i + 1) * (i + 2)) >> 1 <= t) ++i;
            while (((i * (i + 1)) >> 1) > t) --i;
            const int j = t - ((i * (i + 1)) >> 1);
            const int ib = kb + 1 + i, jb = kb + 1 + j;
            double acc[4][2];
            strip_zero(acc);
            tile_mma_strip<false>(acc, T.A + (size_t)(32 * ib) * T.lda + 32 * kb, T.lda, T.A + (size_t)(32 * jb) * T.lda + 32 * kb, T.lda, strip);
            strip_store<true>(acc, T.A + (size_t)(32 * ib) * T.lda + 32 * jb, T.lda, -1.0, strip);
        }
        return;
    }
    for (int t = t0 + part * NW + wid; t < items; t += nparts * NW) {
        int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while (((i + 1) * (i + 2)) >> 1 <= t) ++i;
        while (((i * (i + 1)) >> 1) > t) --i;
        const int j = t - ((i * (i + 1)) >> 1);
        const int ib = kb + 1 + i, jb = kb + 1 + j;
        double acc[4][4][2];
        tile_zero(acc);
        tile_mma<false>(acc, T.A + (size_t)(32 * ib) * T.lda + 32 * kb, T.lda, T.A + (size_t)(32 * jb) * T.lda + 32 * kb, T.lda);
        tile_store<true>(acc, T.A + (size_t)(32 * ib) * T.lda + 32 * jb, T.lda, -1.0);
    }
}

// Cholesky factor of diagonal tile kb and the inverse of that factor (one warp, shared memory), fit CTA only
__device__ void tiles_potrf(const Ctx& c, const Tiles T, int kb) {
    double (*Sd)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(c.sm + GEMM_SMEM_DOUBLES);
    double (*Xd)[XD_LD] = reinterpret_cast<double (*)[XD_LD]>(c.sm + GEMM_SMEM_DOUBLES + NB * (NB + 1));
    const double* D = T.A + (size_t)(32 * kb) * T.lda + 32 * kb;
    for (int e = threadIdx.x; e < NB * NB; e += NT) {
        const int r = e >> 5, q = e & 31;
        Sd[r][q] = (q <= r) ? D[(size_t)r * T.lda + q] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < 32) potrf32_warp(Sd, Xd, NB);
    __syncthreads();
    double* Di = T.Dinv + (size_t)kb * 1024;
    double* Xkk = T.XI + (size_t)(32 * kb) * T.lda + 32 * kb;  // diagonal tile of X = L^-1 (tiles_inv_level starts from these)
    for (int e = threadIdx.x; e < NB * NB; e += NT) {
        const double v = Xd[e >> 5][e & 31];
        Di[e] = v;
        Xkk[(size_t)(e >> 5) * T.lda + (e & 31)] = v;
    }
    __syncthreads();
}
// X = L^-1 by recursive halving over blocks of tiles, bottom-up: at level s = 1, 2, 4, ... the sibling blocks
// [lo, mid) and [mid, hi) (lo = 2 s p, mid = lo + s, hi = min(lo + 2 s, nt)) already hold their own inverses X11, X22 and
//     stage 0:  T[i][j] = sum_{k = j}^{mid - 1} L[i][k] X[k][j]        i in [mid, hi), j in [lo, mid)      (L21 X11)
//     stage 1:  X[i][j] = - sum_{k = mid}^{i} X[i][k] T[k][j]                                            (- X22 T)
// fill the block below the diagonal.  Every output tile is one warp job with a fixed k order, all tiles of a stage are
// independent: log2(nt) levels of two wide jobs each instead of a chain of nt dependent steps per tile column (the
// column-by-column forward substitution this replaces: 7.5 ms of a C3 fit, 56 ms of a C5 fit).  T[i][j] is kept in the
// unused upper triangle of the factor's tile matrix, at tile (j, i); every (i, j) belongs to exactly one level.
__device__ void tiles_inv_level(const Tiles T, int s, int stage, int part, int nparts) {
    const int wid = threadIdx.x >> 5;
    const int npairs = (T.nt + 2 * s - 1) / (2 * s);
    const long long total = (long long)npairs * s * s;
    if (use_strips(total, nparts)) {
        const int strip = wid & 3;
        for (long long idx = (part * NW + wid) >> 2; idx < total; idx += (long long)((nparts * NW) >> 2)) {
            const int ii = (int)(idx % s);
            const long long rest = idx / s;
            const int jj = (int)(rest % s), p = (int)(rest / s);
            const int lo = 2 * s * p, mid = lo + s;
            const int i = mid + ii, j = lo + jj;
            if (i >= T.nt) continue;
            double acc[4][2];
            strip_zero(acc);
            if (stage == 0) {
                for (int k = j; k < mid; ++k)
                    tile_mma_strip<true>(acc, T.A + (size_t)(32 * i) * T.lda + 32 * k, T.lda, T.XI + (size_t)(32 * k) * T.lda + 32 * j, T.lda, strip);
                strip_store<false>(acc, T.A + (size_t)(32 * j) * T.lda + 32 * i, T.lda, 1.0, strip);
            } else {
                for (int k = mid; k <= i; ++k)
                    tile_mma_strip<true>(acc, T.XI + (size_t)(32 * i) * T.lda + 32 * k, T.lda, T.A + (size_t)(32 * j) * T.lda + 32 * k, T.lda, strip);
                strip_store<false>(acc, T.XI + (size_t)(32 * i) * T.lda + 32 * j, T.lda, -1.0, strip);
            }
        }
        return;
    }
    for (long long idx = part * NW + wid; idx < total; idx += (long long)nparts * NW) {
        const int ii = (int)(idx % s);
        const long long rest = idx / s;
        const int jj = (int)(rest % s), p = (int)(rest / s);
        const int lo = 2 * s * p, mid = lo + s;
        const int i = mid + ii, j = lo + jj;
        if (i >= T.nt) continue;
        double acc[4][4][2];
        tile_zero(acc);
        if (stage == 0) {
            for (int k = j; k < mid; ++k)
                tile_mma<true>(acc, T.A + (size_t)(32 * i) * T.lda + 32 * k, T.lda, T.XI + (size_t)(32 * k) * T.lda + 32 * j, T.lda);
            tile_store<false>(acc, T.A + (size_t)(32 * j) * T.lda + 32 * i, T.lda, 1.0);
        } else {
            for (int k = mid; k <= i; ++k)
                tile_mma<true>(acc, T.XI + (size_t)(32 * i) * T.lda + 32 * k, T.lda, T.A + (size_t)(32 * j) * T.lda + 32 * k, T.lda);
            tile_store<false>(acc, T.XI + (size_t)(32 * i) * T.lda + 32 * j, T.lda, -1.0);
        }
    }
}
// X = L^-1 by tile columns jb = part, part + nparts, ... (a column per CTA): X[jb][jb] = Dinv[jb];
// X[ib][jb] = -Dinv[ib] * sum_{kb = jb}^{ib - 1} L[ib][kb] X[kb][jb].  The 16 warps split the k range of the sum (warp w
// takes kb = jb + w, jb + w + NW, ...), their partial tiles are added in warp order through shared memory, then warp w
// multiplies the 8 x 8 sub-tile w of the product with Dinv[ib].
__device__ void tiles_inverse(const Ctx& c, const Tiles T, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lq = lane >> 2, lr = lane & 3;
    double* part_s = c.sm;                                      // [NW][32][32] partial tiles
    double* sum_s = c.sm + NW * 1024;                           // [32][33] their sum
    for (int jb = part; jb < T.nt; jb += nparts) {
        {
            const double* Di = T.Dinv + (size_t)jb * 1024;
            double* Xjj = T.XI + (size_t)(32 * jb) * T.lda + 32 * jb;
            for (int e = threadIdx.x; e < 1024; e += NT) Xjj[(size_t)(e >> 5) * T.lda + (e & 31)] = Di[e];
        }
        __syncthreads();
        for (int ib = jb + 1; ib < T.nt; ++ib) {
            double acc[4][4][2];
            tile_zero(acc);
            for (int kb = jb + wid; kb < ib; kb += NW)
                tile_mma<true>(acc, T.A + (size_t)(32 * ib) * T.lda + 32 * kb, T.lda, T.XI + (size_t)(32 * kb) * T.lda + 32 * jb, T.lda);
            double* mine = part_s + wid * 1024;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mine[(8 * mt + lq) * 32 + 8 * nt + 2 * lr] = acc[mt][nt][0];
                    mine[(8 * mt + lq) * 32 + 8 * nt + 2 * lr + 1] = acc[mt][nt][1];
                }
            __syncthreads();
            const int nw_used = min(NW, ib - jb);
            for (int e = threadIdx.x; e < 1024; e += NT) {
                double v = part_s[e];
                for (int w = 1; w < nw_used; ++w) v += part_s[w * 1024 + e];
                sum_s[(e >> 5) * 33 + (e & 31)] = v;
            }
            __syncthreads();
            // X[ib][jb] = -Dinv[ib] (32 x 32, lower) * sum: warp w computes sub-tile (w / 4, w % 4)
            for (int st = wid; st < 16; st += NW) {
                const int mt = st >> 2, nt = st & 3;
                const double* Di = T.Dinv + (size_t)ib * 1024;
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double a = Di[(8 * mt + lq) * 32 + 4 * ks + lr];
                    const double b = sum_s[(4 * ks + lr) * 33 + 8 * nt + lq];
                    dmma8x8x4(d0, d1, a, b);
                }
                double* X = T.XI + (size_t)(32 * ib + 8 * mt + lq) * T.lda + 32 * jb + 8 * nt + 2 * lr;
                X[0] = -d0; X[1] = -d1;
            }
            __syncthreads();                                    // X[ib][jb] is read (through L2) by the next rows of this column
            __threadfence_block();
        }
    }
}
// w = X b (rows), then mu = X^T w and beta = column sums of squares (columns) on the row-major inverse factor
__device__ void tiles_wvec(const Ctx& c, const Tiles T, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = part * NW + wid; i < T.na; i += nparts * NW) {
        double s = 0.0;
        const double* row = T.XI + (size_t)i * T.lda;
        for (int q = lane; q <= i; q += 32) s += row[q] * c.bvec[q];
        s = warp_sum(s);
        if (lane == 0) c.wvec[i] = s;
    }
}
// mu = X^T w, beta = column sums of squares of X.  Work items = (tile column of 32 columns, row segment): one warp per
// item walks the rows of its segment with coalesced 256-byte reads and leaves partial sums in c.PP (free during the tile
// solve); the fit CTA then adds the segments of every column in ascending order.  The segmentation depends on the
// number of active rows only, so the result does not depend on the number of CTAs.
__device__ __forceinline__ int mubeta_seg(int na) {
    const int nseg = min(56, (na + 255) >> 8);                    // 2 nseg lda doubles must fit c.PP (KSEG_MAX NB (N + ROWPAD))
    return (((na + nseg - 1) / nseg) + 31) & ~31;                 // rows per segment, a multiple of the tile height
}
__device__ void tiles_mubeta(const Ctx& c, const Tiles T, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg = mubeta_seg(T.na), nseg = (T.na + seg - 1) / seg;
    for (int idx = part * NW + wid; idx < T.nt * nseg; idx += nparts * NW) {
        const int jt = idx / nseg, sg = idx - jt * nseg;
        if ((sg + 1) * seg <= 32 * jt) continue;                    // segment entirely above the diagonal tile
        const int cc = 32 * jt + lane;
        const int ilo = max(sg * seg, 32 * jt), ihi = min((sg + 1) * seg, T.na);
        double m = 0.0, v = 0.0;
        const double* col = T.XI + cc;
        for (int i = ilo; i < ihi; i += 4) {
            double x[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = i + u < ihi && i + u >= cc;
                x[u] = ok ? col[(size_t)(i + u) * T.lda] : 0.0;
                w[u] = (i + u < ihi) ? c.wvec[i + u] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { m += x[u] * w[u]; v += x[u] * x[u]; }
        }
        c.PP[(size_t)(2 * sg) * T.lda + cc] = m;
        c.PP[(size_t)(2 * sg + 1) * T.lda + cc] = v;
    }
}
__device__ void tiles_mubeta_sum(const Ctx& c, const Tiles T) {
    const int seg = mubeta_seg(T.na), nseg = (T.na + seg - 1) / seg;
    for (int cc = threadIdx.x; cc < T.na; cc += NT) {
        double m = 0.0, v = 0.0;
        for (int sg = (cc & ~31) / seg; sg < nseg; ++sg) {
            m += c.PP[(size_t)(2 * sg) * T.lda + cc];
            v += c.PP[(size_t)(2 * sg + 1) * T.lda + cc];
        }
        const int n = c.act[cc];
        c.mu[n] = m;
        c.beta[n] = v;
    }
}
// the whole solve, driven by the fit CTA; job types 8..13 are served by helper_loop
__device__ __noinline__ void a2_tiles(const Ctx& c, double sigma, int na) {
    const Tiles T = make_tiles(c, na);
    const bool dist = HELPERS && c.ct > 1;
    const int np = dist ? c.ct : 1;
    if (dist) { if (threadIdx.x == 0) *reinterpret_cast<double*>(c.job + 8) = sigma; post_job(c, 8, na, 0, 0); }
    gram_rows_A(c, T, sigma, 0, np);
    if (dist) wait_helpers(c); else __syncthreads();
    phase_mark(c, 1);
    // Look-ahead: the factor of diagonal tile kb + 1 only needs tile (kb+1, kb+1) of the trailing update of column kb.  With
    // helpers, the fit CTA applies that one tile itself and factors it while the helpers update all the other tiles.
    tiles_potrf(c, T, 0);
    phase_mark(c, 2);                                           // (diagnostics: 2 = diagonal tiles, 4 = panels, 3 = trailing updates)
    for (int kb = 0; kb + 1 < T.nt; ++kb) {
        const bool d1 = dist && (T.nt - kb - 1) > NW;           // more panel tiles than this CTA has warps
        if (d1) post_job(c, 9, na, kb, 0);
        tiles_trsm(T, kb, 0, d1 ? np : 1);
        if (d1) wait_helpers(c); else { __threadfence(); __syncthreads(); }
        phase_mark(c, 4);
        const int m = T.nt - kb - 1;
        const bool d2 = dist && ((m * (m + 1)) >> 1) > NW;
        if (d2) {
            post_job(c, 10, na, kb, 0);                         // helpers: items 1 .. of the trailing update, shared among them
            tiles_update(T, kb, 0, 1, 0, 1);                    // fit CTA: item 0 = tile (kb+1, kb+1) ...
            __threadfence();
            __syncthreads();
            tiles_potrf(c, T, kb + 1);                          // ... and its factor, under the helpers' update
            phase_mark(c, 2);
            wait_helpers(c);
        } else {
            tiles_update(T, kb, 0, 1);
            __threadfence();
            __syncthreads();
            phase_mark(c, 3);
            tiles_potrf(c, T, kb + 1);
            phase_mark(c, 2);
        }
        phase_mark(c, 3);
    }
    if (g_phase_enable & 4096) {                                // diagnostics: the column-by-column forward substitution
        if (dist) post_job(c, 11, na, 0, -1);
        tiles_inverse(c, T, 0, np);
        if (dist) wait_helpers(c); else { __threadfence(); __syncthreads(); }
    } else {
        for (int s = 1; s < T.nt; s <<= 1)
            for (int stage = 0; stage < 2; ++stage) {
                const long long items = (long long)((T.nt + 2 * s - 1) / (2 * s)) * s * s;
                const bool d3 = dist && items > NW;
                if (d3) post_job(c, 11, na, s, stage);
                tiles_inv_level(T, s, stage, 0, d3 ? np : 1);
                if (d3) wait_helpers(c); else { __threadfence(); __syncthreads(); }
            }
    }
    phase_mark(c, 5);
    if (dist) post_job(c, 12, na, 0, 0);
    tiles_wvec(c, T, 0, np);
    if (dist) wait_helpers(c); else { __threadfence(); __syncthreads(); }
    if (dist) post_job(c, 13, na, 0, 0);
    tiles_mubeta(c, T, 0, np);
    if (dist) wait_helpers(c); else { __threadfence(); __syncthreads(); }
    tiles_mubeta_sum(c, T);
    __syncthreads();
    phase_mark(c, 6);
}

// Job types: 1 / 2 panel GEMM (upper / lower; a = i0, b = nb), 3 Newton rows of c.dlist (a = rows), 4 Monte-Carlo means
// (a = key buffer, b = samples), 5 w = X b, 6 mu / beta (a = active rows), 7 Gram rows of a block (a = i0, b = nb, sigma at
// int offset 8), 0 quit.
// chinfo[i].w = 1 when chain row i shares a trial with chain row i - 1 (their steps must not overlap), for the two-team
// chain sweep.  One warp per pair: every entry of row i is looked up in row i - 1 by binary search (the trials of a CSR
// row are ascending).  Structural test (entries count whatever their posterior is): conservative and cheap.
__device__ void chain_deps(const Ctx& c, int nchain, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = 1 + part * NW + wid; i < nchain; i += nparts * NW) {
        const int4 a = c.chinfo[i - 1], b = c.chinfo[i];
        const int* ka = c.col_k + a.y;
        int hit = 0;
        for (int q = lane; q < b.z; q += 32) {
            const int k = c.col_k[b.y + q];
            int lo = 0, hi = a.z;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ka[mid] < k) lo = mid + 1; else hi = mid;
            }
            hit |= (lo < a.z && ka[lo] == k) ? 1 : 0;
        }
        hit = __any_sync(0xffffffffu, hit);
        if (lane == 0) reinterpret_cast<int*>(c.chinfo + i)[3] = hit ? 1 : 0;
    }
}

// ---- O(K) / O(nnz) passes of one iteration as helper jobs (single large fits): pure maps over trials, by-trial entries
// or rows, split into contiguous trial ranges / interleaved rows.  Every output element is computed by one thread (or
// one warp) exactly as in the single-CTA code, so the fit stays bitwise identical whatever the number of helpers.
__device__ __forceinline__ int part_lo(int n, int part, int nparts) { return (int)((long long)n * part / nparts); }
// job 16 (a2 set-up): Gram-expansion records of the by-trial index in use, D and b of the active rows
__device__ void job_a2_rows(const Ctx& c, double sigma, int na, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i0 = part_lo(c.unnz, part, nparts), i1 = part_lo(c.unnz, part + 1, nparts);
#pragma unroll 8
    for (int i = i0 + threadIdx.x; i < i1; i += NT)
        c.cscq[i] = make_double2(__longlong_as_double((long long)c.ainv[c.ucsc_row[i]]), c.lamT[i]);
    for (int ia = part * NW + wid; ia < na; ia += nparts * NW) row_dvec_bvec(c, ia, sigma);
}
// job 14 (a3 set-up): fresh prediction into c.pred (global) and the per-entry constant part of the sigmoid argument
__device__ void job_pred_cst(const Ctx& c, double sigma, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k0 = part_lo(c.K, part, nparts), k1 = part_lo(c.K, part + 1, nparts);
    for (int k = k0 + threadIdx.x; k < k1; k += NT) c.pred[k] = trial_pred(c, c.ucol_ptr[k], c.ucol_ptr[k + 1]);
    for (int n = part * NW + wid; n < c.N; n += nparts * NW)
        if (!c.dcnt[n]) row_cst(c, n, sigma);
}
// job 15 (after the sweep): by-trial copy of the new lam, residual of a6 (caviar.py:238-244) and the
// spontaneous-event mask of a8 (caviar.py:155) for a contiguous range of trials
__device__ void job_lamT_resid(const Ctx& c, double spont_orth, int part, int nparts) {
    const int k0 = part_lo(c.K, part, nparts), k1 = part_lo(c.K, part + 1, nparts);
    const int i0 = c.ucol_ptr[k0], i1 = c.ucol_ptr[k1];
#pragma unroll 8
    for (int i = i0 + threadIdx.x; i < i1; i += NT) c.lamT[i] = c.lam[c.ucsc_pos[i]];
    __syncthreads();
    for (int k = k0 + threadIdx.x; k < k1; k += NT) {
        double s = 0.0;
        unsigned char bl = 0;
        const int end = c.ucol_ptr[k + 1];
        for (int i = c.ucol_ptr[k]; i < end; i += 4) {         // four entries' dependent gathers together, sums in list order
            int r[4];
            double l[4], m[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = i + u < end;
                r[u] = ok ? c.ucsc_row[i + u] : -1;
                l[u] = ok ? c.lamT[i + u] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) m[u] = r[u] >= 0 ? c.mu[r[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r[u] >= 0) { s += m[u] * l[u]; bl |= (l[u] >= spont_orth); }
        }
        c.resid[k] = c.y[k] - s;
        c.blocked[k] = bl;
    }
}
// the passes above are worth a job hand-off (a few microseconds) from this much work on
__device__ __forceinline__ bool dist_passes(const Ctx& c) {
    return HELPERS && c.ct > 1 && !(g_phase_enable & 1024) && (long long)c.K + c.unnz >= 32768;
}

__device__ void helper_loop(const Ctx& c, GemmPipe& gp) {
    __shared__ int s_job[4];
    const int ldr = c.N + ROWPAD;
    const double* powers = reinterpret_cast<const double*>(c.job + 32);
    int seen = 0;
    __syncthreads();                                  // mbarriers initialised
    for (;;) {
        if (threadIdx.x == 0) {
            int type = -1;
            while (type < 0) {
                if (ld_acquire_gpu(&c.job[0]) != seen) type = c.job[1];
                else __nanosleep(200);
            }
            s_job[0] = type; s_job[1] = c.job[2]; s_job[2] = c.job[3]; s_job[3] = c.job[4];
        }
        __syncthreads();
        const int type = s_job[0], a = s_job[1], b = s_job[2], d3 = s_job[3];
        __syncthreads();
        if (type == 0) break;
        ++seen;
        if (type == 1 || type == 2) {
            asm volatile("fence.proxy.async;\n" ::: "memory");      // operands were written through the generic proxy of another SM
            if (type == 1) panel_gemm<true>(c, ldr, a, b, c.PA, c.PB, gp, c.role, c.ct, d3);
            else panel_gemm<false>(c, ldr, a, b, c.PB, c.PA, gp, c.role, c.ct, d3);
        } else if (type == 3) {
            newton_rows(c, powers, c.dlist, a, c.role, c.ct);
        } else if (type == 4) {
            mc_means(c, c.keys + (size_t)a * 2 * c.N, b, powers, c.role, c.ct);
        } else if (type == 5) {
            a2_wvec(c, a, ldr, c.role, c.ct);
        } else if (type == 6) {
            a2_mubeta(c, a, ldr, c.role, c.ct);
        } else if (type >= 8 && type <= 13) {                    // tile solve (a2_tiles): a = active rows, b = tile column
            Ctx h = c;
            if (c.job[5]) { h.ucol_ptr = c.ccol_ptr; h.ucsc_row = c.ccsc_row; h.ucsc_pos = c.ccsc_pos; }
            h.unnz = c.job[6];
            const Tiles T = make_tiles(h, a);
            if (type == 8) gram_rows_A(h, T, *reinterpret_cast<const double*>(c.job + 8), c.role, c.ct);
            else if (type == 9) tiles_trsm(T, b, c.role, c.ct);
            else if (type == 10) tiles_update(T, b, c.role - 1, c.ct - 1, 1);       // item 0 is the fit CTA's (look-ahead)
            else if (type == 11) { if (d3 < 0) tiles_inverse(h, T, c.role, c.ct); else tiles_inv_level(T, b, d3, c.role, c.ct); }
            else if (type == 12) tiles_wvec(h, T, c.role, c.ct);
            else tiles_mubeta(h, T, c.role, c.ct);
        } else if (type == 17) {
            chain_deps(c, a, c.role, c.ct);
        } else if (type == 7 || (type >= 14 && type <= 16)) {
            Ctx h = c;                                           // the by-trial index the fit CTA currently uses
            if (c.job[5]) { h.ucol_ptr = c.ccol_ptr; h.ucsc_row = c.ccsc_row; h.ucsc_pos = c.ccsc_pos; }
            h.unnz = c.job[6];
            const double dv = *reinterpret_cast<const double*>(c.job + 8);
            if (type == 7) gram_rows(h, a, b, dv, c.role, c.ct);
            else if (type == 14) job_pred_cst(h, dv, c.role, c.ct);
            else if (type == 15) job_lamT_resid(h, dv, c.role, c.ct);
            else job_a2_rows(h, dv, a, c.role, c.ct);
        }
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); red_release_gpu(&c.job[16], 1); }
    }
}

__device__ __noinline__ void phase_a2(const Ctx& c, double sigma, int* na_s, GemmPipe& gp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int lq = lane >> 2, lr = lane & 3;
    const int N = c.N;
    const int na = block_compact(N, [&](int i) { return c.rownz[i] > 0; }, c.act, c.ainv, c.red);
    const int ldr = N + ROWPAD;                   // rows per 256-column tile of X
    if (threadIdx.x == 0) *na_s = na;
    // inactive rows decouple: mu = mu0, beta = beta0^2 (variance)
    for (int n = threadIdx.x; n < N; n += NT)
        if (c.rownz[n] == 0) { c.mu[n] = c.mu0[n]; c.beta[n] = c.beta0[n] * c.beta0[n]; }
    if (dist_passes(c)) {
        if (threadIdx.x == 0) *reinterpret_cast<double*>(c.job + 8) = sigma;
        post_job(c, 16, na, 0, 0);
        job_a2_rows(c, sigma, na, 0, c.ct);
        wait_helpers(c);
    } else {
    // per CSC entry: (active index of its row, lam) in one 16-byte record for the Gram expansion
#pragma unroll 8
    for (int i = threadIdx.x; i < c.unnz; i += NT)
        c.cscq[i] = make_double2(__longlong_as_double((long long)c.ainv[c.ucsc_row[i]]), c.lamT[i]);
    // per active row: D = sum lam(1-lam), b = sigma * sum lam*y + mu0/beta0^2
    for (int ia = wid; ia < na; ia += NW) row_dvec_bvec(c, ia, sigma);
    }
    __syncthreads();
    phase_mark(c, 0);
    if (na == 0) return;
    if (!(g_phase_enable & 256) && na <= a2_small_capacity(c.smd)) { a2_small(c, sigma, na); return; }
    if (HELPERS && !(g_phase_enable & 512)) { a2_tiles(c, sigma, na); return; }

    // diagonal block -> its Cholesky factor, and the inverse of that factor (strict upper part zeroed); both live
    // behind the GEMM ring (the Gram conflict tags reuse the same bytes at a different time)
    double (*Sd)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(c.sm + GEMM_SMEM_DOUBLES);
    double (*Xd)[XD_LD] = reinterpret_cast<double (*)[XD_LD]>(c.sm + GEMM_SMEM_DOUBLES + NB * (NB + 1));
    double* X = c.X;
    double* PA = c.PA;
    double* PB = c.PB;
    for (int i0 = 0; i0 < na; i0 += NB) {
        const int nb = min(NB, na - i0);
        if (HELPERS && c.ct > 1 && nb >= 8) {                // rows of the block spread over the helper CTAs
            if (threadIdx.x == 0) *reinterpret_cast<double*>(c.job + 8) = sigma;
            post_job(c, 7, i0, nb, 0);
            gram_rows(c, i0, nb, sigma, 0, c.ct);
            wait_helpers(c);
        } else {
            gram_rows(c, i0, nb, sigma);          // PA[r][0..i0+r] = M[i0+r][.]
            __syncthreads();
        }
        phase_mark(c, 1);
        if (i0 > 0) panel_gemm_dist<true>(c, ldr, i0, nb, PA, PB, gp); // PB = Lrow = A[I,0:i0] X11^T
        phase_mark(c, 2);
        // S = A[I,I] - Lrow Lrow^T : warp w owns the 8x8 tile (w/4, w%4) and runs the whole k range with DMMA,
        // four independent accumulator pairs hide the dependent-issue latency.
        {
            for (int tile = wid; tile < 16; tile += NW) {
                const int mt = tile >> 2, nt = tile & 3;
                if (nt > mt) continue;
                double d[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                if (i0 > 0) {
                    const int ra = 8 * mt + lq, rb = 8 * nt + lq;
                    const bool va = ra < nb, vb = rb < nb;
                    for (int k0 = 0; k0 < i0; k0 += 16) {
                        double av[4], bv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = k0 + 4 * u + lr;
                            av[u] = (va && k < i0) ? PB[pidx(ra, k)] : 0.0;
                            bv[u] = (vb && k < i0) ? PB[pidx(rb, k)] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) dmma8x8x4(d[u][0], d[u][1], av[u], bv[u]);
                    }
                }
                const double s0 = (d[0][0] + d[1][0]) + (d[2][0] + d[3][0]);
                const double s1 = (d[0][1] + d[1][1]) + (d[2][1] + d[3][1]);
                const int r = 8 * mt + lq, cc = 8 * nt + 2 * lr;
                if (r < nb) {
                    Sd[r][cc] = PA[pidx(r, i0 + cc)] - s0;
                    Sd[r][cc + 1] = PA[pidx(r, i0 + cc + 1)] - s1;
                }
            }
        }
        __syncthreads();
        // Cholesky of the nb x nb diagonal block and the inverse of its factor by one warp (potrf32_warp)
        if (threadIdx.x < 32) potrf32_warp(Sd, Xd, nb);
        __syncthreads();
        phase_mark(c, 3);
        if (i0 > 0) {
            panel_gemm_dist<false>(c, ldr, i0, nb, PB, PA, gp);         // PA = W = Lrow X11
            phase_mark(c, 4);
            // X[I, 0:i0] = -Xd W (32x32 lower-triangular times 32 x i0, DMMA), mirrored into the upper half
            const int nsl = (i0 + 15) / 16;
            for (int sl = wid; sl < nsl; sl += NW) {
                const int cb0 = sl * 16;
                double acc[4][2][2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
                double bvv[8][2];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const int r2 = 4 * ks + lr, cc = cb0 + 8 * nt + lq;
                        bvv[ks][nt] = (r2 < nb && cc < i0) ? PA[pidx(r2, cc)] : 0.0;
                    }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    double af[4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) af[mt] = Xd[8 * mt + lq][4 * ks + lr];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bvv[ks][nt]);
                }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    const int r = 8 * mt + lq;
                    if (r >= nb) continue;
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int cc = cb0 + 8 * nt + 2 * lr + e;
                            if (cc < i0) {
                                const double v = -acc[mt][nt][e];
                                X[xidx(i0 + r, cc, ldr)] = v;
                                X[xidx(cc, i0 + r, ldr)] = v;
                            }
                        }
                }
            }
        }
        for (int e = threadIdx.x; e < nb * nb; e += NT) {
            const int r = e / nb, r2 = e - r * nb;
            if (r2 <= r) {
                X[xidx(i0 + r, i0 + r2, ldr)] = Xd[r][r2];
                X[xidx(i0 + r2, i0 + r, ldr)] = Xd[r][r2];
            }
        }
        __syncthreads();
        phase_mark(c, 5);
    }
    // w = X b ; mu = X^T w ; beta = column sums of squares of X
    if (HELPERS && c.ct > 1 && na > 256) {
        post_job(c, 5, na, 0, 0);
        a2_wvec(c, na, ldr, 0, c.ct);
        wait_helpers(c);
        post_job(c, 6, na, 0, 0);
        a2_mubeta(c, na, ldr, 0, c.ct);
        wait_helpers(c);
    } else {
        a2_wvec(c, na, ldr, 0, 1);
        __syncthreads();
        a2_mubeta(c, na, ldr, 0, 1);
        __syncthreads();
    }
    phase_mark(c, 6);
}

// ------------------------------------------------------------------------------------------------ a3
// One neuron of update_lam's sweep (caviar.py:200-227, reduced form App. A.8), executed by one warp.
// cp/cs/lo: the row's packed (trial | power<<27) entries, sigmoid-argument constants (in) / new posteriors (out),
// and old posteriors -- either staged in shared memory (chain) or the global arrays themselves.
// chain=true: the neuron reads and updates the running prediction (mu[n] != 0).
#define NEG_INF (-CUDART_INF)

// what a sweep step needs, held in registers (the big Ctx lives in local memory; keep it off the critical path)
struct RowCtx {
    double *lam, *sp, *slam, *slam2;
    int *n0p, *n1p, *rownz;
    int P;
};
__device__ __forceinline__ RowCtx make_rowctx(const Ctx& c) {
    RowCtx r;
    r.lam = c.lam; r.sp = c.sp; r.slam = c.slam; r.slam2 = c.slam2;
    r.n0p = c.n0p; r.n1p = c.n1p; r.rownz = c.rownz; r.P = c.P;
    return r;
}

template <int PT>
__device__ __forceinline__ void sweep_row(const RowCtx c, int n, int beg, int len, bool chain, double mu_n,
                                          const int* cntp, const int* nmask, const int* __restrict__ cp, double* cs,
                                          const double* lo, double sigma, double thr, double minspk, bool gate,
                                          double* pred) {
    const int lane = threadIdx.x & 31;
    const int dbg = g_phase_enable;                // read once: the stores of the entry loop below could alias it
    const bool prof = chain && dbg && blockIdx.x == 0 && lane == 0;
    long long tp0 = 0;
    if (prof) tp0 = clock64();
    const double coef = sigma * mu_n;
    double tot, accp[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) accp[p] = 0.0;
    bool rare = false;
    int q = lane;
    for (; q + 32 < len; q += 64) {                      // two independent entries per trip (ILP across the exp/div chains)
        const int pk0 = cp[q], pk1 = cp[q + 32];
        const double cj0 = cs[q], cj1 = cs[q + 32];
        const double x0 = chain ? (cj0 - coef * pred[pk0 & 0x7ffffff]) : cj0;
        const double x1 = chain ? (cj1 - coef * pred[pk1 & 0x7ffffff]) : cj1;
        const double e0 = sigmoid_fast(x0), e1 = sigmoid_fast(x1);
        cs[q] = e0; cs[q + 32] = e1;
        const int pw0 = pk0 >> 27, pw1 = pk1 >> 27;
#pragma unroll
        for (int p = 0; p < PT; ++p) { accp[p] += (pw0 == p) ? e0 : 0.0; accp[p] += (pw1 == p) ? e1 : 0.0; }
        rare |= (e0 == 1.0) || (e0 == 0.0) || (e1 == 1.0) || (e1 == 0.0);
    }
    if (q < len) {
        const int pk = cp[q];
        const double cj = cs[q];
        const double x = chain ? (cj - coef * pred[pk & 0x7ffffff]) : cj;
        const double est = sigmoid_fast(x);
        cs[q] = est;
        const int pw = pk >> 27;
#pragma unroll
        for (int p = 0; p < PT; ++p) accp[p] += (pw == p) ? est : 0.0;
        rare |= (est == 1.0) || (est == 0.0);
    }
    if (prof) { const long long t = clock64(); g_phase_cycles[20] += t - tp0; tp0 = t; }
    // per-power sums (warp-uniform guard skips unused slots); the row total is their sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                   // P butterflies interleaved level by level
        double tmpv[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p)
            if (p < c.P) tmpv[p] = __shfl_xor_sync(0xffffffffu, accp[p], o);
#pragma unroll
        for (int p = 0; p < PT; ++p)
            if (p < c.P) accp[p] += tmpv[p];
    }
    tot = 0.0;
#pragma unroll
    for (int p = 0; p < PT; ++p)
        if (p < c.P) tot += accp[p];
    if (prof) { const long long t = clock64(); g_phase_cycles[21] += t - tp0; tp0 = t; }
    int c0[PT], c1[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) { c0[p] = (p < c.P) ? nmask[p] : 0; c1[p] = 0; }
    if (__any_sync(0xffffffffu, rare)) {                 // exact zero / one counts (needed by update_phi's nan_to_num)
        int z0[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) z0[p] = 0;
        for (int q = lane; q < len; q += 32) {
            const int pw = cp[q] >> 27;
            const double est = cs[q];
#pragma unroll
            for (int p = 0; p < PT; ++p) { z0[p] += (pw == p && est == 0.0); c1[p] += (pw == p && est == 1.0); }
        }
#pragma unroll
        for (int p = 0; p < PT; ++p) { c0[p] += warp_sum(z0[p]); c1[p] = warp_sum(c1[p]); }
    }
    bool ok = true;
    if (gate) {
        // spike rate of power p is computed by lane p (one division latency instead of P), then broadcast
        double mine = 0.0;
        {
            const int pl = lane < c.P ? lane : 0;
            const int cnt = cntp[pl];
            double num = accp[0];
#pragma unroll
            for (int p = 1; p < PT; ++p) num = (pl == p) ? accp[p] : num;
            mine = num / ((double)cnt + 1e-4 * (cnt == 0 ? 1.0 : 0.0));
        }
        double sr[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) sr[p] = __shfl_sync(0xffffffffu, mine, p);
        ok = (pava_last_reg<PT>(sr, c.P) >= thr) && (tot >= minspk);
        if (dbg & 32) ok = true;          // debug
    }
    if (prof) { const long long t = clock64(); g_phase_cycles[22] += t - tp0; tp0 = t; }
    // second pass: commit the row, update the running prediction
    const double muok = ok ? mu_n : 0.0;
    double sl2 = 0.0;
    double* lam_row = c.lam + beg;
    for (int q = lane; q < len; q += 32) {
        const double nw = ok ? cs[q] : 0.0;
        const double old = lo[q];
        if (!(dbg & 16)) lam_row[q] = nw;
        if (chain) {
            const int k = cp[q] & 0x7ffffff;
            pred[k] = (pred[k] + muok * nw) - mu_n * old;
        }
        sl2 += nw * nw;
    }
    sl2 = warp_sum(sl2);
    if (lane == 0 && !(dbg & 8)) {
        int zeros = 0;
#pragma unroll
        for (int p = 0; p < PT; ++p)
            if (p < c.P) {
                zeros += c0[p];
                c.sp[n * PMAX + p] = ok ? accp[p] : 0.0;
                c.n0p[n * PMAX + p] = ok ? c0[p] : cntp[p];
                c.n1p[n * PMAX + p] = ok ? c1[p] : 0;
            }
        c.slam[n] = ok ? tot : 0.0;
        c.slam2[n] = sl2;
        int masked = 0;
#pragma unroll
        for (int p = 0; p < PT; ++p) if (p < c.P) masked += nmask[p];
        c.rownz[n] = ok ? (len - (zeros - masked)) : 0;
    }
    __syncwarp();
    if (prof) { const long long t = clock64(); g_phase_cycles[23] += t - tp0; g_phase_cycles[19] += 1; g_phase_cycles[24] += len; }
}

constexpr int NSTAGE = 3;
constexpr int HD = 18;           // header doubles per stage: mu, then 2*PT ints (cntp, nmask)
constexpr int TW = 4;            // warps of the chain team (one per SM sub-partition: four fp64 pipes instead of one)
constexpr int TT = 32 * TW;
constexpr int EPL = RC / TT;     // staged entries per team thread
constexpr int STAGE_DOUBLES = NSTAGE * (RC + RC + RC / 2 + TW * HD);   // cs, lo, cp (ints), one header copy per team warp
__device__ __forceinline__ void team_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(TT) : "memory"); }
__device__ __forceinline__ void team_sync_id(int team) { asm volatile("bar.sync %0, %1;\n" ::"r"(1 + team), "n"(TT) : "memory"); }
constexpr int CM_ECHAIN = 10;    // status: the two chain teams lost each other (never seen; a bounded wait instead of a hang)

// The sequential part of the sweep: neurons with mu != 0, in update order (caviar.py:196-229; quirk A.3 #3: the in-sweep
// zeroing of mu is visible to later neurons through the running prediction).  One step per neuron, steps strictly in
// order, executed by a TEAM of TW warps (one per SM sub-partition):
//   * every team thread owns the entries q = t, t + TT, ... of the row (a C3 / C4 row has ~100 entries: one sigmoid per
//     thread), staged two neurons ahead into shared memory with cp.async by the thread that will consume them;
//   * per-power sums, the sum of squares and the exact-0 / exact-1 counts are reduced inside each warp by shuffles and
//     across the TW warps through shared memory in fixed warp order (deterministic; the same in both CTA variants);
//   * every thread takes the accept / reject decision redundantly (spike-rate divisions by lanes 0..P-1, PAVA in
//     registers), then commits its own entries to lam and to the running prediction in shared memory;
//   * two named-barrier synchronisations of the team per step (after the partial sums, after the commit).
template <int PT>
__device__ __noinline__ void sweep_chain(const Ctx& c, int nchain, double sigma, double thr, double minspk, bool gate, double* pred,
                            double* stage_base) {
    const int lane = threadIdx.x & 31, tw = threadIdx.x >> 5, tt = threadIdx.x;      // the team is warps 0 .. TW-1
    __shared__ double xch[2][TW][PT + 1];
    __shared__ int xci[2][TW][2 * PT + 1];
    const RowCtx rc = make_rowctx(c);
    const int* g_colpw = c.colpw; const double* g_cst = c.cst; const double* g_lam = c.lam; const double* g_mu = c.mu;
    const int* g_cntp = c.cntp; const int* g_nmask = c.nmask; const int4* g_chinfo = c.chinfo;
    double* scs = stage_base;                                   // [NSTAGE][RC]
    double* slo = scs + NSTAGE * RC;                            // [NSTAGE][RC]
    int* scp = reinterpret_cast<int*>(slo + NSTAGE * RC);       // [NSTAGE][RC]
    double* shd = reinterpret_cast<double*>(scp + NSTAGE * RC); // [NSTAGE][TW][HD]: mu, then ints cntp[PT], nmask[PT]
    const bool prof = (g_phase_enable & 64) && blockIdx.x == 0 && tt == 0;     // per-step counters cost ~3k cycles per step: opt-in
    auto stage = [&](int4 inf, int buf) {
        const int n = inf.x, beg = inf.y, len = inf.z;
        if (len <= RC) {
            for (int q = tt; q < len; q += TT) {
                cp_async4(scp + buf * RC + q, g_colpw + beg + q);
                cp_async8(scs + buf * RC + q, g_cst + beg + q);
                cp_async8(slo + buf * RC + q, g_lam + beg + q);
            }
        }
        double* hd = shd + (buf * TW + tw) * HD;
        if (lane == 0) cp_async8(hd, g_mu + n);
        int* hi = reinterpret_cast<int*>(hd + 1);
        for (int q = lane; q < 2 * PT; q += 32)
            cp_async4(hi + q, (q < PT) ? (g_cntp + n * PMAX + q) : (g_nmask + n * PMAX + (q - PT)));
        cp_async_commit();
    };
    int4 infA = make_int4(0, 0, 0, 0), infB = infA;
    if (nchain > 0) stage(g_chinfo[0], 0);
    if (nchain > 1) stage(g_chinfo[1], 1);
    if (nchain > 2) infA = g_chinfo[2];
    if (nchain > 3) infB = g_chinfo[3];
    int4 cur = nchain > 0 ? g_chinfo[0] : infA, nxt = nchain > 1 ? g_chinfo[1] : infA;
    for (int i = 0; i < nchain; ++i) {
        long long tp0 = 0;
        if (prof) tp0 = clock64();
        if (i + 2 < nchain) { stage(infA, (i + 2) % NSTAGE); cp_async_wait<2>(); }
        else if (i + 1 < nchain) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncwarp();
        const int4 after = infA;
        infA = infB;
        if (i + 4 < nchain) infB = g_chinfo[i + 4];
        const int buf = i % NSTAGE, par = i & 1;
        const int n = cur.x, beg = cur.y, len = cur.z;
        const double* hd = shd + (buf * TW + tw) * HD;
        const double mu_n = hd[0];
        const int* cntp = reinterpret_cast<const int*>(hd + 1);
        const int* nmask = cntp + PT;
        const bool staged = len <= RC;
        const int* cp = staged ? scp + buf * RC : g_colpw + beg;
        const double* cs = staged ? scs + buf * RC : g_cst + beg;
        const double* lo = staged ? slo + buf * RC : g_lam + beg;
        const double coef = sigma * mu_n;
        // ---- pass 1: the new posteriors of this thread's entries and their partial statistics ----
        double accp[PT], se2 = 0.0;
        int z0[PT], z1[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p) { accp[p] = 0.0; z0[p] = 0; z1[p] = 0; }
        bool rare = false;
        double ev[EPL];
        int kv[EPL];
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int q = tt + j * TT;
            ev[j] = 0.0; kv[j] = -1;
            if (q < len) {
                const int pk = cp[q];
                const int k = pk & 0x7ffffff, pw = pk >> 27;
                const double e = sigmoid_fast(cs[q] - coef * pred[k]);
                ev[j] = e; kv[j] = k;
#pragma unroll
                for (int p = 0; p < PT; ++p) accp[p] += (pw == p) ? e : 0.0;
                se2 += e * e;
                if (e == 0.0 || e == 1.0) {
                    rare = true;
#pragma unroll
                    for (int p = 0; p < PT; ++p) { z0[p] += (pw == p && e == 0.0); z1[p] += (pw == p && e == 1.0); }
                }
            }
        }
        for (int q = tt + EPL * TT; q < len; q += TT) {          // rows longer than the staging capacity (not at the benchmarked sizes)
            const int pk = cp[q];
            const int k = pk & 0x7ffffff, pw = pk >> 27;
            const double e = sigmoid_fast(cs[q] - coef * pred[k]);
            const_cast<double*>(g_cst)[beg + q] = e;             // parked in the (global) constant array until the commit
#pragma unroll
            for (int p = 0; p < PT; ++p) accp[p] += (pw == p) ? e : 0.0;
            se2 += e * e;
            if (e == 0.0 || e == 1.0) {
                rare = true;
#pragma unroll
                for (int p = 0; p < PT; ++p) { z0[p] += (pw == p && e == 0.0); z1[p] += (pw == p && e == 1.0); }
            }
        }
        if (prof) { const long long t = clock64(); g_phase_cycles[20] += t - tp0; tp0 = t; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                       // P + 1 butterflies interleaved level by level
            double tmpv[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p)
                if (p < rc.P) tmpv[p] = __shfl_xor_sync(0xffffffffu, accp[p], o);
            const double t2 = __shfl_xor_sync(0xffffffffu, se2, o);
#pragma unroll
            for (int p = 0; p < PT; ++p)
                if (p < rc.P) accp[p] += tmpv[p];
            se2 += t2;
        }
        const bool wrare = __any_sync(0xffffffffu, rare);
        if (wrare) {
#pragma unroll
            for (int p = 0; p < PT; ++p) { z0[p] = warp_sum(z0[p]); z1[p] = warp_sum(z1[p]); }
        }
        if (lane == 0) {
#pragma unroll
            for (int p = 0; p < PT; ++p) xch[par][tw][p] = accp[p];
            xch[par][tw][PT] = se2;
            xci[par][tw][2 * PT] = wrare ? 1 : 0;
            if (wrare) {
#pragma unroll
                for (int p = 0; p < PT; ++p) { xci[par][tw][p] = z0[p]; xci[par][tw][PT + p] = z1[p]; }
            }
        }
        team_sync();
        // ---- every thread: totals in fixed warp order, decision ----
        double tot = 0.0;
        se2 = 0.0;
        bool anyrare = false;
#pragma unroll
        for (int p = 0; p < PT; ++p) accp[p] = 0.0;
#pragma unroll
        for (int w = 0; w < TW; ++w) {
#pragma unroll
            for (int p = 0; p < PT; ++p)
                if (p < rc.P) accp[p] += xch[par][w][p];
            se2 += xch[par][w][PT];
            anyrare |= xci[par][w][2 * PT] != 0;
        }
#pragma unroll
        for (int p = 0; p < PT; ++p)
            if (p < rc.P) tot += accp[p];
        if (prof) { const long long t = clock64(); g_phase_cycles[21] += t - tp0; tp0 = t; }
        bool ok = true;
        if (gate) {
            double mine = 0.0;                                   // spike rate of power p by lane p: one division latency
            {
                const int pl = lane < rc.P ? lane : 0;
                const int cnt = cntp[pl];
                double num = accp[0];
#pragma unroll
                for (int p = 1; p < PT; ++p) num = (pl == p) ? accp[p] : num;
                mine = num / ((double)cnt + 1e-4 * (cnt == 0 ? 1.0 : 0.0));
            }
            double sr[PT];
#pragma unroll
            for (int p = 0; p < PT; ++p) sr[p] = __shfl_sync(0xffffffffu, mine, p);
            ok = (pava_last_reg<PT>(sr, rc.P) >= thr) && (tot >= minspk);
        }
        if (prof) { const long long t = clock64(); g_phase_cycles[22] += t - tp0; tp0 = t; }
        // ---- commit this thread's entries ----
        const double muok = ok ? mu_n : 0.0;
        double* lam_row = rc.lam + beg;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int q = tt + j * TT;
            if (q < len) {
                const double nw = ok ? ev[j] : 0.0;
                const double old = lo[q];
                lam_row[q] = nw;
                const int k = kv[j];
                pred[k] = (pred[k] + muok * nw) - mu_n * old;
            }
        }
        for (int q = tt + EPL * TT; q < len; q += TT) {
            const double nw = ok ? g_cst[beg + q] : 0.0;
            const double old = lo[q];
            const int k = cp[q] & 0x7ffffff;
            pred[k] = (pred[k] + muok * nw) - mu_n * old;
            lam_row[q] = nw;                                     // (lo aliases lam_row for unstaged rows: read before write)
        }
        if (tt == 0) {
            int zeros = 0, masked = 0;
#pragma unroll
            for (int p = 0; p < PT; ++p)
                if (p < rc.P) {
                    int c0 = nmask[p], c1 = 0;
                    if (anyrare) {
#pragma unroll
                        for (int w = 0; w < TW; ++w)
                            if (xci[par][w][2 * PT]) { c0 += xci[par][w][p]; c1 += xci[par][w][PT + p]; }
                    }
                    zeros += c0; masked += nmask[p];
                    rc.sp[n * PMAX + p] = ok ? accp[p] : 0.0;
                    rc.n0p[n * PMAX + p] = ok ? c0 : cntp[p];
                    rc.n1p[n * PMAX + p] = ok ? c1 : 0;
                }
            rc.slam[n] = ok ? tot : 0.0;
            rc.slam2[n] = ok ? se2 : 0.0;
            rc.rownz[n] = ok ? (len - (zeros - masked)) : 0;
        }
        team_sync();                                             // the prediction is consistent before the next neuron reads it
        if (prof) { const long long t = clock64(); g_phase_cycles[23] += t - tp0; g_phase_cycles[19] += 1; g_phase_cycles[24] += len; }
        cur = nxt;
        nxt = after;
    }
}


// ---- shared-memory accessors by 32-bit shared address (LDS / STS instead of generic LD / ST on the chain's critical path)
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory"); }

// isotonic_regression(sr)[-1] for P <= 3 in closed form (pava.py:9-61, unit weights): the pools' sums accumulate in merge
// order exactly as the stack algorithm does -- (a + b) + c when the first two merged first, a + (b + c) otherwise --
// and the strict `>` merge tests compare the same quotients (division by 2 is exact; only the final mean of three needs a
// real division, which `third` supplies: exact IEEE division or a reciprocal multiply, see sweep_chain_fast).
template <bool EXACT>
__device__ __forceinline__ double pava3_last(double a, double b, double c, int P) {
    if (P == 1) return a;
    if (P == 2) return (a > b) ? 0.5 * (a + b) : b;
    if (!EXACT) {                                  // straight-line form of the cases below (selects, no branches)
        const double vab = a + b, vbc = b + c;
        const double resA = (0.5 * vab > c) ? (vab + c) * (1.0 / 3.0) : c;
        const double resB = (b > c) ? ((a > 0.5 * vbc) ? (a + vbc) * (1.0 / 3.0) : 0.5 * vbc) : c;
        return (a > b) ? resA : resB;
    }
    if (a > b) {
        const double v = a + b;
        if (0.5 * v > c) { const double s = v + c; return EXACT ? s / 3.0 : s * (1.0 / 3.0); }
        return c;
    }
    if (b > c) {
        const double v = b + c;
        if (a > 0.5 * v) { const double s = a + v; return EXACT ? s / 3.0 : s * (1.0 / 3.0); }
        return 0.5 * v;
    }
    return c;
}

// The chain sweep for the common configuration (P <= 3 powers, prediction vector and every chain row staged in shared
// memory): same team structure as sweep_chain, restructured for the latency of one step --
// what a single fit is bound by (ncu, profiles/r2c: the general version issues ~900 dependent instructions per step):
//   * the three per-power sums and the sum of squares are reduced by ONE split butterfly (6 shuffles instead of 20): after
//     the xor-16 and xor-8 levels every lane carries one of the four values, lanes 0 / 8 / 16 / 24 end up with the totals;
//   * the exact-0 / exact-1 counts travel as 16-bit fields of two integers through redux.sync (one instruction each);
//   * the accept / reject gate multiplies by the reciprocal trial counts prepared at init and takes PAVA in closed form;
//     when the result is within 1e-9 of the threshold -- 1e6 times the rounding error of that shortcut -- the gate is
//     re-evaluated with the reference's own operations (IEEE divisions), so every decision equals the exact one;
//   * shared memory is addressed as shared memory (LDS / STS), per-row statistics are stored by eight lanes in parallel.
// PSM: the running prediction lives in shared memory (K doubles fit) -- otherwise in global memory (L2), e.g. C5 with
// K = 100 000 trials; the value read in pass 1 is kept for the commit either way (nobody else touches the entry in between).
//
// TWO TEAMS (16-warp variant): consecutive chain rows that share no trial neither read nor write a common entry of the
// prediction, so their steps are independent.  Team t takes the steps i = t, t + 2, ...; chinfo[i].w says whether row i
// shares a trial with row i - 1 (chain_deps).  Before a step the team waits until the other team has finished step
// i - 3 (always: this bounds the run-ahead, so that only ADJACENT steps ever overlap) and, if the rows are dependent,
// step i - 1; progress counters in shared memory, st.release.cta by the team leader after the commit barrier,
// ld.acquire.cta by every waiting lane.  Every entry of the prediction still sees the same updates in the same order:
// bitwise identical to the one-team sweep.  (A team only ever waits for steps that precede its own, and the team that
// owns the oldest unfinished step never waits: no deadlock.)
__device__ __forceinline__ int ld_acquire_cta_s(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];\n" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_s(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __noinline__ void sweep_chain_fast(const Ctx& c, int nchain, double sigma, double thr, double minspk, bool gate, double* pred,
                                 double* stage_base, const bool PSM, const int team, const int nteams, int* done) {
    constexpr int PT = 4;
    const int lane = threadIdx.x & 31, tw = (threadIdx.x >> 5) - team * TW, tt = threadIdx.x - team * TT;
    __shared__ __align__(16) double xch_s[2][2][TW][4];
    __shared__ __align__(16) int xci_s[2][2][TW][4];
    double (*xch)[TW][4] = xch_s[team];
    int (*xci)[TW][4] = xci_s[team];
    stage_base += (size_t)team * STAGE_DOUBLES;
    const int P = c.P;
    double* const g_lam = c.lam; double* const g_sp = c.sp; double* const g_slam = c.slam; double* const g_slam2 = c.slam2;
    int* const g_n0p = c.n0p; int* const g_n1p = c.n1p; int* const g_rownz = c.rownz;
    const int* g_colpw = c.colpw; const double* g_cst = c.cst; const double* g_mu = c.mu; const double* g_rcnt = c.rcnt;
    const int* g_cntp = c.cntp; const int* g_nmask = c.nmask; const int4* g_chinfo = c.chinfo;
    double* scs = stage_base;                                   // [NSTAGE][RC]
    double* slo = scs + NSTAGE * RC;                            // [NSTAGE][RC]
    int* scp = reinterpret_cast<int*>(slo + NSTAGE * RC);       // [NSTAGE][RC]
    double* shd = reinterpret_cast<double*>(scp + NSTAGE * RC); // [NSTAGE][TW][HD]: mu, rcnt[4], then ints cntp[4], nmask[4]
    const uint32_t a_pred = PSM ? smem_u32(pred) : 0u, a_scs = smem_u32(scs), a_slo = smem_u32(slo), a_scp = smem_u32(scp),
                   a_shd = smem_u32(shd), a_xch = smem_u32(&xch[0][0][0]), a_xci = smem_u32(&xci[0][0][0]);
    auto stage = [&](int4 inf, int buf) {
        const int n = inf.x, beg = inf.y, len = inf.z;
        for (int q = tt; q < len; q += TT) {
            cp_async4(scp + buf * RC + q, g_colpw + beg + q);
            cp_async8(scs + buf * RC + q, g_cst + beg + q);
            cp_async8(slo + buf * RC + q, g_lam + beg + q);
        }
        double* hd = shd + (buf * TW + tw) * HD;
        if (lane == 0) cp_async8(hd, g_mu + n);
        else if (lane <= PT) cp_async8(hd + lane, g_rcnt + n * PMAX + (lane - 1));
        else if (lane <= 3 * PT) {
            int* hi = reinterpret_cast<int*>(hd + 1 + PT);
            const int q = lane - PT - 1;
            cp_async4(hi + q, (q < PT) ? (g_cntp + n * PMAX + q) : (g_nmask + n * PMAX + (q - PT)));
        }
        cp_async_commit();
    };
    // the team's own steps are i = team + li * nteams, li = 0, 1, ...
    const int g0 = team, g1 = team + nteams, g2 = team + 2 * nteams, g3 = team + 3 * nteams;
    int4 infA = make_int4(0, 0, 0, 0), infB = infA;
    if (g0 < nchain) stage(g_chinfo[g0], 0);
    if (g1 < nchain) stage(g_chinfo[g1], 1);
    if (g2 < nchain) infA = g_chinfo[g2];
    if (g3 < nchain) infB = g_chinfo[g3];
    int4 cur = g0 < nchain ? g_chinfo[g0] : infA, nxt = g1 < nchain ? g_chinfo[g1] : infA;
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    int li = 0;
    for (int i = team; i < nchain; i += nteams, ++li) {
        if (i + 2 * nteams < nchain) { stage(infA, (li + 2) % NSTAGE); cp_async_wait<2>(); }
        else if (i + nteams < nchain) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncwarp();
        const int4 after = infA;
        infA = infB;
        if (i + 4 * nteams < nchain) infB = g_chinfo[i + 4 * nteams];
        const int buf = li % NSTAGE, par = li & 1;
        if (nteams > 1) {
            // steps of the other team that must be complete: i - 3 always, i - 1 if this row shares a trial with it
            const int m = (i + team - 2) >> 1;                   // the other team's local index of step i - 1
            const int need = cur.w ? m + 1 : m;
            if (need > 0) {
                long long spins = 0;
                while (ld_acquire_cta_s(done + (1 - team)) < need)
                    if (++spins > (1ll << 28)) { atomicExch(c.status, CM_ECHAIN); break; }
            }
        }
        const int n = cur.x, beg = cur.y, len = cur.z;
        const uint32_t a_hd = a_shd + (uint32_t)((buf * TW + tw) * HD) * 8u;
        const double mu_n = lds_f64(a_hd);
        const double coef = sigma * mu_n;
        const uint32_t a_cp = a_scp + (uint32_t)(buf * RC) * 4u, a_cs = a_scs + (uint32_t)(buf * RC) * 8u,
                       a_lo = a_slo + (uint32_t)(buf * RC) * 8u;
        // ---- pass 1 ----
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;           // three per-power sums, sum of squares
        unsigned c0a = 0, c1a = 0, c2 = 0;   // 16-bit fields: #(e==0) of power 0 | 1, #(e==1) of power 0 | 1, #(e==0) | #(e==1) of power 2
        double ev[EPL], pv[EPL];
        uint32_t ka[EPL];
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int q = tt + j * TT;
            ev[j] = 0.0; pv[j] = 0.0; ka[j] = 0;
            if (j * TT < len && q < len) {                        // first test warp-uniform: rows of <= TT entries run j = 0 only
                const int pk = lds_s32(a_cp + (uint32_t)q * 4u);
                const uint32_t kk = (uint32_t)(pk & 0x7ffffff);
                const uint32_t kad = PSM ? a_pred + kk * 8u : kk;
                const int pw = pk >> 27;
                const double pk_pred = PSM ? lds_f64(kad) : pred[kk];
                const double e = sigmoid_fast(lds_f64(a_cs + (uint32_t)q * 8u) - coef * pk_pred);
                ev[j] = e; pv[j] = pk_pred; ka[j] = kad;
                v0 += (pw == 0) ? e : 0.0;
                v1 += (pw == 1) ? e : 0.0;
                v2 += (pw == 2) ? e : 0.0;
                v3 = fma(e, e, v3);
                const unsigned one = (e == 1.0) ? 1u : 0u, zer = (e == 0.0) ? 1u : 0u;
                const unsigned sh = (pw == 1) ? 16u : 0u;
                c0a += (pw < 2) ? (zer << sh) : 0u;
                c1a += (pw < 2) ? (one << sh) : 0u;
                c2 += (pw == 2) ? (zer | (one << 16)) : 0u;
            }
        }
        // ---- split butterfly: 4 values, 6 shuffles ----
        {
            double a = hi16 ? v2 : v0, b = hi16 ? v3 : v1;
            const double sa = hi16 ? v0 : v2, sb = hi16 ? v1 : v3;
            a += __shfl_xor_sync(0xffffffffu, sa, 16);
            b += __shfl_xor_sync(0xffffffffu, sb, 16);
            double x = hi8 ? b : a;
            const double sx = hi8 ? a : b;
            x += __shfl_xor_sync(0xffffffffu, sx, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 2);
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            if ((lane & 7) == 0) sts_f64(a_xch + (uint32_t)(((par * TW + tw) * 4) + (lane >> 3)) * 8u, x);
        }
        c0a = __reduce_add_sync(0xffffffffu, c0a);
        c1a = __reduce_add_sync(0xffffffffu, c1a);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        if (lane == 1) sts_s32(a_xci + (uint32_t)((par * TW + tw) * 4) * 4u, (int)c0a);
        if (lane == 2) sts_s32(a_xci + (uint32_t)((par * TW + tw) * 4 + 1) * 4u, (int)c1a);
        if (lane == 3) sts_s32(a_xci + (uint32_t)((par * TW + tw) * 4 + 2) * 4u, (int)c2);
        team_sync_id(team);
        // ---- totals in fixed warp order, gate ----
        double S0 = 0.0, S1 = 0.0, S2 = 0.0, Q = 0.0;
#pragma unroll
        for (int w = 0; w < TW; ++w) {
            const uint32_t a = a_xch + (uint32_t)((par * TW + w) * 4) * 8u;
            S0 += lds_f64(a); S1 += lds_f64(a + 8u); S2 += lds_f64(a + 16u); Q += lds_f64(a + 24u);
        }
        double tot = S0;
        if (P > 1) tot += S1;
        if (P > 2) tot += S2;
        bool ok = true;
        if (gate) {
            const double r0 = lds_f64(a_hd + 8u), r1 = lds_f64(a_hd + 16u), r2 = lds_f64(a_hd + 24u);
            const double pf = pava3_last<false>(S0 * r0, S1 * r1, S2 * r2, P);
            ok = pf >= thr;
            if (!(fabs(pf - thr) > 1e-9)) {                       // the reference's own operations (caviar.py:183, pava.py:42)
                const int n0 = lds_s32(a_hd + 40u), n1 = lds_s32(a_hd + 44u), n2 = lds_s32(a_hd + 48u);
                const double e0 = S0 / ((double)n0 + 1e-4 * (n0 == 0 ? 1.0 : 0.0));
                const double e1 = S1 / ((double)n1 + 1e-4 * (n1 == 0 ? 1.0 : 0.0));
                const double e2 = S2 / ((double)n2 + 1e-4 * (n2 == 0 ? 1.0 : 0.0));
                ok = pava3_last<true>(e0, e1, e2, P) >= thr;
            }
            ok = ok && (tot >= minspk);
        }
        // ---- commit ----
        const double muok = ok ? mu_n : 0.0;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int q = tt + j * TT;
            if (j * TT < len && q < len) {
                const double nw = ok ? ev[j] : 0.0;
                const double old = lds_f64(a_lo + (uint32_t)q * 8u);
                g_lam[beg + q] = nw;
                const double np_ = (pv[j] + muok * nw) - mu_n * old;
                if (PSM) sts_f64(ka[j], np_); else pred[ka[j]] = np_;
            }
        }
        if (tw == TW - 1 && lane < 8) {
            // per-row statistics for update_phi / update_sigma (caviar.py:238-244,275-287), branch-free: lanes 0-2 own the
            // per-power entries, lane 3 the row sum, lane 4 the sum of squares, lane 5 the non-zero count
            const int4 w0 = *reinterpret_cast<const int4*>(&xci[par][0][0]), w1 = *reinterpret_cast<const int4*>(&xci[par][1][0]);
            const int4 w2 = *reinterpret_cast<const int4*>(&xci[par][2][0]), w3 = *reinterpret_cast<const int4*>(&xci[par][3][0]);
            const int c0s = w0.x + w1.x + w2.x + w3.x, c1s = w0.y + w1.y + w2.y + w3.y, cbs = w0.z + w1.z + w2.z + w3.z;
            const int m0 = lds_s32(a_hd + 56u), m1 = lds_s32(a_hd + 60u), m2 = lds_s32(a_hd + 64u);   // masked trials per power
            const int z0 = m0 + (c0s & 0xffff), z1 = m1 + (c0s >> 16), z2 = m2 + (cbs & 0xffff);   // exact zeros incl. masked
            const int pl = lane < 3 ? lane : 0;
            const int zmine = pl == 0 ? z0 : (pl == 1 ? z1 : z2);
            const int omine = pl == 0 ? (c1s & 0xffff) : (pl == 1 ? (c1s >> 16) : (cbs >> 16));
            const int cntmine = lds_s32(a_hd + 40u + 4u * (uint32_t)pl);
            const int zeros = z0 + (P > 1 ? z1 : 0) + (P > 2 ? z2 : 0), masked = m0 + (P > 1 ? m1 : 0) + (P > 2 ? m2 : 0);
            double dv = pl == 0 ? S0 : (pl == 1 ? S1 : S2);
            dv = lane == 3 ? tot : (lane == 4 ? Q : dv);
            dv = ok ? dv : 0.0;
            double* dptr = lane < 3 ? g_sp + n * PMAX + pl : (lane == 3 ? g_slam + n : g_slam2 + n);
            if (lane < P || lane == 3 || lane == 4) *dptr = dv;
            if (lane < P) {
                g_n0p[n * PMAX + pl] = ok ? zmine : cntmine;
                g_n1p[n * PMAX + pl] = ok ? omine : 0;
            }
            if (lane == 5) g_rownz[n] = ok ? (len - (zeros - masked)) : 0;
        }
        team_sync_id(team);                                       // the prediction is consistent before the next neuron reads it
        if (nteams > 1 && tt == 0) st_release_cta_s(done + team, li + 1);
        cur = nxt;
        nxt = after;
    }
}

// PRNG work for one iteration, done by ONE warp: shuffle sub-keys, the N per-neuron sample keys
// (key, key_next = split(key), caviar.py:209) and the key the next iteration starts from (caviar.py:251,304).
__device__ __noinline__ void rng_iteration(int N, int rounds, uint32_t& k0, uint32_t& k1, uint32_t* keys_out, uint32_t* subkeys) {
    const int lane = threadIdx.x & 31;
    uint32_t p0 = k0, p1 = k1;
    for (int r = 0; r < rounds; ++r) {                 // permutation(key): key, subkey = split(key) per round
        uint32_t a, b, s0, s1;
        warp_split(p0, p1, a, b, s0, s1);
        p0 = a; p1 = b;
        if (lane == 0) { subkeys[2 * r] = s0; subkeys[2 * r + 1] = s1; }
    }
    uint32_t c0 = k0, c1 = k1;
    for (int m = 0; m < N; ++m) {
        uint32_t s0, s1, n0, n1;
        warp_split(c0, c1, s0, s1, n0, n1);
        if (lane == 0) { keys_out[2 * m] = s0; keys_out[2 * m + 1] = s1; }
        c0 = n0; c1 = n1;
    }
    uint32_t a, b, n0, n1;
    warp_split(c0, c1, a, b, n0, n1);                  // update_phi returns split(key)[1]
    k0 = n0; k1 = n1;
}

// ------------------------------------------------------------------------------------------------ a7
__device__ __forceinline__ double group_loglik(double f, double cnt, double S, double n0, double n1) {
    if (cnt == 0.0) return 0.0;
    if (f != f) return 0.0;                                   // nan_to_num(nan) = 0
    if (f == 1.0) return -DBL_MAX * (cnt - n1);               // lam<1 elements are -inf -> -DBL_MAX each
    if (f == 0.0) return -DBL_MAX * (cnt - n0);
    return S * log(f) + (cnt - S) * log(1.0 - f);
}

constexpr int GPL = (PMAX + 1 + 3) / 4;       // power groups per lane of a quad (group 0 = untargeted trials)

struct QuadStats {                             // the groups g = m, m+4, ... owned by member m of the quad
    double pv[GPL], cnt[GPL], S[GPL], n0[GPL], n1[GPL];
    int ng;
};

__device__ __forceinline__ double quad_sum_m(double v, unsigned mask) {      // inside a branch taken by whole quads
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    return v;
}
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// update_phi for the rows list[0..nlist): _laplace_approx (caviar.py:253-308), 10 damped Newton steps from the PRIOR
// mean, covariance = H^-1 before the last step.  One row per quad of lanes (the power groups of a row are spread over
// its 4 lanes), eight rows in flight per warp -- as a STATE MACHINE: every trip of the warp loop evaluates the objective
// AND its gradient / Hessian sums at one point per row, the row's current backtracking candidate.  An accepted candidate
// is the next iterate, and the sums just computed there are that step's gradient pass (same inputs, same operations,
// same fp result as evaluating them again), so a step costs 1 + #backtracks evaluations instead of 2 + #backtracks; rows
// advance independently (a warp no longer loops until the slowest of its eight rows has finished EVERY step: the
// lock-step version spent 5.1 trips per step where the rows needed 1.7) and a quad that finishes a row fetches its next
// one from the CTA's shared ticket counter (which quad runs a row does not change its result).
__device__ __noinline__ void newton_rows(const Ctx& c, const double* powers, const int* list, int nlist, int part, int nparts) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int quad = lane >> 2, mem = lane & 3;
    const double t = 10.0, alpha = 0.25, bbeta = 0.5;
    // Rows are handed out dynamically inside the CTA: positions of the list are dealt to the CTAs of the fit in chunks of
    // NW * 8, a quad that is free takes the CTA's next position from a shared counter (rows need 10 .. 400 evaluations:
    // with a static row list per quad the warps ran at 57 - 70 % of their quads on average).
    constexpr int CH = NW * 8;
    __shared__ int s_next;
    __syncthreads();
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const unsigned qmask = 0xFu << (lane & 28);
    bool exhausted = false;
    bool have = false, first = false;
    int n = 0, step = 0, bt = 0;
    QuadStats s;
    s.ng = 0;
#pragma unroll
    for (int i = 0; i < GPL; ++i) { s.pv[i] = 0.0; s.cnt[i] = 0.0; s.S[i] = 0.0; s.n0[i] = 0.0; s.n1[i] = 0.0; }
    double prior[2] = {1.0, 1.0}, prec[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    double p0 = 1.0, p1 = 1.0, basev = 0.0, v0 = 0.0, v1 = 0.0, Jv = 0.0, stp = 1.0;
    long long evals_w = 0, evals_r = 0;
    for (;;) {
        if (!have && !exhausted) {
            int tkt = 0;
            if (mem == 0) tkt = atomicAdd(&s_next, 1);
            tkt = __shfl_sync(qmask, tkt, lane & 28);
            const int idx = ((tkt / CH) * nparts + part) * CH + (tkt % CH);
            if (idx >= nlist) exhausted = true;                   // positions grow with the ticket: nothing left for this CTA
            else {                                               // fetch the row: sufficient statistics, prior, precision
            n = list[idx];
            double ctot = 0.0;
            for (int p = 0; p < c.P; ++p) ctot += (double)c.cntp[n * PMAX + p];
            s.ng = 0;
#pragma unroll
            for (int i = 0; i < GPL; ++i) {
                const int g = mem + 4 * i;
                s.pv[i] = 0.0; s.cnt[i] = 0.0; s.S[i] = 0.0; s.n0[i] = 0.0; s.n1[i] = 0.0;
                if (g <= c.P) {
                    s.ng = i + 1;
                    if (g == 0) { s.cnt[i] = (double)c.K - ctot; s.n0[i] = s.cnt[i]; }
                    else {
                        const int p = g - 1;
                        s.pv[i] = powers[p];
                        s.cnt[i] = (double)c.cntp[n * PMAX + p];
                        s.S[i] = c.sp[n * PMAX + p];
                        s.n0[i] = (double)c.n0p[n * PMAX + p];
                        s.n1[i] = (double)c.n1p[n * PMAX + p];
                    }
                }
            }
            prior[0] = c.phi0[2 * n]; prior[1] = c.phi0[2 * n + 1];
            const double* cov0 = c.phicov0 + 4 * n;
            const double det0 = cov0[0] * cov0[3] - cov0[1] * cov0[2];
            prec[0] = cov0[3] / det0; prec[1] = -cov0[1] / det0; prec[2] = -cov0[2] / det0; prec[3] = cov0[0] / det0;
            p0 = prior[0]; p1 = prior[1];
            hi[0] = hi[1] = hi[2] = hi[3] = 0.0;
            step = 0; bt = 0; stp = 1.0; v0 = 0.0; v1 = 0.0;
            first = true; have = true;
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // the point this trip evaluates: the iterate itself (first trip of a row) or the current candidate p + stp v
        const double x0 = first ? p0 : p0 + stp * v0, x1 = first ? p1 : p1 + stp * v1;
        double j1 = 0, j2 = 0, h11 = 0, h12 = 0, h22 = 0, ll = 0;
#pragma unroll
        for (int i = 0; i < GPL; ++i)
            if (i < s.ng) {
                const double f = sigmoid_d(x0 * s.pv[i] - x1);
                const double r = s.S[i] - s.cnt[i] * f;
                const double w = s.cnt[i] * f * (1.0 - f);
                j1 -= s.pv[i] * r;
                j2 += r;
                h11 += s.pv[i] * s.pv[i] * w;
                h12 -= s.pv[i] * w;
                h22 += w;
                ll += group_loglik(f, s.cnt[i], s.S[i], s.n0[i], s.n1[i]);
            }
        ll = quad_sum(ll);
        const double d0 = x0 - prior[0], d1 = x1 - prior[1];
        const double quadf = 0.5 * (d0 * (prec[0] * d0 + prec[1] * d1) + d1 * (prec[2] * d0 + prec[3] * d1));
        // log(x0) + log(x1): even members of the quad take log(x0), odd ones log(x1) -- one logarithm per trip, not two
        const double lg = log((mem & 1) ? x1 : x0);
        const double lgo = __shfl_xor_sync(0xffffffffu, lg, 1);
        const double val = -ll - (((mem & 1) ? lgo : lg) + ((mem & 1) ? lg : lgo)) / t + quadf;
        ++evals_w;
        if (have) {
            ++evals_r;
            bool accept = first;
            if (!first) {                                        // Armijo test of the candidate (caviar.py:289-299)
                const double rhs = basev + alpha * stp * Jv;
                const bool go = (bt < 40) && ((val != val) || val > rhs);
                accept = !go;
                if (go) { ++bt; stp *= bbeta; }
            }
            if (accept) {
                if (!first) { p0 = x0; p1 = x1; ++step; }
                first = false;
                if (step == 10) {
                    if (mem == 0) {
                        c.phi[2 * n] = p0; c.phi[2 * n + 1] = p1;
                        for (int q = 0; q < 4; ++q) c.phicov[4 * n + q] = hi[q];
                        if (c.rownz[n] == 0) {                   // all-zero rows always give the same answer: cache it
                            c.phiz[2 * n] = p0; c.phiz[2 * n + 1] = p1;
                            for (int q = 0; q < 4; ++q) c.phicovz[4 * n + q] = hi[q];
                            c.phizok[n] = 1;
                        }
                    }
                    have = false;
                } else {                                         // Newton direction at the new iterate
                    basev = val;
                    // gradient / Hessian sums over the quad (a whole quad is in this branch or not: `accept` derives
                    // from quad-uniform values), then three divisions: 1 / p0, 1 / p1, 1 / det
                    const unsigned qm = 0xFu << (lane & 28);
                    j1 = quad_sum_m(j1, qm); j2 = quad_sum_m(j2, qm); h11 = quad_sum_m(h11, qm);
                    h12 = quad_sum_m(h12, qm); h22 = quad_sum_m(h22, qm);
                    const double rp0 = 1.0 / p0, rp1 = 1.0 / p1, rt = 1.0 / t;
                    const double J0 = j1 + (prec[0] * d0 + prec[1] * d1) - rt * rp0;
                    const double J1 = j2 + (prec[2] * d0 + prec[3] * d1) - rt * rp1;
                    const double H00 = h11 + prec[0] + rt * rp0 * rp0;
                    const double H01 = h12 + prec[1];
                    const double H10 = h12 + prec[2];
                    const double H11 = h22 + prec[3] + rt * rp1 * rp1;
                    const double rdet = 1.0 / (H00 * H11 - H01 * H10);
                    hi[0] = H11 * rdet; hi[1] = -H01 * rdet; hi[2] = -H10 * rdet; hi[3] = H00 * rdet;
                    v0 = -(hi[0] * J0 + hi[1] * J1); v1 = -(hi[2] * J0 + hi[3] * J1);
                    Jv = J0 * v0 + J1 * v1;
                    stp = 1.0;
                    bt = 0;
                }
            }
        }
    }
    if ((g_phase_enable & 1) && blockIdx.x == 0) {               // diagnostics: warp trips vs evaluations the rows needed
        const long long need = __reduce_add_sync(0xffffffffu, (unsigned)((mem == 0) ? evals_r : 0));
        if (lane == 0) { atomicAdd((unsigned long long*)&g_phase_cycles[25], (unsigned long long)evals_w);
                         atomicAdd((unsigned long long*)&g_phase_cycles[26], (unsigned long long)need); }
    }
}

// ------------------------------------------------------------------------------------------------ the fit
template <int PT>
__global__ void __launch_bounds__(NT, (NT == 512) ? 1 : 2) caviar_fit_kernel(const FitParams p) {
    extern __shared__ __align__(16) double dyn_smem[];
    __shared__ double red[NW + 2];
    __shared__ double sc_shape, sc_rate, sc_spont;
    __shared__ int sc_na, sc_flag, sc_focus;
    __shared__ uint32_t sc_key[2];
    __shared__ uint32_t sc_subkeys[2 * MAX_SHUFFLE_ROUNDS];
    __shared__ double sc_powers[PMAX];
    __shared__ long long sc_tlast;
    __shared__ __align__(8) uint64_t sc_bar[2 * NST];

    __shared__ int sc_b;
    __shared__ int sc_done[2];         // steps finished by each team of the two-team chain sweep
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&sc_bar[i], 1); mbar_init(&sc_bar[NST + i], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    GemmPipe gp;
    gp.full = sc_bar; gp.empty = sc_bar + NST; gp.seq = 0;
    // Fits are handed out through a device-side queue when the launch has no helper CTAs (p.queue != nullptr): the grid
    // is one resident wave of CTAs and every CTA pulls the next fit index when it finishes one, so a batch that is not
    // a multiple of the resident wave does not leave SMs idle behind the slowest wave.  Which CTA runs a fit does not
    // change a single bit of its result.
    for (bool first = true;; first = false) {
    int b;
    if (p.queue) {
        __syncthreads();
        if (threadIdx.x == 0) sc_b = atomicAdd(p.queue, 1);
        __syncthreads();
        b = sc_b;
        if (b >= p.B) break;
    } else {
        if (!first) break;
        b = blockIdx.x / p.ct;                        // p.ct CTAs per fit: the fit itself and its panel-GEMM helpers
    }
    char* base = p.ws + (size_t)b * p.L.stride;
    const Layout& L = p.L;
    Ctx c;
    c.N = p.N; c.K = p.K; c.P = p.P; c.it = 0;
#define CM_D(name) c.name = reinterpret_cast<double*>(base + L.name)
#define CM_I(name) c.name = reinterpret_cast<int*>(base + L.name)
    CM_D(X); CM_D(XI); CM_D(Dinv); CM_D(PA); CM_D(PB); CM_D(PP); CM_D(lam); CM_D(cst); CM_D(y); CM_D(ss); CM_D(pred); CM_D(resid); CM_D(z); CM_D(mu);
    CM_D(beta); CM_D(bvec); CM_D(dvec); CM_D(wvec); CM_D(slam); CM_D(slam2); CM_D(sp); CM_D(phibar); CM_D(phi);
    CM_D(phicov); CM_D(phiz); CM_D(phicovz); CM_D(lamhist); CM_D(lamT); CM_D(growbuf); CM_D(rcnt); CM_D(mce);
    CM_I(row_ptr); CM_I(col_ptr); CM_I(col_k); CM_I(csc_row); CM_I(csc_pos); CM_I(cntp); CM_I(n0p); CM_I(n1p);
    CM_I(act); CM_I(ainv); CM_I(order); CM_I(order2); CM_I(pos); CM_I(rownz); CM_I(phizok); CM_I(dcnt); CM_I(dlist); CM_I(colpw); CM_I(nmask);
#undef CM_D
#undef CM_I
    c.chinfo = reinterpret_cast<int4*>(base + L.chinfo);
    c.ccol_ptr = reinterpret_cast<int*>(base + L.ccol_ptr); c.ccsc_row = reinterpret_cast<int*>(base + L.ccsc_row);
    c.ccsc_pos = reinterpret_cast<int*>(base + L.ccsc_pos); c.member = reinterpret_cast<int*>(base + L.member);
    c.rowcb = reinterpret_cast<int2*>(base + L.rowcb);
    c.ucol_ptr = c.col_ptr; c.ucsc_row = c.csc_row; c.ucsc_pos = c.csc_pos;
    c.job = reinterpret_cast<int*>(base + L.job);
    c.ct = p.ct;
    c.role = p.queue ? 0 : blockIdx.x - b * p.ct;      // queue mode has no helper CTAs: whoever pulls a fit runs it
    c.cscq = reinterpret_cast<double2*>(base + L.cscq);
    c.sortkeys = reinterpret_cast<uint32_t*>(base + L.sortkeys);
    c.keys = reinterpret_cast<uint32_t*>(base + L.keys);
    c.pw = reinterpret_cast<unsigned char*>(base + L.pw);
    c.mask = reinterpret_cast<unsigned char*>(base + L.mask);
    c.blocked = reinterpret_cast<unsigned char*>(base + L.blocked);
    c.mu0 = p.mu0 + (size_t)b * p.N;
    c.beta0 = p.beta0 + (size_t)b * p.N;
    c.phi0 = p.phi0 + (size_t)b * p.N * 2;
    c.phicov0 = p.phicov0 + (size_t)b * p.N * 4;
    c.red = red;
    c.tlast = &sc_tlast;
    if (threadIdx.x == 0) sc_tlast = clock64();
    c.sm = dyn_smem;
    c.smd = p.smem_doubles;
    const int N = c.N, K = c.K, P = c.P;
    const cm_caviar_options& o = p.opt;
    c.status = p.status + b;
    if (p.status[b] != 0) continue;                   // prologue reported an error for this fit
    if (HELPERS && c.role > 0) { helper_loop(c, gp); return; }
    c.nnz = c.row_ptr[N];
    c.unnz = c.nnz;
    for (int n = threadIdx.x; n < N; n += NT) c.member[n] = 1;
    build_rowcb(c);
    if (HELPERS && c.ct > 1 && threadIdx.x == 0) { c.job[5] = 0; c.job[6] = c.nnz; }
    const int iters = o.iters;
    const int S = o.num_mc_samples;
    // number of shuffle rounds of jax.random.permutation: ceil(3 ln N / ln(2^32-1))
    int rounds = (int)ceil(3.0 * log((double)max(1, N)) / log(4294967295.0));
    rounds = min(rounds, MAX_SHUFFLE_ROUNDS);
    // a3 shared-memory plan: [pred (K doubles, when it fits)] [row staging of the chain warp]
    const int kpad = (K + 1) & ~1;
    const bool pred_smem = kpad + STAGE_DOUBLES <= c.smd;
    double* pred = pred_smem ? c.sm : c.pred;
    double* stage_base = pred_smem ? c.sm + kpad : c.sm;

    // ---------------- init (caviar.py:28-51) ----------------
    if (threadIdx.x < PMAX) {
        sc_powers[threadIdx.x] = threadIdx.x < P ? p.powers[threadIdx.x] : 0.0;
        if (HELPERS && c.ct > 1) reinterpret_cast<double*>(c.job + 32)[threadIdx.x] = sc_powers[threadIdx.x];   // for the helpers
    }
    if (threadIdx.x == 0) {
        sc_shape = p.shape0_arr ? p.shape0_arr[b] : p.shape0;
        sc_rate = p.rate0_arr ? p.rate0_arr[b] : p.rate0;
        sc_spont = 0.0;
        const unsigned long long seed = p.seeds[b];
        sc_key[0] = (uint32_t)(seed >> 32);
        sc_key[1] = (uint32_t)(seed & 0xffffffffull);
    }
    for (int k = threadIdx.x; k < K; k += NT) {
        c.mask[k] = c.ss[k] > o.y_xcorr_thresh ? 1 : 0;
        c.z[k] = 0.0;
    }
    for (int n = threadIdx.x; n < N; n += NT) {
        c.phi[2 * n] = c.phi0[2 * n]; c.phi[2 * n + 1] = c.phi0[2 * n + 1];
        for (int q = 0; q < 4; ++q) c.phicov[4 * n + q] = c.phicov0[4 * n + q];
        c.phizok[n] = 0;
        c.mu[n] = c.mu0[n]; c.beta[n] = c.beta0[n];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * PMAX; i += NT) {   // reciprocal spike-rate denominators (caviar.py:183) for the fast gate
        const int cnt = c.cntp[i];
        c.rcnt[i] = 1.0 / ((double)cnt + 1e-4 * (cnt == 0 ? 1.0 : 0.0));
    }
    for (int n = threadIdx.x; n < N; n += NT) {       // lam0 = 0.95 [I>0] lam_mask: every indexed entry is unmasked
        const int len = c.row_ptr[n + 1] - c.row_ptr[n];
        c.slam[n] = 0.95 * len; c.slam2[n] = 0.95 * 0.95 * len; c.rownz[n] = len;
    }
    for (int j = threadIdx.x; j < c.nnz; j += NT) c.lam[j] = 0.95;
    __syncthreads();
    for (int i = threadIdx.x; i < c.unnz; i += NT) c.lamT[i] = c.lam[c.ucsc_pos[i]];
    double sumy = 0.0, ysq = 0.0;
    for (int k = threadIdx.x; k < K; k += NT) { const double v = c.y[k]; sumy += v; ysq += v * v; }
    sumy = block_sum(sumy, red);
    ysq = block_sum(ysq, red) + 1e-5;
    // PRNG stream for iteration 0
    uint32_t rk0 = sc_key[0], rk1 = sc_key[1];
    if (wid == NW - 1) rng_iteration(N, rounds, rk0, rk1, c.keys, sc_subkeys);
    __syncthreads();
    phase_mark(c, 15);

    for (int it = 0; it < iters; ++it) {
        c.it = it;
        const double sigma = sc_shape / sc_rate;
        // ================= a2: block_update_mu =================
        phase_a2(c, sigma, &sc_na, gp);
        // ================= a3: update_lam =================
        uint32_t* keys_cur = c.keys + (size_t)(it & 1) * 2 * N;
        uint32_t* keys_nxt = c.keys + (size_t)((it + 1) & 1) * 2 * N;
        // update order: permutation(key, N) by `rounds` stable sorts on fresh 32-bit keys (caviar.py:196)
        for (int n = threadIdx.x; n < N; n += NT) c.order[n] = n;
        __syncthreads();
        for (int r = 0; r < rounds; ++r) {
            const uint32_t s0 = sc_subkeys[2 * r], s1 = sc_subkeys[2 * r + 1];
            const int h = (N + 1) / 2;
            // stable sort of the current order by fresh 32-bit keys = sort of the unique 64-bit composites (key, position)
            int npow = 1;
            while (npow < N) npow <<= 1;
            if (npow <= c.smd) {                                // bitonic network in shared memory (C3: 1024, C5: 8192 entries)
                unsigned long long* sk = reinterpret_cast<unsigned long long*>(c.sm);
                for (int q = threadIdx.x; q < h; q += NT) {
                    uint32_t x0 = (uint32_t)q, x1 = (h + q < N) ? (uint32_t)(h + q) : 0u;
                    threefry2x32(s0, s1, x0, x1);
                    sk[q] = ((unsigned long long)x0 << 32) | (uint32_t)q;
                    if (h + q < N) sk[h + q] = ((unsigned long long)x1 << 32) | (uint32_t)(h + q);
                }
                for (int q = N + threadIdx.x; q < npow; q += NT) sk[q] = ~0ull;
                __syncthreads();
                for (int kk = 2; kk <= npow; kk <<= 1)
                    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                        for (int q = threadIdx.x; q < npow; q += NT) {
                            const int partner = q ^ jj;
                            if (partner > q) {
                                const unsigned long long a = sk[q], b = sk[partner];
                                const bool up = (q & kk) == 0;
                                if ((a > b) == up) { sk[q] = b; sk[partner] = a; }
                            }
                        }
                        __syncthreads();
                    }
                for (int q = threadIdx.x; q < N; q += NT) c.order2[q] = c.order[(int)(sk[q] & 0xffffffffull)];
                __syncthreads();
            } else {
                for (int q = threadIdx.x; q < h; q += NT) {
                    uint32_t x0 = (uint32_t)q, x1 = (h + q < N) ? (uint32_t)(h + q) : 0u;
                    threefry2x32(s0, s1, x0, x1);
                    c.sortkeys[q] = x0;
                    if (h + q < N) c.sortkeys[h + q] = x1;
                }
                __syncthreads();
                for (int i = threadIdx.x; i < N; i += NT) {
                    const uint32_t ki = c.sortkeys[i];
                    int rank = 0;
                    for (int j = 0; j < N; ++j) {
                        const uint32_t kj = c.sortkeys[j];
                        rank += (kj < ki) || (kj == ki && j < i);
                    }
                    c.order2[rank] = c.order[i];
                }
                __syncthreads();
            }
            for (int n = threadIdx.x; n < N; n += NT) c.order[n] = c.order2[n];
            __syncthreads();
        }
        for (int m = threadIdx.x; m < N; m += NT) c.pos[c.order[m]] = m;
        __syncthreads();
        phase_mark(c, 7);
        // Rows that are all-zero with mu == 0 and cannot reach minimum_spike_count are rejected again whatever the
        // samples are: est_k <= sigmoid((m0 + 8.3 s0) Imax - sigma beta^2 / 2) because the truncated-normal samples
        // obey phi_0 <= m0 + ndtri(1 - 2^-53) s0 and phi_1 >= 0.  Their row stays zero -> no work at all (exact).
        {
            const bool gate_on = it > o.delay_spont_est;
            const double imax = sc_powers[P - 1];
            for (int n = threadIdx.x; n < N; n += NT) {
                int skip = 0;
                if (gate_on && c.rownz[n] == 0 && c.mu[n] == 0.0) {
                    const double m0 = c.phi[2 * n], s0 = c.phicov[4 * n], be = c.beta[n];
                    const double len = (double)(c.row_ptr[n + 1] - c.row_ptr[n]);
                    if (s0 > 0.0 && m0 == m0) {
                        const double xmax = (m0 + 8.3 * s0) * imax - 0.5 * sigma * be * be;
                        const double ub = len * sigmoid_d(xmax) * (1.0 + 1e-9);
                        skip = (ub < o.minimum_spike_count) ? 1 : 0;
                    }
                    if (len == 0.0) skip = 1;
                }
                c.dcnt[n] = skip;
            }
        }
        __syncthreads();
        // Monte-Carlo means of the truncated-normal sigmoid coefficients (caviar.py:209-215, App. A.2)
        if (HELPERS && c.ct > 1 && N >= 512) {
            post_job(c, 4, it & 1, S, 0);
            mc_means(c, keys_cur, S, sc_powers, 0, c.ct);
            wait_helpers(c);
        } else {
            mc_means(c, keys_cur, S, sc_powers, 0, 1);
        }
        __syncthreads();
        phase_mark(c, 8);
        if (dist_passes(c)) {
            if (threadIdx.x == 0) *reinterpret_cast<double*>(c.job + 8) = sigma;
            post_job(c, 14, 0, 0, 0);
            job_pred_cst(c, sigma, 0, c.ct);
            wait_helpers(c);
            if (pred_smem)
                for (int k = threadIdx.x; k < K; k += NT) pred[k] = c.pred[k];
        } else {
        compute_pred(c, pred);
        __syncthreads();
        // per-entry constant part of the sigmoid argument
        for (int n = wid; n < N; n += NW)
            if (!c.dcnt[n]) row_cst(c, n, sigma);
        }
        __syncthreads();
        phase_mark(c, 9);
        // chain = neurons with mu != 0 in update order; (n, row begin, row length) table for the prefetching warp
        const int nchain = block_compact(N, [&](int m) { return c.mu[c.order[m]] != 0.0; }, c.dlist, nullptr, red);
        if (threadIdx.x == 0) sc_flag = 0;
        __syncthreads();
        {
            int mx = 0;
            for (int i = threadIdx.x; i < nchain; i += NT) {
                const int n = c.order[c.dlist[i]];
                const int len = c.row_ptr[n + 1] - c.row_ptr[n];
                c.chinfo[i] = make_int4(n, c.row_ptr[n], len, 0);
                mx = max(mx, len);
            }
            if (mx > RC) atomicOr(&sc_flag, 1);            // a chain row exceeds the staging capacity -> general sweep
        }
        __syncthreads();
        const bool fast_chain = (PT == 4) && P <= 3 && sc_flag == 0;
        // two chain teams (16-warp variant): rows without a common trial overlap, see sweep_chain_fast
        const bool two_teams = HELPERS && fast_chain && nchain >= 16 && !(g_phase_enable & 2048) &&
                               (pred_smem ? kpad : 0) + 2 * STAGE_DOUBLES <= c.smd;
        if (two_teams) {
            if (threadIdx.x == 0) { sc_done[0] = 0; sc_done[1] = 0; }
            if (c.ct > 1 && nchain >= 128) {
                post_job(c, 17, nchain, 0, 0);
                chain_deps(c, nchain, 0, c.ct);
                wait_helpers(c);
            } else {
                chain_deps(c, nchain, 0, 1);
                __syncthreads();
            }
        }
        const int nteams = two_teams ? 2 : 1;
        {
            const double thr = o.msrmp + sc_spont;
            const bool gate = it > o.delay_spont_est;
            const long long role_t0 = clock64();
            if (wid < TW * nteams) {
                if (fast_chain) sweep_chain_fast(c, nchain, sigma, thr, o.minimum_spike_count, gate, pred, stage_base, pred_smem,
                                                 wid / TW, nteams, sc_done);
                else sweep_chain<PT>(c, nchain, sigma, thr, o.minimum_spike_count, gate, pred, stage_base);
                if (g_phase_enable && blockIdx.x == 0 && threadIdx.x == 0) g_phase_cycles[16] += clock64() - role_t0;
            } else if (wid == NW - 1) {
                if (it + 1 < iters) rng_iteration(N, rounds, rk0, rk1, keys_nxt, sc_subkeys);
                if (g_phase_enable && blockIdx.x == 0 && lane == 0) g_phase_cycles[17] += clock64() - role_t0;
            } else {
                // mu == 0: the row neither reads nor changes the prediction -> order-free, run concurrently
                for (int m = wid - TW * nteams; m < N; m += NW - TW * nteams - 1) {
                    const int n = c.order[m];
                    if (c.mu[n] == 0.0 && !c.dcnt[n]) {
                        const int beg = c.row_ptr[n], len = c.row_ptr[n + 1] - beg;
                        sweep_row<PT>(make_rowctx(c), n, beg, len, false, 0.0, c.cntp + n * PMAX, c.nmask + n * PMAX,
                                      c.colpw + beg, c.cst + beg, c.lam + beg, sigma, thr, o.minimum_spike_count, gate, pred);
                    }
                }
                if (g_phase_enable && blockIdx.x == 0 && wid == TW * nteams && lane == 0) g_phase_cycles[18] += clock64() - role_t0;
            }
        }
        __syncthreads();
        // Pruned rows (posterior identically zero) contribute exact zeros to every by-trial pass -- the prediction, the Gram
        // expansion, the spontaneous-event mask -- so those passes may skip them: once the live rows hold <= 70 % of the
        // entries of the index in use, it is rebuilt over the live rows only (and again from the full index should a row
        // outside it come back).  Bitwise neutral: only +-0.0 terms leave the sums.
        {
            int alive_entries = 0, outside = 0;
            for (int n = threadIdx.x; n < N; n += NT)
                if (c.rownz[n] > 0) {
                    alive_entries += c.row_ptr[n + 1] - c.row_ptr[n];
                    outside |= (c.member[n] == 0);
                }
            alive_entries = block_sum_int(alive_entries, red);
            outside = block_sum_int(outside, red);
            if (!(g_phase_enable & 128) && (outside > 0 || (long long)alive_entries * 10 <= (long long)c.unnz * 7)) {
                compact_csc(c, red);
                if (HELPERS && c.ct > 1 && threadIdx.x == 0) { c.job[5] = 1; c.job[6] = c.unnz; }   // helpers switch index too
            }
        }
        // by-trial copy of the new lam; with helpers also the residual of a6 and the spontaneous-event mask of a8
        // (one fused pass also without helpers: the by-trial lists are walked once instead of three times; bit 10 of the
        // diagnostics word switches the separate passes back on)
        const bool dist15 = !(g_phase_enable & 1024);
        if (dist15) {
            if (dist_passes(c)) {
                if (threadIdx.x == 0) *reinterpret_cast<double*>(c.job + 8) = o.spont_orthogonality;
                post_job(c, 15, 0, 0, 0);
                job_lamT_resid(c, o.spont_orthogonality, 0, c.ct);
                wait_helpers(c);
            } else {
                job_lamT_resid(c, o.spont_orthogonality, 0, 1);
            }
        } else {
#pragma unroll 8
            for (int i = threadIdx.x; i < c.unnz; i += NT) c.lamT[i] = c.lam[c.ucsc_pos[i]];
        }
        __syncthreads();
        phase_mark(c, 10);
        // ================= a6: update_sigma (caviar.py:238-244), with a2's mu =================
        if (!dist15) {
            compute_pred(c, pred);
            __syncthreads();
        }
        {
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int k = threadIdx.x; k < K; k += NT) {
                double r;
                if (dist15) r = c.resid[k];
                else { r = c.y[k] - pred[k]; c.resid[k] = r; }
                s1 += r * r;
            }
            for (int n = threadIdx.x; n < N; n += NT) {
                const double m = c.mu[n], be = c.beta[n];
                s2 += m * m * c.slam2[n];
                s3 += (m * m + be * be) * c.slam[n];
            }
            s1 = block_sum(s1, red); s2 = block_sum(s2, red); s3 = block_sum(s3, red);
            if (threadIdx.x == 0) {
                sc_shape = (p.shape0_arr ? p.shape0_arr[b] : p.shape0) + (double)K / 2.0;
                sc_rate = (p.rate0_arr ? p.rate0_arr[b] : p.rate0) + 0.5 * (s1 - s2 + s3);
            }
        }
        __syncthreads();
        phase_mark(c, 11);
        // ================= a7: update_phi (caviar.py:246-310) =================
        {
            const int nl = block_compact(N, [&](int n) { return !(c.rownz[n] == 0 && c.phizok[n]); }, c.dlist, nullptr, red);
            for (int n = threadIdx.x; n < N; n += NT)
                if (c.rownz[n] == 0 && c.phizok[n]) {
                    c.phi[2 * n] = c.phiz[2 * n]; c.phi[2 * n + 1] = c.phiz[2 * n + 1];
                    for (int q = 0; q < 4; ++q) c.phicov[4 * n + q] = c.phicovz[4 * n + q];
                }
            if (nl > 0) {
                if (HELPERS && c.ct > 1 && nl >= 256) {
                    post_job(c, 3, nl, 0, 0);
                    newton_rows(c, sc_powers, c.dlist, nl, 0, c.ct);
                    wait_helpers(c);
                } else {
                    newton_rows(c, sc_powers, c.dlist, nl, 0, 1);
                }
            }
        }
        __syncthreads();
        phase_mark(c, 12);
        // ================= a8: estimate_spont_act_soft_thresh (caviar.py:146-163, 86-88) =================
        if (!dist15)
            for (int k = threadIdx.x; k < K; k += NT) {
                unsigned char bl = 0;
                for (int i = c.ucol_ptr[k]; i < c.ucol_ptr[k + 1]; ++i) bl |= (c.lamT[i] >= o.spont_orthogonality);
                c.blocked[k] = bl;
            }
        __syncthreads();
        {
            double err = sumy, pen = o.penalty;
            int j = it;
            while (j < o.max_backtrack_iters && err > o.tol) {
                double e = 0.0;
                for (int k = threadIdx.x; k < K; k += NT) {
                    const double r = c.resid[k];
                    double zz = (r < pen) ? 0.0 : r - pen;
                    zz = zz < 0.0 ? 0.0 : zz;
                    if (c.blocked[k]) zz = 0.0;
                    zz *= (double)c.mask[k];
                    c.z[k] = zz;
                    const double d = r - zz;
                    e += d * d;
                }
                err = block_sum(e, red) / ysq;
                ++j;
                pen *= o.scale_factor;
            }
            int nzz = 0;
            for (int k = threadIdx.x; k < K; k += NT) nzz += (c.z[k] != 0.0);
            nzz = block_sum_int(nzz, red);
            if (threadIdx.x == 0) sc_spont = (double)nzz / (double)K;
        }
        __syncthreads();
        phase_mark(c, 13);
        // ================= histories (caviar.py:90-92) =================
        if (o.save_histories) {
            const size_t hb = (size_t)b * iters + it;
            if (p.mu_hist) for (int n = threadIdx.x; n < N; n += NT) p.mu_hist[hb * N + n] = c.mu[n];
            if (p.beta_hist) for (int n = threadIdx.x; n < N; n += NT) p.beta_hist[hb * N + n] = c.beta[n];
            if (p.phi_hist) for (int n = threadIdx.x; n < 2 * N; n += NT) p.phi_hist[hb * 2 * N + n] = c.phi[n];
            if (p.phicov_hist) for (int n = threadIdx.x; n < 4 * N; n += NT) p.phicov_hist[hb * 4 * N + n] = c.phicov[n];
            if (p.z_hist) for (int k = threadIdx.x; k < K; k += NT) p.z_hist[hb * K + k] = c.z[k];
            if (threadIdx.x == 0) {
                if (p.shape_hist) p.shape_hist[hb] = sc_shape;
                if (p.rate_hist) p.rate_hist[hb] = sc_rate;
            }
            if (p.lamhist) for (int j = threadIdx.x; j < c.nnz; j += NT) c.lamhist[(size_t)it * c.nnz + j] = c.lam[j];
            __syncthreads();
        }
    }

    // ================= a9: reconnect_spont_cells (caviar.py:102-144) + final update_phi (caviar.py:98) =================
    if (o.fn_scan) {
        // disconnected cells in ascending order; per-cell count of spontaneous events on its stimulated trials
        const int nd = block_compact(N, [&](int i) { return c.mu[i] == 0.0; }, c.dlist, nullptr, red);
        int* alive = c.order2;               // 1 while the cell is still a candidate
        for (int i = threadIdx.x; i < nd; i += NT) alive[i] = 1;
        for (int n = threadIdx.x; n < N; n += NT) c.pos[n] = 0;      // "row changed" flags for the final update_phi
        __syncthreads();
        int remaining = nd;
        bool recount = true;
        int nzz = 0;
        while (remaining > 0) {
            if (recount) {                                   // #(z != 0) only changes when a cell is reconnected
                nzz = 0;
                for (int k = threadIdx.x; k < K; k += NT) nzz += (c.z[k] != 0.0);
                nzz = block_sum_int(nzz, red);
            }
            if (!((double)nzz > o.minimum_spike_count)) break;
            if (recount) {
                for (int i = wid; i < nd; i += NW) {
                    if (!alive[i]) continue;
                    const int n = c.dlist[i];
                    int cnt = 0;
                    for (int j = c.row_ptr[n] + lane; j < c.row_ptr[n + 1]; j += 32) cnt += (c.z[c.col_k[j]] != 0.0);
                    cnt = warp_sum(cnt);
                    if (lane == 0) c.dcnt[i] = cnt;
                }
                recount = false;
                __syncthreads();
            }
            // focus = first maximum of the counts among remaining cells (np.argmax, caviar.py:117)
            long long best = -1;
            for (int i = threadIdx.x; i < nd; i += NT)
                if (alive[i]) {
                    const long long key = ((long long)c.dcnt[i] << 32) | (unsigned)(0x7fffffff - i);
                    best = key > best ? key : best;
                }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const long long oth = __shfl_xor_sync(0xffffffffu, best, off);
                best = oth > best ? oth : best;
            }
            long long* redl = reinterpret_cast<long long*>(red);
            __syncthreads();
            if (lane == 0) redl[wid] = best;
            __syncthreads();
            if (threadIdx.x == 0) {
                long long bb = -1;
                for (int w = 0; w < NW; ++w) bb = redl[w] > bb ? redl[w] : bb;
                sc_focus = 0x7fffffff - (int)(bb & 0xffffffffll);
                sc_flag = 0;
            }
            __syncthreads();
            const int fi = sc_focus;
            const int focus = c.dlist[fi];
            const int maxcnt = c.dcnt[fi];
            if ((double)maxcnt < o.minimum_spike_count) break;        // no remaining cell can pass (exact shortcut)
            if (wid == 0) {
                int cz[PT];
#pragma unroll
                for (int q = 0; q < PT; ++q) cz[q] = 0;
                for (int j = c.row_ptr[focus] + lane; j < c.row_ptr[focus + 1]; j += 32) {
                    const bool nzv = c.z[c.col_k[j]] != 0.0;
                    const int pw = c.pw[j];
#pragma unroll
                    for (int q = 0; q < PT; ++q) cz[q] += (pw == q && nzv);
                }
                double sr[PMAX];
                int spike_count = 0;
                for (int pp = 0; pp < P; ++pp) {
                    int v = 0;
#pragma unroll
                    for (int q = 0; q < PT; ++q) if (q == pp) v = cz[q];
                    v = warp_sum(v);
                    const int cnt = c.cntp[focus * PMAX + pp];
                    sr[pp] = cnt > 0 ? (double)v / (double)cnt : 0.0;
                    spike_count += v;
                }
                const double pv = pava_last(sr, P);
                if (pv >= o.msrmp && (double)spike_count >= o.minimum_spike_count) {
                    // mu = mean(z[locs]), beta = sem(z[locs]) (ddof=1), lam[focus, locs] = 1, z[locs] = 0
                    double s = 0.0;
                    for (int j = c.row_ptr[focus] + lane; j < c.row_ptr[focus + 1]; j += 32) s += c.z[c.col_k[j]];
                    s = warp_sum(s);
                    const double mean = s / (double)spike_count;
                    double q2 = 0.0;
                    for (int j = c.row_ptr[focus] + lane; j < c.row_ptr[focus + 1]; j += 32) {
                        const double zv = c.z[c.col_k[j]];
                        if (zv != 0.0) q2 += (zv - mean) * (zv - mean);
                    }
                    q2 = warp_sum(q2);
                    __syncwarp();
                    for (int j = c.row_ptr[focus] + lane; j < c.row_ptr[focus + 1]; j += 32) {
                        const int k = c.col_k[j];
                        if (c.z[k] != 0.0) {
                            c.lam[j] = 1.0;
                            c.z[k] = 0.0;
                        }
                    }
                    if (lane == 0) {
                        c.mu[focus] = mean;
                        c.beta[focus] = spike_count > 1 ? sqrt(q2 / (double)(spike_count - 1)) / sqrt((double)spike_count)
                                                        : __longlong_as_double(0x7ff8000000000000ll);
                        c.pos[focus] = 1;
                        sc_flag = 1;
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) alive[fi] = 0;
            if (sc_flag) recount = true;
            --remaining;
            __syncthreads();
        }
        __syncthreads();
        // rows touched by reconnection: exact statistics from the row, then the Laplace/Newton update
        for (int n = wid; n < N; n += NW) {
            if (!c.pos[n]) continue;
            double accp[PT]; int c0[PT], c1[PT];
#pragma unroll
            for (int q = 0; q < PT; ++q) { accp[q] = 0.0; c0[q] = 0; c1[q] = 0; }
            int nz = 0;
            for (int j = c.row_ptr[n] + lane; j < c.row_ptr[n + 1]; j += 32) {
                const double l = c.lam[j];
                const int pw = c.pw[j];
                nz += (l != 0.0);
#pragma unroll
                for (int q = 0; q < PT; ++q) {
                    const bool m = pw == q;
                    accp[q] += m ? l : 0.0; c0[q] += (m && l == 0.0); c1[q] += (m && l == 1.0);
                }
            }
            nz = warp_sum(nz);
            for (int pp = 0; pp < P; ++pp) {
                double s = 0.0; int a0 = 0, a1 = 0;
#pragma unroll
                for (int q = 0; q < PT; ++q) if (q == pp) { s = accp[q]; a0 = c0[q]; a1 = c1[q]; }
                s = warp_sum(s); a0 = warp_sum(a0); a1 = warp_sum(a1);
                if (lane == 0) { c.sp[n * PMAX + pp] = s; c.n0p[n * PMAX + pp] = a0 + c.nmask[n * PMAX + pp]; c.n1p[n * PMAX + pp] = a1; }
            }
            if (lane == 0) c.rownz[n] = nz;
        }
        __syncthreads();
        {
            const int nl = block_compact(N, [&](int n) { return c.pos[n] != 0; }, c.dlist, nullptr, red);
            if (nl > 0) {
                if (HELPERS && c.ct > 1 && nl >= 256) {
                    post_job(c, 3, nl, 0, 0);
                    newton_rows(c, sc_powers, c.dlist, nl, 0, c.ct);
                    wait_helpers(c);
                } else {
                    newton_rows(c, sc_powers, c.dlist, nl, 0, 1);
                }
            }
        }
        __syncthreads();
    }

    phase_mark(c, 14);
    // ---------------- outputs ----------------
    for (int n = threadIdx.x; n < N; n += NT) {
        p.mu_out[(size_t)b * N + n] = c.mu[n];
        p.beta_out[(size_t)b * N + n] = c.beta[n];
    }
    for (int n = threadIdx.x; n < 2 * N; n += NT) p.phi_out[(size_t)b * 2 * N + n] = c.phi[n];
    for (int n = threadIdx.x; n < 4 * N; n += NT) p.phicov_out[(size_t)b * 4 * N + n] = c.phicov[n];
    for (int k = threadIdx.x; k < K; k += NT) p.z_out[(size_t)b * K + k] = c.z[k];
    if (threadIdx.x == 0) { p.shape_out[b] = sc_shape; p.rate_out[b] = sc_rate; }
    if (HELPERS && c.ct > 1 && threadIdx.x == 0) {               // release the helpers
        if (c.job[20]) p.status[b] = CM_EHELPER;      // a helper did not answer: results are not to be trusted
        c.job[1] = 0;
        __threadfence();
        st_release_gpu(&c.job[0], c.job[0] + 1);
    }
    }   // fit queue
}

}  // namespace CM_FITNS
