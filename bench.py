#!/usr/bin/env python
"""bench.py -- the model-X Gaussian knockoff hot path on B200 (BASELINE.json metric, configs[2]).

A "step" is one pass of the hot path over one batch of synthetic input, for BOTH methods BASELINE configs[2] names:
    for method in (:mvr, :sdp_ccd):   s = solve_s(Sigma, method; m)  ->  conditional factor  ->  X~ = condition(X, mu, Sigma, diag(s); m)
on block-equicorrelated Sigma (blocks of 10, rho = 0.5, gamma = 0.25), p = 5000, n = 100000 rows IN TOTAL, m = 5.
`value` = knockoff rows/s = (number of methods) * n / step time, inputs resident in HBM.
`e2e`   = the same through the host-pointer C ABI (kb_solve_s + kb_condition) with pinned HOST buffers: H2D of X and
          Sigma, D2H of X~ inside the timed region.

Multi-GPU (one process per GPU, NCCL): STRONG scaling by default -- n is fixed, rows are sharded n/N per rank, the
independent solves (:mvr || :sdp_ccd) run on different ranks (a single CCD solve is sequential: "replicas only"),
s and the factor blocks are BROADCAST over NVLink (NCCL), sampling needs no collective (counter-based RNG keyed by
the global row).  The long :sdp_ccd solve is the Amdahl term, so the step is STAGGERED: the ranks that do not run it sample
as many of the :mvr rows as fit beside it (knockoffs.jl_b200/dist.py: stagger_fraction / stagger_layout), the rest is shared
by all ranks afterwards; from 4 ranks on the m block columns of the factor are built on different ranks from the closed form
of the recursion (kb_cond_factor_block_dev).  `scaling_alt` reports the weak-scaling run (n rows per GPU) next to it.  Three more objects ride
on the same JSON line: `config5` (row-sharded dosage Gram + ONE NCCL all-reduce of the p x p sums + :maxent +
sampling), `config4` (group :mvr, LD blocks dealt to ranks, the reduce hook over NCCL) and the CPU baseline.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "knockoff rows/s (solve_s + conditional factor + sampling, :mvr and :sdp_ccd, p=5000, m=5)"
UNIT = "rows/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
# ncu --set full capture (profiles/); None until a capture of this exact configuration exists.
NCU_TRAFFIC_BYTES = {
    # profiles/r01_ccd_stream_kernel_ncu_raw.csv: 375.47 GB read + 494.29 GB written per sweep launch (p = 5000)
    "ccd_stream_kernel": 869.76e9,
    # profiles/r02_dgemm_dmma_kernel_batched_sampling_ncu_raw.csv: ONE batched launch = all m = 5 copies of one 4096-row
    # chunk (gridDim.z = 5, 25 280 CTAs; per copy a lower- and an upper-triangular K = 5000 segment): 7.347 GB read +
    # 0.801 GB written in 27.19 ms, DMMA pipe 93.4 % active (algorithmic: ~3.4 GB of operands + 0.82 GB of C)
    "dgemm_dmma_kernel": 8.149e9,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--p", type=int, default=5000)
    ap.add_argument("--n", type=int, default=100000, help="rows in total (strong scaling) / per GPU (weak)")
    ap.add_argument("--m", type=int, default=5)
    ap.add_argument("--methods", default="mvr,sdp_ccd")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stagger", action="store_true", help="N > 1: plain order (both solves, then all sampling)")
    ap.add_argument("--no-split-factor", action="store_true", help="N >= 4: factor on the solver rank and broadcast all blocks (round-2 first form)")
    ap.add_argument("--c4-cold", action="store_true", help="config 4: time the first full-size solve (workspace allocation included)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-host-memory leg of the end-to-end measurement")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="skip the separately timed streaming-kernel sweep")
    ap.add_argument("--no-extra", action="store_true", help="skip the config4 / config5 / weak-scaling sections")
    ap.add_argument("--chunk", type=int, default=0, help="rows per sampling chunk (0 = library default)")
    ap.add_argument("--refresh-every", type=int, default=-1)
    ap.add_argument("--c5-rows", type=int, default=500000)
    ap.add_argument("--pipeline", action="store_true",
                    help="1 GPU experiment: sample :mvr on a background (SM-partitioned) context beside the :sdp_ccd solve instead of "
                         "running the stages one after the other (measured SLOWER on B200, profiles/r02_pipeline_experiment.md)")
    return ap.parse_args()


def methods_of(a):
    return [x for x in a.methods.split(",") if x]


def config_dict(a, world):
    """The workload description -- IDENTICAL in both arms (no run-dependent keys)."""
    return {"workload": (f"BASELINE configs[2]: block-equicorrelated Sigma (blocks of 10, rho=0.5, gamma=0.25), p={a.p}, "
                         f"n={a.n} rows, m={a.m}, methods :{' and :'.join(methods_of(a))}"),
            "p": a.p, "n": a.n, "m": a.m, "methods": methods_of(a), "n_is": "total rows" if a.scaling == "strong" else "rows per GPU",
            "l2": "inputs larger than L2 (X 4 GB, 2 GB of factor blocks per method, 200 MB resident inverse rewritten by every "
                  "rank-256 pass)"}


# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.t = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle (port of the reference's algorithm) on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------------
def host_threads():
    """Use every host core for the BLAS stages, also under torchrun (which exports OMP_NUM_THREADS=1)."""
    n = os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(n)                      # for BLAS copies that are loaded after this point
    try:
        import scipy.linalg  # noqa: F401  (scipy ships its own OpenBLAS: it must be loaded before the limits are raised)
        import scipy.linalg.blas  # noqa: F401
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def golden_sweeps(method, p, m):
    """Sweep counts of the oracle's own FULL solves of this workload (committed fixtures, tests/golden/config_c3_*.npz)."""
    if p == 5000 and m == 5:
        path = os.path.join(ROOT, "tests", "golden", f"config_c3_{method}.npz")
        if os.path.exists(path):
            z = np.load(path)
            return int(json.loads(bytes(z["meta"]).decode())["sweeps"])
    return {"mvr": 5, "maxent": 5, "sdp_ccd": 59}[method]


def ccd_stride(steps):
    """Every k-th coordinate of one whole sweep is timed; k grows with --steps so that the whole reference run stays within
    a few minutes (one full p = 5000 sweep is ~75 s of one core per method)."""
    return int(min(160, max(16, 8 * steps)))


SAMPLE_ROWS = 4096       # one full chunk of rows for the sampling GEMMs


def cpu_reference_step(a, Sigma, sweeps, stride):
    """Times a bounded sample of ONE step (all methods) on the CPU and extrapolates to the full workload.
    Returns (rows_per_s, detail dict).  What is measured and how it is scaled:
      * solve_equi (:mvr): eigvalsh of a leading 2500-block x (p/2500)^3                         (utilities.jl:108)
      * CCD: every stride-th coordinate of ONE WHOLE sweep (a coordinate's cost depends on its position only),
        x stride x the sweep count of the oracle's own full solve; one thread, like the reference's BLAS-2 sweep
      * inv(Sigma): in full, all threads                                                           (modelX.jl:106)
      * the pm x pm Cholesky of modelX.jl:116-120: dpotrf of a 2p x 2p leading block x (m/2)^3, all threads
      * sampling on one 4096-row chunk x n/4096: mean GEMM 2 rows p^2, and randn(rows, mp) * L with L LOWER TRIANGULAR
        (BLAS trmm, modelX.jl:120-121): dtrmm against the 2p x 2p factor x (m/2)^2."""
    from oracle import c_oracle
    import scipy.linalg as sla
    from scipy.linalg import blas
    p, m, n = a.p, a.m, a.n
    kap = (m + 1) / m
    det = {}
    rng = np.random.default_rng(0)
    pe = min(p, 2500)
    t = time.perf_counter()
    lam_e = np.linalg.eigvalsh(Sigma[:pe, :pe]).min()
    t_equi = (time.perf_counter() - t) * (p / pe) ** 3
    s0 = np.full(p, min(1.0, kap * lam_e))       # block family: lambda_min is p-independent
    t = time.perf_counter()
    Sinv = np.linalg.inv(Sigma)
    t_inv = time.perf_counter() - t
    total = 0.0
    for method in methods_of(a):
        d = {}
        r = c_oracle.solve_ccd(Sigma, method, m=m, s_init=s0, max_sweeps=1, coord_stride=stride, lapack_init=True)
        if r["rc"] != 0:
            raise RuntimeError(f"oracle CCD failed rc={r['rc']}")
        d["equi_s"] = t_equi if method != "sdp_ccd" else 0.0
        d["ccd_s_per_sweep"] = r["loop_seconds"] * stride
        d["sweeps"] = sweeps[method]
        d["ccd_s"] = d["ccd_s_per_sweep"] * sweeps[method]
        d["inv_s"] = t_inv
        s_vec = np.full(p, 0.5 * s0[0])
        C = 2 * np.diag(s_vec) - (s_vec[:, None] * Sinv) * s_vec[None, :]
        mm = min(m, 2)
        St = np.empty((mm * p, mm * p))
        for k in range(mm):
            for l in range(mm):
                St[k * p:(k + 1) * p, l * p:(l + 1) * p] = C if k == l else C - np.diag(s_vec)
        t = time.perf_counter()
        L = sla.cholesky(St, lower=True, overwrite_a=True, check_finite=False)
        d["chol_s"] = (time.perf_counter() - t) * (m / mm) ** 3
        rows = min(SAMPLE_ROWS, n)
        Xs = rng.standard_normal((rows, p))
        SiS = Sinv * s_vec[None, :]
        t = time.perf_counter()
        mui = Xs - Xs @ SiS
        t_mean = time.perf_counter() - t
        Z = np.asfortranarray(rng.standard_normal((rows, mm * p)))
        t = time.perf_counter()
        noise = blas.dtrmm(1.0, L, Z, side=1, lower=1, overwrite_b=1)
        t_noise = (time.perf_counter() - t) * (m / mm) ** 2
        d["sample_s"] = (t_mean + t_noise) * (n / rows)
        d["total_s"] = d["equi_s"] + d["ccd_s"] + d["inv_s"] + d["chol_s"] + d["sample_s"]
        total += d["total_s"]
        det[method] = d
        del mui, noise, L, St, Z
    det["total_s_extrapolated"] = total
    return len(methods_of(a)) * n / total, det


def cpu_sample_desc(a, threads, stride):
    return (f"oracle port on host cores ({threads} BLAS threads): eigvalsh on a 2500-block (x(p/2500)^3); every {stride}th "
            f"coordinate of one whole CCD sweep per method (x{stride} x the oracle's full-solve sweep count; 1 thread, like the "
            f"reference's BLAS-2 sweep); full inv(Sigma); dpotrf of a 2p x 2p block of the pm x pm matrix (x(m/2)^3); sampling on "
            f"one {SAMPLE_ROWS}-row chunk (x n/{SAMPLE_ROWS}): mean GEMM + dtrmm against the lower-triangular 2p x 2p factor (x(m/2)^2)")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = host_threads()
    import knockoffs_b200  # noqa: F401  (harness generators only; no GPU object is created on this arm)
    kb = sys.modules["knockoffs_b200"]
    Sigma = kb.harness.block_sigma(a.p, 10, 0.5, 0.25)
    sweeps = {mth: golden_sweeps(mth, a.p, a.m) for mth in methods_of(a)}
    vals, det = [], {}
    t0 = time.perf_counter()
    for _ in range(a.steps):
        v, det = cpu_reference_step(a, Sigma, sweeps, ccd_stride(a.steps))
        vals.append(v)
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    n_rows = len(methods_of(a)) * a.n
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * n_rows / v, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_desc(a, threads, ccd_stride(a.steps)),
                             "detail": det, "sample_wall_s": wall, "extrapolated": True,
                             "note": "the CPU work does not depend on the GPU count: the same value is reported at every N"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of the GPU arm."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        import knockoffs_b200  # noqa: F401
        self.torch, self.dist = torch, dist
        self.kb = sys.modules["knockoffs_b200"]
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.d = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.d)
        self.early_groups = {}
        self.ctx = self.kb.Context(self.local)
        if a.chunk:
            self.ctx.set_sample_chunk_rows(a.chunk)
        self.ext = torch.cuda.ExternalStream(self.ctx.stream, device=self.d)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def min_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.d)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.d)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def bcast(self, t, src):
        if self.world > 1:
            self.dist.broadcast(t, src=src)

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def group_without(self, rank_out):
        """NCCL group of every rank but `rank_out` (created once; every rank must call this)."""
        if rank_out not in self.early_groups:
            ranks = [r for r in range(self.world) if r != rank_out]
            self.early_groups[rank_out] = (self.dist.new_group(ranks=ranks), ranks)
        return self.early_groups[rank_out]


def gen_gaussian_Xt(E, dSigma, p, n, seed):
    """X = Z0 chol(Sigma)' generated on the device (data generation only), stored column-major as Xt[j, i] = X[i, j]."""
    torch = E.torch
    g = torch.Generator(device=E.d)
    g.manual_seed(seed)
    Lc = torch.linalg.cholesky(dSigma)
    Xt = torch.empty((p, n), dtype=torch.float64, device=E.d)
    for r0 in range(0, n, 16384):
        r1 = min(n, r0 + 16384)
        Xt[:, r0:r1] = Lc @ torch.randn((p, r1 - r0), dtype=torch.float64, device=E.d, generator=g)
    return Xt


def run_pipeline(E, a, scaling, steps, warmup, sample_clocks=False):
    """The configs[2] step at E.world GPUs.  Returns a dict of measurements (rank 0 fills the roofline parts)."""
    torch, kb, dev = E.torch, E.kb, E.kb.dev
    ctx, ext, world, rank = E.ctx, E.ext, E.world, E.rank
    p, m, methods = a.p, a.m, methods_of(a)
    f64 = torch.float64
    if scaling == "strong":
        lo, hi = kb.dist.shard_rows(a.n, world, rank)
        n_total = a.n
    else:
        lo, hi = rank * a.n, (rank + 1) * a.n
        n_total = a.n * world
    n_loc = hi - lo
    # Rows per method.  With two methods on N >= 2 ranks the rank that solves the LONG method (:sdp_ccd, 0.52 s, sequential)
    # would keep everyone waiting: instead the other ranks receive the first method's factor early (broadcast inside the group
    # that excludes the long solver) and sample ALL of its rows among themselves while the long solve is still running; the
    # second method's rows are then shared by all ranks.
    owner = [i % world for i in range(len(methods))]
    can_stagger = (world > 1 and len(methods) == 2)
    early = None
    early_ranks = []
    if can_stagger:
        grp, early_ranks = E.group_without(owner[1])
        early = (grp, early_ranks)
    # f = the fraction of the FIRST method's rows that the early ranks (everyone but the long solver) sample while the long
    # solve is still running: as many as fit into that gap (f = 1 from 3-4 ranks on for n = 1e5; ~0.4 on 2 ranks or in the weak
    # run).  The remaining (1 - f) n rows of the first method and all rows of the second are then shared by all ranks.
    # f = 0 is the plain order.  Decided after the first (plain) warm-up step from its measured stage times, see below.
    frac = {"f": 0.0}

    def layout(f):
        return kb.dist.stagger_layout(n_total, world, rank, owner[1] if can_stagger else -1, f if can_stagger else 0.0)

    Sigma_h = kb.harness.block_sigma(p, 10, 0.5, 0.25)
    dSigma = torch.from_numpy(np.ascontiguousarray(Sigma_h)).to(E.d)          # symmetric: row/col-major coincide
    buf = {"X": {}, "Xko_t": None, "n_loc": 0}

    def ensure_buffers(f):
        """This rank's rows of X under layout(f) -- one device matrix per row range (plain shard, early share, late share) --
        and an output buffer for its longest range."""
        e_rng, l_rng = layout(f)
        want = {"plain": (lo, hi), "early": e_rng, "late": l_rng}
        for k, (r0, r1) in want.items():
            have = buf["X"].get(k)
            if r1 <= r0:
                buf["X"].pop(k, None)
            elif have is None or have[0] != r0 or have[1] != r1:
                buf["X"].pop(k, None)
                if k != "plain" and (r0, r1) == (lo, hi):
                    buf["X"][k] = buf["X"]["plain"]
                else:
                    buf["X"][k] = (r0, r1, gen_gaussian_Xt(E, dSigma, p, r1 - r0, 20260003 + rank + 1000 * len(k)))
        need = max(r[1] - r[0] for r in want.values())
        if buf["Xko_t"] is None or need > buf["n_loc"]:
            buf["Xko_t"] = None
            buf["Xko_t"] = torch.empty((m * p, need), dtype=f64, device=E.d)
            buf["n_loc"] = need

    def x_rows(r0, r1):
        """(device pointer of row r0, leading dimension) in the piece of X that holds rows [r0, r1)"""
        for q0, q1, t in buf["X"].values():
            if q0 <= r0 and r1 <= q1:
                return t.data_ptr() + 8 * (r0 - q0), q1 - q0
        raise RuntimeError(f"rows [{r0}, {r1}) of X are not resident on rank {rank}")

    ensure_buffers(0.0)
    Xt, x_lo, n_x, Xko_t, n_loc = buf["X"]["plain"][2], lo, hi - lo, buf["Xko_t"], buf["n_loc"]     # 1-GPU paths below
    dmu = torch.zeros(p, dtype=f64, device=E.d)
    ds = [torch.empty(p, dtype=f64, device=E.d) for _ in methods]
    dSiS = [torch.empty((p, p), dtype=f64, device=E.d) for _ in methods]
    dLb = [torch.empty((2 * m - 1, p, p), dtype=f64, device=E.d) for _ in methods]
    torch.cuda.synchronize()
    st = {"steps": 0, "bcast_ms": 0.0, "sample_ms": 0.0, "staggered": False}
    last = {"solve_factor_ms": 0.0, "sample_ms": 0.0}            # stage times of the most recent step on this rank
    per = {mth: {"solve_ms": 0.0, "ccd_ms": 0.0, "init_ms": 0.0, "factor_ms": 0.0, "ref_bytes": 0.0, "touched_bytes": 0.0,
                 "sweeps": 0, "flops": 0.0, "objective": None} for mth in methods}

    # Factor split (>= 4 ranks, m >= 2): the m block columns of the factor are independent (closed form of the recursion,
    # kb_cond_factor_block_dev), so after the solve only s is broadcast and the ranks build the columns side by side:
    # ~5/3 p^3 flops per rank instead of 5.3 p^3 on the solver, then each column is broadcast from where it was built.
    split_factor = (world >= 4 and m >= 2 and not a.no_split_factor)
    split_model_ms = 16.0            # what the split factor adds after the solve (for the staggering decision only)
    st["split_factor"] = split_factor

    def step(record=False):
        nvtx = torch.cuda.nvtx
        for i, mth in enumerate(methods):
            if owner[i] != rank:
                continue
            e = [E.ev() for _ in range(3)]
            e[0].record(ext)
            nvtx.range_push(f"solve_s:{mth}")
            info = dev.solve_s_dev(ctx, dSigma.data_ptr(), p, mth, m, ds[i].data_ptr(), refresh_every=a.refresh_every)
            nvtx.range_pop()
            e[1].record(ext)
            if not split_factor:
                nvtx.range_push(f"cond_factor:{mth}")
                dev.cond_factor_dev(ctx, dSigma.data_ptr(), p, ds[i].data_ptr(), True, m, dSiS[i].data_ptr(), dLb[i].data_ptr())
                nvtx.range_pop()
            e[2].record(ext)
            ext.synchronize()
            last["solve_factor_ms"] = e[0].elapsed_time(e[2]) + (split_model_ms if split_factor else 0.0)
            if record:
                q = per[mth]
                q["solve_ms"] += e[0].elapsed_time(e[1]); q["factor_ms"] += e[1].elapsed_time(e[2])
                q["ccd_ms"] += info["solve_ms"]; q["init_ms"] += info["init_ms"]; q["ref_bytes"] += info["ref_bytes"]
                q["touched_bytes"] += info["touched_bytes"]; q["sweeps"] = info["sweeps"]; q["mode"] = info["mode"]
                q["flops"] += info["flops"]; q["objective"] = info["objective"]
        f = frac["f"]
        staggered = f > 0.0
        e_rng, l_rng = layout(f)

        def sample(i, r0, r1):
            if r1 <= r0:
                return
            xp, xld = x_rows(r0, r1)
            dev.sample_dev(ctx, xp, r1 - r0, xld, p, dmu.data_ptr(), dSiS[i].data_ptr(),
                           dLb[i].data_ptr(), m, buf["Xko_t"].data_ptr(), buf["n_loc"], seed=20260003 + i, row_offset=r0)

        def bcast_method(i, group=None, ranks=None):
            # s (p doubles) and the factor blocks ((2m) p^2 doubles = 2 GB) from the rank that solved, NCCL over NVLink
            b0, b1 = E.ev(), E.ev()
            b0.record()
            if not split_factor:
                for t in (ds[i], dSiS[i], dLb[i]):
                    E.dist.broadcast(t, src=owner[i], group=group)
                b1.record()
                torch.cuda.current_stream().synchronize()
                return b0.elapsed_time(b1)
            # split factor: only s travels from the solver; block column k of the factor (closed form, independent of the
            # other columns: kb_cond_factor_block_dev) is built by rank ranks[(k - 1) % G] and broadcast from there
            ranks = ranks if ranks is not None else list(range(world))
            G = len(ranks)
            E.dist.broadcast(ds[i], src=owner[i], group=group)
            torch.cuda.current_stream().synchronize()
            f0, f1 = E.ev(), E.ev()
            f0.record(ext)
            for k in range(1, m + 1):
                if ranks[(k - 1) % G] == rank:
                    dev.cond_factor_block_dev(ctx, dSigma.data_ptr(), p, ds[i].data_ptr(), k, m, dSiS[i].data_ptr() if k == 1 else 0,
                                              dLb[i][2 * (k - 1)].data_ptr(), dLb[i][2 * (k - 1) + 1].data_ptr() if k < m else 0)
            f1.record(ext)
            ext.synchronize()
            if record and rank == owner[i]:
                per[methods[i]]["factor_ms"] += f0.elapsed_time(f1)
            b0.record()
            for k in range(1, m + 1):
                src = ranks[(k - 1) % G]
                if k == 1:
                    E.dist.broadcast(dSiS[i], src=src, group=group)
                E.dist.broadcast(dLb[i][2 * (k - 1)], src=src, group=group)
                if k < m:
                    E.dist.broadcast(dLb[i][2 * (k - 1) + 1], src=src, group=group)
            b1.record()
            torch.cuda.current_stream().synchronize()
            return b0.elapsed_time(b1)

        e0, e1 = E.ev(), E.ev()
        samp = 0.0
        bc = 0.0
        if staggered:
            grp, _ = early
            if rank in early_ranks:
                bc += bcast_method(0, grp, early_ranks)
                e0.record(ext); sample(0, *e_rng); e1.record(ext); ext.synchronize()
                samp += e0.elapsed_time(e1)
            bc_all = bcast_method(1)
            if f < 1.0:
                # the rest of the first method's rows is shared by ALL ranks: the long solver has not seen its blocks yet
                b0, b1 = E.ev(), E.ev()
                b0.record()
                for t in (dSiS[0], dLb[0]):
                    E.dist.broadcast(t, src=owner[0])
                b1.record()
                torch.cuda.current_stream().synchronize()
                bc_all += b0.elapsed_time(b1)
            e2, e3 = E.ev(), E.ev()
            e2.record(ext); sample(0, *l_rng); sample(1, lo, hi); e3.record(ext); ext.synchronize()
            samp += e2.elapsed_time(e3)
            if record:
                st["bcast_ms"] += bc_all if rank == owner[1] else bc       # the waiting inside a collective is not transfer time
        else:
            if world > 1:
                for i in range(len(methods)):
                    bc += bcast_method(i)
                if record:
                    st["bcast_ms"] += bc
            e0.record(ext)
            for i in range(len(methods)):
                sample(i, lo, hi)
            e1.record(ext)
            ext.synchronize()
            samp = e0.elapsed_time(e1)
        last["sample_ms"] = samp
        if record:
            st["sample_ms"] += samp
            st["steps"] += 1

    # ---- 1 GPU, two methods, OPT-IN (--pipeline): software pipeline over two contexts.  The :sdp_ccd solve is a chain of small
    #      kernels, the :mvr sampling is pure tensor-core throughput.  A BACKGROUND context
    #      (kb_create_background: every kernel confined to the SM partition that keeps 16 SMs free, lowest priority) samples
    #      the first method's knockoffs on a second host thread WHILE the foreground context solves the second method;
    #      the second method is sampled on the foreground context (whole GPU).  Same work, same outputs, same seeds.
    #      Measured on B200 (r02k): 2476 ms per step against 2397 ms for the plain order -- the solve's own GEMM phases need
    #      about half of the machine, and beside 0.5 ms-long sampling CTAs it slows from 0.53 to 1.44 s.  Hence off by default.
    pipelined = (world == 1 and len(methods) == 2 and a.pipeline)
    ctx_bg = ext_bg = Xko_t2 = None
    if pipelined:
        try:
            ctx_bg = kb.Context(E.local, background=True)
            if a.chunk:
                ctx_bg.set_sample_chunk_rows(a.chunk)
            ext_bg = torch.cuda.ExternalStream(ctx_bg.stream, device=E.d)
            Xko_t2 = torch.empty((m * p, n_loc), dtype=f64, device=E.d)
        except Exception:
            pipelined = False
    st["pipelined"] = pipelined
    st["bg_sample_ms"] = 0.0

    def solve_and_factor(i, mth, record):
        e = [E.ev() for _ in range(3)]
        e[0].record(ext)
        info = dev.solve_s_dev(ctx, dSigma.data_ptr(), p, mth, m, ds[i].data_ptr(), refresh_every=a.refresh_every)
        e[1].record(ext)
        dev.cond_factor_dev(ctx, dSigma.data_ptr(), p, ds[i].data_ptr(), True, m, dSiS[i].data_ptr(), dLb[i].data_ptr())
        e[2].record(ext)
        if record:
            ext.synchronize()
            q = per[mth]
            q["solve_ms"] += e[0].elapsed_time(e[1]); q["factor_ms"] += e[1].elapsed_time(e[2])
            q["ccd_ms"] += info["solve_ms"]; q["init_ms"] += info["init_ms"]; q["ref_bytes"] += info["ref_bytes"]
            q["touched_bytes"] += info["touched_bytes"]; q["sweeps"] = info["sweeps"]; q["mode"] = info["mode"]
            q["flops"] += info["flops"]; q["objective"] = info["objective"]

    def step_pipelined(record=False):
        ready = threading.Event()
        err = []

        def background_sampler():
            try:
                torch.cuda.set_device(E.local)
                ready.wait()
                b0, b1 = E.ev(), E.ev()
                b0.record(ext_bg)
                dev.sample_dev(ctx_bg, Xt.data_ptr(), n_loc, n_loc, p, dmu.data_ptr(), dSiS[0].data_ptr(), dLb[0].data_ptr(), m,
                               Xko_t2.data_ptr(), n_loc, seed=20260003, row_offset=lo)
                b1.record(ext_bg)
                if record:
                    ext_bg.synchronize()
                    st["bg_sample_ms"] += b0.elapsed_time(b1)
            except Exception as ex:      # noqa: BLE001
                err.append(ex)

        th = threading.Thread(target=background_sampler, daemon=True)
        th.start()
        try:
            solve_and_factor(0, methods[0], record)
        finally:
            ready.set()
        solve_and_factor(1, methods[1], record)
        e0, e1 = E.ev(), E.ev()
        e0.record(ext)
        dev.sample_dev(ctx, Xt.data_ptr(), n_loc, n_loc, p, dmu.data_ptr(), dSiS[1].data_ptr(), dLb[1].data_ptr(), m,
                       Xko_t.data_ptr(), n_loc, seed=20260003 + 1, row_offset=lo)
        e1.record(ext)
        th.join()
        if err:
            raise err[0]
        if record:
            ext.synchronize()
            st["sample_ms"] += e0.elapsed_time(e1)
            st["steps"] += 1

    if pipelined:
        step = step_pipelined            # noqa: F811
    for w in range(warmup):
        step()
        if w == 0 and can_stagger and not a.no_stagger:
            # One plain step has been timed on every rank: A / B = solve + factor of the short / long method on their owners,
            # t = this rank's sampling of its plain shard of both methods; sampling all rows of one method on one GPU then
            # takes Ts = t N / 2.  The early ranks can sample  f = (B - A - broadcast) (N - 1) / Ts  of the first method's rows
            # before the long solve ends (capped at 1).  plain: B + 2 Ts / N;  staggered: max(A + f Ts / (N - 1), B) +
            # (2 - f) Ts / N.  Every rank evaluates the same gathered numbers, so all take the same decision.
            gathered = [None] * world
            E.dist.all_gather_object(gathered, (last["solve_factor_ms"], last["sample_ms"]))
            A_ms, B_ms = gathered[owner[0]][0], gathered[owner[1]][0]
            Ts = max(g[1] for g in gathered) * world / 2.0
            f, plain_ms, stag_ms = kb.dist.stagger_fraction(A_ms, B_ms, Ts, world)
            frac["f"] = f
            st["staggered"] = f > 0.0
            st["early_fraction"] = f
            st["stagger_model_ms"] = {"plain": plain_ms, "staggered": stag_ms}
            ensure_buffers(f)
    if warmup == 0 and can_stagger and not a.no_stagger:
        frac["f"] = 1.0 if (scaling == "strong" and world >= 3) else 0.0
        st["staggered"] = frac["f"] > 0.0
        st["early_fraction"] = frac["f"]
        ensure_buffers(frac["f"])
    sampler = ClockSampler(E.local) if (sample_clocks and rank == 0) else None
    E.barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launches
    launches0_bg = ctx_bg.launches if ctx_bg is not None else 0
    t0, t1 = E.ev(), E.ev()
    t0.record(ext)
    for _ in range(steps):
        step(record=True)
    t1.record(ext)                       # pipelined: the background thread has been joined (its stream is drained) before this
    E.barrier()
    ext.synchronize()
    launches = ctx.launches - launches0 + (ctx_bg.launches - launches0_bg if ctx_bg is not None else 0)
    clocks = sampler.stop() if sampler else None
    total_ms = E.max_over_ranks(t0.elapsed_time(t1))
    ms_per_step = total_ms / steps
    out = {"ms_per_step": ms_per_step, "value": len(methods) * n_total / (ms_per_step * 1e-3), "launches": int(launches),
           "clocks": clocks, "n_total": n_total, "rows_per_gpu": buf["n_loc"], "st": st, "per": per, "owner": owner,
           "handles": (dSigma, buf["X"]["plain"][2], dmu, ds, dSiS, dLb, buf["Xko_t"], Sigma_h, lo), "ctx_bg": ctx_bg,
           "shard": (lo, hi, lo), "stag": {"on": frac["f"] == 1.0, "rows0": layout(frac["f"])[0], "early": early,
                                           "X0": buf["X"].get("early", (0, 0, None))[2]}}
    del Xt, Xko_t
    del Xko_t2
    # per-method timings live on their owner: gather them on rank 0
    if world > 1:
        gathered = [None] * world
        E.dist.all_gather_object(gathered, {k: v for k, v in per.items() if owner[methods.index(k)] == rank})
        for g in gathered:
            per.update(g)
        st["sample_ms"] = E.max_over_ranks(st["sample_ms"])
        # a rank that arrives early WAITS inside the collective for the slowest solve: the transfer time itself is what the
        # last rank to arrive measures, i.e. the minimum over ranks
        st["bcast_ms"] = E.min_over_ranks(st["bcast_ms"])
    return out


def run_e2e(E, a, res, steps):
    """The same step through the host-pointer C ABI: kb_solve_s (on the method's owner rank) -> s broadcast (p doubles) ->
    kb_condition on this rank's row shard, pinned host buffers, every H2D / D2H inside the timed region."""
    import ctypes as C
    import psutil
    torch, kb = E.torch, E.kb
    ctx, world, rank = E.ctx, E.world, E.rank
    p, m, methods = a.p, a.m, methods_of(a)
    dSigma, Xt, dmu, ds, dSiS, dLb, Xko_t, Sigma_h, lo = res["handles"]
    lo, hi, x_lo = res["shard"]                   # the plain row shard of this rank; Xt holds rows [x_lo, ...) of X
    n_loc = hi - lo
    need = (p * n_loc + m * p * n_loc) * 8
    avail = psutil.virtual_memory().available / max(1, world)
    n_e = n_loc
    if need > 0.6 * avail:
        n_e = max(1024, int(n_loc * 0.6 * avail / need) // 1024 * 1024)
    Xh_t = torch.empty((p, n_e), dtype=torch.float64, pin_memory=True)
    Xh_t.copy_(Xt[:, lo - x_lo:lo - x_lo + n_e])
    # staggered (N >= 3): like the device-resident step, the ranks that do not run the long solve condition ALL rows of the
    # first method among themselves meanwhile; this rank's share of those rows gets its own pinned X buffer
    stag = res.get("stag") or {"on": False}
    stag_on = bool(stag["on"]) and n_e == n_loc
    r0 = r1 = 0
    Xh0_t = None
    if stag_on:
        r0, r1 = stag["rows0"]
        if r1 > r0:
            Xh0_t = torch.empty((p, r1 - r0), dtype=torch.float64, pin_memory=True)
            Xh0_t.copy_(stag["X0"])
    Xko_h_t = torch.empty((m * p, max(n_e, r1 - r0)), dtype=torch.float64, pin_memory=True)
    mu_h = np.zeros(p)
    Sig_f = np.asfortranarray(Sigma_h)
    s_h = [np.empty(p) for _ in methods]
    opts = kb.make_opts(refresh_every=a.refresh_every)
    info = kb.api.SolveInfo()
    owner = res["owner"]

    def e2e_step():
        for i, mth in enumerate(methods):
            if owner[i] == rank:
                rc = ctx._lib.kb_solve_s(ctx.h, Sig_f.ctypes.data, p, kb.api.METHODS[mth], float(m), C.byref(opts), None,
                                         s_h[i].ctypes.data, C.byref(info))
                kb.api.check(ctx.h, rc)
        if world > 1:
            for i in range(len(methods)):
                t = torch.from_numpy(s_h[i]).to(E.d)
                E.bcast(t, owner[i])
                s_h[i][:] = t.cpu().numpy()
        for i, mth in enumerate(methods):
            rc = ctx._lib.kb_condition(ctx.h, Xh_t.data_ptr(), n_e, p, mu_h.ctypes.data, Sig_f.ctypes.data, s_h[i].ctypes.data, 1, m,
                                       None, 20260003 + i, lo, Xko_h_t.data_ptr())
            kb.api.check(ctx.h, rc)

    def e2e_step_staggered():
        grp, early_ranks = stag["early"]
        if owner[0] == rank:
            rc = ctx._lib.kb_solve_s(ctx.h, Sig_f.ctypes.data, p, kb.api.METHODS[methods[0]], float(m), C.byref(opts), None,
                                     s_h[0].ctypes.data, C.byref(info))
            kb.api.check(ctx.h, rc)
        if rank in early_ranks:
            t = torch.from_numpy(s_h[0]).to(E.d)
            E.dist.broadcast(t, src=owner[0], group=grp)
            s_h[0][:] = t.cpu().numpy()
            if r1 > r0:
                rc = ctx._lib.kb_condition(ctx.h, Xh0_t.data_ptr(), r1 - r0, p, mu_h.ctypes.data, Sig_f.ctypes.data, s_h[0].ctypes.data,
                                           1, m, None, 20260003, r0, Xko_h_t.data_ptr())
                kb.api.check(ctx.h, rc)
        if owner[1] == rank:
            rc = ctx._lib.kb_solve_s(ctx.h, Sig_f.ctypes.data, p, kb.api.METHODS[methods[1]], float(m), C.byref(opts), None,
                                     s_h[1].ctypes.data, C.byref(info))
            kb.api.check(ctx.h, rc)
        t = torch.from_numpy(s_h[1]).to(E.d)
        E.bcast(t, owner[1])
        s_h[1][:] = t.cpu().numpy()
        rc = ctx._lib.kb_condition(ctx.h, Xh_t.data_ptr(), n_e, p, mu_h.ctypes.data, Sig_f.ctypes.data, s_h[1].ctypes.data, 1, m,
                                   None, 20260003 + 1, lo, Xko_h_t.data_ptr())
        kb.api.check(ctx.h, rc)

    # 1 GPU: the same software pipeline as the device-resident step -- kb_condition(:mvr) runs on the background context
    # (second host thread, its own pinned output buffer) while the foreground context solves :sdp_ccd
    ctx_bg = res.get("ctx_bg")
    Xko_h_t2 = None
    if ctx_bg is not None and world == 1 and len(methods) == 2:
        try:
            Xko_h_t2 = torch.empty((m * p, n_e), dtype=torch.float64, pin_memory=True)
        except Exception:
            Xko_h_t2 = None

    def e2e_step_pipelined():
        ready = threading.Event()
        err = []

        def background():
            try:
                torch.cuda.set_device(E.local)
                ready.wait()
                rc = ctx_bg._lib.kb_condition(ctx_bg.h, Xh_t.data_ptr(), n_e, p, mu_h.ctypes.data, Sig_f.ctypes.data,
                                              s_h[0].ctypes.data, 1, m, None, 20260003, lo, Xko_h_t2.data_ptr())
                kb.api.check(ctx_bg.h, rc)
            except Exception as ex:      # noqa: BLE001
                err.append(ex)

        th = threading.Thread(target=background, daemon=True)
        th.start()
        try:
            rc = ctx._lib.kb_solve_s(ctx.h, Sig_f.ctypes.data, p, kb.api.METHODS[methods[0]], float(m), C.byref(opts), None,
                                     s_h[0].ctypes.data, C.byref(info))
            kb.api.check(ctx.h, rc)
        finally:
            ready.set()
        rc = ctx._lib.kb_solve_s(ctx.h, Sig_f.ctypes.data, p, kb.api.METHODS[methods[1]], float(m), C.byref(opts), None,
                                 s_h[1].ctypes.data, C.byref(info))
        kb.api.check(ctx.h, rc)
        rc = ctx._lib.kb_condition(ctx.h, Xh_t.data_ptr(), n_e, p, mu_h.ctypes.data, Sig_f.ctypes.data, s_h[1].ctypes.data, 1, m,
                                   None, 20260003 + 1, lo, Xko_h_t.data_ptr())
        kb.api.check(ctx.h, rc)
        th.join()
        if err:
            raise err[0]

    if Xko_h_t2 is not None:
        e2e_step = e2e_step_pipelined    # noqa: F811
    if stag_on and world > 2 and len(methods) == 2:
        e2e_step = e2e_step_staggered    # noqa: F811
    e2e_step()                     # warm-up
    E.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    E.barrier()
    dt = E.max_over_ranks(time.perf_counter() - t0)
    rows = n_e * world if res["n_total"] != a.n or world == 1 else int(round(a.n * n_e / n_loc))
    nm = len(methods)
    rows_in = (n_e + (r1 - r0)) if e2e_step is e2e_step_staggered else nm * n_e          # rows this rank conditions per step
    out = {"value": nm * rows * steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": int((p * rows_in + nm * (p * p + 2 * p)) * 8 + sum(p * p * 8 for i in range(nm) if owner[i] == rank)),
           "d2h_bytes_per_step": int((m * p * rows_in) * 8 + sum(p * 8 for i in range(nm) if owner[i] == rank)),
           "staggered": e2e_step is e2e_step_staggered,
           "rows_per_gpu": n_e, "ms_per_step": 1000.0 * dt / steps,
           "api": "kb_solve_s + kb_condition (host pointers, pinned); bytes are per rank",
           "pipelined": Xko_h_t2 is not None}
    # the same call sequence from PAGEABLE host memory (plain numpy arrays, what a Julia Matrix{Float64} is): the library
    # stages through its pinned bounce buffers with host copy threads (api.cu host_is_pageable / parallel_copy2d)
    if world == 1 and not a.no_pageable:
        try:
            Xp = np.empty((p, n_e))
            Xp[...] = Xh_t.numpy()
            Xkop = np.empty((m * p, n_e))

            def pageable_step():
                for i, mth in enumerate(methods):
                    rc = ctx._lib.kb_solve_s(ctx.h, Sig_f.ctypes.data, p, kb.api.METHODS[mth], float(m), C.byref(opts), None,
                                             s_h[i].ctypes.data, C.byref(info))
                    kb.api.check(ctx.h, rc)
                    rc = ctx._lib.kb_condition(ctx.h, Xp.ctypes.data, n_e, p, mu_h.ctypes.data, Sig_f.ctypes.data, s_h[i].ctypes.data,
                                               1, m, None, 20260003 + i, lo, Xkop.ctypes.data)
                    kb.api.check(ctx.h, rc)

            pageable_step()            # first touch of the output pages
            t0 = time.perf_counter()
            pageable_step()
            dtp = time.perf_counter() - t0
            same = bool(np.array_equal(Xkop[:, :256], Xko_h_t.numpy()[:, :256]))
            out["pageable"] = {"value": nm * n_e / dtp, "unit": UNIT, "ms_per_step": 1000.0 * dtp, "steps": 1,
                               "host_buffers": "numpy (pageable) X and X-tilde; staged through the library's pinned bounce buffers",
                               "same_bits_as_pinned_run": same}
            del Xp, Xkop
        except MemoryError as ex:
            out["pageable"] = {"unavailable": str(ex)}
    del Xh_t, Xko_h_t, Xko_h_t2, Xh0_t
    return out


def run_config5(E, a):
    """BASELINE configs[4]: genotype-like X (0/1/2 dosages), n = 500000 rows in total sharded over the ranks, p = 5000:
    local Gram partial (DMMA SYRK) -> ONE NCCL all-reduce of the p x p sums (+ p column sums) -> :maxent solve on rank 0
    -> s and the factor broadcast -> sampling on every rank's shard (m = 1)."""
    torch, kb, dev, dist = E.torch, E.kb, E.kb.dev, E.dist
    ctx, ext, world, rank = E.ctx, E.ext, E.world, E.rank
    p, n = 5000, a.c5_rows
    lo, hi = kb.dist.shard_rows(n, world, rank)
    n_loc = hi - lo
    f64 = torch.float64
    g = torch.Generator(device=E.d)
    g.manual_seed(20260005)                                  # same allele frequencies on every rank
    freq = 0.05 + 0.45 * torch.rand(p, dtype=f64, device=E.d, generator=g)
    thr = torch.special.ndtri(freq)[:, None]
    g.manual_seed(20260005 + 1000 * (rank + 1))
    Xt = torch.empty((p, n_loc), dtype=f64, device=E.d)
    for b0 in range(0, p, 50):                               # LD within blocks of 50 through a shared latent
        b1 = min(p, b0 + 50)
        acc = torch.zeros((b1 - b0, n_loc), dtype=f64, device=E.d)
        for _ in range(2):
            z = torch.randn((1, n_loc), dtype=f64, device=E.d, generator=g)
            e = torch.randn((b1 - b0, n_loc), dtype=f64, device=E.d, generator=g)
            acc += ((0.6 ** 0.5) * z + (0.4 ** 0.5) * e < thr[b0:b1]).to(f64)
        Xt[b0:b1] = acc
    del acc, z, e
    G = torch.empty((p, p), dtype=f64, device=E.d)
    cs = torch.empty(p, dtype=f64, device=E.d)
    s = torch.empty(p, dtype=f64, device=E.d)
    SiS = torch.empty((p, p), dtype=f64, device=E.d)
    Lb = torch.empty((1, p, p), dtype=f64, device=E.d)
    Xko = torch.empty((p, n_loc), dtype=f64, device=E.d)
    torch.cuda.synchronize()
    t = {"gram_ms": 0.0, "allreduce_ms": 0.0, "solve_ms": 0.0, "factor_ms": 0.0, "bcast_ms": 0.0, "sample_ms": 0.0}
    info = {}

    def step(record):
        nonlocal info
        e = [E.ev() for _ in range(2)]
        e[0].record(ext)
        dev.gram_partial_dev(ctx, Xt.data_ptr(), n_loc, n_loc, p, G.data_ptr(), cs.data_ptr())
        e[1].record(ext)
        ext.synchronize()
        a0, a1 = E.ev(), E.ev()
        a0.record()
        if world > 1:
            dist.all_reduce(G, op=dist.ReduceOp.SUM)
            dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        a1.record()
        torch.cuda.current_stream().synchronize()
        f = [E.ev() for _ in range(4)]
        f[0].record(ext)
        dev.gram_finish_dev(ctx, G.data_ptr(), cs.data_ptr(), n, p, True, 0.0)
        if rank == 0:
            info = dev.solve_s_dev(ctx, G.data_ptr(), p, "maxent", 1, s.data_ptr())
        f[1].record(ext)
        if rank == 0:
            dev.cond_factor_dev(ctx, G.data_ptr(), p, s.data_ptr(), True, 1, SiS.data_ptr(), Lb.data_ptr())
        f[2].record(ext)
        ext.synchronize()
        b0, b1 = E.ev(), E.ev()
        b0.record()
        E.bcast(s, 0); E.bcast(SiS, 0); E.bcast(Lb, 0)
        b1.record()
        torch.cuda.current_stream().synchronize()
        c0, c1 = E.ev(), E.ev()
        c0.record(ext)
        dev.sample_dev(ctx, Xt.data_ptr(), n_loc, n_loc, p, cs.data_ptr(), SiS.data_ptr(), Lb.data_ptr(), 1, Xko.data_ptr(), n_loc,
                       seed=20260005, row_offset=lo)
        c1.record(ext)
        ext.synchronize()
        if record:
            t["gram_ms"] += e[0].elapsed_time(e[1]); t["allreduce_ms"] += a0.elapsed_time(a1)
            t["solve_ms"] += f[0].elapsed_time(f[1]); t["factor_ms"] += f[1].elapsed_time(f[2])
            t["bcast_ms"] += b0.elapsed_time(b1); t["sample_ms"] += c0.elapsed_time(c1)

    step(False)
    E.barrier()
    l0 = ctx.launches
    t0, t1 = E.ev(), E.ev()
    t0.record(ext)
    K = 2
    for _ in range(K):
        step(True)
    t1.record(ext)
    E.barrier()
    ext.synchronize()
    ms = E.max_over_ranks(t0.elapsed_time(t1)) / K
    for k in t:
        t[k] = (E.min_over_ranks if k in ("bcast_ms", "allreduce_ms") else E.max_over_ranks)(t[k] / K)   # collectives: see run_pipeline
    out = {"workload": f"BASELINE configs[4]: dosage X n={n} rows in total (sharded x{world}), p={p}, Gram + NCCL all-reduce + :maxent + "
                       f"sampling m=1", "value": n / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms, "stages_ms": t,
           "gram_tflops_per_gpu": n_loc * p * (p + 1.0) / (t["gram_ms"] * 1e-3) / 1e12 if t["gram_ms"] > 0 else None,
           "allreduce_bytes": (p * p + p) * 8 if world > 1 else 0,
           "allreduce_busbw_gbs": (2.0 * (world - 1) / world * (p * p + p) * 8 / (t["allreduce_ms"] * 1e-3) / 1e9)
           if (world > 1 and t["allreduce_ms"] > 0) else None,
           "sweeps": info.get("sweeps") if rank == 0 else None, "gpu_launches": int((ctx.launches - l0) / K), "steps": K}
    del Xt, Xko, G, SiS, Lb
    torch.cuda.empty_cache()
    return out


def run_config4(E, a):
    """BASELINE configs[3]: group :mvr knockoff solve, p = 10000 = 8 independent LD blocks of 1250 variables (groups of 5),
    LD blocks dealt round-robin to the ranks.  All blocks sweep in lockstep; the reference's GLOBAL decisions (shared equi
    gamma, per-sweep convergence on the summed objective / max delta) go through the library's reduce hook: a barrier
    between this rank's host threads (one context per LD block) and one NCCL all-reduce of a few doubles between ranks."""
    torch, kb, dist = E.torch, E.kb, E.dist
    world, rank = E.world, E.rank
    nblocks, pb = 8, 1250
    mine = list(range(rank, nblocks, world))
    blocks, groups = {}, {}
    for b in mine:
        g = np.repeat(np.arange(1, pb // 5 + 1), 5)
        rho = 0.7 + 0.02 * b                         # blocks differ, so the global gamma really is a min over blocks
        S = np.where(g[:, None] == g[None, :], rho, rho * 0.25).astype(float)
        np.fill_diagonal(S, 1.0)
        blocks[b], groups[b] = np.asfortranarray(S), g
    ctxs = {b: kb.Context(E.local) for b in mine}
    nth = len(mine)
    bar = threading.Barrier(nth) if nth > 1 else None
    slots = {}
    result = {}
    ops = {0: dist.ReduceOp.MIN, 1: dist.ReduceOp.SUM, 2: dist.ReduceOp.MAX}
    ncoll = [0]

    def make_reduce(b):
        def reduce(op, vals):
            slots[b] = vals.copy()
            if bar is not None:
                bar.wait(timeout=600)
            if b == mine[0]:
                both = np.stack([slots[x] for x in mine])
                loc = both.min(0) if op == 0 else (both.sum(0) if op == 1 else both.max(0))
                if world > 1:
                    t = torch.from_numpy(loc).to(E.d)
                    dist.all_reduce(t, op=ops[op])
                    loc = t.cpu().numpy()
                    ncoll[0] += 1
                slots["res"] = loc
            if bar is not None:
                bar.wait(timeout=600)
            vals[:] = slots["res"]
        return reduce

    def worker(b, small):
        try:
            torch.cuda.set_device(E.local)
            Sg, gg = (blocks[b][:50, :50], groups[b][:50]) if small else (blocks[b], groups[b])
            result[b] = kb.solve_s_group_sharded(Sg, gg, "mvr", make_reduce(b), m=1, ctx=ctxs[b])
        except Exception as ex:      # noqa: BLE001
            result[b] = ex
            if bar is not None:
                bar.abort()

    def run(small):
        th = [threading.Thread(target=worker, args=(b, small), daemon=True) for b in mine]
        for x in th:
            x.start()
        for x in th:
            x.join(900)
        for b in mine:
            if isinstance(result.get(b), Exception):
                raise result[b]

    run(True)                        # warm-up on 50-variable corners (module load, reduce-hook plumbing)
    if not a.c4_cold:
        run(False)                   # and one full-size pass: every context's workspace pool reaches its final size (cudaMalloc is
                                     # device-synchronising, and 8 contexts growing their pools at once serialise each other)
    E.barrier()
    ncoll[0] = 0
    t0 = time.perf_counter()
    run(False)
    E.barrier()
    wall = E.max_over_ranks(time.perf_counter() - t0)
    obj = np.array([sum(result[b][1] for b in mine)])
    tt = torch.from_numpy(obj).to(E.d)
    if world > 1:
        dist.all_reduce(tt)
    infos = [result[b][2] for b in mine]
    out = {"workload": f"BASELINE configs[3]: group :mvr, p={nblocks * pb} = {nblocks} LD blocks of {pb} (groups of 5), m=1, blocks dealt to "
                       f"{world} rank(s), lockstep sweeps", "solve_wall_s": wall, "objective": float(tt.item()),
           "inner_sweeps": infos[0]["inner_sweeps"], "steps_per_block": infos[0]["steps"], "device_ms_per_block": [i["solve_ms"] for i in infos],
           "nccl_allreduces": ncoll[0], "blocks_per_rank": len(mine), "timing": "host wall clock around the lockstep solve, max over ranks; " + ("first full-size solve (cold workspaces)" if a.c4_cold else "second full-size solve (workspaces warm)")}
    for c in ctxs.values():
        c.close() if hasattr(c, "close") else None
    return out


def run_ours(a):
    E = Env(a)
    torch, kb, dev = E.torch, E.kb, E.kb.dev
    world, rank, ctx = E.world, E.rank, E.ctx
    p, m, methods = a.p, a.m, methods_of(a)

    # FP64 tensor peak: MEASURED_PEAKS.json has no f64 entry -> measure cuBLAS DGEMM here (best of 6)
    fp64_peak = None
    if rank == 0:
        A = torch.randn((8192, 8192), dtype=torch.float64, device=E.d)
        B = torch.randn((8192, 8192), dtype=torch.float64, device=E.d)
        best = 1e9
        for _ in range(6):
            e0, e1 = E.ev(), E.ev()
            e0.record(); torch.matmul(A, B); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fp64_peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del A, B
    hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass

    res = run_pipeline(E, a, a.scaling, a.steps, a.warmup, sample_clocks=True)
    dSigma, Xt, dmu, ds, dSiS, dLb, Xko_t, Sigma_h, lo = res["handles"]
    n_loc = min(res["rows_per_gpu"], Xt.shape[1])     # rows of the plain shard: what the stage timed alone below samples

    # ---- the memory-bound formulation, timed alone (outside the step): ONE sweep of the rank-1 streaming kernel
    #      (8 p (p+1)^2 algorithmic bytes) -> achieved HBM GB/s of ccd_stream_kernel
    stream_roof = None
    if rank == 0 and not a.no_stream:
        best = None
        for _ in range(2):
            inf = dev.solve_s_dev(ctx, dSigma.data_ptr(), p, "mvr", m, ds[0].data_ptr(), mode="stream", niter=1, objective=False)
            if best is None or inf["solve_ms"] < best["solve_ms"]:
                best = inf
        gbs = best["ref_bytes"] / (best["solve_ms"] * 1e-3) / 1e9
        stream_roof = {"kernel": "ccd_stream_kernel (KB_MODE_STREAM, one :mvr sweep, timed alone)", "bound": "hbm", "achieved": gbs,
                       "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": NCU_TRAFFIC_BYTES["ccd_stream_kernel"],
                       "peak_source": peak_src, "algorithmic_bytes_per_launch": best["ref_bytes"],
                       "touched_bytes_per_launch": best["touched_bytes"], "launch_ms": best["solve_ms"]}

    # ---- the dominant kernel timed ALONE (one method's sampling stage on the whole GPU, CUDA events on the library's stream):
    #      in the pipelined 1-GPU step the two sampling stages overlap with the :sdp_ccd solve / with each other, so their
    #      in-step durations do not isolate the GEMM
    sample_alone_ms = None
    if rank == 0:
        best = 1e30
        for _ in range(2):
            e0, e1 = E.ev(), E.ev()
            e0.record(E.ext)
            dev.sample_dev(ctx, Xt.data_ptr(), n_loc, Xt.shape[1], p, dmu.data_ptr(), dSiS[0].data_ptr(), dLb[0].data_ptr(), m,
                           Xko_t.data_ptr(), n_loc, seed=20260003, row_offset=lo)
            e1.record(E.ext)
            E.ext.synchronize()
            best = min(best, e0.elapsed_time(e1))
        sample_alone_ms = best

    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(E, a, res, min(a.steps, 10))

    alt = None
    del res["handles"], dSigma, Xt, ds, dSiS, dLb, Xko_t
    torch.cuda.empty_cache()
    extra = {}
    if not a.no_extra:
        if world > 1:
            try:
                other = "weak" if a.scaling == "strong" else "strong"
                r2 = run_pipeline(E, a, other, max(1, min(a.steps, 2)), 1)
                alt = {"scaling": other, "value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms_per_step"], "n_total": r2["n_total"],
                       "rows_per_gpu": r2["rows_per_gpu"], "sampling_ms": r2["st"]["sample_ms"] / max(1, r2["st"]["steps"]),
                       "bcast_ms": r2["st"]["bcast_ms"] / max(1, r2["st"]["steps"])}
                del r2
                torch.cuda.empty_cache()
            except Exception as ex:     # noqa: BLE001
                alt = {"error": repr(ex)}
        for name, fn in (("config5", run_config5), ("config4", run_config4)):
            try:
                extra[name] = fn(E, a)
            except Exception as ex:     # noqa: BLE001
                extra[name] = {"error": repr(ex)}
            torch.cuda.empty_cache()

    if rank == 0:
        k = max(1, res["st"]["steps"])
        per, st = res["per"], res["st"]
        sample_ms = st["sample_ms"] / k                        # foreground sampling inside the step (pipelined: one method)
        n_s = n_loc                                            # rows this rank sampled per method
        structured = m > 1 and os.environ.get("KB_SAMPLE_STRUCTURED", "1") != "0"
        # sampling flops per method.  SURVEY 8(d) counts 3 m n p^2; for Diagonal(s) every M_k = L_kk + (upper triangular N_k), so a
        # copy is two triangular products: executed = algorithmic minimum = (2 m + 1) n p^2.
        survey_1 = 3.0 * m * n_s * float(p) ** 2
        flops_1 = (2.0 * m + 1.0) * n_s * float(p) ** 2 if structured else survey_1
        alone_tfs = flops_1 / (sample_alone_ms * 1e-3) / 1e12 if sample_alone_ms else 0.0
        roof_gemm = {"kernel": "dgemm_dmma_kernel (the sampling stage of one method: mean GEMM + one batched launch per 4096-row chunk)",
                     "bound": "tensor", "achieved": alone_tfs, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": alone_tfs / fp64_peak if fp64_peak else None,
                     "traffic": NCU_TRAFFIC_BYTES["dgemm_dmma_kernel"],
                     "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul f64) measured in this run, best of 6",
                     "algorithmic_flops_per_stage": flops_1, "stage_ms": sample_alone_ms,
                     "timed": "one method's sampling stage alone on the whole GPU right after the timed region, CUDA events on the "
                              "library's stream, best of 2 (inside the pipelined step the stages overlap the :sdp_ccd solve)",
                     "in_step": {"foreground_sampling_ms": sample_ms, "background_sampling_ms": st.get("bg_sample_ms", 0.0) / k,
                                 "pipelined": bool(st.get("pipelined"))},
                     "flop_count": "(2m+1) n p^2 per method: two triangular products per copy (M_k = L_kk + upper N_k for diagonal S)"
                     if structured else "3 m n p^2 (SURVEY 8d)",
                     "survey_3mnp2_equivalent_tflops": survey_1 / (sample_alone_ms * 1e-3) / 1e12 if sample_alone_ms else 0.0}
        solve = {}
        for mth in methods:
            q = per[mth]
            sw = max(1, q["sweeps"])
            ccd_ms = q["ccd_ms"] / k
            fl_alg = float(p) ** 3 * (4.0 / 3.0 + m / 3.0 + 2.0 * (m - 1))
            fl_exe = float(p) ** 3 * (1.0 + m / 3.0 + (m - 1) * 2.0 / 3.0)
            fms = q["factor_ms"] / k
            solve[mth] = {"solve_s_seconds": q["solve_ms"] / k / 1e3, "sweeps": q["sweeps"], "ccd_kernel_ms": ccd_ms,
                          "sweep_ms": ccd_ms / sw, "inverse_builds_ms": q["init_ms"] / k, "objective": q["objective"],
                          "tensor_flops_per_solve": q["flops"] / k,
                          "ccd_tflops_p3_per_sweep": float(p) ** 3 * sw / (ccd_ms * 1e-3) / 1e12 if ccd_ms > 0 else None,
                          "algorithmic_bytes_per_sweep": q["ref_bytes"] / k / sw,
                          "hbm_equivalent_gbs": q["ref_bytes"] / k / (ccd_ms * 1e-3) / 1e9 if ccd_ms > 0 else None,
                          "cond_factor_ms": fms,
                          "cond_factor_executed_tflops": fl_exe / (fms * 1e-3) / 1e12 if fms > 0 else None,
                          "cond_factor_frac_of_peak_executed": (fl_exe / (fms * 1e-3) / 1e12 / fp64_peak) if (fp64_peak and fms > 0) else None,
                          "cond_factor_survey_count_tflops": fl_alg / (fms * 1e-3) / 1e12 if fms > 0 else None,
                          "mode": q.get("mode"), "owner_rank": res["owner"][methods.index(mth)]}
        cfg = config_dict(a, world)
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "parallelism": (f"rows sharded x{world} ({res['shard'][1] - res['shard'][0]} per GPU), solves of the {len(methods)} methods on different ranks, " +
                                ("s NCCL-broadcast, the m block columns of the factor built side by side on different ranks (closed form) and broadcast from there"
                                 if st.get("split_factor") else "s + factor blocks NCCL-broadcast") + (f"; the ranks that do not run the long :sdp_ccd solve sample {100.0 * st.get('early_fraction', 1.0):.0f} % of the :mvr rows meanwhile, "
                                                                        "the rest is shared by all ranks afterwards" if st.get("staggered") else "") if world > 1 else
                                ("single GPU, two contexts: :mvr sampled on the background (SM-partitioned) context beside the :sdp_ccd solve"
                                 if st.get("pipelined") else "single GPU, stages strictly one after the other")),
                "clocks": res["clocks"], "gpu_launches": res["launches"],
                "stages_ms": {"solve_and_factor_critical_path": max(per[x]["solve_ms"] + per[x]["factor_ms"] for x in methods) / k
                              if world > 1 else sum(per[x]["solve_ms"] + per[x]["factor_ms"] for x in methods) / k,
                              "broadcast": st["bcast_ms"] / k, "sampling_foreground": sample_ms,
                              "sampling_background_overlapped": st.get("bg_sample_ms", 0.0) / k},
                "pipelined": bool(st.get("pipelined")),
                "sampling_stage_rows_per_s": (n_loc * world / (sample_alone_ms * 1e-3)) if sample_alone_ms else None,
                "solve": solve, "solve_s_seconds": {x: solve[x]["solve_s_seconds"] for x in methods},
                "roofline": roof_gemm}
        if stream_roof is not None:
            line["roofline_ccd_stream"] = stream_roof
        if e2e is not None:
            line["e2e"] = e2e
        if alt is not None:
            line["scaling_alt"] = alt
        line.update(extra)
        if not a.no_cpu and world == 1:
            threads = host_threads()
            sweeps = {x: (per[x]["sweeps"] or golden_sweeps(x, p, m)) for x in methods}
            v, det = cpu_reference_step(a, Sigma_h, sweeps, 24)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": cpu_sample_desc(a, threads, 24), "detail": det, "extrapolated": True}
        print(json.dumps(line), flush=True)
    if world > 1:
        E.dist.barrier()
        E.dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
