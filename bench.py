#!/usr/bin/env python
"""bench.py — LDA collapsed-Gibbs token-samples/sec at K=1024 (BASELINE.json headline metric).

Workload (BASELINE config 3): LDA, K=1024, V=50,000, 100M-token synthetic corpus (500,000
documents x 200 tokens; phi_k ~ Dirichlet(0.01), theta_d ~ Dirichlet(0.1); beta=0.1 per word,
alpha=0.1), documents sharded contiguously over the ranks, one all-reduce of the count deltas
per sweep.  A "step" is one full sweep over the corpus (every token: remove -> K scores -> draw
-> add, reference src/LDA.jl:113-133).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (one rank per GPU)
  python bench.py --impl reference ...                         the reference's serial CPU sampler
                                                               (oracle port), bounded sample

Prints ONE JSON line on rank 0.
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

K, V, D_TOTAL, L_DOC = 1024, 50_000, 500_000, 200
CORPUS = "dirichlet"            # --corpus zipf: word frequencies ~ 1 / rank (hot words: int32 rows inside the same kernel)
ALPHA, BETA = 0.1, 0.1          # aa_total = BETA * V
PHI_CONC, THETA_CONC = 0.01, 0.1
B_ALG = 4 * K + 12              # algorithmic bytes per token-sample
SEED = 123
CHUNK_DOCS = 10_000             # corpus is generated chunk by chunk, seeded by chunk index


def emit(line: dict):
    """Print the result line on the real stdout (see main: fd 1 points at stderr while the run is in progress)."""
    sys.stdout.flush()
    fd = getattr(emit, "fd", None)
    if fd is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(fd, (json.dumps(line) + "\n").encode())


def sweep_kernel_name(n_hot=0):
    """The kernel bias_model_sweep launches for this workload (csrc/model.cu: use_half16_kernel / use_half_kernel)."""
    force = os.environ.get("BIAS_LDA_KERNEL", "")
    if force.startswith("r"):
        return "lda_md_int_sweep_kernel<8>"
    if force.startswith("h"):
        return "lda_md_int_sweep_half_kernel<16, 8>"
    # corpora with hot words run the instantiation with the int32-row route: 2 CTAs of 8 warps (csrc/lda_half16.cuh, kH16HotWarps)
    return "lda_md_int_sweep_half16_kernel<8, 8, 2, true, true>" if n_hot > 0 else "lda_md_int_sweep_half16_kernel<8, 6, 3, true, false>"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ synthetic corpus
def topic_cdf_torch(device):
    """phi_k ~ Dirichlet(PHI_CONC * 1_V), returned as a flat float64 CDF with topic offsets."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED)
    conc = torch.full((K, V), PHI_CONC, device=device, dtype=torch.float64)
    phi = torch._standard_gamma(conc, generator=g) if hasattr(torch, "_standard_gamma") else torch.distributions.Gamma(conc, 1.0).sample()
    phi = phi + 1e-300
    phi /= phi.sum(1, keepdim=True)
    cdf = torch.cumsum(phi, 1)
    cdf[:, -1] = 1.0
    return (cdf + torch.arange(K, device=device, dtype=torch.float64)[:, None]).reshape(-1)


def gen_docs_torch(cdf_flat, d0, d1, device):
    """Words of documents [d0, d1): theta_d ~ Dirichlet(THETA_CONC), topic ~ theta_d, word ~ phi_topic."""
    import torch
    out = []
    c0 = d0 // CHUNK_DOCS
    c1 = (d1 + CHUNK_DOCS - 1) // CHUNK_DOCS
    for c in range(c0, c1):
        lo, hi = c * CHUNK_DOCS, min((c + 1) * CHUNK_DOCS, D_TOTAL)
        g = torch.Generator(device=device)
        g.manual_seed(SEED * 1_000_003 + c)
        conc = torch.full((hi - lo, K), THETA_CONC, device=device, dtype=torch.float32)
        theta = torch._standard_gamma(conc, generator=g) + 1e-30
        topics = torch.multinomial(theta, L_DOC, replacement=True, generator=g)           # [docs, L]
        u = torch.rand(topics.shape, device=device, dtype=torch.float64, generator=g)
        flat = torch.searchsorted(cdf_flat, topics.to(torch.float64) + u)
        words = (flat - topics * V).clamp_(0, V - 1).to(torch.int32)
        a, b_ = max(d0, lo) - lo, min(d1, hi) - lo
        out.append(words[a:b_].reshape(-1))
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.int32, device=device)


def gen_docs_zipf_torch(d0, d1, device):
    """--corpus zipf: word ids drawn independently with p(w) ~ 1 / (w + 1) (Zipf exponent 1.0), so that a handful of
    words carry a large share of the tokens, as in natural-language corpora (at 100 M tokens over V = 50,000 about 130
    words exceed 65,535 occurrences and hold ~48 % of the tokens)."""
    import torch
    p = 1.0 / torch.arange(1, V + 1, device=device, dtype=torch.float64)
    cdf = torch.cumsum(p / p.sum(), 0)
    out = []
    for c in range(d0 // CHUNK_DOCS, (d1 + CHUNK_DOCS - 1) // CHUNK_DOCS):
        lo, hi = c * CHUNK_DOCS, min((c + 1) * CHUNK_DOCS, D_TOTAL)
        g = torch.Generator(device=device)
        g.manual_seed(SEED * 1_000_003 + c)
        u = torch.rand(((hi - lo), L_DOC), device=device, dtype=torch.float64, generator=g)
        words = torch.searchsorted(cdf, u).clamp_(0, V - 1).to(torch.int32)
        a, b_ = max(d0, lo) - lo, min(d1, hi) - lo
        out.append(words[a:b_].reshape(-1))
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.int32, device=device)


def gen_docs_numpy(d0, d1):
    """CPU generator of the same model (used by the reference arm on a bounded sample)."""
    rng = np.random.default_rng(SEED)
    phi = rng.gamma(PHI_CONC, size=(K, V)) + 1e-300
    phi /= phi.sum(1, keepdims=True)
    cdf = np.cumsum(phi, axis=1)
    rng = np.random.default_rng(SEED + 1 + d0)
    n = d1 - d0
    theta = rng.gamma(THETA_CONC, size=(n, K)) + 1e-300
    tc = np.cumsum(theta, axis=1)
    words = np.empty((n, L_DOC), dtype=np.int64)
    for r in range(n):
        z = np.minimum(np.searchsorted(tc[r], rng.random(L_DOC) * tc[r, -1]), K - 1)
        uw = rng.random(L_DOC)
        for t in range(L_DOC):
            words[r, t] = min(np.searchsorted(cdf[z[t]], uw[t] * cdf[z[t], -1]), V - 1)
    return words.reshape(-1)


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline_run(n_docs: int, n_sweeps: int, warm: int = 0, diag_mode: int = 0):
    """The reference's serial sampler (oracle/bias_oracle.c, a port of src/LDA.jl:113-133) on the
    first n_docs documents of the workload at full K and V, sampling-only (the O(V) monitoring
    diagnostic of src/conjugates.jl:205 is skipped, which favours the reference)."""
    from oracle import oracle as orc
    words = gen_docs_numpy(0, n_docs)
    ptr = np.arange(n_docs + 1, dtype=np.int64) * L_DOC
    rng = np.random.default_rng(7)
    z0 = rng.integers(0, K, len(words))
    state = orc.LDAState(orc.Comps.md(K, V, BETA * V), orc.Data.words(words), ptr, z0, ALPHA, with_diag=False)
    times = []
    for s in range(warm + n_sweeps):
        u = rng.random(len(words))
        t0 = time.perf_counter()
        state.sweep(u, diag_mode=diag_mode)
        if s >= warm:
            times.append(time.perf_counter() - t0)
    return len(words), times


def _ref_worker(widx, n_docs, n_sweeps, warm, barrier, q):
    """One instance of the serial reference sampler on its own documents (own count tables: the reference has no
    shared-memory parallelism, so all-core use means independent instances on disjoint shards)."""
    from oracle import oracle as orc
    words = gen_docs_numpy(widx * n_docs, (widx + 1) * n_docs)
    ptr = np.arange(n_docs + 1, dtype=np.int64) * L_DOC
    rng = np.random.default_rng(7 + widx)
    z0 = rng.integers(0, K, len(words))
    state = orc.LDAState(orc.Comps.md(K, V, BETA * V), orc.Data.words(words), ptr, z0, ALPHA, with_diag=False)
    us = [rng.random(len(words)) for _ in range(warm + n_sweeps)]
    for s_ in range(warm):
        state.sweep(us[s_], diag_mode=0)
    barrier.wait()
    t0 = time.perf_counter()
    for s_ in range(warm, warm + n_sweeps):
        state.sweep(us[s_], diag_mode=0)
    t1 = time.perf_counter()
    q.put((widx, len(words), t0, t1))


def reference_arm(args):
    """The reference's CPU sampler (oracle port of src/LDA.jl:113-133; Julia is not in the image) on the box's host
    cores.  The reference is single-threaded; to use every core, one independent instance runs per core on its own
    shard of documents (the per-instance figure is reported next to the aggregate)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    n_docs = min(args.ref_docs, 400)          # ~4 s of serial CPU work per step and instance
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    mem_gb = 8.0
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                mem_gb = int(ln.split()[1]) / 1e6
    except Exception:
        pass
    procs = max(1, min(cores, 64, int(mem_gb / 1.0)))          # ~0.5 GB per instance (K x V Int64 table)
    mpc = mp.get_context("fork")
    barrier, q = mpc.Barrier(procs), mpc.Queue()
    ws = [mpc.Process(target=_ref_worker, args=(i, n_docs, args.steps, args.warmup, barrier, q)) for i in range(procs)]
    for w_ in ws:
        w_.start()
    res = [q.get() for _ in ws]
    for w_ in ws:
        w_.join()
    n_tok = res[0][1]
    wall = max(r[3] for r in res) - min(r[2] for r in res)
    v = procs * n_tok * args.steps / wall
    per_inst = float(np.mean([n_tok * args.steps / (r[3] - r[2]) for r in res]))
    sample = (f"{procs} independent instances of the serial Float64 sampler (one per host core), each on its own "
              f"{n_docs} documents ({n_tok} tokens) of the K={K}, V={V} workload per step, sampling only; "
              f"one instance alone runs at {per_inst:.0f} tokens/s")
    line = {
        "impl": "reference", "metric": "lda_gibbs_token_samples_per_sec_K1024", "value": v, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "LDA collapsed Gibbs K=1024 V=50000 (BASELINE config 3), bounded sample",
                   "tokens_per_step": n_tok * procs, "K": K, "V": V},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": procs, "kind": "port", "sample": sample,
                         "single_instance_tokens_per_s": per_inst},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------ our arm
def ncu_facts():
    """Per-token figures of the dominant kernel from the committed ncu capture (profiles/lda_sweep_traffic.json)."""
    tp = os.path.join(ROOT, "profiles", "lda_sweep_traffic.json")
    try:
        return json.load(open(tp))
    except Exception:
        return {}


def make_corpus(dev, d0, d1):
    import torch
    if CORPUS == "zipf":
        return gen_docs_zipf_torch(d0, d1, dev)
    cdf = topic_cdf_torch(dev)
    w = gen_docs_torch(cdf, d0, d1, dev)
    del cdf
    torch.cuda.empty_cache()
    return w


def hbm_regime_line(b, ctx, dev, steps, warmup):
    """The same sweep at V = 200,000: the 16-bit rows are a 410 MB table that cannot live in the 126 MB L2, so every
    token's row (2 K bytes) comes from HBM.  This is the regime the metric's roofline was written for; algorithmic
    bytes per token-sample here are 2 K + 12 (16-bit row + word id + z read + z write)."""
    import torch
    global V
    v_saved, V = V, 200_000
    try:
        n_docs = 250_000                          # 50 M tokens: rows are still touched ~250 times each per sweep
        words = make_corpus(dev, 0, n_docs)
        n = int(words.numel())
        g = torch.Generator(device=dev)
        g.manual_seed(SEED + 5)
        z = torch.randint(0, K, (n,), device=dev, dtype=torch.int32, generator=g)
        lda = b.LDA(b.MultinomialDirichlet(V, BETA * V), K, ALPHA)
        data = b.FlatData.from_csr(b._lib.INT, n, group_ptr=np.arange(n_docs + 1, dtype=np.int64) * L_DOC, word=words.cpu().numpy())
        rs = b.ResidentSampler(ctx, lda, data, z.cpu().numpy())
        del words, z
        for _ in range(warmup):
            rs.sweep(1)
        ms = []
        for _ in range(steps):
            rs.sweep(1)
            ms.append(rs.last_sweep_ms()[0])
        rs.close()
        peak, _ = peaks()
        k_ms = float(np.mean(ms))
        b_alg = 2 * K + 12
        achieved = b_alg * n / (k_ms * 1e-3) / 1e9
        return {"workload": f"LDA K={K} V={V} {n} tokens (16-bit rows: {2 * K * V / 1e6:.0f} MB > L2)", "kernel_ms": k_ms,
                "tokens_per_s": n / (k_ms * 1e-3), "alg_bytes_per_token": b_alg, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak}
    finally:
        V = v_saved


def ours(args):
    import torch
    import torch.distributed as dist
    import bias_jl_b200 as b
    from bias_jl_b200.sharding import attach_ranks, device_tensor, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)      # plumbing: barriers, max over ranks, the IPC handles

    n_docs_total = args.docs
    d0, d1 = shard_bounds(n_docs_total, rank, world)
    t_gen = time.perf_counter()
    words_dev = make_corpus(dev, d0, d1)
    n_local = int(words_dev.numel())
    gz = torch.Generator(device=dev)
    gz.manual_seed(SEED + 17 + rank)
    z_dev = torch.randint(0, K, (n_local,), device=dev, dtype=torch.int32, generator=gz)
    # host copies in pinned memory: the C ABI takes host buffers (the Julia arrays)
    words_h = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    z_h = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    words_h.copy_(words_dev)
    z_h.copy_(z_dev)
    torch.cuda.synchronize()
    del z_dev
    torch.cuda.empty_cache()
    ptr = np.arange(d1 - d0 + 1, dtype=np.int64) * L_DOC
    gen_s = time.perf_counter() - t_gen

    stream = torch.cuda.Stream(device=dev)     # a real (non-null) stream: the library launches on it and the events time it
    torch.cuda.set_stream(stream)
    ctx = b.Context(local, seed=SEED, stream=stream.cuda_stream)
    lda = b.LDA(b.MultinomialDirichlet(V, BETA * V), K, ALPHA)
    data = b.FlatData.from_csr(b._lib.INT, n_local, group_ptr=ptr, word=words_h.numpy())

    def make_sampler(c):
        """A resident shard; with more than one rank its sweep call is collective: local sweep + the library's own count
        exchange over NVLink peer memory (csrc/comm.cu) — no NCCL on the data path."""
        rs_ = b.ResidentSampler(c, lda, data, z_h.numpy())
        if world > 1:
            attach_ranks(rs_, n_local)
        return rs_

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs already in HBM when the timed region starts ----
    rs = make_sampler(ctx)
    for _ in range(args.warmup):
        rs.sweep(1)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        rs.sweep(1)
    ev1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launches - launches0
    clock_rec = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    stats = rs.stats()
    narrow_current, n_hot, hot_cap = rs.narrow_info()

    # dominant-kernel duration: the sweep call alone (at N > 1 it includes the exchange), CUDA events on the launching stream
    barrier()
    k_ms = []
    for _ in range(min(3, args.steps)):
        rs.sweep(1)
        k_ms.append(rs.last_sweep_ms()[0])
    kern_ms = max_over_ranks(float(np.mean(k_ms)))
    exchange = None
    if world > 1:
        x_ms, x_bytes = rs.last_exchange()
        x_ms = max_over_ranks(x_ms)
        nvlink_peak = 770.0          # GB/s per direction per GPU, measured peer copy on this pool (B200_PROFILING.md)
        # NVLink traffic of one rank per direction: INCOMING = the rows it owns read from the n - 1 peers (x_bytes) + the n - 1
        # peers' merged rows written into its replica (as many bytes); OUTGOING = the mirror image
        per_dir = 2 * x_bytes
        exchange = {"kernel": "narrow_exchange_kernel", "ms": x_ms, "bytes_read_from_peers_per_rank": x_bytes,
                    "bytes_written_to_peers_per_rank": x_bytes, "nvlink_bytes_per_direction_per_rank": per_dir,
                    "nvlink_gbs_per_direction": per_dir / (x_ms * 1e-3) / 1e9, "peak": nvlink_peak,
                    "frac": per_dir / (x_ms * 1e-3) / 1e9 / nvlink_peak,
                    "note": "device time between the exchange's two barriers (all ranks present), max over ranks; a rank reads the "
                            "rows it owns from every peer and writes the merged rows into every peer's replica, while its peers do "
                            "the same to it: (n-1)/n of the 16-bit table read + as much written arrive at every GPU, and leave it"}

    # ---- every replica against the recount from ALL ranks' assignments (not timed) ----
    z_now = torch.from_numpy(rs.get_zz()).to(dev)
    Kpad = 1024
    recount = torch.bincount(words_dev.to(torch.int64) * Kpad + z_now.to(torch.int64), minlength=V * Kpad).to(torch.int32)
    nk_ref = torch.bincount(z_now.to(torch.int64), minlength=Kpad).to(torch.int32)
    if world > 1:
        dist.all_reduce(recount)
        dist.all_reduce(nk_ref)
    digest = rs.replica_digest()
    rs.delta_begin_zero()                     # zero base => the delta buffer is the live [n_wk | n_k | n_items]
    (p32, n32, _), = rs.delta_pack()
    live = device_tensor(p32, n32, "int32", dev)
    ctx.sync()
    ok = bool(torch.equal(live[:V * Kpad], recount)) and bool(torch.equal(live[V * Kpad:V * Kpad + Kpad], nk_ref))
    chk = torch.tensor([1 if ok else 0, digest & 0x7fffffff, -(digest & 0x7fffffff)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.MIN)
    replica_check = "ok" if int(chk[0]) == 1 and int(chk[1]) == -int(chk[2]) else "FAILED"
    del recount, live, z_now
    ll_sum, ll_n = rs.log_likelihood()           # training log-likelihood per token of the local shard (not timed)
    rs.close()
    del words_dev
    torch.cuda.empty_cache()

    # ---- end-to-end arm: host buffers in, host assignments out, every step ----
    # Streaming use of the public API: two resident models on two contexts (streams) are re-fed in turn
    # (ResidentSampler.reload_u16: H2D of the step's words / zz / offsets from pinned host memory + the
    # initialisation pass, no allocation), swept once (at N > 1 the sweep call first merges the ranks' fresh local
    # counts, then sweeps and exchanges), and read back (get_zz_u16: D2H into pinned host memory).
    # The upload of step s+1 and the download of step s run on the copy engines while step s / s+1 is swept.
    e2e_steps = max(1, args.e2e_steps)
    streams = [stream, torch.cuda.Stream(device=dev)]
    ctxs = [ctx, b.Context(local, seed=SEED + 1000, stream=streams[1].cuda_stream)]
    # K = 1024 and V = 50,000 fit 16 bits: the step's word ids and assignments go up, and the new assignments come
    # back, as uint16 (bias_model_reload_u16 / bias_model_get_zz_u16) - half the PCIe bytes of the int32 calls
    use16 = V <= 65536
    t_fl = time.perf_counter()
    words16_np = words_h.numpy().astype(np.int64).astype(np.uint16 if use16 else np.int32)   # what the shim does: Int64 -> narrow, flat
    flatten_ms = 1e3 * (time.perf_counter() - t_fl)
    words16_h = torch.from_numpy(words16_np).pin_memory()
    z16_h = torch.from_numpy(z_h.numpy().astype(words16_np.dtype)).pin_memory()
    zz_out_h = [torch.from_numpy(np.zeros(n_local, dtype=words16_np.dtype)).pin_memory() for _ in range(2)]
    pipes = []
    for i in range(2):
        with torch.cuda.stream(streams[i]):
            pipes.append(make_sampler(ctxs[i]))

    def e2e_upload(i):
        with torch.cuda.stream(streams[i]):
            if use16:
                pipes[i].reload_u16(words16_h.numpy(), ptr, z16_h.numpy())
            else:
                pipes[i].reload(b.FlatData.from_csr(b._lib.INT, n_local, group_ptr=ptr, word=words16_h.numpy()), z16_h.numpy())

    def e2e_sweep(i):
        with torch.cuda.stream(streams[i]):
            pipes[i].sweep(1)

    def e2e_download(i):
        if use16:
            return pipes[i].get_zz_u16(zz_out_h[i].numpy())      # waits for the model's own stream only
        return pipes[i].get_zz(zz_out_h[i].numpy())

    verbose = bool(os.environ.get("BIAS_BENCH_VERBOSE")) and rank == 0

    def e2e_run(n_steps):
        marks = [time.perf_counter()]
        e2e_upload(0)
        e2e_sweep(0)
        for s_ in range(n_steps):
            cur, nxt = s_ % 2, 1 - (s_ % 2)
            if s_ + 1 < n_steps:
                e2e_upload(nxt)
                marks.append(time.perf_counter())
                e2e_sweep(nxt)
                marks.append(time.perf_counter())
            e2e_download(cur)
            marks.append(time.perf_counter())
        if verbose:
            print("e2e host marks (ms since start; per step: upload done, sweep enqueued, download done):",
                  [round(1e3 * (m_ - marks[0]), 2) for m_ in marks], file=sys.stderr)

    e2e_run(2)                                   # warm both pipelines (first-touch of the pinned buffers)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    moved_e2e = int(pipes[(e2e_steps - 1) % 2].stats().n_moved)
    assert moved_e2e > 0 and int(zz_out_h[(e2e_steps - 1) % 2].numpy().max()) < K
    h2d = words16_h.numel() * words16_h.element_size() + z16_h.numel() * z16_h.element_size() + ptr.nbytes
    d2h = zz_out_h[0].numel() * zz_out_h[0].element_size()
    for rs_i in pipes:
        rs_i.close()
    ctxs[1].close()

    n_total = n_docs_total * L_DOC
    value = n_total * args.steps / (ms_total * 1e-3)
    e2e_value = n_total * e2e_steps / e2e_s
    peak, peak_src = peaks()
    tokens_per_launch = n_local
    achieved = B_ALG * tokens_per_launch / (kern_ms * 1e-3) / 1e9
    facts = ncu_facts()
    traffic = facts.get("dram_bytes_per_token")
    traffic = traffic * tokens_per_launch if traffic is not None else None    # ncu capture scaled to this launch
    l2_bytes = 2 * K + 12 + 4          # what one token-sample pulls through L2: its 16-bit row, word id, z, hot-row index
    hbm_regime = None
    if rank == 0 and world == 1 and not args.no_hbm_regime and V == 50_000 and CORPUS == "dirichlet":
        hbm_regime = hbm_regime_line(b, ctx, dev, 3, 3)

    others = None
    if rank == 0 and world == 1 and not args.no_other_configs and V == 50_000 and CORPUS == "dirichlet" and n_docs_total == D_TOTAL:
        others = {}
        for name in OTHER:
            try:
                o = other_config(name)
                others[name] = {"workload": o["config"]["workload"], "ms_per_sweep": o["ms_per_step"], "items_per_s": o["value"],
                                "e2e_items_per_s": o["e2e"]["value"], "cpu_items_per_s_1core": o["cpu_baseline"]["value"],
                                "cpu_sample": o["cpu_baseline"]["sample"], **o["extra"]}
            except Exception as e:       # a secondary figure must not take the headline line down
                others[name] = {"error": repr(e)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_tok, times = cpu_baseline_run(args.ref_docs, 1, 0)
        cpu = {"value": n_tok / times[0], "unit": "tokens/s", "cores": 1, "kind": "port",
               "sample": f"first {args.ref_docs} documents ({n_tok} tokens) of the same workload, one sweep of the "
                         f"serial Float64 sampler (oracle port of src/LDA.jl:113-133), sampling only"}
        # the reference as it runs: every token also evaluates the O(V) monitoring diagnostic of src/conjugates.jl:205
        f_docs = max(1, args.ref_docs // 50)
        f_tok, f_times = cpu_baseline_run(f_docs, 1, 0, diag_mode=2)
        cpu["faithful_value"] = f_tok / f_times[0]
        cpu["faithful_sample"] = f"first {f_docs} documents ({f_tok} tokens), same sweep with the per-token O(V) diagnostic"
    if rank == 0:
        table_mb = 2 * K * V / 1e6
        line = {
            "metric": "lda_gibbs_token_samples_per_sec_K1024", "value": value, "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"LDA collapsed Gibbs K={K} V={V} {n_total} tokens ({n_docs_total} docs x {L_DOC}), "
                                   f"BASELINE config 3" + ("" if CORPUS == "dirichlet" else f", corpus={CORPUS}"),
                       "K": K, "V": V, "tokens": n_total, "docs": n_docs_total, "corpus": CORPUS,
                       "parallelism": f"documents sharded over {world} GPU(s); per sweep one reduce-scatter + all-gather of the "
                                      f"16-bit count tables by the library's own kernel over NVLink peer memory (no NCCL on the data path)",
                       "l2": f"per sweep the kernel streams {n_total * 12 / world / 1e6:.0f} MB of token arrays (word, z read, z write) per GPU, "
                             f"far larger than the 126 MB L2; the {table_mb:.0f} MB table of 16-bit count rows it gathers from is "
                             + ("smaller than L2 and mostly stays resident between steps (the table is the state being updated, not an input; "
                                "ncu sector hit rate in roofline.l2)" if table_mb < 126 else "larger than L2"),
                       "corpus_gen_s": round(gen_s, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": sweep_kernel_name(n_hot),
                         "kernel_ms": kern_ms, "alg_bytes_per_token": B_ALG, "tokens_per_launch": tokens_per_launch,
                         "note": "achieved uses SURVEY 8(d)'s 4 K + 12 bytes (int32 row from HBM) for continuity; the kernel reads "
                                 "2 K-byte 16-bit rows that mostly hit in L2, so frac > 1 says the metric's HBM bound no longer binds "
                                 "at V = 50,000. What binds: l2 / issue below (from the committed ncu capture); the HBM-bound regime "
                                 "is measured in hbm_regime (V = 200,000, rows cannot stay in L2). kernel_ms spans the sweep call's "
                                 "launches (sweep kernel > 99 %; at N > 1 also the exchange).",
                         "l2": {"bytes_per_token": l2_bytes, "achieved_gbs": l2_bytes * tokens_per_launch / (kern_ms * 1e-3) / 1e9,
                                "sector_hit_rate_pct": facts.get("lts_sector_hit_rate_pct"), "throughput_pct": facts.get("lts_throughput_pct")},
                         "issue": {"issue_slot_pct": facts.get("issue_slot_pct"), "inst_per_token": facts.get("inst_per_token"),
                                   "capture": facts.get("capture")},
                         "hbm_regime": hbm_regime},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "how": "2 resident models on 2 streams, re-fed in turn (reload_u16 = H2D of uint16 word ids / assignments + init pass, "
                           "sweep [N > 1: merge of the fresh local counts + sweep + exchange], get_zz_u16 = D2H); "
                           "wall clock over the steps, barrier + synchronize both sides, max over ranks",
                    "shim_flatten_ms_per_call": flatten_ms,
                    "shim_note": "the Julia arrays are Vector{Vector{Int64}}: flattening / narrowing this rank's tokens to the flat "
                                 "16-bit host arrays fed above costs shim_flatten_ms_per_call on one host core, once per "
                                 "collapsed_gibbs_sampler! call (hundreds of sweeps), not per sweep; it is outside the timed steps"},
            "gpu_launches": int(launches),
            "clocks": clock_rec,
            "replica_check": replica_check,
            "exchange": exchange,
            "other_configs": others,
            "narrow_tables": {"current": bool(narrow_current), "hot_words": int(n_hot), "hot_rows_allocated": int(hot_cap)},
            "sweep_stats": {"n_moved_last": int(stats.n_moved), "diag_last": stats.log_likelihood,
                            "train_loglik_per_token_rank0": ll_sum / max(1, ll_n), "sweeps_done": args.warmup + args.steps},
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if replica_check != "ok":
        raise SystemExit("replica check failed: a rank's count tables differ from the recount of all ranks' assignments")


# ------------------------------------------------------------------ the other BASELINE configs
OTHER = {"c1": "C1 LDA bars K=10 V=25, 1000 docs x 100 tokens", "c2": "C2 BMM x MultinomialDirichlet x Sent, K=64, V=1024, 1M sentences x 64 tokens",
         "c4": "C4 HDP x Gaussian, 100,000 groups x 100 points, Stirling table counts, n_internals=10",
         "c5": "C5 RCRP x Gaussian, 100 epochs x 500 points"}


def other_config(name: str):
    """One of BASELINE.json's other configurations (tools/bench_configs.py builds the workload, sweeps it on cuda:0 and
    times the serial oracle on a bounded sample) as a dict in the shape of the headline line."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    import bias_jl_b200 as b
    ctx = b.Context(0, seed=SEED)
    try:
        r = {"c1": lambda: bc.c1(ctx), "c2": lambda: bc.c2(ctx, 1_000_000), "c4": lambda: bc.c4(ctx, 100_000),
             "c5": lambda: bc.c5(ctx, 100, 500)}[name]()
    finally:
        ctx.close()
    peak, peak_src = peaks()
    ach = r["alg_bytes_per_item"] * r["gpu_items_per_s"] / 1e9
    return {"metric": f"{name}_gibbs_item_samples_per_sec", "value": r["gpu_items_per_s"], "unit": "items/s", "ms_per_step": r["gpu_ms_per_sweep"],
            "config": {"workload": OTHER[name], "detail": r["config"]}, "dtype": "f64" if name != "c1" else "f32",
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "peak_source": peak_src, "alg_bytes_per_item": r["alg_bytes_per_item"],
                         "note": "count tables of this configuration are KB to MB and live in L2 / shared memory: lgamma / fp64-issue / "
                                 "atomic / launch-latency bound by construction (SURVEY 8d), the HBM fraction is stated for completeness"},
            "cpu_baseline": {"value": r["cpu_items_per_s"], "unit": "items/s", "cores": 1, "kind": "port", "sample": r["cpu_sample"]},
            "e2e": r["e2e"], "extra": {k: v for k, v in r.items() if k in ("KK", "KK_true", "KK_serial", "homogeneity", "n_moved_last", "gpu_tokens_per_s")}}


def main():
    global V, CORPUS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=D_TOTAL, help="documents in the corpus (default: the full 100M-token workload)")
    ap.add_argument("--ref-docs", type=int, default=1000, help="documents in the CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=16)
    ap.add_argument("--kernel-timing", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-regime", action="store_true", help="skip the V = 200,000 measurement of roofline.hbm_regime")
    ap.add_argument("--corpus", default="dirichlet", choices=["dirichlet", "zipf"])
    ap.add_argument("--config", default="c3", choices=["c3", "c1", "c2", "c4", "c5"],
                    help="c3 (default): the headline LDA workload; the others print one line for that BASELINE configuration (1 GPU)")
    ap.add_argument("--no-other-configs", action="store_true", help="do not append the other configurations' figures to the headline line")
    ap.add_argument("--vocab", type=int, default=V, help="vocabulary size (default: BASELINE config 3's 50,000)")
    args = ap.parse_args()
    V, CORPUS = args.vocab, args.corpus
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3) if os.environ.get("BIAS_BENCH_STRICT", "1") == "1" else args.warmup
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    emit.fd = saved_stdout
    if args.impl == "reference":
        reference_arm(args)
    elif args.config != "c3":
        line = other_config(args.config)
        line.update({"n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "data": "synthetic"})
        emit(line)
    else:
        ours(args)


if __name__ == "__main__":
    main()
