#!/usr/bin/env python
"""bench.py — Aurora prover hot path (LDE / FRI fold / Merkle) on B200, one JSON line on stdout.

A "step" is one pass of the hot path of ONE Aurora proof over GF(2^192) at N = 2^20 constraints,
RS rate 1/8 (|L| = 2^23), FRI localization [1,2,2,...], non-zk — BASELINE.json configs[1] — i.e. exactly
the sequence of star calls the reference prover makes (traced from the unmodified reference at N = 2^12
and scaled: see WORKLOAD below and DESIGN.md §4):
   9 additive FFTs   (8 onto L = 2^23, 1 onto H = 2^20)          libiop/algebra/fft.tcc:39-124
  11 additive IFFTs  (6 x 2^20, 1 x 2^21, small ones)            libiop/algebra/fft.tcc:126-204
  10 FRI folds       (2^23 -> 2^22 -> 2^20 -> ... -> 2^4)        libiop/protocols/ldt/fri/fri_aux.tcc:37-103
  11 Merkle trees    (4x2^23/2, 1x2^23/2, 1x|L_i|/4, i=1..9)     libiop/bcs/merkle_tree.tcc:93-229
on synthetic (seeded random) oracle contents.  value = N / step time (constraints/s of the hot path).
The elementwise virtual-oracle passes, proof-of-work and transcript assembly of the full prover are
NOT in the step — this is stated in config.workload.

  value : inputs resident in HBM, CUDA-event timed on the context's stream.
  e2e   : the WHOLE Aurora prover at N = 2^20 through the reference-facing API: the reference's own
          aurora_snark_prover (unmodified libiop host code: IOP layer, protocols, BCS transformation, hash chain,
          query phase) over host std::vectors, with this library dropped in under its hot-path headers
          (libiop_b200/lib/libiop_b200_dropin.so, libiop_b200/dropin/).  Every oracle crosses PCIe inside the timed
          region (h2d / d2h bytes are counted by the library).  value = 2^20 / prover wall seconds.  The transcript is
          compared byte for byte with the CPU reference's in the whole_prover_dropin section.
  hot_path_host_buffers : the star calls alone through the host-buffer C ABI (upload, kernel, download per call).
  --impl reference : the unmodified reference aurora_snark_prover on the host CPU (oracle/_ref/libiopref.so, built
          from /root/reference in the build container), single-threaded like the reference is, on a bounded sample
          (N = 2^17) of the same workload; constraints/s.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
RHO = 3
M = LOG_N + RHO  # |L| = 2^23


def workload(log_n=LOG_N):
    """The star-call sequence of aurora_snark_prover (non-zk, rho=3, FRI localization 2), SURVEY.md §3.1."""
    n, m = log_n, log_n + RHO
    N = 1 << n
    ffts = [(32, n), (N - 32, m), (N, m), (N, m), (N, m), (32, m), (N, m), (N, m), (N - 1, m)]
    iffts = [5, n, n, n, n, n, n, 5, n + 1, 4]
    folds = [(m, 1)] + [(k, 2) for k in range(m - 1, 4, -2)]        # (dim of L_i, eta)
    trees = [(4, m, 2), (1, m, 2)] + [(1, k, 4) for k in range(m - 1, 4, -2)]   # (n_oracles, dim, coset)
    return ffts, iffts, folds, trees


def algorithmic_bytes(ffts, iffts, folds, trees):
    """FFT 24*(n_coeffs + 2^m); IFFT 48*2^k; fold 24*|L_i|*(1+2^-eta); Merkle 24*n_or*|L| + 32*(2*leaves-1)."""
    b = sum(24 * (nc + (1 << m)) for nc, m in ffts)
    b += sum(48 * (1 << k) for k in iffts)
    b += sum(24 * (1 << k) * (1 + 2.0 ** -eta) for k, eta in folds)
    b += sum(24 * no * (1 << k) + 32 * (2 * ((1 << k) // cs) - 1) for no, k, cs in trees)
    return b


def reference_field_ops(ffts, iffts, folds):
    """SURVEY.md §8(d) canonical op counts of the reference algorithm (multiplications only)."""
    mul = sum(1.5 * m * (1 << m) for _, m in ffts) + sum(1.5 * k * (1 << k) for k in iffts)
    mul += sum((208 + 4 * (1 << eta)) * ((1 << k) >> eta) for k, eta in folds)
    return mul


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.samples, self.reasons, self.max_mhz = [], set(), None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def run_gpu(args):
    import libiop_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    dev = L.B200(local_rank)
    ffts, iffts, folds, trees = workload()
    rng = np.random.default_rng(1234 + rank)           # independent proof instance per GPU
    n_l = 1 << M

    def rand_dev(n):
        """random field elements generated on the host in chunks (synthetic oracle contents)."""
        buf = dev.alloc(n)
        chunk = 1 << 20
        for off in range(0, n, chunk):
            k = min(chunk, n - off)
            dev.upload(rng.integers(0, 2**64, size=(k, 3), dtype=np.uint64), buf, off)
        return buf

    # ---- device-resident state: 5 full-size oracles + inputs ----
    dom_L = L.Domain.standard(M)
    dom_of = {k: L.Domain.standard(k) for k in set([n for _, n in ffts] + iffts)}
    # FRI domains L_i have non-standard bases (fri_ldt.tcc:310-338); the fold cost does not depend on the basis
    # values, so synthetic random bases are used
    fold_doms = {k: L.Domain.affine(rng.integers(0, 2**64, size=(k, 3), dtype=np.uint64),
                                    rng.integers(0, 2**64, size=3, dtype=np.uint64)) for k, _ in folds}
    coeff_in = rand_dev(1 << LOG_N)
    evals_in = rand_dev(1 << (LOG_N + 1))
    oracles = [dev.alloc(n_l) for _ in range(5)]      # fw, fAz, fBz, fCz / h + FRI chain
    small_out = dev.alloc(1 << (LOG_N + 1))
    fri_bufs = {}
    for k, eta in folds:
        fri_bufs[k - eta] = dev.alloc(1 << (k - eta))
    node_bufs = [dev.alloc_bytes(32 * (2 * ((1 << k) // cs) - 1)) for _, k, cs in trees]
    x_chal = rng.integers(0, 2**64, size=3, dtype=np.uint64)
    launches = [0]

    def step_resident():
        oi = 0
        for nc, m in ffts:
            if m == M:
                dev.gf192_fft_dev(dom_L, coeff_in, nc, oracles[oi % 5]); oi += 1
            else:
                dev.gf192_fft_dev(dom_of[m], coeff_in, nc, small_out)
        for k in iffts:
            dev.gf192_ifft_dev(dom_of[k], evals_in, small_out)
        cur = oracles[4]
        for k, eta in folds:
            dev.fri_fold_dev(1, None, fold_doms[k], eta, x_chal, cur, fri_bufs[k - eta])
            cur = fri_bufs[k - eta]
        for (no, k, cs), nb in zip(trees, node_bufs):
            if no == 4:
                srcs = oracles[:4]
            elif k == M:
                srcs = [oracles[4]]
            else:
                srcs = [fri_bufs[k]]
            dev.merkle_build_dev(srcs, 1 << k, cs, 1, nb)

    # ---- e2e: host buffers through the C ABI (what the C++ adapter does per call) ----
    lib, ctx = dev.lib, dev.ctx
    host_bytes = 24 * n_l
    hp = lib.b200_host_alloc(ctx, C.c_size_t(host_bytes))
    hin = lib.b200_host_alloc(ctx, C.c_size_t(24 << (LOG_N + 1)))
    if not hp or not hin:
        raise RuntimeError("pinned host allocation failed")
    hp_p = C.cast(hp, C.POINTER(C.c_uint64)); hin_p = C.cast(hin, C.POINTER(C.c_uint64))
    np.ctypeslib.as_array(hin_p, shape=(3 << (LOG_N + 1),))[:] = rng.integers(0, 2**64, size=3 << (LOG_N + 1), dtype=np.uint64)
    h2d = [0]; d2h = [0]

    CH = 1 << (LOG_N + 1)

    def up(buf, n):
        for off in range(0, n, CH):
            k = min(CH, n - off)
            dev._chk(lib.b200_upload(ctx, C.c_uint64(buf), C.c_size_t(off), hin_p, C.c_size_t(k))); h2d[0] += 24 * k

    def down(buf, n):
        dev._chk(lib.b200_download(ctx, hp_p, C.c_uint64(buf), C.c_size_t(0), C.c_size_t(n))); d2h[0] += 24 * n

    def step_e2e():
        h2d[0] = d2h[0] = 0
        oi = 0
        for nc, m in ffts:                       # FFT_over_field_subset(host vector, domain) -> host vector
            up(coeff_in, nc)
            out = oracles[oi % 5] if m == M else small_out
            oi += (m == M)
            dev.gf192_fft_dev(dom_L if m == M else dom_of[m], coeff_in, nc, out)
            down(out, 1 << m)
        for k in iffts:
            up(evals_in, 1 << k)
            dev.gf192_ifft_dev(dom_of[k], evals_in, small_out)
            down(small_out, 1 << k)
        cur, cur_n = oracles[4], n_l
        for k, eta in folds:                     # evaluate_next_f_i_over_entire_domain(host vector) -> host vector
            up(cur, cur_n)
            dev.fri_fold_dev(1, None, fold_doms[k], eta, x_chal, cur, fri_bufs[k - eta])
            cur, cur_n = fri_bufs[k - eta], 1 << (k - eta)
            down(cur, cur_n)
        for (no, k, cs), nb in zip(trees, node_bufs):   # merkle_tree::construct(host vectors) -> root to host
            srcs = oracles[:4] if no == 4 else ([oracles[4]] if k == M else [fri_bufs[k]])
            for s in srcs:
                up(s, 1 << k)
            dev.merkle_build_dev(srcs, 1 << k, cs, 1, nb)
            root = (C.c_uint8 * 32)()
            dev._chk(lib.b200_merkle_root(ctx, C.c_uint64(nb), root, C.c_size_t(32))); d2h[0] += 32

    def barrier():
        dev.sync()
        if dist:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        dev.timer_start()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = dev.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        if dist:
            import torch
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, _ = timed(step_resident, args.steps, args.warmup)
    # dominant kernel share: one 2^20 -> 2^23 LDE timed alone with events on the same stream
    for _ in range(2):
        dev.gf192_fft_dev(dom_L, coeff_in, 1 << LOG_N, oracles[0])
    dev.sync(); dev.timer_start()
    reps = 3
    for _ in range(reps):
        dev.gf192_fft_dev(dom_L, coeff_in, 1 << LOG_N, oracles[0])
    lde_ms = dev.timer_stop() / reps
    e2e_steps = max(1, min(args.steps, 3))
    e2e_ms, _ = timed(step_e2e, e2e_steps, 1)
    if rank == 0:
        sampler.stop_flag.set(); sampler.join()

    # launches of our kernels inside one step, counted by the library itself
    lib.b200_launch_count.restype = C.c_ulonglong
    dev.sync(); l0 = lib.b200_launch_count(ctx); step_resident(); dev.sync(); launches_per_step = lib.b200_launch_count(ctx) - l0
    # per-launch CUDA-event timing of the tile passes of one LDE (events on the context's stream)
    lib.b200_profile_begin(ctx)
    dev.gf192_fft_dev(dom_L, coeff_in, 1 << LOG_N, oracles[0])
    pbuf = C.create_string_buffer(1 << 16)
    lib.b200_profile_end(ctx, pbuf, C.c_size_t(len(pbuf)))
    passes = [(l.split("\t")[0], float(l.split("\t")[1])) for l in pbuf.value.decode().strip().split("\n") if "\t" in l]
    bf_passes = [(lab, t) for lab, t in passes if "cosets=8" in lab and "bf=8" in lab]
    dom_label, dom_ms = max(bf_passes, key=lambda x: x[1]) if bf_passes else ("", float("nan"))
    fwd_ms = sum(t for lab, t in passes if "cosets=1" in lab)
    bwd_ms = sum(t for lab, t in passes if "cosets=8" in lab)

    # final transcript gather (the only exchange between ranks): each rank's 11 Merkle roots -> rank 0
    from libiop_b200 import sharding
    roots = b"".join(dev.download_bytes(nb, 32) for nb in node_bufs)
    all_roots = sharding.gather_transcripts(roots)

    ms_per_step = ms / args.steps
    N = 1 << LOG_N
    value = world * N / (ms_per_step * 1e-3)
    e2e_value = world * N / (e2e_ms / e2e_steps * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    lde_bytes = 24 * ((1 << LOG_N) + n_l)
    lde_gbs = lde_bytes / (lde_ms * 1e-3) / 1e9
    # dominant kernel: one in-place butterfly tile pass over the 2^23-element oracle (8 stages, 8 cosets):
    # algorithmic bytes per launch = read 24*2^23 + write 24*2^23
    pass_bytes = 2 * 24 * n_l
    achieved = pass_bytes / (dom_ms * 1e-3) / 1e9
    out = {
        "metric": "Aurora prover constraints/sec at N=2^20 (hot path: LDE + FRI fold + Merkle); LDE GF-ops/sec vs HBM roofline",
        "value": value, "unit": "constraints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "gf2^192 (carry-less u32 limbs) + u64 BLAKE2b", "data": "synthetic",
        "config": {"workload": "aurora_gf192_N2^20_rate1/8_fri_loc2_nonzk: hot-path star calls of one proof "
                               "(9 FFT, 10 IFFT, 10 FRI folds, 11 Merkle trees); elementwise virtual oracles, PoW and transcript "
                               "assembly excluded", "log_n": LOG_N, "codeword_dim": M, "per_gpu": "one independent proof instance",
                   "l2_policy": "working set (5 x 192 MiB oracles + inputs) exceeds the 126 MB L2"},
        "roofline": {"bound": "hbm", "kernel": "additive_tile_pass<0>: butterfly tile pass, 8 stages over the 2^23-point oracle (8 cosets)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                     "traffic": 350.5e6, "traffic_source": "ncu --set full, profiles/r1_ncu_summary_v2.md (dram read+write of one butterfly pass)",
                     "algorithmic_bytes_per_launch": pass_bytes, "launch_ms": dom_ms, "launch_label": dom_label,
                     "gf_mul_per_s_in_kernel": 8 * (n_l // 2) / (dom_ms * 1e-3), "gf_mul_peak_measured": 20.2e9,
                     "frac_of_multiplier_peak": 8 * (n_l // 2) / (dom_ms * 1e-3) / 20.2e9,
                     "lde": {"ms": lde_ms, "forward_half_ms": fwd_ms, "butterfly_half_ms": bwd_ms, "n_launches": len(passes),
                             "algorithmic_bytes": lde_bytes, "achieved_gbs": lde_gbs, "frac": lde_gbs / peak,
                             "gf_mul_per_s_reference_count": 1.5 * M * n_l / (lde_ms * 1e-3),
                             "gf_mul_per_s_executed": 100 * (1 << LOG_N) / (lde_ms * 1e-3)},
                     "note": "integer-pipe bound, not HBM bound: ~950 warp instructions (289 IMAD.WIDE) per GF(2^192) multiplication-butterfly, "
                             "100 multiplications per 24-byte coefficient; see DESIGN.md 3.1-3.2 and profiles/r1_ncu_summary_v2.md"},
        "hot_path_host_buffers": {"value": e2e_value, "unit": "constraints/s", "h2d_bytes_per_step": h2d[0], "d2h_bytes_per_step": d2h[0],
                                  "ms_per_step": e2e_ms / e2e_steps,
                                  "what": "the star calls alone through the host-buffer C ABI (upload, kernel, download per call)"},
        "gpu_launches": None,
        "transcript_gather": {"ranks": len(all_roots), "bytes_per_rank": len(roots),
                              "distinct_instances": len(set(all_roots)) == len(all_roots)},
        "clocks": sampler.summary() if rank == 0 else None,
    }
    # launches per step: counted from the scheduler; measured with ncu in profiles/
    out["gpu_launches"] = int(launches_per_step) * args.steps
    out["gpu_launches_per_step"] = int(launches_per_step)
    # e2e: the whole reference prover over host vectors with the library dropped in, one proof per rank
    e2e = whole_prover_e2e(LOG_N, runs=max(1, min(args.steps, 3)), warmup=2, seed_offset=rank)
    if dist:
        import torch
        t = torch.tensor([e2e.get("prover_s", float("inf"))], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if "prover_s" in e2e:
            e2e["prover_s_max_over_ranks"] = float(t.item())
    if "prover_s" in e2e:
        secs = e2e.get("prover_s_max_over_ranks", e2e["prover_s"])
        e2e = dict({"value": world * N / secs, "unit": "constraints/s"}, **e2e)
    out["e2e"] = e2e
    if rank == 0 and world == 1 and args.concurrency > 1:
        # throughput with several prover processes sharing this GPU (the reference's host code is single-threaded and
        # leaves the device idle 60 % of a proof): P independent instances proving back to back
        out["e2e_throughput"] = whole_prover_throughput("dropin", LOG_N, runs=3, procs=args.concurrency)
    if rank == 0 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sample_log_n=args.cpu_sample_log_n)
        out["cpu_baseline_hot_path"] = cpu_baseline_hot_path(sample_log_n=args.cpu_sample_log_n)
    if rank == 0 and args.dropin_log_n:
        out["whole_prover_dropin"] = whole_prover_dropin(args.dropin_log_n)
    if rank == 0:
        print(json.dumps(out))
    dev.close()
    if dist:
        dist.destroy_process_group()


def estimate_launches(ffts, iffts, folds, trees):
    """Upper-level count of our kernels per step (tile passes per transform are scheduler dependent;
    the ncu launch list under profiles/ is the evidence)."""
    n = 0
    for nc, m in ffts:
        d = max(1, int(np.ceil(np.log2(max(nc, 2)))))
        n += 3 * max(0, d - 12) + 1 + (3 if d > 12 else 1)
    for k in iffts:
        n += 3 * max(0, k - 12) + 1 + (3 if k > 12 else 1)
    n += len(folds)
    for no, k, cs in trees:
        leaves = (1 << k) // cs
        n += 1 + max(0, int(np.log2(max(leaves >> 8, 1))) - 11) + 1
    return n


def cpu_reference_times(sample_log_n, use_ref=True):
    """Times the CPU implementation of the same star calls at N = 2^sample_log_n (bounded sample)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Reference, el, rand_elems, standard_basis
    impl = Reference() if (use_ref and Reference.available()) else Oracle()
    kind = "reference" if isinstance(impl, Reference) else "port"
    rng = np.random.default_rng(5)
    ffts, iffts, folds, trees = workload(sample_log_n)
    t0 = time.perf_counter()
    cache = {}
    for nc, m in ffts:
        c = rand_elems(rng, nc)
        impl.additive_fft(c, m, standard_basis(m), el(1 << m))
    for k in iffts:
        impl.additive_ifft(rand_elems(rng, 1 << k), k, standard_basis(k), el(1 << k))
    for k, eta in folds:
        impl.fri_fold_additive(rand_elems(rng, 1 << k), k, rand_elems(rng, k), rand_elems(rng, 1)[0], eta, rand_elems(rng, 1)[0])
    for no, k, cs in trees:
        orc = [rand_elems(rng, 1 << k) for _ in range(no)]
        impl.merkle_build(orc, cs, 1)
    secs = time.perf_counter() - t0
    return kind, secs


class _quiet_stdout:
    """the reference printf()s its parameter blocks; keep stdout to the one JSON line"""
    def __enter__(self):
        sys.stdout.flush()
        self.devnull = os.open(os.devnull, os.O_WRONLY); self.saved = os.dup(1); os.dup2(self.devnull, 1)
    def __exit__(self, *a):
        os.dup2(self.saved, 1); os.close(self.devnull); os.close(self.saved)


def whole_prover_e2e(log_n, runs=3, warmup=2, seed_offset=0):
    """The reference's aurora_snark_prover over host vectors with the B200 library dropped in (libiop_b200/prover.py):
    wall seconds of the prover call (instance generation excluded, get_transcript included — SURVEY.md §8d metric 1)."""
    try:
        from libiop_b200 import prover
        if not prover.available():
            return {"unavailable": "libiop_b200/lib/libiop_b200_dropin.so was not built (needs /root/reference at build time)"}
        with _quiet_stdout():
            p = prover.DropinProver()
            seed = prover.DEFAULT_SEED + seed_offset          # one independent instance per rank (BASELINE config 5: seeds s, s+1, ...)
            for _ in range(warmup):
                p.aurora_gf192(log_n, seed=seed, verify=False)
            s0 = p.device_stats()
            times, dump = [], b""
            for _ in range(runs):
                secs, dump, _ = p.aurora_gf192(log_n, seed=seed, verify=False)
                times.append(secs)
            s1 = p.device_stats()
        import hashlib
        return {"prover_s": float(np.median(times)), "prover_s_all": times,
                "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // runs, "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // runs,
                "gpu_launches_per_proof": (s1["launches"] - s0["launches"]) // runs,
                "transcript_bytes": len(dump), "transcript_sha256": hashlib.sha256(dump).hexdigest(), "instance_seed": hex(seed),
                "api": "aurora_snark_prover<gf192, binary_hash_digest>(constraint_system, primary_input, auxiliary_input, params), N=2^%d, "
                       "rho=3, FRI localization 2, non-zk, host std::vector oracles" % log_n,
                "residency": p.residency_stats(), "pinned_arena": p.arena_stats()}
    except Exception as e:
        return {"unavailable": repr(e)}


def whole_prover_throughput(impl, log_n, runs, procs):
    """bench/e2e_throughput.py: `procs` prover processes, each proving its own instance `runs` times back to back."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "bench"))
        import e2e_throughput
        r = e2e_throughput.measure(impl, log_n, runs, procs)
        r["unit"] = "constraints/s"
        r["value"] = r.pop("constraints_per_s")
        r["host_cores"] = os.cpu_count()
        return r
    except Exception as e:
        return {"unavailable": repr(e)}


def cpu_reference_prover(sample_log_n):
    """The unmodified reference aurora_snark_prover on one host core (the reference is single-threaded)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Reference
    if not Reference.available():
        return None
    with _quiet_stdout():
        secs, dump, _ = Reference().aurora_gf192(sample_log_n, verify=False)
    return secs


def cpu_baseline(sample_log_n=17):
    secs = cpu_reference_prover(sample_log_n)
    if secs is None:
        return cpu_baseline_hot_path(sample_log_n)
    return {"value": (1 << sample_log_n) / secs, "unit": "constraints/s", "cores": 1, "kind": "reference",
            "sample": f"whole unmodified aurora_snark_prover at N=2^{sample_log_n} ({secs:.2f} s on 1 core; the reference has no threading); "
                      f"same binary measured at N=2^20: 74.6 s = 14.05k constraints/s (profiles/r1_dropin_whole_prover_v2.jsonl)"}


def cpu_baseline_hot_path(sample_log_n=14):
    with _quiet_stdout():
        kind, secs = cpu_reference_times(sample_log_n)
    # extrapolate by the reference's own multiplication counts (O(N log N))
    f_s, i_s, fo_s, _ = workload(sample_log_n)
    f_f, i_f, fo_f, _ = workload(LOG_N)
    scale = reference_field_ops(f_f, i_f, fo_f) / reference_field_ops(f_s, i_s, fo_s)
    est_full = secs * scale
    return {"value": (1 << LOG_N) / est_full, "unit": "constraints/s", "cores": 1, "kind": kind,
            "sample": f"same star-call sequence at N=2^{sample_log_n} ({secs:.2f} s on 1 core incl. input generation), "
                      f"extrapolated x{scale:.1f} by the reference's multiplication count to N=2^{LOG_N}",
            "measured_constraints_per_s_at_sample": (1 << sample_log_n) / secs}


def whole_prover_dropin(log_n):
    """The UNMODIFIED reference aurora_snark_prover (host C++ of libiop) with this library dropped in under its three
    hot-path headers (libiop_b200/lib/libiop_b200_dropin.so, built in the build container) next to the plain CPU reference
    (oracle/_ref/libiopref.so) on the same instance: whole-prover seconds and transcript equality."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import hashlib
        from oracle_lib import Reference
        from libiop_b200 import prover
        if not (Reference.available() and prover.available()):
            return {"unavailable": "oracle/_ref/libiopref.so or libiop_b200/lib/libiop_b200_dropin.so was not built (need /root/reference at build time)"}
        with _quiet_stdout():
            g, r = prover.DropinProver(), Reference()
            g.aurora_gf192(log_n, verify=False)
            gs, gd, acc = g.aurora_gf192(log_n, verify=True)
            cs, cd, _ = r.aurora_gf192(log_n, verify=False)
        return {"log_n": log_n, "cpu_reference_prover_s": cs, "dropin_prover_s": gs, "bit_exact_transcript": cd == gd,
                "transcript_sha256": hashlib.sha256(gd).hexdigest(), "reference_verifier_accepts": acc == 1,
                "cpu_constraints_per_s": (1 << log_n) / cs, "dropin_constraints_per_s": (1 << log_n) / gs,
                "note": "same instance, CPU reference vs drop-in, transcripts compared byte for byte; host-side IOP code is the unmodified "
                        "single-threaded reference in both"}
    except Exception as e:  # never let the optional section break the bench line
        return {"unavailable": repr(e)}


def run_reference(args):
    """Reference arm: the unmodified reference aurora_snark_prover on the host CPU, bounded sample N = 2^cpu_sample_log_n."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    n = args.cpu_sample_log_n
    if args.warmup:
        cpu_reference_prover(max(10, n - 3))
    times = []
    for _ in range(max(1, min(args.steps, 3))):
        secs = cpu_reference_prover(n)
        if secs is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libiopref.so was not built (needs /root/reference at build time)"}))
            return
        times.append(secs)
    secs = float(np.median(times))
    v = (1 << n) / secs
    out = {"impl": "reference",
           "metric": "Aurora prover constraints/sec at N=2^20 (hot path: LDE + FRI fold + Merkle); LDE GF-ops/sec vs HBM roofline",
           "value": v, "unit": "constraints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "gf2^192 (PCLMULQDQ)", "data": "synthetic",
           "config": {"workload": f"aurora_gf192 rate1/8 fri_loc2 nonzk: whole unmodified aurora_snark_prover on the host CPU, bounded sample N=2^{n} "
                                  f"(N=2^20 takes 75 s per proof: 14.05k constraints/s, profiles/r1_dropin_whole_prover_v2.jsonl)",
                      "log_n": n, "codeword_dim": n + RHO},
           "cpu_baseline": {"value": v, "unit": "constraints/s", "cores": 1, "kind": "reference",
                            "sample": f"whole aurora_snark_prover at N=2^{n}, median of {len(times)} runs, single thread (the reference has no threading)"},
           "e2e": {"value": v, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.concurrency > 1:
        # the reference has no threading: one proof uses one core.  For a per-socket figure, one independent prover per core.
        out["cpu_all_cores"] = whole_prover_throughput("cpu", n, runs=1, procs=os.cpu_count() or 1)
    out["wall_s"] = time.perf_counter() - t0
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-log-n", type=int, default=17)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--concurrency", type=int, default=8, help="prover processes sharing the GPU in the e2e_throughput section (1 = skip)")
    ap.add_argument("--dropin-log-n", type=int, default=18, help="side-by-side CPU reference vs drop-in prover with transcript comparison at this size (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
