#!/usr/bin/env python
"""bench.py — Aurora prover over GF(2^192) at N = 2^20 constraints on B200 (BASELINE.json configs[1]); one JSON line on stdout.

Workload (both arms): aurora_snark_prover<gf192> on generate_r1cs_example(2^20, 31, 2^20 - 1), security 128, RS rate 1/8
(|L| = 2^23), FRI localization [1,2,2,...], non-zk, BLAKE2b.

  value : the prover's HOT PATH with every oracle resident in HBM — exactly the sequence of LDE / FRI-fold / Merkle calls one
          proof makes (traced from the unmodified reference; WORKLOAD below), on synthetic oracle contents, CUDA-event timed on
          the context's stream.  constraints/s = 2^20 / step time.  This is the number the roofline explains.
  e2e   : the WHOLE reference prover (unmodified libiop host code: IOP layer, protocols, BCS transformation, hash chain, query
          phase) through the drop-in library (libiop_b200/lib/libiop_b200_dropin.so), R1CS instance and witness in HOST
          memory, every host<->device byte inside the timed region, transcript compared with the CPU reference's fingerprint.
          constraints/s = 2^20 / prover wall seconds.  This is the headline against the reference arm.
  roofline : the dominant kernel (one butterfly tile pass of the LDE over the 2^23-point oracle) timed live with CUDA events;
          DRAM traffic of the same launch read from the committed ncu capture (profiles/r2_ncu/ncu_lde_raw.csv).
  cpu_baseline : the unmodified reference prover (oracle/_ref/libiopref.so, built from /root/reference in the build container)
          on ONE host core — the reference has no threading — on a bounded sample (N = 2^18, ~15 s), and the same instance
          through the drop-in with the transcripts compared byte for byte.
  --impl reference : the unmodified reference prover on the host CPU at the SAME config (N = 2^20): one warm-up proof at
          N = 2^14, then min(steps, 2) timed proofs (75 s each on the GPU box's host); `steps` reports the proofs actually timed.

Under torchrun (N > 1) every rank proves its own instance (seed + rank): weak scaling, no data-path collective; NCCL carries
the barrier, the max-over-ranks time and the gather of the real proof transcripts to rank 0.  At N = 8 the line also carries
BASELINE config 5 (one N = 2^22 proof per GPU, verified, transcripts gathered).
"""
import argparse
import csv
import ctypes as C
import hashlib
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
METRIC = ("Aurora prover constraints/sec at N=2^20 (value: the prover's hot-path calls, oracles resident in HBM; "
          "e2e: whole aurora_snark_prover, host vectors); LDE GF-ops/sec vs HBM roofline")
CONFIG = {"workload": "aurora_gf192_N2^20_rate1/8_fri_loc2_nonzk", "log_n": LOG_N, "codeword_dim": M, "field": "GF(2^192)",
          "security": 128, "zk": False, "instance": "generate_r1cs_example(2^20, 31, 2^20-1), seed 0x9E3779B97F4A7C15 + rank",
          "per_gpu": "one independent proof instance",
          "l2_policy": "working set (5 x 192 MiB oracles + inputs) exceeds the 126 MB L2"}


def workload(log_n=LOG_N):
    """The hot-path call sequence of aurora_snark_prover (non-zk, rho=3, FRI localization 2), SURVEY.md §3.1."""
    n, m = log_n, log_n + RHO
    N = 1 << n
    # the calls the drop-in's prover makes (bench/host_gaps.py lists them; profiles/r2_host_gaps_N20_v4.txt):
    # f_1v over the variable domain, f_1v over L (f_z)
    ffts = [(32, n), (32, m)]
    # evaluations over H / the variable domain straight to L (b200_gf192_lde; the reference's IFFT + FFT pairs): f_w, f_Az, f_Bz,
    # f_Cz (r1cs_rs_iop.tcc:449-479, 497-569), p_alpha' and p_alpha_ABC (basic_lincheck_aux.tcc:86-144)
    ldes = [(n, m)] * 6
    # f_1v (twice: submit_witness_oracles and the f_z oracle), the final FRI polynomial
    iffts = [5, 5, 4]
    # the sumcheck quotient h over L from f on two cosets of H (sumcheck.tcc:344-388: the reference's IFFT(2^21), division by Z_H,
    # FFT(2^23)): two interpolation halves, one evaluation half (b200_gf192_lde_interpolate / _evaluate)
    quotients = [(n, m)]
    folds = [(m, 1)] + [(k, 2) for k in range(m - 1, 4, -2)]        # (dim of L_i, eta)
    trees = [(4, m, 2), (1, m, 2)] + [(1, k, 4) for k in range(m - 1, 4, -2)]   # (n_oracles, dim, coset)
    return ffts, ldes, iffts, folds, trees, quotients


def algorithmic_bytes(ffts, ldes, iffts, folds, trees, quotients=()):
    """FFT / LDE 24*(n_in + 2^m); IFFT 48*2^k; fold 24*|L_i|*(1+2^-eta); Merkle 24*n_or*|L| + 32*(2*leaves-1)."""
    b = sum(24 * (nc + (1 << m)) for nc, m in ffts) + sum(24 * ((1 << d) + (1 << m)) for d, m in ldes)
    b += sum(48 * (1 << k) for k in iffts)
    b += sum(2 * 48 * (1 << d) + 24 * ((1 << d) + (1 << m)) for d, m in quotients)
    b += sum(24 * (1 << k) * (1 + 2.0 ** -eta) for k, eta in folds)
    b += sum(24 * no * (1 << k) + 32 * (2 * ((1 << k) // cs) - 1) for no, k, cs in trees)
    return b


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


class _quiet_stdout:
    """the reference printf()s its parameter blocks; keep stdout to the one JSON line"""
    def __enter__(self):
        sys.stdout.flush()
        self.devnull = os.open(os.devnull, os.O_WRONLY); self.saved = os.dup(1); os.dup2(self.devnull, 1)
    def __exit__(self, *a):
        os.dup2(self.saved, 1); os.close(self.devnull); os.close(self.saved)


def ncu_traffic(kernel_substr="additive_tile_pass<0", grid=2048):
    """DRAM bytes (read + write) of the dominant launch from the committed `ncu --set full` capture of bench/profile_kernels.py
    lde (bench/ncu_families.sh).  The transform has two 8-stage butterfly passes of that kernel with that grid and the same
    duration: the first writes the 2^23-point array out of place from the 2^20 coefficients all cosets share (half the traffic),
    the second is the in-place strided pass the roofline entry is about — the launch with the most DRAM traffic among those
    lasting at least 90 % of the longest.  None when the capture is missing."""
    path = os.path.join(ROOT, "profiles", "r2_ncu", "ncu_lde_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        cands = []
        for r in rows[2:]:
            if kernel_substr not in r[ix["Kernel Name"]] or int(float(r[ix["launch__grid_size"]])) != grid:
                continue
            dur = float(r[ix["gpu__time_duration.sum"]])
            tot = sum(float(r[ix[k]]) * scale[units[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            cands.append((dur, tot))
        if not cands:
            return None, None
        longest = max(d for d, _ in cands)
        return max(t for d, t in cands if d >= 0.9 * longest), os.path.relpath(path, ROOT)
    except Exception:
        return None, None


def gpu_cpu_affinity(index, n_cpus):
    """the CPUs NVML names as closest to GPU `index` (same NUMA node / PCIe root), or None"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpus + 63) // 64)
        return {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
    except Exception:
        return None


def pin_rank_to_cores(local_rank, world):
    """one disjoint slice of the host cores per rank: the reference's host code is single-threaded, the drop-in flattens the
    constraint matrices on a few threads, and every rank drives its own PCIe traffic — unpinned, 8 ranks migrate over the same
    cores (round-1 SCALE: e2e efficiency 0.856 at N = 8).  The slice is taken from the cores NVML reports as closest to the
    rank's GPU (shared evenly with the other ranks whose GPUs report the same set) when that is known, else from an even split
    of the process's cores."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < world:
            return None
        n_cpus = os.cpu_count() or (max(cores) + 1)
        near = [gpu_cpu_affinity(r, n_cpus) for r in range(world)]
        mine = None
        if all(a for a in near):
            group = [r for r in range(world) if near[r] == near[local_rank]]          # ranks sharing this GPU's CPU set
            pool = [c for c in cores if c in near[local_rank]]
            per = len(pool) // len(group)
            if per >= 1:
                k = group.index(local_rank)
                mine = pool[k * per:(k + 1) * per]
        if not mine:
            per = len(cores) // world
            mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, set(mine))
        return mine
    except Exception:
        pass
    return None


def run_gpu(args):
    import libiop_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = pin_rank_to_cores(local_rank, world)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    dev = L.B200(local_rank)
    ffts, ldes, iffts, folds, trees, quotients = workload()
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
    dom_of = {k: L.Domain.standard(k) for k in set([n for _, n in ffts] + iffts + [d for d, _ in ldes])}
    # the two cosets of span(basis[0..d)) the first 2^(d+1) positions of L lie on (shift of L, shift + basis[d])
    coset_of = {d: (L.Domain.standard(d, shift_int=1 << M), L.Domain.standard(d, shift_int=(1 << M) | (1 << d))) for d, _ in quotients}
    half_a, half_b = dev.alloc(1 << LOG_N), dev.alloc(1 << LOG_N)
    # FRI domains L_i have non-standard bases (fri_ldt.tcc:310-338); the fold cost does not depend on the basis values
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

    def step_resident():
        oi = 0
        for nc, m in ffts:
            if m == M:
                dev.gf192_fft_dev(dom_L, coeff_in, nc, oracles[oi % 5]); oi += 1
            else:
                dev.gf192_fft_dev(dom_of[m], coeff_in, nc, small_out)
        for d, m in ldes:
            dev.gf192_lde_dev(dom_of[d], dom_L, evals_in, oracles[oi % 5]); oi += 1
        for k in iffts:
            dev.gf192_ifft_dev(dom_of[k], evals_in, small_out)
        for d, m in quotients:
            dev.gf192_lde_interpolate_dev(coset_of[d][0], evals_in, half_a)
            dev.gf192_lde_interpolate_dev(coset_of[d][1], evals_in, half_b)
            dev.gf192_lde_evaluate_dev(dom_L, d, half_a, oracles[oi % 5]); oi += 1
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

    lib, ctx = dev.lib, dev.ctx

    def barrier():
        dev.sync()
        if dist:
            dist.barrier()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        dev.timer_start()
        for _ in range(steps):
            fn()
        ms = dev.timer_stop()
        barrier()
        if dist:
            import torch
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps, max(args.warmup, 3))
    # the dominant kernel, timed live: one 2^20 -> 2^23 LDE with per-launch CUDA events on the context's stream
    for _ in range(2):
        dev.gf192_fft_dev(dom_L, coeff_in, 1 << LOG_N, oracles[0])
    dev.sync(); dev.timer_start()
    reps = 3
    for _ in range(reps):
        dev.gf192_fft_dev(dom_L, coeff_in, 1 << LOG_N, oracles[0])
    lde_ms = dev.timer_stop() / reps
    lib.b200_launch_count.restype = C.c_ulonglong
    dev.sync(); l0 = lib.b200_launch_count(ctx); step_resident(); dev.sync(); launches_per_step = lib.b200_launch_count(ctx) - l0
    lib.b200_profile_begin(ctx)
    dev.gf192_fft_dev(dom_L, coeff_in, 1 << LOG_N, oracles[0])
    pbuf = C.create_string_buffer(1 << 16)
    lib.b200_profile_end(ctx, pbuf, C.c_size_t(len(pbuf)))
    passes = [(l.split("\t")[0], float(l.split("\t")[1])) for l in pbuf.value.decode().strip().split("\n") if "\t" in l]
    bf_passes = [(lab, t) for lab, t in passes if "cosets=8" in lab and "bf=8" in lab]
    # the in-place strided 8-stage pass (the last of the two 8-stage passes; both take the same time)
    dom_label, dom_ms = bf_passes[-1] if bf_passes else ("", float("nan"))
    fwd_ms = sum(t for lab, t in passes if "cosets=1" in lab)
    bwd_ms = sum(t for lab, t in passes if "cosets=8" in lab)
    # the field multiplier's own rate, measured live (element-wise product of two resident 2^22-element vectors)
    mul_n = 1 << 22
    dev._chk(lib.b200_gf192_mul(ctx, C.c_uint64(oracles[1]), C.c_uint64(oracles[2]), C.c_uint64(oracles[3]), C.c_size_t(mul_n)))
    dev.sync(); dev.timer_start()
    for _ in range(3):
        dev._chk(lib.b200_gf192_mul(ctx, C.c_uint64(oracles[1]), C.c_uint64(oracles[2]), C.c_uint64(oracles[3]), C.c_size_t(mul_n)))
    mul_rate = 3 * mul_n / (dev.timer_stop() * 1e-3)
    if rank == 0:
        sampler.stop_flag.set(); sampler.join()

    ms_per_step = ms / args.steps
    N = 1 << LOG_N
    value = world * N / (ms_per_step * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    lde_bytes = 24 * ((1 << LOG_N) + n_l)
    lde_gbs = lde_bytes / (lde_ms * 1e-3) / 1e9
    # dominant kernel: one in-place butterfly tile pass over the 2^23-element oracle (8 stages, 8 cosets):
    # algorithmic bytes per launch = read 24*2^23 + write 24*2^23 (2*24*2^m per pass over 2^m points)
    pass_bytes = 2 * 24 * n_l
    achieved = pass_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    out = {
        "metric": METRIC, "value": value, "unit": "constraints/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "gf2^192 (carry-less u32 limbs) + u64 BLAKE2b", "data": "synthetic",
        "config": dict(CONFIG),
        "value_scope": "hot-path calls of one proof as the drop-in prover makes them (6 fused LDE, the sumcheck quotient as 2 interpolation + 1 evaluation halves, 2 short FFT, 3 short IFFT, 10 FRI folds, 11 Merkle trees), oracles resident in HBM; "
                       "the whole prover is `e2e`",
        "roofline": {"bound": "hbm", "kernel": "additive_tile_pass<0>: butterfly tile pass, 8 stages over the 2^23-point oracle (8 cosets)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": pass_bytes, "launch_ms": dom_ms, "launch_label": dom_label,
                     "gf_mul_per_s_in_kernel": 8 * (n_l // 2) / (dom_ms * 1e-3),
                     "gf_mul_per_s_elementwise_kernel_live": mul_rate,
                     "lde": {"ms": lde_ms, "forward_half_ms": fwd_ms, "butterfly_half_ms": bwd_ms, "n_launches": len(passes),
                             "algorithmic_bytes": lde_bytes, "achieved_gbs": lde_gbs, "frac": lde_gbs / peak,
                             "gf_mul_per_s_reference_count": 1.5 * M * n_l / (lde_ms * 1e-3),
                             "gf_mul_per_s_executed": 100 * (1 << LOG_N) / (lde_ms * 1e-3)},
                     "note": "integer-pipe bound, not HBM bound: ~850 instructions (292 IMAD.WIDE at quarter rate + 512 LOP3) per GF(2^192) "
                             "multiplication-butterfly, 100 multiplications per 24-byte coefficient; traffic (ncu, in-place strided pass) = 0.87 x the algorithmic bytes: "
                             "part of the 192 MiB array is still in the 126 MB L2 from the previous pass; DESIGN.md 3.1-3.2, profiles/r2_variants.md"},
        "algorithmic_bytes_per_step": algorithmic_bytes(ffts, ldes, iffts, folds, trees, quotients),
        "gpu_launches": int(launches_per_step) * args.steps, "gpu_launches_per_step": int(launches_per_step),
        "clocks": sampler.summary() if rank == 0 else None,
        "host_cores_of_rank": cores,
    }
    # free the synthetic state before the whole provers run
    for b in oracles + [coeff_in, evals_in, small_out, half_a, half_b] + list(fri_bufs.values()):
        dev.free(b)
    dev.sync()

    # ---- e2e: the whole reference prover over host vectors with the library dropped in, one proof per rank ----
    e2e = whole_prover_e2e(LOG_N, runs=max(1, min(args.steps, 3)), warmup=2, seed_offset=rank, verify_last=(world > 1))
    transcript = e2e.pop("_transcript", b"")
    if dist:
        import torch
        t = torch.tensor([e2e.get("prover_s", float("inf"))], device="cuda")
        per_rank = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if "prover_s" in e2e:
            e2e["prover_s_max_over_ranks"] = float(t.item())
            e2e["prover_s_per_rank"] = [round(float(x.item()), 5) for x in per_rank]
    if "prover_s" in e2e:
        secs = e2e.get("prover_s_max_over_ranks", e2e["prover_s"])
        e2e = dict({"value": world * N / secs, "unit": "constraints/s"}, **e2e)
    out["e2e"] = e2e
    # final transcript gather (the only exchange between ranks): every rank's real proof transcript -> rank 0
    from libiop_b200 import sharding
    payload = bytes([1 if e2e.get("reference_verifier_accepts") else 0]) + transcript
    gathered = sharding.gather_transcripts(payload)
    out["transcript_gather"] = {"ranks": len(gathered), "bytes_per_rank": len(transcript),
                                "distinct_instances": len(set(gathered)) == len(gathered),
                                "accepted_by_reference_verifier": (sum(g[0] for g in gathered) if world > 1 else None),
                                "sha256_per_rank": [hashlib.sha256(g[1:]).hexdigest()[:16] for g in gathered]}
    if world >= 8 or args.config5:
        out["config5"] = config5_leg(rank, world, dist)
    if rank == 0 and world == 1 and args.concurrency > 1:
        # throughput with several prover processes sharing this GPU (the reference's host code is single-threaded and leaves the
        # device idle most of a proof): P independent instances proving back to back
        out["e2e_throughput"] = whole_prover_throughput("dropin", LOG_N, runs=3, procs=args.concurrency)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.cpu_sample_log_n)
    if rank == 0:
        print(json.dumps(out))
    dev.close()
    if dist:
        dist.destroy_process_group()


def golden_fingerprint(prover, log_n, zk=False):
    try:
        for c in json.load(open(os.path.join(ROOT, "tests", "golden", "prover_fingerprints_config.json"))):
            if (c["prover"], c["log_n"], c["zk"]) == (prover, log_n, zk):
                return c
    except Exception:
        pass
    return None


def whole_prover_e2e(log_n, runs=3, warmup=2, seed_offset=0, verify_last=False):
    """The reference's aurora_snark_prover over host vectors with the B200 library dropped in (libiop_b200/prover.py):
    wall seconds of the prover call (instance generation excluded, get_transcript included — SURVEY.md §8d metric 1)."""
    try:
        from libiop_b200 import prover
        if not prover.available():
            return {"unavailable": "libiop_b200/lib/libiop_b200_dropin.so was not built (needs /root/reference at build time)"}
        with _quiet_stdout():
            p = prover.DropinProver()
            seed = prover.DEFAULT_SEED + seed_offset          # one independent instance per rank (BASELINE config 5: seeds s, s+1, ...)
            warm = [p.aurora_gf192(log_n, seed=seed, verify=False)[0] for _ in range(warmup)]
            s0, l0 = p.device_stats(), p.lazy_stats()
            times, dump, acc = [], b"", -1
            for r in range(runs):
                secs, dump, _ = p.aurora_gf192(log_n, seed=seed, verify=False)
                times.append(secs)
            s1, l1 = p.device_stats(), p.lazy_stats()
            if verify_last:                                   # outside the timed proofs: the reference verifier on this rank's proof
                _, dump2, acc = p.aurora_gf192(log_n, seed=seed, verify=True)
                acc = acc if dump2 == dump else 0
        sha = hashlib.sha256(dump).hexdigest()
        g = golden_fingerprint("aurora_gf192", log_n) if seed_offset == 0 else None
        return {"prover_s": float(np.median(times)), "prover_s_all": times,
                # the untimed warm-up proofs of this process: the first builds the per-domain tables (twiddles, recursed shifts,
                # x^-1 over the codeword domain, index tables), the pinned slabs and the lazy-vector reservation
                "warmup_prover_s": warm,
                "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"]) // runs, "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"]) // runs,
                "gpu_launches_per_proof": (s1["launches"] - s0["launches"]) // runs,
                "transcript_bytes": len(dump), "transcript_sha256": sha, "instance_seed": hex(seed),
                "equals_cpu_reference_transcript": (sha == g["sha256"]) if g else None,
                "cpu_reference_fingerprint": "tests/golden/prover_fingerprints_config.json (unmodified reference, CPU)" if g else None,
                "reference_verifier_accepts": (acc == 1) if acc >= 0 else None,
                "api": "aurora_snark_prover<gf192, binary_hash_digest>(constraint_system, primary_input, auxiliary_input, params), N=2^%d, "
                       "rho=3, FRI localization 2, non-zk, host std::vector inputs" % log_n,
                "device_resident_oracles": {"host_touched_vectors": (l1["faults"] - l0["faults"]) // runs,
                                            "query_gathers": (l1["gathers"] - l0["gathers"]) // runs, "residency": p.residency_stats()},
                "pinned_arena": p.arena_stats(), "_transcript": dump}
    except Exception as e:
        return {"unavailable": repr(e)}


def config5_leg(rank, world, dist):
    """BASELINE config 5: one Aurora N = 2^22 proof per GPU (seed + rank), verified by the reference verifier on its own rank,
    transcripts gathered on rank 0; rank 0's transcript is compared with the CPU reference's fingerprint."""
    try:
        import psutil
        need = 14 << 30
        if psutil.virtual_memory().available < need * max(1, min(world, 8)):
            return {"skipped": "not enough free host memory for %d provers at N = 2^22" % world}
        from libiop_b200 import prover, sharding
        with _quiet_stdout():
            p = prover.DropinProver()
            seed = prover.DEFAULT_SEED + rank
            p.aurora_gf192(22, seed=seed, verify=False)
            secs, dump, acc = p.aurora_gf192(22, seed=seed, verify=True)
        if dist:
            import torch
            t = torch.tensor([secs], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs_max = float(t.item())
        else:
            secs_max = secs
        gathered = sharding.gather_transcripts(bytes([1 if acc == 1 else 0]) + dump)
        g = golden_fingerprint("aurora_gf192", 22)
        return {"workload": "aurora_gf192_N2^22_rate1/8_fri_loc2_nonzk, one proof per GPU", "log_n": 22, "prover_s_max_over_ranks": secs_max,
                "value": world * (1 << 22) / secs_max, "unit": "constraints/s",
                "transcript_gather": {"ranks": len(gathered), "bytes_per_rank": len(dump), "accepted_by_reference_verifier": sum(x[0] for x in gathered),
                                      "distinct_instances": len(set(gathered)) == len(gathered)},
                "rank0_equals_cpu_reference_transcript": (hashlib.sha256(dump).hexdigest() == g["sha256"]) if (g and rank == 0) else None}
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


def cpu_baseline(sample_log_n=18):
    """The unmodified reference aurora_snark_prover on one host core (the reference is single-threaded) on a bounded sample of the
    workload, and the same instance through the drop-in with the transcripts compared byte for byte."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        from oracle_lib import Reference
        from libiop_b200 import prover
        if not Reference.available():
            return {"unavailable": "oracle/_ref/libiopref.so was not built (needs /root/reference at build time)"}
        with _quiet_stdout():
            cs, cd, _ = Reference().aurora_gf192(sample_log_n, verify=False)
            g = prover.DropinProver()
            g.aurora_gf192(sample_log_n, verify=False)
            gs, gd, acc = g.aurora_gf192(sample_log_n, verify=True)
        return {"value": (1 << sample_log_n) / cs, "unit": "constraints/s", "cores": 1, "kind": "reference",
                "sample": f"whole unmodified aurora_snark_prover at N=2^{sample_log_n} ({cs:.2f} s on 1 core; the reference has no threading); "
                          f"the full N=2^20 CPU run is the reference arm (bench.py --impl reference)",
                "same_instance_through_dropin": {"prover_s": gs, "bit_exact_transcript": cd == gd, "reference_verifier_accepts": acc == 1,
                                                 "transcript_sha256": hashlib.sha256(gd).hexdigest()}}
    except Exception as e:
        return {"unavailable": repr(e)}


def run_reference(args):
    """Reference arm: the unmodified reference aurora_snark_prover on the host CPU at the bench's own config (N = 2^20)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Reference
    if not Reference.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libiopref.so was not built (needs /root/reference at build time)"}))
        return
    t0 = time.perf_counter()
    ref = Reference()
    steps_run = max(1, min(args.steps, args.reference_proofs))
    with _quiet_stdout():
        for _ in range(args.warmup):                            # warm-up proofs on a small instance (~1 s each)
            ref.aurora_gf192(14, verify=False)
        times = ref.aurora_gf192_repeat(LOG_N, steps_run)       # one generated instance, `steps_run` proofs
    secs = float(np.median(times))
    v = (1 << LOG_N) / secs
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "constraints/s", "n_gpus": args.gpus,
           "steps": steps_run, "steps_requested": args.steps, "warmup": args.warmup, "warmup_what": "proofs at N=2^14",
           "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "gf2^192 (PCLMULQDQ)", "data": "synthetic", "config": dict(CONFIG),
           "value_scope": "whole unmodified aurora_snark_prover on the host CPU, one core (the reference has no threading); "
                          "a proof takes ~75 s, so min(steps, %d) proofs are timed (warm-up proofs run at N=2^14)" % args.reference_proofs,
           "prover_s_all": times,
           "cpu_baseline": {"value": v, "unit": "constraints/s", "cores": 1, "kind": "reference",
                            "sample": f"whole aurora_snark_prover at N=2^{LOG_N}, median of {steps_run} proof(s), single thread"},
           "e2e": {"value": v, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.concurrency > 1 and args.cpu_all_cores:
        # the reference has no threading: one proof uses one core.  For a per-socket figure, one independent prover per core (N = 2^17).
        out["cpu_all_cores"] = whole_prover_throughput("cpu", 17, runs=1, procs=os.cpu_count() or 1)
    out["wall_s"] = time.perf_counter() - t0
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-log-n", type=int, default=18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reference-proofs", type=int, default=2, help="N=2^20 CPU proofs timed by the reference arm (75 s each)")
    ap.add_argument("--cpu-all-cores", action="store_true", help="reference arm: also one CPU prover per core at N=2^17")
    ap.add_argument("--config5", action="store_true", help="also prove BASELINE config 5 (N=2^22 per GPU); default at 8 GPUs")
    ap.add_argument("--concurrency", type=int, default=1, help="prover processes sharing the GPU in the e2e_throughput section (1 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
