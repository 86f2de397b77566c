#!/usr/bin/env python
"""bench.py - KING N x N pair-dot-products/s on the BASELINE.json workload (100k samples x 1M SNPs).

A "step" is one pass of the hot path (the KING pair-count kernel, CalcKing's dense loop) over one
batch of `--batch-variants` synthetic variants for ALL sample pairs of this rank's row block; the
counts accumulate in HBM across steps exactly as in a full run (16 steps of 65,536 variants = one
1M-SNP job).

  value   = (pairs x variants_per_step / 1e6) / step time: full-length (1M-SNP) pair-dot-products per
            second, whole job over all ranks, inputs already resident in HBM.
  e2e     = same metric through the C-ABI with HOST buffers: every step copies the step's genotype
            batch from pinned host memory and reads back 1/steps_per_job of the fp64 kinship matrix.
  roofline= int8 tensor pipe: 5 products x 2 ops x pairs x variants / kernel time (CUDA events on the
            library's stream) vs the nominal dense int8 rate (4.5 POP/s); fractions of 2 x the measured
            bf16 cuBLAS rates of MEASURED_PEAKS.json are reported beside it.
  cpu_baseline / --impl reference = the UNMODIFIED reference binary (oracle/_ref/plink2,
            --make-king-table, all host threads) on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): rows of the output triangle are split into equal-area
blocks (ParallelBounds, the reference's --parallel); each rank synthesises 1/G of the step's
variants and ONE NCCL all_gather per step assembles the full genotype tile everywhere (inside the
timed region); outputs stay local.  Total work is fixed => "strong" scaling.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_N = 100_000
FULL_M = 1_000_000
METRIC = "KING NxN pair-dot-products/sec (100k samples x 1M SNPs)"
UNIT = "pair-dot-products/s (1M-SNP pairs)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=FULL_N)
    ap.add_argument("--batch-variants", type=int, default=65536)
    ap.add_argument("--algo", default="tensor_ts", choices=["tensor_ts", "tensor", "popcount"])
    ap.add_argument("--cpu-samples", type=int, default=16384)
    ap.add_argument("--cpu-variants", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short GRM-kernel measurement reported under `secondary`")
    return ap.parse_args()


# ----------------------------------------------------------------------------------- synthetic data
def synth_genovecs(torch, n, v0, v1, device, seed=20260923, miss=0.01):
    """uint8 [v1-v0, 8*ceil(n/32)] PgrGet-layout rows: HWE genotypes, per-variant ALT freq ~U(.02,.98),
    1% missing (the shape of the reference's --dummy generator, 2.0/plink2_import.cc:16326-16460).
    Deterministic in (seed, variant index) so every rank can synthesise its own slice."""
    row_bytes = (n + 31) // 32 * 8
    out = torch.empty((v1 - v0, row_bytes), dtype=torch.uint8, device=device)
    n4 = row_bytes * 4
    chunk = max(1, min(v1 - v0, (1 << 28) // max(n4, 1)))
    for s in range(v0, v1, chunk):
        e = min(v1, s + chunk)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + s)
        f = torch.rand((e - s, 1), generator=g, device=device) * 0.96 + 0.02
        code = (torch.rand((e - s, n4), generator=g, device=device) < f).to(torch.uint8)
        code += (torch.rand((e - s, n4), generator=g, device=device) < f).to(torch.uint8)
        code[torch.rand((e - s, n4), generator=g, device=device) < miss] = 3
        if n4 > n:
            code[:, n:] = 0
        q = code.view(e - s, row_bytes, 4)
        out[s - v0 : e - v0] = q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)
        del code, q, f
    return out


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- reference CPU arm
def run_reference_sample(n, m, threads, workdir, keep_input=None, binary=None):
    """Times the unmodified reference binary on an n x m synthetic .bed: `--make-king-table` with a
    table filter so the text output stays small (the N^2 M/64 popcount loop is unaffected).
    Returns (seconds, pair_snp_per_s)."""
    import numpy as np
    import torch

    plink2 = binary or os.path.join(ROOT, "oracle", "_ref", "plink2")
    if not os.path.exists(plink2):
        raise FileNotFoundError(f"{plink2} missing (oracle/build_ref.sh builds it where /root/reference exists)")
    prefix = keep_input or os.path.join(workdir, f"cpu_{n}_{m}")
    if not os.path.exists(prefix + ".bed"):
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        # PgrGet codes -> .bed codes (0 homALT,1 missing,2 het,3 homREF; pgen_spec.tex:436-438)
        lut = torch.tensor([3, 2, 0, 1], dtype=torch.uint8, device=dev)
        bpv = (n + 3) // 4
        with open(prefix + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            for s0 in range(0, m, 2048):
                s1 = min(m, s0 + 2048)
                by = synth_genovecs(torch, n, s0, s1, dev)[:, :bpv].contiguous()
                c = torch.stack([(by >> sh) & 3 for sh in (0, 2, 4, 6)], dim=-1).to(torch.int32)
                b = lut[c]
                if n % 4:
                    b.view(s1 - s0, -1)[:, n:] = 0
                bed = (b[..., 0] | (b[..., 1] << 2) | (b[..., 2] << 4) | (b[..., 3] << 6)).to(torch.uint8).cpu().numpy()
                bed.tofile(f)
                del by, c, b, bed
        with open(prefix + ".bim", "w") as f:
            f.write("".join(f"1\tsnp{k}\t0\t{k + 1}\tA\tG\n" for k in range(m)))
        with open(prefix + ".fam", "w") as f:
            f.write("".join(f"0\tper{k}\t0\t0\t2\t-9\n" for k in range(n)))
    t0 = time.perf_counter()
    r = subprocess.run([plink2, "--bfile", prefix, "--make-king-table", "--king-table-filter", "0.35", "--threads", str(threads), "--memory", "64000", "--out", prefix + ("_out" if binary is None else "_b200")], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference binary failed: " + r.stdout[-500:] + r.stderr[-500:])
    pairs = n * (n - 1) // 2
    return dt, pairs * m / dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n, m = args.cpu_samples, args.cpu_variants
    tmp = tempfile.mkdtemp(prefix="pl2ref_")
    try:
        times = []
        for it in range(args.warmup + args.steps):
            dt, _ = run_reference_sample(n, m, threads, tmp, keep_input=os.path.join(tmp, "in"))
            if it >= args.warmup:
                times.append(dt)
        t = sum(times) / len(times)
        pairs = n * (n - 1) // 2
        val = pairs * m / 1e6 / t
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 popcount (AVX2)", "data": "synthetic",
            "config": {"workload": "plink2 --make-king-table, bounded sample of the 100k x 1M job", "samples": n, "variants": m, "threads": threads, "pair_snp_per_s": pairs * m / t},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"{n} samples x {m} variants, whole `plink2 --make-king-table` run incl. .bed load"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------- B200 arm
def b200_arm(args):
    import torch
    import torch.distributed as dist

    import plink_ng_b200 as p
    from plink_ng_b200.host import KING_ALGO_POPCOUNT, KING_ALGO_TENSOR, KING_ALGO_TENSOR_TS, KingJob
    from plink_ng_b200.sharding import assemble_block, pairs_in_rows, row_block, variant_slice

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.samples
    mb = args.batch_variants
    row_bytes = (n + 31) // 32 * 8
    r0, r1 = row_block(n, rank, world)
    my_pairs = pairs_in_rows(r0, r1)
    total_pairs = n * (n - 1) // 2
    algo = {"tensor_ts": KING_ALGO_TENSOR_TS, "tensor": KING_ALGO_TENSOR, "popcount": KING_ALGO_POPCOUNT}[args.algo]

    # this rank's slice of the step's variants; all_gather assembles the tile (north_star)
    per, v0, v1 = variant_slice(mb, rank, world)
    slice_dev = torch.zeros((per, row_bytes), dtype=torch.uint8, device=dev)
    if v1 > v0:
        slice_dev[: v1 - v0] = synth_genovecs(torch, n, v0, v1, dev)
    full = slice_dev  # world == 1; otherwise re-assembled by the all_gather of every step
    torch.cuda.synchronize()

    ctx = p.GpuContext(local_rank)
    job = KingJob(ctx, n, r0, r1, algo)
    ctx.synchronize()
    ext_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def step_resident():
        nonlocal full
        with torch.cuda.stream(ext_stream):
            full = assemble_block(dist, torch, slice_dev, per, world)
        job.add_variants_device(full.data_ptr(), row_bytes, mb)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    ctx.event_record(0)
    for _ in range(args.steps):
        step_resident()
    ctx.event_record(1)
    dev_ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    step_ms = float(t_ms.item()) / args.steps
    value = total_pairs * (mb / 1e6) / (step_ms * 1e-3)

    # kernel-only time for the roofline: the same steps without the collective (N=1: identical)
    ctx.event_record(2)
    for _ in range(max(1, args.steps // 2)):
        job.add_variants_device(full.data_ptr(), row_bytes, mb)
    ctx.event_record(3)
    kern_ms = ctx.event_elapsed_ms(2, 3) / max(1, args.steps // 2)

    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        host_batch = torch.empty((mb, row_bytes), dtype=torch.uint8, pin_memory=True)
        host_batch.copy_(full[:mb])
        steps_per_job = max(1, FULL_M // mb)
        rows = r1 - r0
        slices = [(r0 + rows * k // steps_per_job, r0 + rows * (k + 1) // steps_per_job) for k in range(steps_per_job)]
        max_pairs = max((b * (b - 1) - a * (a - 1)) // 2 for a, b in slices)
        host_out = torch.empty((max_pairs,), dtype=torch.float64, pin_memory=True)
        import ctypes as C
        from plink_ng_b200.capi import lib, check

        def step_e2e(k):
            check(lib.pl2gpu_king_add_variants(job._h, C.c_void_p(host_batch.data_ptr()), row_bytes, mb, 0), "add_variants(host)")
            a, b = slices[k % steps_per_job]
            check(lib.pl2gpu_king_get_kinship(job._h, a, b, C.c_void_p(host_out.data_ptr()), 0), "get_kinship(host)")
            return (b * (b - 1) - a * (a - 1)) // 2 * 8

        step_e2e(0)
        barrier()
        t1 = time.perf_counter()
        d2h = 0
        for k in range(args.steps):
            d2h += step_e2e(k)
        barrier()
        e_ms = torch.tensor([(time.perf_counter() - t1) * 1e3], dtype=torch.float64, device=dev)
        d2h_t = torch.tensor([float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(d2h_t, op=dist.ReduceOp.SUM)
        e_step_ms = float(e_ms.item()) / args.steps
        e2e = {
            "value": total_pairs * (mb / 1e6) / (e_step_ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(mb) * int(row_bytes) * world, "d2h_bytes_per_step": int(d2h_t.item() / args.steps),
            "ms_per_step": e_step_ms,
        }

    job.close()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- secondary (not the headline metric): the GRM tensor kernel of the same path, resident inputs ----
    secondary = None
    if world == 1 and not args.no_secondary:
        try:
            import numpy as np
            from plink_ng_b200.host import GrmJob
            n2, m2 = min(n, 16384), 65536
            g2 = synth_genovecs(torch, n2, 0, m2, dev)
            torch.cuda.synchronize()
            rf = np.random.default_rng(0).uniform(0.05, 0.95, size=m2)
            with GrmJob(ctx, n2) as gj:
                gj.add_variants_device(g2.data_ptr(), g2.shape[1], m2, ref_freqs=rf)
                ctx.synchronize()
                ctx.event_record(6)
                for _ in range(3):
                    gj.add_variants_device(g2.data_ptr(), g2.shape[1], m2, ref_freqs=rf)
                ctx.event_record(7)
                g_ms = ctx.event_elapsed_ms(6, 7) / 3
            tops = 11 * 2 * (n2 * (n2 + 1) // 2) * m2 / (g_ms * 1e-3) / 1e12
            secondary = {"grm": {"kernel": "grm_ts_kernel", "workload": f"{n2} samples x {m2} variants per add_variants call (tables + re-tiling + tensor kernel), inputs resident",
                                 "ms_per_call": g_ms, "achieved": tops, "unit": "TOP/s (int8; 10 digit planes + obs = 11 products x 2 ops per pair and variant)", "frac_of_nominal_4500": tops / 4500.0,
                                 "pair_snp_per_s": (n2 * (n2 + 1) // 2) * m2 / (g_ms * 1e-3)}}
            del g2
        except Exception as ex:  # reported, never faked
            secondary = {"grm": {"error": str(ex)[-300:]}}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # int8 tensor peak.  MEASURED_PEAKS.json holds no int8 figure, only cuBLAS bf16 (burst / sustained, the
    # latter power-capped at ~1.37 GHz).  tcgen05 kind::i8 retires 8192 MAC/clk/SM (128 x N x 32 in N/2 clk,
    # B300_MICROARCH.md "tcgen05 floor"): 148 SMs x 16384 op/clk x 1.965 GHz = 4.77 POP/s, NVIDIA's dense
    # int8/fp8 nominal is 4.5 POP/s.  `peak` is the nominal 4500; the bf16-derived figures are given beside it.
    bf16_burst = peaks.get("bf16_tflops") or 1590.0
    bf16_sust = peaks.get("bf16_tflops_sustained") or 1400.0
    peak = 4500.0
    peak_src = ("nominal dense int8 4.5 POP/s (no measured int8 peak exists; 2 x measured bf16 would be "
                f"{2 * bf16_burst:.0f} burst / {2 * bf16_sust:.0f} sustained TOP/s, which this kernel exceeds because cuBLAS bf16 is power-capped)")
    ops = 5 * 2 * my_pairs * mb  # algorithmic: 5 indicator products per pair and variant
    achieved = ops / (kern_ms * 1e-3) / 1e12
    kname = {KING_ALGO_TENSOR_TS: "king_ts_kernel", KING_ALGO_TENSOR: "king_tc_kernel", KING_ALGO_POPCOUNT: "king_popc_kernel"}[algo]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s (int8)", "frac": achieved / peak, "traffic": None,
                "kernel": kname, "kernel_ms": kern_ms, "peak_source": peak_src, "frac_of_2x_bf16_burst": achieved / (2 * bf16_burst),
                "frac_of_2x_bf16_sustained": achieved / (2 * bf16_sust), "algorithmic_ops_per_launch": ops,
                "traffic_note": "not captured at this size; ncu at 16384 samples x 65536 variants: dram read+write 8.85 GB per launch vs 5.98 GB algorithmic (accumulators once each way + 2-bit operands), profiles/r01_ncu_king_ts_v3.md",
                "kernel_ms_note": "one add_variants call = pad + two re-tiling launches (~2%) + the tensor kernel, CUDA events on the library's stream"}
    if algo == KING_ALGO_POPCOUNT:
        roofline["note"] = "popcount kernel: int8-equivalent ops shown for comparability; its own limiter is the POPC pipe"

    cpu_baseline = None
    cli = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            tmp = tempfile.mkdtemp(prefix="pl2cpu_")
            threads = os.cpu_count() or 1
            dt, rate = run_reference_sample(args.cpu_samples, args.cpu_variants, threads, tmp)
            cpu_baseline = {"value": rate / 1e6, "unit": UNIT, "cores": threads, "kind": "reference", "seconds": dt,
                            "sample": f"{args.cpu_samples} samples x {args.cpu_variants} variants, one whole `plink2 --make-king-table --threads {threads}` run (incl. .bed load) of oracle/_ref/plink2"}
            # the same command line through OUR host program (the drop-in face), same files, outputs compared byte for byte
            try:
                ours = os.path.join(ROOT, "plink_ng_b200", "plink2_b200")
                dt2, rate2 = run_reference_sample(args.cpu_samples, args.cpu_variants, threads, tmp, binary=ours)
                pre = os.path.join(tmp, f"cpu_{args.cpu_samples}_{args.cpu_variants}")
                same = open(pre + "_out.kin0", "rb").read() == open(pre + "_b200.kin0", "rb").read()
                cli = {"seconds": dt2, "value": rate2 / 1e6, "unit": UNIT, "speedup_vs_reference_run": dt / dt2, "kin0_identical_to_reference": same,
                       "command": "plink2_b200 --bfile <same> --make-king-table --king-table-filter 0.35 (process start to exit: CUDA init, .bed load, H2D, kernels, table write)"}
            except Exception as ex:
                cli = {"error": str(ex)[-300:]}
            shutil.rmtree(tmp, ignore_errors=True)
        except Exception as ex:  # the baseline is reported, never silently faked
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 popcount" if algo == KING_ALGO_POPCOUNT else "s8 (exact int32 accumulate)",
        "data": "synthetic",
        "config": {"workload": f"--make-king, {n} samples x {FULL_M} SNPs in steps of {mb} variants", "samples": n, "variants_per_step": mb, "steps_per_full_job": FULL_M // mb,
                   "parallelism": f"row-block x{world} (ParallelBounds), 1 all_gather/step" if world > 1 else "single GPU", "algo": args.algo,
                   "l2_policy": "inputs (1.6 GB batch + accumulators) exceed L2; no flush needed", "pair_snp_per_s": total_pairs * mb / (step_ms * 1e-3), "wall_ms_per_step": wall_ms / args.steps},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "cli_same_files": cli, "secondary": secondary, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
