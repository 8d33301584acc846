#!/usr/bin/env python
"""bench.py — batched NN evals/sec of the self-play evaluation hot path (BASELINE.json metric).

A "step" is one forward pass of the engine over one batch of synthetic 19x19 positions
(workload = BASELINE.json configs[1]: 19x19, 10bx128 net, batch 256 per GPU).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, sm_100a), one process per GPU
  python bench.py --impl reference --gpus N ...            the reference's own Eigen CPU forward
                                                           (oracle/_ref, compiled from /root/reference) on host cores

One JSON line on stdout from rank 0.  `value` = whole-job evals/s with inputs resident in HBM, device-timed
with CUDA events on the engine's stream (max over ranks); `e2e` = the same metric through the C ABI
(sb_submit/sb_wait) with pinned HOST buffers, H2D + D2H inside the timed region; `roofline` = the conv3x3
tensor-core kernel against the measured dense bf16/fp16 peak; `cpu_baseline` = the reference Eigen forward
on this box's host cores (rank 0, N=1 only).  Weak scaling: every rank evaluates its own batch; the only
collective is the NCCL broadcast of the packed weight blob from rank 0 at load.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sayuri_b200 import synth  # noqa: E402

METRIC = "nn_evals_per_sec"
UNIT = "evals/s"


def algorithmic_flops_per_eval(blocks, C, P, V, n_se, S=361):
    """direct-conv, 2 flop/MAC, no Winograd discount, no padding waste."""
    conv = 2 * S * (9 * 43 * C + blocks * 2 * 9 * C * C + C * P + 5 * P + C * V + V)
    fc = 2 * (3 * P * P + 5 * P + 9 * V * V + 45 * V)
    se = n_se * 2 * (3 * C * (C // 4) + (C // 4) * 2 * C)
    return conv + fc + se


def conv3x3_flops_per_eval(blocks, C, P, V, S=361):
    """Algorithmic flops of what the dominant kernel computes: the input and tower 3x3 convolutions and the
    head-entry 1x1 convolutions (one single-tap launch of the same kernel)."""
    return 2 * S * (9 * 43 * C + blocks * 2 * 9 * C * C + C * (P + V))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs ~100 ms to print its first line) and every line is stamped with the
    host time it arrived; only samples that fall inside a marked timed window are reported (if a window was too short
    to catch one, the samples taken under load around it are used and the line says so)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []      # (host time, text)
        self.windows = []    # (t0, t1)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 3.0:   # first line = the sampler is live
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                rows.append((t, float(f[0]), float(f[1]), float(f[2]), [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if any(t0 <= r[0] <= t1 + 0.02 for t0, t1 in self.windows)]
        how = "samples inside the timed regions"
        if not inside and rows:
            pmax = max(r[3] for r in rows)
            inside = [r for r in rows if r[3] >= 0.6 * pmax]
            how = "timed regions shorter than the sampling period: samples under load (>= 60 % of peak power) around them"
        reasons = sorted({n for r in inside for n in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": max(r[2] for r in rows) if rows else None, "reasons": reasons, "samples": len(inside),
                "power_w_max": max(r[3] for r in inside) if inside else None, "how": how}


def workload_config(args, P, V):
    """The SAME dict in both arms (the driver compares them): the workload, and how each arm treats caches / precision."""
    return {"workload": "19x19 board, %s net (P=%d,V=%d, SE every 3rd block, mish), batch-%d NN forward per GPU" % (args.net, P, V, args.batch),
            "net": args.net, "board": 19, "batch_per_gpu": args.batch,
            "precision": "reference arm: fp32 Eigen; our arm: --precision (default fp32_split = fp32-faithful fp16 hi/lo split, fp32 accumulate)",
            "l2": "our arm: 256 MiB buffer written between timed iterations (L2 flush); reference arm: CPU, not applicable",
            "parallelism": "our arm: replica per GPU, weights broadcast over NCCL/NVLink; reference arm: host threads"}


def weights_file(net, seed=20260417):
    path = os.path.join(tempfile.gettempdir(), "sb_bench_%s_seed%d.bin" % (net, seed))
    if not os.path.exists(path):
        tmp = path + ".%d.tmp" % os.getpid()
        synth.write_synth_net(tmp, net, seed=seed)
        os.replace(tmp, path)
    return path


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = unmodified
    loader.cc + blas_forward_pipe.cc + Eigen), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle.oracle_py import Reference
    blocks, C, P, V = synth.NETS[args.net]
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0, "config": workload_config(args, P, V)}
    if not Reference.available():
        line["unavailable"] = "oracle/_ref is not built (needs /root/reference at build time)"
        print(json.dumps(line), flush=True)
        return
    ref = Reference(weights_file(args.net), winograd=True)   # reference default (config.cc:32)
    n_pos = 32
    planes = synth.synth_positions(n_pos, 19, seed=20260418)
    sample_s = min(args.ref_seconds, max(0.5, 150.0 / max(args.steps, 1)))   # whole run stays within a few minutes
    for _ in range(min(args.warmup, 1)):
        ref.time_forward(planes, n_pos, 19, cores, 1.0)
    total_n, total_t = 0, 0.0
    for _ in range(args.steps):
        n, t = ref.time_forward(planes, n_pos, 19, cores, sample_s)
        total_n += n
        total_t += t
    value = total_n / total_t
    line.update({"value": value, "ms_per_step": 1e3 * args.batch / value,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                                  "sample": "%d steps x %.1f s of BlasForwardPipe::Forward (Winograd, Eigen %s build) on %d threads, %d synthetic positions round-robin"
                                            % (args.steps, sample_s, Reference.variant, cores, n_pos)},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    from sayuri_b200 import engine
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    blocks, C, P, V = synth.NETS[args.net]
    stack = synth.default_stack(blocks)
    n_se = sum(1 for s in stack if s.endswith("-SE"))
    B = args.batch
    precision = {"fp32_split": engine.PRECISION_FP32_SPLIT, "fp16": engine.PRECISION_FP16}[args.precision]

    # ---- weights: rank 0 parses + packs, everyone else receives the packed blob over NCCL ----------
    pipe = engine.B200ForwardPipe()
    if rank == 0:
        pipe.initialize(weights_file(args.net), 19, B, gpus=[local_rank], precision=precision)
    else:
        desc = dict(version=5, blocks=blocks, channels=C, P=P, V=V, activation=5,
                    se_sizes=[C // 4 if s.endswith("-SE") else 0 for s in stack])
        pipe.initialize_from_tensors(desc, None, 19, B, gpus=[local_rank], precision=precision)
    for kv in args.option:
        k, v = kv.split("=")
        pipe.set_option(k, int(v))
    if world > 1:
        from sayuri_b200.dist import replicate_weights
        replicate_weights(pipe, dist, rank, torch.device("cuda", local_rank))   # the path's only collective

    # ---- inputs: synthetic positions in pinned host memory (two alternating batches) ----------------
    pinned = engine.PinnedArray((2, B, engine.PLANE_FLOATS))
    pos = synth.synth_positions(min(B, 64), 19, seed=20260418 + rank).reshape(-1, engine.PLANE_FLOATS)
    for k in range(2):
        for i in range(B):
            pinned.array[k, i] = pos[(i + 7 * k) % pos.shape[0]]
    sizes = [19] * B
    offs = [0] * B
    outs = [np.zeros(B, dtype=engine.OUTPUT_DTYPE) for _ in range(2)]

    def barrier():
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # upload once: inputs resident in HBM for the device-timed loop
    pipe.submit(0, 0, pinned.array[0], sizes, offs)
    pipe.wait(0, 0, outs[0])
    pipe.submit(0, 1, pinned.array[1], sizes, offs)
    pipe.wait(0, 1, outs[1])
    if not np.isfinite(outs[0]["probabilities"]).all():
        raise RuntimeError("non-finite outputs")

    # ---- (1) device-timed, inputs resident: warm-up W, then exactly K steps -------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    pipe.time_forward(0, 0, args.warmup, flush_l2=True)
    barrier()
    launches0 = pipe.launch_count()
    tw0 = time.perf_counter()
    ms, _, _ = pipe.time_forward(0, 0, args.steps, flush_l2=True)
    sampler.window(tw0, time.perf_counter())
    launches = pipe.launch_count() - launches0
    barrier()
    t_dev = reduce_max(float(ms.sum()) / 1e3)
    value = world * B * args.steps / t_dev

    # ---- roofline of the dominant kernel (conv3x3_tc), events around every launch, separate pass ----
    # (sb_time_forward brackets the NON-convolution kernels of a forward with CUDA events on the engine's stream and
    #  subtracts them from the forward's duration, median of 5 forwards, L2 flushed: the conv launches keep their
    #  programmatic-dependent-launch overlap exactly as in the timed steps)
    prof_ms, conv_ms_prof, conv_n = pipe.time_forward(0, 0, 0, flush_l2=True, profile_conv=True)
    # the profiling pass runs under event overhead and a different power state: take its SHARE of conv time and apply
    # it to the timed steps' own duration
    conv_share = conv_ms_prof / float(prof_ms[0]) if len(prof_ms) and prof_ms[0] > 0 else None
    conv_ms = conv_share * float(ms.mean()) if conv_share else conv_ms_prof
    peaks = measured_peaks()
    conv_flops = conv3x3_flops_per_eval(blocks, C, P, V) * B
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
    peak = peaks["tflops_sustained"]
    # dram bytes per launch come from the committed `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum of
    # one tower-conv launch); the entry names the sha256 of the kernel source it was taken from and is dropped (null) when
    # that is not the source this library was built from, so the number cannot silently go stale
    traffic, traffic_note = None, "no committed capture for this kernel source"
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tpath):
        import hashlib
        with open(os.path.join(ROOT, "sayuri_b200", "csrc", "conv3x3_tc2.cuh"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        with open(tpath) as f:
            ent = json.load(f).get(args.precision, {})
        if ent.get("kernel_source_sha256") == sha:
            traffic, traffic_note = ent.get("dram_bytes_per_launch"), ent.get("source", "")
        elif ent:
            traffic_note = "committed capture is of an older kernel source (sha mismatch): dropped"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "conv3x3_tc2_kernel<%s, mish> (tcgen05 cta_group::2)" % ("split" if precision == engine.PRECISION_FP32_SPLIT else "fp16"),
                "launches_per_step": conv_n, "launches_note": "convolutions per step; the engine chains consecutive ones into fewer kernel launches (option conv_chain), `traffic` and algorithmic_flops_per_launch_avg are per convolution",
                "kernel_ms_per_step": conv_ms,
                "kernel_share_of_step": conv_share,
                "algorithmic_flops_per_launch_avg": conv_flops / max(conv_n, 1),
                "peak_source": peaks["source"] + ", dense bf16/fp16 sustained; kernel timed inside the step",
                "how": "achieved = algorithmic flops of the %d convolutions of a step / (kernel_share_of_step x ms_per_step); the share comes from extra L2-flushed forwards (median of 5) whose non-conv kernels are bracketed with CUDA events on the engine's stream, so the conv launches keep their dependent-launch overlap" % conv_n,
                "tensor_flops_issued_per_algorithmic": 3 * (400.0 / 361.0) if precision == engine.PRECISION_FP32_SPLIT else (400.0 / 361.0),
                "note": "fp32-faithful rung issues 3 fp16 MMAs per algorithmic MAC (hi*hi + lo*hi + hi*lo) on a 400-row/361-cell canvas"}

    # ---- (2) end to end through the C ABI, 2 slots pipelined: host buffers in, host results out ---------------------
    #      headline `e2e`: inputs in ordinary (pageable) host memory, as the front-end holds them; sb_submit packs every
    #      position exactly into its 2.2 KB record in pinned staging (the product path: what crosses PCIe is the record);
    #      `e2e_raw_fp32`: the same loop with 62 KB of fp32 planes per position DMA'd in place from pinned memory.
    def e2e_loop(steps, src):
        for s in range(steps):
            slot = s & 1
            if s >= 2:
                pipe.wait(0, slot, outs[slot])
            pipe.submit(0, slot, src[slot], sizes, offs)
        for s in range(max(steps - 2, 0), steps):
            pipe.wait(0, s & 1, outs[s & 1])

    def e2e_measure(src):
        e2e_loop(max(args.warmup, 2), src)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(args.steps, src)
        t = time.perf_counter() - t0
        sampler.window(t0, t0 + t)
        barrier()
        return world * B * args.steps / reduce_max(t)

    pageable = [np.array(pinned.array[k]) for k in range(2)]    # plain numpy memory: sb_submit takes its packing path
    v_packed = e2e_measure(pageable)
    v_raw = e2e_measure([pinned.array[0], pinned.array[1]])
    out_bytes = B * (2 * 361 + 8) * 4
    e2e = {"value": v_packed, "unit": UNIT, "h2d_bytes_per_step": B * 2252 + 2 * B * 4, "d2h_bytes_per_step": out_bytes,
           "input": "fp32 planes in pageable host memory, packed exactly to 2252-byte records by sb_submit (4 host threads), records DMA'd from pinned staging",
           "timing": "host perf_counter around K pipelined sb_submit/sb_wait steps (2 slots), device idle on both sides"}
    e2e_raw = {"value": v_raw, "unit": UNIT, "h2d_bytes_per_step": B * engine.PLANE_FLOATS * 4 + 2 * B * 4, "d2h_bytes_per_step": out_bytes,
               "input": "fp32 planes in pinned host memory (sb_host_alloc), DMA'd in place"}

    # ---- (3) the single-position call of the plugin interface (NetworkForwardPipe::Forward -> sb_eval): T native
    #      host threads, fp32 planes in pageable memory, packing + H2D + forward + D2H + wake-up inside the region ----
    e2e_eval = None
    if args.eval_threads > 0:
        try:
            pipe.batcher_config(B, 200)
            n_thr = args.eval_threads
            tw0 = time.perf_counter()
            ev = pipe.eval_throughput(pos, 19, n_thr, args.eval_seconds)
            sampler.window(tw0 + 0.2, time.perf_counter())
            st = pipe.batcher_stats()
            e2e_eval = {"value": world * ev, "unit": UNIT, "host_threads": n_thr, "seconds": args.eval_seconds,
                        "mean_batch": st["positions"] / max(st["batches"], 1),
                        "h2d_bytes_per_eval": 2252, "d2h_bytes_per_eval": (2 * 361 + 8) * 4,
                        "call": "sb_eval (blocking, one position per call; the engine's batcher forms the batches)"}
        except RuntimeError as ex:
            e2e_eval = {"value": None, "error": str(ex)}
    clocks = sampler.stop()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16x3-split/f32-accum" if precision == engine.PRECISION_FP32_SPLIT else "f16/f32-accum",
            "data": "synthetic",
            "config": workload_config(args, P, V), "precision": args.precision, "options": args.option,
            "clocks": clocks, "e2e": e2e, "e2e_raw_fp32": e2e_raw, "e2e_eval": e2e_eval, "gpu_launches": int(launches), "roofline": roofline,
            "algorithmic_gflop_per_eval": algorithmic_flops_per_eval(blocks, C, P, V, n_se) / 1e9}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's Eigen forward on host cores -----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle.oracle_py import Reference
            if Reference.available():
                cores = os.cpu_count() or 1
                ref = Reference(weights_file(args.net), winograd=True)
                rp = synth.synth_positions(32, 19, seed=20260418)
                n, t = ref.time_forward(rp, 32, 19, cores, args.cpu_seconds)
                line["cpu_baseline"] = {"value": n / t, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": "%.0f s of BlasForwardPipe::Forward (Winograd, Eigen %s build) on %d threads over 32 of the same synthetic positions"
                                                  % (args.cpu_seconds, Reference.variant, cores)}
            else:
                from oracle.oracle_py import Oracle
                orc = Oracle(weights_file(args.net))
                rp = synth.synth_positions(8, 19, seed=20260418)
                t0 = time.perf_counter()
                k = 0
                while time.perf_counter() - t0 < args.cpu_seconds:
                    orc.forward(rp[k % 8], 19, 0)
                    k += 1
                line["cpu_baseline"] = {"value": k / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": "%.0f s of the scalar C oracle on 1 thread" % args.cpu_seconds}
        except Exception as ex:  # the baseline is reported, never load-bearing
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}
    pinned.free()
    pipe.destroy()
    if dist is not None:
        dist.barrier()
    if rank == 0 and args.selfplay_games > 0:
        # the other half of BASELINE.json's metric: self-play games/hour through the UNMODIFIED reference loop (19x19, 400
        # visits) over our pipe, ONE process driving all N GPUs like the reference front-end (tools/selfplay_bench.py)
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "selfplay_bench.py"), "--preset", "config2", "--net", args.net,
                                "--gpus", ",".join(str(g) for g in range(world)), "--parallel-games", str(args.selfplay_games * world),
                                "--timeout", "3000"] + (["--fp16"] if args.precision == "fp16" else []),
                               capture_output=True, text=True, timeout=3100)
            sp = json.loads(r.stdout.strip().splitlines()[-1])
            line["selfplay"] = sp
            line["games_per_hour"] = sp["games_per_hour"]
        except Exception as ex:
            line["selfplay"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--net", default="10bx128", choices=sorted(synth.NETS))
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default="fp32_split", choices=["fp32_split", "fp16"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=3.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-threads", type=int, default=512, help="host threads of the sb_eval leg (0 = skip)")
    ap.add_argument("--eval-seconds", type=float, default=2.0)
    ap.add_argument("--selfplay-games", type=int, default=0, metavar="G",
                    help="also play G parallel self-play games per GPU to the end through the unmodified reference loop (19x19, "
                         "400 visits) and add games_per_hour to the line; takes minutes, off by default")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE",
                    help="engine knob for A/B runs (sb_set_option), e.g. --option chunk_taps=3; recorded in config")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
