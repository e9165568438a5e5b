#!/usr/bin/env python
"""Throughput benchmark of the Whisper-AT tagging hot path (mel -> encoder -> TL-TR) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model large-v2] [--batch B]

One process per GPU (the driver launches N>1 through torch.distributed.run); a "step" is one pass of the hot
path over one batch of synthetic 30 s clips.  Default workload = the configuration BASELINE.json's metric is quoted
on: large-v2 + full TL-TR head, 128-bin mel, at_time_res=10, 128 clips per GPU (1024 across 8), bf16 operands.
Rank 0 prints ONE JSON line; see DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "whisper-at_b200"))
sys.dont_write_bytecode = True

import torch  # noqa: E402

METRIC = "audio-sec/sec, large-v2 mel+encoder+TL-TR tagging"
UNIT = "audio-s/s"


def flops_per_clip(d: int, L: int, n_mels: int, low: bool, res: float):
    """SURVEY.md §8d formulas (2*M*N*K; attention 4*T^2*d per layer)."""
    T = 1500
    gemm_enc = 2 * 3000 * d * 3 * n_mels + 2 * T * d * 3 * d + L * (2 * T * 4 * d * d + 2 * T * 8 * d * d)
    attn_enc = L * 4 * T * T * d
    di = 512 if low else d
    dw = int(res * 2.5)
    S = math.ceil(75 / dw)
    N = S * L * dw
    gemm_head = (2 * N * d * 512 if low else 0) + N * 24 * di * di + S * L * 24 * di * di + 2 * S * di * 527
    attn_head = S * L * 4 * dw * dw * di + S * 4 * L * L * di
    return dict(gemm=gemm_enc + gemm_head, attn=attn_enc + attn_head, total=gemm_enc + attn_enc + gemm_head + attn_head,
                encoder=gemm_enc + attn_enc)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=float(j["bf16_tflops_sustained"]), hbm=float(j["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1590.0, hbm=6650.0, source="of fallback (B200_PROFILING.md: no MEASURED_PEAKS.json)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def oracle_time_clips(name, n_mels, low, res, n_clips, threads):
    """The CPU arm: the oracle port of the reference (oracle/wat_oracle.py), fp32, one clip per call as the
    reference's AT output requires, on `threads` host threads.  Returns seconds per clip (list)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wat_oracle as O
    from whisper_at import synth
    torch.set_num_threads(threads)
    d, h, L = synth.MODEL_SHAPES[name]
    sd = synth.synth_state_dict(n_mels, d, L, low, seed=1, init="lively")
    times = []
    with torch.no_grad():
        for i in range(n_clips):
            clip = synth.synth_clip(1 + i)
            t0 = time.perf_counter()
            O.tag(clip[None], sd, h, n_mels, res)
            times.append(time.perf_counter() - t0)
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="large-v2")
    ap.add_argument("--n-mels", type=int, default=None)
    ap.add_argument("--low", action="store_true")
    ap.add_argument("--res", type=float, default=10)
    ap.add_argument("--batch", type=int, default=128, help="clips per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=0, help="clips per internal chunk of the library (0 = the whole per-GPU batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--allow-short-warmup", action="store_true", help="profiling runs only: do not force W >= 3")
    args = ap.parse_args()
    n_mels = args.n_mels or (128 if args.model == "large-v2" else 80)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from whisper_at import synth
    d, h, L = synth.MODEL_SHAPES[args.model]
    fl = flops_per_clip(d, L, n_mels, args.low, args.res)
    workload = f"whisper-at {args.model} {'TL-TR-512' if args.low else 'TL-TR'} n_mels={n_mels} at_time_res={args.res:g}, 30 s synthetic clips"
    config = dict(workload=workload, clips_per_gpu=args.batch, clips_per_internal_chunk=(args.chunk or args.batch), global_batch=args.batch * world, at_time_res=args.res,
                  parallelism=f"dp{world} (clips sharded, no data-path collective; NCCL all_gather of logits)",
                  l2="inputs+activations per step >> 126 MB L2 (no flush needed)", flops_per_clip=fl["total"])

    # ------------------------------------------------------------------ reference arm: CPU, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        budget = 240.0
        t_first = oracle_time_clips(args.model, n_mels, args.low, args.res, 1, threads)[0]       # warm-up step (also sizes the run)
        steps = max(1, min(args.steps, int((budget - t_first) / max(t_first, 1e-3))))
        times = oracle_time_clips(args.model, n_mels, args.low, args.res, steps, threads)
        total = sum(times)
        val = 30.0 * steps / total
        out = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=1, ms_per_step=1000 * total / steps,
                   higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                   config=dict(config, clips_per_step=1, steps_requested=args.steps),
                   cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind="port",
                                     sample=f"{steps} step(s) of 1 clip each (the reference's AT path is batch-1), oracle/wat_oracle.py, torch CPU fp32"),
                   e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(out))
        return

    # ------------------------------------------------------------------ our arm
    import whisper_at
    from whisper_at import _lib
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = whisper_at.ModelDimensions(n_mels, 1500, d, h, L, 51865, 448, d, h, L)
    model = whisper_at.Whisper(dims, at_low_compute=args.low, precision=args.precision, max_batch=(args.chunk or args.batch))
    model.load_state_dict(synth.synth_state_dict(n_mels, d, L, args.low, seed=1, init="lively"), strict=False)
    model = model.to(f"cuda:{local}")
    B = args.batch
    # distinct synthetic clips per rank; a few base clips rolled in time keep host generation short
    base = synth.synth_batch(min(B, 8), start=1 + 8 * rank)
    audio_host = torch.stack([torch.roll(base[i % base.shape[0]], 1600 * (i // base.shape[0])) for i in range(B)]).pin_memory()
    audio_dev = audio_host.cuda(non_blocking=True)
    dw = int(args.res * 2.5)
    S = math.ceil(75 / dw)
    gathered = torch.empty((world * B, S, 527), device="cuda") if world > 1 else None
    eng = model.engine()
    Lb = _lib.lib()

    def step_device():
        lg = model.tag_batch(audio_dev, at_time_res=args.res)
        if world > 1:
            dist.all_gather_into_tensor(gathered, lg)
        return lg

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    n_warm = args.warmup if args.allow_short_warmup else max(args.warmup, 3)
    for _ in range(n_warm):
        step_device()
    sync_all()
    launches0 = model.kernel_launches()
    _lib.check(Lb.wat_profile(eng.h, 1))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        lg = step_device()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    n_cls = Lb.wat_profile_classes()
    pms, pcnt = (C.c_double * n_cls)(), (C.c_int64 * n_cls)()
    _lib.check(Lb.wat_profile_read(eng.h, pms, pcnt))
    _lib.check(Lb.wat_profile(eng.h, 0))
    launches = model.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = 30.0 * B * world * args.steps / (ms / 1000.0)

    # ---- e2e: host buffers through the C-ABI host entry (H2D of the PCM + D2H of the logits inside the timed region)
    out_host = torch.empty((B, S, 527), dtype=torch.float32).pin_memory()
    model.tag_batch_host(audio_host, at_time_res=args.res, out=out_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.tag_batch_host(audio_host, at_time_res=args.res, out=out_host)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = 30.0 * B * world * args.steps / float(te.item())
    same = bool(torch.equal(out_host, lg.cpu()))

    if rank == 0:
        pk = peaks()
        prof = {Lb.wat_profile_class_name(i).decode(): dict(ms_per_step=pms[i] / args.steps, launches_per_step=pcnt[i] / args.steps)
                for i in range(n_cls)}
        gemm_ids = [i for i in range(n_cls) if Lb.wat_profile_class_name(i).decode().startswith('gemm')]
        gemm_ms = sum(pms[i] for i in gemm_ids) / args.steps
        achieved = fl["gemm"] * B / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
        attn_ms = prof["attention"]["ms_per_step"]
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=n_warm,
                   ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype="bf16" if args.precision == "bf16" else "f32", data="synthetic", impl="ours", config=config,
                   e2e=dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(audio_host.numel() * 4) * world,
                            d2h_bytes_per_step=int(out_host.numel() * 4) * world,
                            h2d_bytes_per_step_per_gpu=int(audio_host.numel() * 4), matches_device_path=same),
                   gpu_launches=int(launches), clocks=clocks,
                   roofline=dict(bound="tensor", kernel="gemm_tc2_kernel<8|16> / gemm_tc_kernel (all dense tcgen05 GEMMs of the step)", achieved=achieved,
                                 peak=pk["tflops"], unit="TFLOP/s", frac=achieved / pk["tflops"], traffic=None,
                                 peak_source=pk["source"], kernel_ms_per_step=gemm_ms,
                                 kernel_share_of_step=gemm_ms / (ms / args.steps),
                                 attention_tflops=(fl["attn"] * B / (attn_ms / 1000.0) / 1e12) if attn_ms > 0 else None,
                                 whole_step_tflops=fl["total"] * B / (ms / args.steps / 1000.0) / 1e12),
                   kernel_profile=prof)
        T = 1500
        per_kind = {"gemm_qkv": 2 * T * 3 * d * d, "gemm_out": 2 * T * d * d, "gemm_fc1": 2 * T * 4 * d * d, "gemm_fc2": 2 * T * 4 * d * d}
        # DRAM traffic of one launch of each encoder GEMM kind from the committed ncu capture (16 clips per launch, MB:
        # dram__bytes_read.sum + dram__bytes_write.sum) next to the algorithmic bytes of that launch; `traffic` itself stays
        # null because `achieved` aggregates every GEMM launch of the step rather than one launch
        if args.model == "large-v2" and not args.low:
            out["roofline"]["traffic_ncu_MB_per_launch_16clips"] = dict(
                source="profiles/r01e_v5_all_kernels.txt", gemm_qkv=203.6, gemm_out=274.0, gemm_fc1=267.8, gemm_fc2=644.1,
                algorithmic=dict(gemm_qkv=255.6, gemm_out=310.5, gemm_fc1=320.3, gemm_fc2=504.8))
        out["roofline"]["encoder_gemm_tflops"] = {k: (f * L * B / (prof[k]["ms_per_step"] / 1000.0) / 1e12) if prof[k]["ms_per_step"] > 0 else None
                                                  for k, f in per_kind.items()}
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            tt = oracle_time_clips(args.model, n_mels, args.low, args.res, 1, threads)          # first clip also warms torch up
            n_more = max(1, min(8, int(15.0 / max(tt[0], 1e-3))))                                # ~15 s of CPU work
            tt = oracle_time_clips(args.model, n_mels, args.low, args.res, n_more, threads)
            out["cpu_baseline"] = dict(value=30.0 * len(tt) / sum(tt), unit=UNIT, cores=threads, kind="port",
                                       sample=f"{len(tt)} clips of the same workload (after 1 warm-up clip), one per call as the "
                                              "reference's AT path requires, oracle/wat_oracle.py (torch CPU fp32), all host threads")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
