#!/usr/bin/env python3
"""Benchmark of the vocoder hot path (BASELINE.json metric: vocoder audio-sec/sec, xRT).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the CPU restatement of the reference

Workload (config.workload): BASELINE.json configs[2], "batch of 256 independent 64-frame x
16-codebook chunks on 1 B200" -- the configuration the xRT throughput target is quoted on
(configs[1], one chunk at batch 1, is a latency case and is reported as ``latency_ms``).
One *step* = one pass of the hot path (RVQ gather -> ... -> head) over one batch of synthetic
random codes with random-init weights of the named architecture (seeds: weights 0, codes 1).
Multi-GPU: one process per GPU, every rank its own batch (weak scaling, no data-path
collective); the value is all ranks' audio seconds / max-over-ranks device time.

Prints ONE JSON line (see the keys below).  audio seconds are nominal: 64 frames * 1920 / 24000
= 5.12 s per chunk, whatever transconv_trim is.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CHUNK_AUDIO_S = 64 * 1920 / 24000.0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle (the only executable restatement of the reference's model call here)
# ----------------------------------------------------------------------------------------

def time_oracle(cfg, weights, chunks_per_step: int, steps: int, warmup: int, threads: int):
    import torch
    from oracle import vocoder_oracle as VO
    torch.set_num_threads(threads)
    W = VO.Weights(weights)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (chunks_per_step, cfg.chunk_frames, 16),
                                              dtype=np.int64)
    for _ in range(warmup):
        VO.forward(codes, W, cfg)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        VO.forward(codes, W, cfg)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank: int):
    """--impl reference: the reference's own implementation of the path is ONNX Runtime CPU on a
    model file that does not exist in this image; what runs is its CPU
    restatement (oracle/, torch CPU FP32) with all host threads, on a bounded sample."""
    if rank != 0:
        return
    pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
    cfg = pkg.VocoderConfig()
    weights = pkg.init_weights(cfg, 0)
    threads = os.cpu_count() or 1
    sample = 2
    ts = time_oracle(cfg, weights, sample, args.steps, max(1, min(args.warmup, 2)), threads)
    t = sum(ts) / len(ts)
    xrt = sample * CHUNK_AUDIO_S / t
    line = {
        "impl": "reference", "metric": "vocoder_xrt", "value": xrt, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "256 independent 64-frame x 16-codebook chunks per GPU (BASELINE configs[2])",
                   "architecture": "qwen3-tts-12hz decoder, decoder_dim 1536, 114 M params, random init seed 0",
                   "transconv_trim": cfg.transconv_trim},
        "cpu_baseline": {"value": xrt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the 256 chunks per step (torch CPU FP32 oracle; "
                                   "onnxruntime and the .onnx file are absent from this image)"},
        "e2e": {"value": xrt, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops": cfg.flops_per_chunk() * sample / t / 1e9,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------

def run_b200(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
    backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    cfg = pkg.VocoderConfig()
    weights = pkg.init_weights(cfg, 0)
    voc = backend.Vocoder(cfg, weights, device=local_rank, wave=args.wave)
    voc.set_option("gemm", args.gemm)
    B, Lc = args.batch, voc.chunk_samples
    rng = np.random.default_rng(1 + rank)
    h_codes = torch.from_numpy(rng.integers(0, cfg.codebook_size, (B, cfg.chunk_frames, 16), dtype=np.int64)).pin_memory()
    d_codes = h_codes.to(dev)
    d_out = torch.empty(B, Lc, dtype=torch.float32, device=dev)
    h_out = torch.empty(B, Lc, dtype=torch.float32).pin_memory()
    st = torch.cuda.Stream(device=dev)
    own = torch.cuda.ExternalStream(voc.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (1) device-resident throughput: the timed region of `value`
    for _ in range(args.warmup):
        voc.infer_chunks_dev(d_codes, B, d_out, st.cuda_stream)
    voc.check_dev(st.cuda_stream)
    voc.set_option("profile", "1")
    voc.profile_report()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = voc.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        voc.infer_chunks_dev(d_codes, B, d_out, st.cuda_stream)
    e1.record(st)
    barrier()
    clocks = sampler.finish()
    launches = voc.kernel_launches - l0
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    prof = voc.profile_report()
    voc.set_option("profile", "0")
    value = world * B * CHUNK_AUDIO_S / (ms_dev / 1e3)

    # ---- (2) end to end through the host C-ABI call: pinned host codes in, host floats out
    for _ in range(max(1, args.warmup // 2)):
        voc.lib.voc_infer_chunks(voc._h, h_codes.data_ptr(), B, h_out.data_ptr())
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(own)
    for _ in range(args.steps):
        rc = voc.lib.voc_infer_chunks(voc._h, h_codes.data_ptr(), B, h_out.data_ptr())
        if rc:
            raise RuntimeError(voc.lib.voc_last_error(voc._h))
    f1.record(own)
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1) / args.steps)
    e2e = world * B * CHUNK_AUDIO_S / (ms_e2e / 1e3)
    checksum = float(h_out[0, :1000].double().abs().sum())

    # ---- (3) batch-1 streaming latency (BASELINE configs[1]), host to host, p50 over 30
    lat = []
    one = h_codes[:1].contiguous().pin_memory()
    for i in range(35):
        t0 = time.perf_counter()
        voc.lib.voc_infer_chunks(voc._h, one.data_ptr(), 1, h_out.data_ptr())
        if i >= 5:
            lat.append((time.perf_counter() - t0) * 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (the tap-GEMM that runs every conv/linear)
    peaks, peak_src = _peaks()
    gemm = [p for p in prof if p["tag"] not in ("rvq_gather", "head", "stitch", "xf.norm", "xf.attn",
                                                "xf.swiglu", "up.dwconv_ln")]
    g_ms = sum(p["ms"] for p in gemm)
    g_flops = sum(p["flops"] for p in gemm)
    g_calls = sum(p["calls"] for p in gemm)
    all_ms = sum(p["ms"] for p in prof)
    family = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    passes = 1 if args.gemm == "simt" else 3
    # the dominant kernel = the layer (one launch per wave of windows) with the largest share of the step
    top = max(gemm, key=lambda q: q["ms"]) if gemm else None
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if top and args.wave == tr.get("wave") and B % args.wave == 0 and top["tag"] in tr and args.gemm != "simt":
            traffic = tr[top["tag"]]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    achieved = top["flops"] / (top["ms"] / 1e3) / 1e12 if top and top["ms"] > 0 else 0.0
    roofline = {
        "kernel": (f"tapgemm_tc_kernel (tcgen05 tap-GEMM), layer {top['tag']}" if args.gemm != "simt"
                   else f"tapgemm_simt_kernel, layer {top['tag']}") if top else None,
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_note": "DRAM bytes of one launch of that layer (ncu --set full, profiles/r1_traffic.json); "
                        "equals its algorithmic bytes (operand in, operand out)" if traffic else None,
        "peak_source": f"{peak_src} bf16 dense sustained (MEASURED_PEAKS.json)",
        "launches_per_step": top["calls"] / args.steps if top else None,
        "avg_launch_ms": top["ms"] / max(top["calls"], 1) if top else None,
        "share_of_step": top["ms"] / all_ms if top and all_ms else None,
        "tensor_passes_per_flop": passes,
        "mma_issue_tflops": achieved * passes,
        "family": {"what": "all tap-GEMM launches (every conv / transposed-conv / linear layer)",
                   "achieved": family, "frac": family / peak, "mma_issue_tflops": family * passes,
                   "launches_per_step": g_calls / args.steps, "share_of_step": g_ms / all_ms if all_ms else None},
        "note": ("FP32 CUDA-core path" if args.gemm == "simt" else
                 "`achieved` counts algorithmic FLOPs (2*M*N*K*taps); every product is three fp16 tcgen05 passes "
                 "(hi*hi + hi*lo + lo*hi, FP32 accumulate) because one bf16/tf32 pass fails the 60 dB / 1e-4 gate, "
                 "so the tensor pipe runs at 3x `achieved` (mma_issue_tflops)"),
    }
    breakdown = {p["tag"]: {"ms_per_step": p["ms"] / args.steps,
                            "tflops": (p["flops"] / (p["ms"] / 1e3) / 1e12) if p["ms"] > 0 and p["flops"] else None,
                            "gbs": (p["bytes"] / (p["ms"] / 1e3) / 1e9) if p["ms"] > 0 else None}
                 for p in sorted(prof, key=lambda q: -q["ms"])}

    # ---- CPU baseline on this box's host cores (bounded sample: 1 warm-up + 3 chunks)
    threads = os.cpu_count() or 1
    ts = time_oracle(cfg, weights, 1, 3, 1, threads)
    cpu_xrt = CHUNK_AUDIO_S / (sum(ts) / len(ts))

    line = {
        "metric": "vocoder_xrt", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.gemm == "simt" else "f16x3-f32acc",
        "data": "synthetic",
        "config": {"workload": f"{B} independent 64-frame x 16-codebook chunks per GPU (BASELINE configs[2])",
                   "architecture": "qwen3-tts-12hz decoder, decoder_dim 1536, 114 M params, random init seed 0",
                   "transconv_trim": cfg.transconv_trim, "wave": args.wave, "gemm": args.gemm,
                   "l2": "activations streamed per step exceed L2 by >100x; no flush needed",
                   "audio_seconds_per_chunk": CHUNK_AUDIO_S},
        "tflops_algorithmic": world * B * cfg.flops_per_chunk() / (ms_dev / 1e3) / 1e12,
        "e2e": {"value": e2e, "unit": "audio-s/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h_codes.numel() * 8), "d2h_bytes_per_step": int(B * Lc * 4),
                "call": "voc_infer_chunks (host int64 codes -> host float32 audio)", "checksum": checksum},
        "latency_ms": {"workload": "1 chunk, batch 1, host to host (BASELINE configs[1])",
                       "p50": statistics.median(lat), "p95": sorted(lat)[int(0.95 * len(lat)) - 1]},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_xrt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": "3 chunks after 1 warm-up, torch CPU FP32 oracle, all host threads "
                                   "(onnxruntime / the .onnx file are absent from this image)"},
        "breakdown": breakdown,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="chunks per GPU per step")
    ap.add_argument("--wave", type=int, default=32, help="chunks resident in HBM at once")
    ap.add_argument("--gemm", default="auto", choices=["auto", "simt", "tc"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
