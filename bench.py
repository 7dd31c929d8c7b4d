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
import ctypes
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


def bench_config(args, cfg):
    """The `config` object of the JSON line -- the same for both arms (the workload, not the engine)."""
    return {"workload": f"{args.batch} independent 64-frame x 16-codebook chunks per GPU (BASELINE configs[2])",
            "architecture": "qwen3-tts-12hz decoder, decoder_dim 1536, 114 M params, random init seed 0",
            "transconv_trim": cfg.transconv_trim, "wave": args.wave, "gemm": args.gemm,
            "l2": "activations streamed per step exceed L2 by >100x; no flush needed",
            "audio_seconds_per_chunk": CHUNK_AUDIO_S}


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


def code_predictor_leg(device: int, frames: int = 60):
    """the code predictor's frame latency (the reference's "86 % of per-token time") through cp_predict
    -- host hidden state + code_0 embedding in, 15 host int32 codes out, one CUDA-graph launch per frame -- and
    through the level-1 step interface with the reference's host sampler; the CPU oracle's frame time beside it."""
    import importlib
    import torch
    cpm = importlib.import_module("qwen3-tts-axera-russian_b200.code_predictor")
    cfg = cpm.CPConfig()
    w = cpm.init_weights(cfg, 0)
    cp = cpm.CodePredictor(cfg, w, device=device)
    rng = np.random.default_rng(11)
    hs = rng.standard_normal((frames + 16, cfg.hidden)).astype(np.float32)
    es = rng.standard_normal((frames + 16, cfg.hidden)).astype(np.float32)

    def frame_bytes_fn():
        layer_bytes = sum(int(np.prod(sh)) for k, sh in cpm.weight_shapes(cfg).items() if k.startswith("layer_")) * 4
        return (cfg.groups + 1) * layer_bytes + cfg.groups * cfg.vocab * cfg.hidden * 4
    for i in range(5):
        cp.predict(hs[i], es[i], 0.1, 50, seed=i)
    l0 = cp.launches
    lat = []
    for i in range(5, frames + 5):
        t = time.perf_counter()
        cp.predict(hs[i], es[i], 0.1, 50, seed=i)
        lat.append((time.perf_counter() - t) * 1e3)
    per_frame_launches = (cp.launches - l0) // frames
    lat.sort()
    path = cp.predict_path
    # the opt-in persistent whole-frame kernel, same frames
    persistent = None
    try:
        cp.set_option("predict", "persistent")
        for i in range(3):
            cp.predict(hs[i], es[i], 0.1, 50, seed=i)
        lp = []
        for i in range(5, frames + 5):
            t = time.perf_counter()
            cp.predict(hs[i], es[i], 0.1, 50, seed=i)
            lp.append((time.perf_counter() - t) * 1e3)
        lp.sort()
        persistent = {"frame_ms_p50": lp[len(lp) // 2], "gpu_launches_per_frame": 1,
                      "what": "cp_set_option(predict, persistent): one cooperative kernel, 432 grid barriers"}
    except Exception as e:
        persistent = {"error": repr(e)}
    cp.set_option("predict", path)
    # B independent streams per launch (cp_predict_batch): the weights are streamed once for all of them
    batched = {}
    try:
        for B in (2, 4, 8):
            sd = np.arange(B, dtype=np.uint64)
            for i in range(3):
                cp.predict_batch(hs[i:i + B], es[i:i + B], 0.1, 50, seeds=sd)
            lb = []
            for i in range(5, 5 + max(10, frames // 3)):
                t = time.perf_counter()
                cp.predict_batch(hs[i:i + B], es[i:i + B], 0.1, 50, seeds=sd)
                lb.append((time.perf_counter() - t) * 1e3)
            lb.sort()
            ms = lb[len(lb) // 2]
            batched[str(B)] = {"ms_per_launch_p50": ms, "frames_per_s": B * 1e3 / ms,
                               "weight_stream_gbs": frame_bytes_fn() / (ms / 1e3) / 1e9}
    except Exception as e:
        batched["error"] = repr(e)
    # level 1: 16 steps + 15 logits round trips per frame, greedy on the host
    t = time.perf_counter()
    n1 = 10
    for i in range(n1):
        cp.reset()
        cp.step(hs[i][None], 0)
        cp.step(es[i][None], 1)
        for g in range(cfg.groups):
            tok = int(np.argmax(cp.logits(g)))
            if g + 1 < cfg.groups:
                cp.step(w[f"codec_emb_{g}"][tok][None], g + 2)
    ms_l1 = (time.perf_counter() - t) * 1e3 / n1
    frame_bytes = frame_bytes_fn()
    peaks, peak_src = _peaks()
    hbm = float(peaks.get("hbm_gbs_sustained", peaks.get("hbm_gbs", 6400.0)))
    p50 = lat[len(lat) // 2]
    out = {"workload": "code predictor, production shape (5 layers, 1024 hidden, GQA 16/8 x 128, 3072 MLP, 15 groups x 2048), "
                       "random weights, batch 1: one frame = 16 decode steps + 15 lm_heads + 15 top-k draws",
           "call": "cp_predict (host float32 hidden state + code_0 embedding -> 15 host int32 codes), T = 0.1, top_k = 50",
           "frame_ms_p50": p50, "frame_ms_p95": lat[int(0.95 * len(lat)) - 1], "frames_per_s": 1e3 / p50,
           "realtime_factor_at_12.5_frames_per_s": (1e3 / p50) / 12.5,
           "predict_path": path, "gpu_launches_per_frame": int(per_frame_launches), "persistent_kernel": persistent,
           "batched_streams": batched,
           "level1_frame_ms": ms_l1,
           "dtype": "f32",
           "roofline": {"bound": "hbm", "unit": "GB/s", "achieved": frame_bytes / (p50 / 1e3) / 1e9, "peak": hbm,
                        "frac": frame_bytes / (p50 / 1e3) / 1e9 / hbm, "traffic": 5163419896,
                        "traffic_note": "DRAM read + write bytes of one frame, ncu --set full of cp_frame_kernel (one launch = "
                                        "one frame; profiles/r2_ncu_cp_frame_kernel.txt): equals the algorithmic bytes",
                        "algorithmic_bytes_per_frame": int(frame_bytes),
                        "note": "float32 weights streamed once per decode step (314.6 MB of layer weights x 16 steps + 15 "
                                "lm_heads); the whole frame incl. H2D / D2H and the graph launch is in the time",
                        "peak_source": peak_src}}
    try:
        from oracle import code_predictor_oracle as CPO
        ocfg = CPO.CPConfig()
        W = CPO.Weights(w)
        torch.set_num_threads(os.cpu_count() or 1)
        CPO.predict(hs[0], es[0], W, ocfg)
        t = time.perf_counter()
        want = CPO.predict(hs[1], es[1], W, ocfg)
        cpu_ms = (time.perf_counter() - t) * 1e3
        got = [int(c) for c in cp.predict(hs[1], es[1], 0.1, 1, seed=0)]
        out["cpu_baseline"] = {"frame_ms": cpu_ms, "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "1 frame after 1 warm-up, torch CPU FP32 oracle, greedy"}
        out["parity"] = {"greedy_codes_equal_oracle": got == want, "frame": 1}
    except Exception as e:
        out["cpu_baseline"] = {"error": repr(e)}
    cp.close()
    return out


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
    # BASELINE configs[0] as written: intra_op_num_threads = 4 (dual_npu/vocoder_server.py:41), one chunk
    t4 = time_oracle(cfg, weights, 1, 2, 1, 4)
    xrt4 = CHUNK_AUDIO_S / (sum(t4) / len(t4))
    line = {
        "impl": "reference", "metric": "vocoder_xrt", "value": xrt, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, cfg),
        "cpu_baseline": {"value": xrt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the {args.batch} chunks per step, per-chunk rate extrapolated (torch CPU "
                                   "FP32 oracle, all host threads; onnxruntime and the .onnx file are absent from this image)",
                         "threads_4": {"value": xrt4, "cores": 4,
                                       "sample": "1 chunk x 2 after 1 warm-up at torch.set_num_threads(4) = the "
                                                 "reference's intra_op_num_threads (BASELINE configs[0])"}},
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
    if args.profile_only:                      # runs under ncu: the device-timed region above is all that is wanted
        if rank == 0:
            print(json.dumps({"profile_only": True, "value": value, "ms_per_step": ms_dev, "gpu_launches": int(launches)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

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

    pick = sorted({0, min(args.wave - 1, B - 1), min(args.wave, B - 1), B - 1})
    got_pick = h_out[pick].numpy().copy() if rank == 0 else None      # for the parity check (4)
    simt_launches = int(voc.simt_launches)

    # ---- (3) batch-1 streaming latency (BASELINE configs[1]), host to host, p50 / p95 over 200
    lat = []
    one = h_codes[:1].contiguous().pin_memory()
    for i in range(25 if args.no_legs else 220):
        t0 = time.perf_counter()
        voc.lib.voc_infer_chunks(voc._h, one.data_ptr(), 1, h_out.data_ptr())
        if i >= 20:
            lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    # the serving curve between configs[1] and configs[2]: p50 of the same host-to-host call for small batches (what a
    # coalescing server sees); the launcher picks narrower column tiles while the widest would leave SMs idle
    sweep = {}
    if not args.no_legs:
        for b in (2, 4, 8, 16, 32):
            if b > B:
                break
            tb = []
            for i in range(35):
                t0 = time.perf_counter()
                voc.lib.voc_infer_chunks(voc._h, h_codes.data_ptr(), b, h_out.data_ptr())
                if i >= 5:
                    tb.append((time.perf_counter() - t0) * 1e3)
            tb.sort()
            sweep[str(b)] = {"p50_ms": tb[len(tb) // 2], "audio_s_per_s": b * CHUNK_AUDIO_S / (tb[len(tb) // 2] / 1e3)}

    # ---- (4) parity of THIS configuration: four windows of the batch just timed (first, last, either side of
    # a wave boundary) against the CPU oracle, computed after the timed regions; rank 0 only

    # ---- (5) BASELINE configs[3]: one 10-minute utterance (7 500 frames, 157 windows) split into contiguous window
    # ranges over the ranks, each rank stitching the span its windows own, int16 PCM gathered to rank 0 (NCCL),
    # D2H on rank 0.  Timed with CUDA events on one stream that carries the H2D, the kernels, the gather and the D2H.
    S = importlib.import_module("qwen3-tts-axera-russian_b200.sharding")
    utt = None
    if not args.no_legs:
        n_u = 7500
        u_codes = torch.from_numpy(np.random.default_rng(7).integers(0, cfg.codebook_size, (n_u, 16), dtype=np.int64)).pin_memory()
        nw_u = voc.num_windows(n_u)
        total_u = voc.out_samples(n_u)
        ranges = S.window_ranges(nw_u, world)
        w0, w1 = ranges[rank]
        d_u = torch.empty(n_u, 16, dtype=torch.int64, device=dev)
        d_pcm = torch.empty(total_u, dtype=torch.int16, device=dev)
        h_pcm = torch.empty(total_u, dtype=torch.int16).pin_memory()
        counts = None

        def utt_step():
            nonlocal counts
            with torch.cuda.stream(st):
                d_u.copy_(u_codes, non_blocking=True)
                off, cnt = voc.synthesize_range_dev(d_u, n_u, w0, w1, d_out_i16=d_pcm, cap=total_u, stream=st.cuda_stream)
                if world > 1:
                    if counts is None:
                        c = torch.tensor([cnt], dtype=torch.int64, device=dev)
                        allc = [torch.zeros_like(c) for _ in range(world)]
                        dist.all_gather(allc, c)
                        counts = [int(x.item()) for x in allc]
                    mx = max(counts)
                    send = d_pcm[:mx].view(torch.uint8)
                    glist = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
                    dist.gather(send, glist, dst=0)
                    if rank == 0:
                        o = 0
                        for r in range(world):
                            h_pcm[o:o + counts[r]].copy_(glist[r].view(torch.int16)[:counts[r]], non_blocking=True)
                            o += counts[r]
                else:
                    h_pcm[:cnt].copy_(d_pcm[:cnt], non_blocking=True)
            return cnt

        for _ in range(2):
            utt_step()
        voc.check_dev(st.cuda_stream)
        barrier()
        reps_u = 3
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record(st)
        for _ in range(reps_u):
            utt_step()
        u1.record(st)
        barrier()
        ms_u = max_over_ranks(u0.elapsed_time(u1) / reps_u)
        bit_equal = None
        if rank == 0:
            alone = voc.synthesize_pcm16(u_codes.numpy())            # the same request on this GPU alone
            bit_equal = bool(len(alone) == total_u and np.array_equal(alone, h_pcm.numpy()))
        utt = {"workload": "one 10-minute utterance: 7500 frames, 157 windows of 64 frames, stride 48, crossfade 16 "
                           "(BASELINE configs[3]); contiguous window ranges per rank, NCCL gather of int16 PCM to rank 0",
               "value": n_u * 0.08 / (ms_u / 1e3), "unit": "audio-s/s", "ms": ms_u, "n_gpus": world,
               "windows_per_rank": [b - a for a, b in ranges], "out_samples": int(total_u),
               "pcm_gather_bytes": int(2 * total_u) if world > 1 else 0, "bit_equal_to_one_gpu": bit_equal,
               "timed": "H2D codes + kernels + stitch + PCM gather + D2H on rank 0, CUDA events, max over ranks"}

    # ---- (6) BASELINE configs[4]: 1000 mixed-length utterances (lengths round(exp(N(ln 100, 0.8^2))) clipped to
    # [8, 3750] frames, seed 2), sharded over the ranks by longest-processing-time balance; every rank synthesises
    # its whole shard with ONE voc_synthesize_batch_pcm16 call (host codes in, host PCM out); no collective.
    corpus = None
    if not args.no_legs:
        rng2 = np.random.default_rng(2)
        lengths = np.clip(np.round(np.exp(rng2.normal(np.log(100), 0.8, args.corpus))), 8, 3750).astype(int)
        bins = S.shard_corpus(lengths, world, cfg.chunk_frames)
        mine = bins[rank]
        my_lens = np.asarray([int(lengths[i]) for i in mine], dtype=np.int32)
        frames_mine = int(my_lens.sum())
        c_codes = torch.from_numpy(np.random.default_rng(1000 + rank).integers(
            0, cfg.codebook_size, (max(frames_mine, 1), 16), dtype=np.int64)).pin_memory()
        cap_c = int(sum(voc.out_samples(int(n)) for n in my_lens))
        c_out = torch.empty(max(cap_c, 1), dtype=torch.int16).pin_memory()
        offs = np.zeros(len(mine) + 1, dtype=np.int64)

        def corpus_step():
            if not len(mine):
                return
            rc = voc.lib.voc_synthesize_batch_pcm16(voc._h, c_codes.data_ptr(), my_lens.ctypes.data, len(mine),
                                                    c_out.data_ptr(), cap_c, offs.ctypes.data)
            if rc:
                raise RuntimeError(voc.lib.voc_last_error(voc._h))

        corpus_step()
        barrier()
        reps_c = 2
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(own)
        for _ in range(reps_c):
            corpus_step()
        c1.record(own)
        barrier()
        ms_c = max_over_ranks(c0.elapsed_time(c1) / reps_c)
        # spot check on rank 0: the first utterance of the shard alone gives the same PCM
        same = None
        if rank == 0 and len(mine):
            a = voc.synthesize_pcm16(c_codes[: int(my_lens[0])].numpy())
            same = bool(np.array_equal(a, c_out[: int(offs[1])].numpy()))
        nwin_c = int(sum(voc.num_windows(int(n)) for n in lengths))
        corpus = {"workload": f"{args.corpus} utterances, lengths round(exp(N(ln 100, 0.8^2))) in [8, 3750] frames, seed 2 "
                              "(BASELINE configs[4]); LPT shard per rank, one batched call per rank, no collective",
                  "value": float(lengths.sum()) * 0.08 / (ms_c / 1e3), "unit": "audio-s/s", "ms": ms_c, "n_gpus": world,
                  "frames": int(lengths.sum()), "windows": nwin_c,
                  "window_frames_over_utterance_frames": nwin_c * 64 / float(lengths.sum()),
                  "first_utterance_equals_single_request": same,
                  "timed": "host int64 codes -> host int16 PCM through voc_synthesize_batch_pcm16, CUDA events on the "
                           "handle's stream, max over ranks"}

    # ---- (7) the same 10-minute utterance through the opt-in carried-state decode: no windows, no
    # recomputed overlap, no crossfade -- the output is the un-chunked decoder's, NOT the reference's stitched one, so
    # this is reported beside the windowed figure, never instead of it.  Needs the causal trim; rank 0 only.
    stream = None
    if not args.no_legs and rank == 0:
        import dataclasses
        cfg_r = dataclasses.replace(cfg, transconv_trim="right")
        voc.close()                                   # one set of activation pools at a time
        voc_r = backend.Vocoder(cfg_r, weights, device=local_rank, wave=args.wave)
        voc_r.set_option("gemm", args.gemm)
        own_r = torch.cuda.ExternalStream(voc_r.stream, device=dev)
        n_u = 7500
        s_codes = torch.from_numpy(np.random.default_rng(7).integers(0, cfg.codebook_size, (n_u, 16), dtype=np.int64)).pin_memory()
        s_out = torch.empty(n_u * 1920, dtype=torch.int16).pin_memory()
        cnt = ctypes.c_longlong(0)

        def stream_step():
            voc_r.stream_reset()
            rc = voc_r.lib.voc_stream_decode_pcm16(voc_r._h, s_codes.data_ptr(), n_u, s_out.data_ptr(), s_out.numel(), ctypes.byref(cnt))
            if rc:
                raise RuntimeError(voc_r.lib.voc_last_error(voc_r._h))

        stream_step()
        torch.cuda.synchronize(dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(own_r)
        for _ in range(3):
            stream_step()
        s1.record(own_r)
        torch.cuda.synchronize(dev)
        ms_s = s0.elapsed_time(s1) / 3
        # windowed decode of the same codes on the same (right-trim) handle, for the ratio
        w_out = torch.empty(voc_r.out_samples(n_u), dtype=torch.int16).pin_memory()
        wn = ctypes.c_longlong(0)
        voc_r.lib.voc_synthesize_pcm16(voc_r._h, s_codes.data_ptr(), n_u, w_out.data_ptr(), w_out.numel(), ctypes.byref(wn))
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(own_r)
        for _ in range(3):
            voc_r.lib.voc_synthesize_pcm16(voc_r._h, s_codes.data_ptr(), n_u, w_out.data_ptr(), w_out.numel(), ctypes.byref(wn))
        w1.record(own_r)
        torch.cuda.synchronize(dev)
        ms_w = w0.elapsed_time(w1) / 3
        stream = {"workload": "the 10-minute utterance (7500 frames) through voc_stream_decode_pcm16: carried per-layer state, "
                              "segments of wave*64-1 frames, transconv_trim=right; host codes -> host PCM",
                  "value": n_u * 0.08 / (ms_s / 1e3), "unit": "audio-s/s", "ms": ms_s, "out_samples": int(cnt.value),
                  "windowed_same_handle": {"value": n_u * 0.08 / (ms_w / 1e3), "ms": ms_w, "out_samples": int(wn.value)},
                  "speedup_over_windowed": ms_w / ms_s,
                  "note": "opt-in mode: output = un-chunked decode (tests/test_gpu_stream.py), not the reference's 64/48/16 stitching"}
        # live streaming: the talker emits 12.5 frames/s; decode k new frames per call with the carried state (the
        # reference's client instead waits for 64-frame windows, tts_client.py:188-197).  p50 host-to-host per call.
        inc = {}
        try:
            for k in (0, 1, 4, 16):
                # k = 0: an untimed pass of one-frame calls -- straight after the 157-window legs the SM clock is still
                # recovering from the power cap, which would be charged to whichever piece length came first
                timed = k > 0
                k = max(k, 1)
                voc_r.stream_reset()
                tk = []
                for i in range(40 if timed else 80):
                    t0 = time.perf_counter()
                    rc = voc_r.lib.voc_stream_decode_pcm16(voc_r._h, s_codes[i * k:(i + 1) * k].data_ptr(), k, s_out.data_ptr(),
                                                          s_out.numel(), ctypes.byref(cnt))
                    if rc:
                        raise RuntimeError(voc_r.lib.voc_last_error(voc_r._h))
                    if i >= 8:
                        tk.append((time.perf_counter() - t0) * 1e3)
                if not timed:
                    continue
                tk.sort()
                inc[str(k)] = {"p50_ms": tk[len(tk) // 2], "audio_ms_per_call": k * 80.0,
                               "realtime_factor": k * 80.0 / tk[len(tk) // 2]}
        except Exception as e:                                # a side measurement must not cost the headline line
            inc["error"] = repr(e)
        stream["incremental"] = inc
        stream["incremental_note"] = ("k new frames per voc_stream_decode_pcm16 call after a reset, 32 timed calls each: "
                                      "latency of the first audio of a live stream = the k = 1 figure")
        voc_r.close()

    cp_leg = None
    if not args.no_legs and rank == 0:
        try:
            cp_leg = code_predictor_leg(local_rank)
        except Exception as e:                                # an N4 failure must not cost the headline line
            cp_leg = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    parity = None
    try:
        from oracle import vocoder_oracle as VO
        torch.set_num_threads(os.cpu_count() or 1)
        ref, _ = VO.forward(h_codes[pick].numpy(), VO.Weights(weights), cfg)
        ref = ref.numpy().astype(np.float64)
        err = ref - got_pick
        snrs = [10 * np.log10((ref[i] ** 2).sum() / max((err[i] ** 2).sum(), 1e-300)) for i in range(len(pick))]
        parity = {"windows": pick, "snr_db": float(min(snrs)), "max_abs": float(np.abs(err).max()),
                  "gate": "snr_db >= 60 and max_abs <= 1e-4 vs the FP32 CPU oracle on the same codes",
                  "pass": bool(min(snrs) >= 60.0 and np.abs(err).max() <= 1e-4),
                  "what": "windows of the timed 256-window batch (wave 32, front wave 256), host-to-host call"}
    except Exception as e:                                    # the bench line must still print
        parity = {"error": repr(e)}

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
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if top and args.wave == tr.get("wave") and B % args.wave == 0 and top["tag"] in tr and args.gemm != "simt":
            traffic = tr[top["tag"]]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    achieved = top["flops"] / (top["ms"] / 1e3) / 1e12 if top and top["ms"] > 0 else 0.0
    roofline = {
        "kernel": (f"tapgemm_simt_kernel, layer {top['tag']}" if args.gemm == "simt"
                   else f"ru_fused_kernel (tcgen05: conv7 -> Snake -> conv1 -> + residual in one launch), layer {top['tag']}"
                   if top["tag"].endswith(".fused") else f"tapgemm_tc_kernel (tcgen05 tap-GEMM), layer {top['tag']}") if top else None,
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_note": "DRAM bytes of one launch of that layer (ncu --set full, profiles/r2_traffic.json)" if traffic else None,
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
    t4 = time_oracle(cfg, weights, 1, 2, 1, 4)                 # the reference's intra_op_num_threads = 4
    cpu_xrt4 = CHUNK_AUDIO_S / (sum(t4) / len(t4))

    line = {
        "metric": "vocoder_xrt", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.gemm == "simt" else "f16x3-f32acc",
        "data": "synthetic",
        "config": bench_config(args, cfg),
        "tflops_algorithmic_nominal": world * B * cfg.flops_per_chunk() / (ms_dev / 1e3) / 1e12,
        "tflops_note": "nominal 317.49 GFLOP per window (SURVEY 8d); the executed graph (transconv_trim) does "
                       f"{cfg.flops_per_chunk(nominal=False) / 1e9:.2f} GFLOP",
        "e2e": {"value": e2e, "unit": "audio-s/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h_codes.numel() * 8), "d2h_bytes_per_step": int(B * Lc * 4),
                "call": "voc_infer_chunks (host int64 codes -> host float32 audio)", "checksum": checksum,
                "parity": parity,
                "latency_ms": {"workload": "1 chunk, batch 1, host to host incl. H2D of 8 KB codes and D2H of the "
                                           "window (BASELINE configs[1]); 200 calls after 20 warm-ups",
                               "p50": lat[len(lat) // 2], "p95": lat[int(0.95 * len(lat)) - 1],
                               "small_batches": sweep,
                               "small_batches_note": "p50 of 30 calls of the same host-to-host call with 2..32 windows"},
                "utterance_10min": utt, "corpus_1k": corpus, "stream_10min": stream,
                "code_predictor": cp_leg},
        "gpu_launches": int(launches),
        "simt_launches": simt_launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_xrt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": "3 chunks after 1 warm-up, torch CPU FP32 oracle, all host threads "
                                   "(onnxruntime / the .onnx file are absent from this image)",
                         "threads_4": {"value": cpu_xrt4, "cores": 4,
                                       "sample": "2 chunks after 1 warm-up at torch.set_num_threads(4), the reference's "
                                                 "intra_op_num_threads (dual_npu/vocoder_server.py:41; BASELINE configs[0])"}},
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
    ap.add_argument("--gemm", default="tc", choices=["auto", "simt", "tc"])
    ap.add_argument("--corpus", type=int, default=1000, help="utterances of the corpus leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the 10-minute-utterance and corpus legs")
    ap.add_argument("--profile-only", action="store_true", help="stop after the timed regions (for runs under ncu)")
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
