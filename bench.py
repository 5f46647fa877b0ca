#!/usr/bin/env python
"""bench.py -- raw MB/s of the TEZip predict-delta-encode hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (config.workload): BASELINE.json configs[1] -- error-bounded lossy compression (`-m abs -b 2`, i.e.
floor(1e-2 * 255) levels) of 1000 synthetic 128x160x3 uint8 frames with a random-init 4-layer
(3,48,96,192) PredNet (weight set (ii): U(+-0.1) biases), static window 10, p = 0.  A "step" is one compress
pass over the whole sequence; the decompress pass over the same container is timed the same way and reported
under "decompress".  N > 1 (torchrun, one rank per GPU): every rank owns one 1000-frame shard of an N*1000-frame
sequence (sharded by whole windows, weak scaling); NCCL only all-reduces the symbol histogram and all-gathers
halo / sizes.

  value  : compress MB/s, frames resident in HBM -> int16 stream + key plane resident in HBM.
  e2e    : same through the array-level API with HOST buffers (pinned): H2D of the frames and D2H of the stream
           and key plane inside the timed region, every step.  Boundary = packed int16 stream + key plane (before
           zstd), the same boundary the reference arm times.  e2e.value is the streaming form of that API (two
           alternating sets of pinned output buffers; step i's download runs under step i+1's kernels; one
           barrier + synchronize around the K steps); e2e.synchronised_each_step is the same call with a device
           synchronize after every step.
  --impl reference : the CPU oracle port of the reference (oracle/, torch-CPU fp32 PredNet + C loops) on a bounded
           sample of the same workload, all host threads.
"""
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

STACK = (3, 48, 96, 192)
H, W, C = 128, 160, 3
WORKLOAD = "lossy abs bound 2 levels (~1e-2), 1000x128x160x3 u8 synthetic frames, 4-layer PredNet (3,48,96,192), SWP window 10, p=0"


def workload_text(args):
    """The default is BASELINE.json's configs[1]; other --frames/--window/--mode/--bound are named as they are."""
    if args.frames == 1000 and args.window == 10 and args.mode == "abs" and list(args.bound) == [2.0]:
        return WORKLOAD
    b = list(args.bound)
    kind = "lossless" if (b and b[0] == 0) else "lossy %s bound %s" % (args.mode, "/".join("%g" % v for v in b))
    return "%s, %dx128x160x3 u8 synthetic frames, 4-layer PredNet (3,48,96,192), SWP window %d, p=0" % (
        kind, args.frames, args.window)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--window", type=int, default=10)
    ap.add_argument("--mode", default="abs")
    ap.add_argument("--bound", type=float, nargs="*", default=[2.0])
    ap.add_argument("--cpu-sample-frames", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dwp", action="store_true",
                    help="BASELINE config 3: dynamic windows (-t T, T calibrated as SURVEY 8(d)) instead of -w; "
                         "--chains sub-ranges per GPU run as a batch, each starting with a forced key frame")
    ap.add_argument("--chains", type=int, default=100)
    ap.add_argument("--no-subrecords", action="store_true", help="skip the bounded config-3 / config-4 sub-records")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="2 (default): BASELINE configs[1], the headline line (with a bounded config-4 sub-record); "
                         "4: BASELINE configs[3], 5000 x 1024x1024x1 u16 frames, lossless, window 10, the sequence "
                         "split by whole windows over the ranks (strong scaling)")
    ap.add_argument("--c4-frames", type=int, default=None, help="config 4: total frames (default 5000; sub-record 200)")
    ap.add_argument("--c4-batch", type=int, default=16, help="config 4: windows in flight per PredNet step")
    ap.add_argument("--c4-cpu-frames", type=int, default=3, help="config 4: frames of the CPU oracle sample")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tf_burst": float(p["bf16_tflops"]),
                "tf_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML from a thread every
    5 ms (nvidia-smi -lms needs > 100 ms for its first line on an 8-GPU box -- longer than a short timed region), with
    the nvidia-smi loop as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.samples, self.stop_flag = None, [], False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])          # CUDA ordinal -> NVML index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nvml = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(self.reasons_fn(self.h))))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            nv = self.nvml
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = [c for c, _r in self.samples]
            reasons = sorted(nm for nm, bit in bits.items() if any(r & bit for _c, r in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ================================================================================================ reference arm
def cpu_port_run(frames, ws, window, mode, bound, threads=None):
    """One compress + one decompress of `frames` with the oracle port; returns seconds and stream sizes."""
    import torch
    from oracle import codec_oracle as co
    from oracle.prednet_oracle import PredNetOracle
    if threads:
        torch.set_num_threads(threads)
    net = PredNetOracle(ws, STACK, STACK)
    tc, td = {}, {}
    t0 = time.perf_counter()
    r = co.compress_arrays(frames, net, 0, window, None, mode, bound, True, timers=tc)
    t1 = time.perf_counter()
    out, _ = co.decompress_arrays(r["key_plane"], r["payload"], net, timers=td)
    t2 = time.perf_counter()
    return {"compress_s": t1 - t0, "decompress_s": t2 - t1, "r": r, "out": out, "stages_c": tc, "stages_d": td}


def run_reference(args):
    import torch
    from tezip_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build as obuild
    obuild.build()
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core
    if args.config == 4:
        return run_reference_config4(args)
    n = args.cpu_sample_frames
    ws = synth.make_weights(STACK, bias="uniform", seed=7)
    frames = synth.make_frames(n, H, W, C, seed=1)
    raw_mb = frames.size / 1e6
    for _ in range(args.warmup):
        cpu_port_run(frames[:min(n, 11)], ws, args.window, args.mode, args.bound)
    tcs, tds = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_port_run(frames, ws, args.window, args.mode, args.bound)
        tcs.append(r["compress_s"]); tds.append(r["decompress_s"])
    wall = time.perf_counter() - t0
    v = raw_mb / float(np.mean(tcs))
    vd = raw_mb / float(np.mean(tds))
    cores = torch.get_num_threads()
    sample = "first %d frames of the workload (%.2f MB raw) per step, oracle port, %d torch threads" % (n, raw_mb, cores)
    line = {"impl": "reference", "metric": "raw_MB_per_s_compress", "value": v, "unit": "MB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(tcs)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 + int64/int16 (CPU)",
            "data": "synthetic", "config": {"workload": workload_text(args), "boundary": "packed int16 stream + key plane, before zstd",
                                            "sample_frames": n},
            "decompress": {"value": vd, "unit": "MB/s", "ms_per_step": 1e3 * float(np.mean(tds))},
            "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores, "kind": "port", "sample": sample,
                             "decompress_value": vd, "host_cpus": os.cpu_count()},
            "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line))


# ================================================================================================ config 4
def run_reference_config4(args):
    """--impl reference --config 4: the oracle extended to 16-bit samples (oracle/wide_oracle.py; the reference itself
    refuses such input, compress.py:106-110) on a short prefix of the workload, all host threads."""
    import torch
    from oracle import wide_oracle as wo
    from oracle.prednet_oracle import PredNetOracle
    from tezip_b200 import synth
    n = max(2, args.c4_cpu_frames)
    ws = synth.make_weights(STACK4, bias="uniform", seed=WSEED4)
    frames = synth.make_frames(n, H4, W4, 1, seed=11, dtype=np.uint16)
    net = PredNetOracle(ws, STACK4, STACK4)
    mb = frames.size * 2 / 1e6
    tcs, tds = [], []
    for i in range(max(1, min(args.warmup, 1)) + max(1, min(args.steps, 2))):
        t0 = time.perf_counter()
        r = wo.compress_arrays(frames, net, 0, 10, None, "abs", [0.0], True)
        t1 = time.perf_counter()
        out, _ = wo.decompress_arrays(r["key_plane"], r["payload"], net)
        t2 = time.perf_counter()
        if i >= 1:
            tcs.append(t1 - t0); tds.append(t2 - t1)
    v, vd, cores = mb / float(np.mean(tcs)), mb / float(np.mean(tds)), torch.get_num_threads()
    sample = "first %d frames of the workload (%.1f MB raw) per step, oracle/wide_oracle.py, %d torch threads" % (n, mb, cores)
    print(json.dumps({"impl": "reference", "metric": "raw_MB_per_s_compress", "value": v, "unit": "MB/s",
                      "n_gpus": args.gpus, "steps": len(tcs), "warmup": 1, "ms_per_step": 1e3 * float(np.mean(tcs)),
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 + int64/int32 (CPU)",
                      "data": "synthetic",
                      "config": {"workload": "lossless, 5000x1024x1024x1 u16 synthetic frames, 4-layer PredNet (1,48,96,192), "
                                             "SWP window 10, p=0, container v2 (int32 codes)", "sample_frames": n},
                      "decompress": {"value": vd, "unit": "MB/s"},
                      "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores, "kind": "port", "sample": sample,
                                       "decompress_value": vd, "roundtrip_exact": bool(np.array_equal(out, frames))},
                      "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


STACK4 = (1, 48, 96, 192)
H4 = W4 = 1024
WSEED4 = 4      # weight seed of the one-channel net: seed 7 (the 3-channel runs) happens to draw an Ahat_0 bias that
                # clips every prediction to exactly 0 for this architecture -- legal, but a degenerate parity check


def synth_frames16_device(nt, dev, seed, t0=0, chunk=50):
    """SURVEY.md 8(d) 16-bit variant, generated on the device (5000 frames are 10.5 GB): 32768 + 20000 * pattern +
    N(0, 64), clipped, u16 [nt, 1024, 1024, 1].  Plumbing, not the timed path."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((nt, H4, W4, 1), dtype=torch.uint16, device=dev)
    y = torch.arange(H4, dtype=torch.float32, device=dev)[None, :, None]
    x = torch.arange(W4, dtype=torch.float32, device=dev)[None, None, :]
    for a in range(0, nt, chunk):
        b = min(nt, a + chunk)
        t = torch.arange(t0 + a, t0 + b, dtype=torch.float32, device=dev)[:, None, None]
        pat = torch.sin(2 * np.pi * (x + 2 * t) / 32.0) * torch.cos(2 * np.pi * (y + t) / 24.0)
        f = 32768.0 + 20000.0 * pat + 64.0 * torch.randn(pat.shape, generator=g, device=dev)
        v = f.round().clamp_(0, 65535).to(torch.int32)
        out[a:b].view(torch.int16).copy_(v.to(torch.int16).view(b - a, H4, W4, 1))   # wraps mod 2^16 == the u16 bits
    return out


def config4_record(args, dev, rank, world, comm, nt_total, with_e2e, cpu_frames):
    """BASELINE configs[3]: lossless compress + decompress of 1024x1024x1 u16 frames (container v2: int32 codes),
    (1,48,96,192) PredNet, window 10; the sequence is split by whole windows over the ranks."""
    import torch
    import torch.distributed as dist
    from tezip_b200 import synth, codec, ops, _lib
    from tezip_b200.prednet import PredNet
    from tezip_b200.dist import shard_ranges
    lib = _lib.load()
    Wn = 10
    a, b = shard_ranges(nt_total, 0, Wn, world)[rank]
    nt = b - a
    ws = synth.make_weights(STACK4, bias="uniform", seed=WSEED4)
    net = PredNet(STACK4, STACK4, weights=ws, input_hw=(H4, W4), max_batch=args.c4_batch, device=dev.index)
    frames = synth_frames16_device(nt, dev, seed=11, t0=a)
    N = nt * H4 * W4
    raw_total = nt_total * H4 * W4 * 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warm, flush=None):
        for _ in range(warm):
            fn()
        if flush is not None:
            flush()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.tz_launch_count()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        if flush is not None:
            flush()
        e1.record()
        torch.cuda.synchronize(dev)
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps, float(t[1]) / steps, lib.tz_launch_count() - l0

    keep = {}

    def compress_dev():
        keep["enc"] = codec.encode_frames(frames, net, 0, Wn, None, "abs", [0.0], True, comm=comm)

    first_mode = 0 if rank == 0 else 1

    def decompress_dev():
        e = keep["enc"]
        keep["out"] = codec.decode_arrays(e.key_plane, e.body, e.table, e.shape, 0, net, first_mode=first_mode)[0]

    steps = max(1, min(args.steps, 2))
    ms_c, _w, launches = timed(compress_dev, steps, 1)
    ms_d, _w, launches_d = timed(decompress_dev, steps, 1)
    exact = bool(torch.equal(keep["out"].view(torch.int16), frames.view(torch.int16)))   # lossless round trip, full size
    t = torch.tensor([1.0 if exact else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    exact = bool(t.item() == 1.0)
    rec = {"workload": "lossless, %dx1024x1024x1 u16 synthetic frames (%.2f GB raw), 4-layer PredNet (1,48,96,192), "
                       "SWP window 10, p=0, container v2 (int32 codes)" % (nt_total, raw_total / 1e9),
           "n_gpus": world, "scaling": "strong (the %d frames are split by whole windows over the ranks)" % nt_total,
           "frames_per_gpu": nt, "windows_in_flight": args.c4_batch,
           "compress": {"value": raw_total / 1e6 / (ms_c * 1e-3), "unit": "MB/s", "ms_per_step": ms_c},
           "decompress": {"value": raw_total / 1e6 / (ms_d * 1e-3), "unit": "MB/s", "ms_per_step": ms_d},
           "lossless_roundtrip_exact": exact, "gpu_launches": int(launches), "gpu_launches_decompress": int(launches_d),
           "table_symbols": int(len(keep["enc"].table)), "prednet_gflop_per_frame": net.flops_per_frame() / 1e9}
    if with_e2e:
        fh = torch.empty((nt, H4, W4, 1), dtype=torch.uint16).pin_memory()
        fh.copy_(frames)
        kh = torch.empty_like(fh).pin_memory()
        bh = torch.empty(N, dtype=torch.int32).pin_memory()
        oh = torch.empty_like(fh).pin_memory()

        def compress_e2e():
            keep["ench"] = codec.encode_frames_host(fh, net, 0, Wn, None, "abs", [0.0], kh, bh, True, comm=comm)
            torch.cuda.synchronize(dev)

        def decompress_e2e():
            e = keep["ench"]
            codec.decode_arrays_host(kh, bh, e.table, e.shape, 0, net, oh, first_mode=first_mode)
            torch.cuda.synchronize(dev)

        _m, wall_ce, _l = timed(compress_e2e, 1, 1)
        _m, wall_de, _l = timed(decompress_e2e, 1, 0)
        rec["compress"]["e2e"] = {"value": raw_total / 1e6 / (wall_ce * 1e-3), "unit": "MB/s",
                                  "h2d_bytes_per_step": int(N * 2), "d2h_bytes_per_step": int(N * 4 + N * 2)}
        rec["decompress"]["e2e"] = {"value": raw_total / 1e6 / (wall_de * 1e-3), "unit": "MB/s",
                                    "h2d_bytes_per_step": int(N * 4 + N * 2), "d2h_bytes_per_step": int(N * 2)}
        rec["e2e_roundtrip_exact"] = bool(np.array_equal(oh.numpy(), fh.numpy())) if nt <= 400 else \
            bool(torch.equal(oh.view(torch.int16)[:200], fh.view(torch.int16)[:200]))
        del fh, kh, bh, oh
    # ---- rooflines: the dominant PredNet kernel (tensor pipe) and the fused wide codec kernels (HBM, 10 B/sample)
    pk = peaks()
    Bk = args.c4_batch
    xin = ops.pad_normalize(frames, torch.arange(Bk, dtype=torch.int32, device=dev) * Wn, H4, W4)
    xout = torch.empty_like(xin)
    names = net.kernels()
    acc = np.zeros(len(names))
    for i in range(4):
        ms = net.next_timed(xin, xout)
        if i >= 1:
            acc += np.array(ms) / 3
    kern = [{"kernel": nm, "ms": float(ms), "tflops": (fl * Bk / (ms * 1e-3) / 1e12) if ms > 0 else 0.0}
            for (nm, fl), ms in zip(names, acc)]
    dom = max(range(len(kern)), key=lambda i: kern[i]["ms"] if kern[i]["kernel"].startswith("conv_tc") else -1)
    total_ms = float(acc.sum())
    rec["roofline"] = {"bound": "tensor", "kernel": kern[dom]["kernel"], "achieved": kern[dom]["tflops"],
                       "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / pk["tf_burst"],
                       "traffic": None, "launch_ms": kern[dom]["ms"],
                       "next_step": {"ms": total_ms, "frames": Bk,
                                     "tflops": net.flops_per_frame() * Bk / (total_ms * 1e-3) / 1e12},
                       "whole_compress_tflops": net.flops_per_frame() * (nt - -(-nt // Wn)) / (ms_c * 1e-3) / 1e12,
                       "kernels": kern, "peak_source": pk["src"] + " bf16 dense burst; fp16 operands; algorithmic FLOPs"}
    e = keep["enc"]
    enc_keep = codec.encode_frames(frames[:min(nt, 200)].contiguous(), net, 0, Wn, None, "abs", [0.0], True, keep_pool=True)
    nn = enc_keep.body.numel()
    pool, slot = enc_keep.pool, torch.from_numpy(enc_keep.pred_slot).to(dev)
    hist = torch.zeros(_lib.TZ_WIDE_BINS, dtype=torch.int64, device=dev)
    ovf = torch.zeros(1, dtype=torch.int64, device=dev)
    lut = torch.empty(_lib.TZ_WIDE_BINS, dtype=torch.int32, device=dev)
    tab = torch.empty(_lib.TZ_WIDE_BINS, dtype=torch.int32, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    obuf = torch.empty(nn, dtype=torch.int32, device=dev)
    fsub = frames[:min(nt, 200)]

    def ev_time(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(reps):
            fn()
        b_.record()
        torch.cuda.synchronize(dev)
        return a_.elapsed_time(b_) / reps

    ms_h = ev_time(lambda: ops.encode16(fsub, pool, slot, None, 0, hist=hist, overflow=ovf))
    ops.build_table16_device(hist, tab, lut, meta)
    ms_t = ev_time(lambda: ops.build_table16_device(hist, tab, lut, meta))
    ms_r = ev_time(lambda: ops.encode16(fsub, pool, slot, None, 1, lut=lut, out=obuf))
    lut_d = torch.from_numpy(ops.decode_lut16(enc_keep.table)).to(dev)
    ms_x = ev_time(lambda: ops.reconstruct(enc_keep.body, tuple(fsub.shape), H4, W4, len(enc_keep.table), lut_d, pool,
                                           slot, enc_keep.key_plane))
    rec["roofline_codec"] = {"bound": "hbm", "kernel": "delta_rank16 (fused residual + delta + rank map)",
                             "achieved": 10.0 * nn / (ms_r * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": 10.0 * nn / (ms_r * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                             "samples": int(nn), "algorithmic_bytes_per_sample": 10,
                             "ms": {"hist_pass": ms_h, "table": ms_t, "rank_pass": ms_r, "reconstruct": ms_x},
                             "reconstruct_achieved": 10.0 * nn / (ms_x * 1e-3) / 1e9,
                             "peak_source": pk["src"] + " copy bandwidth"}
    # ---- CPU baseline: the oracle extended the same way (oracle/wide_oracle.py) on a short prefix
    if rank == 0 and world == 1 and cpu_frames >= 2 and not args.no_cpu_baseline:
        from oracle import build as obuild, wide_oracle as wo
        from oracle.prednet_oracle import PredNetOracle
        obuild.build()
        torch.set_num_threads(os.cpu_count() or 1)
        fr_np = frames[:cpu_frames].cpu().numpy()
        onet = PredNetOracle(ws, STACK4, STACK4)
        t0 = time.perf_counter()
        r = wo.compress_arrays(fr_np, onet, 0, Wn, None, "abs", [0.0], True)
        t1 = time.perf_counter()
        out, _i = wo.decompress_arrays(r["key_plane"], r["payload"], onet)
        t2 = time.perf_counter()
        mb = fr_np.size * 2 / 1e6
        # parity at full frame size: same key plane; residuals within the fp16-operand tolerance of the predictor
        encs = codec.encode_frames(frames[:cpu_frames].contiguous(), net, 0, Wn, None, "abs", [0.0], True, keep_x=True)
        dx = int(np.abs(encs.x.cpu().numpy().ravel().astype(np.int64) - r["x"]).max())
        rec["cpu_baseline"] = {"value": mb / (t1 - t0), "unit": "MB/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "first %d frames (%.1f MB raw), oracle/wide_oracle.py + torch-CPU fp32 PredNet"
                                         % (cpu_frames, mb),
                               "decompress_value": mb / (t2 - t1), "oracle_roundtrip_exact": bool(np.array_equal(out, fr_np)),
                               "max_residual_difference_vs_oracle_levels": dx,
                               "residual_tolerance_levels": int(6e-3 * 65535)}
    net.close()
    return rec


def run_config4(args):
    import torch
    import torch.distributed as dist
    from tezip_b200 import build
    from tezip_b200.dist import ShardComm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = ShardComm(None, dev)
    build.build()
    sampler = ClockSampler(local)
    sampler.start()
    rec = config4_record(args, dev, rank, world, comm, args.c4_frames or 5000, True, max(args.c4_cpu_frames, 0))
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": "raw_MB_per_s_compress", "value": rec["compress"]["value"], "unit": "MB/s", "n_gpus": world,
                "steps": max(1, min(args.steps, 2)), "warmup": 1, "ms_per_step": rec["compress"]["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 x f16 -> f32 (PredNet, tcgen05); int32 codec", "data": "synthetic",
                "config": {"workload": rec["workload"], "frames_per_gpu": rec["frames_per_gpu"],
                           "windows_in_flight": rec["windows_in_flight"],
                           "boundary": "packed int32 stream + u16 key plane, before zstd",
                           "l2": "inputs far larger than L2 (%.1f GB per pass)" % (rec["frames_per_gpu"] * H4 * W4 * 10 / 1e9)},
                "decompress": rec["decompress"], "e2e": rec["compress"].get("e2e"), "gpu_launches": rec["gpu_launches"],
                "clocks": clocks, "roofline": rec["roofline"], "roofline_codec": rec["roofline_codec"],
                "cpu_baseline": rec.get("cpu_baseline"), "lossless_roundtrip_exact": rec["lossless_roundtrip_exact"],
                "e2e_roundtrip_exact": rec.get("e2e_roundtrip_exact"), "table_symbols": rec["table_symbols"],
                "prednet_gflop_per_frame": rec["prednet_gflop_per_frame"]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ================================================================================================ native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from tezip_b200 import synth, codec, ops, _lib, build
    from tezip_b200.prednet import PredNet
    from tezip_b200.dist import ShardComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = ShardComm(None, dev)
    build.build()
    lib = _lib.load()

    nt, Wn, mode, bound = args.frames, args.window, args.mode, list(args.bound)
    ws = synth.make_weights(STACK, bias="uniform", seed=7)
    frames_np = synth.make_frames(nt, H, W, C, seed=1 + rank)
    raw_bytes = frames_np.size
    n_win = -(-nt // Wn)
    net = PredNet(STACK, STACK, weights=ws, input_hw=(H, W), max_batch=min(n_win, 128), device=local)
    frames_host = torch.from_numpy(frames_np).pin_memory()
    frames_dev = frames_host.to(dev)
    N = nt * H * W * C
    body_host = torch.empty(N, dtype=torch.int16).pin_memory()
    keyp_host = torch.empty((nt, H, W, C), dtype=torch.uint8).pin_memory()
    out_host = torch.empty((nt, H, W, C), dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    thr, chains, win = None, 1, Wn
    if args.dwp:
        # T = mean cumulative window MSE at step Wn of an unbounded window (first 2*Wn frames), SURVEY 8(d)
        pool_c = torch.empty((2 * Wn + 2, H, W, C), dtype=torch.float32, device=dev)
        _k, slot_c, _a, _n = codec.run_dwp(net, frames_dev[:2 * Wn].contiguous(), 0, 1e30, pool_c, 1)
        idx = torch.arange(1, Wn + 1, dtype=torch.int32, device=dev)
        sse = ops.window_sse(frames_dev, idx, pool_c[slot_c[1:Wn + 1].to(torch.int64)])
        thr = float(sse.sum().item() / (Wn * H * W * C))
        chains, win = args.chains, None

    def compress_dev():
        return codec.encode_frames(frames_dev, net, 0, win, thr, mode, bound, True, dwp_chains=chains, comm=comm)

    # Back-to-back steps: a step returns as soon as its kernels are queued (defer=True) and the host-side end of the
    # PREVIOUS step's entropy stage (reading its table and flags) runs after the next step has been queued, so the
    # GPU does not idle between steps while the host reads one table and launches the next step's first kernels.
    # Every step's table is still read and checked inside the timed region (flush()).
    prev_enc = [None]

    def settle(enc):
        if prev_enc[0] is not None:
            prev_enc[0].finalize()
        prev_enc[0] = enc
        return enc

    def flush():
        settle(None)

    def compress_dev_stream():
        return settle(codec.encode_frames(frames_dev, net, 0, win, thr, mode, bound, True, dwp_chains=chains, comm=comm,
                                          defer=True))

    def compress_e2e():
        enc = codec.encode_frames_host(frames_host, net, 0, win, thr, mode, bound, keyp_host, body_host, True,
                                       dwp_chains=chains, comm=comm)
        if comm is not None:
            offsets[0] = comm.stream_offsets_async(enc.body.numel())    # sizes / offsets all-gather: queued, not awaited
        torch.cuda.synchronize(dev)
        return enc

    offsets = [None]
    # streaming variant of the host-buffer path: consecutive sequences alternate between two sets of pinned output
    # buffers and the next sequence's kernels do not queue behind the previous one's device->host copies
    host_sets = [(keyp_host, body_host),
                 (torch.empty_like(keyp_host).pin_memory(), torch.empty_like(body_host).pin_memory())]
    seq = [0]

    def compress_e2e_pipelined():
        kh, bh = host_sets[seq[0] & 1]
        seq[0] += 1
        enc = codec.encode_frames_host(frames_host, net, 0, win, thr, mode, bound, kh, bh, True, dwp_chains=chains,
                                       comm=comm, wait_copies=False, defer=True)
        if comm is not None:
            offsets[0] = comm.stream_offsets_async(enc.body.numel())
        in_flight[seq[0] & 1] = enc   # the record owns the device tensors its copies read: keep it until the next but one
        return settle(enc)

    in_flight = [None, None]
    enc0 = compress_dev()
    first_mode, first_x = (0, 0) if rank == 0 else (1, 0)

    def decompress_dev():
        return codec.decode_arrays(enc0.key_plane, enc0.body, enc0.table, enc0.shape, 0, net, first_mode=first_mode,
                                   first_x=first_x)[0]

    def decompress_e2e():
        out, _plan = codec.decode_arrays_host(keyp_host, body_host, enc0.table, enc0.shape, 0, net, out_host,
                                              first_mode=first_mode, first_x=first_x)
        torch.cuda.synchronize(dev)
        return out

    out_sets = [out_host, torch.empty_like(out_host).pin_memory()]
    dseq = [0]
    d_in_flight = [None, None]

    def decompress_e2e_pipelined():
        oh = out_sets[dseq[0] & 1]
        dseq[0] += 1
        d_in_flight[dseq[0] & 1] = codec.decode_arrays_host(keyp_host, body_host, enc0.table, enc0.shape, 0, net, oh,
                                                            first_mode=first_mode, first_x=first_x, wait_copies=False)

    def timed(fn, steps, warm, flush=None):
        for _ in range(warm):
            fn()
        if flush is not None:
            flush()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.tz_launch_count()
        t0 = time.perf_counter()
        e0.record()
        marks = []
        for _ in range(steps):
            fn()
            marks.append(time.perf_counter() - t0)
            if os.environ.get("TZ_BENCH_DEBUG"):
                st = torch.cuda.memory_stats(dev)
                marks[-1] = (marks[-1], st.get("num_device_alloc", -1), st.get("num_device_free", -1),
                             st.get("num_alloc_retries", -1), int(st["reserved_bytes.all.current"] / 1e6))
        if flush is not None:
            flush()
        e1.record()
        torch.cuda.synchronize(dev)
        barrier()
        wall = time.perf_counter() - t0
        if os.environ.get("TZ_BENCH_DEBUG"):
            sys.stderr.write("timed %s: host returned at %s ms (cudaMallocs, cudaFrees, retries, reserved MB), all "
                             "done at %.2f ms\n" % (getattr(fn, "__name__", "?"),
                                                    " ".join("%.2f %r" % (m[0] * 1e3, m[1:]) for m in marks), wall * 1e3))
        ms = e0.elapsed_time(e1)
        launches = lib.tz_launch_count() - l0
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps, float(t[1]) / steps, launches

    sampler = ClockSampler(local)
    sampler.start()
    ms_c, wall_c, launches = timed(compress_dev_stream, args.steps, args.warmup, flush)
    ms_d, wall_d, launches_d = timed(decompress_dev, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_c_sync, _w, _l = timed(compress_dev, args.steps, 1)
    _ms, wall_ce, _ = timed(compress_e2e, args.steps, max(1, args.warmup - 1))
    _ms, wall_de, _ = timed(decompress_e2e, args.steps, max(1, args.warmup - 1))
    # (the streaming paths keep two sequences in flight and three staging buffers cycling: four warm-up steps bring the
    # caching allocator to its steady state -- a cudaMalloc inside the timed region was seen to cost 100+ ms on some
    # boxes)
    _ms, wall_cp, _ = timed(compress_e2e_pipelined, args.steps, max(4, args.warmup), flush)
    torch.cuda.synchronize(dev)
    pipelined_ok = bool(torch.equal(host_sets[0][1], host_sets[1][1]) and torch.equal(host_sets[0][0], host_sets[1][0]))
    # the streaming decoder reads the container the synchronised compress_e2e left in keyp_host / body_host
    compress_e2e()
    _ms, wall_dp, _ = timed(decompress_e2e_pipelined, args.steps, max(4, args.warmup))
    torch.cuda.synchronize(dev)
    dec_ref = decompress_dev()
    pipelined_dec_ok = bool(torch.equal(out_sets[0], out_sets[1]) and torch.equal(out_sets[0], dec_ref.cpu()))
    d_in_flight[:] = [None, None]

    # ---- the platform's ceiling for the e2e number: the same bytes over the same pinned buffers, NO kernels (frames
    # in on one stream, stream + key frames out on another, all ranks at once).  e2e / this = what the path costs on
    # top of the bus.
    n_keys = len(enc0.keys)
    key_bytes = n_keys * H * W * C
    scratch_body = torch.empty(N, dtype=torch.int16, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copy_only():
        with torch.cuda.stream(s_in):
            frames_dev.copy_(frames_host, non_blocking=True)
        with torch.cuda.stream(s_out):
            body_host.copy_(scratch_body, non_blocking=True)
            keyp_host.view(-1)[:key_bytes].copy_(frames_dev.view(-1)[:key_bytes], non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s_in)
        torch.cuda.current_stream(dev).wait_stream(s_out)
        torch.cuda.synchronize(dev)

    _ms, wall_copy, _ = timed(copy_only, args.steps, 1)
    frames_dev.copy_(frames_host)
    keyp_host._tz_key_state = None      # (the ceiling run scribbled over the key plane buffer)

    # ---- N > 1: the concatenation of the shards' streams must be the stream one process writes for the whole sequence
    stream_check = None
    if world > 1:
        nc = 8 * Wn                                    # frames per rank of a small check sequence
        small = torch.from_numpy(synth.make_frames(nc, H, W, C, seed=1000 + rank)).to(dev)
        enc_s = codec.encode_frames(small, net, 0, Wn, None, mode, bound, True, comm=comm)
        gathered = [torch.empty_like(enc_s.body) for _ in range(world)] if rank == 0 else None
        gframes = [torch.empty_like(small) for _ in range(world)] if rank == 0 else None
        dist.gather(enc_s.body.view(torch.uint8).view(-1), [g.view(torch.uint8).view(-1) for g in gathered] if rank == 0 else None, dst=0)
        dist.gather(small.view(-1), [g.view(-1) for g in gframes] if rank == 0 else None, dst=0)
        if rank == 0:
            whole = codec.encode_frames(torch.cat(gframes), net, 0, Wn, None, mode, bound, True)
            stream_check = {"frames": nc * world, "shards": world,
                            "sharded_stream_equals_single_process": bool(torch.equal(torch.cat(gathered), whole.body)),
                            "same_table": bool(np.array_equal(enc_s.table, whole.table))}

    # correctness inside the bench: the decoded frames respect the bound
    dec = decompress_dev()
    maxerr = int((dec.to(torch.int16) - frames_dev.to(torch.int16)).abs().max().item())

    # ---- per-kernel roofline (device events between the launches of one next(); B = windows in flight)
    pk = peaks()
    Bk = min(n_win, net.max_batch)
    xin = ops.pad_normalize(frames_dev, torch.arange(Bk, dtype=torch.int32, device=dev) * Wn, H, W)
    xout = torch.empty_like(xin)
    names = net.kernels()
    acc = np.zeros(len(names))
    reps = 6
    for i in range(reps + 2):
        ms = net.next_timed(xin, xout)
        if i >= 2:
            acc += np.array(ms)
    acc /= reps
    kern = []
    # executed (not algorithmic) share of the MMA FLOPs: the gate conv of layer l has Cin = 2 S_l + R_l + R_{l+1}
    # (prednet.py:220-222); its R_l slice multiplies the constant r(t=0) and is folded into the bias maps at create
    Lh = len(STACK)
    exec_share = {}
    for l in range(Lh):
        cin = 2 * STACK[l] + STACK[l] + (STACK[l + 1] if l < Lh - 1 else 0)
        exec_share["conv_tc_gates%d" % l] = (cin - STACK[l]) / cin
    for (nm, fl), ms in zip(names, acc):
        tf = (fl * Bk / (ms * 1e-3) / 1e12) if ms > 0 else 0.0
        kern.append({"kernel": nm, "ms": float(ms), "tflops": tf, "tflops_executed": tf * exec_share.get(nm, 1.0)})
    dom = max(range(len(kern)), key=lambda i: kern[i]["ms"] if kern[i]["kernel"].startswith("conv_tc") else -1)
    total_ms = float(acc.sum())
    achieved = kern[dom]["tflops"]
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (same B, same shape)
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if Bk == 100 and nt == 1000:
            traffic = tj["per_launch_dram_bytes"].get(kern[dom]["kernel"])
    except Exception:
        traffic = None
    # executed (not algorithmic) MMA FLOPs: the r_{t-1} slice of every gate kernel is hoisted into the bias maps
    # `achieved` follows SURVEY.md 8(d): ALGORITHMIC FLOPs (full concatenated K for the gate convs) / launch duration.
    # The kernel executes fewer (the r(t-1) K-slice is hoisted into per-pixel bias maps), so the honest utilisation
    # figure is `frac` = EXECUTED FLOPs against the BURST peak (a ~0.2 ms kernel timed between events is a burst
    # measurement); the algorithmic figure against the sustained peak is kept beside it.
    roofline = {"bound": "tensor", "kernel": kern[dom]["kernel"], "achieved": achieved, "peak": pk["tf_burst"],
                "unit": "TFLOP/s", "frac": kern[dom]["tflops_executed"] / pk["tf_burst"],
                "achieved_executed": kern[dom]["tflops_executed"],
                "frac_algorithmic_vs_sustained": achieved / pk["tf_sustained"], "peak_sustained": pk["tf_sustained"],
                "traffic": traffic,
                "traffic_source": "static: DRAM bytes per launch from the committed ncu --set full capture of this "
                                  "command (profiles/r2_traffic.json, same B and shape); not re-measured in this run",
                "flops_convention": "achieved = algorithmic FLOPs of SURVEY.md 8(d); frac = executed FLOPs / burst peak",
                "peak_source": "%s cuBLAS bf16 dense burst (fp16 operands run at the same rate)" % pk["src"],
                "launch_ms": kern[dom]["ms"], "share_of_next": kern[dom]["ms"] / total_ms,
                "next_step": {"ms": total_ms, "tflops": net.flops_per_frame() * Bk / (total_ms * 1e-3) / 1e12,
                              "frac_algorithmic_vs_sustained": net.flops_per_frame() * Bk / (total_ms * 1e-3) / 1e12 / pk["tf_sustained"]},
                "kernels": kern}

    # ---- codec kernels against the HBM roofline (7 B/sample algorithmic)
    enc_keep = codec.encode_frames(frames_dev, net, 0, Wn, None, mode, bound, True, keep_pool=True, comm=None)
    pool, slot = enc_keep.pool, torch.from_numpy(enc_keep.pred_slot).to(dev)
    apply_t = torch.from_numpy((enc_keep.pred_slot >= 0).astype(np.uint8)).to(dev)
    hist = torch.zeros(_lib.TZ_HIST_BINS, dtype=torch.int64, device=dev)
    ovf = torch.zeros(1, dtype=torch.int64, device=dev)
    lut = torch.from_numpy(ops.encode_lut(enc_keep.table)).to(dev)
    xbuf = torch.empty((nt, H, W, C), dtype=torch.int16, device=dev)
    obuf = torch.empty(N, dtype=torch.int16, device=dev)

    def ev_time(fn, reps=5):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    codec_ms = {
        "residual": ev_time(lambda: ops.residual(frames_dev, pool, slot, out=xbuf)),
        "residual+error_bound": ev_time(lambda: (ops.residual(frames_dev, pool, slot, out=xbuf),
                                                 ops.error_bound(frames_dev, xbuf, apply_t, mode, bound))),
        "delta_hist": ev_time(lambda: ops.finding_difference_hist(xbuf, hist, ovf)),
        "delta_rank": ev_time(lambda: ops.finding_difference_rank(xbuf, lut, out=obuf)),
        "fused_lossless_hist": ev_time(lambda: ops.encode_lossless(frames_dev, pool, slot, 0, hist=hist, overflow=ovf)),
        "fused_lossless_rank": ev_time(lambda: ops.encode_lossless(frames_dev, pool, slot, 1, lut=lut, out=obuf)),
    }
    if ops.encode_lossy_supported(frames_dev, mode) and not codec.is_lossless(mode, bound):
        hl = torch.zeros(_lib.TZ_HIST_BINS + 2, dtype=torch.int64, device=dev)

        def fused_lossy():
            hl[-1:].zero_()
            ops.encode_lossy(frames_dev, pool, slot, apply_t, mode, bound, hl[:_lib.TZ_HIST_BINS], hl[-2:-1],
                             hl[-1:].view(torch.int32), x=xbuf)
        codec_ms["fused_lossy_pass(+4B memset)"] = ev_time(fused_lossy)
    lut_d = torch.from_numpy(ops.decode_lut(enc_keep.table)).to(dev)
    codec_ms["reconstruct"] = ev_time(lambda: ops.reconstruct(enc_keep.body, (nt, H, W, C), H, W, len(enc_keep.table),
                                                              lut_d, pool, slot, enc_keep.key_plane))
    codec_ms["error_bound"] = codec_ms["residual+error_bound"] - codec_ms["residual"]
    enc_total = codec_ms["residual+error_bound"] + codec_ms["delta_hist"] + codec_ms["delta_rank"]
    enc_total_separate = enc_total
    if "fused_lossy_pass(+4B memset)" in codec_ms:     # what encode_with_pool runs: one fused pass + the rank map
        enc_total = codec_ms["fused_lossy_pass(+4B memset)"] + codec_ms["delta_rank"]
    roofline_codec = {"bound": "hbm", "kernel": "fused_lossless_rank (residual+delta+rank map)",
                      "achieved": 7.0 * N / (codec_ms["fused_lossless_rank"] * 1e-3) / 1e9, "peak": pk["hbm"],
                      "unit": "GB/s", "traffic": None,
                      "lossy_encode_total": {"ms": enc_total, "achieved": 7.0 * N / (enc_total * 1e-3) / 1e9,
                                             "frac": 7.0 * N / (enc_total * 1e-3) / 1e9 / pk["hbm"],
                                             "launches": "tz_encode_lossy (residual + error bound + delta histogram) "
                                                         "+ tz_delta_rank",
                                             "separate_kernels_ms": enc_total_separate},
                      "reconstruct": {"ms": codec_ms["reconstruct"],
                                      "achieved": 7.0 * N / (codec_ms["reconstruct"] * 1e-3) / 1e9},
                      "ms": codec_ms, "peak_source": pk["src"] + " copy bandwidth"}
    roofline_codec["frac"] = roofline_codec["achieved"] / pk["hbm"]

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample
    cpu = None
    ratio = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import build as obuild
        from tezip_b200 import container
        obuild.build()
        torch.set_num_threads(os.cpu_count() or 1)
        ns = args.cpu_sample_frames
        t0 = time.perf_counter()
        r = cpu_port_run(frames_np[:ns], ws, Wn, mode, bound)
        sample_mb = ns * H * W * C / 1e6
        cores = torch.get_num_threads()
        cpu = {"value": sample_mb / r["compress_s"], "unit": "MB/s", "cores": cores, "kind": "port",
               "sample": "first %d frames (%.2f MB raw), one compress + one decompress, oracle port" % (ns, sample_mb),
               "decompress_value": sample_mb / r["decompress_s"], "host_cpus": os.cpu_count(),
               "stages_compress_s": {k: round(v, 4) for k, v in r["stages_c"].items()},
               "stages_decompress_s": {k: round(v, 4) for k, v in r["stages_d"].items()}}
        # compression ratio, same sample, same bound: native vs oracle (zstd level 9 like the reference)
        encs = codec.encode_frames(frames_dev[:ns].contiguous(), net, 0, Wn, None, mode, bound, True)
        nat = len(container.zstd_compress(encs.payload())) + len(container.zstd_compress(encs.key_plane.cpu().numpy()))
        orc = len(container.zstd_compress(r["r"]["payload"])) + len(container.zstd_compress(r["r"]["key_plane"]))
        ratio = {"native": ns * H * W * C / nat, "oracle": ns * H * W * C / orc, "native_over_oracle": orc / nat,
                 "sample_frames": ns}

    # ---- BASELINE config 5: decompress-only throughput for windows 5 / 10 / 20 (device-resident, same frames)
    sweep = {}
    for wn in (5, 20):
        nwin = -(-nt // wn)
        if nwin > net.max_batch:
            net_w = PredNet(STACK, STACK, weights=ws, input_hw=(H, W), max_batch=min(nwin, 256), device=local)
        else:
            net_w = net
        enc_w = codec.encode_frames(frames_dev, net_w, 0, wn, None, mode, bound, True, comm=comm)

        def dec_w():
            return codec.decode_arrays(enc_w.key_plane, enc_w.body, enc_w.table, enc_w.shape, 0, net_w,
                                       first_mode=first_mode, first_x=first_x)[0]
        ms_w, _w, _l = timed(dec_w, max(2, args.steps // 2), 1)
        sweep["w%d" % wn] = {"decompress_MB_per_s": world * raw_bytes / 1e6 / (ms_w * 1e-3), "ms": ms_w}
        if net_w is not net:
            net_w.close()
    sweep["w%d" % Wn] = {"decompress_MB_per_s": world * raw_bytes / 1e6 / (ms_d * 1e-3), "ms": ms_d}

    # ---- BASELINE configs[0] (lossless, 100 frames, window 10 -- the reference's own CPU-runnable case) as a sub-record:
    # only ten windows are in flight, so this is the small-batch end of the same kernels
    c1_rec = None
    if not args.dwp and not args.no_subrecords:
        n1 = min(100, nt)
        fr1 = frames_dev[:n1].contiguous()
        keep1 = {}

        def c1_compress():
            keep1["enc"] = codec.encode_frames(fr1, net, 0, 10, None, "abs", [0.0], True, comm=comm)

        def c1_decompress():
            e = keep1["enc"]
            keep1["out"] = codec.decode_arrays(e.key_plane, e.body, e.table, e.shape, 0, net, first_mode=first_mode,
                                               first_x=first_x)[0]
        ms1c, _w, l1c = timed(c1_compress, args.steps, 2)
        ms1d, _w, _l = timed(c1_decompress, args.steps, 2)
        c1_rec = {"workload": "lossless (abs 0), %d x 128x160x3 u8 frames per GPU, window 10, p=0 (BASELINE configs[0])" % n1,
                  "compress": {"value": world * fr1.numel() / 1e6 / (ms1c * 1e-3), "unit": "MB/s", "ms_per_step": ms1c},
                  "decompress": {"value": world * fr1.numel() / 1e6 / (ms1d * 1e-3), "unit": "MB/s", "ms_per_step": ms1d},
                  "lossless_roundtrip_exact": bool(torch.equal(keep1["out"], fr1)), "windows_in_flight": -(-n1 // 10),
                  "table_symbols": int(len(keep1["enc"].table)), "gpu_launches": int(l1c)}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from tezip_b200 import container as tzc1
            e1 = keep1["enc"]
            c1_rec["ratio_zstd9"] = fr1.numel() / float(len(tzc1.zstd_compress(e1.payload())) +
                                                      len(tzc1.zstd_compress(e1.key_plane.cpu().numpy())))
        del keep1, fr1

    # ---- BASELINE configs[2] (dynamic windows, 10k frames over 8 GPUs = 1250 per GPU) as a sub-record of every line
    dwp_rec = None
    if not args.dwp and not args.no_subrecords:
        nt3, ch3 = 1250, 100
        fr3_np = synth.make_frames(nt3, H, W, C, seed=2001 + rank)
        fr3 = torch.from_numpy(fr3_np).to(dev)
        pool_c = torch.empty((2 * Wn + 2, H, W, C), dtype=torch.float32, device=dev)
        _k, slot_c, _a, _n = codec.run_dwp(net, fr3[:2 * Wn].contiguous(), 0, 1e30, pool_c, 1)
        idx3 = torch.arange(1, Wn + 1, dtype=torch.int32, device=dev)
        sse3 = ops.window_sse(fr3, idx3, pool_c[slot_c[1:Wn + 1].to(torch.int64)])
        thr3 = float(sse3.sum().item() / (Wn * H * W * C))       # cumulative window MSE at step Wn
        if world > 1:                                             # one T for the whole job: rank 0's
            t3 = torch.tensor([thr3], dtype=torch.float64, device=dev)
            dist.broadcast(t3, 0)
            thr3 = float(t3.item())
        keep3 = {}

        def compress_dwp():
            keep3["enc"] = codec.encode_frames(fr3, net, 0, None, thr3, mode, bound, True, dwp_chains=ch3, comm=comm)

        ms3, _w3, l3 = timed(compress_dwp, max(2, args.steps // 2), 2)
        enc3 = keep3["enc"]
        dec3 = codec.decode_arrays(enc3.key_plane, enc3.body, enc3.table, enc3.shape, 0, net, first_mode=first_mode,
                                   first_x=first_x)[0]
        err3 = int((dec3.to(torch.int16) - fr3.to(torch.int16)).abs().max().item())
        dwp_rec = {"workload": "dynamic windows (-t), %d frames per GPU x %d GPUs, 128x160x3 u8, %s, %d chains per GPU "
                               "(forced key frame at every chain start)" % (nt3, world, workload_text(args).split(",")[0], ch3),
                   "threshold": thr3, "threshold_rule": "cumulative window MSE at step %d of an unbounded window "
                                                        "(SURVEY.md 8(d)), rank 0's frames" % Wn,
                   "compress": {"value": world * fr3_np.size / 1e6 / (ms3 * 1e-3), "unit": "MB/s", "ms_per_step": ms3},
                   "key_frames_rank0": len(enc3.keys), "max_abs_error_levels": err3, "gpu_launches": int(l3),
                   "host_syncs_per_step": "none inside the scan (tz_dwp_gather / tz_window_sse / tz_dwp_update)"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import codec_oracle as co
            from oracle.prednet_oracle import PredNetOracle
            from tezip_b200 import container as tzc
            npre = 60
            r3 = co.compress_arrays(fr3_np[:npre], PredNetOracle(ws, STACK, STACK), 0, None, thr3, mode, bound, True)
            seq = codec.encode_frames(fr3[:npre].contiguous(), net, 0, None, thr3, mode, bound, True, dwp_chains=1)
            bat = codec.encode_frames(fr3[:npre].contiguous(), net, 0, None, thr3, mode, bound, True,
                                      dwp_chains=max(1, npre * ch3 // nt3))

            def size(payload, kp):
                return len(tzc.zstd_compress(payload)) + len(tzc.zstd_compress(kp))
            so = size(r3["payload"], r3["key_plane"])
            dwp_rec["ratio_vs_sequential_oracle"] = {
                "prefix_frames": npre, "oracle_keys": [int(k) for k in r3["keys"]],
                "native_sequential_keys_equal_oracle": seq.keys == [int(k) for k in r3["keys"]],
                "oracle": npre * H * W * C / so,
                "native_sequential": npre * H * W * C / size(seq.payload(), seq.key_plane.cpu().numpy()),
                "native_batched_chains": npre * H * W * C / size(bat.payload(), bat.key_plane.cpu().numpy()),
                "batched_key_frames": len(bat.keys), "sequential_key_frames": len(seq.keys)}
        del fr3, keep3, enc3, dec3

    # ---- BASELINE configs[3] (1024x1024x1 u16, lossless, container v2) as a bounded sub-record
    c4_rec = None
    if not args.dwp and not args.no_subrecords:
        torch.cuda.empty_cache()
        c4_rec = config4_record(args, dev, rank, world, comm, (args.c4_frames or 200) * world, False, 0)
        c4_rec["note"] = "bounded sample of BASELINE configs[3] (%d frames per GPU); `bench.py --config 4` runs the " \
                         "5000-frame sequence" % (args.c4_frames or 200)

    # ---- the container stage that follows the hot path: zstd level 9, one frame, all host cores
    cont = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tezip_b200 import container as tzc
        payload = enc0.payload()
        kp_np = enc0.key_plane.cpu().numpy()
        nw = os.cpu_count() or 1
        cont = {"workers": nw, "note": "single zstd frames with content size (what the reference decoder needs), all "
                                       "host threads (ZSTD_c_nbWorkers); not part of value/e2e.  Level 9 is the "
                                       "reference's; TEZIP_ZSTD_LEVEL selects another"}
        for lvl in (9, 3, 1):
            t0 = time.perf_counter()
            zb = tzc.zstd_compress(payload, lvl, workers=nw)
            zk = tzc.zstd_compress(kp_np, lvl, workers=max(1, nw // 4))
            tz_s = time.perf_counter() - t0
            cont["level%d" % lvl] = {"seconds": tz_s, "raw_MB_per_s": raw_bytes / 1e6 / tz_s,
                                     "ratio": raw_bytes / float(len(zb) + len(zk))}
        # the same two frames written on the GPU (tezip_b200/zstd_frames.py, TEZIP_ZSTD_LEVEL=gpu): Huffman-coded
        # literal blocks, no match finding -- the decoder is still libzstd's
        from tezip_b200 import zstd_frames as tzf
        tail_dev = torch.from_numpy(payload[enc0.body.numel():]).to(dev)
        pay_dev = torch.cat([enc0.body.reshape(-1), tail_dev])
        for _warm in range(2):
            tzf.frame_device(pay_dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        marks = []
        fb, nb_ = tzf.frame_device(pay_dev, marks)
        fk, nk_ = tzf.frame_device(enc0.key_plane, marks)
        torch.cuda.synchronize(dev)
        t_dev = time.perf_counter() - t0
        kern_ms = {}
        for nm, a_, b_ in marks:
            kern_ms[nm] = kern_ms.get(nm, 0.0) + a_.elapsed_time(b_)
        src_bytes = pay_dev.numel() * 2 + enc0.key_plane.numel()
        zb_g, zk_g = fb[:nb_].cpu().numpy(), fk[:nk_].cpu().numpy()
        tzf.frame_host(pay_dev)                                      # (allocates the pinned landing buffer)
        t0 = time.perf_counter()
        n_h = tzf.frame_host(pay_dev).size + tzf.frame_host(enc0.key_plane).size
        t_all = time.perf_counter() - t0
        assert n_h == nb_ + nk_
        cont["gpu"] = {"kernels_ms": kern_ms, "source_bytes": src_bytes,
                       "kernels_GB_per_s_of_source": src_bytes / 1e6 / max(sum(kern_ms.values()), 1e-9),
                       "seconds_device": t_dev, "seconds_with_download": t_all, "raw_MB_per_s": raw_bytes / 1e6 / t_all,
                       "raw_MB_per_s_device": raw_bytes / 1e6 / t_dev, "ratio": raw_bytes / float(nb_ + nk_),
                       "decodes_with_libzstd": bool(np.array_equal(tzc.zstd_decompress(zb_g).view("<i2"), payload) and
                                                    np.array_equal(tzc.zstd_decompress(zk_g), kp_np.reshape(-1))),
                       "note": "frames written by CUDA kernels from the device copies: per 128 KB block RLE, raw or "
                               "Huffman literals + zero sequences (RFC 8878); seconds_device includes the two small "
                               "host round trips (histogram, size), seconds_with_download the copy of the frames to "
                               "host memory"}
        # ... and read back: header walk on the host, all blocks decoded at once on the device, against libzstd's decoder
        # on the same two frames (one host thread, as zstd.decompress)
        t0 = time.perf_counter()
        tzc.zstd_decompress(zb_g), tzc.zstd_decompress(zk_g)
        t_lib = time.perf_counter() - t0
        tzf.decompress_device(zb_g, dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        dmarks = []
        db, dk = tzf.decompress_device(zb_g, dev, marks=dmarks), tzf.decompress_device(zk_g, dev, marks=dmarks)
        torch.cuda.synchronize(dev)
        t_gdec = time.perf_counter() - t0
        t0 = time.perf_counter()
        tzf.parse_frame(zb_g), tzf.parse_frame(zk_g)
        t_parse = time.perf_counter() - t0
        cont["gpu"]["decode"] = {"seconds": t_gdec, "kernels_ms": sum(a_.elapsed_time(b_) for _n, a_, b_ in dmarks),
                                 "host_header_walk_seconds": t_parse, "raw_MB_per_s": raw_bytes / 1e6 / t_gdec,
                                 "libzstd_seconds_same_frames": t_lib,
                                 "identical_to_source": bool(torch.equal(db.view(torch.int16), pay_dev) and
                                                             torch.equal(dk, enc0.key_plane.reshape(-1))),
                                 "note": "host buffers in (pageable), decoded content left in device memory where "
                                         "the decompressor needs it; includes the upload of the compressed frames"}
        del fb, fk, pay_dev, db, dk
        cont["zstd_level"], cont["seconds"] = 9, cont["level9"]["seconds"]
        cont["raw_MB_per_s"], cont["ratio"] = cont["level9"]["raw_MB_per_s"], cont["level9"]["ratio"]

    if rank == 0:
        total_mb = world * raw_bytes / 1e6
        line = {
            "metric": "raw_MB_per_s_compress", "value": total_mb / (ms_c * 1e-3), "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_c, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 x f16 -> f32 (PredNet, tcgen05); int16/f64 codec",
            "data": "synthetic",
            "config": {"workload": workload_text(args) if not args.dwp else workload_text(args).replace(
                           "SWP window %d" % args.window, "DWP threshold %.6g (calibrated), %d chains per GPU, %d key frames" %
                           (thr, chains, len(enc0.keys))),
                       "frames_per_gpu": nt, "windows_in_flight": Bk,
                       "boundary": "packed int16 stream + key plane, before zstd",
                       "l2": "no explicit flush: one step streams ~%.1f GB (prediction pool + activations) >> 126 MB L2"
                             % ((enc_keep.pool.numel() * 4 + net.device_bytes()) / 1e9),
                       "sharding": "by window, %d ranks" % world},
            "decompress": {"value": total_mb / (ms_d * 1e-3), "unit": "MB/s", "ms_per_step": ms_d,
                           "e2e": {"value": total_mb / (wall_dp * 1e-3), "unit": "MB/s",
                                   "h2d_bytes_per_step": int(N * 3), "d2h_bytes_per_step": int(N),
                                   "mode": "streaming decoder (codec.decode_arrays_host, wait_copies=False, two "
                                           "alternating pinned output buffers): the next sequence's key plane is "
                                           "uploaded and scanned on the upload stream while the current one is "
                                           "predicted, downloads run on their own stream; one synchronize around the "
                                           "K steps",
                                   "outputs_identical_to_device_path": pipelined_dec_ok,
                                   "synchronised_each_step": {"value": total_mb / (wall_de * 1e-3), "unit": "MB/s"}}},
            "e2e": {"value": total_mb / (wall_cp * 1e-3), "unit": "MB/s", "h2d_bytes_per_step": int(N),
                    "d2h_bytes_per_step": int(N * 2 + key_bytes),
                    "mode": "streaming compressor through the host-buffer API (codec.encode_frames_host, "
                            "wait_copies=False): every step copies its frames in from pinned host memory and its "
                            "int16 stream + key frames out to one of two alternating sets of pinned buffers; the "
                            "device->host copies of step i run under the kernels of step i+1; the K steps are "
                            "bracketed by a barrier + synchronize on both sides, so every copy of every step is "
                            "inside the timed region",
                    "both_buffer_sets_identical": pipelined_ok,
                    "synchronised_each_step": {
                        "value": total_mb / (wall_ce * 1e-3), "unit": "MB/s",
                        "note": "the same call with a device synchronize after every step (one isolated sequence: "
                                "its 123 MB stream download cannot start before the table exists, i.e. after the "
                                "last kernel, and nothing hides it)"},
                    "copy_ceiling": {"value": total_mb / (wall_copy * 1e-3), "unit": "MB/s", "ms_per_step": wall_copy,
                                     "note": "the same H2D + D2H bytes over the same pinned buffers with no kernels, "
                                             "all ranks at once: the bus / host-memory limit of this box",
                                     "e2e_over_ceiling": wall_copy / wall_cp,
                                     "synchronised_over_ceiling": wall_copy / wall_ce}},
            "stream_check": stream_check,
            "gpu_launches": int(launches), "gpu_launches_decompress": int(launches_d),
            "wall_ms_per_step": wall_c,
            "ms_per_step_host_synchronised": ms_c_sync,   # encode_frames(defer=False): the host reads every step's
                                                          # table before it queues the next step (GPU idles meanwhile)
            "clocks": clocks, "max_abs_error_levels": maxerr,
            "roofline": roofline, "roofline_codec": roofline_codec, "cpu_baseline": cpu, "ratio": ratio,
            "decompress_sweep": sweep, "container": cont, "config1": c1_rec, "dwp": dwp_rec, "config4": c4_rec,
            "prednet_gflop_per_frame": net.flops_per_frame() / 1e9,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == 4:
        run_config4(a)
    else:
        run_native(a)
