#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 YOLOv8 detector (contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W              # this repo's engine
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload (BASELINE.json `metric`, configs[2]): YOLOv8n 640x640, batch 64 per GPU, bf16 tensor-core
path, nc=80, synthetic frames + seeded random-init weights.  One "step" = one pass of the whole hot
path (preprocess -> 63 convs -> DFL decode -> filter -> NMS -> D2H of the detections) over one batch.
`value` = frames/s with the frames already resident in HBM (4 rotating input sets, 315 MB > L2),
timed with CUDA events on the engine's stream inside the C library; `e2e` = the same metric through
zl_infer_batch() with HOST (pinned) frames, H2D and D2H copies inside the timed region.
Multi-GPU: frames never share state, so ranks are independent replicas fed their own batches
(no collective on the data path; NCCL is used only for the barrier and the max-over-ranks time).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))

import numpy as np  # noqa: E402

SCALE, NC, HW, BATCH = "n", 80, 640, 64
CONF, IOU = 0.5, 0.45
WORKLOAD = f"YOLOv8{SCALE} {HW}x{HW} batch {BATCH} per GPU, bf16 tcgen05 path, nc={NC}, conf {CONF} iou {IOU}"
METRIC, UNIT = "frames_per_sec_yolov8n_640_b64", "frames/s"
FLOPS_PER_FRAME = 8.7429e9          # conv FLOPs (2*MAC), v8n 640 nc=80


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled back to back from a thread
    (nvidia-smi -lms cannot start fast enough for a region of tens of milliseconds); nvidia-smi as a fallback."""

    def __init__(self, gpu_index):
        self.idx, self.samples, self.reasons, self.max_mhz = gpu_index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        self.src = "none"

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            for _ in range(3):            # the first queries after nvmlInit take tens of ms (a whole timed region): take them here
                pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)
                try:
                    pynvml.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            self.src = "nvml"
        except Exception:
            self._nvml = None
            self.src = "nvidia-smi"
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        time.sleep(0.02)          # first samples land before the timed region opens

    def _run(self):
        if self._nvml is not None:
            n = self._nvml
            bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self._stop.is_set():
                try:
                    self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                    try:
                        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.0005)        # an NVML query takes 1-30 ms on a loaded box: poll nearly back to back
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self._phys_index()), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                time.sleep(0.05)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": self.src}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": self.src}


def shard_frames(n_total, rank, world):
    """Frame-sharding rule of the multi-GPU path: frames are independent units, rank r
    takes the contiguous slice [lo, hi) and never exchanges anything with the other ranks."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_max(value, dist, device="cpu"):
    """Max over ranks of a per-rank scalar (the only collective in the bench: timing, not data)."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_inputs(n_sets, seed0=5678):
    from oracle import synth
    return [synth.frames_structured(BATCH, HW, HW, seed=seed0 + s) for s in range(n_sets)]


def make_model():
    from oracle import yolov8_ref, zlw
    tensors = yolov8_ref.synthetic_model(SCALE, NC, seed=0)
    return tensors, zlw.dumps(tensors, SCALE, NC)


# ----------------------------------------------------------------------------- CPU legs (oracle port)
def cpu_pipeline_fps(tensors, frames, threads, batch):
    """Reference path on host cores: P1 (C) -> YOLOv8 (torch CPU fp32, stand-in for the ORT CPU session)
    -> F1/N1 (C).  batch=1 is how the reference runs (onnx_engine.cpp:348-352 never batches)."""
    import torch
    from oracle import oracle_c, yolov8_ref
    torch.set_num_threads(threads)
    sess = yolov8_ref.Session(tensors, SCALE, NC)
    t0 = time.perf_counter()
    ndet = 0
    for i0 in range(0, len(frames), batch):
        chunk = frames[i0:i0 + batch]
        x = np.stack([oracle_c.preprocess(f, HW, HW, HW, HW)[1] for f in chunk])
        raw = sess.run(x)
        for j in range(len(chunk)):
            d, _ = oracle_c.postprocess(raw[j], HW, HW, CONF, IOU)
            ndet += len(d)
    dt = time.perf_counter() - t0
    return len(frames) / dt, dt, ndet


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  It cannot be compiled here
    (ONNX Runtime and the model are absent, sources have compile errors: SURVEY.md §0 fact 5), so this
    is the oracle port, b=1 sequential like the reference, all host threads."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    tensors, _ = make_model()
    sample = 8                                       # frames per step: bounded sample of the 64-frame batch
    frames = list(make_inputs(1)[0][:sample])
    for _ in range(max(args.warmup, 1)):
        cpu_pipeline_fps(tensors, frames[:2], cores, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline_fps(tensors, frames, cores, 1)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.replace("bf16 tcgen05 path", "fp32 CPU path"), "sample": f"{sample} frames per step, b=1 sequential"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps x {sample} frames, b=1 sequential, torch-CPU stand-in for ORT-CPU + C pre/post"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
# algorithmic work per FRAME at 640x640, nc = 80 (conv FLOPs = 2*MAC; activation bytes = every conv
# output written once and read once, 16-bit) — the roofline numerators, never the engine's per-layer operand sums.
ALGO = {"n": dict(flops=8.743e9, act_bytes=60.4e6), "s": dict(flops=28.602e9, act_bytes=109.5e6), "m": dict(flops=78.936e9, act_bytes=193.3e6)}
# north_star's 16-bit gate, as measured by tests/test_gpu_engine.py::test_16bit_contract_* and scripts/gpu_diag.py on B200
GATE = {"fp16": "pass (every matched detection IoU >= 0.99; 2-4 % NMS knife-edge flips; profiles/accuracy_r02.md)",
        "bf16": "fail (19-45 % of the matched detections reach IoU 0.99, all reach 0.9; profiles/accuracy_r02.md)"}


def conv_family(prof):
    # tcgen05 conv kernels: persistent (9), fallback (1) and the fused head kernel (12: the last six 1x1 convs of the Detect
    # head + decode; its whole time is charged to the family, so the roofline numerator still covers all 60 convs)
    tc = [p for p in prof if p["kind"] in (1, 9, 12)]
    return tc, sum(p["ms"] for p in tc)


def measure_resident(eng, n_sets, steps, warmup):
    eng.run_resident(n_sets, max(warmup, 3))
    ms, launches, _ = eng.run_resident(n_sets, steps)
    return ms, launches


def side_config(zlb200, scale, batch, prec, device, steps, peaks, dtype):
    """BASELINE configs[3] models (YOLOv8s / YOLOv8m 640x640, batch 256 sharded over the ranks): frames/s of this rank's
    shard with the frames resident, plus the tensor-roof fraction (the compute-bound models of the family)."""
    from oracle import synth, yolov8_ref, zlw
    t = yolov8_ref.synthetic_model(scale, NC, seed=0)
    lanes = 2
    e = zlb200.Engine(HW, HW, NC, scale, precision=prec, conf=CONF, iou=IOU, max_batch=batch, device=device, num_lanes=lanes)
    e.load_weights_blob(zlw.dumps(t, scale, NC))
    e.warmup(1)
    frames = synth.frames_structured(min(batch, 32), HW, HW, seed=900)
    reps = [frames[i % len(frames)] for i in range(batch)]
    for sidx in range(2):
        e.upload_resident(sidx, reps)
    ms, _ = measure_resident(e, 2, steps, 3)
    fps = batch * steps / (ms / 1e3)
    prof = e.profile(0, 1)
    tc, tc_ms = conv_family(prof)
    e.close()
    return {"model": f"yolov8{scale}", "batch_per_gpu": batch, "dtype": dtype, "frames_per_s_this_gpu": fps, "ms_per_step": ms / steps,
            "tensor_tflops_whole_step": fps * ALGO[scale]["flops"] / 1e12, "tensor_frac_of_sustained": fps * ALGO[scale]["flops"] / (peaks["tf_sust"] * 1e12),
            "conv_tflops_kernels_only": ALGO[scale]["flops"] * batch / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
            "conv_hbm_gbs_algorithmic": ALGO[scale]["act_bytes"] * batch / (tc_ms * 1e-3) / 1e9 if tc_ms else None, "conv_launches": len(tc)}


def run_ours(args, rank, local_rank, world):
    import torch
    import zlb200
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    # keep this rank's host threads (and the pinned buffers they first touch) on the cores next to its GPU
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].strip().isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        if cpus and world > 1:
            share = max(len(cpus) // world, 1)
            mine = cpus[(local_rank * share) % len(cpus):][:share] or cpus
            os.sched_setaffinity(0, set(mine))
    except Exception:
        pass

    peaks = load_peaks()
    tensors, blob = make_model()
    n_sets = 4
    sets = make_inputs(n_sets, seed0=5678 + 100 * rank)
    PREC = {"fp16": zlb200.FP16, "bf16": zlb200.BF16}
    prec = PREC[args.dtype]
    LANES = args.e2e_threads   # host threads in the end-to-end leg: each owns a lane, so H2D of one batch overlaps compute of another
    eng = zlb200.Engine(HW, HW, NC, SCALE, precision=prec, conf=CONF, iou=IOU, max_batch=BATCH, device=local_rank, num_lanes=LANES)
    eng.load_weights_blob(blob)
    eng.warmup(1)
    for s in range(n_sets):
        eng.upload_resident(s, list(sets[s]))

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    eng.run_resident(n_sets, max(args.warmup, 3))
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ms, launches, _ = eng.run_resident(n_sets, args.steps)
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    barrier()
    max_ms = reduce_max(ms, dist, "cuda")
    value = world * BATCH * args.steps / (max_ms / 1e3)

    # ---- end to end through the public C-ABI call (zl_infer_batch) with pinned HOST frames, looped in C so no interpreter
    # sits in the timed path: every step copies its 64 frames host->device (one copy: they are contiguous) and reads the
    # detections back; LANES host threads keep one batch each in flight
    pinned = [zlb200.pinned_array((BATCH, HW, HW, 3)) for _ in range(LANES)]
    for t, buf in enumerate(pinned):
        buf[:] = sets[t % n_sets]
    eng.bench_e2e(pinned, 2 * LANES)
    barrier()
    e2e_s, n_det = eng.bench_e2e(pinned, args.steps)
    barrier()
    e2e_fps = world * BATCH * args.steps / reduce_max(e2e_s, dist, "cuda")
    d2h = (4 + 2 * BATCH) * 4 + min(BATCH * 64, BATCH * eng.A) * 24
    # the e2e leg's ceiling: pinned H2D bandwidth of this rank while every rank measures at the same time
    barrier()
    h2d_peak = eng.bench_h2d(256 << 20, 6)
    h2d_peak_min = -reduce_max(-h2d_peak, dist, "cuda")
    barrier()

    # ---- config 4 (YOLOv8s / YOLOv8m 640x640, batch 256 sharded over the ranks): every rank runs its shard, rank 0 reports
    cfg4 = None
    if not args.quick:
        cfg4 = []
        shard = max(256 // world, 1)
        for sc in ("s", "m"):
            try:
                barrier()
                r = side_config(zlb200, sc, shard, prec, local_rank, 6, peaks, args.dtype)
                slow = reduce_max(r["ms_per_step"], dist, "cuda")
                r["frames_per_s_all_gpus"] = 256 / (slow / 1e3) if world * shard == 256 else world * shard / (slow / 1e3)
                r["global_batch"] = world * shard
                r["scaling"] = "strong (global batch 256 split over the ranks)"
                cfg4.append(r)
            except Exception as ex:
                cfg4.append({"model": f"yolov8{sc}", "error": str(ex)})

    if rank != 0:
        eng.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline (rank 0): the same pass un-captured on ONE stream, a CUDA-event pair around every kernel
    # (serial and un-overlapped, so its sum exceeds the timed step, which overlaps LANES streams with graphs + PDL)
    prof = eng.profile(0, 3)
    tc, tc_ms = conv_family(prof)
    step_ms_prof = sum(p["ms"] for p in prof)
    algo = ALGO[SCALE]
    tflops = algo["flops"] * BATCH / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    gbs = algo["act_bytes"] * BATCH / (tc_ms * 1e-3) / 1e9 if tc_ms > 0 else 0.0
    ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)
    ai = algo["flops"] / algo["act_bytes"]
    hbm_bound = ai < ridge
    traffic = None
    try:   # ncu dram__bytes_{read,write}.sum of the conv family per step (profiles/)
        for name in ("step_metrics_summary_r02.json", "step_metrics_summary_r01.json"):
            pth = os.path.join(ROOT, "profiles", name)
            if os.path.exists(pth):
                sm = json.load(open(pth))
                traffic = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in sm.items()
                              if k.startswith("conv_halo") or k.startswith("conv_tc") or k.startswith("head_decode"))
                break
    except Exception:
        traffic = None
    roofline = {
        "kernel": "conv_halo_kernel (persistent tcgen05 implicit-GEMM conv) + head_decode_kernel (last 1x1 convs of the head + decode), %d launches/step" % len(tc),
        "bound": "hbm" if hbm_bound else "tensor",
        "achieved": gbs if hbm_bound else tflops,
        "peak": peaks["hbm"] if hbm_bound else peaks["tf_sust"],
        "unit": "GB/s" if hbm_bound else "TFLOP/s",
        "frac": (gbs / peaks["hbm"]) if hbm_bound else (tflops / peaks["tf_sust"]),
        "traffic": traffic,
        "algorithmic_bytes_per_step": algo["act_bytes"] * BATCH,
        "algorithmic_note": "SURVEY.md 8(d): 60.4 MB of 16-bit activations per frame (every conv output written once and read once) x 64 frames; "
                            "8.743 GFLOP per frame",
        "unit_of_work": "one step = all %d tcgen05 conv launches of a 64-frame batch (54 x conv_halo_kernel + the fused head kernel holding the last six 1x1 convs)" % len(tc),
        "peak_source": peaks["src"] + (", sustained figure: kernel timed inside a long step" if not hbm_bound else ""),
        "tensor_tflops": tflops, "tensor_frac_of_sustained": tflops / peaks["tf_sust"],
        "hbm_gbs_algorithmic": gbs, "hbm_frac": gbs / peaks["hbm"],
        "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
        "conv_ms_per_step_serial": tc_ms, "all_kernels_ms_per_step_serial": step_ms_prof,
        "share_of_step": tc_ms / step_ms_prof if step_ms_prof else None,
        "how": "zl_engine_profile: CUDA-event pair around every kernel on ONE un-captured stream, 3 passes after the timed region; serial, "
               "so the sum exceeds ms_per_step (timed region = %d streams overlapping, CUDA graphs + programmatic dependent launch)" % LANES,
        "whole_step_hbm_frac": value / world * algo["act_bytes"] / 1e9 / peaks["hbm"],
        "whole_step_tensor_frac_of_sustained": value / world * algo["flops"] / (peaks["tf_sust"] * 1e12),
    }
    top = sorted(prof, key=lambda p: -p["ms"])[:8]
    roofline["top_kernels_ms"] = [{"name": p["name"], "ms": round(p["ms"], 4)} for p in top]
    # P1 stand-alone: achieved GB/s against the HBM peak
    pre_ms, pre_bytes = eng.bench_preprocess(HW, HW, BATCH, 20)
    roofline["preprocess"] = {"kernel": "preprocess_kernel stand-alone, 64 frames 640x640 -> NHWC4 16-bit", "ms": pre_ms,
                              "gbs": pre_bytes / (pre_ms * 1e-3) / 1e9, "hbm_frac": pre_bytes / (pre_ms * 1e-3) / 1e9 / peaks["hbm"],
                              "bytes_per_launch": pre_bytes}

    # ---- steady state: >= 5 s of back-to-back steps, clocks and power sampled over the whole window
    steady = None
    if not args.quick:
        n_long = int(max(5.5e3 / (max_ms / args.steps), 200))
        smp = ClockSampler(local_rank)
        smp.start()
        ms_long, _, _ = eng.run_resident(n_sets, n_long)
        ck = smp.stop()
        steady = {"seconds": ms_long / 1e3, "steps": n_long, "frames_per_s": BATCH * n_long / (ms_long / 1e3), "clocks": ck}

    # ---- the other 16-bit format (same kernels, the UMMA instruction descriptor differs): BASELINE configs[2] says bf16
    other = None
    if not args.quick:
        od = "bf16" if args.dtype == "fp16" else "fp16"
        try:
            eng.close()
            eng = zlb200.Engine(HW, HW, NC, SCALE, precision=PREC[od], conf=CONF, iou=IOU, max_batch=BATCH, device=local_rank, num_lanes=LANES)
            eng.load_weights_blob(blob)
            eng.warmup(1)
            for s in range(n_sets):
                eng.upload_resident(s, list(sets[s]))
            oms, _ = measure_resident(eng, n_sets, args.steps, args.warmup)
            other = {"dtype": od, "value": BATCH * args.steps / (oms / 1e3), "unit": UNIT, "ms_per_step": oms / args.steps, "n_gpus": 1, "gate": GATE[od]}
        except Exception as ex:
            other = {"dtype": od, "error": str(ex)}

    # ---- cfg5 (BASELINE configs[4]): decode/NMS stress, 8400 anchors x 80 classes x batch 128, conf 0.01
    cfg5 = None
    if not args.quick:
        try:
            from oracle import synth
            raw = synth.stress_head(128, 80, 8400)
            mf, mn, kept = eng.bench_decode_nms(raw, 0.01, 0.45, iters=5)
            b = 128 * 2.822e6
            cfg5 = {"workload": "head tensor [128, 84, 8400] fp32, conf 0.01, iou 0.45 (SURVEY 8d stress set)", "filter_ms": mf, "nms_ms": mn, "kept": kept,
                    "filter_gbs_vs_2.822MB_per_frame": b / (mf * 1e-3) / 1e9, "filter_hbm_frac": b / (mf * 1e-3) / 1e9 / peaks["hbm"],
                    "decode_plus_nms_frames_per_s": 128 / ((mf + mn) * 1e-3)}
            # the same call at a small batch: launch_nms deals each frame's classes to a thread-block cluster (4 CTAs at 32 frames)
            mf32, mn32, kept32 = eng.bench_decode_nms(raw[:32], 0.01, 0.45, iters=5)
            cfg5["batch32_cluster_of_4"] = {"filter_ms": mf32, "nms_ms": mn32, "kept": kept32}
        except Exception as ex:
            cfg5 = {"error": str(ex)}

    # ---- b=1 416x416 latency (BASELINE configs[1]), CUDA graph, frame in pinned host memory
    latency = None
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        from oracle import synth, yolov8_ref, zlw
        t4 = yolov8_ref.synthetic_model("n", 4, seed=0)
        e1 = zlb200.Engine(416, 416, 4, "n", precision=prec, max_batch=1, device=local_rank)
        e1.load_weights_blob(zlw.dumps(t4, "n", 4))
        e1.warmup(3)
        pf = zlb200.pinned_array((416, 416, 3))
        pf[:] = synth.frames_structured(1, 416, 416)[0]
        lat = np.sort(e1.bench_latency(pf, warmup=200, iters=2000))
        latency = {"workload": "YOLOv8n 416x416 b=1 nc=4, CUDA graph, pinned host frame -> detections on host",
                   "p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]), "mean_ms": float(lat.mean()),
                   "device_ms": e1.stats()["avg_device_time_ms"], "iters": 2000}
        e1.close()
    except Exception as ex:  # the headline number must still print
        latency = {"error": str(ex)}

    # ---- CPU baseline on the box's host cores (N=1 only): bounded sample of the same workload
    cpu_baseline = None
    if world == 1 and not args.quick:
        cores = os.cpu_count() or 1
        sample = list(sets[0][:16])
        cpu_pipeline_fps(tensors, sample[:2], cores, 1)
        fps1, dt1, _ = cpu_pipeline_fps(tensors, sample, cores, 1)
        fps8, dt8, _ = cpu_pipeline_fps(tensors, sample, cores, 8)
        cpu_baseline = {"value": fps1, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"16 frames b=1 sequential ({dt1:.1f}s), as the reference runs; batched b=8: {fps8:.1f} frames/s ({dt8:.1f}s)",
                        "note": "torch-CPU fp32 stand-in for the ORT-CPU session + the oracle's C pre/post (pinned to the reference's own compiled text, oracle/_ref)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD.replace("bf16", args.dtype), "inputs": f"{n_sets} rotating resident input sets of {BATCH} frames ({n_sets * BATCH * HW * HW * 3 / 1e6:.0f} MB > 126 MB L2)",
                   "frames_per_step_per_gpu": BATCH, "streams": LANES, "parallelism": f"frame-sharded replicas x{world}, no collective"},
        "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": BATCH * HW * HW * 3, "d2h_bytes_per_step": d2h,
                "api": f"zl_infer_batch (C-ABI) on pinned host frames, looped in C (zl_bench_e2e) from {LANES} host threads (one lane each)", "detections_last_step": int(n_det),
                "h2d_gbs_achieved_per_gpu": e2e_fps / world * HW * HW * 3 / 1e9, "h2d_gbs_pinned_peak_per_gpu_all_ranks_busy": h2d_peak_min},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "latency_b1_416": latency,
        "gate_16bit": {"fp16": GATE["fp16"], "bf16": GATE["bf16"], "bf16_gate": "fail", "fp16_gate": "pass"},
        "other_dtype": other,
        "steady_state": steady,
        "cfg4_yolov8s_m_b256": cfg4,
        "cfg5_decode_nms": cfg5,
        "tensor_roof_frac_whole_step": value * FLOPS_PER_FRAME / (world * peaks["tf_sust"] * 1e12),
        "wall_ms_per_step_rank0": wall_ms / args.steps,
    }
    print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"], help="16-bit tensor-core format (same speed; fp16 meets the IoU>=0.99 parity gate)")
    ap.add_argument("--e2e-threads", type=int, default=4, help="host threads (= engine lanes) feeding zl_infer_batch in the e2e leg")
    ap.add_argument("--quick", action="store_true", help="skip the latency and CPU-baseline legs (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
