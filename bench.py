#!/usr/bin/env python
"""bench.py -- headline measurement of the QOI hot path (BASELINE.json: raw-pixel GB/s encode & decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = encode + decode of the workload's synthetic images through the device-pointer C ABI
(include/qoipp_b200.h), inputs already resident in HBM, L2 flushed between the timed regions.
`value` = raw pixel bytes through the codec per second (encode pass + decode pass, each counting the raw image
once), whole job over all ranks.  `e2e` = the same metric through qoipp_b200_encode_host / _decode_host with
pinned HOST buffers (H2D and D2H inside the timed region).  See DESIGN.md "Measurement".

Workloads (BASELINE.json configs):
  4k_rgb_photo    configs[1]  single 3840x2160 RGB photo-like image (default; the config the metric is quoted on)
  8k_rgba_photo   the "single 8K image" of the target sentence, 7680x4320 RGBA, opaque
  16k_rgba_noise  configs[2]  16384x16384 RGBA noise (decode stress)
  batch512        configs[3]  batch of 512x512 RGBA photo-like images, sharded by image across ranks
Under torchrun every rank runs the workload on its own images (sharded by image, no collective on the data path).
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

WORKLOADS = {
    # name: (kind, w, h, ch, images per rank, distinct images generated)
    "4k_rgb_photo": ("photo", 3840, 2160, 3, 1, 1),
    "8k_rgba_photo": ("photo_opaque", 7680, 4320, 4, 1, 1),
    "16k_rgba_noise": ("noise", 16384, 16384, 4, 1, 1),
    "batch512": ("photo_opaque", 512, 512, 4, 1024, 64),
}


def make_images(name: str, rank: int):
    from qoipp_b200 import synth

    kind, w, h, ch, n, distinct = WORKLOADS[name]
    imgs = []
    for k in range(distinct):
        seed = 0x51F0 + 1000 * rank + k
        if kind == "photo_opaque":  # RGBA photo-like with alpha 255 (the blob variant is a parity-test class)
            rgb = synth.generate("photo", w, h, 3, seed=seed).reshape(-1, 3)
            img = np.concatenate([rgb, np.full((rgb.shape[0], 1), 255, np.uint8)], axis=1).reshape(-1)
        elif kind == "noise" and w * h > (1 << 26):  # 1 GiB: tile a 64 Mi-pixel noise block (content class is what matters)
            block = synth.generate("noise", 8192, 8192, ch, seed=seed)
            img = np.tile(block, (w * h) // (8192 * 8192))
        else:
            img = synth.generate(kind, w, h, ch, seed=seed)
        imgs.append(img)
    return kind, w, h, ch, n, imgs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(name: str, imgs, w, h, ch, budget_s: float = 20.0):
    """The reference's own CPU path (oracle/_ref when it was built, else the C port) on the host cores:
    single thread and thread-per-image with every core.  Returns the cpu_baseline object + raw GB/s."""
    from oracle.pyoracle import Oracle, Ref

    impl, kind = (Ref, "reference") if Ref.available() else (Oracle, "port")
    raw = imgs[0]
    # 16k images are sampled by their first 4096 rows: same content class, bounded time
    rows = h if w * h <= (1 << 25) else max(1, (1 << 25) // w)
    raw = raw[: w * rows * ch]
    cap = (ch + 1) * w * rows + 22
    enc = impl.encode(raw, w, rows, ch)

    def one():  # encode_into a pre-allocated worst-size buffer + decode, like 04_bench.cpp:445-510
        e, out, n, ok = impl.encode_into(raw, w, rows, ch, 0, cap=cap)
        impl.decode(enc)

    one()
    t0 = time.perf_counter()
    one()
    t_one = time.perf_counter() - t0
    reps1 = max(1, min(10, int(budget_s / 3 / t_one)))
    t0 = time.perf_counter()
    for _ in range(reps1):
        one()
    t_single = (time.perf_counter() - t0) / reps1
    cores = os.cpu_count() or 1
    repsT = max(1, min(4, int(budget_s / 2 / t_one)))

    def worker():
        for _ in range(repsT):
            one()

    th = [threading.Thread(target=worker) for _ in range(cores)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    t_multi = time.perf_counter() - t0
    gbps_single = 2 * raw.size / t_single / 1e9
    gbps_multi = 2 * raw.size * cores * repsT / t_multi / 1e9
    return {
        "value": round(gbps_multi, 4), "unit": "GB/s", "cores": cores, "kind": kind,
        "sample": f"{name}: {w}x{rows}x{ch} encode_into+decode, thread-per-image on {cores} threads x {repsT} reps "
                  f"(single thread: {gbps_single:.4f} GB/s, {reps1} reps)",
        "single_thread_value": round(gbps_single, 4), "sample_raw_bytes": int(raw.size),
    }


def run_reference_arm(args, rank):
    if rank != 0:
        return
    kind, w, h, ch, n, imgs = make_images(args.workload, 0)
    for _ in range(max(0, args.warmup - 1)):
        pass  # cpu_reference() does its own untimed call
    cb = cpu_reference(args.workload, imgs, w, h, ch, budget_s=min(60.0, 8.0 * max(1, args.steps)))
    line = {
        "impl": "reference", "metric": "raw_pixel_GBps_encode_decode", "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(2 * cb["sample_raw_bytes"] / (cb["value"] * 1e9) * 1e3, 5),
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "width": w, "height": h, "channels": ch, "images_per_rank": n},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="4k_rgb_photo", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    from qoipp_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: qoipp_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    kind, w, h, ch, n_img, imgs = make_images(args.workload, rank)
    raw_one = w * h * ch
    worst = (ch + 1) * w * h + 22
    stride = (worst + 255) // 256 * 256
    ctx = api.Context(local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- inputs resident in HBM
    host_raw = np.concatenate([imgs[k % len(imgs)] for k in range(n_img)])
    d_raw = torch.from_numpy(host_raw).cuda()
    d_qoi = torch.empty(stride * n_img, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(raw_one * n_img, dtype=torch.uint8, device="cuda")
    d_written = torch.zeros(n_img, dtype=torch.int64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    batch = n_img > 1

    def encode():
        if batch:
            ctx.encode_batch_dev(d_raw, raw_one, n_img, w, h, ch, 0, d_qoi, stride, stride, d_written, stream)
        else:
            ctx.encode_dev(d_raw, w, h, ch, 0, d_qoi, worst, stream)

    encode()
    torch.cuda.synchronize()
    if batch:
        sizes = d_written.cpu().numpy().astype(np.uint64)
        # decode takes the streams packed back to back: gather them once (setup, untimed)
        offs = np.zeros(n_img + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(sizes)
        packed = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device="cuda")
        for k in range(n_img):
            packed[int(offs[k]): int(offs[k + 1])] = d_qoi[k * stride: k * stride + int(sizes[k])]
        enc_bytes = int(offs[-1])
    else:
        enc_bytes, ok = ctx.encode_status(stream)
        assert ok
        offs, packed = None, d_qoi

    def decode():
        if batch:
            ctx.decode_batch_dev(packed, offs, w, h, ch, 0, 0, d_out, raw_one, stream)
        else:
            ctx.decode_dev(packed, enc_bytes, w, h, ch, 0, 0, False, d_out, raw_one, stream)

    decode()
    torch.cuda.synchronize()
    path = ctx.decode_status(stream)
    assert torch.equal(d_out, d_raw), "round trip mismatch"  # encode -> decode returns the input bit for bit

    def step(ev):
        flush.fill_(1)
        ev[0].record()
        encode()
        ev[1].record()
        flush.fill_(2)
        ev[2].record()
        decode()
        ev[3].record()

    sampler = ClockSampler(local_rank)  # nvidia-smi needs ~100 ms to produce its first line: start before the warm-up
    sampler.start()
    for _ in range(args.warmup):
        step([torch.cuda.Event(enable_timing=True) for _ in range(4)])
    torch.cuda.synchronize()
    # keep the GPU under the same load until the sampler has delivered a few lines, then time
    t_load = time.perf_counter()
    while len(sampler.lines) < 3 and time.perf_counter() - t_load < 1.5:
        step([torch.cuda.Event(enable_timing=True) for _ in range(4)])
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for ev in evs:
        step(ev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    enc_ms = float(np.sum([ev[0].elapsed_time(ev[1]) for ev in evs]))
    dec_ms = float(np.sum([ev[2].elapsed_time(ev[3]) for ev in evs]))
    tot = torch.tensor([enc_ms + dec_ms, enc_ms, dec_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)  # slowest rank defines the job time
    tot_ms, enc_ms_max, dec_ms_max = (float(x) for x in tot.cpu())

    # ---- end to end through the host-pointer C ABI, pinned buffers, one image at a time
    e2e_steps = max(3, min(args.steps, 10))
    h_raw = torch.from_numpy(imgs[0]).pin_memory()
    h_qoi = torch.empty(worst, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(raw_one, dtype=torch.uint8).pin_memory()
    np_raw, np_qoi, np_out = h_raw.numpy(), h_qoi.numpy(), h_out.numpy()
    import ctypes as C
    from qoipp_b200._lib import Desc, lib

    def e2e_once():
        written, complete = C.c_uint64(0), C.c_int32(0)
        e = lib.qoipp_b200_encode_host(ctx._h, C.c_void_p(np_raw.ctypes.data), np_raw.size, C.byref(Desc(w, h, ch, 0)),
                                       C.c_void_p(np_qoi.ctypes.data), worst, C.byref(written), C.byref(complete))
        assert e == 0 and complete.value
        d = Desc()
        e = lib.qoipp_b200_decode_host(ctx._h, C.c_void_p(np_qoi.ctypes.data), written.value, 0, 0, C.c_void_p(np_out.ctypes.data),
                                       raw_one, C.byref(d))
        assert e == 0
        return written.value

    e2e_imgs = min(n_img, 16)
    for _ in range(2):
        e2e_once()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for _ in range(e2e_imgs):
            e2e_enc = e2e_once()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert np.array_equal(np_out, imgs[0])
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_gbps = 2 * raw_one * e2e_imgs * world / float(e2e_t.item()) / 1e9

    if rank == 0:
        raw_total = raw_one * n_img
        ms_per_step = tot_ms / args.steps
        value = 2 * raw_total * world / (ms_per_step * 1e-3) / 1e9
        enc_gbps = raw_total / (enc_ms_max / args.steps * 1e-3) / 1e9
        dec_gbps = raw_total / (dec_ms_max / args.steps * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # dominant kernel = the slower direction; algorithmic bytes = raw + encoded (each byte crosses HBM once)
        alg = raw_total + enc_bytes
        dom = "decode_kernel" if dec_ms_max >= enc_ms_max else "encode_kernel"
        dom_ms = max(dec_ms_max, enc_ms_max) / args.steps
        achieved = alg / (dom_ms * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel, per launch, from the committed ncu --set full capture of this workload
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tr.get("workload") == args.workload and dom in tr["kernels"]:
                traffic = int(tr["kernels"][dom]["dram_bytes_read"] + tr["kernels"][dom]["dram_bytes_write"])
        except Exception:
            traffic = None
        line = {
            "metric": "raw_pixel_GBps_encode_decode", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "class": kind, "width": w, "height": h, "channels": ch, "images_per_rank": n_img,
                       "raw_bytes_per_rank": raw_total, "encoded_bytes_per_rank": enc_bytes, "l2": "flushed (512 MiB fill) before every timed region",
                       "decode_path": "parallel" if path == 0 else "sequential", "sharding": "by image, no collective"},
            "encode_GBps": round(enc_gbps, 3), "decode_GBps": round(dec_gbps, 3),
            "encode_ms": round(enc_ms_max / args.steps, 5), "decode_ms": round(dec_ms_max / args.steps, 5),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes": alg, "encode_frac": round((alg / (enc_ms_max / args.steps * 1e-3) / 1e9) / peak, 5),
                         "decode_frac": round((alg / (dec_ms_max / args.steps * 1e-3) / 1e9) / peak, 5)},
            "e2e": {"value": round(e2e_gbps, 3), "unit": "GB/s", "h2d_bytes_per_step": (raw_one + e2e_enc) * e2e_imgs,
                    "d2h_bytes_per_step": (e2e_enc + raw_one) * e2e_imgs, "images_per_step": e2e_imgs},
            "gpu_launches": 4 * args.steps, "clocks": clocks, "wall_s": round(t_wall, 3),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args.workload, imgs, w, h, ch)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
