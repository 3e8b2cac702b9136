#!/usr/bin/env python
"""bench.py -- headline measurement of the QOI hot path (BASELINE.json: raw-pixel GB/s encode & decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--also LIST|none] [--impl reference]

One "step" = encode + decode of the workload's synthetic images through the device-pointer C ABI
(include/qoipp_b200.h), inputs already resident in HBM, L2 flushed between the timed regions.
`value` = raw pixel bytes through the codec per second (encode pass + decode pass, each counting the raw image
once), whole job over all ranks.  `e2e` = the same metric through the host-pointer C ABI with pinned HOST buffers
(H2D and D2H inside the timed region); `e2e_pageable` = through qoipp::encode / qoipp::decode of libqoipp.so on ordinary
host memory (the call a user of the reference makes).  See DESIGN.md "Measurement".

Workloads (BASELINE.json configs):
  8k_rgba_photo         the "single 8K image" of the target sentence: 7680x4320 RGBA `photo` (soft alpha blobs)   [default, 1 GPU]
  8k_rgba_photo_opaque  the same content with alpha 255
  4k_rgb_photo          configs[1]  single 3840x2160 RGB photo-like image
  16k_rgba_noise / 16k_rgba_photo / 16k_rgba_resync   configs[2]  16384x16384 RGBA decode stress
  batch8192             configs[3]  8192 x 512x512 RGBA `photo` (seed 0x51F0 + k), image k on rank floor(k * G / 8192)
                        (qoipp_b200.sharding.shard_range), strong scaling                                         [default under torchrun]
  batch8192_opaque      the same with alpha 255
The default 1-GPU run prints ONE line whose headline is 8k_rgba_photo and whose "also" object holds the other
workloads (fewer steps each) and whose cpu_baseline holds configs[0] (1080p RGBA round trip on the CPU).
Single-image workloads under torchrun run one replica per rank (a single image stays on one GPU).
"""
from __future__ import annotations

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

WORKLOADS = {
    # name: kind, w, h, ch, images (whole job), BASELINE config
    "8k_rgba_photo": ("photo", 7680, 4320, 4, 1, "target sentence: single 8K image"),
    "8k_rgba_photo_opaque": ("photo_opaque", 7680, 4320, 4, 1, "target sentence, alpha 255"),
    "4k_rgb_photo": ("photo", 3840, 2160, 3, 1, "configs[1]"),
    "16k_rgba_noise": ("noise", 16384, 16384, 4, 1, "configs[2]"),
    "16k_rgba_photo": ("photo", 16384, 16384, 4, 1, "configs[2]"),
    "16k_rgba_resync": ("resync", 16384, 16384, 4, 1, "configs[2]"),
    "batch8192": ("photo", 512, 512, 4, 8192, "configs[3]"),
    "batch8192_opaque": ("photo_opaque", 512, 512, 4, 8192, "configs[3], alpha 255"),
}
ALSO_DEFAULT = ["4k_rgb_photo", "8k_rgba_photo_opaque", "16k_rgba_noise", "16k_rgba_photo", "16k_rgba_resync", "batch8192",
                "batch8192_opaque"]
METRIC = "raw_pixel_GBps_encode_decode"


def workload_config(name: str, world: int) -> dict:
    """The `config` object -- identical in the b200 and the reference arm."""
    kind, w, h, ch, images, cfg = WORKLOADS[name]
    return {"workload": name, "baseline_config": cfg, "class": kind, "width": w, "height": h, "channels": ch, "images": images,
            "sharding": f"image k on rank floor(k*{world}/{images}), no collective" if images > 1 else "one replica per rank",
            "l2": "flushed (512 MiB fill) before every timed region"}


def shard(name: str, rank: int, world: int):
    """-> (first, last) image indices of this rank and the seeds of those images"""
    from qoipp_b200 import sharding, synth

    kind, w, h, ch, images, _ = WORKLOADS[name]
    if images == 1:
        base = "photo" if kind == "photo_opaque" else kind
        return 0, 1, [synth.BASE_SEED + synth.CLASSES.index(base) + 1000 * rank]
    first, last = sharding.shard_range(images, rank, world)
    return first, last, [synth.BASE_SEED + k for k in range(first, last)]


def generate_device(name: str, seeds, device):
    """[len(seeds), raw_one] uint8 on `device`, bit-identical to qoipp_b200.synth (tests/test_synth_torch.py)"""
    import torch

    from qoipp_b200 import synth_torch

    kind, w, h, ch, _, _ = WORKLOADS[name]
    out = torch.empty((len(seeds), w * h * ch), dtype=torch.uint8, device=device)
    step = max(1, (1 << 26) // (w * h))  # images per generator call
    for i in range(0, len(seeds), step):
        out[i: i + step] = synth_torch.generate(kind, w, h, ch, seeds=seeds[i: i + step], device=device)
    return out


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


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) timed by its own benchmark method (oracle/ref_shim.cpp ref_bench =
# example/source/04_bench.cpp:445-510,733-754 / BASELINE.md section 4)
# ----------------------------------------------------------------------------------------------------------------------
def host_sample(name: str, n_images: int, max_rows: int | None):
    """`n_images` host images of the workload's class, at most `max_rows` rows each (a bounded sample of a large image
    keeps the content class: the first rows of the same image)."""
    import torch

    kind, w, h, ch, images, _ = WORKLOADS[name]
    rows = h if max_rows is None else min(h, max_rows)
    _, _, seeds0 = shard(name, 0, 1)
    seeds = [seeds0[k % len(seeds0)] if images > 1 else seeds0[0] + 7919 * k for k in range(n_images)]
    from qoipp_b200 import synth_torch

    dev = "cuda" if torch.cuda.is_available() else "cpu"
    out = []  # rows [0, rows) of an image are a function of (seed, w) alone: generate exactly those
    for i in range(0, n_images, 64):
        st = torch.as_tensor([synth_torch._s64(int(s)) for s in seeds[i: i + 64]], dtype=torch.int64)
        if kind in ("photo", "photo_opaque"):
            t = synth_torch._photo_rows(w, 0, rows, ch, st, dev, opaque=kind == "photo_opaque")
        else:
            t = synth_torch._noise_rows(w, 0, rows, ch, st, dev, resync=kind == "resync")
        out.extend(np.ascontiguousarray(x) for x in t.cpu().numpy())
    return out, w, rows, ch


def cpu_reference(name: str, reps: int = 5, single: bool = True):
    """cpu_baseline object for `name`: thread-per-image on every host thread (+ one thread), bounded sample."""
    from oracle.pyoracle import Oracle, Ref

    kind, w, h, ch, images, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    max_rows = None if w * h <= (1 << 22) else max(1, (1 << 23) // w)  # <= 8 Mi pixels per image
    per_thread = 4 if images > 1 else 1
    imgs, w, rows, ch = host_sample(name, cores * per_thread, max_rows)
    raw_one = imgs[0].size
    sample = f"{name}: {len(imgs)} images of {w}x{rows}x{ch}" + (f" (first {rows} rows of the {h}-row image)" if rows != h else "")
    if Ref.available():
        r = Ref.bench(imgs, w, rows, ch, threads=0, warmups=3, reps=reps)
        enc_s, dec_s, T = r["enc_s"], r["dec_s"], r["threads"]
        tot = raw_one * len(imgs) * reps
        out = {"value": round(2 * tot / (enc_s + dec_s) / 1e9, 4), "unit": "GB/s", "cores": T, "kind": "reference",
               "encode_GBps": round(tot / enc_s / 1e9, 4), "decode_GBps": round(tot / dec_s / 1e9, 4),
               "sample": sample + f"; qoipp::encode_into (pre-allocated, pre-touched worst_size buffer) + qoipp::decode_into, thread-per-image on "
                                  f"T={T} threads (hardware_concurrency), 1 + 3 warm-ups, {reps} timed passes, steady_clock; reference built "
                                  f"-O3 -march={r['march']} (built in the dev container: /root/reference is not on this box, so not -march=native)",
               "encoded_over_raw": round(r["enc_bytes"] / (raw_one * len(imgs)), 4), "sample_raw_bytes": int(raw_one * len(imgs))}
        if single:
            r1 = Ref.bench(imgs[:1], w, rows, ch, threads=1, warmups=3, reps=reps)
            out["single_thread"] = {"encode_GBps": round(raw_one * reps / r1["enc_s"] / 1e9, 4), "decode_GBps": round(raw_one * reps / r1["dec_s"] / 1e9, 4),
                                    "value": round(2 * raw_one * reps / (r1["enc_s"] + r1["dec_s"]) / 1e9, 4)}
        return out
    # the C restatement, single thread (no compiled reference on this box)
    t0 = time.perf_counter()
    for im in imgs[:2]:
        Oracle.decode(Oracle.encode(im, w, rows, ch))
    dt = time.perf_counter() - t0
    return {"value": round(2 * raw_one * 2 / dt / 1e9, 4), "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample + "; oracle port, 1 thread"}


def cpu_config1():
    """configs[0]: 1920x1080 RGBA round trip on the CPU, classes noise / gradient / flat / photo."""
    from oracle.pyoracle import Ref
    from qoipp_b200 import synth

    if not Ref.available():
        return None
    out = {}
    for cls in ("noise", "gradient", "flat", "photo"):
        img = synth.generate(cls, 1920, 1080, 4)
        r = Ref.bench([img], 1920, 1080, 4, threads=1, warmups=3, reps=5)
        out[cls] = {"encode_ms": round(r["enc_s"] / 5 * 1e3, 3), "decode_ms": round(r["dec_s"] / 5 * 1e3, 3),
                    "encode_GBps": round(img.size * 5 / r["enc_s"] / 1e9, 4), "decode_GBps": round(img.size * 5 / r["dec_s"] / 1e9, 4),
                    "encoded_over_raw": round(r["enc_bytes"] / img.size, 4)}
    return {"config": "configs[0]: single 1920x1080 RGBA image, encode + decode on the CPU, 1 thread, 3 warm-ups + mean of 5", "classes": out}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    name = args.workload
    reps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    cb = cpu_reference(name, reps=reps, single=False)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": reps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(2 * cb["sample_raw_bytes"] / (cb["value"] * 1e9) * 1e3, 5) if cb.get("sample_raw_bytes") else None,
        "higher_is_better": True, "scaling": "strong" if WORKLOADS[name][4] > 1 else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(name, world),
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 2),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------------
def numa_pin(local_rank: int):
    """Run this rank's host threads on the cores next to its GPU, so that pinned buffers (first touch) and the staging
    threads are NUMA-local to the PCIe root of the device.  Returns a description for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)], capture_output=True, text=True,
                             timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"numa_node": None}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        ids = [i for i in ids if i in allowed]
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus}
    except Exception as ex:  # best effort: the measurement still runs
        return {"numa_node": None, "note": str(ex)[:80]}


def measure(name, ctx, rank, local_rank, world, steps, warmup, dist, do_e2e=True, sampler=None):
    """One workload on this rank's shard.  Returns a dict of job-wide numbers (identical on every rank)."""
    import torch

    from qoipp_b200._lib import Desc, lib

    kind, w, h, ch, images, _ = WORKLOADS[name]
    first, last, seeds = shard(name, rank, world)
    n_img = last - first
    raw_one = w * h * ch
    worst = (ch + 1) * w * h + 22
    stride = (worst + 255) // 256 * 256
    stream = torch.cuda.current_stream().cuda_stream
    dev = torch.device("cuda", local_rank)
    batch = images > 1

    d_raw = generate_device(name, seeds, dev).reshape(-1)
    d_qoi = torch.empty(stride * n_img, dtype=torch.uint8, device=dev)
    d_out = torch.empty(raw_one * n_img, dtype=torch.uint8, device=dev)
    d_written = torch.zeros(n_img, dtype=torch.int64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def encode():
        if batch:
            ctx.encode_batch_dev(d_raw, raw_one, n_img, w, h, ch, 0, d_qoi, stride, stride, d_written, stream)
        else:
            ctx.encode_dev(d_raw, w, h, ch, 0, d_qoi, worst, stream)

    encode()
    torch.cuda.synchronize()
    if batch:
        sizes = d_written.cpu().numpy().astype(np.uint64)
        offs = np.zeros(n_img + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(sizes)
        packed = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device=dev)  # decode takes the streams back to back (setup, untimed)
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
    paths = ctx.decode_status_batch(n_img, stream) if batch else np.array([ctx.decode_status(stream)])
    assert torch.equal(d_out, d_raw), f"{name}: round trip mismatch"  # encode -> decode returns the input bit for bit
    d_out.zero_()

    def step(ev):
        flush.fill_(1)
        ev[0].record()
        encode()
        ev[1].record()
        flush.fill_(2)
        ev[2].record()
        decode()
        ev[3].record()

    def events():
        return [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    for _ in range(warmup):
        step(events())
    torch.cuda.synchronize()
    if sampler is not None:  # keep the GPU under the same load until nvidia-smi has delivered a few lines
        t_load = time.perf_counter()
        while len(sampler.lines) < 3 and time.perf_counter() - t_load < 1.5:
            step(events())
            torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [events() for _ in range(steps)]
    t_wall0 = time.perf_counter()
    for ev in evs:
        step(ev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    assert torch.equal(d_out, d_raw), f"{name}: round trip mismatch after the timed steps"
    enc_ms = float(np.sum([ev[0].elapsed_time(ev[1]) for ev in evs]))
    dec_ms = float(np.sum([ev[2].elapsed_time(ev[3]) for ev in evs]))
    red = torch.tensor([enc_ms + dec_ms, enc_ms, dec_ms, float(paths.max())], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(raw_one * n_img), float(enc_bytes), float((paths > 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)  # the slowest rank defines the job time
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tot_ms, enc_ms_max, dec_ms_max, path_max = (float(x) for x in red.cpu())
    raw_job, enc_job, retried = (int(x) for x in tot.cpu())
    res = {
        "images_this_rank": n_img, "raw_bytes": raw_job, "encoded_bytes": enc_job, "encoded_over_raw": round(enc_job / raw_job, 4),
        "ms_per_step": tot_ms / steps, "encode_ms": enc_ms_max / steps, "decode_ms": dec_ms_max / steps,
        "decode_path_max": int(path_max), "images_needing_retry_rounds": retried, "wall_s": round(t_wall, 3), "steps": steps,
    }
    del flush, d_out

    # ---- end to end through the host-pointer C ABI with pinned buffers (H2D + D2H inside the timed region)
    if do_e2e:
        e2e_imgs = min(n_img, 1024)
        e2e_steps = max(3, min(steps, 10 if batch else 20))
        h_raw = torch.empty(raw_one * e2e_imgs, dtype=torch.uint8).pin_memory()
        h_raw.copy_(d_raw[: raw_one * e2e_imgs])
        h_qoi = torch.empty(stride * e2e_imgs, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(raw_one * e2e_imgs, dtype=torch.uint8).pin_memory()
        np_raw, np_qoi, np_out = h_raw.numpy(), h_qoi.numpy(), h_out.numpy()
        h_written = np.zeros(e2e_imgs, dtype=np.uint64)
        desc = Desc(w, h, ch, 0)
        state = {"enc": 0}

        def e2e_once():
            if batch:
                e = lib.qoipp_b200_encode_batch_host(ctx._h, C.c_void_p(np_raw.ctypes.data), raw_one, e2e_imgs, C.byref(desc), C.c_void_p(np_qoi.ctypes.data),
                                                     stride, stride, h_written.ctypes.data_as(C.POINTER(C.c_uint64)))
                assert e == 0, e
                state["enc"] = int(h_written.sum())
                # the encoded slots are decoded where they lie (stream k at k * stride, h_written[k] bytes)
                e = lib.qoipp_b200_decode_batch_host(ctx._h, C.c_void_p(np_qoi.ctypes.data), stride, h_written.ctypes.data_as(C.POINTER(C.c_uint64)), e2e_imgs,
                                                     C.byref(desc), 0, C.c_void_p(np_out.ctypes.data), raw_one)
                assert e == 0, e
            else:
                written, complete = C.c_uint64(0), C.c_int32(0)
                e = lib.qoipp_b200_encode_host(ctx._h, C.c_void_p(np_raw.ctypes.data), np_raw.size, C.byref(desc), C.c_void_p(np_qoi.ctypes.data), worst,
                                               C.byref(written), C.byref(complete))
                assert e == 0 and complete.value
                d = Desc()
                e = lib.qoipp_b200_decode_host(ctx._h, C.c_void_p(np_qoi.ctypes.data), written.value, 0, 0, C.c_void_p(np_out.ctypes.data), raw_one,
                                               C.byref(d))
                assert e == 0
                state["enc"] = written.value

        for _ in range(2):
            e2e_once()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_once()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert np.array_equal(np_out, np_raw), f"{name}: e2e round trip mismatch"
        e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        cnt = torch.tensor([float(e2e_imgs), float(state["enc"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        imgs_job, enc_e2e = int(cnt[0].item()), int(cnt[1].item())
        res["e2e"] = {"value": round(2 * raw_one * imgs_job / float(e2e_t.item()) / 1e9, 3), "unit": "GB/s",
                      "h2d_bytes_per_step": raw_one * imgs_job + enc_e2e, "d2h_bytes_per_step": enc_e2e + raw_one * imgs_job,
                      "images_per_step": imgs_job, "steps": e2e_steps, "ms_per_step": round(float(e2e_t.item()) * 1e3, 4),
                      "path": "qoipp_b200_encode%s_host + qoipp_b200_decode%s_host on page-locked host buffers" % (("_batch",) * 2 if batch else ("",) * 2)}
        del h_raw, h_qoi, h_out

        # ---- the reference's own call on ordinary host memory: qoipp::encode / qoipp::decode of libqoipp.so
        so = os.path.join(ROOT, "qoipp_b200", "libqoipp_e2e.so")
        if os.path.exists(so):
            L = C.CDLL(so)
            L.qoipp_cxx_roundtrip.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint8, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
            img = d_raw[:raw_one].cpu().numpy()  # pageable
            es, ds, nb = C.c_double(0), C.c_double(0), C.c_uint64(0)
            reps = 5 if raw_one > (64 << 20) else 10
            if world > 1:
                dist.barrier()
            e = L.qoipp_cxx_roundtrip(C.c_void_p(img.ctypes.data), img.size, w, h, ch, local_rank, 2, reps, C.byref(es), C.byref(ds), C.byref(nb))
            assert e == 0, f"qoipp_cxx_roundtrip: {e}"
            tp = torch.tensor([(es.value + ds.value) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            res["e2e_pageable"] = {"value": round(2 * raw_one * world / float(tp.item()) / 1e9, 3), "unit": "GB/s", "images_per_step": world,
                                   "encode_ms": round(es.value / reps * 1e3, 3), "decode_ms": round(ds.value / reps * 1e3, 3),
                                   "path": "qoipp::encode(ByteCSpan, Desc) + qoipp::decode(ByteCSpan) of libqoipp.so, pageable std::vector, one image per call"}
    del d_raw, d_qoi, packed
    torch.cuda.empty_cache()
    return res


def roofline_of(name, res, peak, peak_src, world):
    """Dominant direction of the workload against the measured HBM peak; algorithmic bytes = raw + encoded per direction."""
    alg = res["raw_bytes"] + res["encoded_bytes"]  # whole job
    dom = "decode_wt_kernel" if res["decode_ms"] >= res["encode_ms"] else "encode_ts_kernel+encode_ts_copy_kernel"
    dom_ms = max(res["decode_ms"], res["encode_ms"])
    achieved = alg / world / (dom_ms * 1e-3) / 1e9  # per GPU: the peak is one GPU's
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        k = tr.get(name, {}).get("decode" if dom.startswith("decode") else "encode")
        if k:
            traffic = int(k["dram_bytes_read"] + k["dram_bytes_write"])
    except Exception:
        traffic = None
    return {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": alg // world,
            "encode_frac": round(alg / world / (res["encode_ms"] * 1e-3) / 1e9 / peak, 5),
            "decode_frac": round(alg / world / (res["decode_ms"] * 1e-3) / 1e9 / peak, 5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--also", default="auto", help="comma list of further workloads measured into the line's `also` object, or none")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    auto = args.workload == "auto"
    if auto:
        args.workload = "8k_rgba_photo" if max(world, args.gpus) == 1 else "batch8192"
    if args.also == "auto":
        also = ALSO_DEFAULT if (auto and world == 1) else []
    else:
        also = [a for a in args.also.split(",") if a and a != "none"]
    also = [a for a in also if a != args.workload]

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from qoipp_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: qoipp_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    numa = numa_pin(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the one JSON line
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    ctx = api.Context(local_rank)

    sampler = ClockSampler(local_rank)  # nvidia-smi needs ~100 ms to produce its first line: start before the warm-up
    sampler.start()
    res = measure(args.workload, ctx, rank, local_rank, world, args.steps, args.warmup, dist, do_e2e=not args.no_e2e, sampler=sampler)
    clocks = sampler.stop()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def summary(name, r):
        v = 2 * r["raw_bytes"] / (r["ms_per_step"] * 1e-3) / 1e9
        rf = roofline_of(name, r, peak, peak_src, world)
        out = {"config": WORKLOADS[name][5], "value": round(v, 3), "unit": "GB/s", "steps": r["steps"], "ms_per_step": round(r["ms_per_step"], 5),
               "encode_ms": round(r["encode_ms"], 5), "decode_ms": round(r["decode_ms"], 5),
               "encode_GBps": round(r["raw_bytes"] / (r["encode_ms"] * 1e-3) / 1e9, 3), "decode_GBps": round(r["raw_bytes"] / (r["decode_ms"] * 1e-3) / 1e9, 3),
               "encoded_over_raw": r["encoded_over_raw"], "encode_frac": rf["encode_frac"], "decode_frac": rf["decode_frac"], "traffic": rf["traffic"],
               "decode_path_max": r["decode_path_max"], "images_needing_retry_rounds": r["images_needing_retry_rounds"]}
        for k in ("e2e", "e2e_pageable"):
            if k in r:
                out[k] = r[k]
        return out

    also_out = {}
    for a in also:
        try:
            r = measure(a, ctx, rank, local_rank, world, max(3, min(args.steps, 5)), 3, dist, do_e2e=not args.no_e2e and a.startswith("batch"))
            also_out[a] = summary(a, r)
        except Exception as ex:  # a secondary workload never takes the headline down
            also_out[a] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
            torch.cuda.empty_cache()

    if rank == 0:
        images = WORKLOADS[args.workload][4]
        value = 2 * res["raw_bytes"] / (res["ms_per_step"] * 1e-3) / 1e9
        rf = roofline_of(args.workload, res, peak, peak_src, world)
        cfg = workload_config(args.workload, world)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(res["ms_per_step"], 5), "higher_is_better": True,
            "scaling": "strong" if images > 1 else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic (qoipp_b200/synth_torch.py, bit-identical to synth.py)",
            "config": cfg,
            "encode_GBps": round(res["raw_bytes"] / (res["encode_ms"] * 1e-3) / 1e9, 3),
            "decode_GBps": round(res["raw_bytes"] / (res["decode_ms"] * 1e-3) / 1e9, 3),
            "encode_ms": round(res["encode_ms"], 5), "decode_ms": round(res["decode_ms"], 5),
            "raw_bytes": res["raw_bytes"], "encoded_bytes": res["encoded_bytes"],
            "decode_path_max": res["decode_path_max"], "images_needing_retry_rounds": res["images_needing_retry_rounds"],
            "roofline": rf,
            "e2e": res.get("e2e"), "e2e_pageable": res.get("e2e_pageable"),
            "gpu_launches": 4 * args.steps, "clocks": clocks, "wall_s": res["wall_s"], "numa": numa,
        }
        if also_out:
            line["also"] = also_out
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args.workload)
            c1 = cpu_config1()
            if c1:
                line["cpu_baseline"]["config1_1080p_rgba"] = c1
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
