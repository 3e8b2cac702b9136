"""Development aid (not the bench contract): device-timed encode / decode of the bench images through the device-pointer
C ABI, CUDA events, L2 flushed between calls, median of N.  Used for the A/B numbers in profiles/r02_experiments.md.

    python tools/ab_probe.py [--reps 10] [--only decode|encode] [--cases 4k,8k,8kblob,...]
"""
import argparse
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qoipp_b200 import api, synth  # noqa: E402

CASES = {
    "4k": ("photo", 3840, 2160, 3, False),
    "8k": ("photo", 7680, 4320, 4, True),        # opaque
    "8kblob": ("photo", 7680, 4320, 4, False),   # alpha blobs (SURVEY's RGBA photo class)
    "4knoise": ("noise", 3840, 2160, 4, False),
    "8kflat": ("flat", 7680, 4320, 4, False),
    "8kgrad": ("gradient", 7680, 4320, 3, False),
    "4kpal": ("palette", 3840, 2160, 4, False),
    "4kdither": ("dither", 3840, 2160, 3, False),
    "512": ("photo", 512, 512, 4, True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--cases", default="4k,8k,8kblob")
    a = ap.parse_args()
    ctx = api.Context(0)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for name in a.cases.split(","):
        kind, w, h, ch, opaque = CASES[name]
        raw = synth.generate(kind, w, h, ch)
        if opaque and ch == 4:
            raw = raw.copy()
            raw[3::4] = 255
        d_raw = torch.from_numpy(raw).cuda()
        cap = (ch + 1) * w * h + 22
        d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros(raw.size, dtype=torch.uint8, device="cuda")
        ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st)
        n, ok = ctx.encode_status(st)
        ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, raw.size, st)
        path = ctx.decode_status(st)
        same = bool(torch.equal(d_out, d_raw))

        def timed(fn):
            ts = []
            for i in range(a.reps + 3):
                flush.fill_(i & 255)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            return float(np.median(ts)), float(np.min(ts))

        alg = raw.size + n
        line = f"{name:8s} {kind} {w}x{h}x{ch} E/N={n / (w * h):.3f} roundtrip={'ok' if same else 'MISMATCH'} path={path}"
        if a.only != "decode":
            med, mn = timed(lambda: ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st))
            line += f" | encode {med:8.1f} us (min {mn:.1f}) {alg / med / 1e3 / 6548.8 * 100:5.2f}% hbm"
        if a.only != "encode":
            med, mn = timed(lambda: ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, raw.size, st))
            line += f" | decode {med:8.1f} us (min {mn:.1f}) {alg / med / 1e3 / 6548.8 * 100:5.2f}% hbm"
        print(line, flush=True)


if __name__ == "__main__":
    main()
