"""Development aid: repeated batch decodes of RGBA `photo` (alpha blobs) images, every result compared with the input."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth_torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
w = h = 512; ch = 4
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
d_raw = torch.cat([synth_torch.generate("photo", w, h, ch, seeds=[0x51F0 + k for k in range(i, min(B, i + 256))], device="cuda") for i in range(0, B, 256)]).reshape(-1)
raw_one = w * h * ch
stride = (5 * w * h + 22 + 255) // 256 * 256
d_q = torch.empty(stride * B, dtype=torch.uint8, device="cuda"); d_written = torch.zeros(B, dtype=torch.int64, device="cuda")
ctx.encode_batch_dev(d_raw, raw_one, B, w, h, ch, 0, d_q, stride, stride, d_written, st); torch.cuda.synchronize()
sizes = d_written.cpu().numpy().astype(np.uint64); offs = np.zeros(B + 1, np.uint64); offs[1:] = np.cumsum(sizes)
packed = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device="cuda")
for k in range(B): packed[int(offs[k]): int(offs[k + 1])] = d_q[k * stride: k * stride + int(sizes[k])]
d_out = torch.zeros(raw_one * B, dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
bad_runs = 0
times = []
for it in range(reps):
    d_out.zero_(); flush.fill_(it)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.decode_batch_dev(packed, offs, w, h, ch, 0, 0, d_out, raw_one, st); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    paths = ctx.decode_status_batch(B, st)
    eq = (d_out.view(B, raw_one) == d_raw.view(B, raw_one)).all(dim=1).cpu().numpy()
    if not eq.all():
        bad_runs += 1
        ids = np.nonzero(~eq)[0]
        k = int(ids[0])
        diff = (d_out.view(B, raw_one)[k] != d_raw.view(B, raw_one)[k]).nonzero().flatten()
        print(f"run {it}: {len(ids)} images differ, ids {ids[:8]}, paths {paths[ids[:8]]}; image {k}: {diff.numel()} bytes differ, first at pixel {int(diff[0]) // 4}, last {int(diff[-1]) // 4}", flush=True)
    elif it < 3:
        print(f"run {it}: ok; paths hist {np.bincount(paths)}", flush=True)
print("bad runs:", bad_runs, "of", reps, "median decode ms %.3f" % np.median(times[2:]))
