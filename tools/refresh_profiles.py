"""Development aid: turn the files a GPU run left in gpurun_out/ into the tracked evidence under profiles/ and the
results table of README.md.

    python tools/refresh_profiles.py

Inputs (gpurun_out/): r02_{8k_blob,8k_opaque,4k_rgb}.ncu-rep (+ .log) from tools/prof_run.py under `ncu --set full`,
r02_launches_bench_8k_rgba_photo.csv (ncu launch list of a short bench run), r02_bench_default.json, r02_bench_reference.json.
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
REPS = (("8k_rgba_photo", "r02_8k_blob"), ("8k_rgba_photo_opaque", "r02_8k_opaque"), ("4k_rgb_photo", "r02_4k_rgb"))


def raw_csv(rep):
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def summary():
    out = ["ncu --set full --clock-control none --import-source on, B200, round 2 (tools/prof_run.py: warm encode + decode calls of one workload,",
           "L2 flushed before each; the third call is captured).  Raw reports stay in gpurun_out/ (scratch); this is the summary."]
    for _, rep in REPS:
        path = os.path.join(G, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        log = open(os.path.join(G, rep + ".log")).read().strip().splitlines()
        line = next((ln for ln in log if "encode + decode calls" in ln), log[-1] if log else "")
        out += ["", f"######## {rep}  ({line})"]
        s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")], input=raw_csv(path), capture_output=True, text=True).stdout
        out.append(s.rstrip())
    open(os.path.join(P, "r02_ncu_full_summary.txt"), "w").write("\n".join(out) + "\n")


def traffic():
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    T = {}
    for wl, rep in REPS:
        path = os.path.join(G, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(io.StringIO(raw_csv(path))))
        hdr, units = rows[0], rows[1]
        ks = []
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))

            def b(name):
                return float(d[name].replace(",", "")) * mult[u[name]]

            ks.append((d["Kernel Name"].split("(")[0], b("dram__bytes_read.sum"), b("dram__bytes_write.sum")))
        T[wl] = {}
        for side in ("encode", "decode"):
            sel = [k for k in ks if side in k[0]]
            T[wl][side] = {"kernels": [k[0] for k in sel], "dram_bytes_read": int(sum(k[1] for k in sel)), "dram_bytes_write": int(sum(k[2] for k in sel))}
        T[wl]["source"] = f"ncu --set full, one launch each, gpurun_out/{rep}.ncu-rep (summary: profiles/r02_ncu_full_summary.txt)"
    json.dump(T, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)


def launches():
    src = os.path.join(G, "r02_launches_bench_8k_rgba_photo.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, P)
    rows = list(csv.reader(open(src)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    tot = collections.OrderedDict()
    for d in data:
        k = d["Kernel Name"].split("(")[0].split("<unnamed>::")[-1][:70]
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1000 if d["Metric Unit"] == "ns" else (v * 1000 if d["Metric Unit"] == "ms" else v)
        tot.setdefault(k, []).append(v)
    out = ["Launch list of `python bench.py --steps 2 --warmup 3 --also none --no-cpu-baseline --no-e2e` (8k_rgba_photo) under",
           "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"encode_|decode_\" (profiles/r02_launches_bench_8k_rgba_photo.csv).",
           "Per-launch times under ncu are cold-cache and serialised: the SHARE of a step is what compares with the bench line.", ""]
    for k, v in tot.items():
        out.append(f"{k:72s} launches {len(v):3d}  mean {sum(v) / len(v):9.1f} us")
    ours = {k: sum(v) / len(v) for k, v in tot.items()}
    s = sum(ours.values())
    out += ["", "share of one step (encode + decode of the image):"]
    for k, v in ours.items():
        out.append(f"  {k:60s} {v:9.1f} us  {v / s * 100:5.1f} %")
    open(os.path.join(P, "r02_launch_shares.txt"), "w").write("\n".join(out) + "\n")


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def readme():
    bd, br = os.path.join(G, "r02_bench_default.json"), os.path.join(G, "r02_bench_reference.json")
    if not (os.path.exists(bd) and os.path.exists(br)):
        return
    shutil.copy(bd, os.path.join(P, "r02_bench_8k_rgba_photo_default.json"))
    shutil.copy(br, os.path.join(P, "r02_bench_reference_8k_rgba_photo.json"))
    d, r = last_json(bd), last_json(br)
    a, cb = d.get("also", {}), d.get("cpu_baseline", {})
    rows = [("8k_rgba_photo (headline, alpha blobs)", d["encode_ms"], d["decode_ms"], d["value"], d["roofline"]["encode_frac"], d["roofline"]["decode_frac"],
             d["decode_path_max"], (d.get("e2e") or {}).get("value"))]
    for k, v in a.items():
        if v.get("value") is None:
            continue
        rows.append((k, v["encode_ms"], v["decode_ms"], v["value"], v["encode_frac"], v["decode_frac"], v["decode_path_max"], (v.get("e2e") or {}).get("value")))
    t = ["| workload (bench.py) | encode ms | decode ms | value GB/s (2·raw / step) | (raw+E)/t of 6548.8 GB/s: encode | decode | decode path | e2e GB/s (pinned host buffers) |",
         "|---|---|---|---|---|---|---|---|"]
    for w, e, dd, v, ef, df, p, e2 in rows:
        t.append(f"| {w} | {e:.3f} | {dd:.3f} | {v:.1f} | {ef * 100:.1f} % | {df * 100:.1f} % | {p} | {e2 if e2 is not None else ''} |")
    t.append("")
    ep = d.get("e2e_pageable") or {}
    t.append(f"`e2e_pageable` (8K image through `qoipp::encode` + `qoipp::decode` of libqoipp.so, ordinary `std::vector`): {ep.get('value')} GB/s "
             f"({ep.get('encode_ms')} + {ep.get('decode_ms')} ms).")
    st = cb.get("single_thread", {})
    t.append(f"Reference on the same box's host cores (`bench.py --impl reference`, `oracle/_ref`, {r['cpu_baseline']['cores']} threads, thread-per-image): "
             f"{r['value']} GB/s; one thread: {st.get('value')} GB/s (encode {st.get('encode_GBps')}, decode {st.get('decode_GBps')}).")
    c1 = cb.get("config1_1080p_rgba", {}).get("classes", {})
    t.append("configs[0] (1920×1080 RGBA round trip on the CPU, one thread): " + "; ".join(f"{k} {v['encode_ms']} + {v['decode_ms']} ms" for k, v in c1.items()) + ".")
    txt = "\n".join(t)
    p = os.path.join(ROOT, "README.md")
    s = open(p).read()
    s = re.sub(r"\| workload \(bench\.py\).*?configs\[0\] \(1920×1080[^\n]*\n", lambda m: txt + "\n", s, flags=re.S)
    open(p, "w").write(s)
    print(txt)


if __name__ == "__main__":
    summary()
    traffic()
    launches()
    readme()
