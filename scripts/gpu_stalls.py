"""Where the persistent conv kernel's cycles go, per layer (instrumented instantiation) next to the CUDA-event times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200  # noqa: E402
from oracle import synth, yolov8_ref, zlw  # noqa: E402

if __name__ == "__main__":
    scale = sys.argv[1] if len(sys.argv) > 1 else "n"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    hw = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    nc = int(sys.argv[4]) if len(sys.argv) > 4 else 80
    t = yolov8_ref.synthetic_model(scale, nc, 0)
    e = zlb200.Engine(hw, hw, nc, scale, precision=zlb200.FP16, max_batch=batch)
    e.load_weights_blob(zlw.dumps(t, scale, nc))
    e.upload_resident(0, list(synth.frames_structured(batch, hw, hw)))
    prof = e.profile(0, 5)
    st = e.profile_stalls(0)
    rows = []
    print(f"{'op':26s} {'us':>7s} {'TF/s':>6s} {'GB/s':>6s} | kc ch sub ns nt stg tiles acc | {'cta_kc':>7s} {'slow_kc':>7s} {'pro_c':>6s} {'wW%':>4s} | {'Pwait%':>6s} {'Mpatch%':>7s} {'Macc%':>6s} {'Miss%':>6s} {'Ewait%':>6s} {'Ebusy%':>6s} | {'ldtm':>5s} {'stgw':>5s} {'st+f':>5s} {'back':>5s}  (cycles per tile)")
    for p, s in zip(prof, st):
        s = [int(v) for v in s]
        line = f"{p['name'][:26]:26s} {p['ms']*1e3:7.1f} {p['flops']/max(p['ms'],1e-9)/1e9:6.0f} {p['bytes']/max(p['ms'],1e-9)/1e6:6.0f}"
        if s[10]:
            n = s[10]
            life = s[6] / n
            plan = s[11]
            kc, ch, sub, ns, stg, nt, tiles = plan & 255, (plan >> 8) & 255, (plan >> 16) & 15, (plan >> 20) & 15, (plan >> 24) & 255, (plan >> 32) & 1023, (plan >> 42) & 0x3fff
            nacc, wstream = (plan >> 56) & 15, (plan >> 62) & 1
            pc = lambda v: 100.0 * v / n / life
            line += (f" | {kc:2d} {ch:2d} {sub:3d} {ns:2d} {nt:3d} {stg:3d} {tiles:5d} a{nacc}{'S' if wstream else 'R'} | {life/1e3:7.1f} {s[9]/1e3:7.1f} {s[7]/n:6.0f} {pc(s[8]):4.0f} | "
                     f"{pc(s[0]):6.1f} {pc(s[1]):7.1f} {pc(s[2]):6.1f} {pc(s[3]):6.1f} {pc(s[4]):6.1f} {pc(s[5]):6.1f}")
            tpc = max(tiles * ns / n, 1e-9)     # work units per CTA
            line += f" | {s[12]/n/tpc:5.0f} {s[13]/n/tpc:5.0f} {s[14]/n/tpc:5.0f} {s[15]/n/tpc:5.0f}  busy/tile {s[5]/n/tpc:6.0f} life/tile {life/tpc:6.0f}"
            rows.append(dict(name=p["name"], us=p["ms"] * 1e3, kc=kc, cchunks=ch, sub=sub, nsplit=ns, nt=nt, stages=stg, tiles=tiles, ctas=n,
                             cta_cycles=life, slowest=s[9], prologue=s[7] / n, producer_wait=pc(s[0]), mma_wait_patch=pc(s[1]), mma_wait_acc=pc(s[2]),
                             mma_issue=pc(s[3]), epi_wait=pc(s[4]), epi_busy=pc(s[5]), weights_wait=pc(s[8])))
        print(line)
    print("total us", sum(p["ms"] for p in prof) * 1e3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"stalls_{scale}_{batch}_{hw}.json"), "w"), indent=1)
    e.close()
