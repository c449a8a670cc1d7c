import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
print("N swz sbo nacc shift ksteps grid -> issue/mma total/mma (cycles)")
for grid in (1, 148):
    for N in (16, 32, 64, 128, 256):
        for swz in (128, 32):
            for (sbo, shift) in ((8 * swz, 0), (10 * swz, 1)):
                for nacc in (1, 2, 4):
                    if nacc * N > 512: continue
                    ks = swz // 32
                    i, t = zlb200.probe_umma(N, swz, sbo, nacc, 576, shift, ks, grid)
                    print(f"{N:4d} {swz:4d} {sbo:5d} {nacc:2d} {shift} {ks} {grid:4d} -> {i:7.1f} {t:7.1f}")
