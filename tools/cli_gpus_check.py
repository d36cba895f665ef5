"""wolkencli --gpus N file to file at size: N strip files of the C3 scene (POINTS in total) are written to /tmp, then
`wolkencli --gpus N` and `wolkencli --gpus 1` classify them (lossless writer, one output file); the two outputs must be
byte-identical.  Prints the CLI's own per-GPU report and phase times.
    python tools/cli_gpus_check.py N POINTS"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wolkenbase_b200 import synth  # noqa: E402

CLI = os.path.join(ROOT, "wolkenbase_b200", "host", "wolkencli")


def main():
    n_gpus, points = int(sys.argv[1]), int(sys.argv[2])
    d = synth.describe(3, points)
    names, base = [], 0
    for k in range(n_gpus):
        c0, c1 = d.grid_nx * k // n_gpus, d.grid_nx * (k + 1) // n_gpus
        c = synth.generate(3, points, seed=3, region=(c0, 0, c1 - c0, d.grid_ny), gps_base=base)
        base += c.n
        name = "/tmp/wb_strip_%d.las" % k
        c.write(name)
        names.append(name)
    print("%d points in %d files" % (base, n_gpus), flush=True)
    outs = {}
    for g in (n_gpus, 1):
        t = time.time()
        o = subprocess.run([CLI, "--gpus", str(g), "-o", "/tmp/wb_out_%d" % g, "--lossless", "--separate-classes", "0",
                            "--dump", "/tmp/wb_dump_%d" % g, "--timing"] + names, capture_output=True, text=True)
        outs[g] = o
        print("--gpus %d: rc %d, %.1f s wall" % (g, o.returncode, time.time() - t))
        print("\n".join(l for l in o.stdout.splitlines() if "GPU" in l or l.startswith("{") or "Classified" in l or l[:2] in ("1 ", "2 ")))
        if o.returncode:
            print(o.stderr[-1500:])
    same = open("/tmp/wb_out_%d.las" % n_gpus, "rb").read() == open("/tmp/wb_out_1.las", "rb").read()
    print("outputs identical:", same)
    for k in range(n_gpus):
        os.unlink(names[k])


if __name__ == "__main__":
    main()
