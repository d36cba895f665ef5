"""File-to-file wall time of wolkencli (LAS in -> classified LAS 1.4 out), GPU writer and CPU writer.

Usage (GPU box): python tools/file_to_file.py [points] [scene] [writers, e.g. device,lossless]
Writes the inputs under /tmp, prints one JSON line per writer.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wolkenbase_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
scene = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cli = os.path.join(ROOT, "wolkenbase_b200", "host", "wolkencli")
work = "/tmp/wb_f2f"
os.makedirs(work, exist_ok=True)
cloud = synth.generate(scene, n, seed=scene)
las = os.path.join(work, "in.las")
cloud.write(las)
writers = sys.argv[3].split(",") if len(sys.argv) > 3 else ["device", "host", "lossless"]
for writer in writers:
    d = os.path.join(work, writer)
    os.makedirs(d, exist_ok=True)
    cmd = [cli, "-o", os.path.join(d, "out"), "--dump", os.path.join(d, "dump"), "--timing", las]
    env = dict(os.environ)
    if writer == "device-mmap":
        env["WB_WRITE_MODE"] = "mmap"
    if writer == "host":
        cmd.append("--host-writer")
    if writer == "lossless":
        cmd.append("--lossless")
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=d, env=env)
    wall = time.time() - t
    if r.returncode != 0:
        print(writer, "failed", r.stdout[-500:], r.stderr[-500:])
        continue
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    rep.update({"writer": writer, "points": cloud.n, "fmt": cloud.fmt, "process_wall_s": round(wall, 3),
                "points_per_s_file_to_file": round(cloud.n / wall),
                "out_bytes": sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".las"))})
    print(json.dumps(rep))
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
