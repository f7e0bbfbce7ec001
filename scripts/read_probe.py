"""Sensitivity of a plain read-only stream to the bytes in flight per SM.  One bench.py
process per variant, because the variant is read from the environment (BIOEN_B200_READ_VARIANT = "U,blocks/SM")."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for var in ("8,2", "4,2", "2,2", "1,2", "4,1", "2,1"):
    env = dict(os.environ, BIOEN_B200_READ_VARIANT=var)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-dropin",
                          "--no-optimum", "--steps", "20", "--warmup", "3"], env=env, capture_output=True, text=True)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    u, b = map(int, var.split(","))
    print("U=%d blocks/SM=%d  in flight %4d KB/SM  read stream %.0f GB/s   (stream_pass %.0f GB/s)"
          % (u, b, u * b * 16, d["roofline"]["read_only_stream"]["gbs"], d["roofline"]["achieved"]), flush=True)
