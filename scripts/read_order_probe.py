"""What the tile ORDER of the matrix passes costs at the DRAM, without TMA: k_read_tiles in the four orders of
csrc/vector_kernels.cuh next to the plain linear read stream."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"-1": "linear 16-byte stream", "0": "row pass, chunk per block", "1": "row pass, tiles round-robin",
         "2": "column pass, chunk per block", "3": "column pass, runs round-robin"}
for bps in ("2",):
    for order in ("-1", "0", "1", "2", "3"):
        env = dict(os.environ, BIOEN_B200_READ_ORDER=order, BIOEN_B200_READ_VARIANT="4," + bps)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-dropin",
                              "--no-optimum", "--steps", "10", "--warmup", "3"], env=env, capture_output=True, text=True)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("blocks/SM=%s  %-30s %.0f GB/s   (stream_pass %.0f GB/s)"
              % (bps, NAMES[order], d["roofline"]["read_only_stream"]["gbs"], d["roofline"]["achieved"]), flush=True)
