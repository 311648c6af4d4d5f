"""Development aid: compile the current csrc/ into variants/<name>.so (git-ignored, travels to the GPU box)
so that several builds can be timed back to back in ONE gpurun call (boxes differ by 10-15 %)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import build as b
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "variants"), exist_ok=True)
out = os.path.join(root, "variants", sys.argv[1] + ".so")
cmd = [b._nvcc()] + b.NVCC_FLAGS + sys.argv[2:] + ["-o", out] + [os.path.join(b.CSRC, s) for s in b.CU_SOURCES]
r = subprocess.run(cmd, capture_output=True, text=True)
print(r.stdout[-2000:], r.stderr[-2000:], out if r.returncode == 0 else "FAILED")
