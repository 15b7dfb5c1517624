"""Two 128-frame 640x480 extraction calls (profiling target: ncu -k regex:... --launch-skip over the first call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ucoslam_b200, orb_oracle as oo
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = ucoslam_b200.Context(0)
frames = [oo.synth_frame(i % 8) for i in range(n)]
for _ in range(2):
    ctx.orb_extract_batch(frames, ucoslam_b200.OrbParams(2000))
print("ok")
