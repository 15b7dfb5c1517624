"""One bench step (64 frames: ORB extract + k-NN vs predecessor + 8 local-BA windows) bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/step python scripts/one_step.py
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/one_step.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np, torch
import bench, ucoslam_b200
F, KPTS, K = 64, bench.KPTS, bench.K_NN
ctx, ctx_ba = ucoslam_b200.Context(0), ucoslam_b200.Context(0)
prm = ucoslam_b200.OrbParams(KPTS)
clip = torch.from_numpy(bench.synth_clip(F, 1234)).cuda()
kps = torch.zeros((F, KPTS, 28), dtype=torch.uint8, device="cuda")
desc = torch.zeros((F, KPTS, 32), dtype=torch.uint8, device="cuda")
nout = torch.zeros(F, dtype=torch.int32, device="cuda")
idx = torch.empty((F, KPTS, K), dtype=torch.int32, device="cuda")
dist = torch.empty_like(idx)
packed = ctx_ba.ba_pack_batch(bench.ba_windows(8, 500), bench.BA_ITERS)
torch.cuda.synchronize()

def step():
    ctx.orb_extract_batch_dev(clip.data_ptr(), F, bench.W, bench.H, bench.W, bench.W * bench.H, prm, kps.data_ptr(), desc.data_ptr(), nout.data_ptr())
    ctx.hamming_knn_batch_dev(F - 1, desc[1].data_ptr(), KPTS * 32, KPTS, nout[1:].data_ptr(), desc[0].data_ptr(), KPTS * 32, KPTS,
                              nout.data_ptr(), K, 0, idx[1].data_ptr(), dist[1].data_ptr())
    ctx.sync()
    ctx_ba.ba_solve_batch(None, bench.BA_ITERS, packed=packed)

step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
os._exit(0)
