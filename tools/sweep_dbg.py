import os, sys
os.environ["S360_SWEEP_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from surround360_b200 import api, synth
ctx = api.Context(0)
i0, i1 = synth.pair_512()
big0 = np.ascontiguousarray(np.tile(i0, (3, 2, 1))[:1769, :1214]); big1 = np.ascontiguousarray(np.tile(i1, (3, 2, 1))[:1769, :1214])
print("single flow 1214x1769"); ctx.computeOpticalFlow("pixflow_low", big0, big1, hint=3)
print("batch of 28"); ctx.computeOpticalFlowBatch("pixflow_low", [big0] * 28, [big1] * 28, [3] * 28)
flat0 = np.ascontiguousarray(np.tile(i0, (1, 4, 1))[:64, :2000]); flat1 = np.ascontiguousarray(np.tile(i1, (1, 4, 1))[:64, :2000])
print("wide 2000x64 (4 bands at level 0)"); ctx.computeOpticalFlow("pixflow_low", flat0, flat1, hint=3)
