"""Developer tool (GPU box, needs tools/_exp/libscn_TRACE.so: `tools/build_exp.sh TRACE`): clock64 timeline of CTA 0's roles in one
gather-GEMM launch - where does a tile's time go?"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SCN_B200_LIB_PATH"] = os.path.join(ROOT, "tools", "_exp", "libscn_TRACE.so")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "automatic-as-built-reconstruction_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import sparseconvnet as scn  # noqa: E402
from sparseconvnet import _lib  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
s = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
xyz = bench.building(300000)
a = xyz * (50 / 2 ** s)
a -= a.min(0)
locs = torch.from_numpy(a).long()
locs = torch.cat([locs, torch.zeros(len(locs), 1, dtype=torch.long)], 1)
ss = [4096 >> s, 4096 >> s, 512 >> s]
dev = torch.device("cuda", 0)
feats = torch.randn(len(locs), C, device=dev)
conv = scn.SubmanifoldConvolution(3, C, C, 3, False).to(dev)
scn.set_conv_precision(prec)
x = scn.InputLayer(3, ss, 4)([locs, feats])
lib = ctypes.CDLL(os.environ["SCN_B200_LIB_PATH"])
buf = np.zeros((5, 8192), dtype=np.uint64)
cnt = np.zeros(5, dtype=np.int32)
with torch.no_grad():
    for _ in range(3):
        conv(x)
    lib.scn_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p), 1)
    conv(x)
    lib.scn_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p), 1)
names = ["loader", "producer0", "mma", "epilogue6"]
ev = {}
for r in range(4):
    v = buf[r, :cnt[r]]
    ev[r] = [(int(x >> np.uint64(56)), int(x & np.uint64(0xffffffffffffff))) for x in v]
t0 = min(e[0][1] for e in ev.values() if e)
print("counts", cnt)
for r in range(4):
    print(names[r], " ".join("%d:%.2f" % (tag, (t - t0) / 1965.0) for tag, t in ev[r][:60]))
# per-tile summary for the MMA role: tags 1 (item start) 2 (tmem free) 3 (first operands ready) 4 (last commit)
m = ev[2]
tiles = []
cur = {}
for tag, t in m:
    if tag == 1 and cur:
        tiles.append(cur)
        cur = {}
    cur[tag] = t
tiles.append(cur)
d = lambda a, b: np.mean([(x[b] - x[a]) / 1965.0 for x in tiles if a in x and b in x])
print("MMA role, us per tile: start->tmem free %.2f  tmem free->first operands %.2f  first operands->last commit %.2f" %
      (d(1, 2), d(2, 3), d(3, 4)))
steps = [int(x >> np.uint64(56)) for x in buf[4, :cnt[4]]]
per = [((x[4] - x[3]) / 1965.0, n) for x, n in zip(tiles, steps) if 3 in x and 4 in x]
print("steps per tile and us/step: " + " ".join("%d:%.3f" % (n, t / n) for t, n in per))
print("mean us/step %.3f" % (sum(t for t, n in per) / sum(n for t, n in per)))
starts = [x[1] for x in tiles if 1 in x]
print("tile period (us): mean %.2f over %d tiles" % (np.mean(np.diff(starts)) / 1965.0, len(starts)))
e = ev[3]
et = []
cur = {}
for tag, t in e:
    if tag == 1 and cur:
        et.append(cur)
        cur = {}
    cur[tag] = t
et.append(cur)
de = lambda a, b: np.mean([(x[b] - x[a]) / 1965.0 for x in et if a in x and b in x])
print("epilogue, us per tile: start->accumulator ready %.2f  ready->stores issued %.2f" % (de(1, 2), de(2, 3)))
