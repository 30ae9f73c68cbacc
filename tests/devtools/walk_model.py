"""Development aid: event counts of the walk schedule on the CPU (see walk_model.c).  Usage: python tests/devtools/walk_model.py [N] [groups]"""
import ctypes as C, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import okd

here = os.path.dirname(os.path.abspath(__file__))
so = "/tmp/walk_model.so"
subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(here, "walk_model.c"), "-lm"], check=True)
lib = C.CDLL(so)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
o = okd.Oracle()
parts = o.circular_orbits(N)
nodes, idx, last = o.build_tree_canonical(parts, threads=8)
n = len(parts)
ngroups = (n + 31) // 32
stride = max(1, ngroups // G)
names = "batches popped far near mixed leaves entries lanes mixed_rounds mixed_round_nodes leaf_rounds leaf_round_nodes max_sp max_pend drains part_entries tests pack2 pack4 pack32".split() + [f"h{i}" for i in range(33)] + [f"tl{i}" for i in range(24)] + [f"sw{i}" for i in range(16)]
for defer, hard in ((0, 320), (1, 320)):
    st = (C.c_double * len(names))()
    lib.wm_run(C.c_void_p(nodes.ctypes.data), C.c_void_p(parts.ctypes.data), C.c_void_p(idx.ctypes.data), C.c_uint64(n),
               C.c_double(0.3), C.c_uint64(0), C.c_uint64(G), C.c_uint64(stride), C.c_int(defer), C.c_int(hard), st)
    g = min(G, ngroups)
    d = dict(zip(names, st))
    print(f"defer={defer} hard={hard}: " + " ".join(f"{k}={v / g:.1f}" if not k.startswith('max') else f"{k}={v:.0f}" for k, v in zip(names, st) if not k.startswith('h') and not k.startswith('tl') and not k.startswith('sw')))
    if defer: print("   hist popc:", " ".join(f"{int(d[f'h{i}']/g)}" for i in range(33)))
    print(f"   lane efficiency {d['lanes'] / d['entries'] / 32:.3f}  batch fill {d['popped'] / d['batches']:.1f}"
          + (f"  mixed round fill {d['mixed_round_nodes'] / max(1, d['mixed_rounds']):.1f} leaf round fill {d['leaf_round_nodes'] / max(1, d['leaf_rounds']):.1f}" if defer else ""))
    T = [33, 24, 28, 30, 28, 28, 32, 28]; CAP = [96, 96, 96, 96, 64, 128, 96, 192]
    for c in range(8):
        dn, sr, sd = d[f"tl{3*c}"] / g, d[f"tl{3*c+1}"] / g, d[f"tl{3*c+2}"] / g
        print(f"   two-list T={T[c]} cap={CAP[c]}: dense iterations {dn:.0f}  sparse rounds {sr:.0f} ({sd:.1f} drains)  est. cycles {dn*35.7 + sr*39.5 + (d['entries']/g - dn)*1.0:.0f} vs {d['entries']/g*35.7:.0f}")
    T = [33, 33, 33, 28, 28, 28, 24, 33]; W = [64, 96, 128, 64, 96, 128, 64, 256]
    for c in range(8):
        dn, sr = d[f"sw{2*c}"] / g, d[f"sw{2*c+1}"] / g
        print(f"   sliding window T={T[c]} W={W[c]}: dense iterations {dn:.0f}  rounds {sr:.0f}  (ideal {(d['lanes'] / g) / 32:.0f} at T=33)")
    ti = (C.c_double * 32).in_dll(lib, "g_ti")
    for c, cap in enumerate([96, 128, 192, 256]):
        r = [ti[8 * c + k] / g for k in range(8)]
        print(f"   tile drain cap={cap}: today {r[0]:.0f} | g16/U4 {r[1]:.0f}  g8/U4 {r[2]:.0f}  g8/U2 {r[3]:.0f}  g8/U1 {r[6]:.0f}  g4/U2 {r[4]:.0f}  g4/U1 {r[5]:.0f}  ({r[7]:.1f} drains)")
