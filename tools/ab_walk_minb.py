"""A/B of the walk's register budget on shard-sized grids (one GPU): walk ms per step for 24 and 32 one-warp CTAs
per SM (KDNB_WALK_MINB) at sizes whose grid equals a 1/8, 1/4, 1/2 shard of N=1M, plus N=100k and N=1M.
usage (GPU box): python tools/ab_walk_minb.py [out.txt]        — one child process per setting (the knob is read once)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import multilanguagekdtree_b200 as kd
n = int(sys.argv[1])
with kd.KDTreeSim(flags=kd.FLAG_PROFILE) as sim:
    sim.upload(kd.circular_orbits(n, seed=12345))
    sim.simple_sim(1e-3, 3)
    best = None
    for rep in range(2):
        sim.stage_reset()
        sim.simple_sim(1e-3, 10)
        ms, steps = sim.stage_ms()
        w = ms["walk"] / steps
        best = w if best is None else min(best, w)
    print("%%.4f %%.4f" %% (best, ms["build"] / steps))
""" % ROOT


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
    for n in (125_000, 250_000, 500_000, 100_000, 1_000_000):
        row = []
        for minb in ("24", "32"):
            r = subprocess.run([sys.executable, "-c", CHILD, str(n)], capture_output=True, text=True,
                               env={**os.environ, "KDNB_WALK_MINB": minb}, timeout=120)
            row.append(r.stdout.strip().split()[0] if r.returncode == 0 and r.stdout.strip() else "ERR:" + r.stderr[-200:])
        line = f"N={n:8d} grid={(n + 32) // 32:6d}  walk ms/step  minb24 {row[0]}  minb32 {row[1]}"
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()


if __name__ == "__main__":
    main()
