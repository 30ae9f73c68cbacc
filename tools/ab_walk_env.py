"""A/B of a walk launch knob on shard-sized grids (one GPU): walk ms per step for each value of an environment variable
(KDNB_WALK_MINB 24 32, KDNB_WALK_PF 0 1, ...) at sizes whose grid equals a 1/8, 1/4, 1/2 shard of N=1M, plus N=100k and N=1M.
usage (GPU box): python tools/ab_walk_env.py VAR value [value ...] [--out file]   — one child process per setting"""
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
    args = sys.argv[1:]
    out = None
    if "--out" in args:
        k = args.index("--out")
        out = open(args[k + 1], "w")
        del args[k:k + 2]
    var, values = args[0], args[1:]
    sizes = [int(x) for x in os.environ.get("AB_SIZES", "125000 250000 500000 100000 1000000").split()]
    for n in sizes:
        row = []
        for val in values:
            r = subprocess.run([sys.executable, "-c", CHILD, str(n)], capture_output=True, text=True,
                               env={**os.environ, var: val}, timeout=120)
            row.append(r.stdout.strip().split()[0] if r.returncode == 0 and r.stdout.strip() else "ERR:" + r.stderr[-200:])
        line = f"N={n:8d} grid={(n + 32) // 32:6d}  walk ms/step  " + "  ".join(f"{var}={v} {t}" for v, t in zip(values, row))
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()


if __name__ == "__main__":
    main()
